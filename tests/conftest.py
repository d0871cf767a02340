import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def spec_from_meta(meta):
    from oracle.pipeline import FrameSpec
    fk = meta.get("filter_kwargs") or {}
    lk = meta.get("label_kwargs") or {}
    return FrameSpec(dim_res=meta["dim_res"], no_z=meta["no_z"], sigmas=meta.get("explicit_sigmas"),
                     run_mask=bool(meta.get("run_mask", True)), **fk, **lk)


@pytest.fixture(scope="session")
def golden():
    return load_golden
