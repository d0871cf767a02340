"""Label thresholds with histogram_nbins != 256 without a GPU: the oracle against the executed reference
(oracle/make_golden.py::label_nbins_cases) and csrc/histn.cu, compiled for the host through oracle/cuda_emu.h, against
numpy's histogram / the oracle's triangle and Otsu.  The GPU run (tests/test_zz_histn_gpu.py) adds the Label class itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import histn_checks as K
from conftest import ROOT


@pytest.mark.parametrize("name", K.NBINS_CASES)
def test_oracle_label_with_other_bin_counts_matches_executed_reference(name):
    from oracle import pipeline as P
    g = K.load_nbins_case(name)
    spec = P.FrameSpec(dim_res=g["dim_res"], no_z=g["no_z"], **g["kw"])
    it, ft = P.label_thresholds(g["raw"], g["frangi"], spec)
    assert (it is None and np.isnan(g["intensity_thresh"])) or float(it) == g["intensity_thresh"]
    assert float(ft) == g["frangi_thresh"]
    labels = P.label_frame(g["frangi"], spec, ft, raw=g["raw"] if it is not None else None, intensity_thresh=it)
    assert np.array_equal(labels, g["labels"])


@pytest.fixture(scope="module")
def emu_lib():
    from nellie_b200 import _cabi
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "histn_host.so")
    srcs = [os.path.join(ROOT, "oracle", "histn_host.cpp"), os.path.join(ROOT, "oracle", "cuda_emu.h"),
            os.path.join(ROOT, "nellie_b200", "csrc", "histn.cu"), os.path.join(ROOT, "nellie_b200", "csrc", "devmath.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                        f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++", srcs[0],
                        "-o", so], check=True)
    lib = C.CDLL(so)
    for name in ("nb200_histn_workspace_bytes", "nb200_histn_threshold", "nb200_hist_reset", "nb200_hist_minmax"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = _cabi._SIGS[name]
    return lib


def test_emulated_thresholds_match_numpy(emu_lib):
    K.check_against_numpy(emu_lib, "cpu")


@pytest.mark.parametrize("name", K.NBINS_CASES)
def test_emulated_thresholds_match_executed_reference(emu_lib, name):
    """The sampled values of the fixture (oracle's restatement of _sample_nonzero) through the emulated kernels."""
    from oracle import pipeline as P
    g = K.load_nbins_case(name)
    spec = P.FrameSpec(dim_res=g["dim_res"], no_z=g["no_z"], **g["kw"])
    nbins = spec.histogram_nbins
    it = None
    if spec.otsu_thresh_intensity:
        vals = P.label_sample(g["raw"], spec)
        integer = g["raw"].dtype.kind in "iu"
        out = K.thresholds_of(emu_lib, "cpu", vals.astype(np.float32), nbins, 0, int(integer), 1)
        it = np.float64(out[0]) if integer else np.float32(out[0])
        assert float(it) == g["intensity_thresh"]
    vals = P.label_sample(g["frangi"], spec, g["raw"] if it is not None else None, it)
    out = K.thresholds_of(emu_lib, "cpu", vals, nbins, 1, 0, 0)
    ft = float(min(10 ** np.float32(out[5]), 10 ** np.float32(out[6])))      # labelling.py:452-455 on the two scalars
    assert ft == g["frangi_thresh"]
