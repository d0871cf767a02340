"""Network stage: the three array kernels of its GPU backend against fixtures produced by executing the
reference's own methods (oracle/make_golden.py::network_cases) and against the oracle on fresh random skeletons."""
import json
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["network3d", "network2d"])
def test_network_kernels_match_reference(name):
    from nellie_b200.networking import NetworkKernels
    z = np.load(f"{GOLDEN_DIR}/{name}.npz")
    no_z = json.loads(str(z["meta"]))["no_z"]
    net = NetworkKernels(SimpleNamespace(no_z=no_z))
    assert np.array_equal(net._remove_connected_label_pixels(z["skel"]), z["cleaned"])
    pc = net._get_pixel_class(z["skel"])
    assert pc.dtype == np.uint8 and np.array_equal(pc, z["pixel_class"])
    br = net._get_branch_skel_labels(pc)
    assert br.dtype == np.int32 and np.array_equal(br, z["branch"])


@pytest.mark.parametrize("shape", [(19, 45, 67), (1, 50, 33), (40, 70)])
def test_network_kernels_match_oracle_on_random_skeletons(shape):
    import torch
    from nellie_b200.networking import NetworkKernels
    from oracle import pipeline as P
    no_z = len(shape) == 2
    rng = np.random.default_rng(sum(shape))
    skel = (rng.random(shape) < 0.06).astype(np.int32) * rng.integers(1, 9, shape).astype(np.int32)
    net = NetworkKernels(SimpleNamespace(no_z=no_z))
    assert np.array_equal(net._remove_connected_label_pixels(skel), P.network_remove_connected(skel, no_z))
    pc = net._get_pixel_class(skel)
    assert np.array_equal(pc, P.network_pixel_class(skel, no_z))
    assert np.array_equal(net._get_branch_skel_labels(pc), P.network_branch_labels(pc, no_z))
    # device tensors in -> device tensors out
    d = net._get_pixel_class(torch.from_numpy(skel).cuda())
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), pc)
