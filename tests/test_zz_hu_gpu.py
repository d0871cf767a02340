"""HuMomentTracking feature extraction on the GPU: the checks of tests/hu_checks.py through the CUDA library
(executed-reference fixtures in both ROI modes, frame transforms, boxes, statistics per dtype) and the mirror class on
files.  Sorts last on purpose: added after the last full GPU session of round 2; verified through the host emulation
(tests/test_hu_cpu.py) and by scripts/markers_quick.py on a B200."""
import numpy as np
import pytest

import hu_checks as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    from nellie_b200 import _cabi
    assert torch.cuda.is_available()
    return K.Backend(_cabi.load(), "cuda")


@pytest.mark.parametrize("name", K.HU_CASES)
def test_hu_features_match_executed_reference(cuda, name):
    K.check_fixture(cuda, name)


@pytest.mark.parametrize("shape", [(7, 20, 33), (1, 30, 31), (40, 50)])
def test_frame_transforms_match_oracle(cuda, shape):
    K.check_frame_transforms(cuda, shape)


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("shape", [(14, 30, 33), (60, 70)])
def test_stats_and_bounds_match_oracle(cuda, shape, dtype):
    K.check_stats_and_bounds(cuda, shape, dtype)


def test_markers_then_hu_features_on_files(cuda, tmp_path):
    from nellie_b200 import Markers
    from nellie_b200.hu_tracking import HuMomentFeatures
    K.check_markers_then_hu_on_files(Markers, HuMomentFeatures, tmp_path)
