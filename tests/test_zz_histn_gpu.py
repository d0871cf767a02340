"""Label thresholds with histogram_nbins != 256 on the GPU (csrc/histn.cu): against numpy / the oracle, and the Label class
against fixtures of the executed reference.  Sorts last: written after the GPU budget of round 2 was spent."""
import ctypes as C

import pytest

import histn_checks as K

pytestmark = pytest.mark.gpu


def test_thresholds_match_numpy():
    import torch
    from nellie_b200 import _cabi
    K.check_against_numpy(_cabi.load(), "cuda", C.c_void_p(torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("name", K.NBINS_CASES)
def test_label_class_with_other_bin_counts_matches_executed_reference(name):
    K.check_label_class_on_fixture(name)
