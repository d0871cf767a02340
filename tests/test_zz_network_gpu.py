"""Network stage on the GPU: csrc/network.cu — the arg-max of the missing skeleton labels and the per-object
feature transform with scipy's tie-breaking — and the Network stage class against scipy, the oracle and fixtures produced by
executing the reference's _run_frame_backend (oracle/make_golden.py::network_frame_cases).  Same checks as the host-emulated
run in tests/test_network_cpu.py (tests/network_checks.py).  Sorts last on purpose: written after the GPU budget of round 2 was
spent, so this file's first execution on a GPU is the driver's."""
import pytest

import network_checks as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import torch
    from nellie_b200 import _cabi
    assert torch.cuda.is_available()
    return _cabi.load()


@pytest.mark.parametrize("name", K.FRAME_CASES)
def test_host_steps_match_executed_reference(lib, name):
    K.check_host_steps_on_fixture(lib, "cuda", name)


def test_relabel_matches_scipy_feature_transform(lib):
    K.check_relabel_against_scipy(lib, "cuda", trials=36)


def test_add_missing_matches_oracle(lib):
    K.check_add_missing_against_oracle(lib, "cuda")
    K.check_add_missing_tie_rule(lib, "cuda")


@pytest.mark.parametrize("name", K.FRAME_CASES)
def test_stage_class_matches_executed_reference(lib, name, tmp_path):
    from nellie_b200.networking import Network
    K.check_stage_class_on_fixture(Network, name, tmp_path if name.endswith("2d_half") else None)
