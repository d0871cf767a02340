"""Network stage without a GPU: the oracle's restatement of the host steps against the executed reference
(oracle/make_golden.py::network_frame_cases), and csrc/network.cu compiled for the host through oracle/cuda_emu.h — the
per-object feature transform with scipy's tie-breaking, the arg-max of the missing skeleton labels — against scipy, the
oracle and the fixtures, driven by the product's NetworkEngine / Network host code.  Shared checks: tests/network_checks.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import network_checks as K
from conftest import ROOT


@pytest.mark.parametrize("name", K.FRAME_CASES)
def test_oracle_network_frame_matches_executed_reference(name):
    from oracle import pipeline as P
    g = K.load_frame_case(name)
    no_z = g["meta"]["no_z"]
    assert np.array_equal(P.network_add_missing(g["cleaned"], g["labels"], g["frangi"]), g["added"])
    assert np.array_equal(P.network_relabel_objects(g["branch"], g["labels"], tuple(g["scaling"])), g["relabelled"])
    branch, pixel_class, relabelled = P.network_frame(g["labels"], g["frangi"], g["skeleton"], tuple(g["scaling"]), no_z)
    assert np.array_equal(branch, g["branch"]) and np.array_equal(pixel_class, g["pixel_class"])
    assert relabelled.dtype == np.uint32 and np.array_equal(relabelled, g["relabelled"])


@pytest.fixture(scope="module")
def emu_lib():
    from nellie_b200 import _cabi
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "network_host.so")
    srcs = [os.path.join(ROOT, "oracle", "network_host.cpp"), os.path.join(ROOT, "oracle", "cuda_emu.h"),
            os.path.join(ROOT, "nellie_b200", "csrc", "network.cu")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                        f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++", srcs[0],
                        "-o", so], check=True)
    lib = C.CDLL(so)
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith("nb200_network_"):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return lib


@pytest.mark.parametrize("name", K.FRAME_CASES)
def test_emulated_host_steps_match_executed_reference(emu_lib, name):
    K.check_host_steps_on_fixture(emu_lib, "cpu", name)


def test_emulated_relabel_matches_scipy_feature_transform(emu_lib):
    K.check_relabel_against_scipy(emu_lib, "cpu")


def test_emulated_add_missing_matches_oracle(emu_lib):
    K.check_add_missing_against_oracle(emu_lib, "cpu")
    K.check_add_missing_tie_rule(emu_lib, "cpu")


def _emu_network(emu_lib):
    """nellie_b200.Network on the CPU: network.cu host-emulated; the three label.cu kernels (shared-memory CCL, not
    emulatable; GPU-tested in tests/test_network_gpu.py) answered by the oracle."""
    import torch
    from nellie_b200 import networking as N
    from oracle import pipeline as P

    class Emu(N.Network):
        @property
        def device(self):
            return torch.device("cpu")

        def _engine(self):
            if self._net is None:
                self._net = N.NetworkEngine(self.im_info.no_z, self.scaling, "cpu", lib=emu_lib)
            return self._net

        def _get_pixel_class(self, skel):
            return torch.from_numpy(P.network_pixel_class(skel.numpy(), self.im_info.no_z).astype(np.uint8))

        def _get_branch_skel_labels(self, pixel_class):
            return torch.from_numpy(P.network_branch_labels(pixel_class.numpy(), self.im_info.no_z).astype(np.int32))

        def _remove_connected_label_pixels(self, skel):
            return torch.from_numpy(P.network_remove_connected(skel.numpy(), self.im_info.no_z).astype(np.int32))

    return Emu


@pytest.mark.parametrize("name", ["network_frame_cfg3_half", "network_frame_phantom2d_half", "network_frame_sample_crop"])
def test_stage_class_on_emulated_kernels(emu_lib, name, tmp_path):
    K.check_stage_class_on_fixture(_emu_network(emu_lib), name, tmp_path if name == "network_frame_phantom2d_half" else None)


def test_network_refuses_cpu_and_needs_a_skeletonizer():
    from types import SimpleNamespace
    from nellie_b200.networking import Network
    info = SimpleNamespace(no_t=True, no_z=True, shape=(1, 9, 9), axes="TYX", dim_res={"X": 0.2, "Y": 0.2, "Z": None})
    with pytest.raises(ValueError):
        Network(info, device="cpu")
    net = Network(info)
    try:
        import skimage  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="scikit-image"):
            net._skeletonize(np.ones((9, 9), np.int32))
