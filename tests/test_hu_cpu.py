"""HuMomentTracking feature extraction without a GPU: the oracle against fixtures produced by executing the
unmodified reference in both of its ROI modes (oracle/make_golden.py::hu_cases), the reference's own test of _log_hu, and the
CUDA kernels of csrc/hu.cu compiled for the host through oracle/cuda_emu.h, driven by the product's HuFeatureEngine.
The shared checks live in tests/hu_checks.py; tests/test_zz_hu_gpu.py runs them on the GPU."""
import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import pytest

import hu_checks as K
from conftest import ROOT


@pytest.mark.parametrize("name", K.HU_CASES)
def test_oracle_hu_features_match_executed_reference(name):
    from oracle import pipeline as P
    g = K.load_hu_case(name)
    for tag, dense in (("dense", True), ("stream", False)):
        coords, phys, stats, hu = P.hu_frame_features(g["raw"], g["frangi"], g["distance"], g["marker"],
                                                      K.scaling_of(g["meta"]), g["meta"]["no_z"], dense=dense)
        assert np.array_equal(coords, g[f"coords_{tag}"]) and np.array_equal(phys, g[f"phys_{tag}"])
        assert stats.dtype == np.float32 and np.array_equal(stats, g[f"stats_{tag}"])
        assert hu.dtype == g[f"hu_{tag}"].dtype and np.array_equal(hu, g[f"hu_{tag}"])


def test_oracle_replays_reference_log_hu_test():
    """tests/test_hu_tracking.py:16-24 of the reference."""
    from oracle import pipeline as P
    log_hu = P.hu_log(np.array([[0.0, 1e-12, -1e-6]], dtype=np.float32))
    assert np.all(np.isfinite(log_hu)) and np.isclose(log_hu[0, 0], 0.0)


@pytest.fixture(scope="module")
def emu():
    from nellie_b200 import _cabi
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "hu_host.so")
    srcs = [os.path.join(ROOT, "oracle", "hu_host.cpp"), os.path.join(ROOT, "oracle", "cuda_emu.h"),
            os.path.join(ROOT, "nellie_b200", "csrc", "hu.cu"), os.path.join(ROOT, "nellie_b200", "csrc", "devmath.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                        f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++", srcs[0],
                        "-o", so], check=True)
    lib = C.CDLL(so)
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith("nb200_hu_"):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return K.Backend(lib, "cpu")


@pytest.mark.parametrize("name", K.HU_CASES)
def test_emulated_hu_kernels_match_executed_reference(emu, name):
    K.check_fixture(emu, name)


@pytest.mark.parametrize("shape", [(7, 20, 33), (1, 30, 31), (40, 50)])
def test_emulated_frame_transforms_match_oracle(emu, shape):
    K.check_frame_transforms(emu, shape)


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("shape", [(14, 30, 33), (60, 70)])
def test_emulated_stats_and_bounds_match_oracle(emu, shape, dtype):
    K.check_stats_and_bounds(emu, shape, dtype)


def test_markers_then_hu_features_on_files_emulated(emu, tmp_path):
    """Markers.run() -> files -> HuMomentFeatures, both stages on host-emulated kernels."""
    import torch
    import markers_checks as MK
    from nellie_b200 import _cabi, hu_tracking as H, mocap_marking as M
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                    f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++",
                    os.path.join(ROOT, "oracle", "markers_host.cpp"), "-o",
                    os.path.join(ROOT, "oracle", "_build", "markers_host.so")], check=True)
    mlib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "markers_host.so"))
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith("nb200_markers_") or name in ("nb200_gauss_axis", "nb200_gauss_yx"):
            fn = getattr(mlib, name)
            fn.argtypes, fn.restype = argtypes, restype

    class EmuMarkers(M.Markers):
        def _torch_device(self):
            return torch.device("cpu")

        def _engine_for(self, frame_shape):
            if not self.sigmas:
                self._set_default_sigmas()
            if self._engine is None:
                self._engine = M.MarkerEngine(tuple(frame_shape), self.im_info.no_z, tuple(float(s) for s in self.sigmas),
                                              self.z_ratio, self.max_radius_px, self.peak_min_distance, "cpu", lib=mlib)
            return self._engine

    class EmuHu(H.HuMomentFeatures):
        def _torch_device(self):
            return torch.device("cpu")

        def _engine_for(self, frame_shape):
            if self._engine is None:
                self._engine = H.HuFeatureEngine(tuple(frame_shape), self.im_info.no_z, "cpu", lib=emu.lib)
            return self._engine

    K.check_markers_then_hu_on_files(EmuMarkers, EmuHu, tmp_path)


def test_mirror_class_on_emulated_kernels(emu):
    """HuMomentFeatures._get_frame_features(t): the reference's _FrameFeatures contract, empty frames included."""
    import torch
    from nellie_b200 import hu_tracking as H

    class Emu(H.HuMomentFeatures):
        def _torch_device(self):
            return torch.device("cpu")

        def _engine_for(self, frame_shape):
            key = tuple(int(s) for s in frame_shape)
            if self._engine is None or self._engine.shape != key:
                self._engine = H.HuFeatureEngine(key, self.im_info.no_z, "cpu", lib=emu.lib)
            return self._engine

    g = K.load_hu_case("hu_phantom3d_iso")
    info = SimpleNamespace(no_t=False, no_z=False, shape=(2,) + g["raw"].shape, axes="TZYX", dim_res=g["meta"]["dim_res"])
    m = Emu(info, num_t=2, dense_limit=int(5e7))
    m.im_memmap = np.stack([g["raw"], g["raw"]])
    m.im_frangi_memmap = np.stack([g["frangi"], g["frangi"]])
    m.im_distance_memmap = np.stack([g["distance"], g["distance"]])
    m.im_marker_memmap = np.stack([g["marker"], np.zeros_like(g["marker"])])
    ff = m._get_frame_features(0)
    assert np.array_equal(ff.coords_voxel, g["coords_dense"]) and np.array_equal(ff.coords_phys, g["phys_dense"])
    assert np.array_equal(ff.stats, g["stats_dense"]) and ff.hu.shape == g["hu_dense"].shape
    empty = m._get_frame_features(1)
    assert empty.coords_voxel.shape == (0, 3) and empty.stats.shape == (0, 0) and empty.hu.shape == (0, 0)
    low = Emu(info, num_t=2, low_memory=True)
    low.im_memmap, low.im_frangi_memmap = m.im_memmap, m.im_frangi_memmap
    low.im_distance_memmap, low.im_marker_memmap = m.im_distance_memmap, m.im_marker_memmap
    ff = low._get_frame_features(0)
    assert np.array_equal(ff.stats, g["stats_stream"]) and ff.hu.dtype == np.float32
    # a uint16 raw frame travels in its own type and takes numpy's integer rules on the device
    gu = K.load_hu_case("hu_sample_crop")
    assert gu["raw"].dtype == np.uint16
    iu = SimpleNamespace(no_t=False, no_z=False, shape=(2,) + gu["raw"].shape, axes="TZYX", dim_res=gu["meta"]["dim_res"])
    mu = Emu(iu, num_t=2, dense_limit=int(5e7))
    mu.im_memmap, mu.im_frangi_memmap = gu["raw"][None], gu["frangi"][None]
    mu.im_distance_memmap, mu.im_marker_memmap = gu["distance"][None], gu["marker"][None]
    fu = mu._get_frame_features(0)
    assert np.array_equal(fu.stats, gu["stats_dense"]) and np.allclose(fu.hu, gu["hu_dense"], rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        H.HuMomentFeatures(info, device="cpu")
    with pytest.raises(NotImplementedError):
        H.integer_bits(np.int32)
