"""Markers stage without a GPU.

(1) the oracle restatement (oracle/pipeline.py marker_*) against fixtures produced by executing the unmodified reference
    (oracle/make_golden.py::marker_cases) and against the reference's own two tests (tests/test_mocap_marking.py);
(2) the CUDA kernels of csrc/markers.cu compiled for the host through oracle/cuda_emu.h (same kernel bodies, index
    arithmetic, grid-stride loops and C entry points, run serially) against scipy, the oracle and the fixtures, driven
    by the product's own ``MarkerEngine`` / ``Markers`` host code with the emulated library injected.
The checks themselves live in tests/markers_checks.py; tests/test_zmarkers_gpu.py runs the same ones on the GPU.
"""
import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import pytest

import markers_checks as K
from conftest import ROOT


# ---- (1) oracle ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", K.MARKER_CASES)
def test_oracle_markers_match_executed_reference(name):
    from oracle import pipeline as P
    g = K.load_marker_case(name)
    spec = K.marker_spec(g["meta"])
    assert np.array_equal(np.asarray(P.marker_sigmas(spec)), g["sigmas"])
    marker, distance, border = P.marker_frame(g["raw"], g["labels"], spec, frangi=g["frangi"])
    assert marker.dtype == np.uint8 and distance.dtype == np.float32 and border.dtype == np.uint8
    assert np.array_equal(marker, g["marker"])
    assert np.array_equal(distance, g["distance"])
    assert np.array_equal(border, g["border"])


def test_oracle_replays_reference_marker_tests():
    from oracle import pipeline as P
    intensity, labels, dim_res = K.reference_test_inputs()
    spec = P.MarkerSpec(dim_res=dim_res, no_z=True, num_sigma=3)
    marker, distance, border = P.marker_frame(intensity, labels, spec)
    assert marker.sum() == 1 and marker[4, 4] == 1          # the one bright voxel at the centre of the square
    assert distance.max() == 3.0
    # tests/test_mocap_marking.py:61-72: the border never overlaps the mask
    mask = np.zeros((7, 7), dtype=bool)
    mask[2:5, 2:5] = True
    _, b = P.marker_distance(mask, spec)
    assert b.shape == mask.shape and not np.any(b & mask)


# ---- (2) the CUDA kernels, host-emulated ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    """oracle/markers_host.cpp: csrc/markers.cu compiled by g++ behind the same C entry points."""
    from nellie_b200 import _cabi
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "markers_host.so")
    srcs = [os.path.join(ROOT, "oracle", "markers_host.cpp"), os.path.join(ROOT, "oracle", "cuda_emu.h"),
            os.path.join(ROOT, "nellie_b200", "csrc", "markers.cu")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                        f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++", srcs[0],
                        "-o", so], check=True)
    lib = C.CDLL(so)
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith("nb200_markers_") or name in ("nb200_gauss_axis", "nb200_gauss_yx"):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return K.Backend(lib, "cpu")


def _emu_markers(be):
    """nellie_b200.Markers with the emulated library injected (tests only): exercises the mirror class's host code."""
    import torch
    from nellie_b200 import mocap_marking as M

    class Emu(M.Markers):
        def _torch_device(self):
            return torch.device("cpu")

        def _engine_for(self, frame_shape):
            key = tuple(int(s) for s in frame_shape)
            if not self.sigmas:
                self._set_default_sigmas()
            sig = tuple(float(s) for s in self.sigmas)
            if self._engine is None or self._engine.shape != key or tuple(self._engine.sigmas) != sig:
                self._engine = M.MarkerEngine(key, self.im_info.no_z, sig, self.z_ratio, self.max_radius_px,
                                              self.peak_min_distance, "cpu", lib=be.lib, truncate=self.truncate)
            return self._engine

    return Emu


@pytest.mark.parametrize("name", K.MARKER_CASES)
def test_emulated_kernels_match_executed_reference(emu, name):
    K.check_fixture(emu, name)


@pytest.mark.parametrize("shape,clamp", K.EDT_CASES)
def test_emulated_edt_and_border_match_scipy(emu, shape, clamp):
    K.check_edt_and_border(emu, shape, clamp)


@pytest.mark.parametrize("shape,z_res", K.PEAK_CASES)
def test_emulated_peaks_and_nms_match_oracle(emu, shape, z_res):
    K.check_peaks_and_nms(emu, shape, z_res)


@pytest.mark.parametrize("case", K.OPTION_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c["kw"].items()) or "aniso")
def test_emulated_option_matrix_matches_oracle(emu, case):
    K.check_option_matrix(emu, case)


def test_mirror_class_replays_reference_marker_tests(emu):
    K.check_mirror_class_replays_reference_tests(_emu_markers(emu))


def test_mirror_class_helpers_and_run_on_files(emu, tmp_path):
    K.check_mirror_class_helpers_and_run_on_files(_emu_markers(emu), tmp_path)


def test_z_sharded_markers_equal_the_whole_frame(emu, tmp_path):
    K.check_z_sharded_equals_whole_frame(_emu_markers(emu), tmp_path)


def test_markers_refuses_cpu_and_unknown_images():
    from nellie_b200.mocap_marking import Markers
    info = SimpleNamespace(no_t=True, no_z=True, shape=(1, 9, 9), axes="TYX", dim_res={"X": 0.2, "Y": 0.2})
    with pytest.raises(ValueError):
        Markers(info, device="cpu")
    with pytest.raises(ValueError):
        Markers(info, prefer_gpu=False)
    with pytest.raises(ValueError):
        Markers(info, use_im="raw")
