"""Markers stage on the GPU: the checks of tests/markers_checks.py through the CUDA library — executed-reference
fixtures, scipy / oracle comparisons of every kernel, the reference's own tests replayed on the mirror class, run() on
files — plus one frame of production size against the oracle.  (The file sorts last on purpose: these kernels were added
after the last GPU session of round 2 and were verified through the host emulation only, tests/test_markers_cpu.py.)"""
import numpy as np
import pytest

import markers_checks as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    from nellie_b200 import _cabi
    assert torch.cuda.is_available()
    return K.Backend(_cabi.load(), "cuda")


@pytest.mark.parametrize("sigma", [2.2, 2.45, 2.87, 3.1, 3.4, 4.4])
@pytest.mark.parametrize("shape", [(21, 45, 68), (7, 30, 131)])
def test_gauss_axis_large_radii_and_derivative_taps_match_scipy(cuda, shape, sigma):
    """The Markers scales reach radii 9..18 (truncate 4.0), beyond what the Filter's cascade uses, and filter with
    second-derivative taps: every axis of nb200_gauss_axis against scipy.ndimage.gaussian_filter1d, orders 0 and 2, incl.
    lines shorter than the radius."""
    import ctypes as C
    import scipy.ndimage as ndi
    import torch
    from nellie_b200 import _cabi
    from nellie_b200.engine import gaussian_taps
    from nellie_b200.engine2d import gaussian_taps_order2
    rng = np.random.default_rng(int(sigma * 100) + shape[2])
    x = (rng.random(shape, dtype=np.float32) * 20.0).astype(np.float32)
    a = torch.from_numpy(x).cuda()
    b = torch.empty_like(a)
    v = _cabi.Vol.whole(*shape)
    for order, (w, r) in ((0, gaussian_taps(sigma, 4.0)), (2, gaussian_taps_order2(sigma, 4.0))):
        assert r == int(4.0 * sigma + 0.5) and 9 <= r <= 18
        for axis in range(3):
            _cabi.call("nb200_gauss_axis", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.byref(v), axis,
                       w.ctypes.data_as(C.POINTER(C.c_double)), r, cuda.stream())
            ref = ndi.gaussian_filter1d(x, sigma, axis=axis, order=order, truncate=4.0, mode="reflect")
            assert np.array_equal(b.cpu().numpy(), ref), (order, axis, r)


@pytest.mark.parametrize("name", K.MARKER_CASES)
def test_markers_match_executed_reference(cuda, name):
    K.check_fixture(cuda, name)


@pytest.mark.parametrize("shape,clamp", K.EDT_CASES)
def test_edt_and_border_match_scipy(cuda, shape, clamp):
    K.check_edt_and_border(cuda, shape, clamp)


@pytest.mark.parametrize("shape,z_res", K.PEAK_CASES)
def test_peaks_and_nms_match_oracle(cuda, shape, z_res):
    K.check_peaks_and_nms(cuda, shape, z_res)


@pytest.mark.parametrize("case", K.OPTION_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c["kw"].items()) or "aniso")
def test_option_matrix_matches_oracle(cuda, case):
    K.check_option_matrix(cuda, case)


def test_mirror_class_replays_reference_marker_tests(cuda):
    from nellie_b200.mocap_marking import Markers
    K.check_mirror_class_replays_reference_tests(Markers)


def test_mirror_class_helpers_and_run_on_files(cuda, tmp_path):
    from nellie_b200.mocap_marking import Markers
    K.check_mirror_class_helpers_and_run_on_files(Markers, tmp_path)


def test_z_sharded_markers_equal_the_whole_frame(cuda, tmp_path):
    from nellie_b200.mocap_marking import Markers
    K.check_z_sharded_equals_whole_frame(Markers, tmp_path)


def test_markers_frame_of_production_size_matches_oracle(cuda):
    """A 96 x 320 x 384 frame (1.2e7 voxels: every grid-stride loop wraps many times) of labelled tubes and blobs."""
    from types import SimpleNamespace
    from nellie_b200.mocap_marking import Markers
    from nellie_b200.phantoms import tubular_phantom_np
    from oracle import pipeline as P
    shape = (96, 320, 384)
    raw = tubular_phantom_np(shape, seed=41, n_tubes=60)
    rng = np.random.default_rng(41)
    labels = (raw > np.percentile(raw, 93)).astype(np.int32)
    labels[K.blob_mask(shape, rng, n_blobs=12, r_max=28.0)] = 2
    dim_res = {"X": 0.1, "Y": 0.1, "Z": 0.15, "T": 1.0}
    info = SimpleNamespace(no_t=True, no_z=False, shape=(1,) + shape, axes="TZYX", dim_res=dim_res)
    m = Markers(info, num_t=1)
    m.im_memmap, m.label_memmap = raw[None], labels[None]
    m.shape = m.label_memmap.shape
    m._set_default_sigmas()
    marker, distance, border = m._run_frame_impl(0)
    ref = P.marker_frame(raw, labels, P.MarkerSpec(dim_res=dim_res))
    assert np.array_equal(distance, ref[1])
    assert np.array_equal(border, ref[2])
    assert np.array_equal(marker, ref[0]) and marker.sum() > 100
