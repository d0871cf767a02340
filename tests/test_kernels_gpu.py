"""Kernel-level GPU parity through the raw C ABI: blur, Hessian components, Frobenius samples, thresholds."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, spec_from_meta

pytestmark = pytest.mark.gpu


def _vp(t):
    return C.c_void_p(t.data_ptr())


def _setup(name):
    import torch
    from nellie_b200.engine import FilterParams, FrangiEngine3D
    from oracle import pipeline as P
    g = load_golden(name)
    spec = spec_from_meta(g["meta"])
    trace = []
    P.frangi_frame(g["raw"], spec, trace=trace)
    params = FilterParams(dim_res=g["meta"]["dim_res"], no_z=False, sigmas=g["meta"].get("explicit_sigmas"))
    eng = FrangiEngine3D(g["raw"].shape, params, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return g, spec, trace, eng, st


@pytest.mark.parametrize("name", ["sample_crop", "phantom3d_iso", "phantom3d_aniso"])
def test_gaussian_cascade_is_bit_exact(name):
    import torch
    g, spec, trace, eng, st = _setup(name)
    eng.load_frame(torch.from_numpy(g["raw"].astype(np.float32)).cuda())
    # run only the blur of the first sigma through the engine's own sequence
    from nellie_b200 import _cabi
    for axis, t in enumerate(eng.steps[0]):
        if t is None or t[1] == 0:
            continue
        w, r = t
        v = eng.vol(2, 2)
        _cabi.call("nb200_gauss_axis", _vp(eng.gauss[eng.cur]), _vp(eng.gauss[1 - eng.cur]), C.byref(v), axis,
                   w.ctypes.data_as(C.POINTER(C.c_double)), r, st)
        eng.cur = 1 - eng.cur
    got = eng.gauss[eng.cur].cpu().numpy()
    assert np.array_equal(got, trace[0]["gauss"])


@pytest.mark.parametrize("name", ["sample_crop", "phantom3d_iso", "phantom3d_aniso", "phantom3d_strided", "phantom3d_pow2"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_hessian_components_and_frob_samples_are_bit_exact(name, mode):
    import torch
    from nellie_b200 import _cabi
    from oracle import pipeline as P
    g, spec, trace, eng, st = _setup(name)
    if mode == 1 and eng.div_mode == 0:
        pytest.skip("fast division not verified for these spacings")
    if mode == 2 and eng.div_mode != 2:
        pytest.skip("spacings are not powers of two")
    for rec in (trace[0], trace[-1]):
        gauss = torch.from_numpy(rec["gauss"]).cuda()
        out6 = torch.zeros((6,) + rec["gauss"].shape, dtype=torch.float32, device="cuda")
        own = eng.vol()
        _cabi.call("nb200_hessian_components", _vp(gauss), C.byref(own), eng._fd_c, mode, _vp(out6), st)
        got = out6.cpu().numpy()
        for j, nm in enumerate(["hxx", "hxy", "hxz", "hyy", "hyz", "hzz"]):
            assert np.array_equal(got[j], rec["comp"][nm]), (nm, int((got[j] != rec["comp"][nm]).sum()))
        # K2: max|H|, and sqrt(frob_sq) at the lattice points
        comp, frob_sq, max_abs, frob = P.hessian(rec["gauss"], spec)
        sz, sy, sx = eng.strides
        eng.samples.zero_()
        _cabi.call("nb200_hstats_reset", _vp(eng.hstats), st)
        _cabi.call("nb200_hessian_stats", _vp(gauss), C.byref(own), eng._fd_c, mode, None, sz, sy, sx,
                   _vp(eng.samples), _vp(eng.hstats), st)
        hs = eng.hstats.cpu().numpy()
        assert np.array([hs[0]], dtype=np.uint32).view(np.float32)[0] == np.float32(max_abs)
        assert np.array([hs[1]], dtype=np.uint32).view(np.float32)[0] == frob_sq.max()
        ref = np.sqrt(frob_sq)[::sz, ::sy, ::sx]
        got_s = eng.samples.cpu().numpy()[:ref.size].reshape(ref.shape)
        assert np.array_equal(got_s, ref), int((got_s != ref).sum())
        # the fast statistics pass (approximate march + exact re-evaluation of the near-maximum sub-chunks)
        if mode in (1, 2) and eng.fast_path:
            eng.samples.zero_()
            _cabi.call("nb200_hstats_reset", _vp(eng.hstats), st)
            _cabi.call("nb200_hessian_stats_fast", _vp(gauss), C.byref(own), eng._fd_c, mode, sz, sy, sx,
                       _vp(eng.samples), _vp(eng.hstats), _vp(eng.fast_ws), st)
            hs = eng.hstats.cpu().numpy()
            assert hs[_cabi.HS_FALLBACK] == 0
            assert np.array([hs[0]], dtype=np.uint32).view(np.float32)[0] == np.float32(max_abs)
            # the approximate maximum covers the interior of the march only (the shell is evaluated exactly)
            approx = np.array([hs[_cabi.HS_APPROX_MAX_BITS]], dtype=np.uint32).view(np.float32)[0]
            nx = rec["gauss"].shape[2]
            inner = max(float(np.abs(c[2:-2, 2:-2, 4:4 * ((nx - 2) // 4)]).max()) for c in rec["comp"].values())
            assert abs(approx - inner) <= 1e-4 * inner, (approx, inner)
            assert np.array([hs[3]], dtype=np.uint32).view(np.float32)[0] == np.abs(rec["gauss"]).max()
            got_s = eng.samples.cpu().numpy()[:ref.size].reshape(ref.shape)
            assert np.array_equal(got_s, ref), int((got_s != ref).sum())


def test_divisor_modes():
    import torch
    from nellie_b200 import _cabi
    lib = _cabi.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    m = C.c_int(-1)
    for d, want in [(0.25, 2), (0.5, 2), (1.0, 2), (0.1, 1), (0.2, 1), (0.4, 1), (0.0655, 1), (0.131, 1)]:
        _cabi.check(lib.nb200_divisor_mode(C.c_float(np.float32(d)), C.byref(m), st), "nb200_divisor_mode")
        assert m.value == want, (d, m.value)


# scipy.ndimage.gaussian_filter is the third-party primitive the reference calls (filtering.py:828-835);
# the vectorised Z march and the fused Y+X tile kernel must reproduce it bit for bit at every radius,
# for shapes that are not multiples of the tile and for lines shorter than the radius (repeated reflection).
@pytest.mark.parametrize("shape", [(37, 70, 132), (9, 33, 248), (5, 3, 12), (40, 130, 131)])
@pytest.mark.parametrize("sigma", [0.3, 0.6, 0.98, 1.3, 1.7, 2.0, 2.3, 2.7])
def test_gauss_z_vec_and_fused_yx_match_scipy(shape, sigma):
    import scipy.ndimage as ndi
    import torch
    from nellie_b200 import _cabi
    from nellie_b200._cabi import Vol
    from nellie_b200.engine import gaussian_taps
    rng = np.random.default_rng(int(sigma * 100) + shape[2])
    x = (rng.random(shape, dtype=np.float32) * 500.0).astype(np.float32)
    w, r = gaussian_taps(sigma, 3.0)
    assert 1 <= r <= 8
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    dp = C.POINTER(C.c_double)
    v = Vol.whole(*shape)
    a = torch.from_numpy(x).cuda()
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    _cabi.call("nb200_gauss_axis", _vp(a), _vp(b), C.byref(v), 0, w.ctypes.data_as(dp), r, st)
    ref_z = ndi.gaussian_filter1d(x, sigma, axis=0, truncate=3.0, mode="reflect")
    assert np.array_equal(b.cpu().numpy(), ref_z)
    _cabi.call("nb200_gauss_yx", _vp(b), _vp(c), C.byref(v), w.ctypes.data_as(dp), w.ctypes.data_as(dp), r, st)
    ref = ndi.gaussian_filter(x, sigma, truncate=3.0, mode="reflect")
    assert np.array_equal(c.cpu().numpy(), ref)
    # and the per-axis kernels agree with the fused one
    d = torch.empty_like(a)
    _cabi.call("nb200_gauss_axis", _vp(b), _vp(d), C.byref(v), 1, w.ctypes.data_as(dp), r, st)
    _cabi.call("nb200_gauss_axis", _vp(d), _vp(b), C.byref(v), 2, w.ctypes.data_as(dp), r, st)
    assert torch.equal(b, c)


# _mask_volume's opening (filtering.py:964-966) on the marching bit-plane kernel (nx % 4 == 0) and on the
# brick fallback (any nx), against scipy.ndimage.binary_opening itself; also as a Z window of a taller buffer.
@pytest.mark.parametrize("shape", [(70, 75, 256), (5, 9, 12), (33, 40, 131), (140, 33, 124)])
@pytest.mark.parametrize("density", [0.5, 0.85])
def test_opening_matches_scipy(shape, density):
    import scipy.ndimage as ndi
    import torch
    from nellie_b200 import _cabi
    from nellie_b200._cabi import Vol
    rng = np.random.default_rng(shape[2])
    vol = rng.random(shape, dtype=np.float32)
    vol = ndi.uniform_filter(vol, 3).astype(np.float32)          # blobs, so that the opening keeps something
    cut = float(np.quantile(vol, 1.0 - density))
    vol[rng.random(shape) < 0.05] = -1.0                          # dead voxels of the accumulator
    v_pos = np.maximum(vol, 0.0)
    ref = v_pos * ndi.binary_opening(v_pos > np.float32(cut))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    acc = torch.from_numpy(vol).cuda()
    out = torch.full_like(acc, 7.0)
    thr = torch.tensor([np.float32(cut), 1.0], dtype=torch.float64, device="cuda")
    v = Vol.whole(*shape)
    _cabi.call("nb200_finalize_opening", _vp(acc), _vp(out), C.byref(v), _vp(thr), st)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert ref.any()
    # pass-through when there is no positive sample
    thr0 = torch.tensor([0.0, 0.0], dtype=torch.float64, device="cuda")
    _cabi.call("nb200_finalize_opening", _vp(acc), _vp(out), C.byref(v), _vp(thr0), st)
    assert np.array_equal(out.cpu().numpy(), v_pos)
    # slab window: planes [3, nz-2) of the same buffer must equal the same planes of the whole-frame result
    if shape[0] > 8:
        out2 = torch.full_like(acc, 7.0)
        w = Vol(shape[0], shape[1], shape[2], 3, shape[0] - 2, 0, shape[0])
        _cabi.call("nb200_finalize_opening", _vp(acc), _vp(out2), C.byref(w), _vp(thr), st)
        got = out2.cpu().numpy()
        assert np.array_equal(got[3:shape[0] - 2], ref[3:shape[0] - 2])
        assert (got[:3] == 7.0).all() and (got[shape[0] - 2:] == 7.0).all()


def test_remove_edges_kernel_matches_the_reference_loop():
    """nb200_remove_edges against the oracle's restatement of filtering.py:969-1000, 3-D (per slice) and 2-D,
    with empty slices, boxes lower than the margin, dead (-1) voxels and responses touching the frame border."""
    import torch
    from nellie_b200 import _cabi

    def remove_edge_bands_(t):
        d = t.cuda().contiguous()
        shp = d.shape if d.dim() == 3 else (1,) + tuple(d.shape)
        _cabi.call("nb200_remove_edges", C.c_void_p(d.data_ptr()), int(shp[0]), int(shp[1]), int(shp[2]), 15,
                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return d.cpu()

    from oracle import pipeline as P
    rng = np.random.default_rng(3)
    for shape in [(6, 50, 20), (4, 17, 9), (1, 40, 8), (3, 12, 5)]:
        v = np.zeros(shape, np.float32)
        for z in range(shape[0]):
            if z == 1:
                continue                                         # an empty slice
            r0 = int(rng.integers(0, shape[1] - 2))
            r1 = int(rng.integers(r0, shape[1]))
            v[z, r0:r1 + 1] = rng.random((r1 + 1 - r0, shape[2])) * (rng.random((r1 + 1 - r0, shape[2])) < 0.4)
            v[z, r0, 0] = 1.0
            v[z, r1, -1] = 1.0
        spec3 = P.FrameSpec(dim_res={"X": 1.0, "Y": 1.0, "Z": 1.0, "T": 1.0}, no_z=False)
        want = P.remove_edges(v.copy(), spec3)
        acc = v.copy()
        acc[(v == 0) & (rng.random(shape) < 0.5)] = -1.0        # the engines' "dead voxel" marker
        got = remove_edge_bands_(torch.from_numpy(acc)).numpy()
        assert np.array_equal(np.maximum(got, 0.0), want), shape
        spec2 = P.FrameSpec(dim_res={"X": 1.0, "Y": 1.0, "T": 1.0}, no_z=True)
        want2 = P.remove_edges(v[0].copy(), spec2)
        got2 = remove_edge_bands_(torch.from_numpy(v[0].copy())).numpy()
        assert np.array_equal(got2, want2), shape
    assert not remove_edge_bands_(torch.zeros((3, 8, 8))).any()


def test_filter_with_remove_edges_matches_the_oracle():
    """Filter(remove_edges=True) (filtering.py:931-932) end to end, 3-D and 2-D, against the oracle."""
    from types import SimpleNamespace
    from nellie_b200 import Filter
    from nellie_b200.phantoms import tubular_phantom_np
    from oracle import pipeline as P
    for shape, no_z in [((20, 64, 72), False), ((96, 104), True)]:
        dim_res = {"X": 0.1, "Y": 0.1, "Z": None if no_z else 0.1, "T": 1.0}
        raw = tubular_phantom_np(shape if not no_z else (1,) + shape, seed=41, n_tubes=5)
        raw = raw if not no_z else raw[0]
        info = SimpleNamespace(no_t=True, no_z=no_z, shape=(1,) + raw.shape, axes="TYX" if no_z else "TZYX", dim_res=dim_res)
        f = Filter(info, device="b200", remove_edges=True)
        f._get_t()
        f._set_default_sigmas()
        got = f.filter_frame_host(raw)
        ref = P.filter_frame(raw, P.FrameSpec(dim_res=dim_res, no_z=no_z, remove_edges=True))
        assert np.array_equal(got > 0, ref > 0), shape
        if no_z:
            assert (np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-6 * np.abs(ref).max()).all()
        else:
            assert np.array_equal(got, ref), shape

