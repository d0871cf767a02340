"""The fast Hessian path (csrc/hessian_fast.cu: approximate classification with proven margins + exact candidates in
one barrier-free TMA march) must be bit-identical to the exact kernels of round 1 (frangi.cu / sparse.cu), which are
themselves pinned to the executed reference by tests/test_filter_gpu.py and tests/test_kernels_gpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ISO = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
ANISO = {"X": 0.0655, "Y": 0.0655, "Z": 0.25, "T": 1.0}
POW2 = {"X": 0.125, "Y": 0.125, "Z": 0.125, "T": 1.0}
POW2_ANISO = {"X": 0.125, "Y": 0.125, "Z": 0.25, "T": 1.0}


def _engines(shape, dim_res, sigmas=None, **kw):
    import torch
    from nellie_b200.engine import FilterParams, FrangiEngine3D
    fast = FrangiEngine3D(shape, FilterParams(dim_res=dim_res, sigmas=sigmas, **kw), device="cuda")
    exact = FrangiEngine3D(shape, FilterParams(dim_res=dim_res, sigmas=sigmas, **kw), device="cuda")
    assert fast.fast_path, "shape / spacing should qualify for the fast path"
    exact.fast_path = False
    fast.diag = torch.zeros(8, dtype=torch.int64, device="cuda")
    return fast, exact


def _compare(fast, exact, frame, expect_fallback=False):
    import torch
    from nellie_b200 import _cabi
    of = fast.filter_frame(frame).clone()
    acc_f = fast.acc.clone()
    oe = exact.filter_frame(frame).clone()
    rf, re_ = fast.sigma_records(), exact.sigma_records()
    for col in (_cabi.SP_GAMMA, _cabi.SP_GAMMA_SQ, _cabi.SP_FROB_THR, _cabi.SP_FROB_CUT, _cabi.SP_MAX_ABS, _cabi.SP_SKIP,
                _cabi.SP_FROBSQ_MIN):
        assert np.array_equal(rf[:, col], re_[:, col], equal_nan=True), (col, rf[:, col], re_[:, col])
    own = slice(fast.pad_lo, fast.pad_lo + fast.nz_own)
    assert torch.equal(acc_f[own], exact.acc[own]), int((acc_f[own] != exact.acc[own]).sum())
    assert torch.equal(of, oe)
    if expect_fallback:
        assert (rf[:, _cabi.SP_UNSAFE] == 1).all()
    else:
        assert (rf[:, _cabi.SP_UNSAFE] == 0).all(), rf[:, _cabi.SP_UNSAFE]
    return fast.diag.cpu().numpy(), rf


@pytest.mark.parametrize("shape,dim_res,sigmas", [
    ((24, 48, 64), ISO, None),
    ((40, 75, 136), ANISO, None),                    # anisotropic: approximate tests run in Hessian units
    ((33, 61, 140), POW2, [1.0, 1.2, 1.4, 1.6]),     # division mode POW2, two X tiles, ragged rows
    ((21, 30, 260), POW2_ANISO, [1.0, 1.5]),         # three X tiles, partial last tile
    ((70, 200, 256), ISO, [1.0, 1.4, 2.2, 3.0]),     # several Z chunks per tile, lattice strides > 1
    ((9, 19, 12), ISO, [1.0]),                       # barely larger than the stencil
])
def test_fast_path_is_bit_identical_to_the_exact_kernels(shape, dim_res, sigmas):
    import torch
    from nellie_b200.phantoms import tubular_phantom_np
    raw = tubular_phantom_np(shape, seed=sum(shape), n_tubes=max(3, int(np.prod(shape)) // 40000))
    fast, exact = _engines(shape, dim_res, sigmas)
    diag, _ = _compare(fast, exact, torch.from_numpy(raw).cuda())
    print(f"{shape}: candidates={diag[0]} uncertain kills={diag[1]} survivors={diag[2]} of {raw.size * len(fast.sigmas)}")
    if raw.size > 20000:
        assert 0 < diag[2] <= diag[0] < raw.size * len(fast.sigmas) // 2


def test_fast_path_integer_image_and_flat_regions():
    """uint8-like data: many exactly-equal neighbours (zero Hessians, frob_sq exactly at thresholds) and a frame that is
    constant over half of the volume."""
    import torch
    from nellie_b200.phantoms import tubular_phantom_np
    shape = (36, 90, 128)
    raw = np.clip(np.round(tubular_phantom_np(shape, seed=5, n_tubes=12) / 16.0), 0, 255).astype(np.float32)
    raw[:, :, 64:] = 7.0
    fast, exact = _engines(shape, ISO, [1.0, 1.8])
    _compare(fast, exact, torch.from_numpy(raw).cuda())


def test_fast_path_mask_disabled_and_fixed_threshold():
    import torch
    from nellie_b200.phantoms import tubular_phantom_np
    shape = (28, 56, 72)
    t = torch.from_numpy(tubular_phantom_np(shape, seed=9, n_tubes=6)).cuda()
    fast, exact = _engines(shape, ISO, [1.0, 2.0], mask=False)
    diag, _ = _compare(fast, exact, t)
    assert diag[1] == 0                                # nothing dies without the Frobenius gate
    # fixed thresholds: one well inside the range, one that empties the mask (skipped sigma), one in the band where
    # the bounds 1 <= max frob <= 3 cannot decide (exact max frob^2 through the gated pass)
    for thr in (0.2, 7.0, 2.2, 3.9):
        fast, exact = _engines(shape, ISO, [1.0, 2.0], frob_thresh=thr)
        _compare(fast, exact, t)


def test_fast_path_falls_back_when_its_error_bound_is_useless():
    """A large offset makes the proven error bound (proportional to max|g|) comparable with the Hessian itself: the
    statistics pass must notice (hstats[FALLBACK] -> sp[UNSAFE]) and the exact kernels must take over."""
    import torch
    from nellie_b200.phantoms import tubular_phantom_np
    shape = (24, 48, 64)
    raw = tubular_phantom_np(shape, seed=3, n_tubes=4) * 1e-3 + 4.0e3
    fast, exact = _engines(shape, ISO, [1.0, 1.6])
    _compare(fast, exact, torch.from_numpy(raw.astype(np.float32)).cuda(), expect_fallback=True)


def test_fast_path_slab_windows():
    """Z-slab windows (nb200_vol): the fast kernels on three slabs with halo planes must reproduce the un-sharded
    acc planes bit for bit (statistics are reduced by hand like the NCCL all-reduce does)."""
    import ctypes as C
    import torch
    from nellie_b200 import _cabi
    from nellie_b200.engine import FilterParams, FrangiEngine3D
    from nellie_b200.phantoms import tubular_phantom_np
    shape = (48, 60, 136)
    raw = torch.from_numpy(tubular_phantom_np(shape, seed=12, n_tubes=10)).cuda()
    p = FilterParams(dim_res=ISO, sigmas=[1.0, 1.6])
    whole = FrangiEngine3D(shape, p, device="cuda")
    ref = whole.filter_frame(raw).clone()
    bounds = [(0, 17), (17, 30), (30, 48)]
    slabs = [FrangiEngine3D(shape, p, device="cuda", slab=(48, a, b - a)) for a, b in bounds]
    assert all(s.fast_path for s in slabs)

    def exchange(buf_of, depth):
        for k, s in enumerate(slabs):
            a, b = bounds[k]
            lo, hi = min(depth, s.pad_lo), min(depth, s.pad_hi)
            for j, t in enumerate(slabs):
                if j == k:
                    continue
                ta, tb = bounds[j]
                for z in list(range(a - lo, a)) + list(range(b, b + hi)):
                    if ta <= z < tb:
                        buf_of(s)[z - s.zg_off].copy_(buf_of(t)[z - t.zg_off])

    # The slabs advance in lock step through the per-sigma sequence of FrangiEngine3D._analyse_sigma_fast, with plain
    # copies / torch reductions standing in for the NCCL halo exchange and all-reduces of sharding.ZComm.
    for s in slabs:
        s.load_frame(raw[s.z0:s.z0 + s.nz_own])
        s.acc.zero_()
    src = [s.cur for s in slabs]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    for i in range(len(whole.sigmas)):
        # blur: exchange halos of the source volume, then every slab blurs owned+2 planes
        rz = whole.steps[i][0][1] if whole.steps[i][0] is not None else 0
        exchange(lambda e: e.gauss[src[slabs.index(e)]], rz + 2)
        for k, s in enumerate(slabs):
            s.exchange_halo = lambda buf, depth: None
            src[k] = s._blur_sigma(i, src[k], {})
        # analysis with explicit reductions: run each slab up to a reduction, reduce, continue
        sz, sy, sx = whole.strides
        for k, s in enumerate(slabs):
            g = s.gauss[src[k]]
            own = s.vol()
            s._call("nb200_lattice_sample", vp(g), C.byref(own), sz, sy, sx, vp(s.samples), st)
            s._call("nb200_hist_reset", vp(s.hist), st)
            s._call("nb200_hist_minmax", vp(s.samples), s.n_samples, _cabi.TF_NONE, None, vp(s.hist), st)
        _reduce_minmax([s.hist for s in slabs])
        for s in slabs:
            s._call("nb200_hist_bins", vp(s.samples), s.n_samples, _cabi.TF_NONE, None, vp(s.hist), st)
        _reduce_sum([s.hist for s in slabs])
        for k, s in enumerate(slabs):
            g = s.gauss[src[k]]
            own = s.vol()
            s._call("nb200_finalize_gamma", vp(s.hist), vp(s.sp[i]), st)
            s._call("nb200_hstats_reset", vp(s.hstats), st)
            s._call("nb200_hessian_stats_fast", vp(g), C.byref(own), s._fd_c, s.div_mode, sz, sy, sx, vp(s.samples),
                    vp(s.hstats), vp(s.fast_ws), st)
        _reduce_max([s.hstats for s in slabs])
        for s in slabs:
            s._call("nb200_finalize_max_abs", vp(s.hstats), vp(s.sp[i]), st)
            s._call("nb200_hist_reset", vp(s.hist), st)
            dp = C.c_void_p(s.sp[i].data_ptr() + 8 * _cabi.SP_MAX_ABS)
            s._call("nb200_hist_minmax", vp(s.samples), s.n_samples, _cabi.TF_DIV, dp, vp(s.hist), st)
        _reduce_minmax([s.hist for s in slabs])
        for s in slabs:
            dp = C.c_void_p(s.sp[i].data_ptr() + 8 * _cabi.SP_MAX_ABS)
            s._call("nb200_hist_bins", vp(s.samples), s.n_samples, _cabi.TF_DIV, dp, vp(s.hist), st)
        _reduce_sum([s.hist for s in slabs])
        for k, s in enumerate(slabs):
            g = s.gauss[src[k]]
            own = s.vol()
            s._call("nb200_finalize_frob_fast", vp(s.hist), vp(s.hstats), float("nan"), 2.0, s.max_scale, 1, vp(s.sp[i]), st)
            s._call("nb200_frangi_fast", vp(g), vp(s.acc), C.byref(own), s._fd_c, s.div_mode, 0.5, 0.5, vp(s.sp[i]), None, st)
        assert all((s.sp[i].cpu().numpy()[[_cabi.SP_UNSAFE, _cabi.SP_AMBIG]] == 0).all() for s in slabs)
        assert all(np.array_equal(s.sp[i].cpu().numpy()[:11], whole.sp[i].cpu().numpy()[:11]) for s in slabs)
    for k, s in enumerate(slabs):
        a, b = bounds[k]
        assert torch.equal(s.acc[s.pad_lo:s.pad_lo + s.nz_own], whole.acc[a:b]), k
    assert ref is not None


def _reduce_minmax(states):
    import torch
    lo = torch.stack([s[0] for s in states]).min()
    hi = torch.stack([s[1] for s in states]).max()
    for s in states:
        s[0] = lo
        s[1] = hi


def _reduce_sum(states):
    import torch
    tot = torch.stack([s[2:] for s in states]).sum(0)
    for s in states:
        s[2:] = tot


def _reduce_max(states):
    import torch
    m = torch.stack(list(states)).max(0).values
    for s in states:
        s.copy_(m)
