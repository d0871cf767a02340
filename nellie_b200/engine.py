"""Device-side orchestration of one frame of the Filter hot path (single GPU or one Z slab).

Everything here is plumbing: torch owns the device buffers and the stream, every arithmetic
step is a kernel of ``libnellie_b200.so`` called through the C ABI, and no scalar comes back to
the host inside a frame (gamma, thresholds and max|H| stay in the device record ``sp``).

Reference control flow restated: nellie/segmentation/filtering.py:806-853 (_compute_vesselness),
:910-933 (_run_frame), :952-967 (_mask_volume), :1005-1031 (_run_filter).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import Vol


# ---------------------------------------------------------------------------------------------
# host-side scalar logic (F0): sigma schedule, spacing, sampling strides, Gaussian taps
# ---------------------------------------------------------------------------------------------
@dataclass
class FilterParams:
    """Constructor knobs of the reference Filter (filtering.py:23-40) plus the physical pixel sizes."""
    dim_res: dict
    no_z: bool = False
    min_radius_um: float = 0.25
    max_radius_um: float = 1.0
    alpha_sq: float = 0.5
    beta_sq: float = 0.5
    frob_thresh: Optional[float] = None
    frob_thresh_division: float = 2
    max_threshold_samples: int = 1_000_000
    truncate: float = 3.0
    remove_edges: bool = False
    sigmas: Optional[Sequence[float]] = None
    mask: bool = True          # Filter._run_frame(mask=False): no Frobenius gate (filtering.py:910-933)

    def z_ratio(self) -> float:  # filtering.py:75-78
        z_res = self.dim_res.get("Z") or self.dim_res.get("X") or 1.0
        x_res = self.dim_res.get("X") or 1.0
        return float(z_res) / float(x_res)

    def spacing(self):  # filtering.py:265-275
        y = float(self.dim_res.get("Y") or 1.0)
        x = float(self.dim_res.get("X") or 1.0)
        if self.no_z:
            return (y, x)
        z = float(self.dim_res.get("Z") or self.dim_res.get("X") or 1.0)
        return (z, y, x)

    def sigma_list(self):  # filtering.py:288-316
        if self.sigmas is not None:
            return sorted(float(s) for s in self.sigmas)
        lo_px = self.min_radius_um / self.dim_res["X"]
        hi_px = self.max_radius_um / self.dim_res["X"]
        a, b = lo_px / 2.0, hi_px / 3.0
        s_min, s_max = min(a, b), max(a, b)
        if s_max <= s_min:
            s_max = s_min + 0.2
        step = max(0.2, (s_max - s_min) / 5.0)
        vals = list(np.arange(s_min, s_max, step, dtype=float))
        vals.sort()
        return [float(v) for v in vals]

    def sigma_vec(self, sigma):  # filtering.py:277-286
        if self.no_z:
            return (float(sigma), float(sigma))
        return (float(sigma) / self.z_ratio(), float(sigma), float(sigma))

    def delta_sigma_vec(self, prev, cur):  # filtering.py:816-825
        return tuple(float(np.sqrt(max(0.0, float(c) ** 2 - float(p) ** 2)))
                     for p, c in zip(self.sigma_vec(prev), self.sigma_vec(cur)))

    def fd_spacing_f32(self):
        """[fl32(h), fl32(2h)] per axis: numpy.gradient divides a float32 array by the Python floats
        ``h`` (edges) and ``2. * h`` (interior), both cast to float32."""
        out = []
        for h in self.spacing():
            out += [np.float32(h), np.float32(2.0 * h)]
        return np.asarray(out, dtype=np.float32)


def sample_strides(shape, max_samples):
    """filtering.py:328-340 (_sample_strides)."""
    nd = len(shape)
    if max_samples is None or max_samples <= 0:
        return (1,) * nd
    total = int(np.prod(shape))
    if total <= max_samples:
        return (1,) * nd
    s0 = max(1, int(np.ceil((total / max_samples) ** (1.0 / nd))))
    st = [s0] * nd
    while int(np.prod([int(np.ceil(n / s)) for n, s in zip(shape, st)])) > max_samples:
        k = int(np.argmax([n / s for n, s in zip(shape, st)]))
        st[k] += 1
    return tuple(st)


def gaussian_taps(delta_sigma: float, truncate: float):
    """scipy.ndimage._gaussian_kernel1d(order=0): radius int(truncate*sd+0.5); returns (w[0..r], r)."""
    sd = float(delta_sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64), radius


_DIV_MODE_CACHE = {}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------
# 3-D engine
# ---------------------------------------------------------------------------------------------
class FrangiEngine3D:
    """Buffers + kernel sequence for frames of one shape ``(nz, ny, nx)`` on one GPU.

    ``slab`` describes a Z slab of a larger frame for multi-GPU runs (see sharding.py); by default the
    engine owns the whole frame.  Buffers: two blurred volumes (ping-pong), the accumulator and the
    output — 16 B/voxel resident, 8 GiB... 16 GiB for a 1024^3 frame, well inside 180 GB of HBM3e.
    """

    def __init__(self, shape, params: FilterParams, device=None, slab=None):
        if params.no_z or len(shape) != 3:
            raise ValueError("FrangiEngine3D needs a (Z, Y, X) frame")
        self.p = params
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = _cabi.load()
        self.sigmas = params.sigma_list()
        self.steps = []
        prev = 0.0
        for s in self.sigmas:
            dvec = params.delta_sigma_vec(prev, s)
            taps = [gaussian_taps(d, params.truncate) if d > 1e-15 else None for d in dvec]
            self.steps.append(taps)
            prev = s
        self.halo_z = 2 + max([t[0][1] for t in self.steps if t[0] is not None] + [0])
        # slab geometry --------------------------------------------------------------------
        if slab is None:
            nz, ny, nx = (int(s) for s in shape)
            self.nz_glob, self.z0, self.nz_own = nz, 0, nz
        else:
            self.nz_glob, self.z0, self.nz_own = slab
            ny, nx = int(shape[-2]), int(shape[-1])
        self.ny, self.nx = ny, nx
        self.pad_lo = min(self.halo_z, self.z0)
        self.pad_hi = min(self.halo_z, self.nz_glob - (self.z0 + self.nz_own))
        self.nz_buf = self.pad_lo + self.nz_own + self.pad_hi
        self.zg_off = self.z0 - self.pad_lo
        if min(self.nz_glob, ny, nx) < 2:
            raise ValueError("numpy.gradient needs at least 2 samples along every axis")
        self.strides = sample_strides((self.nz_glob, ny, nx), params.max_threshold_samples)
        sz, sy, sx = self.strides
        g0, g1 = self.z0, self.z0 + self.nz_own
        first = ((g0 + sz - 1) // sz) * sz
        self.n_lat_z = (g1 - 1 - first) // sz + 1 if first < g1 else 0
        self.n_samples = self.n_lat_z * math.ceil(ny / sy) * math.ceil(nx / sx)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.gauss = [torch.empty((self.nz_buf, ny, nx), **f32) for _ in range(3)]   # sigma i, scratch, sigma i+1
        self.code = torch.empty((self.nz_buf, ny, nx), **f32)                        # K2's per-voxel record
        self.side_stream = torch.cuda.Stream(dev)
        # blur of sigma i+1 on a side stream under K2/K3 of sigma i.  Measured: no gain on B200 — K2 holds 61 K of
        # the 64 K registers of an SM (2 CTAs x 256 threads x 120), so the block scheduler cannot co-schedule the
        # blur CTAs and the two streams serialise (84.2 vs 84.4 ms/step); kept as an option, off by default
        self.overlap_blur = False
        self.acc = torch.empty((self.nz_buf, ny, nx), **f32)
        self.out = torch.empty((self.nz_own, ny, nx), **f32)
        self.samples = torch.empty(max(1, self.n_samples), **f32)
        self.frob_samples = torch.empty(max(1, self.n_samples), **f32)   # fast path: frob samples next to the gauss samples
        # histogram state and Hessian stats live in ONE record: a Z-sharded run moves both with one all-gather
        self.state = torch.zeros(_cabi.STATE_WORDS, dtype=torch.int64, device=dev)
        self.hist = self.state[:_cabi.HIST_WORDS]
        self.hstats = self.state[_cabi.HIST_WORDS:]
        self.sp = torch.zeros((len(self.sigmas), _cabi.SP_WORDS), dtype=torch.float64, device=dev)
        self.select = torch.zeros(_cabi.SELECT_WORDS, dtype=torch.int64, device=dev)
        self.pct = torch.zeros(2, dtype=torch.float64, device=dev)
        self.fd = params.fd_spacing_f32()
        self._fd_c = self.fd.ctypes.data_as(C.POINTER(C.c_float))
        self.div_mode = self._pick_div_mode()
        self.sparse_list = True  # sparse K3 as stream + solve kernels over a global candidate list (else: one kernel, smem queues)
        self.list_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sparse_k3 = True  # K3 from K2's per-voxel record (nb200_frangi_sparse); False = dense march (nb200_frangi_accumulate)
        # Fast Hessian path (hessian_fast.cu): approximate classification with proven margins + exact candidates in
        # one barrier-free TMA march.  Needs nx % 4 == 0, a verified division mode and < 2^32 voxels per buffer;
        # otherwise (and whenever the device flags sp[UNSAFE]) the exact kernels above run.
        self.fast_path = (nx % 4 == 0 and self.div_mode in (_cabi.DIV_FAST, _cabi.DIV_POW2)
                          and self.nz_buf * ny * nx < 2 ** 32 and ny >= 5 and nx >= 12)
        self.fast_ws = torch.zeros(int(self.lib.nb200_hessian_fast_workspace_bytes()) // 4, dtype=torch.int32, device=dev)
        self.max_scale = float(1.0 / (np.float64(self.fd[1::2].min()) ** 2))
        self.use_graph = False   # replay the per-frame sequence as a CUDA graph (see filter_frame); ZShardedFilter turns it on
        self.batch_sigmas = False  # run every reduction point once for ALL sigmas (_run_sigmas_batched); ZShardedFilter turns it on
        self._batch = None
        self._graphs, self._eager_done = {}, {}
        self.diag = None      # set to a zeroed int64[8] device tensor to collect candidate / survivor counts (tests)
        self.fuse_yx = True   # Y and X blur passes in one kernel (nb200_gauss_yx); False = one kernel per axis
        self.launches = 0     # C-ABI calls
        self.kernels = 0      # CUDA kernels enqueued by those calls
        self.profile = None   # set to a list to record (name, start, end) CUDA events per C-ABI call
        # hooks for the multi-GPU driver (identity on one GPU)
        self.exchange_halo = lambda buf, depth: None
        self.reduce_hist_minmax = lambda state: None
        self.reduce_hist_bins = lambda state: None
        self.reduce_hstats = lambda hs: None
        self.fold_state = lambda state, stage: None      # fast path: one packed reduction per reduction point
        self.gather_samples = lambda s, n: (s, n)

    def _pick_div_mode(self):
        """Weakest division mode over the six grid-spacing divisors (verified on the device, once)."""
        modes = []
        with torch.cuda.device(self.device):
            for d in self.fd:
                key = float(d)
                if key not in _DIV_MODE_CACHE:
                    m = C.c_int(0)
                    _cabi.check(self.lib.nb200_divisor_mode(C.c_float(key), C.byref(m), _stream()), "nb200_divisor_mode")
                    _DIV_MODE_CACHE[key] = int(m.value)
                modes.append(_DIV_MODE_CACHE[key])
        if all(m == _cabi.DIV_POW2 for m in modes):
            return _cabi.DIV_POW2
        if all(m in (_cabi.DIV_POW2, _cabi.DIV_FAST) for m in modes):
            return _cabi.DIV_FAST
        return _cabi.DIV_IEEE

    # -- geometry helpers ---------------------------------------------------------------------
    def vol(self, extra_lo=0, extra_hi=0) -> Vol:
        """Window over the owned planes, optionally widened into the halo (clipped at the frame)."""
        zc0 = self.pad_lo - min(extra_lo, self.pad_lo)
        zc1 = self.pad_lo + self.nz_own + min(extra_hi, self.pad_hi)
        return Vol(self.nz_buf, self.ny, self.nx, zc0, zc1, self.zg_off, self.nz_glob)

    # CUDA kernels behind one C-ABI call (default 1): K2 = march + border shell, redo = reset + IEEE twin + shell,
    # sparse K3 = stream + solve, dense K3 = march + IEEE twin + shell, percentile = radix select passes
    KERNELS_PER_CALL = {"nb200_hessian_stats_code": 2, "nb200_hessian_stats_redo": 3, "nb200_frangi_sparse": 2,
                        "nb200_frangi_accumulate": 3, "nb200_percentile": 17, "nb200_hessian_stats_fast": 4,
                        "nb200_hessian_stats_ambig": 2, "nb200_frangi_fast": 2, "nb200_frangi_sparse_gated": 2}

    def _call(self, name, *args):
        self.launches += 1
        self.kernels += self.KERNELS_PER_CALL.get(name, 1)
        if self.profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(getattr(self.lib, name)(*args), name)
            e1.record()
            self.profile.append((name, e0, e1))
            return
        _cabi.check(getattr(self.lib, name)(*args), name)

    def profile_summary(self):
        """{call name: (launches, total ms)} from the CUDA events recorded while ``self.profile`` was a list."""
        torch.cuda.synchronize(self.device)
        out = {}
        for name, e0, e1 in self.profile or []:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1))
        return out

    # -- thresholds from a sample buffer --------------------------------------------------------
    def _histogram(self, samples, n, transform, divisor_ptr):
        st = _stream()
        self._call("nb200_hist_reset", _ptr(self.hist), st)
        self._call("nb200_hist_minmax", _ptr(samples), n, transform, divisor_ptr, _ptr(self.hist), st)
        self.reduce_hist_minmax(self.hist)
        self._call("nb200_hist_bins", _ptr(samples), n, transform, divisor_ptr, _ptr(self.hist), st)
        self.reduce_hist_bins(self.hist)

    # -- one frame ----------------------------------------------------------------------------
    def load_frame(self, frame: torch.Tensor):
        """Copy the owned planes of ``frame`` (any real dtype, device tensor) into the blur buffer as
        float32 (filtering.py:924: xp.asarray(frame, dtype=float32)); the caller's tensor is never
        written (the reference aliases and overwrites float32 inputs — SURVEY App. C-1)."""
        own = self.gauss[0][self.pad_lo:self.pad_lo + self.nz_own]
        own.copy_(frame)
        self.cur = 0

    def _blur_sigma(self, i, src_idx, last_reader):
        """F1 for sigma i on the CURRENT stream: incremental blur of ``gauss[src_idx]`` (axes Z, Y, X in scipy's
        order) into the other volumes; returns the index of the volume holding the result.  ``gauss[src_idx]`` itself
        is never written (the previous sigma's K2/K3 may still be reading it on another stream); a destination
        volume is only written after the event of its last reader."""
        taps = self.steps[i]
        if not any(t is not None and t[1] > 0 for t in taps):
            return src_idx
        st = _stream()
        cur_stream = torch.cuda.current_stream(self.device)
        rz = taps[0][1] if taps[0] is not None else 0
        self.exchange_halo(self.gauss[src_idx], rz + 2)
        # Z pass over owned+2 planes (reads the exchanged halo); Y/X passes likewise so the Hessian stencil finds
        # blurred neighbours without a second exchange
        v = self.vol(2, 2)
        dp = C.POINTER(C.c_double)
        fuse_yx = (self.fuse_yx and taps[1] is not None and taps[2] is not None
                   and taps[1][1] == taps[2][1] and 1 <= taps[1][1] <= 8)
        free = [k for k in range(len(self.gauss)) if k != src_idx]
        cur, n_pass = src_idx, 0
        for axis, t in enumerate(taps):
            if t is None or t[1] == 0:
                continue
            w, r = t
            dst = free[n_pass % 2] if cur == src_idx or cur != free[n_pass % 2] else free[(n_pass + 1) % 2]
            ev = last_reader.pop(dst, None)
            if ev is not None:
                cur_stream.wait_event(ev)
            if axis == 1 and fuse_yx:
                self._call("nb200_gauss_yx", _ptr(self.gauss[cur]), _ptr(self.gauss[dst]), C.byref(v), w.ctypes.data_as(dp),
                           taps[2][0].ctypes.data_as(dp), r, st)
                cur = dst
                break
            self._call("nb200_gauss_axis", _ptr(self.gauss[cur]), _ptr(self.gauss[dst]), C.byref(v), axis,
                       w.ctypes.data_as(dp), r, st)
            cur = dst
            n_pass += 1
        return cur

    def _analyse_sigma(self, i, g):
        """F2-F9 for sigma i on the CURRENT stream: gamma, Hessian statistics, Frobenius threshold, K3."""
        st = _stream()
        sp_i = self.sp[i]
        sz, sy, sx = self.strides
        own = self.vol()
        fixed = float("nan") if self.p.frob_thresh is None else float(self.p.frob_thresh)
        division = float(self.p.frob_thresh_division or 0.0)
        mask_on = 1 if self.p.mask else 0
        if self.fast_path:
            self._analyse_sigma_fast(i, g, own, sp_i, fixed, division, mask_on, st)
            return
        # F2/F3: gamma from the positive lattice sample of the blurred volume
        self._call("nb200_lattice_sample", _ptr(g), C.byref(own), sz, sy, sx, _ptr(self.samples), st)
        self._histogram(self.samples, self.n_samples, _cabi.TF_NONE, None)
        self._call("nb200_finalize_gamma", _ptr(self.hist), _ptr(sp_i), st)
        self._call("nb200_hstats_reset", _ptr(self.hstats), st)
        # F4: Hessian statistics (max|H|, max frob^2, frob samples) + K2's per-voxel record for the sparse K3
        code = self.code if self.sparse_k3 else None
        self._call("nb200_hessian_stats_code", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                   sz, sy, sx, _ptr(self.samples), _ptr(self.hstats), _ptr(code), st)
        self.reduce_hstats(self.hstats)
        self._call("nb200_finalize_max_abs", _ptr(self.hstats), _ptr(sp_i), st)
        if self.div_mode == _cabi.DIV_FAST:
            # safety net of the fast division: kernels that return at once unless sp[UNSAFE] was just set
            self._call("nb200_hessian_stats_redo", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                       sz, sy, sx, _ptr(self.samples), _ptr(self.hstats), _ptr(code), st)
            self.reduce_hstats(self.hstats)
            self._call("nb200_finalize_max_abs", _ptr(self.hstats), _ptr(sp_i), st)
        # F5: Frobenius threshold
        self._frob_histogram(sp_i, st)
        if mask_on:
            self._call("nb200_finalize_frob", _ptr(self.hist), _ptr(self.hstats), fixed, division, _ptr(sp_i), st)
        else:
            self._call("nb200_finalize_frob_fast", _ptr(self.hist), _ptr(self.hstats), fixed, division, 0.0, 0,
                       _ptr(sp_i), st)
        # F4-F9 fused
        if self.sparse_k3:
            # candidate list: the output volume is idle until finalize() and holds one word per owned voxel
            lst = self.out if self.sparse_list else None
            self._call("nb200_frangi_sparse", _ptr(g), _ptr(code), _ptr(self.acc), C.byref(own), self._fd_c,
                       self.div_mode, float(self.p.alpha_sq), float(self.p.beta_sq), _ptr(sp_i),
                       _ptr(lst), self.out.numel(), _ptr(self.list_count), st)
        else:
            self._call("nb200_frangi_accumulate", _ptr(g), _ptr(self.acc), C.byref(own), self._fd_c, self.div_mode,
                       float(self.p.alpha_sq), float(self.p.beta_sq), _ptr(sp_i), st)

    def _frob_histogram(self, sp_i, st):
        """Histogram of the frob samples / max|H| (filtering.py:407-444); nothing to do for a fixed threshold."""
        division = float(self.p.frob_thresh_division or 0.0)
        if self.p.mask and self.p.frob_thresh is None and division != 0.0:
            div_ptr = C.c_void_p(sp_i.data_ptr() + 8 * _cabi.SP_MAX_ABS)
            self._histogram(self.samples, self.n_samples, _cabi.TF_DIV, div_ptr)
        else:
            self._call("nb200_hist_reset", _ptr(self.hist), st)

    def _analyse_sigma_fast(self, i, g, own, sp_i, fixed, division, mask_on, st):
        """F2-F9 through hessian_fast.cu.  Every exact fallback is enqueued behind a device flag: the kernels return at
        once unless the statistics pass raised sp[UNSAFE] (value range / exactness argument) or sp[AMBIG].

        Reduction points of a Z-sharded frame (``fold_state``: one all-gather of the 267-word record + one fold kernel
        each; identity on one GPU): [gauss min/max + Hessian stats] -> [gauss bins + stats of the exact redo] ->
        [frob min/max] -> [frob bins] -> [exact max frob^2 when the mask's emptiness was undecided]."""
        sz, sy, sx = self.strides
        a_sq, b_sq = float(self.p.alpha_sq), float(self.p.beta_sq)
        n = self.n_samples
        auto_thr = bool(mask_on) and self.p.frob_thresh is None and division != 0.0
        # F2/F3 (first half) + F4 statistics
        self._call("nb200_lattice_sample", _ptr(g), C.byref(own), sz, sy, sx, _ptr(self.samples), st)
        self._call("nb200_hist_reset", _ptr(self.hist), st)
        self._call("nb200_hist_minmax", _ptr(self.samples), n, _cabi.TF_NONE, None, _ptr(self.hist), st)
        self._call("nb200_hstats_reset", _ptr(self.hstats), st)
        self._call("nb200_hessian_stats_fast", _ptr(g), C.byref(own), self._fd_c, self.div_mode, sz, sy, sx,
                   _ptr(self.frob_samples), _ptr(self.hstats), _ptr(self.fast_ws), st)
        self.fold_state(self.state, _cabi.FOLD_MINMAX)
        self._call("nb200_hist_bins", _ptr(self.samples), n, _cabi.TF_NONE, None, _ptr(self.hist), st)
        self._call("nb200_finalize_max_abs", _ptr(self.hstats), _ptr(sp_i), st)
        self._call("nb200_hessian_stats_redo", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                   sz, sy, sx, _ptr(self.frob_samples), _ptr(self.hstats), _ptr(self.code), st)
        self.fold_state(self.state, _cabi.FOLD_BINS)
        self._call("nb200_finalize_gamma", _ptr(self.hist), _ptr(sp_i), st)
        self._call("nb200_finalize_max_abs", _ptr(self.hstats), _ptr(sp_i), st)
        # F5: Frobenius threshold
        self._call("nb200_hist_reset", _ptr(self.hist), st)
        if auto_thr:
            div_ptr = C.c_void_p(sp_i.data_ptr() + 8 * _cabi.SP_MAX_ABS)
            self._call("nb200_hist_minmax", _ptr(self.frob_samples), n, _cabi.TF_DIV, div_ptr, _ptr(self.hist), st)
            self.fold_state(self.state, _cabi.FOLD_MINMAX)
            self._call("nb200_hist_bins", _ptr(self.frob_samples), n, _cabi.TF_DIV, div_ptr, _ptr(self.hist), st)
            self.fold_state(self.state, _cabi.FOLD_BINS)
        self._call("nb200_finalize_frob_fast", _ptr(self.hist), _ptr(self.hstats), fixed, division, self.max_scale,
                   mask_on, _ptr(sp_i), st)
        if mask_on:
            # emptiness of the mask undecided by the bounds (cut in [0.99, 3.01)): exact max frob^2, gated on sp[AMBIG]
            self._call("nb200_hessian_stats_ambig", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                       sz, sy, sx, _ptr(self.frob_samples), _ptr(self.hstats), st)
            self.fold_state(self.state, _cabi.FOLD_MINMAX)
            self._call("nb200_finalize_frob_resolve", _ptr(self.hstats), _ptr(sp_i), st)
        self._call("nb200_frangi_fast", _ptr(g), _ptr(self.acc), C.byref(own), self._fd_c, self.div_mode, a_sq, b_sq,
                   _ptr(sp_i), _ptr(self.diag), st)
        lst = self.out if self.sparse_list else None
        self._call("nb200_frangi_sparse_gated", _ptr(g), _ptr(self.code), _ptr(self.acc), C.byref(own), self._fd_c,
                   self.div_mode, a_sq, b_sq, _ptr(sp_i), _ptr(lst), self.out.numel(), _ptr(self.list_count), st)

    def run_sigmas(self):
        """filtering.py:814-851 for every sigma; leaves max-over-sigma / dead flags in ``acc``.

        Default: blur, then K2 / thresholds / K3, sigma after sigma on the caller's stream.  With ``overlap_blur``
        the blur of sigma i+1 (which only needs the blurred volume of sigma i) is issued on a side stream under
        K2 / K3 of sigma i; three blur volumes rotate (source of truth of sigma i, scratch, result of sigma i+1) and
        CUDA events order their reuse.  Results are identical either way; on B200 the overlap buys nothing because
        K2 leaves no registers for a second kernel on the SM (see ``__init__``)."""
        if self.batch_sigmas and self.fast_path and not self.overlap_blur:
            return self._run_sigmas_batched()
        main = torch.cuda.current_stream(self.device)
        self.acc.zero_()
        nsig = len(self.steps)
        if not self.overlap_blur:
            src = self.cur
            for i in range(nsig):
                src = self._blur_sigma(i, src, {})
                self._analyse_sigma(i, self.gauss[src])
            self.cur = src
            self._remove_edges()
            return
        side = self.side_stream
        side.wait_stream(main)                       # the frame has been loaded on the caller's stream
        last_reader = {}
        src = self.cur
        for i in range(nsig):
            with torch.cuda.stream(side):
                src = self._blur_sigma(i, src, last_reader)
                ev_blur = torch.cuda.Event()
                ev_blur.record(side)
            main.wait_event(ev_blur)
            self._analyse_sigma(i, self.gauss[src])
            ev_read = torch.cuda.Event()
            ev_read.record(main)
            last_reader[src] = ev_read
        main.wait_stream(side)
        self.cur = src
        self._remove_edges()

    # -- all sigmas at once (Z-sharded frames) ------------------------------------------------------------------------
    def _blur_sigma_batched(self, i, src_idx, dst_idx, scratch_idx):
        """F1 for sigma i: ``gauss[src_idx]`` -> ``gauss[dst_idx]`` (intermediate pass through ``gauss[scratch_idx]``);
        the source is kept (every sigma's blurred volume stays resident until its K3).  Returns the index of the result
        (``src_idx`` itself when sigma i adds no blur)."""
        taps = self.steps[i]
        passes = []
        fuse_yx = (self.fuse_yx and taps[1] is not None and taps[2] is not None
                   and taps[1][1] == taps[2][1] and 1 <= taps[1][1] <= 8)
        for axis, t in enumerate(taps):
            if t is None or t[1] == 0:
                continue
            if axis == 1 and fuse_yx:
                passes.append(("yx", t))
                break
            passes.append((axis, t))
        if not passes:
            return src_idx
        st = _stream()
        rz = taps[0][1] if taps[0] is not None else 0
        self.exchange_halo(self.gauss[src_idx], rz + 2)
        v = self.vol(2, 2)
        dp = C.POINTER(C.c_double)
        cur = src_idx
        for k, (axis, (w, r)) in enumerate(passes):
            dst = dst_idx if (len(passes) - 1 - k) % 2 == 0 else scratch_idx
            if axis == "yx":
                self._call("nb200_gauss_yx", _ptr(self.gauss[cur]), _ptr(self.gauss[dst]), C.byref(v), w.ctypes.data_as(dp),
                           taps[2][0].ctypes.data_as(dp), r, st)
            else:
                self._call("nb200_gauss_axis", _ptr(self.gauss[cur]), _ptr(self.gauss[dst]), C.byref(v), axis,
                           w.ctypes.data_as(dp), r, st)
            cur = dst
        return cur

    def _run_sigmas_batched(self):
        """The fast path of ``run_sigmas`` with the loops interchanged: every stage between two reduction points runs for
        ALL sigmas, then ONE ``fold_state`` reduces the records of all sigmas (``nb200_fold_records_n``).  A Z-sharded
        frame therefore issues 5 all-gathers instead of 5 per sigma (on 8 GPUs the ~45 NCCL operations of a frame cost
        3.5 of 15 ms).  Sigmas are independent until K3 (max / AND into ``acc``, in sigma order as before), so the result
        is the one of the sigma-by-sigma loop, bit for bit; the price is one resident blurred volume per sigma."""
        st = _stream()
        nsig = len(self.steps)
        if self._batch is None:
            dev = self.device
            while len(self.gauss) < nsig + 2:
                self.gauss.append(torch.empty_like(self.gauss[0]))
            n = max(1, self.n_samples)
            self._batch = dict(
                state=torch.zeros((nsig, _cabi.STATE_WORDS), dtype=torch.int64, device=dev),
                samples=torch.empty((nsig, n), dtype=torch.float32, device=dev),
                frob=torch.empty((nsig, n), dtype=torch.float32, device=dev))
        B = self._batch
        state = B["state"]
        hist = [state[i, :_cabi.HIST_WORDS] for i in range(nsig)]
        hstats = [state[i, _cabi.HIST_WORDS:] for i in range(nsig)]
        sz, sy, sx = self.strides
        own = self.vol()
        fixed = float("nan") if self.p.frob_thresh is None else float(self.p.frob_thresh)
        division = float(self.p.frob_thresh_division or 0.0)
        mask_on = 1 if self.p.mask else 0
        auto_thr = bool(mask_on) and self.p.frob_thresh is None and division != 0.0
        a_sq, b_sq = float(self.p.alpha_sq), float(self.p.beta_sq)
        n = self.n_samples
        self.acc.zero_()
        # ---- blur chain, gamma samples, Hessian statistics ----
        if self.cur != 0:                                # the frame is loaded into gauss[0]; keep the chain's layout fixed
            raise RuntimeError("batched sigmas expect the frame in gauss[0] (load_frame)")
        src, g_idx = 0, []
        for i in range(nsig):
            src = self._blur_sigma_batched(i, src, i + 1, nsig + 1)
            g_idx.append(src)
            g = self.gauss[src]
            self._call("nb200_lattice_sample", _ptr(g), C.byref(own), sz, sy, sx, _ptr(B["samples"][i]), st)
            self._call("nb200_hist_reset", _ptr(hist[i]), st)
            self._call("nb200_hist_minmax", _ptr(B["samples"][i]), n, _cabi.TF_NONE, None, _ptr(hist[i]), st)
            self._call("nb200_hstats_reset", _ptr(hstats[i]), st)
            self._call("nb200_hessian_stats_fast", _ptr(g), C.byref(own), self._fd_c, self.div_mode, sz, sy, sx,
                       _ptr(B["frob"][i]), _ptr(hstats[i]), _ptr(self.fast_ws), st)
        self.fold_state(state, _cabi.FOLD_MINMAX)
        for i in range(nsig):
            g, sp_i = self.gauss[g_idx[i]], self.sp[i]
            self._call("nb200_hist_bins", _ptr(B["samples"][i]), n, _cabi.TF_NONE, None, _ptr(hist[i]), st)
            self._call("nb200_finalize_max_abs", _ptr(hstats[i]), _ptr(sp_i), st)
            self._call("nb200_hessian_stats_redo", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                       sz, sy, sx, _ptr(B["frob"][i]), _ptr(hstats[i]), _ptr(self.code), st)
        self.fold_state(state, _cabi.FOLD_BINS)
        # ---- gamma, max|H|, Frobenius threshold ----
        for i in range(nsig):
            sp_i = self.sp[i]
            self._call("nb200_finalize_gamma", _ptr(hist[i]), _ptr(sp_i), st)
            self._call("nb200_finalize_max_abs", _ptr(hstats[i]), _ptr(sp_i), st)
            self._call("nb200_hist_reset", _ptr(hist[i]), st)
            if auto_thr:
                div_ptr = C.c_void_p(sp_i.data_ptr() + 8 * _cabi.SP_MAX_ABS)
                self._call("nb200_hist_minmax", _ptr(B["frob"][i]), n, _cabi.TF_DIV, div_ptr, _ptr(hist[i]), st)
        if auto_thr:
            self.fold_state(state, _cabi.FOLD_MINMAX)
            for i in range(nsig):
                div_ptr = C.c_void_p(self.sp[i].data_ptr() + 8 * _cabi.SP_MAX_ABS)
                self._call("nb200_hist_bins", _ptr(B["frob"][i]), n, _cabi.TF_DIV, div_ptr, _ptr(hist[i]), st)
            self.fold_state(state, _cabi.FOLD_BINS)
        for i in range(nsig):
            g, sp_i = self.gauss[g_idx[i]], self.sp[i]
            self._call("nb200_finalize_frob_fast", _ptr(hist[i]), _ptr(hstats[i]), fixed, division, self.max_scale,
                       mask_on, _ptr(sp_i), st)
            if mask_on:
                self._call("nb200_hessian_stats_ambig", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                           sz, sy, sx, _ptr(B["frob"][i]), _ptr(hstats[i]), st)
        if mask_on:
            self.fold_state(state, _cabi.FOLD_MINMAX)
        # ---- K3, sigma after sigma (max / AND into acc) ----
        lst = self.out if self.sparse_list else None
        for i in range(nsig):
            g, sp_i = self.gauss[g_idx[i]], self.sp[i]
            if mask_on:
                self._call("nb200_finalize_frob_resolve", _ptr(hstats[i]), _ptr(sp_i), st)
            self._call("nb200_frangi_fast", _ptr(g), _ptr(self.acc), C.byref(own), self._fd_c, self.div_mode, a_sq, b_sq,
                       _ptr(sp_i), _ptr(self.diag), st)
            # the per-voxel record of the exact path is ONE volume shared by all sigmas: a sigma that fell back
            # (sp[UNSAFE]) writes it again right before its exact K3 (both kernels return at once otherwise)
            self._call("nb200_hessian_stats_redo", _ptr(g), C.byref(own), self._fd_c, self.div_mode, _ptr(sp_i),
                       sz, sy, sx, _ptr(B["frob"][i]), _ptr(hstats[i]), _ptr(self.code), st)
            self._call("nb200_frangi_sparse_gated", _ptr(g), _ptr(self.code), _ptr(self.acc), C.byref(own), self._fd_c,
                       self.div_mode, a_sq, b_sq, _ptr(sp_i), _ptr(lst), self.out.numel(), _ptr(self.list_count), st)
        self.cur = 0                                     # gauss[0] still holds the frame
        self._remove_edges()

    def _remove_edges(self):
        """filtering.py:931-932: optional (off by default) zeroing of the top / bottom rows of every slice's bounding box,
        applied to the owned planes of the accumulator before _mask_volume (per-slice: no exchange between slabs)."""
        if self.p.remove_edges:
            own = self.acc[self.pad_lo:self.pad_lo + self.nz_own]
            self._call("nb200_remove_edges", _ptr(own), self.nz_own, self.ny, self.nx, 15, _stream())

    def finalize(self, apply_mask_volume=True, out=None):
        """filtering.py:926 (V*masks) + :1014-1018 / :952-967 (_mask_volume).  ``out``: optional device buffer of
        the owned-planes shape that receives the frame instead of the engine's own ``self.out``."""
        st = _stream()
        own = self.vol()
        sz, sy, sx = self.strides
        if apply_mask_volume:
            self._call("nb200_lattice_sample", _ptr(self.acc), C.byref(own), sz, sy, sx, _ptr(self.samples), st)
            samples, n = self.gather_samples(self.samples, self.n_samples)
            self._call("nb200_percentile", _ptr(samples), n, 1.0, _ptr(self.select), _ptr(self.pct), st)
        else:
            self.pct.zero_()   # pct[1] == 0 -> pass-through
        self.exchange_halo(self.acc, 2)
        # output is written for the owned planes only; give the kernel a view whose plane 0 is buffer plane 0
        dst = self.out if out is None else out
        if tuple(dst.shape) != tuple(self.out.shape) or dst.dtype != torch.float32 or not dst.is_contiguous():
            raise ValueError("finalize: out must be a contiguous float32 tensor of the owned-planes shape")
        # the kernel indexes its output like the accumulator (buffer planes) but only writes the planes [zc0, zc1): hand
        # it the address at which buffer plane 0 WOULD sit, so that the owned planes land in `dst` without a staging copy
        base = dst.data_ptr() - self.pad_lo * self.ny * self.nx * 4
        self._call("nb200_finalize_opening", _ptr(self.acc), C.c_void_p(base), C.byref(own), _ptr(self.pct), st)
        return dst

    def filter_frame(self, frame: torch.Tensor, apply_mask_volume=True, out=None) -> torch.Tensor:
        """Device tensor in, device tensor out (the engine's own output buffer unless ``out`` is given).

        With ``use_graph`` the kernel + collective sequence of a frame (~250 launches, plus the NCCL all-gathers and
        halo exchanges of a Z-sharded run) is captured as one CUDA graph on the second call and replayed afterwards:
        on 8 GPUs a slab's kernels take 11 ms while issuing them from Python took 16 ms.  All buffers are engine-owned
        and static; only the upload of the frame stays outside the graph.  Profiling (``self.profile``) and a
        caller-provided ``out`` buffer other than the captured one fall back to eager launches."""
        self.load_frame(frame)
        key = (bool(apply_mask_volume), bool(self.p.mask), self.fast_path, None if out is None else out.data_ptr())
        if not self.use_graph or self.profile is not None:
            self.run_sigmas()
            return self.finalize(apply_mask_volume, out=out)
        if key in self._graphs:
            graph, res, launches, kernels = self._graphs[key]
            self.cur = 0
            graph.replay()
            self.launches += launches
            self.kernels += kernels
            return res
        if not self._eager_done.get(key):
            self._eager_done[key] = True                    # first call: plain launches (lazy allocations, warm caches)
            self.run_sigmas()
            return self.finalize(apply_mask_volume, out=out)
        l0, k0 = self.launches, self.kernels
        try:
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self.cur = 0
                self.run_sigmas()
                res = self.finalize(apply_mask_volume, out=out)
        except Exception:                                   # capture is an optimisation: eager launches from now on
            self.use_graph = False
            torch.cuda.synchronize(self.device)
            self.load_frame(frame)
            self.run_sigmas()
            return self.finalize(apply_mask_volume, out=out)
        self._graphs[key] = (graph, res, self.launches - l0, self.kernels - k0)
        self.cur = 0
        graph.replay()                                      # capture does not execute: run the frame now
        return res

    def mask_volume(self, v: torch.Tensor) -> torch.Tensor:
        """_mask_volume (filtering.py:952-967) of an already computed response ``v`` (>= 0)."""
        self.acc[self.pad_lo:self.pad_lo + self.nz_own].copy_(v)
        return self.finalize(True)

    def sigma_records(self):
        """Per-sigma device scalars copied to the host (tests / diagnostics; forces a sync)."""
        return self.sp.cpu().numpy()
