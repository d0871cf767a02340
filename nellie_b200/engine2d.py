"""Device-side orchestration of one 2-D frame of the Filter path (BASELINE config #4 shape class).

Reference control flow: filtering.py:806-853 (sigma loop, closed-form 2x2 eigenvalues :676-690),
:924-930 (LoG blobness on the blurred frame — SURVEY App. C-2), :952-967 (_mask_volume).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _cabi
from ._cabi import Vol
from .engine import FilterParams, _ptr, _stream, gaussian_taps, sample_strides


def gaussian_taps_order2(sigma: float, truncate: float = 4.0):
    """scipy.ndimage._gaussian_kernel1d(sigma, order=2, radius=int(truncate*sigma+0.5)), restated with
    the same numpy operations so the taps are bit-identical; returns (w[0..r], r)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    order = 2
    exponent_range = np.arange(order + 1)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    q = np.zeros(order + 1)
    q[0] = 1
    D = np.diag(exponent_range[1:], 1)
    P = np.diag(np.ones(order) / -sigma2, -1)
    Q = D + P
    for _ in range(order):
        q = Q.dot(q)
    q = (x[:, None] ** exponent_range).dot(q)
    w = q * phi
    return np.ascontiguousarray(w[radius:], dtype=np.float64), radius


class FrangiEngine2D:
    def __init__(self, shape, params: FilterParams, device=None):
        if not params.no_z or len(shape) != 2:
            raise ValueError("FrangiEngine2D needs a (Y, X) frame and no_z=True")
        self.p = params
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = _cabi.load()
        self.ny, self.nx = int(shape[0]), int(shape[1])
        if min(self.ny, self.nx) < 2:
            raise ValueError("numpy.gradient needs at least 2 samples along every axis")
        self.n = self.ny * self.nx
        self.sigmas = params.sigma_list()
        self.steps = []
        prev = 0.0
        for s in self.sigmas:
            dvec = params.delta_sigma_vec(prev, s)
            self.steps.append([gaussian_taps(d, params.truncate) if d > 1e-15 else None for d in dvec])
            prev = s
        # LoG taps per sigma: (order-0 taps, order-2 taps), truncate 4.0 (scipy default)
        self.log_taps = [(gaussian_taps(s, 4.0), gaussian_taps_order2(s, 4.0)) for s in self.sigmas]
        self.strides = sample_strides((self.ny, self.nx), params.max_threshold_samples)
        sy, sx = self.strides
        self.n_samples = math.ceil(self.ny / sy) * math.ceil(self.nx / sx)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.gauss = [torch.empty((self.ny, self.nx), **f32) for _ in range(2)]
        self.t = [torch.empty((self.ny, self.nx), **f32) for _ in range(3)]
        self.acc = torch.empty((self.ny, self.nx), **f32)
        self.L = torch.empty((self.ny, self.nx), **f32)
        self.v = torch.empty((self.ny, self.nx), **f32)
        self.out = torch.empty((self.ny, self.nx), **f32)
        self.samples = torch.empty(max(1, self.n_samples), **f32)
        self.hist = torch.zeros(_cabi.HIST_WORDS, dtype=torch.int64, device=dev)
        self.hstats = torch.zeros(_cabi.HS_WORDS, dtype=torch.int64, device=dev)
        self.word = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sp = torch.zeros((len(self.sigmas), _cabi.SP_WORDS), dtype=torch.float64, device=dev)
        self.select = torch.zeros(_cabi.SELECT_WORDS, dtype=torch.int64, device=dev)
        self.pct = torch.zeros(2, dtype=torch.float64, device=dev)
        self.fd = params.fd_spacing_f32()
        self._fd_c = self.fd.ctypes.data_as(C.POINTER(C.c_float))
        self.vol = Vol(1, self.ny, self.nx, 0, 1, 0, 1)
        self.launches = 0
        self.profile = None
        self.use_graph = True      # replay the per-frame kernel sequence as a CUDA graph (see filter_frame)
        self._graphs, self._graph_calls, self._eager_done = {}, {}, {}

    def _call(self, name, *args):
        self.launches += 1
        _cabi.check(getattr(self.lib, name)(*args), name)

    def _blur_axis(self, src, dst, axis, taps):
        w, r = taps
        self._call("nb200_gauss_axis", _ptr(src), _ptr(dst), C.byref(self.vol), axis,
                   w.ctypes.data_as(C.POINTER(C.c_double)), r, _stream())

    def _histogram(self, n, transform, divisor_ptr):
        st = _stream()
        self._call("nb200_hist_reset", _ptr(self.hist), st)
        self._call("nb200_hist_minmax", _ptr(self.samples), n, transform, divisor_ptr, _ptr(self.hist), st)
        self._call("nb200_hist_bins", _ptr(self.samples), n, transform, divisor_ptr, _ptr(self.hist), st)

    def filter_frame(self, frame: torch.Tensor, apply_mask_volume=True, out=None) -> torch.Tensor:
        """One 2-D frame.  A 2048^2 frame is ~180 small kernels (3 ms of GPU time, launch-bound), so from the third
        call on the whole per-frame sequence is replayed as ONE CUDA graph (captured on the second call, after an
        eager warm-up; all buffers are engine-owned and static, only the upload of the frame stays outside)."""
        key = (bool(apply_mask_volume), bool(getattr(self.p, "mask", True)))
        self.gauss[0].copy_(frame)
        if not self.use_graph:
            res = self._filter_frame(apply_mask_volume)
        elif key in self._graphs:
            graph, res = self._graphs[key]
            graph.replay()
            self.launches += self._graph_calls[key]
        elif self._eager_done.get(key):
            res = self._capture(key)
        else:
            res = self._filter_frame(apply_mask_volume)
            self._eager_done[key] = True
        if out is None:
            return res
        out.copy_(res)
        return out

    def _capture(self, key):
        """Capture the per-frame sequence (stream capture of the C-ABI launches), then replay it once."""
        before = self.launches
        try:
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                res = self._filter_frame(key[0])
        except Exception:                       # capture is an optimisation: fall back to eager launches for good
            self.use_graph = False
            torch.cuda.synchronize(self.device)
            return self._filter_frame(key[0])
        self._graph_calls[key] = self.launches - before
        self._graphs[key] = (graph, res)
        graph.replay()
        return res

    def _filter_frame(self, apply_mask_volume=True) -> torch.Tensor:
        """The kernel sequence of one frame; the frame is already in ``gauss[0]``."""
        st = _stream()
        cur = 0
        self.acc.zero_()
        sy, sx = self.strides
        for i, taps in enumerate(self.steps):
            sp_i = self.sp[i]
            for axis, t in enumerate(taps):          # axes Y (1), X (2) of a 1-plane volume
                if t is None or t[1] == 0:
                    continue
                self._blur_axis(self.gauss[cur], self.gauss[1 - cur], axis + 1, t)
                cur = 1 - cur
            g = self.gauss[cur]
            self._call("nb200_lattice_sample", _ptr(g), C.byref(self.vol), 1, sy, sx, _ptr(self.samples), st)
            self._histogram(self.n_samples, _cabi.TF_NONE, None)
            self._call("nb200_finalize_gamma", _ptr(self.hist), _ptr(sp_i), st)
            self._call("nb200_hstats_reset", _ptr(self.hstats), st)
            self._call("nb200_hessian_stats_2d", _ptr(g), self.ny, self.nx, self._fd_c, sy, sx, _ptr(self.samples),
                       _ptr(self.hstats), st)
            self._call("nb200_finalize_max_abs", _ptr(self.hstats), _ptr(sp_i), st)
            fixed = float("nan") if self.p.frob_thresh is None else float(self.p.frob_thresh)
            division = float(self.p.frob_thresh_division or 0.0)
            if self.p.frob_thresh is None and division != 0.0:
                self._histogram(self.n_samples, _cabi.TF_DIV, C.c_void_p(sp_i.data_ptr() + 8 * _cabi.SP_MAX_ABS))
            else:
                self._call("nb200_hist_reset", _ptr(self.hist), st)
            self._call("nb200_finalize_frob_fast", _ptr(self.hist), _ptr(self.hstats), fixed, division, 0.0,
                       1 if getattr(self.p, "mask", True) else 0, _ptr(sp_i), st)
            self._call("nb200_frangi_accumulate_2d", _ptr(g), _ptr(self.acc), self.ny, self.nx, self._fd_c,
                       float(self.p.beta_sq), _ptr(sp_i), st)
        # F10: LoG blobness on the sigma_max-blurred frame
        g = self.gauss[cur]
        for i, (s, (t_o0, t_o2)) in enumerate(zip(self.sigmas, self.log_taps)):
            a, b, c = self.t
            self._blur_axis(g, c, 1, t_o2)      # T0: order 2 along Y, then order 0 along X
            self._blur_axis(c, a, 2, t_o0)
            self._blur_axis(g, c, 1, t_o0)      # T1: order 0 along Y, then order 2 along X
            self._blur_axis(c, b, 2, t_o2)
            self._call("nb200_log2d_accumulate", _ptr(a), _ptr(b), _ptr(self.acc), float(np.float32(float(s) ** 2)),
                       int(i == 0), self.n, _ptr(self.L), st)
        self._call("nb200_log2d_combine", _ptr(self.acc), _ptr(self.L), self.n, _ptr(self.word), _ptr(self.v), st)
        if self.p.remove_edges:                 # filtering.py:931-932 (off by default)
            self._call("nb200_remove_edges", _ptr(self.v), 1, self.ny, self.nx, 15, st)
        if not apply_mask_volume:
            return self.v
        return self.mask_volume(self.v)

    def mask_volume(self, v: torch.Tensor) -> torch.Tensor:
        st = _stream()
        sy, sx = self.strides
        self._call("nb200_lattice_sample", _ptr(v), C.byref(self.vol), 1, sy, sx, _ptr(self.samples), st)
        self._call("nb200_percentile", _ptr(self.samples), self.n_samples, 1.0, _ptr(self.select), _ptr(self.pct), st)
        self._call("nb200_finalize_opening_2d", _ptr(v), _ptr(self.out), self.ny, self.nx, _ptr(self.pct), st)
        return self.out

    def sigma_records(self):
        return self.sp.cpu().numpy()
