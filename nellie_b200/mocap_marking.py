"""B200-native drop-in for ``nellie.segmentation.mocap_marking.Markers`` (reference: mocap_marking.py:17-836).

Same constructor keywords, helper names and ``.run()`` contract: reads ``im_instance_label`` and the raw image (and
``im_preprocessed`` for ``use_im='frangi'``) through the ``im_info`` memmaps and writes ``im_marker`` (uint8),
``im_distance`` (float32) and ``im_border`` (uint8) per frame.  Per frame (mocap_marking.py:648-703):

    mask = labels > 0; border shell; exact Euclidean distance transform clamped to 2 * max_radius_px;
    per scale: -LoG * sigma^2 of the distance (or Frangi) image, 3^d local maxima inside the mask, best scale wins;
    non-maximum suppression of the peaks on the raw intensity.

All of it runs in CUDA kernels behind ``include/nellie_b200.h`` (``nb200_markers_*`` in csrc/markers.cu, the Gaussian
derivative passes of ``scipy.ndimage.gaussian_laplace`` through ``nb200_gauss_axis`` / ``nb200_gauss_yx`` with order-2 taps);
the outputs are bit-identical to the reference's.  Only the full-volume branch is reproduced — the reference's own test
(tests/test_mocap_marking.py) asserts that its chunked low-memory branch gives the same result; ``low_memory`` /
``max_chunk_voxels`` are accepted and ignored.  There is no CPU path: ``device='cpu'`` raises.  One pathological input
differs: a label frame without a single background voxel gets the clamp as distance everywhere, where scipy measures to the
virtual voxel (-1, -1, -1).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import logging
import math

import numpy as np
import torch

from . import _cabi
from .engine import gaussian_taps
from .engine2d import gaussian_taps_order2

logger = logging.getLogger("nellie_b200")

_DEVICES = ("auto", "gpu", "cuda", "b200")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _on(device):
    """Make ``device`` the current CUDA device for the enclosed launches."""
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


class MarkerEngine:
    """Device buffers + kernel sequence for frames of one shape on one GPU.

    ``lib`` / ``device``: the C ABI and the torch device holding the buffers.  The product always passes the CUDA
    library and a CUDA device; the CPU test-suite injects the host build of the same kernels (oracle/markers_host.cpp)
    with ``device='cpu'`` to check this sequence without a GPU."""

    def __init__(self, frame_shape, no_z, sigmas, z_ratio, max_radius_px, peak_min_distance, device, lib=None,
                 truncate=4.0):
        self.lib = _cabi.load() if lib is None else lib
        self.device = torch.device(device)
        if lib is None and self.device.type != "cuda":
            raise RuntimeError("nellie_b200 has no CPU path: %s needs a CUDA device (a host device is only accepted "
                               "together with the test suite's emulated kernel library)" % type(self).__name__)
        self.shape = tuple(int(s) for s in frame_shape)
        self.no_z = bool(no_z)
        if self.no_z:
            assert len(self.shape) == 2
            self.nz, (self.ny, self.nx) = 1, self.shape
        else:
            assert len(self.shape) == 3
            self.nz, self.ny, self.nx = self.shape
        self.n = self.nz * self.ny * self.nx
        self.sigmas = [float(s) for s in sigmas]
        self.z_ratio = float(z_ratio)
        self.radius = int(peak_min_distance)                        # mocap_marking.py:599
        # np.minimum(float32 array, python float): the scalar is applied as float32 (mocap_marking.py:447)
        self.clamp = float(np.float32(float(max_radius_px) * 2.0))
        self.window = max(1, int(math.ceil(self.clamp)))
        # taps of scipy's gaussian_filter1d per scale and axis (Z, Y, X): (order-0 taps, order-2 taps, radius)
        self.taps = []
        for s in self.sigmas:
            vec = (s, s) if self.no_z else (s / self.z_ratio, s, s)       # mocap_marking.py:318-338
            if min(vec) <= 1e-15:
                raise NotImplementedError("scipy skips axes with sigma <= 1e-15; not reproduced")
            per_axis = []
            for sd in vec:
                w0, r = gaussian_taps(sd, truncate)
                w2, r2 = gaussian_taps_order2(sd, truncate)
                assert r == r2
                per_axis.append((w0, w2, r))
            if self.no_z:
                per_axis = [None] + per_axis
            self.taps.append(per_axis)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.mask = torch.empty(self.shape, **u8)
        self.border = torch.empty(self.shape, **u8)
        self.peak = torch.empty(self.shape, **u8)
        self.marker = torch.empty(self.shape, **u8)
        self.distance = torch.empty(self.shape, **f32)
        self.best = torch.empty(self.shape, **f32)
        self.scratch = torch.empty(2 * self.n, dtype=torch.int16, device=dev)     # uint16 squared distances
        self.t = [torch.empty(self.shape, **f32) for _ in range(3)]                # Z passes / unfused Y pass
        self.d = [torch.empty(self.shape, **f32) for _ in range(2 if self.no_z else 3)]
        self.vol = _cabi.Vol.whole(self.nz, self.ny, self.nx)
        self.launches = 0

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        self.launches += 1
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            _cabi.check(rc, name)

    def _axis(self, src, dst, axis, w, r):
        self._call("nb200_gauss_axis", _ptr(src), _ptr(dst), C.byref(self.vol), axis,
                   w.ctypes.data_as(C.POINTER(C.c_double)), r, self._stream())

    def _yx(self, src, dst, wy, wx, r, tmp):
        """Y pass then X pass (scipy stores float32 between them); one fused kernel for radii 1..8."""
        dp = C.POINTER(C.c_double)
        if 1 <= r <= 8:
            self._call("nb200_gauss_yx", _ptr(src), _ptr(dst), C.byref(self.vol), wy.ctypes.data_as(dp),
                       wx.ctypes.data_as(dp), r, self._stream())
        else:
            self._axis(src, tmp, 1, wy, r)
            self._axis(tmp, dst, 2, wx, r)

    # ---- stage steps ------------------------------------------------------------------------------------------------
    def distance_and_border(self, labels_i32):
        """mocap_marking.py:419-450 on the label frame: fills ``mask``, ``border``, ``distance``."""
        st = self._stream()
        self._call("nb200_markers_mask_border", _ptr(labels_i32), self.nz, self.ny, self.nx, _ptr(self.mask),
                   _ptr(self.border), st)
        self._call("nb200_markers_edt", _ptr(self.mask), self.nz, self.ny, self.nx, self.window, C.c_float(self.clamp),
                   _ptr(self.scratch), _ptr(self.distance), st)
        return self.distance, self.border

    def laplace_terms(self, base, taps):
        """The separable terms of ``scipy.ndimage.gaussian_laplace(base, sigma_vec)`` in axis order: term a = order 2
        along axis a, order 0 along the others, axes filtered in the order Z, Y, X with a float32 store after each."""
        if self.no_z:
            _, (w0y, w2y, r), (w0x, w2x, _) = taps
            self._yx(base, self.d[0], w2y, w0x, r, self.t[0])
            self._yx(base, self.d[1], w0y, w2x, r, self.t[0])
            return self.d[0], self.d[1], None
        (w0z, w2z, rz), (w0y, w2y, r), (w0x, w2x, _) = taps
        self._axis(base, self.t[0], 0, w2z, rz)
        self._yx(self.t[0], self.d[0], w0y, w0x, r, self.t[2])
        self._axis(base, self.t[1], 0, w0z, rz)            # shared by the Y and the X term
        self._yx(self.t[1], self.d[1], w2y, w0x, r, self.t[2])
        self._yx(self.t[1], self.d[2], w0y, w2x, r, self.t[2])
        return self.d[0], self.d[1], self.d[2]

    def peaks(self, base, fused=True):
        """mocap_marking.py:452-512 (_local_max_peak): fills ``peak`` (uint8) and ``best``.  ``fused`` (default): response
        and local-maximum test in one kernel, no response volume; ``fused=False``: the two-step form (same result)."""
        st = self._stream()
        self.best.zero_()
        self.peak.zero_()
        for s, taps in zip(self.sigmas, self.taps):
            d0, d1, d2 = self.laplace_terms(base, taps)
            sigma_sq = C.c_float(float(np.float32(s ** 2)))       # float32 array * python float (mocap_marking.py:490)
            if fused:
                self._call("nb200_markers_peak_update_fused", _ptr(d0), _ptr(d1), _ptr(d2), sigma_sq, _ptr(self.mask),
                           _ptr(self.distance), self.nz, self.ny, self.nx, _ptr(self.best), _ptr(self.peak), st)
                continue
            self._call("nb200_markers_log_response", _ptr(d0), _ptr(d1), _ptr(d2), self.n, sigma_sq, _ptr(d0), st)
            self._call("nb200_markers_peak_update", _ptr(d0), _ptr(self.mask), _ptr(self.distance), self.nz, self.ny,
                       self.nx, _ptr(self.best), _ptr(self.peak), st)
        return self.peak

    def suppress(self, peak_u8, intensity_f32):
        """mocap_marking.py:569-606 (_remove_close_peaks): fills ``marker``."""
        self._call("nb200_markers_nms", _ptr(peak_u8), _ptr(intensity_f32), self.nz, self.ny, self.nx, self.radius,
                   _ptr(self.marker), self._stream())
        return self.marker

    def run_frame(self, labels_i32, intensity_f32, frangi_f32=None):
        """mocap_marking.py:648-703: returns the engine's (marker uint8, distance float32, border uint8) tensors."""
        self.distance_and_border(labels_i32)
        base = self.distance if frangi_f32 is None else frangi_f32
        self.peaks(base)
        self.suppress(self.peak, intensity_f32)
        return self.marker, self.distance, self.border


class Markers:
    def __init__(self, im_info, num_t=None, min_radius_um=0.20, max_radius_um=1, use_im="distance", num_sigma=5,
                 viewer=None, prefer_gpu=True, peak_min_distance=2, device="auto", low_memory=False,
                 max_chunk_voxels=int(1e6), cuda_device=None, t_shard=None, fallback=None, z_shard=None):
        dev = (device or "auto").lower()
        if dev == "cpu" or (dev == "auto" and not prefer_gpu):
            raise ValueError("nellie_b200.Markers implements the CUDA path only; device='cpu' belongs to "
                             "nellie.segmentation.mocap_marking.Markers")
        if dev not in _DEVICES:
            raise ValueError(f"Unsupported device '{device}'. Use 'auto', 'gpu' or 'b200'.")
        if use_im not in ("distance", "frangi"):
            raise ValueError(f"Unknown use_im value: {use_im}")
        self.im_info = im_info
        self.num_t = num_t
        if self.im_info.no_t:
            self.num_t = 1
        elif num_t is None:
            self.num_t = im_info.shape[im_info.axes.index("T")]
        # mocap_marking.py:124-135
        x_res = self.im_info.dim_res.get("X") or 1.0
        z_res = self.im_info.dim_res.get("Z") or x_res
        self.z_ratio = float(z_res) / float(x_res) if not self.im_info.no_z else 1.0
        self.min_radius_um = max(min_radius_um, float(x_res))
        self.max_radius_um = max_radius_um
        self.min_radius_px = self.min_radius_um / float(x_res)
        self.max_radius_px = self.max_radius_um / float(x_res)
        self.use_im = use_im
        self.num_sigma = num_sigma
        self.sigmas = []
        self.shape = ()
        self.im_memmap = None
        self.im_frangi_memmap = None
        self.label_memmap = None
        self.im_marker_memmap = None
        self.im_distance_memmap = None
        self.im_border_memmap = None
        self.debug = None
        self.viewer = viewer
        self.device = device or "auto"
        self.device_type = "cuda"
        self.use_gpu = True
        self.peak_min_distance = peak_min_distance
        self.low_memory = bool(low_memory)
        self.max_chunk_voxels = int(max_chunk_voxels)
        self.truncate = 4.0
        self._cuda_device = cuda_device
        self.t_shard = None if t_shard is None else (int(t_shard[0]), int(t_shard[1]))
        # Z-sharding (rank, world) of every frame (for frames too large for one GPU): the stage has no global
        # step, so rank r runs the ordinary single-GPU sequence on its planes plus a halo and keeps its own planes
        # (_z_extent); no collective, the ranks only share the output files
        self.z_shard = None if z_shard is None else (int(z_shard[0]), int(z_shard[1]))
        if self.z_shard is not None and t_shard is not None:
            raise ValueError("a Markers stage is sharded over T or over Z, not both")
        if self.z_shard is not None and im_info.no_z:
            raise ValueError("2-D frames are T-sharded; Z-sharding needs a Z axis")
        self.fallback = fallback
        self._engine = None
        self._ctor_kwargs = dict(num_t=num_t, min_radius_um=min_radius_um, max_radius_um=max_radius_um, use_im=use_im,
                                 num_sigma=num_sigma, viewer=viewer, peak_min_distance=peak_min_distance,
                                 max_chunk_voxels=max_chunk_voxels)
        if low_memory:
            logger.warning("nellie_b200.Markers: low_memory is accepted for compatibility and ignored (the reference's "
                           "chunked branch computes the same result, tests/test_mocap_marking.py)")
        if fallback != "reference":
            _cabi.load()

    # ---- host scalars (reference: mocap_marking.py:318-390) -----------------------------------------------------------
    def _get_sigma_vec(self, sigma):
        if self.im_info.no_z:
            return (sigma, sigma)
        return (sigma / self.z_ratio, sigma, sigma)

    def _set_default_sigmas(self):
        min_sigma_step_size = 0.2
        self.sigma_min = self.min_radius_px / 2.0
        self.sigma_max = self.max_radius_px / 3.0
        sigma_range = self.sigma_max - self.sigma_min
        if sigma_range <= 0:
            logger.warning("Non-positive sigma range (min=%f, max=%f). Check radius settings.", self.sigma_min,
                           self.sigma_max)
            self.sigmas = [self.sigma_min]
            return
        sigma_step_size = max(min_sigma_step_size, sigma_range / max(self.num_sigma, 1))
        self.sigmas = list(np.arange(self.sigma_min, self.sigma_max, sigma_step_size))
        if len(self.sigmas) == 0:
            self.sigmas = [self.sigma_min]
            logger.warning("No sigma values generated; falling back to a single sigma=%f.", self.sigma_min)

    def _get_t(self):
        if self.num_t is None:
            self.num_t = 1 if self.im_info.no_t else self.im_info.shape[self.im_info.axes.index("T")]

    def _allocate_memory(self):
        """mocap_marking.py:371-417; a T-sharded stage creates its three output files on rank 0 only."""
        from .sharding import allocate_shared_output
        paths = self.im_info.pipeline_paths
        self.label_memmap = self.im_info.get_memmap(paths["im_instance_label"])
        self.im_memmap = self.im_info.get_memmap(self.im_info.im_path)
        self.shape = self.label_memmap.shape
        self.im_frangi_memmap = self.im_info.get_memmap(paths["im_preprocessed"]) if self.use_im == "frangi" else None
        shard = self.t_shard or self.z_shard
        self.im_marker_memmap = allocate_shared_output(self.im_info, paths["im_marker"], "uint8", "mocap marker image", shard)
        self.im_distance_memmap = allocate_shared_output(self.im_info, paths["im_distance"], "float32",
                                                         "distance transform image", shard)
        self.im_border_memmap = allocate_shared_output(self.im_info, paths["im_border"], "uint8", "border image", shard)

    # ---- device plumbing ------------------------------------------------------------------------------------------------
    def _torch_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested but CUDA is not available. (nellie_b200 has no CPU path)")
        if self._cuda_device is not None:
            return torch.device(self._cuda_device)
        return torch.device("cuda", torch.cuda.current_device())

    def _engine_for(self, frame_shape):
        key = tuple(int(s) for s in frame_shape)
        if not self.sigmas:
            self._set_default_sigmas()
        sig = tuple(float(s) for s in self.sigmas)
        if self._engine is None or self._engine.shape != key or tuple(self._engine.sigmas) != sig:
            dev = self._torch_device()
            with _on(dev):
                self._engine = MarkerEngine(key, self.im_info.no_z, sig, self.z_ratio, self.max_radius_px,
                                            self.peak_min_distance, dev, truncate=self.truncate)
        return self._engine

    def _dev(self, arr, dtype):
        """Host frame (memmap slice / ndarray) or tensor -> contiguous device tensor of ``dtype``."""
        if isinstance(arr, torch.Tensor):
            return arr.to(self._torch_device(), dtype=dtype).contiguous()
        a = np.asarray(arr)
        if not a.dtype.isnative:
            a = a.astype(a.dtype.newbyteorder("="))
        if a.dtype in (np.uint32, np.uint64):
            a = a.astype(np.int64)
        # uint16 frames travel as uint16 (half the bytes of a host-side widening) and are cast on the device
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self._torch_device())
        return t.to(dtype).contiguous()

    def _labels_dev(self, labels):
        """The label frame as the int32 the kernel thresholds (``> 0``)."""
        if isinstance(labels, torch.Tensor):
            t = labels.to(self._torch_device())
        else:
            a = np.asarray(labels)
            if not a.dtype.isnative:
                a = a.astype(a.dtype.newbyteorder("="))
            if a.dtype.kind == "u" and a.dtype.itemsize > 1:
                a = a > 0                               # keep ids above the int32 range positive
            t = torch.from_numpy(np.ascontiguousarray(a)).to(self._torch_device())
        if t.dtype != torch.int32:
            t = (t > 0).to(torch.int32)
        return t.contiguous()

    # ---- stage steps with the reference's signatures ---------------------------------------------------------------------
    def _distance_im(self, mask):
        """mocap_marking.py:419-450: (distance float32, border bool) for a bool mask; numpy in -> numpy out."""
        was_np = not isinstance(mask, torch.Tensor)
        eng = self._engine_for(mask.shape)
        with _on(eng.device):
            distance, border = eng.distance_and_border(self._labels_dev(mask))
            if was_np:
                return distance.cpu().numpy(), border.cpu().numpy().astype(bool)
            return distance.clone(), border.to(torch.bool)

    def _local_max_peak(self, use_im, mask, distance_im, low_memory=False, chunk_voxels=None):
        """mocap_marking.py:452-512: coordinates (N, ndim) of the multi-scale LoG peaks, raster order."""
        was_np = not isinstance(use_im, torch.Tensor)
        eng = self._engine_for(use_im.shape)
        with _on(eng.device):
            eng.mask.copy_(self._dev(mask, torch.uint8) != 0)
            eng.distance.copy_(self._dev(distance_im, torch.float32))
            coords = torch.nonzero(eng.peaks(self._dev(use_im, torch.float32)))
            return coords.cpu().numpy() if was_np else coords

    def _remove_close_peaks(self, coords, intensity_im, low_memory=False, chunk_voxels=None):
        """mocap_marking.py:569-606: coordinates of the peaks that survive the suppression on ``intensity_im``."""
        was_np = not isinstance(coords, torch.Tensor)
        if (coords.numel() if isinstance(coords, torch.Tensor) else coords.size) == 0:
            return coords
        eng = self._engine_for(intensity_im.shape)
        with _on(eng.device):
            c = self._dev(coords, torch.int64)
            eng.peak.zero_()
            eng.peak[tuple(c.T)] = 1
            kept = torch.nonzero(eng.suppress(eng.peak, self._dev(intensity_im, torch.float32)))
            return kept.cpu().numpy() if was_np else kept

    def marker_frame_device(self, labels, intensity, frangi=None):
        """Device-resident frame: (marker uint8, distance float32, border uint8) tensors of the engine."""
        eng = self._engine_for(tuple(labels.shape))
        with _on(eng.device):
            if self.use_im == "frangi" and frangi is None:
                raise RuntimeError("Frangi image requested for peak detection but not available.")
            return eng.run_frame(self._labels_dev(labels), self._dev(intensity, torch.float32),
                                 self._dev(frangi, torch.float32) if self.use_im == "frangi" else None)

    def z_halo(self):
        """Planes a Z slab needs beyond its own for its own planes to come out exactly as in the whole frame: the kept
        peaks of a plane depend on peaks within ``d`` planes (suppression window), those on the response within one more
        plane (3^d maximum), the response on the distance image within the Z radius of the widest kernel, and the clamped
        distance on the label field within the scan window.  Reflection at the slab's own ends only touches planes
        outside that chain; at the frame's ends the slab's end IS the frame's end."""
        if not self.sigmas:
            self._set_default_sigmas()
        clamp = float(np.float32(float(self.max_radius_px) * 2.0))
        window = max(1, int(math.ceil(clamp)))
        rz = max(int(self.truncate * (float(s) / self.z_ratio) + 0.5) for s in self.sigmas)
        return window + rz + int(self.peak_min_distance) + 1

    def _z_extent(self, nz):
        """(first plane, last plane + 1) of the slab this rank computes and of the part it owns."""
        from .sharding import z_partition
        rank, world = self.z_shard
        z0, z1 = z_partition(nz, world)[rank]
        h = self.z_halo()
        return max(0, z0 - h), min(nz, z1 + h), z0, z1

    def _run_frame_impl(self, t, low_memory=False, chunk_voxels=None):
        """mocap_marking.py:648-703: numpy (marker uint8, distance float32, border uint8) of frame ``t`` — of this rank's
        own planes when the stage is Z-sharded."""
        logger.info("Running motion capture marking, volume %s/%s", t, (self.num_t or 1) - 1)
        if self.use_im == "frangi" and self.im_frangi_memmap is None:
            raise RuntimeError("Frangi image requested for peak detection but not available.")
        if self.z_shard is None:
            frangi = self.im_frangi_memmap[t] if self.use_im == "frangi" else None
            marker, distance, border = self.marker_frame_device(self.label_memmap[t], self.im_memmap[t], frangi)
            return marker.cpu().numpy(), distance.cpu().numpy(), border.cpu().numpy()
        e0, e1, z0, z1 = self._z_extent(int(self.label_memmap.shape[1]))
        frangi = self.im_frangi_memmap[t, e0:e1] if self.use_im == "frangi" else None
        marker, distance, border = self.marker_frame_device(self.label_memmap[t, e0:e1], self.im_memmap[t, e0:e1], frangi)
        own = slice(z0 - e0, z1 - e0)
        return marker[own].cpu().numpy(), distance[own].cpu().numpy(), border[own].cpu().numpy()

    def _run_frame(self, t):
        return self._run_frame_impl(t)

    def _run_mocap_marking(self):
        """T loop of mocap_marking.py:756-785."""
        from .pipeline import parallel_copyto
        from .sharding import frames_of_rank
        frames = range(self.num_t) if self.t_shard is None else frames_of_rank(self.num_t, *self.t_shard)
        outs = (self.im_marker_memmap, self.im_distance_memmap, self.im_border_memmap)
        for t in frames:
            if self.viewer is not None:
                self.viewer.status = f"Mocap marking. Frame: {t + 1} of {self.num_t}."
            results = self._run_frame(t)
            whole = self.im_marker_memmap.shape != self.shape and self.im_info.no_t      # mocap_marking.py:767
            own = slice(None) if self.z_shard is None else slice(*self._z_extent(int(self.label_memmap.shape[1]))[2:])
            for mm, frame in zip(outs, results):
                if mm is None:
                    continue
                if whole:
                    mm[own] = frame
                else:
                    parallel_copyto(mm[t][own], frame)
                if hasattr(mm, "flush"):
                    mm.flush()

    def _run_b200(self):
        self._torch_device()
        _cabi.load()
        self._get_t()
        self._allocate_memory()
        self._set_default_sigmas()
        self._run_mocap_marking()

    def _run_reference(self, device, low_memory):
        from nellie.segmentation.mocap_marking import Markers as ReferenceMarkers
        ReferenceMarkers(self.im_info, device=device, low_memory=low_memory, **self._ctor_kwargs).run()

    def run(self):
        logger.info("Running motion capture marking (nellie_b200).")
        from .adaptive import run_with_ladder
        run_with_ladder("Markers", self._run_b200, self._run_reference, self.fallback)
