"""B200-native per-frame feature extraction of ``nellie.tracking.hu_tracking.HuMomentTracking`` (SURVEY §8f-4;
reference: hu_tracking.py:225-421, :544-750).

``HuMomentFeatures`` has the constructor keywords of the reference class that matter for features and its
``_get_frame_features(t)`` contract: for frame ``t`` of the raw image, ``im_preprocessed``, ``im_distance`` and ``im_marker``
it returns ``_FrameFeatures(coords_voxel, coords_phys, stats, hu)`` — the marker voxels in raster order, their physical
coordinates, ``stats`` (N, 4) float32 = mean / variance of the non-zero ROI voxels of the raw and of the log-transformed
Frangi frame, and ``hu`` (N, 6 | 18) = the log-Hu invariants of the ROI (2-D) or of its three maximum projections (3-D).
The matching / flow-vector part of the tracker (cost matrix, assignment, interpolation) stays with the reference; a
maintainer plugs these features in by overriding ``HuMomentTracking._get_frame_features``.

Parity: coordinates and ``stats`` are bit-identical to the reference in both of its modes (dense ROI cube: numpy reduces a
zero-padded cube; streaming: the ROI box), for float32 and for uint8 / uint16 raw frames; ``hu`` agrees to float64 rounding
(numpy's SIMD ``pow`` is not reproducible), see csrc/hu.cu.  No CPU path: ``device='cpu'`` raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import logging
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi

logger = logging.getLogger("nellie_b200")

_DEVICES = ("auto", "gpu", "cuda", "b200")


@dataclass
class _FrameFeatures:
    """hu_tracking.py:21-28."""
    coords_voxel: np.ndarray
    coords_phys: np.ndarray
    stats: object
    hu: object


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _on(device):
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def integer_bits(dtype) -> int:
    """0 for a floating-point frame, 8 / 16 for uint8 / uint16 (numpy's integer arithmetic is reproduced for these)."""
    dt = np.dtype(str(dtype).replace("torch.", "")) if isinstance(dtype, torch.dtype) else np.dtype(dtype)
    if dt.kind == "f":
        if dt != np.float32:
            raise NotImplementedError(f"raw frames of type {dt} are not supported (float32, uint8, uint16)")
        return 0
    if dt == np.uint8:
        return 8
    if dt == np.uint16:
        return 16
    raise NotImplementedError(f"raw frames of type {dt} are not supported (float32, uint8, uint16): numpy's integer "
                              "overflow rules differ per type")


class HuFeatureEngine:
    """Kernel sequence for frames of one shape.  ``lib`` / ``device``: the product passes the CUDA library and a CUDA
    device; the CPU tests inject the host build of the same kernels (oracle/hu_host.cpp) with ``device='cpu'``."""

    PROJ_BYTES = 1 << 30          # markers are processed in groups whose projection scratch stays below this

    def __init__(self, frame_shape, no_z, device, lib=None):
        self.lib = _cabi.load() if lib is None else lib
        self.device = torch.device(device)
        if lib is None and self.device.type != "cuda":
            raise RuntimeError("nellie_b200 has no CPU path: %s needs a CUDA device (a host device is only accepted "
                               "together with the test suite's emulated kernel library)" % type(self).__name__)
        self.shape = tuple(int(s) for s in frame_shape)
        self.no_z = bool(no_z)
        self.ndim = 2 if self.no_z else 3
        assert len(self.shape) == self.ndim
        self.nz, self.ny, self.nx = ((1,) + self.shape) if self.no_z else self.shape
        self.n = self.nz * self.ny * self.nx
        f32 = dict(dtype=torch.float32, device=self.device)
        self.frangi_t = torch.empty(self.shape, **f32)
        self.distance_max = torch.empty(self.shape, **f32)
        self.word = torch.zeros(2, dtype=torch.int32, device=self.device)     # [min of the negatives (ordered bits), max half width]
        self.launches = 0

    def _stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        self.launches += 1
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            _cabi.check(rc, name)

    def transform_frangi(self, frangi_f32):
        """hu_tracking.py:604-612."""
        self.word[0] = -1                                  # 0xFFFFFFFF: no negative value seen
        self._call("nb200_hu_frangi_transform", _ptr(frangi_f32), self.n, _ptr(self.frangi_t), _ptr(self.word), self._stream())
        return self.frangi_t

    def max_distance(self, distance_f32):
        """hu_tracking.py:614-616."""
        self._call("nb200_hu_distance_max", _ptr(distance_f32), self.nz, self.ny, self.nx, _ptr(self.distance_max), self._stream())
        return self.distance_max

    def bounds(self, coords_i64):
        """hu_tracking.py:392-421: int32 (N, 6) boxes and the largest half width."""
        coords_i64 = coords_i64.to(torch.int64).contiguous()       # np.argwhere hands out a transposed view
        n = int(coords_i64.shape[0])
        b = torch.empty((n, 6), dtype=torch.int32, device=self.device)
        self.word[1] = 0
        self._call("nb200_hu_bounds", _ptr(coords_i64), n, self.ndim, _ptr(self.distance_max), self.nz, self.ny, self.nx,
                   _ptr(b), C.c_void_p(self.word.data_ptr() + 4), self._stream())
        return b, int(self.word[1].item())

    def roi_stats(self, frame_f32, bounds, cube, int_bits):
        n = int(bounds.shape[0])
        out = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        self._call("nb200_hu_roi_stats", _ptr(frame_f32), self.nz, self.ny, self.nx, _ptr(bounds), n, self.ndim, int(cube),
                   int(int_bits), _ptr(out), self._stream())
        return out

    def log_hu(self, frame_f32, bounds, side, cube, integer_frame):
        n = int(bounds.shape[0])
        n_proj = 1 if self.no_z else 3
        out = torch.empty((n, 6 * n_proj), dtype=torch.float64, device=self.device)
        group = max(1, self.PROJ_BYTES // (4 * n_proj * side * side))
        proj = torch.empty(min(n, group) * n_proj * side * side, dtype=torch.float32, device=self.device)
        for a in range(0, n, group):
            b = bounds[a:a + group]
            self._call("nb200_hu_log_moments", _ptr(frame_f32), self.nz, self.ny, self.nx, _ptr(b), int(b.shape[0]),
                       self.ndim, int(side), int(cube), int(bool(integer_frame)), _ptr(proj), _ptr(out[a:a + group]),
                       self._stream())
        return out

    def frame_features(self, intensity_f32, int_bits, frangi_f32, distance_f32, marker, dense_limit, low_memory=False):
        """hu_tracking.py:585-680: (coords (N, ndim) int64 tensor, stats (N, 4) float32, log-Hu (N, 6 | 18), used_dense)."""
        fr = self.transform_frangi(frangi_f32)
        self.max_distance(distance_f32)
        coords = torch.nonzero(marker > 0)                 # argwhere: raster order
        n = int(coords.shape[0])
        if n == 0:
            return coords, None, None, True
        bounds, max_half = self.bounds(coords.contiguous())
        side = 2 * max_half + 1                            # hu_tracking.py:633 (max_radius)
        use_dense = (n * side ** self.ndim) <= int(dense_limit) and not low_memory      # :636-641
        cube = side if use_dense else 0
        stats = torch.cat([self.roi_stats(intensity_f32, bounds, cube, int_bits),
                           self.roi_stats(fr, bounds, cube, 0)], dim=1)
        hu = self.log_hu(intensity_f32, bounds, side, cube, int_bits != 0)
        if not use_dense:
            hu = hu.to(torch.float32)                      # streaming rows are stored into a float32 matrix (:698, :747)
        return coords, stats, hu, use_dense


class HuMomentFeatures:
    def __init__(self, im_info, num_t=None, max_distance_um=1.0, viewer=None, device="auto", mode="auto",
                 max_dense_pairs=int(1e7), max_dense_roi_voxels_cpu=int(5e7), max_dense_roi_voxels_gpu=int(2e7),
                 low_memory=False, cuda_device=None, dense_limit=None):
        dev = (device or "auto").lower()
        if dev == "cpu":
            raise ValueError("nellie_b200.HuMomentFeatures implements the CUDA path only; device='cpu' belongs to "
                             "nellie.tracking.hu_tracking.HuMomentTracking")
        if dev not in _DEVICES:
            raise ValueError(f"Unsupported device '{device}'. Use 'auto', 'gpu' or 'b200'.")
        self.im_info = im_info
        self.num_t = num_t
        if num_t is None and not im_info.no_t:
            self.num_t = im_info.shape[im_info.axes.index("T")]
        if im_info.no_z:                                   # hu_tracking.py:117-121
            self.scaling = (im_info.dim_res["Y"], im_info.dim_res["X"])
        else:
            self.scaling = (im_info.dim_res["Z"], im_info.dim_res["Y"], im_info.dim_res["X"])
        self.shape = ()
        self.im_memmap = None
        self.im_frangi_memmap = None
        self.im_distance_memmap = None
        self.im_marker_memmap = None
        self.viewer = viewer
        self.device = device
        self.device_type = "cuda"
        self.low_memory = bool(low_memory)
        self.max_dense_roi_voxels_cpu = int(max_dense_roi_voxels_cpu)
        self.max_dense_roi_voxels_gpu = int(max_dense_roi_voxels_gpu)
        # which of the reference's two ROI modes a frame takes decides the bits of `stats` (numpy reduces different arrays):
        # by default the threshold of the reference's GPU backend; dense_limit=max_dense_roi_voxels_cpu reproduces its CPU runs
        self.dense_limit = int(max_dense_roi_voxels_gpu if dense_limit is None else dense_limit)
        self._cuda_device = cuda_device
        self._engine = None
        _cabi.load()

    def _get_t(self):
        if self.num_t is None:
            self.num_t = 1 if self.im_info.no_t else self.im_info.shape[self.im_info.axes.index("T")]

    def _allocate_memory(self):
        """hu_tracking.py:528-542 (inputs only; the flow-vector output belongs to the tracker)."""
        paths = self.im_info.pipeline_paths
        self.label_memmap = self.im_info.get_memmap(paths["im_instance_label"])
        self.im_memmap = self.im_info.get_memmap(self.im_info.im_path)
        self.im_frangi_memmap = self.im_info.get_memmap(paths["im_preprocessed"])
        self.im_marker_memmap = self.im_info.get_memmap(paths["im_marker"])
        self.im_distance_memmap = self.im_info.get_memmap(paths["im_distance"])
        self.shape = self.label_memmap.shape

    def _torch_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested but CUDA is not available. (nellie_b200 has no CPU path)")
        if self._cuda_device is not None:
            return torch.device(self._cuda_device)
        return torch.device("cuda", torch.cuda.current_device())

    def _engine_for(self, frame_shape):
        key = tuple(int(s) for s in frame_shape)
        if self._engine is None or self._engine.shape != key:
            dev = self._torch_device()
            with _on(dev):
                self._engine = HuFeatureEngine(key, self.im_info.no_z, dev)
        return self._engine

    def _dev_f32(self, arr):
        if isinstance(arr, torch.Tensor):
            return arr.to(self._torch_device(), dtype=torch.float32).contiguous()
        a = np.asarray(arr)
        if not a.dtype.isnative:
            a = a.astype(a.dtype.newbyteorder("="))
        # uint8 / uint16 frames travel in their own type and are cast on the device
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self._torch_device())
        return t.to(torch.float32).contiguous()

    def frame_features_device(self, intensity, frangi, distance, marker):
        """Frames (host arrays or tensors) -> (coords, stats, hu) device tensors of the engine (None, None when no marker)."""
        bits = integer_bits(intensity.dtype)
        eng = self._engine_for(tuple(marker.shape))
        with _on(eng.device):
            mk = marker if isinstance(marker, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(marker)))
            coords, stats, hu, _ = eng.frame_features(self._dev_f32(intensity), bits, self._dev_f32(frangi),
                                                      self._dev_f32(distance), mk.to(eng.device), self.dense_limit,
                                                      self.low_memory)
        return coords, stats, hu

    def _get_frame_features_impl(self, t) -> _FrameFeatures:
        """hu_tracking.py:585-680; arrays come back as numpy (the reference keeps stats / hu in its array module)."""
        coords, stats, hu = self.frame_features_device(self.im_memmap[t], self.im_frangi_memmap[t],
                                                       self.im_distance_memmap[t], self.im_marker_memmap[t])
        dims = 2 if self.im_info.no_z else 3
        if coords.shape[0] == 0:                           # hu_tracking.py:620-626
            return _FrameFeatures(np.zeros((0, dims), dtype=int), np.zeros((0, dims), dtype=float),
                                  np.zeros((0, 0), dtype=np.float32), np.zeros((0, 0), dtype=np.float32))
        coords_np = coords.cpu().numpy()
        coords_phys = coords_np * np.asarray(self.scaling, dtype=float)
        return _FrameFeatures(coords_np.astype(int), coords_phys, stats.cpu().numpy(), hu.cpu().numpy())

    def _get_frame_features(self, t) -> _FrameFeatures:
        return self._get_frame_features_impl(t)
