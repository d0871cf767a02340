"""B200-native drop-in for ``nellie.segmentation.labelling.Label`` (reference: labelling.py:17-778).

Same constructor keywords and ``.run()`` contract; reads ``im_preprocessed`` (float32) and the raw
image through ``im_info`` memmaps and writes ``im_instance_label`` (int32, ids restart at 1 every
frame, 0 = background).  Per frame (labelling.py:701-734): log-domain triangle/Otsu threshold of a
strided sample, then threshold -> fill holes (3-D) -> 26/8-connected components -> size filter ->
3^d majority smoothing -> components again, all in CUDA kernels behind ``include/nellie_b200.h``.

Only the reference's full-volume branch is reproduced.  Its Z-chunked low-memory branch
(labelling.py:585-691) merges labels across the seam with 6-connectivity and runs fill-holes / the
size filter per chunk, i.e. it computes a different segmentation;
``chunk_z`` / ``low_memory`` are accepted for signature compatibility and ignored.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np
import torch

from . import _cabi

logger = logging.getLogger("nellie_b200")

_DEVICES = ("auto", "gpu", "cuda", "b200")
_UNSET = object()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)



def gate_threshold_f32(thresh, data_dtype) -> float:
    """The float32 value T with ``float32(x) > T``  <=>  ``x > thresh`` as the reference evaluates it (labelling.py:419,
    :550): numpy compares an integer frame, or any frame against a float64 *numpy* scalar, in float64 — exact — which for
    operands representable in float32 (uint8 / uint16 / float32 frames) is the comparison against the largest float32
    not above ``thresh``; a float32 frame against a Python float (the ``threshold=`` argument) or a float32 scalar (its
    own Otsu threshold) is compared in float32, i.e. against ``thresh`` rounded to nearest."""
    dt = np.dtype(data_dtype) if not isinstance(data_dtype, torch.dtype) else np.dtype(str(data_dtype).replace("torch.", ""))
    exact = dt.kind in "iu" or (isinstance(thresh, np.floating) and thresh.dtype == np.float64 and dt == np.float32)
    t32 = np.float32(thresh)
    if exact and float(t32) > float(thresh):
        t32 = np.nextafter(t32, np.float32(-np.inf), dtype=np.float32)
    return float(t32)


class LabelEngine:
    """Device buffers + kernel sequence for frames of one shape on one GPU."""

    def __init__(self, frame_shape, no_z, min_area, sampling_pixels, device, nbins=256):
        self.lib = _cabi.load()
        self.nbins = int(nbins)          # labelling.py:23-35 histogram_nbins; 256 = the production kernels of thresholds.cu
        self.device = torch.device(device)
        self.shape = tuple(int(s) for s in frame_shape)
        if no_z:
            assert len(self.shape) == 2
            self.nz, (self.ny, self.nx) = 1, self.shape
        else:
            assert len(self.shape) == 3
            self.nz, self.ny, self.nx = self.shape
        self.no_z = bool(no_z)
        self.n = self.nz * self.ny * self.nx
        self.min_area = int(min_area)
        self.step = max(self.n // max(1, int(sampling_pixels)), 1)   # labelling.py:406-407
        dev = self.device
        ws_bytes = self.lib.nb200_label_workspace_bytes(self.nz, self.ny, self.nx)
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self.labels = torch.empty(self.shape, dtype=torch.int32, device=dev)
        self.n_labels = torch.zeros(1, dtype=torch.int64, device=dev)
        n_samp = (self.n + self.step - 1) // self.step
        self.samples = torch.empty(max(1, n_samp), dtype=torch.float32, device=dev)
        self.hist = torch.zeros(_cabi.HIST_WORDS, dtype=torch.int64, device=dev)
        self.thr = torch.zeros(7, dtype=torch.float64, device=dev)
        self.launches = 0
        self._hist_vals = None           # (sample buffer, count) of the last _hist_of call: the general-bin path re-reads it
        self._histn_ws = None
        if self.nbins != 256:
            self._histn_ws = torch.empty(int(self.lib.nb200_histn_workspace_bytes(self.nbins)), dtype=torch.uint8, device=dev)

    def _call(self, name, *args):
        self.launches += 1
        _cabi.check(getattr(self.lib, name)(*args), name)

    def _hist_of(self, vals, n, transform, f64=False):
        st = _stream()
        self._call("nb200_hist_reset", _ptr(self.hist), st)
        self._call("nb200_hist_minmax", _ptr(vals), n, transform, None, _ptr(self.hist), st)
        self._hist_vals = (vals, n)
        if self.nbins != 256:
            return   # bins + thresholds in one call later (nb200_histn_threshold)
        if f64:      # integer frame: numpy bins with float64 edges
            self._call("nb200_hist_bins_f64", _ptr(vals), n, _ptr(self.hist), st)
        else:
            self._call("nb200_hist_bins", _ptr(vals), n, transform, None, _ptr(self.hist), st)

    def _sample_hist(self, frame, transform, gate=None, gate_thresh=0.0, f64=False):
        """labelling.py:385-438 (_sample_nonzero) + the histogram of the kept values. Returns count."""
        st = _stream()
        offsets = (0, self.step // 2) if self.step > 1 and self.step // 2 > 0 else (0,)
        for off in offsets:
            n_out = (self.n - off + self.step - 1) // self.step if off < self.n else 0
            self._call("nb200_strided_sample", _ptr(frame), self.n, off, self.step, _ptr(gate),
                       float(gate_thresh), _ptr(self.samples), st)
            self._hist_of(self.samples, n_out, transform, f64)
            count = int(self.hist[2].item())
            if count > 0 or self.step == 1:
                return count
        # full-scan fallback (labelling.py:425-438): every positive (gated) voxel
        full = frame.reshape(-1)
        if gate is not None:
            tmp = torch.empty_like(full)
            self._call("nb200_strided_sample", _ptr(full), self.n, 0, 1, _ptr(gate), float(gate_thresh), _ptr(tmp), st)
            full = tmp
        self._hist_of(full, self.n, transform, f64)
        return int(self.hist[2].item())

    def _histn(self, log_domain, f64, otsu_only):
        """Bins + thresholds of the last sampled buffer for histogram_nbins != 256 (csrc/histn.cu)."""
        vals, n = self._hist_vals
        self._call("nb200_histn_threshold", _ptr(vals), n, int(log_domain), int(f64), int(otsu_only), self.nbins,
                   _ptr(self.hist), _ptr(self._histn_ws), _ptr(self.thr), _stream())

    def frangi_threshold(self, frangi, gate=None, gate_thresh=None):
        """labelling.py:440-455; returns the float32 threshold as a Python float, or None."""
        self._sample_hist(frangi, _cabi.TF_LOG10, gate if gate_thresh is not None else None,
                          0.0 if gate_thresh is None else gate_thresh)
        if self.nbins == 256:
            self._call("nb200_finalize_label_threshold", _ptr(self.hist), 1, _ptr(self.thr), _stream())
        else:
            self._histn(log_domain=1, f64=0, otsu_only=0)
        out = self.thr.cpu().numpy()
        if out[3] != 0.0:
            return None
        if out[4] != 0.0:
            raise ValueError("attempt to get argmax of an empty sequence")  # what the reference raises
        # labelling.py:452-455 on the two scalars themselves: ``10 ** np.float32`` is numpy's scalar float32 power (libm
        # powf); doing it here reproduces its bits on any host, which a device pow could only approximate
        triangle = 10 ** np.float32(out[5])
        otsu = 10 ** np.float32(out[6])
        return float(min(triangle, otsu))

    def intensity_otsu(self, raw_f32, integer_frame=False):
        """labelling.py:457-465.  float32 frames: float32 edges and centre, returned as np.float32 (what numpy yields);
        integer frames (``integer_frame``; the values must be exact in float32: uint8 / uint16): float64 edges and centre,
        returned as np.float64 — the scalar type decides how the gate compares (``gate_threshold_f32``)."""
        count = self._sample_hist(raw_f32, _cabi.TF_NONE, f64=integer_frame)
        if count == 0:
            return None
        if self.nbins != 256:
            self._histn(log_domain=0, f64=int(bool(integer_frame)), otsu_only=1)
            v = self.thr.cpu().numpy()[0]
            return np.float64(v) if integer_frame else np.float32(v)
        if integer_frame:
            self._call("nb200_finalize_otsu_f64", _ptr(self.hist), _ptr(self.thr), _stream())
            return np.float64(self.thr.cpu().numpy()[0])
        self._call("nb200_finalize_label_threshold", _ptr(self.hist), 0, _ptr(self.thr), _stream())
        return np.float32(self.thr.cpu().numpy()[0])

    def label(self, frangi, frangi_thresh, raw=None, intensity_thresh=None):
        """labelling.py:467-509 + :546-556. Returns the engine's int32 label tensor (device)."""
        thr = torch.tensor([0.0 if frangi_thresh is None else float(np.float32(frangi_thresh)), 0, 0,
                            1.0 if frangi_thresh is None else 0.0, 0], dtype=torch.float64, device=self.device)
        use_int = intensity_thresh is not None
        self._call("nb200_label_frame", _ptr(frangi), _ptr(raw) if use_int else None, int(use_int),
                   float(np.float32(intensity_thresh)) if use_int else 0.0, _ptr(thr), self.nz, self.ny, self.nx,
                   self.min_area, int(not self.no_z), _ptr(self.labels), _ptr(self.workspace),
                   _ptr(self.n_labels), _stream())
        self._thr_keepalive = thr
        return self.labels


class Label:
    def __init__(self, im_info, num_t=None, threshold=None, otsu_thresh_intensity=False, viewer=None,
                 chunk_z=None, flush_interval=1, min_radius_um=0.25, threshold_sampling_pixels=1_000_000,
                 histogram_nbins=256, device="auto", low_memory: bool = False, max_chunk_voxels: int = int(1e6),
                 cuda_device=None, t_shard=None, fallback=None, z_shard=None):
        dev = (device or "auto").lower()
        if dev == "cpu":
            raise ValueError("nellie_b200.Label implements the CUDA path only; device='cpu' belongs to "
                             "nellie.segmentation.labelling.Label")
        if dev not in _DEVICES:
            raise ValueError(f"Unsupported device '{device}'. Use 'auto', 'gpu' or 'b200'.")
        if int(histogram_nbins) < 2:
            raise ValueError("histogram_nbins must be at least 2")
        if int(histogram_nbins) != 256 and z_shard is not None:
            raise NotImplementedError("the Z-sharded Label reduces the 256-bin histogram record; histogram_nbins != 256 is "
                                      "supported for whole frames and T-sharded runs")
        self.im_info = im_info
        self.device = device
        self.device_type = "cuda"
        self.num_t = num_t
        if num_t is None and not im_info.no_t:
            self.num_t = im_info.shape[im_info.axes.index("T")]
        self.threshold = threshold
        self.otsu_thresh_intensity = otsu_thresh_intensity
        self.im_memmap = None
        self.frangi_memmap = None
        self.instance_label_memmap = None
        self.shape = ()
        self.debug = {}
        self.viewer = viewer
        self.chunk_z = None
        self.flush_interval = max(1, int(flush_interval))
        x_res = im_info.dim_res.get("X") or 1.0
        self.min_radius_um = max(float(min_radius_um), float(x_res))     # labelling.py:95-97
        self.threshold_sampling_pixels = int(threshold_sampling_pixels)
        self.histogram_nbins = int(histogram_nbins)
        self.low_memory = bool(low_memory)
        self.max_chunk_voxels = int(max_chunk_voxels)
        self.ndim = 2 if im_info.no_z else 3
        self.min_area_pixels = self._compute_min_area_pixels()
        self._cuda_device = cuda_device
        # T-sharding: (rank, world) -> frames t with t % world == rank (label ids restart per frame, labelling.py:701-706)
        self.t_shard = None if t_shard is None else (int(t_shard[0]), int(t_shard[1]))
        # Z-sharding of every frame over the ranks of the default torch.distributed group: (rank, world);
        # rank r labels planes z_partition(nz, world)[r] of each frame with sharded_label.label_frame_z_sharded (local
        # CUDA CCL + seam merge) and writes its slab of the shared output file; ids equal the single-GPU numbering
        self.z_shard = None if z_shard is None else (int(z_shard[0]), int(z_shard[1]))
        if self.z_shard is not None and t_shard is not None:
            raise ValueError("a Label stage is sharded over T or over Z, not both")
        self._engine = None
        self.fallback = fallback       # retry ladder of run(): see adaptive.py / Filter
        self._ctor_kwargs = dict(num_t=num_t, threshold=threshold, otsu_thresh_intensity=otsu_thresh_intensity, viewer=viewer,
                                 chunk_z=chunk_z, flush_interval=flush_interval, min_radius_um=min_radius_um,
                                 threshold_sampling_pixels=threshold_sampling_pixels, histogram_nbins=histogram_nbins,
                                 max_chunk_voxels=max_chunk_voxels)
        if low_memory or chunk_z is not None:
            logger.warning("nellie_b200.Label: low_memory / chunk_z are accepted for compatibility and ignored (results "
                           "always equal the reference's full-volume branch, labelling.py:538-583)")
        if fallback != "reference":
            _cabi.load()

    # ---- host scalars ---------------------------------------------------------------------------
    def _compute_min_area_pixels(self):
        """labelling.py:209-219."""
        x_res = self.im_info.dim_res.get("X") or 1.0
        y_res = self.im_info.dim_res.get("Y") or x_res
        if self.im_info.no_z:
            area_px = np.pi * (self.min_radius_um ** 2) / (float(x_res) * float(y_res))
            return max(1, int(np.ceil(area_px)))
        z_res = self.im_info.dim_res.get("Z") or x_res
        volume_px = (4.0 / 3.0) * np.pi * (self.min_radius_um ** 3) / (float(x_res) * float(y_res) * float(z_res))
        return max(1, int(np.ceil(volume_px)))

    def _get_t(self):
        if self.num_t is None:
            self.num_t = 1 if self.im_info.no_t else self.im_info.shape[self.im_info.axes.index("T")]

    def _allocate_memory(self):
        self.im_memmap = self.im_info.get_memmap(self.im_info.im_path)
        self.frangi_memmap = self.im_info.get_memmap(self.im_info.pipeline_paths["im_preprocessed"])
        self.shape = self.frangi_memmap.shape
        from .sharding import allocate_shared_output
        self.instance_label_memmap = allocate_shared_output(self.im_info, self.im_info.pipeline_paths["im_instance_label"],
                                                            "int32", "instance segmentation", self.t_shard or self.z_shard)

    # ---- device plumbing --------------------------------------------------------------------------
    def _torch_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested but CUDA is not available. (nellie_b200 has no CPU path)")
        if self._cuda_device is not None:
            return torch.device(self._cuda_device)
        return torch.device("cuda", torch.cuda.current_device())

    def _engine_for(self, frame_shape):
        key = tuple(frame_shape)
        if self._engine is None or self._engine.shape != key or self._engine.min_area != self.min_area_pixels:
            dev = self._torch_device()
            with torch.cuda.device(dev):
                self._engine = LabelEngine(key, self.im_info.no_z, self.min_area_pixels,
                                           self.threshold_sampling_pixels, dev, nbins=self.histogram_nbins)
        return self._engine

    def _dev_f32(self, arr):
        if isinstance(arr, torch.Tensor):
            return arr.to(self._torch_device(), dtype=torch.float32).contiguous()
        a = np.ascontiguousarray(arr)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        if a.dtype in (np.uint32, np.uint64, np.float64):
            a = a.astype(np.float32)
        t = torch.from_numpy(a).pin_memory().to(self._torch_device(), non_blocking=True)
        return t if t.dtype == torch.float32 else t.to(torch.float32)

    # ---- thresholds (reference: labelling.py:440-465, :511-532) -------------------------------------
    def _compute_frangi_threshold(self, frame, mask_frame=None, mask_thresh=None):
        eng = self._engine_for(frame.shape)
        with torch.cuda.device(eng.device):
            gate = self._dev_f32(mask_frame) if (mask_frame is not None and mask_thresh is not None) else None
            if gate is not None:
                mask_thresh = gate_threshold_f32(mask_thresh, mask_frame.dtype)
            return eng.frangi_threshold(self._dev_f32(frame), gate, mask_thresh)

    def _compute_intensity_otsu_threshold(self, frame):
        eng = self._engine_for(frame.shape)
        with torch.cuda.device(eng.device):
            integer = (not frame.dtype.is_floating_point) if isinstance(frame, torch.Tensor) else np.dtype(frame.dtype).kind in "iu"
            return eng.intensity_otsu(self._dev_f32(frame), integer_frame=integer)

    def _compute_frame_thresholds(self, original_view, frangi_view):
        intensity_thresh = None
        if self.otsu_thresh_intensity:
            intensity_thresh = self._compute_intensity_otsu_threshold(original_view)
            if intensity_thresh is None:
                intensity_thresh = 0
        elif self.threshold is not None:
            intensity_thresh = self.threshold
        if intensity_thresh is not None:
            frangi_thresh = self._compute_frangi_threshold(frangi_view, mask_frame=original_view,
                                                           mask_thresh=intensity_thresh)
        else:
            frangi_thresh = self._compute_frangi_threshold(frangi_view)
        return intensity_thresh, frangi_thresh

    # ---- labelling (reference: labelling.py:467-509, :538-583) --------------------------------------
    def _get_labels(self, frame, frangi_thresh=_UNSET):
        if frangi_thresh is _UNSET:
            frangi_thresh = self._compute_frangi_threshold(frame)
        eng = self._engine_for(frame.shape)
        with torch.cuda.device(eng.device):
            labels = eng.label(self._dev_f32(frame), frangi_thresh).cpu().numpy()
        return labels > 0, labels

    def _run_frame_full_volume(self, t, original_view, frangi_view, intensity_thresh, frangi_thresh):
        logger.info("Running semantic segmentation, volume %s/%s", t, (self.num_t or 1) - 1)
        eng = self._engine_for(frangi_view.shape)
        with torch.cuda.device(eng.device):
            raw = self._dev_f32(original_view) if intensity_thresh is not None else None
            if intensity_thresh is not None:
                intensity_thresh = gate_threshold_f32(intensity_thresh, original_view.dtype)
            return eng.label(self._dev_f32(frangi_view), frangi_thresh, raw, intensity_thresh).cpu().numpy()

    def label_frame_device(self, frangi: torch.Tensor, raw: torch.Tensor = None):
        """Device-resident fast path: thresholds + labels for one frame; returns (labels, frangi_thresh)."""
        eng = self._engine_for(tuple(frangi.shape))
        with torch.cuda.device(eng.device):
            it = None
            rawf = None
            if self.otsu_thresh_intensity or self.threshold is not None:
                rawf = raw.to(torch.float32)
                integer = not raw.dtype.is_floating_point
                it = (eng.intensity_otsu(rawf, integer_frame=integer) or 0) if self.otsu_thresh_intensity else self.threshold
                it = gate_threshold_f32(it, raw.dtype)
            ft = eng.frangi_threshold(frangi, rawf, it)
            return eng.label(frangi, ft, rawf, it), ft

    # ---- top level (reference: labelling.py:697-778) -------------------------------------------------
    def _stage(self, view, slot):
        """Host frame (memmap slice / ndarray) -> device float32 through a persistent pinned buffer (one per slot)."""
        arr = np.asarray(view)
        if not arr.dtype.isnative:
            arr = arr.astype(arr.dtype.newbyteorder("="))
        if arr.dtype in (np.uint32, np.uint64, np.float64):
            arr = arr.astype(np.float32)          # xp.asarray(frame, dtype=float32) rounds once; same value here
        src = torch.from_numpy(np.ascontiguousarray(arr))
        bufs = self.__dict__.setdefault("_pinned", {})
        buf = bufs.get(slot)
        if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
            buf = bufs[slot] = torch.empty(src.shape, dtype=src.dtype).pin_memory()
        from .pipeline import parallel_copyto
        parallel_copyto(buf.numpy(), src.numpy())
        t = buf.to(self._torch_device(), non_blocking=True)
        return t if t.dtype == torch.float32 else t.to(torch.float32)

    def _run_segmentation(self):
        """T loop of labelling.py:697-734: every frame is uploaded once (the thresholds and the labelling share the
        device copies), labelled on the device, downloaded through a pinned buffer and written to the memmap."""
        need_raw = bool(self.otsu_thresh_intensity) or self.threshold is not None
        if self.z_shard is not None:
            return self._run_segmentation_z_sharded(need_raw)
        from .sharding import frames_of_rank
        frames = range(self.num_t) if self.t_shard is None else frames_of_rank(self.num_t, *self.t_shard)
        for t in frames:
            if self.viewer is not None:
                self.viewer.status = f"Extracting organelles. Frame: {t + 1} of {self.num_t}."
            original_view = self.im_memmap[t, ...]
            frangi_view = self.frangi_memmap[t, ...]
            logger.info("Running semantic segmentation, volume %s/%s", t, (self.num_t or 1) - 1)
            eng = self._engine_for(frangi_view.shape)
            with torch.cuda.device(eng.device):
                frangi = self._stage(frangi_view, "frangi")
                raw = self._stage(original_view, "raw") if need_raw else None
                labels, _ = self.label_frame_device(frangi, raw)      # host syncs inside (threshold scalars)
                bufs = self.__dict__.setdefault("_pinned", {})
                host = bufs.get("labels")
                if host is None or host.shape != labels.shape:
                    host = bufs["labels"] = torch.empty(labels.shape, dtype=torch.int32).pin_memory()
                host.copy_(labels)
            from .pipeline import parallel_copyto
            parallel_copyto(self.instance_label_memmap[t, ...], host.numpy())
            if (t + 1) % self.flush_interval == 0 and hasattr(self.instance_label_memmap, "flush"):
                self.instance_label_memmap.flush()
        if hasattr(self.instance_label_memmap, "flush"):
            self.instance_label_memmap.flush()

    def _run_segmentation_z_sharded(self, need_raw):
        """T loop with every frame split into Z slabs over the ranks (needs an initialised process group)."""
        from .pipeline import parallel_copyto
        from .sharded_label import label_frame_z_sharded
        from .sharding import z_partition
        if self.im_info.no_z:
            raise ValueError("2-D frames are T-sharded; Z-sharding needs a Z axis")
        rank, world = self.z_shard
        nz = int(self.frangi_memmap.shape[1])
        z0, z1 = z_partition(nz, world)[rank]
        dev = self._torch_device()
        for t in range(self.num_t):
            if self.viewer is not None:
                self.viewer.status = f"Extracting organelles. Frame: {t + 1} of {self.num_t}."
            with torch.cuda.device(dev):
                frangi = self._stage(self.frangi_memmap[t, z0:z1], "frangi")
                raw = self._stage(self.im_memmap[t, z0:z1], "raw") if need_raw else None
                labels = label_frame_z_sharded(frangi, z0, nz, self.min_area_pixels, self.threshold_sampling_pixels, raw,
                                               bool(self.otsu_thresh_intensity), self.threshold)
                host = labels.cpu()
            parallel_copyto(self.instance_label_memmap[t, z0:z1], host.numpy())
            if hasattr(self.instance_label_memmap, "flush"):
                self.instance_label_memmap.flush()

    def _run_b200(self):
        self._torch_device()
        _cabi.load()
        self._get_t()
        self._allocate_memory()
        self._run_segmentation()

    def _run_reference(self, device, low_memory):
        from nellie.segmentation.labelling import Label as ReferenceLabel
        ReferenceLabel(self.im_info, device=device, low_memory=low_memory, **self._ctor_kwargs).run()

    def run(self):
        logger.info("Running semantic segmentation (nellie_b200).")
        from .adaptive import run_with_ladder
        run_with_ladder("Label", self._run_b200, self._run_reference, self.fallback)
