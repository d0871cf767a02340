"""B200-native drop-in for ``nellie.segmentation.filtering.Filter`` (reference: filtering.py:17-1076).

Same constructor keywords, same ``.run()`` contract, same on-disk intermediate
(``pipeline_paths['im_preprocessed']``, float32, written per frame through the ``im_info`` memmaps),
same private helper names the reference's callers and tests touch (``_get_t``,
``_set_default_sigmas``, ``_run_frame``, ``_mask_volume``, ``_run_filter``).  The arithmetic runs in
hand-written sm_100a kernels behind the C ABI (``include/nellie_b200.h``); there is no numpy/cupy
fallback — if the CUDA library or a GPU is missing, construction or ``run()`` raises.

Deliberate deviations from the reference:
  * the input frame is never mutated (the reference overwrites float32 inputs in place, C-1);
  * the broken low-memory / chunked branches are not reproduced: ``low_memory`` is accepted and
    ignored, results always equal the reference's full-volume branch (C-4);
  * ``sigmas=`` lets a caller pin the scale list (BASELINE config #3 needs 6 sigmas).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _cabi
from .engine import FilterParams, FrangiEngine3D, sample_strides

logger = logging.getLogger("nellie_b200")

_DEVICES = ("auto", "gpu", "cuda", "b200")


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("GPU backend requested but CUDA is not available. "
                           "(nellie_b200 has no CPU path; use nellie's own Filter on CPU hosts)")


class Filter:
    def __init__(self, im_info, num_t=None, remove_edges: bool = False, min_radius_um: float = 0.25,
                 max_radius_um: float = 1.0, alpha_sq: float = 0.5, beta_sq: float = 0.5, frob_thresh=None,
                 frob_thresh_division=2, viewer=None, device: str = "auto", low_memory: bool = False,
                 max_chunk_voxels: int = int(1e6), max_threshold_samples: int = int(1e6), sigmas=None,
                 cuda_device=None, t_shard=None, fallback=None):
        dev = (device or "auto").lower()
        if dev == "cpu":
            raise ValueError("nellie_b200.Filter implements the CUDA path only; device='cpu' belongs to "
                             "nellie.segmentation.filtering.Filter")
        if dev not in _DEVICES:
            raise ValueError(f"Unsupported device '{device}'. Use 'auto', 'gpu' or 'b200'.")
        self.im_info = im_info
        self.device = device
        self.device_type = "cuda"
        self.truncate = 3.0
        if not im_info.no_z:
            z_res = im_info.dim_res.get("Z") or im_info.dim_res.get("X") or 1.0
            x_res = im_info.dim_res.get("X") or 1.0
            self.z_ratio = float(z_res) / float(x_res)
        self.num_t = num_t
        if num_t is None and not im_info.no_t:
            self.num_t = im_info.shape[im_info.axes.index("T")]
        self.remove_edges = remove_edges
        self.min_radius_um = min_radius_um
        self.max_radius_um = max_radius_um
        self.min_radius_px = self.min_radius_um / im_info.dim_res["X"]
        self.max_radius_px = self.max_radius_um / im_info.dim_res["X"]
        self.im_memmap = None
        self.frangi_memmap = None
        self.sigma_vec = None
        self.sigmas = None
        self.alpha_sq = float(alpha_sq)
        self.beta_sq = float(beta_sq)
        self.frob_thresh = frob_thresh
        self.frob_thresh_division = frob_thresh_division
        self.viewer = viewer
        self.low_memory = low_memory
        self.max_chunk_voxels = int(max_chunk_voxels)
        self.max_threshold_samples = int(max_threshold_samples)
        self.work_dtype = "float32"
        self.out_dtype = "float32"
        self.halo = None
        self._explicit_sigmas = None if sigmas is None else [float(s) for s in sigmas]
        self._cuda_device = cuda_device
        # T-sharding: (rank, world) -> this object handles frames t with t % world == rank; frames are
        # independent (per-frame gamma / thresholds, filtering.py:1007-1012), so ranks share nothing but the output file
        self.t_shard = None if t_shard is None else (int(t_shard[0]), int(t_shard[1]))
        self._engine = None
        self._engine_key = None
        # retry ladder of run() (adaptive.py): None = the B200 path or an exception; "reference" = continue with the
        # reference's own CPU classes on an OOM / GPU-unavailable failure, like filtering.py:1054-1076
        self.fallback = fallback
        self._ctor_kwargs = dict(num_t=num_t, remove_edges=remove_edges, min_radius_um=min_radius_um,
                                 max_radius_um=max_radius_um, alpha_sq=alpha_sq, beta_sq=beta_sq, frob_thresh=frob_thresh,
                                 frob_thresh_division=frob_thresh_division, viewer=viewer,
                                 max_chunk_voxels=max_chunk_voxels, max_threshold_samples=max_threshold_samples)
        if low_memory:
            logger.warning("nellie_b200.Filter: low_memory is accepted for compatibility and ignored (results always "
                           "equal the reference's full-volume branch)")
        if fallback != "reference":
            _cabi.load()  # fail at construction when the CUDA library is absent

    # ---- parameters ---------------------------------------------------------------------------
    def _params(self) -> FilterParams:
        return FilterParams(dim_res=dict(self.im_info.dim_res), no_z=bool(self.im_info.no_z),
                            min_radius_um=self.min_radius_um, max_radius_um=self.max_radius_um,
                            alpha_sq=self.alpha_sq, beta_sq=self.beta_sq, frob_thresh=self.frob_thresh,
                            frob_thresh_division=self.frob_thresh_division,
                            max_threshold_samples=self.max_threshold_samples, truncate=self.truncate,
                            remove_edges=self.remove_edges,
                            sigmas=self.sigmas if self.sigmas is not None else self._explicit_sigmas)

    def _get_t(self):
        if self.num_t is None:
            self.num_t = 1 if self.im_info.no_t else self.im_info.shape[self.im_info.axes.index("T")]

    def _get_sigma_vec(self, sigma):
        self.sigma_vec = self._params().sigma_vec(sigma)
        return self.sigma_vec

    def _set_default_sigmas(self):
        p = self._params()
        p.sigmas = self._explicit_sigmas
        self.sigmas = p.sigma_list()
        self.sigma_min, self.sigma_max = min(self.sigmas), max(self.sigmas)
        self.halo = self._compute_halo()

    def _compute_halo(self):
        if not self.sigmas:
            return None
        return tuple(int(np.ceil(self.truncate * float(s))) for s in self._get_sigma_vec(max(self.sigmas)))

    def _sample_strides(self, shape, max_samples):
        return sample_strides(shape, max_samples)

    def _allocate_memory(self):
        self.im_memmap = self.im_info.get_memmap(self.im_info.im_path)
        self.shape = self.im_memmap.shape
        from .sharding import allocate_shared_output
        self.frangi_memmap = allocate_shared_output(self.im_info, self.im_info.pipeline_paths["im_preprocessed"],
                                                    self.out_dtype, "frangi filtered im", self.t_shard)

    # ---- device plumbing ------------------------------------------------------------------------
    def _torch_device(self):
        _require_cuda()
        if self._cuda_device is not None:
            return torch.device(self._cuda_device)
        return torch.device("cuda", torch.cuda.current_device())

    def _engine_for(self, frame_shape):
        if self.sigmas is None:
            self._set_default_sigmas()
        key = (tuple(frame_shape), tuple(self.sigmas), self.alpha_sq, self.beta_sq, self.frob_thresh,
               self.frob_thresh_division, self.max_threshold_samples, self.remove_edges)
        if self._engine is None or self._engine_key != key:
            dev = self._torch_device()
            with torch.cuda.device(dev):
                if self.im_info.no_z:
                    from .engine2d import FrangiEngine2D
                    self._engine = FrangiEngine2D(frame_shape, self._params(), device=dev)
                else:
                    self._engine = FrangiEngine3D(frame_shape, self._params(), device=dev)
            self._engine_key = key
        return self._engine

    def _to_device(self, frame_cpu):
        """Host frame (memmap slice / ndarray, any real dtype) -> device tensor, via pinned staging."""
        arr = np.ascontiguousarray(frame_cpu)
        if arr.dtype.byteorder == ">" or (arr.dtype.byteorder == "=" and not np.little_endian):
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        if arr.dtype == np.uint32 or arr.dtype == np.uint64:
            arr = arr.astype(np.float32)
        elif arr.dtype == np.float64:
            arr = arr.astype(np.float32)   # xp.asarray(frame, dtype=float32) rounds once, same here
        t = torch.from_numpy(arr)
        return t.pin_memory().to(self._torch_device(), non_blocking=True)

    # ---- per-frame API (reference: filtering.py:910-967) ----------------------------------------
    def _run_frame(self, t, mask=True):
        """Vesselness of frame ``t`` BEFORE _mask_volume, as a host float32 array."""
        logger.info("Running Frangi filter on t=%s.", t)
        frame_cpu = self.im_memmap[t, ...]
        eng = self._engine_for(frame_cpu.shape)
        eng.p.mask = bool(mask)          # mask=False: every voxel passes the Frobenius gate (filtering.py:563-566)
        with torch.cuda.device(eng.device):
            out = eng.filter_frame(self._to_device(frame_cpu), apply_mask_volume=False)
            return out.cpu().numpy()

    def _mask_volume(self, frangi_frame):
        """Percentile threshold + binary opening (filtering.py:952-967) of a host or device frame."""
        is_np = isinstance(frangi_frame, np.ndarray)
        eng = self._engine_for(tuple(frangi_frame.shape))
        with torch.cuda.device(eng.device):
            dev_frame = self._to_device(frangi_frame) if is_np else frangi_frame
            out = eng.mask_volume(dev_frame)
            return out.cpu().numpy() if is_np else out.clone()

    def filter_frame_device(self, frame: torch.Tensor) -> torch.Tensor:
        """Device-resident fast path: one frame in, final ``im_preprocessed`` frame out (engine buffer)."""
        eng = self._engine_for(tuple(frame.shape))
        eng.p.mask = True
        with torch.cuda.device(eng.device):
            return eng.filter_frame(frame, apply_mask_volume=True)

    def filter_frame_host(self, frame_cpu, mask=True) -> np.ndarray:
        """Host array in, host array out: H2D, the whole per-frame path, D2H."""
        eng = self._engine_for(tuple(frame_cpu.shape))
        eng.p.mask = bool(mask)
        with torch.cuda.device(eng.device):
            out = eng.filter_frame(self._to_device(frame_cpu), apply_mask_volume=True)
            return out.cpu().numpy()

    # ---- top level (reference: filtering.py:1005-1076) -------------------------------------------
    def _run_filter(self, mask=True):
        """T loop of filtering.py:1005-1031 as a pipelined frame stream (pipeline.py): the upload of frame
        t+1 and the download + memmap write of frame t-1 overlap the kernels of frame t."""
        from .pipeline import FramePipeline
        single = bool(self.im_info.no_t) or self.num_t == 1
        frame_shape = tuple(self.im_memmap.shape[1:])          # the reference indexes im_memmap[t, ...] (T always present)
        eng = self._engine_for(frame_shape)
        eng.p.mask = bool(mask)

        def get_in(t):
            return self.im_memmap[t, ...]

        def get_out(t):
            if self.frangi_memmap.ndim == len(frame_shape):
                return self.frangi_memmap
            return self.frangi_memmap[0 if single else t]

        def on_frame(t):
            if self.viewer is not None:
                self.viewer.status = f"Preprocessing. Frame: {t + 1} of {self.num_t}."

        def after_store(t):
            if single and self.frangi_memmap.ndim > len(frame_shape) and self.frangi_memmap.shape[0] > 1:
                # filtering.py:1026-1027: ``frangi_memmap[:] = filtered_im[:]`` broadcasts the one frame over every T
                self.frangi_memmap[1:] = self.frangi_memmap[0]
            if hasattr(self.frangi_memmap, "flush"):
                self.frangi_memmap.flush()

        from .sharding import frames_of_rank
        frames = list(range(self.num_t)) if self.t_shard is None else frames_of_rank(self.num_t, *self.t_shard)
        with torch.cuda.device(eng.device):
            FramePipeline(eng).run(len(frames), lambda k: get_in(frames[k]), lambda k: get_out(frames[k]),
                                   apply_mask_volume=True, on_frame=lambda k: on_frame(frames[k]),
                                   after_store=lambda k: after_store(frames[k]))

    def _run_b200(self, mask=True):
        _require_cuda()
        _cabi.load()
        self._get_t()
        self._allocate_memory()
        self._set_default_sigmas()
        self._run_filter(mask=mask)

    def _run_reference(self, device, low_memory, mask=True):
        """A rung of the ladder below the B200 one: the reference's own class, untouched (only with fallback='reference')."""
        from nellie.segmentation.filtering import Filter as ReferenceFilter
        ref = ReferenceFilter(self.im_info, device=device, low_memory=low_memory, **self._ctor_kwargs)
        ref.run(mask=mask)

    def run(self, mask=True):
        logger.info("Running Frangi filter (nellie_b200).")
        from .adaptive import run_with_ladder
        run_with_ladder("Filter", lambda: self._run_b200(mask), lambda dev, low: self._run_reference(dev, low, mask),
                        self.fallback)
