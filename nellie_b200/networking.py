"""The array kernels of ``nellie.segmentation.networking.Network`` that its GPU backend runs:

    _get_pixel_class                 networking.py:669-680   3^d neighbour count of the skeleton (classes 1..4)
    _get_branch_skel_labels          networking.py:758-797   connected components of the non-junction skeleton
    _remove_connected_label_pixels   networking.py:261-296   drop skeleton voxels that touch two objects

Same method names and results as the reference (ids of ``scipy.ndimage.label``); numpy arrays or CUDA tensors in, the same
kind out.

``Network`` (below) is the stage class: it adds the two steps the reference keeps on the host even in its GPU backend —
``_add_missing_skeleton_labels`` (networking.py:315-392) and the per-object nearest-seed relabel ``_relabel_objects``
(:485-577, scipy's feature transform restated with its tie-breaking, csrc/network.cu) — as device kernels, and the frame
loop / ``run()`` of the reference (:802-977).  The one step that stays a host dependency, exactly as in the reference's
GPU backend, is ``skimage.morphology.skeletonize`` (:394-410, Lee-94 thinning): ``Network._skeletonize`` calls
scikit-image when it is importable (or the ``skeletonize=`` callable given to the constructor) and raises otherwise.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import logging

import numpy as np
import torch

from . import _cabi


logger = logging.getLogger("nellie_b200")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _on(device):
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


class NetworkKernels:
    def __init__(self, im_info, cuda_device=None):
        self.im_info = im_info
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested but CUDA is not available. (nellie_b200 has no CPU path)")
        self.device = torch.device(cuda_device if cuda_device is not None else "cuda")
        self._ws = None

    def _dims(self, shape):
        if self.im_info.no_z:
            if len(shape) != 2:
                raise ValueError("2-D stack: frames are (Y, X)")
            return 1, int(shape[0]), int(shape[1])
        if len(shape) != 3:
            raise ValueError("3-D stack: frames are (Z, Y, X)")
        return tuple(int(s) for s in shape)

    def _dev_i32(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device, torch.int32).contiguous(), False
        return torch.from_numpy(np.ascontiguousarray(a).astype(np.int32)).to(self.device), True

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _get_pixel_class(self, skel):
        """networking.py:669-680: uint8 classes — 0 background, 1 isolated, 2 tip, 3 edge, 4 junction."""
        nz, ny, nx = self._dims(skel.shape)
        d, was_np = self._dev_i32(skel)
        out = torch.empty(d.shape, dtype=torch.uint8, device=self.device)
        with _on(self.device):
            _cabi.check(self.lib.nb200_pixel_class(_ptr(d), nz, ny, nx, _ptr(out), self._stream()), "nb200_pixel_class")
        return out.cpu().numpy() if was_np else out

    def _get_branch_skel_labels(self, pixel_class):
        """networking.py:758-797: int32 labels of the connected components of (class > 0) & (class != 4)."""
        nz, ny, nx = self._dims(pixel_class.shape)
        was_np = not isinstance(pixel_class, torch.Tensor)
        pc = (torch.from_numpy(np.ascontiguousarray(pixel_class).astype(np.uint8)) if was_np else pixel_class).to(
            self.device, torch.uint8).contiguous()
        need = int(self.lib.nb200_label_workspace_bytes(nz, ny, nx))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        labels = torch.empty(pc.shape, dtype=torch.int32, device=self.device)
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        with _on(self.device):
            _cabi.check(self.lib.nb200_branch_labels(_ptr(pc), nz, ny, nx, _ptr(labels), _ptr(self._ws), _ptr(n),
                                                     self._stream()), "nb200_branch_labels")
        return labels.cpu().numpy() if was_np else labels

    def _remove_connected_label_pixels(self, skel_labels):
        """networking.py:261-296: labelled voxels off the frame boundary whose 3^d window holds two different positive
        labels are set to 0."""
        nz, ny, nx = self._dims(skel_labels.shape)
        d, was_np = self._dev_i32(skel_labels)
        if not self.im_info.no_z and nz == 1:
            # a one-plane 3-D stack: every voxel lies on the Z boundary, which the reference never modifies
            # (networking.py:285-291); the C entry point reads nz == 1 as a 2-D frame, so this case stays here
            return d.cpu().numpy() if was_np else d.clone()
        out = torch.empty_like(d)
        with _on(self.device):
            _cabi.check(self.lib.nb200_remove_connected_label_pixels(_ptr(d), nz, ny, nx, _ptr(out), self._stream()),
                        "nb200_remove_connected_label_pixels")
        return out.cpu().numpy() if was_np else out


_INT_MAX = 2 ** 31 - 1


class NetworkEngine:
    """Device steps of one Network frame on top of the C ABI.  ``lib`` / ``device``: the product passes the CUDA library and
    a CUDA device; the CPU tests inject host builds of the same kernels (oracle/network_host.cpp) with ``device='cpu'``."""

    def __init__(self, no_z, scaling, device, lib=None):
        self.lib = _cabi.load() if lib is None else lib
        self.device = torch.device(device)
        if lib is None and self.device.type != "cuda":
            raise RuntimeError("nellie_b200 has no CPU path: %s needs a CUDA device (a host device is only accepted "
                               "together with the test suite's emulated kernel library)" % type(self).__name__)
        self.no_z = bool(no_z)
        sc = [float(v) for v in scaling]
        self.sampling = (C.c_double * 3)(*( [1.0] + sc if self.no_z else sc ))
        self.crop_voxels = 0          # of the last relabel (diagnostics)
        self.max_crop_voxels = 1 << 28   # per group of objects: 7.5 GB of workspace
        self._ws = None

    def _stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dims(self, t):
        return (1,) + tuple(t.shape) if self.no_z else tuple(t.shape)

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            _cabi.check(rc, name)

    def add_missing(self, skel, labels, frangi, max_label):
        """networking.py:315-392, in place on ``skel`` (int32 device tensor)."""
        if max_label <= 0:
            return skel
        nz, ny, nx = self._dims(labels)
        key = torch.zeros(max_label + 1, dtype=torch.int64, device=self.device)
        in_skel = torch.zeros(max_label + 1, dtype=torch.uint8, device=self.device)
        self._call("nb200_network_add_missing", _ptr(labels), _ptr(frangi), _ptr(skel), nz, ny, nx, int(max_label),
                   _ptr(key), _ptr(in_skel), self._stream())
        return skel

    def skeleton_labels(self, skel, labels):
        """networking.py:835: (skel > 0) * labels."""
        out = torch.empty_like(labels)
        self._call("nb200_network_skeleton_labels", _ptr(skel), _ptr(labels), labels.numel(), _ptr(out), self._stream())
        return out

    def relabel(self, branch, labels, max_label):
        """networking.py:485-577: uint32 branch label of every object voxel (as an int32 tensor holding the same bits)."""
        out = torch.zeros(labels.shape, dtype=torch.int32, device=self.device)
        self.crop_voxels = 0
        if max_label <= 0:
            return out
        nz, ny, nx = self._dims(labels)
        boxes = torch.empty((max_label + 1, 6), dtype=torch.int32, device=self.device)
        boxes[:, :3] = _INT_MAX
        boxes[:, 3:] = -1
        seeded = torch.zeros(max_label + 1, dtype=torch.uint8, device=self.device)
        self._call("nb200_network_object_boxes", _ptr(labels), _ptr(branch), nz, ny, nx, int(max_label), _ptr(boxes),
                   _ptr(seeded), self._stream())
        # crop table (plumbing on a few thousand rows): objects that exist and hold a seed, ascending label
        keep = torch.nonzero((seeded != 0) & (boxes[:, 3] >= 0)).flatten()
        m = int(keep.numel())
        if m == 0:
            return out
        b = boxes[keep].to(torch.int64)
        ext = b[:, 3:] - b[:, :3] + 1
        vol = ext[:, 0] * ext[:, 1] * ext[:, 2]
        # the crops of all objects of a group lie end to end in crop space; groups keep the workspace (28 B per crop voxel)
        # bounded when long diagonal objects have boxes as large as the frame (the reference handles one object at a time)
        vol_host = vol.cpu().tolist()
        groups, start, acc = [], 0, 0
        for i, v in enumerate(vol_host):
            if i > start and acc + v > self.max_crop_voxels:
                groups.append((start, i))
                start, acc = i, 0
            acc += v
        groups.append((start, m))
        for g0, g1 in groups:
            gvol, gext = vol[g0:g1], ext[g0:g1]
            off = torch.cumsum(gvol, 0) - gvol
            crops = torch.cat([keep[g0:g1, None], b[g0:g1, :3], gext, off[:, None]], dim=1).contiguous()
            lines = torch.stack([gvol // gext[:, a] for a in range(3)])              # (3, objects of the group)
            line_starts = (torch.cumsum(lines, 1) - lines).contiguous()
            totals = lines.sum(1).cpu()
            V = int(sum(vol_host[g0:g1]))
            self.crop_voxels += V
            n_lines = (C.c_longlong * 3)(*[int(v) for v in totals])
            if self._ws is None or self._ws.numel() < 7 * V:
                self._ws = None                                                      # release before growing
                self._ws = torch.empty(7 * V, dtype=torch.int32, device=self.device)
            ft_a, ft_b, stack = self._ws[:3 * V], self._ws[3 * V:6 * V], self._ws[6 * V:7 * V]
            self._call("nb200_network_relabel", _ptr(labels), _ptr(branch), nz, ny, nx, _ptr(crops), g1 - g0, V,
                       _ptr(line_starts), n_lines, self.sampling, _ptr(ft_a), _ptr(ft_b), _ptr(stack), _ptr(out),
                       self._stream())
        return out


class Network(NetworkKernels):
    """Drop-in for ``nellie.segmentation.networking.Network`` (networking.py:22-977): same constructor keywords, helper
    names, outputs (``im_skel`` int32, ``im_pixel_class`` uint8, ``im_skel_relabelled`` uint32) and ``run()`` contract."""

    def __init__(self, im_info, num_t=None, min_radius_um=0.20, max_radius_um=1, viewer=None, device="auto",
                 low_memory: bool = False, max_chunk_voxels: int = int(1e6), cuda_device=None, t_shard=None, fallback=None,
                 skeletonize=None):
        dev = (device or "auto").lower()
        if dev == "cpu":
            raise ValueError("nellie_b200.Network implements the CUDA path only; device='cpu' belongs to "
                             "nellie.segmentation.networking.Network")
        if dev not in ("auto", "gpu", "cuda", "b200"):
            raise ValueError(f"Unsupported device '{device}'. Use 'auto', 'gpu' or 'b200'.")
        self.im_info = im_info
        self.lib = None
        self.device_name = device
        self.device_type = "cuda"
        self._cuda_device = cuda_device
        self._ws = None
        self.low_memory = bool(low_memory)
        self.max_chunk_voxels = int(max_chunk_voxels)
        self.num_t = num_t
        if num_t is None and not im_info.no_t:
            self.num_t = im_info.shape[im_info.axes.index("T")]
        if not im_info.no_z:                                 # networking.py:72-85
            self.z_ratio = im_info.dim_res["Z"] / im_info.dim_res["X"]
            self.scaling = (im_info.dim_res["Z"], im_info.dim_res["Y"], im_info.dim_res["X"])
        else:
            self.scaling = (im_info.dim_res["Y"], im_info.dim_res["X"])
        self.min_radius_um = max(min_radius_um, im_info.dim_res["X"])
        self.max_radius_um = max_radius_um
        self.min_radius_px = self.min_radius_um / im_info.dim_res["X"]
        self.max_radius_px = self.max_radius_um / im_info.dim_res["X"]
        self.shape = ()
        self.im_memmap = None
        self.im_frangi_memmap = None
        self.label_memmap = None
        self.pixel_class_memmap = None
        self.skel_memmap = None
        self.skel_relabelled_memmap = None
        self.viewer = viewer
        self.t_shard = None if t_shard is None else (int(t_shard[0]), int(t_shard[1]))
        self.fallback = fallback
        self._skeletonize_fn = skeletonize
        self._net = None
        self._ctor_kwargs = dict(num_t=num_t, min_radius_um=min_radius_um, max_radius_um=max_radius_um, viewer=viewer,
                                 max_chunk_voxels=max_chunk_voxels)
        if low_memory:
            logger.warning("nellie_b200.Network: low_memory is accepted for compatibility and ignored (the reference's "
                           "chunked helpers compute the same arrays)")
        if fallback != "reference":
            self.lib = _cabi.load()

    # ---- device plumbing ---------------------------------------------------------------------------------------------
    @property
    def device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested but CUDA is not available. (nellie_b200 has no CPU path)")
        if self._cuda_device is not None:
            return torch.device(self._cuda_device)
        return torch.device("cuda", torch.cuda.current_device())

    @device.setter
    def device(self, value):
        self._cuda_device = value

    def _engine(self):
        if self._net is None:
            self._net = NetworkEngine(self.im_info.no_z, self.scaling, self.device)
        return self._net

    def _get_t(self):
        if self.num_t is None:
            self.num_t = 1 if self.im_info.no_t else self.im_info.shape[self.im_info.axes.index("T")]

    def _allocate_memory(self):
        """networking.py:721-753; a T-sharded stage creates its output files on rank 0 only."""
        from .sharding import allocate_shared_output
        paths = self.im_info.pipeline_paths
        self.label_memmap = self.im_info.get_memmap(paths["im_instance_label"])
        self.im_memmap = self.im_info.get_memmap(self.im_info.im_path)
        self.im_frangi_memmap = self.im_info.get_memmap(paths["im_preprocessed"])
        self.shape = self.label_memmap.shape
        self.skel_memmap = allocate_shared_output(self.im_info, paths["im_skel"], "int32", "skeleton image", self.t_shard)
        self.pixel_class_memmap = allocate_shared_output(self.im_info, paths["im_pixel_class"], "uint8",
                                                         "pixel class image", self.t_shard)
        self.skel_relabelled_memmap = allocate_shared_output(self.im_info, paths["im_skel_relabelled"], "uint32",
                                                             "skeleton relabelled image", self.t_shard)

    # ---- stage steps ---------------------------------------------------------------------------------------------------
    def _skeletonize(self, label_frame):
        """networking.py:394-410: labels * skeletonize(labels > 0) — host thinning, as in the reference's GPU backend."""
        labels = np.asarray(label_frame)
        fn = self._skeletonize_fn
        if fn is None:
            try:
                from skimage import morphology as morph
            except ImportError as exc:
                raise RuntimeError("Network needs scikit-image for skimage.morphology.skeletonize (a host step of the "
                                   "reference as well), or a skeletonize= callable") from exc
            fn = morph.skeletonize
        return labels * np.asarray(fn(labels > 0)).astype(bool)

    def _add_missing_skeleton_labels(self, skel_frame, label_frame, frangi_frame):
        """networking.py:315-392; numpy in -> numpy out."""
        eng = self._engine()
        with _on(eng.device):
            labels, _ = self._dev_i32(label_frame)
            skel = self._dev_i32(skel_frame)[0].clone()
            frangi = torch.from_numpy(np.ascontiguousarray(np.asarray(frangi_frame, dtype=np.float32))).to(eng.device)
            return eng.add_missing(skel, labels, frangi, int(labels.max().item()) if labels.numel() else 0).cpu().numpy()

    def _relabel_objects(self, branch_skel_labels, label_frame):
        """networking.py:485-577; returns a uint32 numpy array."""
        eng = self._engine()
        with _on(eng.device):
            labels, _ = self._dev_i32(label_frame)
            branch, _ = self._dev_i32(branch_skel_labels)
            out = eng.relabel(branch, labels, int(labels.max().item()) if labels.numel() else 0)
            return out.cpu().numpy().view(np.uint32)

    def _run_frame(self, t):
        """networking.py:802-851: (branch_skel_labels int32, pixel_class uint8, branch_labels uint32) of frame ``t``."""
        logger.info("Running network analysis, volume %s/%s", t, (self.num_t or 1) - 1)
        label_np = np.asarray(self.label_memmap[t])
        skel_np = self._skeletonize(label_np)
        eng = self._engine()
        with _on(eng.device):
            labels, _ = self._dev_i32(label_np)
            frangi = torch.from_numpy(np.ascontiguousarray(np.asarray(self.im_frangi_memmap[t], dtype=np.float32))).to(eng.device)
            max_label = int(labels.max().item()) if labels.numel() else 0
            skel = self._remove_connected_label_pixels(self._dev_i32(skel_np)[0])
            skel = eng.add_missing(skel.clone(), labels, frangi, max_label)
            skel_pre = eng.skeleton_labels(skel, labels)
            pixel_class = self._get_pixel_class(skel_pre)
            branch = self._get_branch_skel_labels(pixel_class)
            relabelled = eng.relabel(branch, labels, max_label)
            return branch.cpu().numpy(), pixel_class.cpu().numpy(), relabelled.cpu().numpy().view(np.uint32)

    def _run_networking(self):
        """T loop of networking.py:905-936."""
        from .sharding import frames_of_rank
        frames = range(self.num_t) if self.t_shard is None else frames_of_rank(self.num_t, *self.t_shard)
        for t in frames:
            if self.viewer is not None:
                self.viewer.status = f"Extracting branches. Frame: {t + 1} of {self.num_t}."
            skel, pixel_class, relabelled = self._run_frame(t)
            single = self.im_info.no_t or self.num_t == 1                      # networking.py:913
            for mm, frame in ((self.skel_memmap, skel), (self.pixel_class_memmap, pixel_class),
                              (self.skel_relabelled_memmap, relabelled)):
                if single:
                    mm[:] = frame
                else:
                    mm[t] = frame
                if hasattr(mm, "flush"):
                    mm.flush()

    def _run_b200(self):
        _ = self.device
        self.lib = _cabi.load()
        self._get_t()
        self._allocate_memory()
        self._run_networking()

    def _run_reference(self, device, low_memory):
        from nellie.segmentation.networking import Network as ReferenceNetwork
        ReferenceNetwork(self.im_info, device=device, low_memory=low_memory, **self._ctor_kwargs).run()

    def run(self):
        logger.info("Running network analysis (nellie_b200).")
        from .adaptive import run_with_ladder
        run_with_ladder("Network", self._run_b200, self._run_reference, self.fallback)
