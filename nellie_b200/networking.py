"""The array kernels of ``nellie.segmentation.networking.Network`` that its GPU backend runs:

    _get_pixel_class                 networking.py:669-680   3^d neighbour count of the skeleton (classes 1..4)
    _get_branch_skel_labels          networking.py:758-797   connected components of the non-junction skeleton
    _remove_connected_label_pixels   networking.py:261-296   drop skeleton voxels that touch two objects

Same method names and results as the reference (ids of ``scipy.ndimage.label``); numpy arrays or CUDA tensors in, the same
kind out.  The rest of the stage — ``skimage.morphology.skeletonize`` (Lee-94 thinning), ``_add_missing_skeleton_labels``
and the EDT-based ``_relabel_objects`` — runs on the host in the reference as well and is not implemented here, so this is
not a drop-in ``Network`` class yet: it is the part of it that maps onto the Label kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi


def _ptr(t):
    return C.c_void_p(t.data_ptr())


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
        with torch.cuda.device(self.device):
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
        with torch.cuda.device(self.device):
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
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.nb200_remove_connected_label_pixels(_ptr(d), nz, ny, nx, _ptr(out), self._stream()),
                        "nb200_remove_connected_label_pixels")
        return out.cpu().numpy() if was_np else out
