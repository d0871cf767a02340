"""Z-sharded Label: ``Label._get_labels`` (nellie/segmentation/labelling.py:467-509) of ONE frame
whose planes are split over the ranks of a process group.

Every rank labels its own slab with the local CCL kernel (``nb200_ccl_label``) and the slabs are stitched by a
**seam merge**: a component's canonical id is the GLOBAL linear index of its first voxel in raster order (+1), ranks
exchange the id planes on both sides of every seam, the (id, id) pairs of voxels that touch across a seam (26-/8- or
6-/4-connectivity) are all-gathered, a small union-find over those ids gives old id → merged id (the minimum of the
set), and every rank remaps its slab.  Canonical ids are therefore identical on all ranks and sorting them reproduces
``scipy.ndimage.label``'s numbering (raster order of first voxels) for the whole frame.  On top of that primitive:

* fill-holes: background components (6-connectivity) that do not touch the GLOBAL frame border are filled;
* size filter: per-id voxel counts are summed over the ranks;
* 3^d majority: one plane of halo from each neighbour, replicate padding at the global border (a one-voxel reflect);
* final labels: rank of the canonical id among all ids of the frame.

This module is plumbing: torch tensors, ``torch.distributed`` collectives and the local CCL callback; seam planes and
id lists are the only data that travel (no volume collective).  ``tests/test_sharded_label_cpu.py`` runs it on CPU
with gloo (world sizes 2 and 3) against ``scipy.ndimage`` on the whole frame, with scipy standing in for the local
CCL kernel; the CUDA callback is :func:`cuda_local_label`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

LocalLabelFn = Callable[[torch.Tensor, bool], Tuple[torch.Tensor, int]]


def cuda_local_label(mask: torch.Tensor, full_conn: bool) -> Tuple[torch.Tensor, int]:
    """Local CCL of a (nz, ny, nx) bool/uint8 CUDA tensor through the C ABI: int32 labels 1..n in local raster order."""
    import ctypes as C

    from . import _cabi
    lib = _cabi.load()
    m = mask.to(torch.uint8).contiguous()
    nz, ny, nx = (int(s) for s in m.shape)
    ws = torch.empty(int(lib.nb200_label_workspace_bytes(nz, ny, nx)), dtype=torch.uint8, device=m.device)
    out = torch.empty(m.shape, dtype=torch.int32, device=m.device)
    n = torch.zeros(1, dtype=torch.int64, device=m.device)
    with torch.cuda.device(m.device):
        _cabi.call("nb200_ccl_label", C.c_void_p(m.data_ptr()), nz, ny, nx, int(bool(full_conn)), C.c_void_p(out.data_ptr()),
                   C.c_void_p(ws.data_ptr()), C.c_void_p(n.data_ptr()),
                   C.c_void_p(torch.cuda.current_stream(m.device).cuda_stream))
    return out, int(n.item())


def cuda_threshold_from_samples(samples: torch.Tensor, log_domain: bool) -> Optional[float]:
    """min(triangle, Otsu) of the gathered sample through the same device kernels as the single-GPU path
    (``LabelEngine.frangi_threshold`` / ``intensity_otsu``: labelling.py:440-465): ``log_domain`` = thresholds of
    log10(sample), returned as 10**t (Frangi); else plain Otsu (intensity).  None for an empty sample.
    Every rank calls it on the same gathered sample and gets the same scalar.  (Composition of verified C-ABI calls;
    not yet run in a multi-GPU job.)"""
    import ctypes as C

    from . import _cabi
    lib = _cabi.load()
    n = int(samples.numel())
    if n == 0:
        return None
    dev = samples.device
    vals = samples.to(torch.float32).contiguous()
    hist = torch.zeros(_cabi.HIST_WORDS, dtype=torch.int64, device=dev)
    thr = torch.zeros(7, dtype=torch.float64, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())                                     # noqa: E731
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        tf = _cabi.TF_LOG10 if log_domain else _cabi.TF_NONE
        _cabi.call("nb200_hist_reset", vp(hist), st)
        _cabi.call("nb200_hist_minmax", vp(vals), n, tf, None, vp(hist), st)
        _cabi.call("nb200_hist_bins", vp(vals), n, tf, None, vp(hist), st)
        _cabi.call("nb200_finalize_label_threshold", vp(hist), int(bool(log_domain)), vp(thr), st)
    out = thr.cpu().numpy()
    if out[3] != 0.0:
        return None
    if out[4] != 0.0:
        raise ValueError("attempt to get argmax of an empty sequence")           # what the reference raises
    if log_domain:                       # labelling.py:452-455 on the scalars themselves (numpy's float32 power)
        return float(min(10 ** np.float32(out[5]), 10 ** np.float32(out[6])))
    return float(np.float32(out[0]))


def _all_gather_ragged(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenation over ranks of 1-D / 2-D int64 tensors of different lengths (dim 0)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    m = max(1, max(int(c.item()) for c in counts))
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:int(c.item())] for b, c in zip(bufs, counts)], dim=0)


def _merge_ids(pairs: np.ndarray):
    """Union-find over the ids in ``pairs`` (k, 2): returns (old ids sorted, merged id = minimum of the set)."""
    if pairs.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    ids = np.unique(pairs)
    idx = {int(v): i for i, v in enumerate(ids)}
    parent = list(range(len(ids)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in pairs:
        ra, rb = find(idx[int(a)]), find(idx[int(b)])
        if ra != rb:
            if ra < rb:
                parent[rb] = ra       # ids is sorted: the smaller index is the smaller id
            else:
                parent[ra] = rb
    return ids.astype(np.int64), np.array([ids[find(i)] for i in range(len(ids))], dtype=np.int64)


def _remap(vol: torch.Tensor, old: torch.Tensor, new: torch.Tensor) -> torch.Tensor:
    if old.numel() == 0:
        return vol
    flat = vol.reshape(-1)
    pos = torch.searchsorted(old, flat).clamp_(max=old.numel() - 1)
    hit = old[pos] == flat
    return torch.where(hit, new[pos], flat).reshape(vol.shape)


def sharded_sample_nonzero(slab: torch.Tensor, z0: int, nz_glob: int, sampling_pixels: int,
                           gate: Optional[torch.Tensor] = None, gate_thresh: Optional[float] = None, group=None) -> torch.Tensor:
    """``Label._sample_nonzero`` (labelling.py:385-438) of a Z-sharded frame: ``flat[off::step]`` of the GLOBAL
    flattened frame, ``step = size // sampling_pixels``, offsets 0 then ``step // 2``, keeping positive values (and
    ``gate > gate_thresh``); full scan when both offsets come up empty.  Every rank takes the lattice points that fall
    into its slab; the kept values are all-gathered in rank order, so every rank holds the same sample, in the same
    order as the single-GPU path, and derives the same thresholds from it."""
    nz, ny, nx = (int(v) for v in slab.shape)
    plane = ny * nx
    size = int(nz_glob) * plane
    step = max(size // max(1, int(sampling_pixels)), 1)
    offsets = (0, step // 2) if step > 1 and step // 2 > 0 else (0,)
    flat = slab.reshape(-1)
    gflat = gate.reshape(-1) if (gate is not None and gate_thresh is not None) else None
    g0 = int(z0) * plane                                 # global flat index of my first voxel

    def gather(vals):
        bits = vals.contiguous().view(torch.int32).to(torch.int64)        # ragged all-gather works on int64
        return _all_gather_ragged(bits, group).to(torch.int32).view(torch.float32)

    for off in offsets:
        first = (off - g0) % step                        # first local index with (g0 + i - off) % step == 0, i >= 0
        if g0 + first < off:
            first += ((off - g0 - first) + step - 1) // step * step
        vals = flat[first::step]
        keep = vals > 0
        if gflat is not None:
            keep &= gflat[first::step] > gate_thresh
        allv = gather(vals[keep])
        if allv.numel() > 0 or step == 1:
            return allv
    keep = flat > 0
    if gflat is not None:
        keep &= gflat > gate_thresh
    return gather(flat[keep])


class ZShardedLabeller:
    """Distributed ``_get_labels`` for the slab ``[z0, z0 + nz_own)`` of a frame with ``nz_glob`` planes."""

    def __init__(self, z0: int, nz_own: int, nz_glob: int, ny: int, nx: int, local_label: LocalLabelFn, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.z0, self.nz, self.nz_glob, self.ny, self.nx = int(z0), int(nz_own), int(nz_glob), int(ny), int(nx)
        self.local_label = local_label
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank < self.world - 1 else None

    # -- one plane to / from each neighbour ---------------------------------------------------------------------
    def _exchange_planes(self, first: torch.Tensor, last: torch.Tensor):
        """Send my first plane down and my last plane up; returns (plane below my slab, plane above) or None."""
        below = torch.zeros_like(first) if self.lower is not None else None
        above = torch.zeros_like(last) if self.upper is not None else None
        ops = []
        if self.lower is not None:
            ops += [dist.P2POp(dist.isend, first.contiguous(), self.lower, self.group),
                    dist.P2POp(dist.irecv, below, self.lower, self.group)]
        if self.upper is not None:
            ops += [dist.P2POp(dist.isend, last.contiguous(), self.upper, self.group),
                    dist.P2POp(dist.irecv, above, self.upper, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return below, above

    # -- distributed connected components ------------------------------------------------------------------------
    def components(self, mask: torch.Tensor, full_conn: bool) -> torch.Tensor:
        """int64 volume: canonical id (global linear index of the component's first voxel, +1) or 0."""
        plane = self.ny * self.nx
        lab, n_local = self.local_label(mask, full_conn)
        lab = lab.to(torch.int64)
        lin = torch.arange(self.nz * plane, dtype=torch.int64, device=mask.device) + self.z0 * plane + 1
        first = torch.full((n_local + 1,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=mask.device)
        first.scatter_reduce_(0, lab.reshape(-1), lin, "amin", include_self=True)
        first[0] = 0
        canon = first[lab]
        if self.world == 1:
            return canon
        # seam pairs: my last plane against the first plane of the slab above
        _, above = self._exchange_planes(canon[0], canon[-1])
        pairs = torch.zeros((0, 2), dtype=torch.int64, device=mask.device)
        if above is not None:
            mine = canon[-1]
            offs = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)] if full_conn else [(0, 0)]
            found = []
            q = F.pad(above[None, None].to(torch.float64), (1, 1, 1, 1))[0, 0].to(torch.int64)   # zero border
            for dy, dx in offs:
                nb = q[1 + dy:1 + dy + self.ny, 1 + dx:1 + dx + self.nx]
                both = (mine > 0) & (nb > 0)
                if bool(both.any()):
                    found.append(torch.stack((mine[both], nb[both]), dim=1))
            if found:
                pairs = torch.unique(torch.cat(found, dim=0), dim=0)
        all_pairs = _all_gather_ragged(pairs, self.group)
        old, new = _merge_ids(all_pairs.cpu().numpy())
        return _remap(canon, torch.from_numpy(old).to(mask.device), torch.from_numpy(new).to(mask.device))

    # -- labelling.py:467-509 ----------------------------------------------------------------------------------------
    def label(self, mask: torch.Tensor, min_area: int, fill_holes: bool = True) -> torch.Tensor:
        """``mask``: bool slab of the thresholded frame.  Returns int32 labels of the slab, numbered as
        ``scipy.ndimage.label`` numbers the whole frame."""
        dev = mask.device
        mask = mask.to(torch.bool)
        if mask.dim() != 3 or self.nz_glob < 2:
            raise ValueError("ZShardedLabeller shards 3-D frames along Z; 2-D frames are T-sharded")
        if fill_holes:
            bg = self.components(~mask, full_conn=False)
            edge = torch.zeros_like(mask)
            edge[:, 0, :] = True
            edge[:, -1, :] = True
            edge[:, :, 0] = True
            edge[:, :, -1] = True
            if self.z0 == 0:
                edge[0] = True
            if self.z0 + self.nz == self.nz_glob:
                edge[-1] = True
            touching = torch.unique(bg[edge & (bg > 0)])
            touching = torch.unique(_all_gather_ragged(touching, self.group))
            outside = torch.isin(bg, touching)
            mask = mask | ((bg > 0) & ~outside)
        comp = self.components(mask, full_conn=True)
        ids, counts = torch.unique(comp[comp > 0], return_counts=True)
        table = _all_gather_ragged(torch.stack((ids, counts), dim=1), self.group)
        if table.shape[0]:
            uid, inv = torch.unique(table[:, 0], return_inverse=True)
            area = torch.zeros(uid.shape[0], dtype=torch.int64, device=dev).index_add_(0, inv, table[:, 1])
            keep_ids = uid[area >= int(min_area)]
        else:
            keep_ids = torch.zeros(0, dtype=torch.int64, device=dev)
        keep = torch.isin(comp, keep_ids) & (comp > 0)
        # uniform_filter(float32(mask), 3, mode="reflect") > 0.5  ==  at least 14 (3-D) / 5 (2-D) of the window set
        k = keep.to(torch.float32)
        below, above = self._exchange_planes(k[0], k[-1])
        lo = below if below is not None else k[0]               # global border: reflect of one voxel = the plane itself
        hi = above if above is not None else k[-1]
        vol = torch.cat((lo[None], k, hi[None]), dim=0)
        vol = F.pad(vol[None, None], (1, 1, 1, 1, 0, 0), mode="replicate")
        cnt = F.conv3d(vol, torch.ones((1, 1, 3, 3, 3), dtype=torch.float32, device=dev))[0, 0]
        smooth = cnt >= 13.5
        comp2 = self.components(smooth, full_conn=True)
        all_ids = torch.unique(_all_gather_ragged(torch.unique(comp2[comp2 > 0]), self.group))
        labels = torch.zeros(comp2.shape, dtype=torch.int32, device=dev)
        if all_ids.numel():
            pos = torch.searchsorted(all_ids, comp2.reshape(-1)).clamp_(max=all_ids.numel() - 1).reshape(comp2.shape)
            labels = torch.where(comp2 > 0, (pos + 1).to(torch.int32), labels)
        return labels


def label_frame_z_sharded(frangi_slab: torch.Tensor, z0: int, nz_glob: int, min_area: int, sampling_pixels: int = 1_000_000,
                          raw_slab: Optional[torch.Tensor] = None, otsu_thresh_intensity: bool = False,
                          threshold: Optional[float] = None, group=None) -> torch.Tensor:
    """``_compute_frame_thresholds`` + ``_run_frame_full_volume`` (labelling.py:511-556) for one Z slab on CUDA: the
    thresholds come from the all-gathered strided sample, the labels from :class:`ZShardedLabeller`.  Returns the int32
    labels of the slab, numbered as the single-GPU path numbers the whole frame."""
    nz, ny, nx = (int(v) for v in frangi_slab.shape)
    it = None
    if otsu_thresh_intensity:
        it = cuda_threshold_from_samples(sharded_sample_nonzero(raw_slab.to(torch.float32), z0, nz_glob, sampling_pixels,
                                                                 group=group), log_domain=False) or 0
    elif threshold is not None:
        it = threshold
    gate = raw_slab.to(torch.float32) if it is not None else None
    ft = cuda_threshold_from_samples(sharded_sample_nonzero(frangi_slab, z0, nz_glob, sampling_pixels, gate, it, group=group),
                                     log_domain=True)
    frangi = frangi_slab if it is None else frangi_slab * (gate > np.float32(it))           # labelling.py:550-552
    mask = torch.zeros_like(frangi, dtype=torch.bool) if ft is None else frangi > np.float32(ft)
    return ZShardedLabeller(z0, nz, nz_glob, ny, nx, cuda_local_label, group).label(mask, min_area)
