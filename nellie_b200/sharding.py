"""Multi-GPU execution of the Filter path: one process per GPU over torch.distributed (NCCL).

Two shardings:

* **T** — frames are independent (gamma / thresholds / labels are per frame, filtering.py:1007-1012,
  labelling.py:701-706): frame ``t`` belongs to rank ``t mod G``; no data-path collective.
* **Z** — one frame split into Z slabs (BASELINE config #3).  Rank g owns planes ``[z0, z1)``.  Per sigma:
  one neighbour halo exchange of ``r_z + 2`` planes of the blurred volume (send/recv, both directions),
  then the threshold reductions: MAX over ranks of (−min, max) of the sample range, SUM of count + 256
  bins, MAX of (max|H|, max frob²) — three tiny all-reduces per threshold, stream-ordered, no host round
  trip.  ``_mask_volume`` all-gathers the ≤ 1e6 lattice samples and exchanges 2 planes of the
  accumulator.  Every rank runs the same kernels on the same global lattice with global border rules,
  so the N-GPU output equals the 1-GPU output bit for bit.

:class:`ZComm` holds only torch.distributed calls on tensors, so its logic is exercised on CPU with the
gloo backend (tests/test_sharding_cpu.py); :class:`ZShardedFilter` wires it into the CUDA engine.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

from . import _cabi


def z_partition(nz: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal Z ranges, one per rank (first ``nz % world`` ranks get one extra plane)."""
    base, extra = divmod(int(nz), int(world))
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


def frames_of_rank(num_t: int, rank: int, world: int) -> List[int]:
    """T-sharding: frame t -> rank t mod world."""
    return [t for t in range(int(num_t)) if t % world == rank]


class ZComm:
    """Collectives of a Z-sharded frame.  ``pad_lo``/``pad_hi`` = halo planes present in the local buffers."""

    def __init__(self, nz_own: int, pad_lo: int, pad_hi: int, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nz_own, self.pad_lo, self.pad_hi = int(nz_own), int(pad_lo), int(pad_hi)
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank < self.world - 1 else None

    # -- halo ---------------------------------------------------------------------------------
    def exchange_halo(self, buf: torch.Tensor, depth: int):
        """Fill ``depth`` halo planes on each interior side of ``buf`` (planes along dim 0) with the
        neighbours' owned boundary planes; the owned planes are [pad_lo, pad_lo + nz_own)."""
        depth = int(depth)
        if depth <= 0 or self.world == 1:
            return
        a, b = self.pad_lo, self.pad_lo + self.nz_own
        if depth > self.nz_own:
            raise ValueError(f"halo depth {depth} exceeds the {self.nz_own} planes a rank owns; use fewer GPUs")
        ops = []
        if self.lower is not None:
            if depth > self.pad_lo:
                raise ValueError("low halo buffer too small")
            ops.append(dist.P2POp(dist.isend, buf[a:a + depth], self.lower, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[a - depth:a], self.lower, self.group))
        if self.upper is not None:
            if depth > self.pad_hi:
                raise ValueError("high halo buffer too small")
            ops.append(dist.P2POp(dist.isend, buf[b - depth:b], self.upper, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[b:b + depth], self.upper, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    # -- threshold state (int64[3 + 256]: min key, max key, count, bins) ---------------------------
    def reduce_hist_minmax(self, state: torch.Tensor):
        if self.world == 1:
            return
        mm = torch.stack((-state[0], state[1]))
        dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=self.group)
        state[0] = -mm[0]
        state[1] = mm[1]

    def reduce_hist_bins(self, state: torch.Tensor):
        if self.world == 1:
            return
        dist.all_reduce(state[2:], op=dist.ReduceOp.SUM, group=self.group)   # count + 256 bins

    def reduce_hstats(self, hstats: torch.Tensor):
        if self.world == 1:
            return
        dist.all_reduce(hstats, op=dist.ReduceOp.MAX, group=self.group)      # float bits of non-negative floats

    def gather_samples(self, samples: torch.Tensor, n: int):
        """All ranks' lattice samples concatenated (zero padded: consumers keep values > 0 only)."""
        if self.world == 1:
            return samples, n
        cnt = torch.tensor([int(n)], dtype=torch.int64, device=samples.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=self.group)
        m = max(1, int(cnt.item()))
        mine = torch.zeros(m, dtype=samples.dtype, device=samples.device)
        mine[:n] = samples[:n]
        out = torch.empty(m * self.world, dtype=samples.dtype, device=samples.device)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        return out, m * self.world


class ZShardedFilter:
    """Filter path of ONE frame split over the ranks of the default process group along Z."""

    def __init__(self, shape, params, device, group=None):
        from .engine import FrangiEngine3D
        nz, ny, nx = (int(s) for s in shape)
        self.shape = (nz, ny, nx)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.z0, self.z1 = z_partition(nz, self.world)[self.rank]
        self.device = torch.device(device)
        self.engine = FrangiEngine3D(shape, params, device=self.device, slab=(nz, self.z0, self.z1 - self.z0))
        e = self.engine
        if self.world > 1 and e.nz_own < e.halo_z:
            raise ValueError(f"{e.nz_own} planes per rank < halo {e.halo_z}: too many GPUs for this frame")
        self.comm = ZComm(e.nz_own, e.pad_lo, e.pad_hi, group)
        e.exchange_halo = self.comm.exchange_halo
        e.reduce_hist_minmax = self.comm.reduce_hist_minmax
        e.reduce_hist_bins = self.comm.reduce_hist_bins
        e.reduce_hstats = self.comm.reduce_hstats
        e.gather_samples = self._gather_samples
        self._pinned = None

    def _gather_samples(self, samples, n):
        # the size is fixed by the geometry: no device->host sync on the frame path
        if self.world == 1:
            return samples, n
        if not hasattr(self, "_gbuf"):
            cnt = torch.tensor([int(n)], dtype=torch.int64, device=samples.device)
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=self.comm.group)
            self._gmax = max(1, int(cnt.item()))
            self._gmine = torch.zeros(self._gmax, dtype=samples.dtype, device=samples.device)
            self._gbuf = torch.empty(self._gmax * self.world, dtype=samples.dtype, device=samples.device)
        self._gmine.zero_()
        self._gmine[:n] = samples[:n]
        dist.all_gather_into_tensor(self._gbuf, self._gmine, group=self.comm.group)
        return self._gbuf, self._gmax * self.world

    def slab_of(self, frame: torch.Tensor) -> torch.Tensor:
        return frame[self.z0:self.z1]

    def filter_frame(self, slab: torch.Tensor) -> torch.Tensor:
        """``slab`` = this rank's planes [z0, z1) of the frame (device tensor); returns its output planes."""
        return self.engine.filter_frame(slab)

    def make_phantom_slab(self, seed: int, n_tubes=None) -> torch.Tensor:
        from .phantoms import tubular_phantom
        full = tubular_phantom(self.shape, seed=seed, device=self.device, n_tubes=n_tubes)
        slab = full[self.z0:self.z1].clone()
        del full
        torch.cuda.empty_cache()
        return slab

    def e2e(self, steps: int, slab: torch.Tensor):
        """Whole-job voxels/s through pinned host buffers (H2D + filter + D2H per step), max over ranks."""
        import time
        host_in = torch.empty(slab.shape, dtype=torch.float32).pin_memory()
        host_in.copy_(slab)
        host_out = torch.empty(slab.shape, dtype=torch.float32).pin_memory()
        staging = torch.empty_like(slab)

        def step():
            staging.copy_(host_in, non_blocking=True)
            out = self.engine.filter_frame(staging)
            host_out.copy_(out, non_blocking=True)

        step()
        torch.cuda.synchronize(self.device)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize(self.device)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=self.device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        voxels = float(self.shape[0]) * self.shape[1] * self.shape[2]
        return {"value": voxels * steps / float(dt.item()), "unit": "voxels/s",
                "h2d_bytes_per_step": int(voxels * 4), "d2h_bytes_per_step": int(voxels * 4)}
