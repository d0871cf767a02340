"""Multi-GPU execution of the Filter path: one process per GPU over torch.distributed (NCCL).

Two shardings:

* **T** — frames are independent (gamma / thresholds / labels are per frame, filtering.py:1007-1012,
  labelling.py:701-706): frame ``t`` belongs to rank ``t mod G``; no data-path collective.
* **Z** — one frame split into Z slabs (BASELINE config #3).  Rank g owns planes ``[z0, z1)``.  Per sigma:
  one neighbour halo exchange of ``r_z + 2`` planes of the blurred volume (send/recv, both directions),
  then the threshold reductions.  The fast path keeps the histogram state and the Hessian stats in one 267-word
  record and reduces it at five points per sigma with ONE all-gather + one fold kernel each (``fold_state``:
  MIN / MAX of the sample ranges, SUM of count + 256 bins, MAX of every Hessian-stats word); the exact fallback
  path still uses the three separate all-reduces.  Everything is stream-ordered, no host round trip.  ``_mask_volume`` all-gathers the ≤ 1e6 lattice samples and exchanges 2 planes of the
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


def allocate_shared_output(im_info, path, dtype, description, t_shard):
    """Output file of a T-sharded stage: created by rank 0 ONLY, opened by the others.

    ``im_info.allocate_memory`` (verifier.py:992-1070) re-creates the file (``tifffile.imwrite`` / ``open(path, 'wb')``);
    if every rank called it, a rank starting later would truncate frames an earlier rank has already written.  Protocol:
    rank 0 allocates; with an initialised ``torch.distributed`` group all ranks then meet at a barrier and ranks > 0 map
    the existing file; without a process group (sequential use of the shards in one process) ranks > 0 simply require the
    file to exist already."""
    rank, world = (0, 1) if t_shard is None else t_shard
    use_dist = world > 1 and dist.is_available() and dist.is_initialized()
    mm = None
    if rank == 0:
        mm = im_info.allocate_memory(path, dtype=dtype, description=description, return_memmap=True)
        if hasattr(mm, "flush"):
            mm.flush()
    if use_dist:
        dist.barrier()
    if rank != 0:
        import os
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: the output of a T-sharded stage is created by rank 0; run shard (0, {world}) first "
                                    "or initialise torch.distributed so that the ranks can wait for it")
        mm = im_info.get_memmap(path)
    return mm


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

    def fold_state(self, state: torch.Tensor, stage: int):
        """One packed reduction point of the threshold state ("pack into <= 2 calls"): all-gather the
        267-word record [histogram state | Hessian stats] of every rank, then fold the gathered records with the right
        operator per word — ``stage`` 0: HIST_MIN -> min, HIST_MAX -> max; 1: count + bins -> sum; Hessian stats -> max in
        both.  One collective + one kernel instead of the three all-reduces (plus slicing glue) of the first version;
        the result is bit-identical on every rank (integers)."""
        if self.world == 1:
            return
        flat = state.reshape(-1)                      # one record, or one per sigma (batched reduction points): a view
        count = flat.numel() // _cabi.STATE_WORDS
        bufs = self.__dict__.setdefault("_gath_bufs", {})     # one per record count: captured graphs keep their pointers
        key = (flat.numel(), str(flat.device))
        if key not in bufs:
            bufs[key] = torch.empty(self.world * flat.numel(), dtype=flat.dtype, device=flat.device)
        self._gath = bufs[key]
        dist.all_gather_into_tensor(self._gath, flat, group=self.group)
        if flat.is_cuda:
            import ctypes as C
            lib = _cabi.load()
            _cabi.check(lib.nb200_fold_records_n(C.c_void_p(self._gath.data_ptr()), self.world, count, int(stage),
                                                 C.c_void_p(flat.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)),
                        "nb200_fold_records_n")
            return
        # host tensors (gloo tests of the plumbing): the same fold with torch ops
        g = self._gath.view(self.world, count, _cabi.STATE_WORDS)
        rec = flat.view(count, _cabi.STATE_WORDS)
        hw = _cabi.HIST_WORDS
        if stage == _cabi.FOLD_MINMAX:
            rec[:, 0] = g[:, :, 0].min(0).values
            rec[:, 1] = g[:, :, 1].max(0).values
        else:
            rec[:, 2:hw] = g[:, :, 2:hw].sum(0)
        rec[:, hw:] = g[:, :, hw:].max(0).values

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
        e.fold_state = self.comm.fold_state
        # a slab's kernels are short: issuing ~250 launches + collectives per frame from Python costs more than they run
        # for on 8 GPUs; capture them once and replay (NB200_NO_GRAPH=1 keeps eager launches)
        import os
        e.use_graph = self.world > 1 and not os.environ.get("NB200_NO_GRAPH")
        # every reduction point once per frame for all sigmas instead of once per sigma (engine._run_sigmas_batched)
        e.batch_sigmas = self.world > 1 and not os.environ.get("NB200_NO_BATCH")
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
