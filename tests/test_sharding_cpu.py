"""Host-side logic of the multi-GPU path, exercised on CPU with the gloo backend (world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nz, halo, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nellie_b200.sharding import ZComm, z_partition
        parts = z_partition(nz, world)
        z0, z1 = parts[rank]
        pad_lo, pad_hi = min(halo, z0), min(halo, nz - z1)
        ny, nx = 5, 7
        full = torch.arange(nz * ny * nx, dtype=torch.float32).reshape(nz, ny, nx)     # the "global frame"
        buf = torch.full((pad_lo + (z1 - z0) + pad_hi, ny, nx), -1.0)
        buf[pad_lo:pad_lo + z1 - z0] = full[z0:z1]
        comm = ZComm(z1 - z0, pad_lo, pad_hi)
        for depth in (1, halo):
            buf2 = buf.clone()
            comm.exchange_halo(buf2, depth)
            lo = depth if rank > 0 else 0
            hi = depth if rank < world - 1 else 0
            want = full[z0 - lo:z1 + hi]
            got = buf2[pad_lo - lo:pad_lo + (z1 - z0) + hi]
            assert torch.equal(got, want), (rank, depth)
        # threshold state: min key / max key / count / bins
        state = torch.zeros(3 + 256, dtype=torch.int64)
        state[0] = 1000 + 10 * rank if rank != 1 else 0xFFFFFFFF      # rank 1 has no samples
        state[1] = 5000 + rank if rank != 1 else 0
        state[2] = 3 * (rank + 1) if rank != 1 else 0
        state[3:] = rank + 1 if rank != 1 else 0
        comm.reduce_hist_minmax(state)
        comm.reduce_hist_bins(state)
        contributing = [r for r in range(world) if r != 1]
        assert int(state[0]) == min(1000 + 10 * r for r in contributing)
        assert int(state[1]) == max(5000 + r for r in contributing)
        assert int(state[2]) == sum(3 * (r + 1) for r in contributing)
        assert bool((state[3:] == sum(r + 1 for r in contributing)).all())
        hs = torch.tensor([np.float32(1.5 + rank).view(np.uint32), np.float32(9.0 - rank).view(np.uint32)], dtype=torch.int64)
        comm.reduce_hstats(hs)
        assert int(hs[0]) == int(np.float32(1.5 + world - 1).view(np.uint32))
        assert int(hs[1]) == int(np.float32(9.0).view(np.uint32))
        # packed record [histogram state | Hessian stats]: one all-gather + fold per reduction point (fast path)
        from nellie_b200 import _cabi
        rec = torch.zeros(_cabi.STATE_WORDS, dtype=torch.int64)
        hw = _cabi.HIST_WORDS
        rec[0] = 1000 + 10 * rank if rank != 1 else 0xFFFFFFFF
        rec[1] = 5000 + rank if rank != 1 else 0
        rec[2] = 3 * (rank + 1) if rank != 1 else 0
        rec[3:hw] = rank + 1 if rank != 1 else 0
        rec[hw:] = torch.arange(_cabi.HS_WORDS) * 7 + rank
        before = rec.clone()
        comm.fold_state(rec, _cabi.FOLD_MINMAX)
        assert int(rec[0]) == min(1000 + 10 * r for r in contributing) and int(rec[1]) == max(5000 + r for r in contributing)
        assert torch.equal(rec[2:hw], before[2:hw]), "stage 0 must not touch count / bins"
        assert torch.equal(rec[hw:], torch.arange(_cabi.HS_WORDS) * 7 + world - 1)
        comm.fold_state(rec, _cabi.FOLD_BINS)
        assert int(rec[2]) == sum(3 * (r + 1) for r in contributing)
        assert bool((rec[3:hw] == sum(r + 1 for r in contributing)).all())
        assert int(rec[0]) == min(1000 + 10 * r for r in contributing), "stage 1 must not touch min / max"
        # one reduction point for several sigmas at once: `count` records folded record by record
        recs = torch.zeros((3, _cabi.STATE_WORDS), dtype=torch.int64)
        for k in range(3):
            recs[k, 0] = 100 * k + 10 * rank
            recs[k, 1] = 100 * k + rank
            recs[k, 2:hw] = (k + 1) * (rank + 1)
            recs[k, hw:] = 1000 * k + rank
        comm.fold_state(recs, _cabi.FOLD_MINMAX)
        for k in range(3):
            assert int(recs[k, 0]) == 100 * k and int(recs[k, 1]) == 100 * k + world - 1
            assert bool((recs[k, 2:hw] == (k + 1) * (rank + 1)).all()), "stage 0 must not touch count / bins"
            assert bool((recs[k, hw:] == 1000 * k + world - 1).all())
        comm.fold_state(recs, _cabi.FOLD_BINS)
        for k in range(3):
            assert bool((recs[k, 2:hw] == (k + 1) * sum(r + 1 for r in range(world))).all())
        # lattice samples: ragged lengths, zero padded
        n = 4 + rank
        s = torch.arange(1, n + 1, dtype=torch.float32) + 100 * rank
        allv, total = comm.gather_samples(s, n)
        pos = allv[:total][allv[:total] > 0]
        want = torch.cat([torch.arange(1, 5 + r, dtype=torch.float32) + 100 * r for r in range(world)])
        assert torch.equal(torch.sort(pos).values, torch.sort(want).values)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz,halo", [(2, 16, 3), (3, 17, 4)])
def test_zcomm_collectives_gloo(world, nz, halo, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), nz, halo, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_partitions():
    from nellie_b200.sharding import frames_of_rank, z_partition
    assert z_partition(1024, 8) == [(128 * i, 128 * (i + 1)) for i in range(8)]
    p = z_partition(17, 3)
    assert p == [(0, 6), (6, 12), (12, 17)] and p[-1][1] == 17
    frames = [frames_of_rank(16, r, 4) for r in range(4)]
    assert sorted(sum(frames, [])) == list(range(16)) and frames[1] == [1, 5, 9, 13]
