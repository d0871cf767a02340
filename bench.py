"""Benchmark of the Filter hot path (Frangi + eigen) — BASELINE.json metric: voxels/s on a 1024^3 fp32
synthetic volume, 6 sigmas, at 1/2/4/8 B200 (Z-sharded with halo exchange), plus achieved HBM GB/s
of the fused Hessian+eigen kernel (K3) against the measured copy peak.

    python bench.py --gpus 1 --steps 3 --warmup 3            # our arm (one JSON line on stdout)
    python bench.py --impl reference --steps 1 --warmup 0    # CPU arm: the oracle port on host cores

A "step" is one pass of the whole per-frame path (filtering.py:1007-1031: cascaded blur, per-sigma
thresholds, fused Hessian/eig/vesselness, percentile + opening) over the resident volume.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "voxels/s Frangi+eig (Filter) on 1024^3 fp32, 6 sigmas"
SIGMAS_CFG3 = [1.0, 1.4, 1.8, 2.2, 2.6, 3.0]
DIM_RES_CFG3 = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy/scipy restatement of the reference) on host cores
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    size, seed = args
    from nellie_b200.phantoms import tubular_phantom_np
    from oracle import pipeline as P
    raw = tubular_phantom_np((size,) * 3, seed=seed)
    spec = P.FrameSpec(dim_res=DIM_RES_CFG3, no_z=False, sigmas=SIGMAS_CFG3)
    t0 = time.perf_counter()
    out = P.filter_frame(raw, spec)
    return time.perf_counter() - t0, int(raw.size), float(out.max())


def _cpu_warm(_):
    from nellie_b200.phantoms import tubular_phantom_np  # noqa: F401  (imports paid before the clock starts)
    from oracle import pipeline  # noqa: F401
    return os.getpid()


def cpu_baseline(size=112, procs=1, pool=None):
    """Time the oracle on `procs` independent size^3 crops of the workload (one process per crop: the
    reference path is single-threaded, frames/crops are its only parallel axis).  Process start-up and imports
    are outside the timed region (`pool` = a warmed multiprocessing pool)."""
    if procs == 1:
        try:                                   # `cores: 1` must be true: keep BLAS/OpenMP pools at one thread
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1):
                res = [_cpu_worker((size, 1000))]
        except ImportError:
            res = [_cpu_worker((size, 1000))]
        wall = res[0][0]
    else:
        own = pool is None
        if own:
            for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
                os.environ[var] = "1"
            import multiprocessing as mp
            pool = mp.get_context("spawn").Pool(procs)
            pool.map(_cpu_warm, range(procs))
        try:
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(size, 1000 + i) for i in range(procs)], chunksize=1)
            wall = time.perf_counter() - t0
        finally:
            if own:
                pool.close()
                pool.join()
    vox = sum(r[1] for r in res)
    return {"value": vox / wall, "unit": "voxels/s", "cores": procs, "kind": "port",
            "sample": f"{procs} x {size}^3 crop(s) of the tubular phantom, 6 sigmas, oracle.pipeline.filter_frame, "
                      f"{wall:.1f} s wall"}


def _host_procs(size):
    """All host cores, bounded by memory: one oracle process peaks at ~1.8 GB for a 192^3 crop (measured), scaling
    with the crop volume; never plan for more than half of what the host / cgroup has available."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    need = 1.8e9 * (size / 192.0) ** 3 * 1.3 + 0.6e9        # + the interpreter with numpy / scipy / torch loaded
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = float(line.split()[1]) * 1024.0
    except OSError:
        pass
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v.isdigit():
                avail = min(avail, float(v)) if avail else float(v)
        except OSError:
            pass
    if avail:
        cores = min(cores, max(1, int(0.5 * avail / need)))
    return max(1, min(cores, 128))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = _host_procs(args.cpu_size)
    # one single-threaded oracle process per core: without this every process starts a BLAS/OpenMP pool as wide as the
    # machine and the oversubscription costs the reference arm 8x (measured)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    steps = max(1, args.steps)
    vals = []
    warm = min(1, max(0, args.warmup))      # numpy/scipy need no more than one pass to page everything in
    pool = None
    if procs > 1:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(procs)
        pool.map(_cpu_warm, range(procs))   # interpreter start-up and imports are not the reference's work
    try:
        for _ in range(warm):
            cpu_baseline(args.cpu_size, procs, pool)
        t_all0 = time.perf_counter()
        for _ in range(steps):
            vals.append(cpu_baseline(args.cpu_size, procs, pool))
        wall = time.perf_counter() - t_all0
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    v = float(np.mean([x["value"] for x in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (f64 accumulate / eigvalsh)", "data": "synthetic",
            "config": {"workload": f"reference CPU path (oracle port) on {procs} x {args.cpu_size}^3 crops of the "
                                   "1024^3 tubular phantom workload, 6 sigmas", "sigmas": SIGMAS_CFG3},
            "cpu_baseline": dict(vals[-1], value=v),
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nellie_b200.engine import FilterParams, FrangiEngine3D
    from nellie_b200.phantoms import tubular_phantom

    n = args.size
    shape = (n, n, n)
    params = FilterParams(dim_res=DIM_RES_CFG3, no_z=False, sigmas=SIGMAS_CFG3)
    if world > 1:
        from nellie_b200.sharding import ZShardedFilter
        runner = ZShardedFilter(shape, params, dev)
        frame = runner.make_phantom_slab(seed=3)
        eng = runner.engine
        step = lambda: runner.filter_frame(frame)                     # noqa: E731
    else:
        eng = FrangiEngine3D(shape, params, device=dev)
        frame = tubular_phantom(shape, seed=3, device=dev, n_tubes=args.tubes)
        step = lambda: eng.filter_frame(frame)                        # noqa: E731
        runner = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launches = 0
    eng.kernels = 0
    eng.profile = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_summary()
    # per-sigma average launch time of the per-sigma kernels (calls come in sigma order, once per sigma and step)
    per_sigma = {}
    nsig = len(SIGMAS_CFG3)
    for name in ("nb200_gauss_axis", "nb200_gauss_yx", "nb200_hessian_stats_code", "nb200_frangi_sparse", "nb200_frangi_accumulate",
                 "nb200_hessian_stats_fast", "nb200_frangi_fast"):
        evs = [(a, b) for nm, a, b in (eng.profile or []) if nm == name]
        if evs and len(evs) % nsig == 0:
            per_sigma[name] = [round(float(np.mean([a.elapsed_time(b) for a, b in evs[i::nsig]])), 3) for i in range(nsig)]
    eng.profile = None
    timed_launches = eng.kernels
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    voxels = float(n) ** 3
    value = voxels * args.steps / (ms * 1e-3)

    # ---- end to end through the public frame-stream API with HOST buffers (pinned): every step uploads its
    # input frame and downloads its result inside the timed region; nellie_b200.pipeline.FramePipeline (the T loop
    # of Filter.run) overlaps the upload of step t+1 and the download of step t-1 with the kernels of step t
    from nellie_b200.pipeline import FramePipeline
    own_shape = tuple(eng.out.shape)
    host_in = torch.empty(own_shape, dtype=torch.float32).pin_memory()
    host_in.copy_(frame)
    host_out = [torch.empty(own_shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = FramePipeline(eng, depth=2)
    pipe.run(2, lambda t: host_in, lambda t: host_out[t % 2])               # warm-up: allocates the staging buffers
    barrier()
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    # device-timed like the resident figure: run() makes the caller's stream wait for both copy streams before it returns,
    # so an event recorded after it completes after the last download
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    pipe.run(args.steps, lambda t: host_in, lambda t: host_out[t % 2])
    p1.record()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([p0.elapsed_time(p1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": voxels * args.steps / float(dt.item()), "unit": "voxels/s",
           "h2d_bytes_per_step": int(pipe.h2d_bytes // args.steps) * world, "d2h_bytes_per_step": int(pipe.d2h_bytes // args.steps) * world,
           "how": "FramePipeline: pinned host frame -> H2D -> Filter path -> D2H -> pinned host frame, every step; "
                  "transfers of neighbouring steps overlap the kernels (depth 2)"}
    del pipe, host_in, host_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, which = measured_peaks()
    # roofline of every volume kernel: algorithmic bytes per launch over the average
    # CUDA-event time of its launches inside the timed region; `roofline` is the dominant one, K3 is always listed
    vox_per_launch = voxels / world
    ALG = {"nb200_gauss_axis": (8.0, "K1z gauss_z_vec: blur along Z (R 4 + W 4)"),
           "nb200_gauss_yx": (8.0, "K1yx gauss_yx_tile: blur along Y and X fused (R 4 + W 4)"),
           "nb200_hessian_stats_code": (8.0, "K2 march_kernel<StatsEpi>: dense Hessian, max|H|, frob samples, per-voxel record (R gauss 4 + W code 4)"),
           "nb200_frangi_sparse": (12.0, "K3 sparse_stream + sparse_solve: mask + eigenvalues + vesselness + max/AND (R code 4 + R acc 4 + W acc 4)"),
           "nb200_frangi_accumulate": (12.0, "K3 march_kernel<FrangiEpi> (dense form)"),
           "nb200_hessian_stats_fast": (4.0, "K2 stats_fast_kernel + fixup + shell: max|H|, frob samples, value range (R gauss 4)"),
           "nb200_frangi_fast": (12.0, "K3 frangi_fast_kernel + shell: mask + eigenvalues + vesselness + max/AND (R gauss 4 + R acc 4 + W acc 4)"),
           "nb200_finalize_opening": (8.0, "K5 opening_march: percentile mask + binary opening (R 4 + W 4)")}
    # DRAM bytes per launch from the ncu --set full captures under profiles/ (1024^3, one GPU): read + write
    # (profiles/r1e_ncu_full_1024.md; K3 = stream + solve of the sigma-1.4 launch, the sigma-1.0 launch moves 36.7e9)
    TRAFFIC_1024 = {"nb200_gauss_axis": 8.62e9, "nb200_gauss_yx": 8.55e9, "nb200_hessian_stats_code": 8.75e9,
                    "nb200_frangi_sparse": 17.6e9, "nb200_finalize_opening": 9.32e9}
    roofs = {}
    for name, (bpv, label) in ALG.items():
        cnt, tot = prof.get(name, (0, 0.0))
        if not cnt:
            continue
        achieved = bpv * vox_per_launch / (tot / cnt * 1e-3) / 1e9
        roofs[name] = {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": which,
                       "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                       "traffic": TRAFFIC_1024.get(name) if (world == 1 and n == 1024) else None,
                       "algorithmic_bytes_per_launch": bpv * vox_per_launch, "avg_launch_ms": tot / cnt,
                       "share_of_step": tot / ms}
    roofline = max(roofs.values(), key=lambda r: r["share_of_step"]) if roofs else None
    breakdown = {k: {"launches": v[0], "ms_per_step": v[1] / args.steps} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    cpu = cpu_baseline(args.cpu_size, 1) if (world == 1 and not args.no_cpu) else None
    line = {"metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (f64 blur accumulate, f64-polished eigenvalues)", "data": "synthetic",
            "config": {"workload": f"synthetic {n}^3 fp32 tubular phantom, {len(SIGMAS_CFG3)} sigmas "
                                   f"{SIGMAS_CFG3}, dim_res 0.1 um isotropic" + (f", Z-sharded over {world} GPUs with halo exchange" if world > 1 else ""),
                       "l2_policy": "inputs larger than L2 (4 B x voxels per buffer >> 126 MB)", "seed": 3},
            "clocks": clocks, "e2e": e2e, "gpu_launches": timed_launches, "roofline": roofline, "roofline_kernels": roofs, "cpu_baseline": cpu,
            "kernel_ms_per_step": breakdown, "kernel_ms_per_sigma": per_sigma}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries write to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is written to
    the saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="edge of the cubic volume (1024 = BASELINE config #3)")
    ap.add_argument("--tubes", type=int, default=None)
    ap.add_argument("--cpu-size", type=int, default=None,
                    help="edge of the crop the CPU arms run (default: 192 for the cpu_baseline sample; the reference arm "
                         "shrinks it so that steps+warmup passes end within ~3 minutes)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.cpu_size is None:
        args.cpu_size = 192
        if args.impl == "reference":
            # one 192^3 crop takes ~40 s per pass (all cores busy with one crop each); keep the whole run near 3 minutes
            budget = 180.0 / (max(1, args.steps) + min(1, max(0, args.warmup)))
            args.cpu_size = int(min(192, max(96, 16 * round(192.0 * (budget / 40.0) ** (1.0 / 3.0) / 16))))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
