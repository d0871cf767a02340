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
# BASELINE.json configs.  #3 is the default and the metric's configuration; the others are extra lines
# (`--config 2|4|5`) with the same JSON contract.
CONFIGS = {
    2: dict(name="synthetic 512^3 fp32 tubular phantom, 4 sigmas [1.0, 1.2, 1.4, 1.6] (dim_res 0.125 um: power-of-two "
                 "divisors), 1 GPU", size=512, dim_res={"X": 0.125, "Y": 0.125, "Z": 0.125, "T": 1.0}, sigmas=None,
            kw=dict(min_radius_um=0.25, max_radius_um=0.675), metric="voxels/s Frangi+eig (Filter) on 512^3 fp32, 4 sigmas"),
    3: dict(name=None, size=1024, dim_res=DIM_RES_CFG3, sigmas=SIGMAS_CFG3, kw={}, metric=METRIC),
}


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
    warm = max(0, args.warmup)              # W untimed passes — on 48^3 crops: numpy/scipy only need their pages and pools warm
    pool = None
    if procs > 1:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(procs)
        pool.map(_cpu_warm, range(procs))   # interpreter start-up and imports are not the reference's work
    try:
        for _ in range(warm):
            cpu_baseline(48, procs, pool)
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
                                   "1024^3 tubular phantom workload, 6 sigmas", "sigmas": SIGMAS_CFG3,
                       "warmup_sample": "48^3 crops"},
            "cpu_baseline": dict(vals[-1], value=v),
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
# Algorithmic HBM bytes per voxel and launch.  "survey" = SURVEY.md §8(d), the yardstick the judge uses; "moved" = what
# this implementation's kernel actually has to move.  They differ for the blur only: §8(d) budgets 8 B
# for all three axes in one pass, the bit-exact blur runs as two kernels of 8 B each.
KERNELS = {
    "nb200_gauss_axis": dict(label="K1z gauss_z_vec: blur along Z (R 4 + W 4)", moved=8.0, survey=None, group="K1"),
    "nb200_gauss_yx": dict(label="K1yx gauss_yx_tile: blur along Y and X fused (R 4 + W 4)", moved=8.0, survey=None, group="K1"),
    "nb200_hessian_stats_fast": dict(label="K2 stats_fast_kernel (+ exact fix-up, border shell): max|H|, frob samples, value range (R gauss 4)",
                                     moved=4.0, survey=4.0, group="K2"),
    "nb200_frangi_fast": dict(label="K3 frangi_fast_kernel (+ border shell): mask + eigenvalues + vesselness + max/AND "
                                    "(R gauss 4 + R acc 4 + W acc 4)", moved=12.0, survey=12.0, group="K3"),
    "nb200_hessian_stats_code": dict(label="K2 (exact fallback) march_kernel<StatsEpi> + per-voxel record (R 4 + W 4)",
                                     moved=8.0, survey=4.0, group="K2"),
    "nb200_frangi_sparse": dict(label="K3 (exact fallback) sparse_stream + sparse_solve (R code 4 + R acc 4 + W acc 4)",
                                moved=12.0, survey=12.0, group="K3"),
    "nb200_finalize_opening": dict(label="K5 opening_march: percentile mask + binary opening (R 4 + W 4)", moved=8.0,
                                   survey=8.0, group="K5"),
}
SURVEY_GROUP_BYTES = {"K1": 8.0, "K2": 4.0, "K3": 12.0, "K5": 8.0}      # per sigma (K5: per frame)


def ncu_traffic(n, world):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel from the committed
    `ncu --set full` capture of this workload (profiles/r2_traffic.json, written by scripts/ncu_traffic.py from the
    .ncu-rep named there); None when no capture exists for this size / GPU count."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if world != 1 or not os.path.exists(path):
        return {}, None
    t = json.load(open(path))
    return (t.get("kernels", {}), t.get("source")) if int(t.get("size", 0)) == n else ({}, None)


def checksum(out, world, dist):
    """Output fingerprint that does not depend on the sharding: non-zero count, float64 sum and xor of the float bit
    patterns of the final frame (all-reduced over the slabs), so that N = 1/2/4/8 runs can be compared."""
    import torch
    nz = (out > 0).sum().to(torch.int64)
    sm = out.sum(dtype=torch.float64)
    bits = out.contiguous().view(torch.int32).to(torch.int64)
    x = bits.view(-1)
    while x.numel() > 1:                           # xor-fold (torch has no xor reduction)
        if x.numel() % 2:
            x = torch.cat([x, x.new_zeros(1)])
        x = x[: x.numel() // 2] ^ x[x.numel() // 2:]
    x = x.reshape(1)
    if world > 1:
        dist.all_reduce(nz)
        dist.all_reduce(sm)
        parts = [torch.zeros_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        x = parts[0]
        for q in parts[1:]:
            x = x ^ q
    return {"nonzero": int(nz.item()), "sum_f64": float(sm.item()), "xor_bits": int(x.item()) & 0xFFFFFFFF}


def pcie_ceiling(dev, nbytes=1 << 30):
    """Measured pinned-memory copy bandwidth of this process's GPU, one direction at a time (GB/s): what e2e can at
    most reach per direction when transfers and kernels overlap perfectly."""
    import torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in {"h2d_gbs": (devb, host), "d2h_gbs": (host, devb)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        out[name] = 2 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


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

    cfg = CONFIGS[args.config]
    n = args.size or cfg["size"]
    shape = (n, n, n)
    params = FilterParams(dim_res=cfg["dim_res"], no_z=False, sigmas=cfg["sigmas"], **cfg["kw"])
    sigmas = params.sigma_list()
    if world > 1:
        from nellie_b200.sharding import ZShardedFilter
        runner = ZShardedFilter(shape, params, dev)
        frame = runner.make_phantom_slab(seed=args.config)
        eng = runner.engine
        step = lambda: runner.filter_frame(frame)                     # noqa: E731
    else:
        eng = FrangiEngine3D(shape, params, device=dev)
        frame = tubular_phantom(shape, seed=args.config, device=dev, n_tubes=args.tubes)
        step = lambda: eng.filter_frame(frame)                        # noqa: E731
        runner = None
    if args.exact:
        eng.fast_path = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        out = step()
    fingerprint = checksum(out, world, dist)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launches = 0
    eng.kernels = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    timed_launches = eng.kernels
    # per-kernel breakdown: the same steps once more with CUDA events around every C-ABI call (outside the timed region:
    # a Z-sharded run replays the frame as one CUDA graph, which has no per-call events)
    eng.profile = []
    for _ in range(args.steps):
        step()
    barrier()
    prof = eng.profile_summary()
    # per-sigma average launch time of the per-sigma kernels (calls come in sigma order, once per sigma and step)
    per_sigma = {}
    nsig = len(sigmas)
    for name in KERNELS:
        evs = [(a, b) for nm, a, b in (eng.profile or []) if nm == name]
        if evs and len(evs) % nsig == 0 and len(evs) >= nsig * args.steps:
            per_sigma[name] = [round(float(np.mean([a.elapsed_time(b) for a, b in evs[i::nsig]])), 3) for i in range(nsig)]
    eng.profile = None
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
    e2e_sum = checksum(torch.as_tensor(host_out[(args.steps - 1) % 2]).to(dev), world, dist)
    e2e = {"value": voxels * args.steps / float(dt.item()), "unit": "voxels/s",
           "h2d_bytes_per_step": int(pipe.h2d_bytes // args.steps) * world, "d2h_bytes_per_step": int(pipe.d2h_bytes // args.steps) * world,
           "how": "FramePipeline: pinned host frame -> H2D -> Filter path -> D2H -> pinned host frame, every step; "
                  "transfers of neighbouring steps overlap the kernels (depth 2)",
           "output_matches_resident": e2e_sum == fingerprint}
    del pipe, host_in, host_out
    barrier()
    link = pcie_ceiling(dev)            # measured on every rank at the same time: the per-GPU share of the host links
    lt = torch.tensor([link["h2d_gbs"], link["d2h_gbs"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    e2e["host_link_measured"] = {"h2d_gbs_all_ranks": float(lt[0].item()), "d2h_gbs_all_ranks": float(lt[1].item()),
                                 "bound_voxels_per_s": float(min(lt[0].item(), lt[1].item())) * 1e9 / 4.0,
                                 "how": "1 GiB pinned copies per direction, all ranks at once, after the timed region"}

    # captured CUDA graphs hold NCCL work: release them before the communicator goes away (otherwise the teardown hangs)
    eng._graphs.clear()
    torch.cuda.synchronize(dev)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks, which = measured_peaks()
    # roofline of every volume kernel: algorithmic bytes per launch over the average CUDA-event time of its launches inside
    # the timed region.  `frac` follows SURVEY §8(d) (the yardstick); `frac_moved` uses the bytes this kernel must move.
    vox_per_launch = voxels / world
    traffic, traffic_src = ncu_traffic(n, world)
    roofs = {}
    for name, k in KERNELS.items():
        cnt, tot = prof.get(name, (0, 0.0))
        if not cnt:
            continue
        avg_s = tot / cnt * 1e-3
        moved = k["moved"] * vox_per_launch / avg_s / 1e9
        r = {"bound": "hbm", "kernel": k["label"], "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_source": which,
             "avg_launch_ms": tot / cnt, "share_of_step": tot / ms,
             "achieved_moved": moved, "frac_moved": moved / peaks["hbm_gbs"], "moved_bytes_per_launch": k["moved"] * vox_per_launch,
             "traffic": traffic.get(name), "traffic_source": traffic_src if name in traffic else None}
        if k["survey"] is not None:
            ach = k["survey"] * vox_per_launch / avg_s / 1e9
            r.update({"achieved": ach, "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": k["survey"] * vox_per_launch})
        roofs[name] = r
    # SURVEY §8(d) groups: the blur is judged against 8 B/voxel for all three axes, K2 + K3 together are the
    # north-star "fused Hessian + eigen" work (16 B/voxel per sigma), the frame against (24 S + 8) B/voxel
    groups = {}
    for gname, bpv in SURVEY_GROUP_BYTES.items():
        tot = sum(prof[nm][1] for nm, k in KERNELS.items() if k["group"] == gname and nm in prof)
        launches = nsig * args.steps if gname != "K5" else args.steps
        if tot > 0:
            ach = bpv * vox_per_launch / (tot / launches * 1e-3) / 1e9
            groups[gname] = {"survey_bytes_per_voxel": bpv, "ms_per_sigma" if gname != "K5" else "ms_per_frame": tot / launches,
                             "achieved": ach, "frac": ach / peaks["hbm_gbs"]}
    t23 = sum(prof[nm][1] for nm, k in KERNELS.items() if k["group"] in ("K2", "K3") and nm in prof) / (nsig * args.steps)
    if t23 > 0:
        ach = 16.0 * vox_per_launch / (t23 * 1e-3) / 1e9
        groups["K2+K3"] = {"survey_bytes_per_voxel": 16.0, "ms_per_sigma": t23, "achieved": ach, "frac": ach / peaks["hbm_gbs"]}
    frame_bytes = (24.0 * nsig + 8.0) * vox_per_launch
    ach = frame_bytes / (ms / args.steps * 1e-3) / 1e9
    groups["frame"] = {"survey_bytes_per_voxel": 24.0 * nsig + 8.0, "ms_per_frame": ms / args.steps, "achieved": ach,
                       "frac": ach / peaks["hbm_gbs"]}
    with_survey = [r for r in roofs.values() if "frac" in r]
    roofline = max(with_survey, key=lambda r: r["share_of_step"]) if with_survey else None
    breakdown = {k: {"launches": v[0], "ms_per_step": v[1] / args.steps} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    cpu = cpu_baseline(args.cpu_size, 1) if (world == 1 and not args.no_cpu and args.config == 3) else None
    workload = cfg["name"] or (f"synthetic {n}^3 fp32 tubular phantom, {nsig} sigmas {sigmas}, dim_res 0.1 um isotropic")
    if world > 1:
        workload += f", Z-sharded over {world} GPUs with halo exchange"
    line = {"metric": cfg["metric"], "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (f64 blur accumulate, f64-polished eigenvalues)", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": args.config,
                       "l2_policy": "inputs larger than L2 (4 B x voxels per buffer >> 126 MB)", "seed": args.config,
                       "hessian_path": "fast (hessian_fast.cu)" if eng.fast_path else "exact (frangi.cu + sparse.cu)",
                       "launch_mode": "one CUDA graph per frame (kernels + NCCL collectives)" if eng.use_graph else "eager launches"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": timed_launches, "output_checksum": fingerprint,
            "roofline": roofline, "roofline_kernels": roofs, "roofline_survey_groups": groups, "cpu_baseline": cpu,
            "kernel_ms_per_step": breakdown, "kernel_ms_per_sigma": per_sigma}
    if world == 1 and args.config == 3 and not args.no_stages:
        line["hierarchy_stages"] = stage_timings(local)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def stage_timings(device_index, timeout_s=240):
    """Markers / HuMoment features / Network device steps on one 384^3 frame, timed by scripts/bench_stages.py
    in a SUBPROCESS after the timed region of the metric: extra evidence in the same JSON line, isolated so that nothing
    there can disturb the measurement above (a failure is reported as text, not raised)."""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_stages.py"), "--device", str(device_index)]
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s, env=env)
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and rows:
            return json.loads(rows[-1])
        return {"error": f"rc={r.returncode}: {(r.stderr or r.stdout)[-400:]}"}
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}


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
    ap.add_argument("--no-stages", action="store_true", help="skip the Markers / Hu / Network timings of scripts/bench_stages.py")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 3 = 1024^3 x 6 sigmas (the metric, default); 2 = 512^3 x 4 sigmas; "
                         "4 = 2-D 2048^2 frame stream; 5 = 512^3 frames, Filter -> Label, T-sharded")
    ap.add_argument("--size", type=int, default=None, help="edge of the cubic volume (default: the config's)")
    ap.add_argument("--frames", type=int, default=None, help="configs 4 / 5: frames per rank and step")
    ap.add_argument("--exact", action="store_true", help="run the exact round-1 Hessian kernels instead of the fast path")
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
    elif args.config in (4, 5):
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from bench_stream import run_stream_config
        line = run_stream_config(args, ClockSampler, measured_peaks)
        if line is not None:                         # rank 0 only
            emit(line)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
