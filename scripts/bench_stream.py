"""bench.py --config 4 | 5: the frame-stream configurations of BASELINE.json.

  4: 2-D + T stream of 2048 x 2048 frames (per-frame 2x2 Hessian eigen path incl. LoG blobness), T-sharded
  5: 3-D + T, 512^3 frames, Filter -> Label per frame (the label hierarchy's first level), T-sharded

One process per GPU; frames are independent (per-frame gamma / thresholds / label ids), so rank r owns its own frames
and there is no data-path collective ("scaling": "weak": every rank processes `--frames` frames per step).
`value` = frames resident on the device, CUDA events; `e2e` = the public classes (`Filter.run`, `Label.run`) on host
arrays (the memmap contract of the reference): host -> device -> host every frame, inside the timed region.
"""
from __future__ import annotations

import os
import time

import numpy as np


class _MemInfo:
    """Duck-typed im_info with in-memory 'memmaps' (reference: tests/test_labelling.py:16-22)."""

    def __init__(self, raw_t, dim_res, no_z):
        self.no_t, self.no_z = False, no_z
        self.shape = raw_t.shape
        self.axes = "TYX" if no_z else "TZYX"
        self.dim_res = dim_res
        self.im_path = "raw"
        self.pipeline_paths = {"im_preprocessed": "pre", "im_instance_label": "lab"}
        self.store = {"raw": raw_t}

    def get_memmap(self, path, read_mode="r+"):
        return self.store[path]

    def allocate_memory(self, path, dtype="float32", description="", return_memmap=True, data=None, read_mode="r+"):
        self.store[path] = np.zeros(self.shape, dtype=dtype)
        return self.store[path]


def run_stream_config(args, ClockSampler, measured_peaks):
    import torch
    import torch.distributed as dist
    from nellie_b200 import Filter, Label
    from nellie_b200.phantoms import tubular_phantom_np

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    two_d = args.config == 4
    if two_d:
        n = args.size or 2048
        frames_n = args.frames or 16
        frame_shape = (n, n)
        dim_res = {"X": 0.1, "Y": 0.1, "Z": None, "T": 1.0}
        mk = lambda t: tubular_phantom_np((1, n, n), seed=4000 + t, n_tubes=200)[0]          # noqa: E731
        metric, unit = "pixels/s 2-D Frangi + LoG (Filter) on a 2048^2 frame stream", "pixels/s"
    else:
        n = args.size or 512
        frames_n = args.frames or 2
        frame_shape = (n, n, n)
        dim_res = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
        mk = lambda t: tubular_phantom_np(frame_shape, seed=5000 + t)                          # noqa: E731
        metric, unit = "voxels/s Filter -> Label (segmentation + label hierarchy level 1) on 512^3 frames", "voxels/s"
    # distinct frames per rank (T-sharding: frame t of the global stream belongs to rank t mod world)
    host = np.stack([mk(rank + world * k) for k in range(frames_n)]).astype(np.float32)
    info = _MemInfo(host, dim_res, two_d)
    filt = Filter(info, device="b200", cuda_device=dev)
    filt._get_t()
    filt._set_default_sigmas()
    lab = None if two_d else Label(info, device="b200", cuda_device=dev)
    dev_frames = [torch.from_numpy(host[k]).to(dev) for k in range(frames_n)]
    vox = float(np.prod(frame_shape))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    lab_ms = [0.0]

    def step():
        for k in range(frames_n):
            out = filt.filter_frame_device(dev_frames[k])
            if lab is not None:
                l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0.record()
                labels, _ = lab.label_frame_device(out)
                l1.record()
                lab_events.append((l0, l1))
        return out if lab is None else labels

    lab_events = []
    for _ in range(max(3, args.warmup)):
        res = step()
    n_labels = None if lab is None else int(res.max().item())
    nz = int((res > 0).sum().item())
    lab_events.clear()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng = filt._engine_for(frame_shape)
    eng.launches = 0
    if hasattr(eng, "kernels"):
        eng.kernels = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    lab_ms = sum(a.elapsed_time(b) for a, b in lab_events) / max(1, len(lab_events)) if lab_events else None
    launches = getattr(eng, "kernels", 0) or eng.launches
    if lab is not None:
        launches += lab._engine.launches if lab._engine is not None else 0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = vox * frames_n * world * args.steps / (ms * 1e-3)

    # ---- end to end: the public run() methods on host arrays (upload, compute, download, store — every frame) ----
    filt.run()                      # warm-up (allocates staging buffers, output arrays)
    if lab is not None:
        lab.run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        filt.run()
        if lab is not None:
            lab.run()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    in_bytes = host[0].nbytes * frames_n * (1 if lab is None else 2)        # Label re-reads the float32 frangi frame
    out_bytes = frames_n * int(vox) * (4 if lab is None else 8)             # float32 frame (+ int32 labels)
    e2e = {"value": vox * frames_n * world * args.steps / float(dt.item()), "unit": unit,
           "h2d_bytes_per_step": int(in_bytes) * world, "d2h_bytes_per_step": int(out_bytes) * world,
           "how": "Filter.run()" + (" + Label.run()" if lab is not None else "") + " on in-memory im_info arrays: every frame "
                  "goes host -> device -> host inside the timed region (host wall clock around run(), max over ranks)"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    peaks, which = measured_peaks()
    nsig = len(filt.sigmas)
    survey_bytes = (24.0 * nsig + 8.0) * vox if lab is None else (24.0 * nsig + 8.0 + 8.0) * vox
    per_frame_ms = ms / (args.steps * frames_n)
    roofline = {"bound": "hbm", "kernel": "whole frame (SURVEY 8d: (24 S + 8) B/voxel Filter" + (" + 8 B/voxel Label)" if lab else ")"),
                "achieved": survey_bytes / (per_frame_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "peak_source": which, "unit": "GB/s",
                "traffic": None}
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 blur accumulate, f64-polished eigenvalues), int32 labels", "data": "synthetic",
            "config": {"workload": f"{frames_n} frames of {'x'.join(map(str, frame_shape))} per rank and step, {nsig} sigmas "
                                   f"{[round(s, 3) for s in filt.sigmas]}, T-sharded over {world} GPU(s)", "baseline_config": args.config,
                       "l2_policy": "frames larger than L2" if vox * 4 > 126e6 else "16 MiB frames: the working set of one frame fits L2 "
                                    "(consecutive frames differ, the stream is 256 MiB per step)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
            "ms_per_frame": per_frame_ms, "label_ms_per_frame": lab_ms, "labels_in_last_frame": n_labels, "nonzero_in_last_frame": nz}
    if world > 1:
        dist.destroy_process_group()
    return line
