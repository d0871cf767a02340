"""Frame stream of the Filter stage: the T loop of filtering.py:1005-1031 as a three-stage pipeline.

The reference handles one frame at a time: ``frame = memmap[t]`` → compute → ``memmap[t] = result;
flush()``.  On a B200 the per-frame compute of a 1024^3 volume (~0.1 s) is of the same order as one PCIe
transfer of the frame (4 GiB at ~55 GB/s), so the stage is laid out as

    host read → [pinned] → H2D (copy stream) → compute (caller's stream) → D2H (copy stream) → [pinned] → host write

with ``depth`` buffers per stage: the upload of frame t+1 and the download of frame t−1 overlap the
kernels of frame t (PCIe is full duplex), and host-side reads/writes of pageable memory (memmaps) run on
worker threads.  Ordering between the stages is carried by CUDA events only; the host never waits for
the GPU except when it must reuse a staging buffer.  Results are identical to the one-frame-at-a-time
loop (same kernels, same order per frame).

Works for one GPU (``FrangiEngine3D`` / ``FrangiEngine2D``) and for a Z slab per rank
(``ZShardedFilter.engine``: the collectives are issued on the compute stream like the kernels).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Optional

import numpy as np
import torch


_COPY_POOL = None


def parallel_copyto(dst: np.ndarray, src: np.ndarray, threads: int = 8, min_bytes: int = 16 << 20):
    """``np.copyto(dst, src)`` split along the first axis over a few threads (numpy releases the GIL inside the copy).
    One thread moves ~4 GB/s into pageable memory; a 512^3 int32 label frame is 512 MiB, so the single-threaded store
    of a frame cost more host time than the whole frame costs on the GPU."""
    global _COPY_POOL
    if dst.ndim == 0 or dst.nbytes < min_bytes or dst.shape[0] < 2:
        np.copyto(dst, src, casting="same_kind")
        return
    if _COPY_POOL is None:
        _COPY_POOL = ThreadPoolExecutor(threads, thread_name_prefix="nb200-copy")
    n = dst.shape[0]
    cuts = np.linspace(0, n, min(threads, n) + 1).astype(int)
    futs = [_COPY_POOL.submit(np.copyto, dst[a:b], src[a:b], "same_kind") for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    for f in futs:
        f.result()


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    arr = x if isinstance(x, np.ndarray) else np.asarray(x)
    if not arr.dtype.isnative:
        arr = arr.astype(arr.dtype.newbyteorder("="))
    if arr.dtype in (np.uint32, np.uint64, np.float64):
        arr = arr.astype(np.float32)     # xp.asarray(frame, dtype=float32) rounds once; same value here
    if not arr.flags.writeable:
        arr = arr.view()
        try:
            arr.flags.writeable = True   # torch.from_numpy warns on read-only views; we never write to inputs
        except ValueError:
            arr = np.array(arr)
    return torch.from_numpy(arr)


class FramePipeline:
    """Double-buffered H2D / compute / D2H pipeline around an engine's ``filter_frame``."""

    def __init__(self, engine, depth: int = 2):
        self.eng = engine
        self.dev = engine.device
        self.depth = int(depth)
        if self.depth < 2:
            # with one buffer the reader thread would refill the pinned input of frame t while its upload is still queued
            raise ValueError("FramePipeline needs depth >= 2 (double buffering)")
        self.frame_shape = tuple(engine.out.shape)
        with torch.cuda.device(self.dev):
            self.s_in = torch.cuda.Stream(self.dev)
            self.s_out = torch.cuda.Stream(self.dev)
            self.dev_out = [torch.empty(self.frame_shape, dtype=torch.float32, device=self.dev) for _ in range(self.depth)]
            mk = lambda: [torch.cuda.Event() for _ in range(self.depth)]          # noqa: E731
            self.ev_h2d, self.ev_loaded, self.ev_done, self.ev_d2h = mk(), mk(), mk(), mk()
        self.dev_in = [None] * self.depth
        self.pin_in = [None] * self.depth
        self.pin_out = [None] * self.depth
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -- staging buffers (allocated on first use, reused for every frame) ---------------------------
    def _dev_in(self, k, dtype):
        if self.dev_in[k] is None or self.dev_in[k].dtype != dtype:
            self.dev_in[k] = torch.empty(self.frame_shape, dtype=dtype, device=self.dev)
        return self.dev_in[k]

    def _pin_in(self, k, dtype):
        if self.pin_in[k] is None or self.pin_in[k].dtype != dtype:
            self.pin_in[k] = torch.empty(self.frame_shape, dtype=dtype).pin_memory()
        return self.pin_in[k]

    def _pin_out(self, k):
        if self.pin_out[k] is None:
            self.pin_out[k] = torch.empty(self.frame_shape, dtype=torch.float32).pin_memory()
        return self.pin_out[k]

    def run(self, num_t: int, get_in: Callable[[int], object], get_out: Callable[[int], object],
            apply_mask_volume: bool = True, on_frame: Optional[Callable[[int], None]] = None,
            after_store: Optional[Callable[[int], None]] = None):
        """Process frames 0..num_t-1.  ``get_in(t)`` returns the host frame (ndarray / memmap slice / CPU
        tensor; pinned tensors are uploaded in place, anything else goes through a pinned staging buffer
        filled on a worker thread); ``get_out(t)`` returns the host destination (pinned float32 tensor:
        downloaded in place; ndarray / memmap slice: written by a worker thread, then ``after_store(t)``)."""
        eng, depth = self.eng, self.depth
        with torch.cuda.device(self.dev):
            compute = torch.cuda.current_stream(self.dev)
            for ev in self.ev_h2d + self.ev_loaded + self.ev_done + self.ev_d2h:
                ev.record(compute)                     # "never used" state: every wait below passes
            reader = ThreadPoolExecutor(1, thread_name_prefix="nb200-read")
            writer = ThreadPoolExecutor(1, thread_name_prefix="nb200-write")
            reads, writes = {}, [None] * depth

            def stage_in(t):
                """Host frame t → something the copy engine can read: the pinned source itself, or a staging
                buffer (after the upload that last used it has finished)."""
                src = _as_tensor(get_in(t))
                if tuple(src.shape) != self.frame_shape:
                    raise ValueError(f"frame {t} has shape {tuple(src.shape)}, pipeline was built for {self.frame_shape}")
                if src.is_pinned():
                    return src
                k = t % depth
                with torch.cuda.device(self.dev):      # worker threads start on device 0: pin against OUR device
                    self.ev_h2d[k].synchronize()
                    buf = self._pin_in(k, src.dtype)
                parallel_copyto(buf.numpy(), src.numpy())
                return buf

            def store(t, k, dst):
                with torch.cuda.device(self.dev):
                    self.ev_d2h[k].synchronize()
                if isinstance(dst, torch.Tensor):
                    dst.copy_(self.pin_out[k])
                else:
                    # memmap slice / ndarray of any byte order or real dtype: numpy converts while assigning IN PLACE
                    # (going through a converted temporary would leave a big-endian or float64 destination untouched)
                    if not isinstance(dst, np.ndarray) or not dst.flags.writeable:
                        raise TypeError("FramePipeline: the destination of a frame must be a writable ndarray / memmap slice")
                    parallel_copyto(dst, self.pin_out[k].numpy())
                if after_store is not None:
                    after_store(t)

            try:
                if num_t > 0:
                    reads[0] = reader.submit(stage_in, 0)
                for t in range(num_t):
                    k = t % depth
                    if on_frame is not None:
                        on_frame(t)
                    hsrc = reads.pop(t).result()
                    if t + 1 < num_t:
                        reads[t + 1] = reader.submit(stage_in, t + 1)      # overlaps the enqueue + kernels of frame t
                    # ---- H2D on the upload stream (after frame t-depth has left dev_in[k]) ----
                    din = self._dev_in(k, hsrc.dtype)
                    self.s_in.wait_event(self.ev_loaded[k])
                    with torch.cuda.stream(self.s_in):
                        din.copy_(hsrc, non_blocking=True)
                        self.ev_h2d[k].record(self.s_in)
                    self.h2d_bytes += hsrc.numel() * hsrc.element_size()
                    # ---- compute on the caller's stream ----
                    compute.wait_event(self.ev_h2d[k])
                    eng_out = self.dev_out[k]
                    compute.wait_event(self.ev_d2h[k])                     # frame t-depth has left dev_out[k]
                    if hasattr(eng, "load_frame"):
                        eng.load_frame(din)
                        self.ev_loaded[k].record(compute)
                        eng.run_sigmas()
                        eng.finalize(apply_mask_volume, out=eng_out)
                    else:
                        eng.filter_frame(din, apply_mask_volume, out=eng_out)
                        self.ev_loaded[k].record(compute)
                    self.ev_done[k].record(compute)
                    # ---- D2H on the download stream ----
                    dst = get_out(t)
                    direct = isinstance(dst, torch.Tensor) and dst.is_pinned()
                    if writes[k] is not None:
                        writes[k].result()                                 # pin_out[k] is free again
                        writes[k] = None
                    hdst = dst if direct else self._pin_out(k)
                    self.s_out.wait_event(self.ev_done[k])
                    with torch.cuda.stream(self.s_out):
                        hdst.copy_(eng_out, non_blocking=True)
                        self.ev_d2h[k].record(self.s_out)
                    self.d2h_bytes += eng_out.numel() * 4
                    if not direct:
                        writes[k] = writer.submit(store, t, k, dst)
                for w in writes:
                    if w is not None:
                        w.result()
                self.s_out.synchronize()
                compute.wait_stream(self.s_in)
                compute.wait_stream(self.s_out)
            finally:
                reader.shutdown(wait=True)
                writer.shutdown(wait=True)
