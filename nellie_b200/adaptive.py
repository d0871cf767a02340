"""Retry ladder of the stage classes' ``run()`` (reference: nellie/utils/adaptive_run.py:103-141,
filtering.py:1033-1076, labelling.py:736-778).

The reference walks ``(device, low_memory)`` candidates — gpu/high, gpu/low, cpu/high, cpu/low — and moves to the next one
when a run dies of an out-of-memory error or because the GPU backend is unavailable; any other exception aborts the
stage.  nellie_b200 implements ONE rung itself (the B200 path) and has no CPU code of its own: with
``fallback="reference"`` the remaining rungs are the reference's own classes (``nellie.segmentation``), imported only
at that moment; without it (the default) a failure of the B200 rung raises, loudly.
"""
from __future__ import annotations

import logging

logger = logging.getLogger("nellie_b200")


def is_oom_error(exc: Exception) -> bool:
    """adaptive_run.py:116-127: MemoryError, or any exception whose repr mentions an out-of-memory condition
    (``_cabi.NellieB200OutOfMemory`` derives from MemoryError; torch's CUDA OOM text matches the second test)."""
    if isinstance(exc, MemoryError):
        return True
    msg = repr(exc).lower()
    return "out of memory" in msg or "outofmemory" in msg


def is_gpu_unavailable_error(exc: Exception) -> bool:
    """adaptive_run.py:130-141 (the cupy import test is replaced by the missing CUDA library of this package)."""
    msg = repr(exc).lower()
    return ("gpu backend requested" in msg or "cuda is not available" in msg or "no cuda devices" in msg
            or "libnellie_b200.so is missing" in msg or "nellie_b200.so is missing" in msg)


def run_with_ladder(stage_name, run_b200, run_reference, fallback):
    """``run_b200()`` first; on an OOM / GPU-unavailable failure continue with ``run_reference(device, low_memory)`` for
    ("cpu", False) then ("cpu", True) when ``fallback == "reference"``.  Other exceptions propagate at once (they abort
    the stage in the reference too, nellie_processor.py:568-605)."""
    rungs = [("b200", False)]
    if fallback == "reference":
        rungs += [("cpu", False), ("cpu", True)]
    elif fallback is not None:
        raise ValueError(f"fallback must be None or 'reference', not {fallback!r}")
    last_exc = None
    for k, (dev, low) in enumerate(rungs):
        try:
            if dev == "b200":
                run_b200()
            else:
                run_reference(dev, low)
            return
        except Exception as exc:                       # noqa: BLE001 - same breadth as the reference's ladder
            last_exc = exc
            more = k + 1 < len(rungs)
            if more and dev == "b200" and is_gpu_unavailable_error(exc):
                logger.warning("%s: B200 backend unavailable (%s); retrying with the reference on CPU.", stage_name, exc)
                continue
            if more and is_oom_error(exc):
                logger.warning("%s: out of memory on %s/%s; retrying with lower settings.", stage_name, dev,
                               "low-memory" if low else "high-memory")
                continue
            raise
    raise last_exc
