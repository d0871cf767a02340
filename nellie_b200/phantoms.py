"""Seeded synthetic inputs for the BASELINE.json configurations.

Tubular phantom: background 100, line segments with a Gaussian cross-section (radius
U(2,5) px, amplitude 400) combined by max, plus N(0,10) noise, float32.  The segment
parameters come from ``numpy.random.default_rng(seed)`` on the host; rasterisation and
noise run in torch on whatever device is asked for (CPU for the small parity cases, the
GPU for the 512^3 / 1024^3 bench volumes).  The same array is always handed to the CPU
oracle and to the CUDA path, so CPU-vs-GPU RNG differences never enter a comparison.
"""
from __future__ import annotations

import numpy as np
import torch

_REF_VOXELS = 96 * 192 * 192  # density anchor: 40 tubes per 96x192x192


def _segments(shape, seed, n_tubes=None):
    rng = np.random.default_rng(seed)
    nd = len(shape)
    vox = int(np.prod(shape))
    if n_tubes is None:
        n_tubes = int(min(2048, max(1, round(40 * vox / _REF_VOXELS))))
    hi = np.asarray(shape, dtype=np.float64) - 1.0
    centre = rng.uniform(0.0, 1.0, size=(n_tubes, nd)) * hi
    direction = rng.standard_normal(size=(n_tubes, nd))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True) + 1e-12
    half = rng.uniform(0.2, 0.6, size=(n_tubes, 1)) * float(min(shape))
    radius = rng.uniform(2.0, 5.0, size=n_tubes)
    return centre - half * direction, centre + half * direction, radius


def tubular_phantom(shape, seed, device="cpu", n_tubes=None, noise_sd=10.0, background=100.0,
                    amplitude=400.0, piece=24.0):
    """Return a float32 torch tensor of ``shape`` (2-D or 3-D) on ``device``."""
    shape = tuple(int(s) for s in shape)
    nd = len(shape)
    dev = torch.device(device)
    p0, p1, radius = _segments(shape, seed, n_tubes)
    vol = torch.zeros(shape, dtype=torch.float32, device=dev)
    for a, b, r in zip(p0, p1, radius):
        length = float(np.linalg.norm(b - a))
        n_piece = max(1, int(np.ceil(length / piece)))
        ts = np.linspace(0.0, 1.0, n_piece + 1)
        reach = 4.0 * r
        for k in range(n_piece):
            qa = a + ts[k] * (b - a)
            qb = a + ts[k + 1] * (b - a)
            lo = np.floor(np.minimum(qa, qb) - reach).astype(int)
            hi = np.ceil(np.maximum(qa, qb) + reach).astype(int) + 1
            lo = np.maximum(lo, 0)
            hi = np.minimum(hi, np.asarray(shape))
            if np.any(hi <= lo):
                continue
            axes = [torch.arange(int(l), int(h), device=dev, dtype=torch.float32) for l, h in zip(lo, hi)]
            grid = torch.meshgrid(*axes, indexing="ij")
            d = [g - float(q) for g, q in zip(grid, qa)]
            seg = qb - qa
            seg_len2 = float(np.dot(seg, seg)) + 1e-12
            t = sum(di * float(si) for di, si in zip(d, seg)) / seg_len2
            t = t.clamp_(0.0, 1.0)
            dist2 = sum((di - t * float(si)) ** 2 for di, si in zip(d, seg))
            val = amplitude * torch.exp(-dist2 / (2.0 * float(r) ** 2))
            sl = tuple(slice(int(l), int(h)) for l, h in zip(lo, hi))
            vol[sl] = torch.maximum(vol[sl], val)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    # noise in slabs so a 1024^3 volume never needs a second full-size temporary
    step = max(1, (1 << 26) // max(1, int(np.prod(shape[1:]))))
    for z0 in range(0, shape[0], step):
        sl = vol[z0:z0 + step]
        sl += background
        sl += noise_sd * torch.randn(sl.shape, generator=gen, device=dev, dtype=torch.float32)
    return vol


def tubular_phantom_np(shape, seed, **kw):
    """Host numpy array of the same phantom (small parity cases)."""
    return tubular_phantom(shape, seed, device="cpu", **kw).numpy()
