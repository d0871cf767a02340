"""``Filter._remove_edges`` (nellie/segmentation/filtering.py:969-1000, bbox :227-250): in every Z slice (or in the
2-D frame) the rows of the bounding box of the non-zero response are found and a band of ``min(15, height)`` rows is
zeroed at its top and at its bottom.  Off by default in the reference (``remove_edges=False``).

Device-agnostic tensor plumbing without a host round trip (row reductions of one frame; not a hot kernel): the same
function runs on the CUDA accumulator inside the engines and on CPU tensors in ``tests/test_host_logic.py``, where it is
checked against the oracle's restatement of the reference loop."""
from __future__ import annotations

import torch

MARGIN = 15


def remove_edge_bands_(v: torch.Tensor, margin: int = MARGIN) -> torch.Tensor:
    """In place on a (Z, Y, X) or (Y, X) response whose non-positive entries count as empty (the engines keep
    -1 = "dead voxel" in the accumulator, the reference has zeros there)."""
    vol = v if v.dim() == 3 else v[None]
    ny = vol.shape[1]
    rows = (vol > 0).any(dim=2)                                     # (Z, Y): rows holding any response
    idx = torch.arange(ny, device=v.device)
    rmin = torch.where(rows, idx, torch.full_like(idx, ny)).amin(dim=1)
    rmax = torch.where(rows, idx, torch.full_like(idx, -1)).amax(dim=1)
    height = (rmax - rmin + 1).clamp(min=0)
    m = torch.clamp(height, max=int(margin))
    top = (idx >= rmin[:, None]) & (idx < (rmin + m)[:, None])
    bottom = (idx > (rmax - m)[:, None]) & (idx <= rmax[:, None])
    band = (top | bottom) & (height > 0)[:, None]
    vol.masked_fill_(band[:, :, None], 0)
    return v
