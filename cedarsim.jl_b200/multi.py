"""Multi-GPU partitioning of a sweep: sweep points are independent, so rank g of G
owns the contiguous block [g*ceil(B/G), min(B, (g+1)*ceil(B/G))) and there is no collective on
the solve path; only the final waveforms / operating points are gathered to rank 0 (NCCL on the
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def partition(B: int, world: int, rank: int) -> Tuple[int, int]:
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def gather_blocks(local, world: int, rank: int, sizes: List[int], axis: int = -1, dst: int = 0):
    """Gather per-rank result blocks (torch tensors, same shape except along `axis`) to `dst` and
    concatenate them in rank order.  Uses the default torch.distributed process group."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    axis = axis % local.dim()
    pad = max(sizes)
    shape = list(local.shape)
    shape[axis] = pad
    buf = torch.zeros(shape, dtype=local.dtype, device=local.device)
    idx = [slice(None)] * local.dim()
    idx[axis] = slice(0, local.shape[axis])
    buf[tuple(idx)] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    parts = []
    for r, n in enumerate(sizes):
        idx[axis] = slice(0, n)
        parts.append(out[r][tuple(idx)])
    return torch.cat(parts, dim=axis)
