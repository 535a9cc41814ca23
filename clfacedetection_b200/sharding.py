"""Frame-level data parallelism across GPUs.

Frames are independent, so a stream of N frames is split into contiguous chunks, one per
rank (one process per GPU).  There is no collective in the hot loop; the only exchange is ONE
gather of the variable-length detection-rect lists at the end (all_gather of the counts, then
of the padded rect arrays), with frame indices made global.  Works on any torch.distributed
backend: NCCL tensors on the GPU box, gloo CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) of the frames rank `rank` owns: contiguous, sizes differ by at most 1."""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def rects_to_array(rects: np.ndarray, frame_offset: int = 0) -> np.ndarray:
    """structured clfd rect records -> int32 [n, 6] (x, y, w, h, GLOBAL frame, cascade)"""
    out = np.zeros((len(rects), 6), np.int32)
    for i, k in enumerate(("x", "y", "w", "h", "frame", "cascade")):
        out[:, i] = rects[k]
    out[:, 4] += frame_offset
    return out


def gather_rects(local: np.ndarray, device=None):
    """The single gather: every rank contributes int32 [n_i, 6]; every rank gets the
    concatenation in rank order (rank 0 is the consumer).  Without an initialised process
    group this is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    mine = torch.zeros((cap, 6), dtype=torch.int32, device=dev)
    if len(local):
        mine[:len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0)
