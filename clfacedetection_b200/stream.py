"""A frame STREAM sharded over GPUs (BASELINE.json configs[4]: 8192 1080p frames over 1/2/4/8 B200).

The stream is a deterministic function of the global frame index, so any rank can produce any
frame and an N-rank run can be compared with a 1-rank run frame by frame:

    frame g = canvas[(g // 64) % K][dy : dy + H, dx : dx + W],   dy = g % 64,  dx = (g // (64 * K)) % 16

with K fixed-seed octave-noise canvases of (H + 63) x (W + 15) pixels.  64 consecutive frames are 64
overlapping VIEWS of one canvas, one row apart: a batch is handed to clfd_detect_submit as a base
pointer with frame_stride = row_stride, i.e. 8192 distinct frames cost K canvases of pinned host
memory and no host-side copies.

Rank r owns the contiguous chunk sharding.shard_range(n_frames, r, world); there is no collective in
the loop, and ONE gather of the rect lists at the end (sharding.gather_rects).
"""
from __future__ import annotations

import time

import numpy as np

from . import sharding
from .frames import octave_frame

VIEWS = 64          # views per canvas position (consecutive frames, one row apart)
DX = 16             # horizontal shifts


class StreamSource:
    def __init__(self, W: int, H: int, n_canvases: int = 8, pinned: bool = True):
        import torch
        self.W, self.H, self.K = W, H, n_canvases
        self.pitch = (W + DX - 1 + 63) // 64 * 64
        rows = H + VIEWS - 1
        t = torch.empty((n_canvases, rows, self.pitch), dtype=torch.uint8)
        self.canvas = t.pin_memory() if pinned else t
        for k in range(n_canvases):
            self.canvas[k, :, :W + DX - 1] = torch.from_numpy(octave_frame(W + DX - 1, rows, 1000 + k))

    def locate(self, g: int):
        """global frame index -> (canvas, dy, dx)"""
        return (g // VIEWS) % self.K, g % VIEWS, (g // (VIEWS * self.K)) % DX

    def frame(self, g: int) -> np.ndarray:
        k, dy, dx = self.locate(g)
        return self.canvas[k, dy:dy + self.H, dx:dx + self.W].numpy()

    def runs(self, first: int, last: int, max_batch: int):
        """[first, last) cut into batches of consecutive frames that are views of one canvas position:
        yields (g0, n, tensor view of frame g0)"""
        g = first
        while g < last:
            n = min(max_batch, last - g, VIEWS - g % VIEWS)
            k, dy, dx = self.locate(g)
            yield g, n, self.canvas[k, dy:, dx:]
            g += n


def run_stream(det, src: StreamSource, first: int, last: int, max_batch: int):
    """Rank-local part of the stream through clfd_detect_submit / _collect (two batches in flight).
    Returns int32 [n, 6] rects (x, y, w, h, GLOBAL frame, cascade)."""
    out = []
    pending = []
    for g0, n, view in src.runs(first, last, max_batch):
        det.submit_views(view, n, src.pitch, src.pitch)   # frame_stride = row_stride: views one row apart
        pending.append(g0)
        if len(pending) == 2:
            out.append(sharding.rects_to_array(det.collect().rects, frame_offset=pending.pop(0)))
    while pending:
        out.append(sharding.rects_to_array(det.collect().rects, frame_offset=pending.pop(0)))
    return np.concatenate(out, axis=0) if out else np.zeros((0, 6), np.int32)


def sorted_rects(r: np.ndarray) -> np.ndarray:
    """canonical order: frame, cascade, w, h, y, x (the device appends in no particular order)"""
    r = np.asarray(r, np.int32).reshape(-1, 6)
    return r[np.lexsort((r[:, 0], r[:, 1], r[:, 3], r[:, 2], r[:, 5], r[:, 4]))] if len(r) else r


def timed_stream(det, src, n_frames, rank, world, max_batch, barrier, device="cuda"):
    """-> (seconds of this rank incl. the gather, gathered rects on every rank)"""
    first, last = sharding.shard_range(n_frames, rank, world)
    barrier()
    t0 = time.perf_counter()
    local = run_stream(det, src, first, last, max_batch)
    gathered = sharding.gather_rects(local, device=device)
    barrier()
    return time.perf_counter() - t0, gathered
