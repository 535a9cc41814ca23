"""Fixed-seed synthetic grayscale frames.

S1 "octave": value noise -- a sum over octaves o = 2..8 of a (H>>o + 2) x (W>>o + 2)
uniform [0,1) grid upsampled bicubically to W x H, amplitudes 1, .6, .36, .216 (o <= 5) then
x1.4 per octave, min-max normalised to 0..255.  S2 "uniform": i.i.d. uniform 0..255.
Seed = 0xC0FFEE + frame_index, numpy PCG64.
"""
from __future__ import annotations

import numpy as np

SEED0 = 0xC0FFEE


def _upsample_bicubic(grid: np.ndarray, W: int, H: int) -> np.ndarray:
    """Separable Catmull-Rom upsample in float64; pure numpy so the frames do not depend
    on any image library version."""
    def axis(src, n_out):
        n_in = src.shape[-1]
        pos = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5
        i0 = np.floor(pos).astype(np.int64)
        t = pos - i0
        idx = np.clip(i0[None, :] + np.arange(-1, 3)[:, None], 0, n_in - 1)
        w = np.stack([((-t + 2) * t - 1) * t * 0.5, ((3 * t - 5) * t * t + 2) * 0.5,
                      ((-3 * t + 4) * t + 1) * t * 0.5, (t - 1) * t * t * 0.5])
        return (src[..., idx] * w).sum(axis=-2)
    tmp = axis(grid, W)
    return axis(tmp.T, H).T


def octave_frame(W: int, H: int, index: int = 0) -> np.ndarray:
    rng = np.random.default_rng(SEED0 + index)
    acc = np.zeros((H, W), np.float64)
    amp = 1.0
    for o in range(2, 9):
        gh, gw = (H >> o) + 2, (W >> o) + 2
        grid = rng.random((gh, gw))
        acc += amp * _upsample_bicubic(grid, W, H)
        amp = amp * 0.6 if o < 5 else amp * 1.4
    lo, hi = acc.min(), acc.max()
    return np.clip(np.rint((acc - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def uniform_frame(W: int, H: int, index: int = 0) -> np.ndarray:
    rng = np.random.default_rng(SEED0 + index)
    return rng.integers(0, 256, size=(H, W), dtype=np.uint8)


def make_frames(kind: str, W: int, H: int, count: int, first: int = 0) -> np.ndarray:
    gen = {"octave": octave_frame, "S1": octave_frame, "uniform": uniform_frame, "S2": uniform_frame}[kind]
    return np.stack([gen(W, H, first + i) for i in range(count)])
