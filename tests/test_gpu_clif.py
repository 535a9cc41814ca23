"""GPU parity: pyramid resize and integral images are BIT-EXACT against the oracle
(integer work).  Called through the C ABI."""
import numpy as np
import pytest

import oracle
from clfacedetection_b200.frames import octave_frame, uniform_frame

pytestmark = pytest.mark.gpu


def _check_integral(ctx, img, tilted):
    s, q, t = ctx.integral(img, tilted=tilted)
    os_, oq, ot = oracle.integral(img, tilted=tilted)
    assert np.array_equal(s, os_)
    assert np.array_equal(q, oq.astype(np.uint64))
    if tilted:
        assert np.array_equal(t, ot)


@pytest.mark.parametrize("shape", [(1, 1), (5, 7), (31, 33), (257, 64), (640, 480), (1333, 750), (1920, 1080)])
@pytest.mark.parametrize("tilted", [False, True])
def test_integral_uniform_noise(gpu_ctx, shape, tilted):
    w, h = shape
    _check_integral(gpu_ctx, uniform_frame(w, h, 3), tilted)


def test_integral_4k_all_255_needs_64_bit(gpu_ctx):
    img = np.full((2160, 3840), 255, np.uint8)
    s, q, _ = gpu_ctx.integral(img)
    assert int(q[-1, -1]) == 255 * 255 * 3840 * 2160 > 2 ** 32
    assert int(s[-1, -1]) == 255 * 3840 * 2160
    _check_integral(gpu_ctx, img, False)


def test_integral_all_zero_and_strided(gpu_ctx):
    _check_integral(gpu_ctx, np.zeros((480, 640), np.uint8), True)
    big = uniform_frame(700, 300, 9)
    view = big[:, :611]          # row stride > width
    s, q, _ = gpu_ctx.integral(np.ascontiguousarray(view))
    os_, oq, _ = oracle.integral(np.ascontiguousarray(view))
    assert np.array_equal(s, os_) and np.array_equal(q, oq.astype(np.uint64))


@pytest.mark.parametrize("src,dst", [((640, 480), (533, 400)), ((640, 480), (640, 480)), ((640, 480), (42, 31)),
                                     ((1920, 1080), (1600, 900)), ((1920, 1080), (926, 521)),
                                     ((1920, 1080), (42, 23)), ((3840, 2160), (3200, 1800)), ((37, 29), (20, 20))])
def test_resize_bit_exact(gpu_ctx, src, dst):
    img = octave_frame(src[0], src[1], 1) if src[0] >= 640 else uniform_frame(src[0], src[1], 1)
    got = gpu_ctx.resize(img, dst[0], dst[1])
    ref = oracle.resize_linear(img, dst[0], dst[1])
    assert np.array_equal(got, ref)


def test_resize_uniform_noise_1080p_levels(gpu_ctx):
    img = uniform_frame(1920, 1080, 5)
    f = 1.0
    for _ in range(6):
        f *= 1.2
        dw, dh = int(np.rint(1920 / f)), int(np.rint(1080 / f))
        assert np.array_equal(gpu_ctx.resize(img, dw, dh), oracle.resize_linear(img, dw, dh))


def test_integral_linearity_property_1080p(gpu_ctx):
    """size-independent property: integral(a) + integral(b) == integral(a+b) for a+b <= 255"""
    a = (uniform_frame(1920, 1080, 1) // 2).astype(np.uint8)
    b = (uniform_frame(1920, 1080, 2) // 2).astype(np.uint8)
    sa, _, _ = gpu_ctx.integral(a)
    sb, _, _ = gpu_ctx.integral(b)
    sab, _, _ = gpu_ctx.integral((a + b).astype(np.uint8))
    assert np.array_equal(sa.astype(np.int64) + sb, sab.astype(np.int64))
