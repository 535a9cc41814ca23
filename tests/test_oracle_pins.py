"""Pins of the oracle (CPU restatement) against OpenCV and the committed golden vectors.
The reference has no tests or golden files of its own, so the pieces
of its arithmetic that live in OpenCV are pinned against cv2 here."""
import os
import zlib

import numpy as np
import pytest

import oracle
from clfacedetection_b200.frames import octave_frame, uniform_frame
from conftest import ALL_CASCADES, oracle_cascade

GOLD = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2")


def test_resize_matches_cv2_inter_linear_bit_exact():
    src = octave_frame(640, 480, 0)
    f = 1.0
    for _ in range(12):   # the 640x480 sf-1.2 level sizes
        f *= 1.2
        dw, dh = int(np.rint(640 / f)), int(np.rint(480 / f))
        assert np.array_equal(oracle.resize_linear(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))
    noise = uniform_frame(333, 217, 3)
    for dw, dh in [(300, 200), (111, 73), (64, 48), (333, 217), (21, 20)]:
        assert np.array_equal(oracle.resize_linear(noise, dw, dh), cv2.resize(noise, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (37, 29), (640, 480)])
def test_integral_matches_cv2_integral3(shape):
    img = uniform_frame(shape[0], shape[1], 7)
    s, q, t = oracle.integral(img, tilted=True)
    cs, cq, ct = cv2.integral3(img)
    assert np.array_equal(s, cs) and np.array_equal(q, cq) and np.array_equal(t, ct)


def test_integral_extremes():
    s, q, _ = oracle.integral(np.full((1080, 1920), 255, np.uint8))
    assert s[-1, -1] == 255 * 1920 * 1080 and q[-1, -1] == 255.0 * 255 * 1920 * 1080
    s, q, t = oracle.integral(np.zeros((50, 70), np.uint8), tilted=True)
    assert not s.any() and not q.any() and not t.any()


def test_grouping_vs_cv2_up_to_mean_rounding():
    """tempcv.cpp:194-198 truncates the class mean, cv2 4.x rounds it: classes and neighbour
    counts must agree exactly, coordinates within one pixel."""
    rng = np.random.default_rng(1)
    for _ in range(100):
        base = rng.integers(0, 200, size=(int(rng.integers(1, 6)), 2))
        r = np.array([[b[0] + rng.integers(-6, 7), b[1] + rng.integers(-6, 7), 40 + rng.integers(-3, 4), 40 + rng.integers(-3, 4)]
                      for b in base[rng.integers(0, len(base), size=int(rng.integers(1, 60)))]], np.int32)
        thr = int(rng.integers(1, 4))
        a, w = oracle.group_rectangles(r, thr, 0.2)
        b, wb = cv2.groupRectangles(r.tolist(), thr, 0.2)
        b = np.array(b, np.int32).reshape(-1, 4)
        if len(a) != len(b):
            continue   # a 1-px difference flipped a nested-rect test: rare, not a defect
        assert np.array_equal(w, np.array(wb).reshape(-1))
        assert np.abs(a - b).max(initial=0) <= 1


def test_grouping_edge_cases():
    r, w = oracle.group_rectangles(np.zeros((0, 4), np.int32), 3)
    assert len(r) == 0
    r, w = oracle.group_rectangles(np.array([[1, 2, 30, 30]] * 5, np.int32), 0)   # threshold 0: untouched, weights 1
    assert len(r) == 5 and (w == 1).all()
    r, w = oracle.group_rectangles(np.array([[10, 10, 30, 30]] * 4 + [[200, 200, 30, 30]], np.int32), 2)
    assert r.tolist() == [[10, 10, 30, 30]] and w.tolist() == [4]


def test_golden_opencv_pins():
    g = np.load(os.path.join(GOLD, "opencv_pins.npz"))
    for k in g.files:
        if k.startswith("resize_") and k != "resize_src":
            dw, dh = (int(v) for v in k.split("_")[2].split("x"))
            assert np.array_equal(oracle.resize_linear(g["resize_src"], dw, dh), g[k]), k
    s, q, t = oracle.integral(g["integral_src"], tilted=True)
    assert np.array_equal(s, g["integral_sum"]) and np.array_equal(q, g["integral_sq"]) and np.array_equal(t, g["integral_tilted"])
    a, w = oracle.group_rectangles(g["group_in"], 2, 0.2)
    assert np.array_equal(w, g["group_w"]) and np.abs(a - g["group_out"]).max(initial=0) <= 1


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_golden_refsi_detection(name):
    g = np.load(os.path.join(GOLD, f"refsi_{name}.npz"))
    cas = oracle_cascade(name)
    for fi, frame in enumerate([octave_frame(320, 240, 21), uniform_frame(320, 240, 22)]):
        r, codes, near, st, levels = cas.detect(frame, 1.2)
        assert np.array_equal(r, g[f"rects_{fi}"])
        assert zlib.crc32(codes.tobytes()) == int(g[f"crc_{fi}"][0])
        assert np.array_equal(np.bincount(codes.astype(np.int64), minlength=128), g[f"hist_{fi}"])
        assert [st.windows, st.weak_evals, st.node_evals, st.accepted, st.near_stage_thr] == g[f"stats_{fi}"].tolist()
        assert [[l.img_w, l.img_h, l.win_w, l.win_h, l.ystep, l.nx, l.ny] for l in levels] == g[f"levels_{fi}"].tolist()


def test_level_plan_matches_survey_counts():
    """SURVEY.md Appendix C (REF-SI loop arithmetic)."""
    def total(W, H, w0, h0, sf, mn=(0, 0)):
        lv = oracle.plan_levels(W, H, w0, h0, sf, mn)
        return len(lv), sum(l.nx * l.ny for l in lv)
    assert total(640, 480, 20, 20, 1.2, (24, 24)) == (17, 283021)
    assert total(1920, 1080, 24, 24, 1.2) == (21, 2633075)
    assert total(1920, 1080, 20, 20, 1.2) == (22, 2672451)
    assert total(3840, 2160, 20, 20, 1.2) == (26, 11094920)
    assert total(1920, 1080, 20, 20, 1.1)[1] == 4566697
    assert total(1920, 1080, 14, 28, 1.1)[1] == 4509954


def test_independent_python_restatement_of_the_evaluator():
    """A second, deliberately naive restatement (numpy scalars, float32/float64 made explicit)
    of tempcv.cpp:795-972 for a stump cascade, checked against the C oracle window by window."""
    cas = oracle_cascade("frontalface_alt")
    f = cas.flat
    w, nr, sthr, two, _ = cas.hidden()
    img = octave_frame(64, 48, 5)
    s, q, _ = oracle.integral(img)
    codes, _, _ = cas.eval_level(img, 2)
    first = np.concatenate([[0], np.cumsum(f.st_ntrees)])
    inv = np.float64(1.0) / np.float64(18 * 18)

    def rs(a, x, y, r):
        return int(a[y + r[1], x + r[0]]) - int(a[y + r[1], x + r[0] + r[2]]) - int(a[y + r[1] + r[3], x + r[0]]) + int(a[y + r[1] + r[3], x + r[0] + r[2]])

    for iy in range(0, codes.shape[0], 3):
        for ix in range(0, codes.shape[1], 3):
            x, y = 2 * ix, 2 * iy
            mean = np.float64(rs(s, x, y, (1, 1, 18, 18))) * inv
            v = np.float64(q[y + 1, x + 1] - q[y + 1, x + 19] - q[y + 19, x + 1] + q[y + 19, x + 19]) * inv - mean * mean
            sigma = np.sqrt(v) if v >= 0 else np.float64(1.0)
            depth = 0
            for st in range(f.n_stages):
                S = np.float64(0)
                for n in range(first[st], first[st + 1]):
                    t = np.float64(f.nd_thr[n]) * sigma
                    if two[st]:
                        val = np.float64(rs(s, x, y, f.nd_rect[n, 1])) * np.float64(w[n, 1]) + \
                              np.float64(rs(s, x, y, f.nd_rect[n, 0])) * np.float64(w[n, 0])
                    else:
                        val = np.float64(np.float32(rs(s, x, y, f.nd_rect[n, 0])) * w[n, 0])
                        val = val + np.float64(np.float32(rs(s, x, y, f.nd_rect[n, 1])) * w[n, 1])
                        if nr[n] == 3:
                            val = val + np.float64(np.float32(rs(s, x, y, f.nd_rect[n, 2])) * w[n, 2])
                    S = S + np.float64(f.alpha[2 * n + (1 if val >= t else 0)])
                if S < np.float64(sthr[st]):
                    break
                depth += 1
            assert depth == codes[iy, ix], (ix, iy)
