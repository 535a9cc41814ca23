"""The reference-facing C++ API (include/clif.h, clod.h over the C ABI) on the GPU: the demo
client examples/clod_demo and -- when it was built in the dev container -- the reference's own
main.cpp, compiled byte for byte against the shim (tests/_build/ref_main)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import cascade_path, oracle_cascade

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5 %d %d 255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


@pytest.mark.parametrize("min_neighbors", [0, 2])
def test_clod_demo_matches_oracle(tmp_path, min_neighbors):
    from clfacedetection_b200.frames import octave_frame
    img = octave_frame(640, 480, 0)
    pgm = str(tmp_path / "frame.pgm")
    _write_pgm(pgm, img)
    out = subprocess.run([os.path.join(ROOT, "examples", "clod_demo"), cascade_path("frontalface_default"), pgm, "1.2",
                          str(min_neighbors), "24", "24"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    # clifIntegral of the (gray -> BGR -> gray) frame: bottom-right corner of sum and squared sum.
    # gray replicated to BGR converts back to itself: (v*1868 + v*9617 + v*4899 + 8192) >> 14 == v
    s, q, _ = oracle.integral(img)
    assert lines[0].split() == ["integral", str(int(s[-1, -1])), str(int(q[-1, -1]))]
    n = int(lines[1].split()[1])
    got = np.array([[int(v) for v in ln.split()[:4]] for ln in lines[2:2 + n]], np.int32).reshape(-1, 4)
    rects, _, _, _, _ = oracle_cascade("frontalface_default").detect(img, 1.2, (24, 24))
    if min_neighbors:
        rects, weights = oracle.group_rectangles(rects, min_neighbors)
        w_got = [float(ln.split()[4]) for ln in lines[2:2 + n]]
        assert w_got == [float(w) for w in weights]
        assert np.array_equal(got, rects)
    else:
        key = lambda r: r[np.lexsort((r[:, 0], r[:, 1], r[:, 2]))] if len(r) else r
        assert np.array_equal(key(got), key(rects))


def test_clod_demo_scale_cascade_mode_matches_ref_sc(tmp_path):
    """clodSetDetectionMode(data, 1): the semantics of the cvHaarDetectObjects call of main.cpp:145."""
    from clfacedetection_b200.frames import octave_frame
    img = octave_frame(640, 480, 3)
    pgm = str(tmp_path / "frame.pgm")
    _write_pgm(pgm, img)
    out = subprocess.run([os.path.join(ROOT, "examples", "clod_demo"), cascade_path("frontalface_alt"), pgm, "1.2", "0", "0", "0", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    n = int(lines[1].split()[1])
    got = np.array([[int(v) for v in ln.split()[:4]] for ln in lines[2:2 + n]], np.int32).reshape(-1, 4)
    rects, _, _, _ = oracle_cascade("frontalface_alt").detect_sc(img, 1.2)
    key = lambda r: r[np.lexsort((r[:, 0], r[:, 1], r[:, 2]))] if len(r) else r
    assert len(rects) > 0 and np.array_equal(key(got), key(rects))


def test_reference_main_runs_unchanged():
    exe = os.path.join(ROOT, "tests", "_build", "ref_main")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/ref_main is built by the CPU suite where /root/reference is mounted")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    for label in ("OpenCV:", "OpenCL (optimized):", "OpenCL (per-stage):"):
        assert label in out.stdout
