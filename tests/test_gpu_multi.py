"""an N-GPU sharded frame stream gives, frame by frame, the rects of a single-GPU run
(and those are the oracle's).  N = the GPUs visible on the box; skipped with fewer than two."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import cascade_path, oracle_cascade

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "multi_gpu_stream.py")
N_FRAMES, W, H, NAME = 150, 480, 360, "eye"   # 150 frames: uneven shards, runs that cross canvas boundaries


def _run(world, out):
    if world == 1:
        cmd = [sys.executable, SCRIPT]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29731", SCRIPT]
    r = subprocess.run(cmd + [out, str(N_FRAMES), str(W), str(H), NAME], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)["rects"]


def test_single_gpu_stream_equals_oracle(tmp_path):
    """the stream machinery itself: overlapping canvas views, submit / collect, global frame indices"""
    from clfacedetection_b200 import stream
    got = _run(1, str(tmp_path / "one.npz"))
    src = stream.StreamSource(W, H, n_canvases=2, pinned=False)
    oc = oracle_cascade(NAME)
    want = []
    for g in list(range(0, 6)) + [63, 64, 65, 127, 128, 149]:
        r, _, _, _, _ = oc.detect(np.ascontiguousarray(src.frame(g)), 1.2, want_codes=False)
        mine = got[got[:, 4] == g][:, :4]
        a = r[np.lexsort((r[:, 0], r[:, 1], r[:, 3], r[:, 2]))] if len(r) else r
        assert np.array_equal(mine, a), g
        want.append(len(r))
    assert sum(want) > 0 and got[:, 4].max() < N_FRAMES


def test_n_gpu_stream_equals_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    one = _run(1, str(tmp_path / "one.npz"))
    many = _run(n, str(tmp_path / "many.npz"))
    assert np.array_equal(one, many)
    assert len(one) > 0
