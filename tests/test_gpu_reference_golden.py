"""The CUDA path against vectors recorded from the REFERENCE'S OWN compiled functions
(tests/golden/reference_tempcv.npz, made by tests/golden/make_ref_golden.py from
tempcv.cpp:40-1516): no oracle in between.  All 19 cascade files the reference ships."""
import os

import numpy as np
import pytest

import clfacedetection_b200 as clfd
from clfacedetection_b200.frames import octave_frame, uniform_frame
from conftest import ALL_CASCADES, cascade_path

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tempcv.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _sorted(r):
    r = np.asarray(r, np.int32).reshape(-1, 4)
    return r[np.lexsort((r[:, 0], r[:, 1], r[:, 2]))] if len(r) else r


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_pyramid_mode_equals_reference_vectors(gpu_ctx, gold, name):
    """raw rect sets of cvHaarDetectObjectsForROC(CV_HAAR_SCALE_IMAGE), the grouped output, the
    reject-level output, and at the unscaled level every window's return code of
    cvRunHaarClassifierCascadeSum"""
    frames = np.stack([octave_frame(320, 240, 21), uniform_frame(320, 240, 22)])
    cas = clfd.Cascade(cascade_path(name))
    det = clfd.Detector(gpu_ctx, cas, 320, 240, max_batch=2, scale_factor=1.2, want_codes=True)
    res = det.detect(frames)
    codes = det.codes(0, 2)
    lv0 = det.levels(0)[0]
    n_stages = cas.info.n_stages
    for fi in range(2):
        raw = res.frame_rects(fi)
        assert np.array_equal(raw, _sorted(gold[f"{name}/si_{fi}"])), f"{name} frame {fi}: raw rects"
        # level 0 = the frame itself on the ystep-2 grid = every other window of the golden map
        ref_map = gold[f"{name}/lvl_{fi}"].astype(np.int32)[::2, ::2]
        assert ref_map.shape == (lv0.ny, lv0.nx)
        c = codes[fi][:lv0.nx * lv0.ny].astype(np.int32).reshape(lv0.ny, lv0.nx)
        exp = (c & 1) if cas.info.is_tree else np.where(c == n_stages, 1, -c)
        assert np.array_equal(exp, ref_map), f"{name} frame {fi}: {int((exp != ref_map).sum())} window results differ"
        # grouped (AgroupRectangles, minNeighbors 2) in the reference's output order
        src = gold[f"{name}/si_{fi}"].reshape(-1, 4)
        g, w = clfd.group_rectangles(src, 2)
        assert np.array_equal(np.concatenate([g, w[:, None]], 1), gold[f"{name}/sig_{fi}"].reshape(-1, 5))
    # reject levels of the last frame... the API reports the whole last batch, frame by frame
    r, lv, wt = det.reject_levels(0)
    for fi in range(2):
        roc = gold[f"{name}/roc_{fi}"].reshape(-1, 6)
        m = r["frame"] == fi
        got = np.stack([r["x"][m], r["y"][m], r["w"][m], r["h"][m]], 1).astype(np.int32).reshape(-1, 4)
        assert np.array_equal(got, roc[:, :4].astype(np.int32)), f"{name} frame {fi}: ROC candidates"
        assert np.array_equal(lv[m], roc[:, 4].astype(np.int32))
        assert wt[m].tobytes() == np.ascontiguousarray(roc[:, 5]).tobytes(), f"{name} frame {fi}: stage sums"
    det.close()


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_scale_cascade_mode_equals_reference_vectors(gpu_ctx, gold, name):
    """raw rect sets of cvHaarDetectObjectsForROC(flags = 0): scaled features + skip rule"""
    frames = np.stack([octave_frame(320, 240, 21), uniform_frame(320, 240, 22)])
    cas = clfd.Cascade(cascade_path(name))
    det = clfd.Detector(gpu_ctx, cas, 320, 240, max_batch=2, scale_factor=1.2, scale_cascade=True)
    res = det.detect(frames)
    for fi in range(2):
        assert np.array_equal(res.frame_rects(fi), _sorted(gold[f"{name}/sc_{fi}"])), f"{name} frame {fi}"
    det.close()
