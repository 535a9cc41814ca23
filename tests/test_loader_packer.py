"""The product's C++ cascade loader / hidden-cascade builder (csrc/haar_xml.cpp,
haar_pack.cpp, reached through the C ABI -- no GPU needed) against the oracle's independent
Python reader + C hidden cascade, and against the census of SURVEY.md Appendix B."""
import os
import tempfile

import numpy as np
import pytest

import clfacedetection_b200 as clfd
import oracle
from conftest import ALL_CASCADES, cascade_path, oracle_cascade

# file: (win_w, win_h, stages, trees, nodes, max trees/stage, max nodes/tree, tilted nodes, 3-rect nodes, stage tree)
CENSUS = {
    "eye": (20, 20, 24, 1066, 1066, 93, 1, 0, 167, 0),
    "eye_tree_eyeglasses": (20, 20, 30, 851, 2553, 47, 3, 577, 295, 0),
    "frontalface_alt": (20, 20, 22, 2135, 2135, 213, 1, 0, 360, 0),
    "frontalface_alt2": (20, 20, 20, 1047, 2094, 109, 2, 0, 347, 0),
    "frontalface_alt_tree": (20, 20, 47, 8468, 8468, 406, 1, 0, 1545, 1),
    "frontalface_default": (24, 24, 25, 2913, 2913, 211, 1, 0, 557, 0),
    "fullbody": (14, 28, 30, 1464, 1464, 107, 1, 201, 227, 0),
    "mcs_nose": (18, 15, 20, 3365, 3365, 377, 1, 990, 557, 0),
    "profileface": (20, 20, 26, 2609, 2609, 195, 1, 0, 415, 0),
    "lefteye_2splits": (20, 20, 20, 366, 732, 33, 2, 185, 58, 0),
    "righteye_2splits": (20, 20, 20, 368, 736, 34, 2, 186, 47, 0),
    "lowerbody": (19, 23, 27, 1221, 1221, 89, 1, 110, 128, 0),
    "upperbody": (22, 18, 30, 2423, 2423, 152, 1, 474, 368, 0),
    "mcs_eyepair_big": (45, 11, 19, 748, 748, 85, 1, 135, 127, 0),
    "mcs_eyepair_small": (22, 5, 17, 860, 860, 133, 1, 76, 177, 0),
    "mcs_lefteye": (18, 12, 14, 1648, 1648, 279, 1, 346, 273, 0),
    "mcs_mouth": (25, 15, 17, 1515, 1515, 218, 1, 223, 295, 0),
    "mcs_righteye": (18, 12, 18, 2942, 2942, 415, 1, 672, 433, 0),
    "mcs_upperbody": (22, 20, 19, 3224, 3224, 334, 1, 657, 495, 0),
}


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_census(name):
    i = clfd.Cascade(cascade_path(name)).info
    got = (i.win_w, i.win_h, i.n_stages, i.n_trees, i.n_nodes, i.max_trees_per_stage, i.max_nodes_per_tree,
           i.n_tilted_nodes, i.n_three_rect_nodes, i.is_tree)
    assert got == CENSUS[name]
    assert bool(i.is_stump_based) == (i.max_nodes_per_tree == 1)
    assert bool(i.has_tilted) == (i.n_tilted_nodes > 0)


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_loader_equals_oracle_reader_bit_for_bit(name):
    ours = clfd.Cascade(cascade_path(name)).arrays()
    ref = oracle.load_cascade_xml(cascade_path(name))
    for k in ("st_ntrees", "st_thr", "st_parent", "st_next", "tr_nnodes", "nd_tilted", "nd_rect", "nd_weight",
              "nd_thr", "nd_left", "nd_right", "alpha"):
        a, b = ours[k], getattr(ref, k)
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), k


@pytest.mark.parametrize("name", ALL_CASCADES)
def test_hidden_cascade_equals_oracle(name):
    """scale-1 weights, kept-rect counts, biased stage thresholds, two_rects (tempcv.cpp:419,453-458,752-760)"""
    w, nr, thr, two = clfd.Cascade(cascade_path(name)).hidden()
    ow, onr, othr, otwo, _ = oracle_cascade(name).hidden()
    assert w.tobytes() == ow.tobytes()
    assert np.array_equal(nr, onr) and thr.tobytes() == othr.tobytes() and np.array_equal(two, otwo)


def test_from_arrays_round_trip_and_child_links():
    flat = oracle.load_cascade_xml(cascade_path("frontalface_alt_tree"))
    c = clfd.Cascade(flat=flat)
    a = c.arrays()
    assert c.info.is_tree == 1
    child = a["st_child"]
    # SURVEY Appendix B topology: 0-4 linear, stage 4 -> child 5 (next 6); odd chain ends at 39
    assert child[:5].tolist() == [1, 2, 3, 4, 5] and a["st_next"][5] == 6 and child[39] == -1
    assert c.info.dense_stages == 47  # a stage tree of stumps is walked by the tile kernel (5 linear stages + 42 routed)


def test_dense_prefix_rules():
    assert clfd.Cascade(cascade_path("frontalface_alt")).info.dense_stages >= 8
    assert clfd.Cascade(cascade_path("frontalface_alt2")).info.dense_stages == 20    # 2-node trees: per-window node state
    assert clfd.Cascade(cascade_path("fullbody")).info.dense_stages == 30            # tilted stumps: second smem tile
    assert clfd.Cascade(cascade_path("mcs_nose")).info.dense_stages == 20
    assert clfd.Cascade(cascade_path("eye_tree_eyeglasses")).info.dense_stages == 30  # 3-node trees + tilted


def _mutate(name, old, new, count=1):
    text = open(cascade_path(name), encoding="latin-1").read()
    assert old in text
    f = tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False, encoding="latin-1")
    f.write(text.replace(old, new, count))
    f.close()
    return f.name


@pytest.mark.parametrize("old,new,msg", [
    ("<size>20 20</size>", "<size>20</size>", "size node is not a valid sequence"),
    ("<_>3 7 14 4 -1.</_>", "<_>3 7 19 4 -1.</_>", "width must be positive integer and (x + width) must not exceed window width"),
    ("<_>3 7 14 4 -1.</_>", "<_>3 7 14 4 -1</_>", "weight must be real number"),
    ("<tilted>0</tilted>", "", "tilted must be 0 or 1"),
    ("<left_val>", "<left_vall>", None),
    ("<parent>-1</parent>", "<parent>99</parent>", "parent must be integer number. (stage 0)"),
    ("<stage_threshold>", "<stage_thresholdx>", None),
])
def test_malformed_cascades_fail_loudly(old, new, msg):
    path = _mutate("frontalface_alt", old, new)
    try:
        with pytest.raises(clfd.ClfdError) as e:
            clfd.Cascade(path)
        assert e.value.status == -3
        if msg:
            assert msg in str(e.value)
        with pytest.raises(Exception):
            oracle.load_cascade_xml(path)
    finally:
        os.unlink(path)


def test_missing_and_foreign_files():
    with pytest.raises(clfd.ClfdError) as e:
        clfd.Cascade("/nonexistent/haarcascade.xml")
    assert e.value.status == -4
    f = tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False)
    f.write("<?xml version='1.0'?><opencv_storage><cascade type_id=\"opencv-ml-svm\"></cascade></opencv_storage>")
    f.close()
    with pytest.raises(clfd.ClfdError) as e:
        clfd.Cascade(f.name)
    assert "opencv-haar-classifier" in str(e.value)
    open(f.name, "w").write("<?xml version='1.0'?><opencv_storage><cascade type_id=\"opencv-cascade-classifier\"></cascade></opencv_storage>")
    with pytest.raises(clfd.ClfdError) as e:   # new format, but empty
        clfd.Cascade(f.name)
    assert "stageType" in str(e.value)
    os.unlink(f.name)


def _write_new_format(arrays, win, path, tilted_tag=True):
    """old-format arrays -> an 'opencv-cascade-classifier' (HAAR, BOOST) file, the way OpenCV's
    own converter lays it out: one feature per node, internalNodes quadruples, leafValues."""
    a = arrays
    out = ["<?xml version=\"1.0\"?>", "<opencv_storage>", "<cascade type_id=\"opencv-cascade-classifier\"><stageType>BOOST</stageType>",
           "  <featureType>HAAR</featureType>", f"  <height>{win[1]}</height>", f"  <width>{win[0]}</width>",
           f"  <stageNum>{len(a['st_ntrees'])}</stageNum>", "  <stages>"]
    t = n = al = 0
    feats = []
    for s, nt in enumerate(a["st_ntrees"]):
        out += ["    <_>", f"      <maxWeakCount>{nt}</maxWeakCount>",
                f"      <stageThreshold>{float(a['st_thr'][s])!r}</stageThreshold>", "      <weakClassifiers>"]
        for _ in range(nt):
            cnt = int(a["tr_nnodes"][t])
            quad = []
            for k in range(cnt):
                quad.append(f"{int(a['nd_left'][n])} {int(a['nd_right'][n])} {len(feats)} {float(a['nd_thr'][n])!r}")
                feats.append(n)
                n += 1
            leaves = " ".join(repr(float(v)) for v in a["alpha"][al:al + cnt + 1])
            al += cnt + 1
            out += ["        <_>", "          <internalNodes>", "            " + " ".join(quad) + "</internalNodes>",
                    "          <leafValues>", "            " + leaves + "</leafValues></_>"]
            t += 1
        out += ["      </weakClassifiers></_>"]
    out += ["  </stages>", "  <features>"]
    rect = np.asarray(a["nd_rect"]).reshape(-1, 3, 4)
    wt = np.asarray(a["nd_weight"]).reshape(-1, 3)
    for n in feats:
        out += ["    <_>", "      <rects>"]
        for k in range(3):
            if rect[n, k, 2] == 0 or rect[n, k, 3] == 0:
                continue
            x, y, w, h = (int(v) for v in rect[n, k])
            out += ["        <_>", f"          {x} {y} {w} {h} {float(wt[n, k])!r}</_>"]
        out[-1] += "</rects>"
        if a["nd_tilted"][n] and tilted_tag:
            out += ["      <tilted>1</tilted>"]
        out[-1] += "</_>"
    out += ["  </features></cascade>", "</opencv_storage>", ""]
    open(path, "w").write("\n".join(out))


@pytest.mark.parametrize("name", ["frontalface_alt", "frontalface_alt2", "fullbody"])
def test_new_format_cascade_equals_old_format(name, tmp_path):
    """the 'opencv-cascade-classifier' (HAAR) on-disk format carries the same
    content; the reader must produce the same arrays, hidden cascade and packing as the old file."""
    old = clfd.Cascade(cascade_path(name))
    p = str(tmp_path / f"new_{name}.xml")
    _write_new_format(old.arrays(), (old.info.win_w, old.info.win_h), p)
    new = clfd.Cascade(p)
    a, b = old.arrays(), new.arrays()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    for x, y in zip(old.hidden(), new.hidden()):
        assert np.array_equal(x, y)
    for f in ("n_stages", "n_trees", "n_nodes", "is_tree", "is_stump_based", "has_tilted", "dense_stages", "dense_stumps"):
        assert getattr(old.info, f) == getattr(new.info, f), f


def test_new_format_rejects_lbp_and_broken_trees(tmp_path):
    old = clfd.Cascade(cascade_path("eye"))
    p = str(tmp_path / "x.xml")
    _write_new_format(old.arrays(), (20, 20), p)
    text = open(p).read()
    for bad, msg in [(text.replace("<featureType>HAAR", "<featureType>LBP"), "LBP"),
                     (text.replace("<stageType>BOOST", "<stageType>TREE"), "stageType"),
                     (text.replace("</leafValues>", " 0.5</leafValues>", 1), "leaf values"),
                     (text.replace("<width>20</width>", "<width>10</width>"), "does not fit")]:
        q = str(tmp_path / "bad.xml")
        open(q, "w").write(bad)
        with pytest.raises(clfd.ClfdError) as e:
            clfd.Cascade(q)
        assert msg in str(e.value)


def test_new_format_files_shipped_with_cv2_if_present():
    cv2 = pytest.importorskip("cv2")
    import os
    d = getattr(getattr(cv2, "data", None), "haarcascades", None)
    if not d or not os.path.exists(os.path.join(d, "haarcascade_eye.xml")):
        pytest.skip("cv2 ships no cascade data here")
    for name in ("eye", "frontalface_default", "profileface"):
        old, new = clfd.Cascade(cascade_path(name)), clfd.Cascade(os.path.join(d, f"haarcascade_{name}.xml"))
        a, b = old.arrays(), new.arrays()
        assert all(np.array_equal(a[k], b[k]) for k in a)


def test_packer_tile_eligibility_of_synthetic_tree_shapes(monkeypatch):
    """Host-side packing decisions need no GPU: trees of up to four nodes (children after their parent)
    are tile-evaluated with padded node records; the test hooks switch each tile-kernel extension off."""
    from test_gpu_clod import _mixed_tree_cascade
    flat = _mixed_tree_cascade()
    c = clfd.Cascade(flat=flat)
    assert c.info.max_nodes_per_tree == 4 and c.info.dense_stages == c.info.n_stages
    assert c.info.dense_stumps == 4 * c.info.n_trees   # every tree padded to 4 records
    monkeypatch.setenv("CLFD_NO_NODE_TILES", "1")
    assert clfd.Cascade(flat=flat).info.dense_stages == 0
    assert clfd.Cascade(cascade_path("frontalface_alt2")).info.dense_stages == 0
    monkeypatch.delenv("CLFD_NO_NODE_TILES")
    monkeypatch.setenv("CLFD_NO_TREE_TILES", "1")
    assert clfd.Cascade(cascade_path("frontalface_alt_tree")).info.dense_stages == 5   # the linear prefix only
    monkeypatch.delenv("CLFD_NO_TREE_TILES")
    monkeypatch.setenv("CLFD_NO_TILTED_TILE", "1")
    assert clfd.Cascade(cascade_path("fullbody")).info.dense_stages == 2               # first tilted stump in stage 2
    assert clfd.Cascade(cascade_path("mcs_nose")).info.dense_stages == 0
    # a child that precedes its parent is left to the generic kernels
    bad = _mixed_tree_cascade()
    import numpy as _np
    nn = _np.array(bad.tr_nnodes)
    first4 = int(_np.flatnonzero(nn == 4)[0])
    n0 = int(nn[:first4].sum())
    left, right = _np.array(bad.nd_left).copy(), _np.array(bad.nd_right).copy()
    # 4-node shape (0,1),(2,3),(-1,-2),(-3,-4): make node 2 point back to node 1 -> cycle-free but out of order
    left[n0 + 2] = 1
    import dataclasses
    bad = dataclasses.replace(bad, nd_left=left, nd_right=right)
    monkeypatch.delenv("CLFD_NO_TILTED_TILE")
    try:
        info = clfd.Cascade(flat=bad).info
        assert info.dense_stages == 0   # stage 0 already holds a 4-node tree
    except clfd.ClfdError:
        pass                            # or rejected outright as a broken tree
