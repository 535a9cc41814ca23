"""Oracle-side reader for OpenCV old-format Haar cascades (``opencv-haar-classifier``).

TEST INFRASTRUCTURE ONLY -- the product has its own C++ loader
(clfacedetection_b200/csrc/haar_xml.cpp); this independent Python restatement exists so the
two can be checked against each other.

Restates icvReadHaarClassifier (/root/reference/CLFaceDetection/tempcv.cpp:1750-2089):
``<size>``, ``<stages>/<_>/<trees>/<_>/<_>{feature{rects,tilted},threshold,
left_val|left_node,right_val|right_node}``, ``stage_threshold``, ``parent``, ``next``.
Reals are parsed as double and cast to float (tempcv.cpp:1932,1958,1995,2033,2054); leaf
values are numbered in the order they are met, left before right (tempcv.cpp:1994-1995,
2032-2033).  The seven ``mcs_*`` files open with a ``<!-----`` comment that strict XML
parsers reject, so everything before ``<opencv_storage>`` is dropped and comments are removed
with a tolerant regex first (loader notes).
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np


@dataclass
class FlatCascade:
    name: str
    win_w: int
    win_h: int
    st_ntrees: np.ndarray   # int32 [S]
    st_thr: np.ndarray      # float32 [S]  (XML value, bias NOT applied)
    st_parent: np.ndarray   # int32 [S]
    st_next: np.ndarray     # int32 [S]
    tr_nnodes: np.ndarray   # int32 [T]
    nd_tilted: np.ndarray   # int32 [N]
    nd_rect: np.ndarray     # int32 [N,3,4] x,y,w,h
    nd_weight: np.ndarray   # float32 [N,3]
    nd_thr: np.ndarray      # float32 [N]
    nd_left: np.ndarray     # int32 [N]  (>0 node index, <=0 leaf -idx)
    nd_right: np.ndarray    # int32 [N]
    alpha: np.ndarray       # float32 [N+T]

    @property
    def n_stages(self) -> int:
        return int(self.st_ntrees.size)

    @property
    def n_trees(self) -> int:
        return int(self.tr_nnodes.size)

    @property
    def n_nodes(self) -> int:
        return int(self.nd_thr.size)


class CascadeFormatError(ValueError):
    pass


_INT_TOKEN = re.compile(r"^[+-]?\d+$")


def _f32(text: str) -> np.float32:
    """CvFileStorage types a bare integer token as INT, and the reference then rejects it where
    it wants CV_NODE_IS_REAL (tempcv.cpp:1925,1952,1981,2019,2049); reals carry '.' or an exponent."""
    tok = text.strip()
    if _INT_TOKEN.match(tok):
        raise CascadeFormatError(f"value must be real number, got integer token {tok!r}")
    return np.float32(float(tok))


def load_cascade_xml(path: str) -> FlatCascade:
    with open(path, "r", encoding="latin-1") as fh:
        text = fh.read()
    start = text.find("<opencv_storage>")
    if start < 0:
        raise CascadeFormatError(f"{path}: no <opencv_storage> element")
    body = re.sub(r"<!--.*?-->", "", text[start:], flags=re.S)
    root = ET.fromstring(body)
    node = None
    for child in root:
        if child.get("type_id") == "opencv-haar-classifier":
            node = child
            break
    if node is None:
        raise CascadeFormatError(f"{path}: no opencv-haar-classifier node")

    size = node.find("size")
    if size is None or len(size.text.split()) != 2:
        raise CascadeFormatError("size node is not a valid sequence.")
    win_w, win_h = (int(v) for v in size.text.split())
    if win_w <= 0 or win_h <= 0:
        raise CascadeFormatError("Invalid size node: width/height must be positive integer")
    stages = node.find("stages")
    if stages is None or len(stages) == 0:
        raise CascadeFormatError("Invalid stages node")

    st_ntrees, st_thr, st_parent, st_next = [], [], [], []
    tr_nnodes, nd_tilted, nd_rect, nd_weight, nd_thr, nd_left, nd_right, alpha = ([] for _ in range(8))
    n_stages = len(stages)
    for i, stage in enumerate(stages):
        trees = stage.find("trees")
        if trees is None or len(trees) == 0:
            raise CascadeFormatError(f"Trees node is not a valid sequence. (stage {i})")
        st_ntrees.append(len(trees))
        for j, tree in enumerate(trees):
            nodes = list(tree)
            if not nodes:
                raise CascadeFormatError(f"Tree node is not a valid sequence. (stage {i}, tree {j})")
            tr_nnodes.append(len(nodes))
            leaves = []
            for k, nd in enumerate(nodes):
                feature = nd.find("feature")
                rects = feature.find("rects") if feature is not None else None
                if rects is None or not (1 <= len(rects) <= 3):
                    raise CascadeFormatError(
                        f"Rects node is not a valid sequence. (stage {i}, tree {j}, node {k})")
                rr = np.zeros((3, 4), np.int32)
                ww = np.zeros(3, np.float32)
                for l, r in enumerate(rects):
                    tok = r.text.split()
                    if len(tok) != 5:
                        raise CascadeFormatError(
                            f"Rect {l} is not a valid sequence. (stage {i}, tree {j}, node {k})")
                    x, y, w, h = (int(t) for t in tok[:4])
                    if x < 0 or y < 0 or w <= 0 or h <= 0 or x + w > win_w or y + h > win_h:
                        raise CascadeFormatError(
                            f"rect out of window (stage {i}, tree {j}, node {k}, rect {l})")
                    rr[l] = (x, y, w, h)
                    ww[l] = _f32(tok[4])
                tilted = feature.find("tilted")
                if tilted is None:
                    raise CascadeFormatError(f"tilted must be 0 or 1. (stage {i}, tree {j}, node {k})")
                nd_tilted.append(int(int(tilted.text) != 0))
                nd_rect.append(rr)
                nd_weight.append(ww)
                thr = nd.find("threshold")
                if thr is None:
                    raise CascadeFormatError(
                        f"threshold must be real number. (stage {i}, tree {j}, node {k})")
                nd_thr.append(_f32(thr.text))
                for side, out in (("left", nd_left), ("right", nd_right)):
                    child = nd.find(side + "_node")
                    if child is not None:
                        idx = int(child.text)
                        if idx <= k or idx >= len(nodes):
                            raise CascadeFormatError(
                                f"{side} node must be valid node number. (stage {i}, tree {j}, node {k})")
                        out.append(idx)
                    else:
                        val = nd.find(side + "_val")
                        if val is None:
                            raise CascadeFormatError(
                                f"{side} node or {side} value must be specified. "
                                f"(stage {i}, tree {j}, node {k})")
                        if len(leaves) >= len(nodes) + 1:
                            raise CascadeFormatError(
                                f"Tree structure is broken: too many values. (stage {i}, tree {j}, node {k})")
                        out.append(-len(leaves))
                        leaves.append(_f32(val.text))
            if len(leaves) != len(nodes) + 1:
                raise CascadeFormatError(
                    f"Tree structure is broken: too few values. (stage {i}, tree {j})")
            alpha.extend(leaves)
        thr = stage.find("stage_threshold")
        if thr is None:
            raise CascadeFormatError(f"stage threshold must be real number. (stage {i})")
        st_thr.append(_f32(thr.text))
        for tag, out in (("parent", st_parent), ("next", st_next)):
            el = stage.find(tag)
            if el is None or not (-1 <= int(el.text) < n_stages):
                raise CascadeFormatError(f"{tag} must be integer number. (stage {i})")
            out.append(int(el.text))

    return FlatCascade(
        name=node.tag, win_w=win_w, win_h=win_h,
        st_ntrees=np.asarray(st_ntrees, np.int32), st_thr=np.asarray(st_thr, np.float32),
        st_parent=np.asarray(st_parent, np.int32), st_next=np.asarray(st_next, np.int32),
        tr_nnodes=np.asarray(tr_nnodes, np.int32), nd_tilted=np.asarray(nd_tilted, np.int32),
        nd_rect=np.ascontiguousarray(np.stack(nd_rect).astype(np.int32)),
        nd_weight=np.ascontiguousarray(np.stack(nd_weight).astype(np.float32)),
        nd_thr=np.asarray(nd_thr, np.float32),
        nd_left=np.asarray(nd_left, np.int32), nd_right=np.asarray(nd_right, np.int32),
        alpha=np.asarray(alpha, np.float32))
