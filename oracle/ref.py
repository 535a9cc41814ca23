"""ctypes front end of oracle/_ref/libtempcv_ref.so -- the REFERENCE'S OWN Haar code
(tempcv.cpp:40-1516, 1702-2089) compiled by oracle/build_ref.py against a test-only OpenCV
stand-in (oracle/ref_shim/).

TEST INFRASTRUCTURE ONLY: used to pin oracle/vj_oracle.c (the restatement every GPU test is
compared with) to the reference itself, and to produce the golden vectors under tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref
from .cascade_xml import FlatCascade

_lib = None

CV_HAAR_SCALE_IMAGE = 2


def available() -> bool:
    return build_ref.build() is not None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libtempcv_ref.so: no reference sources and no prebuilt library")
        L = C.CDLL(path)
        ip, fp, i32, i64, dp, u8 = (C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8))
        L.tcv_last_error.restype = C.c_char_p
        L.tcv_cascade_load_xml.restype = C.c_void_p
        L.tcv_cascade_load_xml.argtypes = [C.c_char_p]
        L.tcv_cascade_from_arrays.restype = C.c_void_p
        L.tcv_cascade_from_arrays.argtypes = [C.c_int, C.c_int, C.c_int, ip, fp, ip, ip, ip, ip, ip, fp, fp, ip, ip, fp]
        L.tcv_cascade_free.argtypes = [C.c_void_p]
        L.tcv_cascade_counts.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.tcv_cascade_dump.argtypes = [C.c_void_p, ip, fp, ip, ip, ip, ip, ip, ip, fp, fp, ip, ip, fp]
        L.tcv_cascade_hid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, fp, ip, i64, fp, ip, ip, dp, i64]
        L.tcv_eval_level.restype = C.c_int64
        L.tcv_eval_level.argtypes = [C.c_void_p, u8, C.c_int, C.c_int, C.c_int, C.c_int, i32, dp, C.c_int]
        L.tcv_eval_scaled.restype = C.c_int64
        L.tcv_eval_scaled.argtypes = [C.c_void_p, u8, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_int, C.c_int, i32, C.c_int]
        L.tcv_detect.restype = C.c_int64
        L.tcv_detect.argtypes = [C.c_void_p, u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32, i32, i32, dp, C.c_int64]
        L.tcv_group_rectangles.argtypes = [i32, C.c_int, C.c_int, C.c_double, i32]
        L.tcv_group_rectangles_roc.argtypes = [i32, C.c_int, C.c_int, C.c_double, i32, dp]
        L.tcv_clod_detect.restype = C.c_int64
        L.tcv_clod_detect.argtypes = [C.c_void_p, u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_uint, i32, C.c_int64]
        _lib = L
    return _lib


def _err() -> str:
    return lib().tcv_last_error().decode(errors="replace")


class RefCascade:
    """A CvHaarClassifierCascade owned by the reference's code: read by its icvReadHaarClassifier
    (from a path) or assembled from a FlatCascade (synthetic cascades)."""

    def __init__(self, src):
        L = lib()
        if isinstance(src, (str, os.PathLike)):
            self._h = L.tcv_cascade_load_xml(os.fspath(src).encode())
        else:
            f: FlatCascade = src
            self._h = L.tcv_cascade_from_arrays(
                f.win_w, f.win_h, f.n_stages, _p(f.st_ntrees, C.c_int), _p(f.st_thr, C.c_float),
                _p(f.st_parent, C.c_int), _p(f.st_next, C.c_int), _p(f.tr_nnodes, C.c_int),
                _p(f.nd_tilted, C.c_int), _p(f.nd_rect, C.c_int), _p(f.nd_weight, C.c_float),
                _p(f.nd_thr, C.c_float), _p(f.nd_left, C.c_int), _p(f.nd_right, C.c_int), _p(f.alpha, C.c_float))
        if not self._h:
            raise ValueError(_err())
        v = [C.c_int() for _ in range(5)]
        L.tcv_cascade_counts(self._h, *[C.byref(x) for x in v])
        self.win_w, self.win_h, self.n_stages, self.n_trees, self.n_nodes = [x.value for x in v]

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.tcv_cascade_free(self._h)
            self._h = None

    def flat(self, name: str = "ref") -> tuple[FlatCascade, np.ndarray]:
        """(FlatCascade as the reference's reader built it, st_child)"""
        S, T, N = self.n_stages, self.n_trees, self.n_nodes
        a = dict(st_ntrees=np.zeros(S, np.int32), st_thr=np.zeros(S, np.float32), st_parent=np.zeros(S, np.int32),
                 st_next=np.zeros(S, np.int32), st_child=np.zeros(S, np.int32), tr_nnodes=np.zeros(T, np.int32),
                 nd_tilted=np.zeros(N, np.int32), nd_rect=np.zeros((N, 3, 4), np.int32),
                 nd_weight=np.zeros((N, 3), np.float32), nd_thr=np.zeros(N, np.float32),
                 nd_left=np.zeros(N, np.int32), nd_right=np.zeros(N, np.int32), alpha=np.zeros(N + T, np.float32))
        lib().tcv_cascade_dump(self._h, _p(a["st_ntrees"], C.c_int), _p(a["st_thr"], C.c_float),
                               _p(a["st_parent"], C.c_int), _p(a["st_next"], C.c_int), _p(a["st_child"], C.c_int),
                               _p(a["tr_nnodes"], C.c_int), _p(a["nd_tilted"], C.c_int), _p(a["nd_rect"], C.c_int),
                               _p(a["nd_weight"], C.c_float), _p(a["nd_thr"], C.c_float), _p(a["nd_left"], C.c_int),
                               _p(a["nd_right"], C.c_int), _p(a["alpha"], C.c_float))
        child = a.pop("st_child")
        return FlatCascade(name=name, win_w=self.win_w, win_h=self.win_h, **a), child

    def hidden(self, W: int, H: int, scale: float = 1.0):
        """Hidden cascade after cvSetImagesForHaarClassifierCascade(scale) on (W+1)x(H+1) integrals ->
        dict(weights[N,3], nrects[N], corners[N,3,4] (element offsets p0..p3), stage_thr[S],
        two_rects[S], flags, inv_window_area, eq_corners[4])"""
        N, S = self.n_nodes, self.n_stages
        w = np.zeros((N, 3), np.float32)
        nr = np.zeros(N, np.int32)
        co = np.zeros((N, 3, 4), np.int64)
        st = np.zeros(S, np.float32)
        two = np.zeros(S, np.int32)
        fl = C.c_int()
        inv = C.c_double()
        eq = np.zeros(4, np.int64)
        if lib().tcv_cascade_hid(self._h, W, H, scale, _p(w, C.c_float), _p(nr, C.c_int), _p(co, C.c_int64),
                                 _p(st, C.c_float), _p(two, C.c_int), C.byref(fl), C.byref(inv), _p(eq, C.c_int64)):
            raise ValueError(_err())
        return dict(weights=w, nrects=nr, corners=co, stage_thr=st, two_rects=two, flags=fl.value,
                    inv_window_area=inv.value, eq_corners=eq)

    def eval_level(self, img: np.ndarray, ystep: int, n_threads: int = 0):
        """cvRunHaarClassifierCascadeSum on every grid window of one image taken as a level
        -> (results int32 [ny,nx] raw return values, stage_sums float64 [ny,nx])"""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        nx = -(-(W - self.win_w) // ystep) if W > self.win_w else 0
        ny = -(-(H - self.win_h) // ystep) if H > self.win_h else 0
        res = np.zeros(max(nx * ny, 1), np.int32)
        ss = np.zeros(max(nx * ny, 1), np.float64)
        n = lib().tcv_eval_level(self._h, _p(img, C.c_uint8), W, H, img.strides[0], ystep, _p(res, C.c_int32),
                                 _p(ss, C.c_double), n_threads)
        if n < 0:
            raise ValueError(_err())
        assert n == nx * ny
        return res[:n].reshape(ny, nx), ss[:n].reshape(ny, nx)

    def eval_scaled(self, img: np.ndarray, factor: float, step: float, nx: int, ny: int, n_threads: int = 0):
        """cvSetImages(scale=factor) + cvRunHaarClassifierCascade at (cvRound(ix*step), cvRound(iy*step))
        for every ix < nx, iy < ny -> results int32 [ny,nx]"""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        res = np.zeros(max(nx * ny, 1), np.int32)
        n = lib().tcv_eval_scaled(self._h, _p(img, C.c_uint8), W, H, img.strides[0], factor, step, nx, ny,
                                  _p(res, C.c_int32), n_threads)
        if n < 0:
            raise ValueError(_err())
        return res[:nx * ny].reshape(ny, nx)

    def detect(self, img: np.ndarray, scale_factor: float, min_neighbors: int = 0, flags: int = CV_HAAR_SCALE_IMAGE,
               min_size=(0, 0), max_size=(0, 0), reject_levels: bool = False):
        """The reference's cvHaarDetectObjectsForROC, whole -> (rects[n,4], neighbors[n], levels, weights).
        img: [H,W] gray or [H,W,3] BGR."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        cap = 1 << 16
        while True:
            rects = np.zeros((cap, 4), np.int32)
            nb = np.zeros(cap, np.int32)
            lv = np.zeros(cap, np.int32)
            wt = np.zeros(cap, np.float64)
            n = lib().tcv_detect(self._h, _p(img, C.c_uint8), W, H, img.strides[0], ch, scale_factor, min_neighbors,
                                 flags, min_size[0], min_size[1], max_size[0], max_size[1], int(reject_levels),
                                 _p(rects, C.c_int32), _p(nb, C.c_int32), _p(lv, C.c_int32), _p(wt, C.c_double), cap)
            if n < 0:
                raise ValueError(_err())
            if n <= cap:
                break
            cap = int(n)
        return rects[:n].copy(), nb[:n].copy(), lv[:n].copy(), wt[:n].copy()


def _clod_detect(self, img: np.ndarray, flags: int = (2 << 2) | (2 << 0), min_size=(0, 0), max_size=(0, 0)):
    """The reference's clodDetectObjects(use_cl=FALSE, min_neighbors=0), clod.cpp:1339-1500 (scale factor 1.1, hard-coded
    at clod.cpp:1349) -> raw matches [n,4].  flags: CLOD_PER_STAGE_ITERATIONS | CLOD_PRECOMPUTE_FEATURES as main.cpp:79."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    cap = 1 << 16
    while True:
        rects = np.zeros((cap, 4), np.int32)
        n = lib().tcv_clod_detect(self._h, _p(img, C.c_uint8), W, H, img.strides[0], min_size[0], min_size[1],
                                  max_size[0], max_size[1], flags, _p(rects, C.c_int32), cap)
        if n <= cap:
            break
        cap = int(n)
    return rects[:n].copy()


RefCascade.clod_detect = _clod_detect


def group_rectangles(rects: np.ndarray, group_threshold: int, eps: float = 0.2):
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4).copy()
    w = np.zeros(max(len(r), 1), np.int32)
    if len(r) == 0:
        return r, w[:0]
    m = lib().tcv_group_rectangles(_p(r, C.c_int32), len(r), group_threshold, eps, _p(w, C.c_int32))
    if m < 0:
        raise ValueError(_err())
    return r[:m].copy(), w[:m].copy()


def group_rectangles_roc(rects, reject_levels, level_weights, group_threshold: int, eps: float = 0.2):
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4).copy()
    lv = np.ascontiguousarray(reject_levels, np.int32).copy()
    wt = np.ascontiguousarray(level_weights, np.float64).copy()
    if len(r) == 0:
        return r, lv, wt
    m = lib().tcv_group_rectangles_roc(_p(r, C.c_int32), len(r), group_threshold, eps, _p(lv, C.c_int32),
                                       _p(wt, C.c_double))
    if m < 0:
        raise ValueError(_err())
    return r[:m].copy(), lv[:m].copy(), wt[:m].copy()
