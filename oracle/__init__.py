"""ctypes front end of the CPU oracle (oracle/vj_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never by clfacedetection_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from .cascade_xml import FlatCascade, load_cascade_xml  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvj_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile."""
    deps = [os.path.join(_HERE, f) for f in ("vj_oracle.c", "vj_oracle.h", "clod_cpu.c", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(d) for d in deps)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


CLOD_PRECOMPUTE_FEATURES = 2 << 0     # clod.h:17
CLOD_PER_STAGE_ITERATIONS = 2 << 2    # clod.h:19


class _Level(C.Structure):
    _fields_ = [("factor", C.c_double), ("img_w", C.c_int), ("img_h", C.c_int),
                ("win_w", C.c_int), ("win_h", C.c_int), ("ystep", C.c_int),
                ("nx", C.c_int), ("ny", C.c_int)]


class _Stats(C.Structure):
    _fields_ = [("windows", C.c_int64), ("weak_evals", C.c_int64), ("node_evals", C.c_int64),
                ("accepted", C.c_int64), ("near_stage_thr", C.c_int64),
                ("stage_reach", C.c_int64 * 64), ("near_stage_events", C.c_int64)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.vjo_cascade_create.restype = C.c_void_p
        L.vjo_cascade_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, fp, ip, ip, ip, ip, ip, fp, fp, ip, ip, fp]
        L.vjo_cascade_free.argtypes = [C.c_void_p]
        L.vjo_last_error.restype = C.c_char_p
        L.vjo_cascade_flags.argtypes = [C.c_void_p]
        L.vjo_cascade_hid.argtypes = [C.c_void_p, fp, ip, fp, ip, ip]
        L.vjo_resize_linear.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int]
        L.vjo_integral.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.vjo_plan_levels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Level), C.c_int]
        L.vjo_detect.restype = C.c_int64
        L.vjo_detect.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int64,
                                 C.POINTER(C.c_int16), C.POINTER(C.c_uint8), C.POINTER(_Stats), C.c_int]
        L.vjo_detect_roc.restype = C.c_int64
        L.vjo_detect_roc.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int16), C.POINTER(C.c_uint8),
                                     C.POINTER(_Stats), C.c_int]
        L.vjo_group_rectangles_roc.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_double,
                                               C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.vjo_plan_sc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                  C.POINTER(_Level), C.c_int]
        L.vjo_detect_sc.restype = C.c_int64
        L.vjo_detect_sc.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int16),
                                    C.POINTER(_Stats), C.c_int]
        L.vjo_eval_level.restype = C.c_int64
        L.vjo_eval_level.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int16), C.POINTER(C.c_uint8), C.POINTER(_Stats), C.c_int]
        L.vjo_group_rectangles.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_double,
                                           C.POINTER(C.c_int32)]
        i64p, u8p = C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
        cas = [C.c_int, C.c_int, C.c_int, ip, fp, ip, ip, fp, fp, fp]
        L.clodcpu_detect.restype = C.c_int64
        L.clodcpu_detect.argtypes = cas + [u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_uint, C.POINTER(C.c_int32), C.c_int64, i64p, i64p]
        L.clodcpu_detect_batch.restype = C.c_int64
        L.clodcpu_detect_batch.argtypes = cas + [u8p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, i64p, i64p, i64p, C.c_int]
        _lib = L
    return _lib


@dataclass
class Level:
    factor: float
    img_w: int
    img_h: int
    win_w: int
    win_h: int
    ystep: int
    nx: int
    ny: int


@dataclass
class Stats:
    windows: int
    weak_evals: int
    node_evals: int
    accepted: int
    near_stage_thr: int
    stage_reach: list
    near_stage_events: int = 0


def _stats(s: _Stats) -> Stats:
    return Stats(s.windows, s.weak_evals, s.node_evals, s.accepted, s.near_stage_thr,
                 list(s.stage_reach), s.near_stage_events)


CODE_SKIPPED, CODE_OUTSIDE = -32768, -32767


class Cascade:
    """Oracle-side cascade: FlatCascade -> hidden cascade (tempcv.cpp:308-467,549-768)."""

    def __init__(self, flat: FlatCascade | str):
        if isinstance(flat, str):
            flat = load_cascade_xml(flat)
        self.flat = flat
        L = lib()
        self._h = L.vjo_cascade_create(
            flat.win_w, flat.win_h, flat.n_stages,
            _p(flat.st_ntrees, C.c_int), _p(flat.st_thr, C.c_float), _p(flat.st_parent, C.c_int),
            _p(flat.st_next, C.c_int), _p(flat.tr_nnodes, C.c_int), _p(flat.nd_tilted, C.c_int),
            _p(flat.nd_rect, C.c_int), _p(flat.nd_weight, C.c_float), _p(flat.nd_thr, C.c_float),
            _p(flat.nd_left, C.c_int), _p(flat.nd_right, C.c_int), _p(flat.alpha, C.c_float))
        if not self._h:
            raise ValueError(L.vjo_last_error().decode())
        fl = L.vjo_cascade_flags(self._h)
        self.is_tree, self.is_stump_based, self.has_tilted = bool(fl & 1), bool(fl & 2), bool(fl & 4)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vjo_cascade_free(self._h)
            self._h = None

    @property
    def win(self):
        return self.flat.win_w, self.flat.win_h

    def hidden(self):
        """(node_weights[N,3], node_nrects[N], stage_thr[S], stage_two_rects[S], stage_child[S])"""
        f = self.flat
        w = np.zeros((f.n_nodes, 3), np.float32)
        nr = np.zeros(f.n_nodes, np.int32)
        st = np.zeros(f.n_stages, np.float32)
        two = np.zeros(f.n_stages, np.int32)
        ch = np.zeros(f.n_stages, np.int32)
        lib().vjo_cascade_hid(self._h, _p(w, C.c_float), _p(nr, C.c_int), _p(st, C.c_float),
                              _p(two, C.c_int), _p(ch, C.c_int))
        return w, nr, st, two, ch

    def plan_levels(self, W, H, scale_factor, min_size=(0, 0), max_size=(0, 0)):
        return plan_levels(W, H, self.flat.win_w, self.flat.win_h, scale_factor, min_size, max_size)

    def detect(self, img: np.ndarray, scale_factor: float, min_size=(0, 0), max_size=(0, 0),
               want_codes: bool = True, n_threads: int = 0):
        """REF-SI detection of one gray frame -> (rects[n,4], codes, near, Stats, levels)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        levels = self.plan_levels(W, H, scale_factor, min_size, max_size)
        nwin = sum(l.nx * l.ny for l in levels)
        codes = np.zeros(nwin, np.int16) if want_codes else None
        near = np.zeros(nwin, np.uint8) if want_codes else None
        cap = 1 << 16
        st = _Stats()
        while True:
            rects = np.zeros((cap, 4), np.int32)
            n = lib().vjo_detect(self._h, _p(img, C.c_uint8), W, H, img.strides[0], scale_factor,
                                 min_size[0], min_size[1], max_size[0], max_size[1],
                                 _p(rects, C.c_int32), cap, _p(codes, C.c_int16), _p(near, C.c_uint8),
                                 C.byref(st), n_threads)
            if n < 0:
                raise ValueError(lib().vjo_last_error().decode())
            if n <= cap:
                break
            cap = int(n)
        return rects[:n].copy(), codes, near, _stats(st), levels

    def _clod_args(self):
        f = self.flat
        return [f.win_w, f.win_h, f.n_stages, _p(f.st_ntrees, C.c_int), _p(f.st_thr, C.c_float), _p(f.tr_nnodes, C.c_int),
                _p(f.nd_rect, C.c_int), _p(f.nd_weight, C.c_float), _p(f.nd_thr, C.c_float), _p(f.alpha, C.c_float)]

    def clod_cpu_detect(self, img: np.ndarray, scale_factor: float = 1.1, min_size=(0, 0), max_size=(0, 0),
                        flags: int = CLOD_PER_STAGE_ITERATIONS | CLOD_PRECOMPUTE_FEATURES):
        """CLOD-CPU (oracle/clod_cpu.c: clodDetectObjects(use_cl=FALSE), clod.cpp:1339-1500) on one gray frame
        -> (raw matches [n,4], windows evaluated, classifier evaluations).  Stump cascades only."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = 1 << 16
        nw, ne = C.c_int64(), C.c_int64()
        while True:
            rects = np.zeros((cap, 4), np.int32)
            n = lib().clodcpu_detect(*self._clod_args(), _p(img, C.c_uint8), W, H, img.strides[0], scale_factor,
                                     min_size[0], min_size[1], max_size[0], max_size[1], flags,
                                     _p(rects, C.c_int32), cap, C.byref(nw), C.byref(ne))
            if n < 0:
                raise ValueError("CLOD-CPU runs stump cascades with at most 220 classifiers per stage (clod.cpp:13,458)")
            if n <= cap:
                break
            cap = int(n)
        return rects[:n].copy(), int(nw.value), int(ne.value)

    def clod_cpu_detect_batch(self, frames: np.ndarray, scale_factor: float, n_threads: int,
                              flags: int = CLOD_PER_STAGE_ITERATIONS | CLOD_PRECOMPUTE_FEATURES):
        """frames [n, H, W] uint8, one frame per thread at a time -> (match counts [n], windows, classifier evaluations)"""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, H, W = frames.shape
        counts = np.zeros(n, np.int64)
        nw, ne = C.c_int64(), C.c_int64()
        r = lib().clodcpu_detect_batch(*self._clod_args(), _p(frames, C.c_uint8), n, frames.strides[0], W, H, frames.strides[1],
                                       scale_factor, 0, 0, 0, 0, flags, _p(counts, C.c_int64), C.byref(nw), C.byref(ne), n_threads)
        if r < 0:
            raise ValueError("CLOD-CPU runs stump cascades with at most 220 classifiers per stage")
        return counts, int(nw.value), int(ne.value)

    def detect_roc(self, img: np.ndarray, scale_factor: float, min_size=(0, 0), max_size=(0, 0), n_threads: int = 0):
        """REF-SI with outputRejectLevels (tempcv.cpp:1084-1094) -> (rects[n,4], reject_levels[n],
        level_weights[n]) in the reference's scan order."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = 1 << 16
        st = _Stats()
        while True:
            rects = np.zeros((cap, 4), np.int32)
            lv = np.zeros(cap, np.int32)
            wt = np.zeros(cap, np.float64)
            n = lib().vjo_detect_roc(self._h, _p(img, C.c_uint8), W, H, img.strides[0], scale_factor,
                                     min_size[0], min_size[1], max_size[0], max_size[1],
                                     _p(rects, C.c_int32), _p(lv, C.c_int32), _p(wt, C.c_double), cap,
                                     None, None, C.byref(st), n_threads)
            if n < 0:
                raise ValueError(lib().vjo_last_error().decode())
            if n <= cap:
                break
            cap = int(n)
        return rects[:n].copy(), lv[:n].copy(), wt[:n].copy()

    def detect_sc(self, img: np.ndarray, scale_factor: float, min_size=(0, 0), want_codes: bool = True,
                  n_threads: int = 0):
        """REF-SC (scale-cascade, tempcv.cpp:1330-1456) detection of one gray frame
        -> (rects[n,4], codes, Stats, levels); codes use CODE_SKIPPED / CODE_OUTSIDE."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        buf = (_Level * 256)()
        nl = lib().vjo_plan_sc(W, H, self.flat.win_w, self.flat.win_h, scale_factor, min_size[0], min_size[1], buf, 256)
        if nl < 0:
            raise ValueError(lib().vjo_last_error().decode())
        levels = [Level(b.factor, b.img_w, b.img_h, b.win_w, b.win_h, b.ystep, b.nx, b.ny) for b in buf[:nl]]
        nwin = sum(l.nx * l.ny for l in levels)
        codes = np.zeros(nwin, np.int16) if want_codes else None
        cap = 1 << 16
        st = _Stats()
        while True:
            rects = np.zeros((cap, 4), np.int32)
            n = lib().vjo_detect_sc(self._h, _p(img, C.c_uint8), W, H, img.strides[0], scale_factor,
                                    min_size[0], min_size[1], _p(rects, C.c_int32), cap, _p(codes, C.c_int16),
                                    C.byref(st), n_threads)
            if n < 0:
                raise ValueError(lib().vjo_last_error().decode())
            if n <= cap:
                break
            cap = int(n)
        return rects[:n].copy(), codes, _stats(st), levels

    def eval_level(self, img: np.ndarray, ystep: int, n_threads: int = 0):
        """All grid windows of one image evaluated as a single level (no resize)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        w0, h0 = self.win
        nx = max(0, -(-(W - w0) // ystep)) if W > w0 else 0
        ny = max(0, -(-(H - h0) // ystep)) if H > h0 else 0
        codes = np.zeros(nx * ny, np.int16)
        near = np.zeros(nx * ny, np.uint8)
        st = _Stats()
        lib().vjo_eval_level(self._h, _p(img, C.c_uint8), W, H, img.strides[0], ystep,
                             _p(codes, C.c_int16), _p(near, C.c_uint8), C.byref(st), n_threads)
        return codes.reshape(ny, nx), near.reshape(ny, nx), _stats(st)


def plan_levels(W, H, w0, h0, scale_factor, min_size=(0, 0), max_size=(0, 0)):
    buf = (_Level * 256)()
    n = lib().vjo_plan_levels(W, H, w0, h0, scale_factor, min_size[0], min_size[1],
                              max_size[0], max_size[1], buf, 256)
    if n < 0:
        raise ValueError(lib().vjo_last_error().decode())
    return [Level(b.factor, b.img_w, b.img_h, b.win_w, b.win_h, b.ystep, b.nx, b.ny) for b in buf[:n]]


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw = src.shape
    dst = np.zeros((dh, dw), np.uint8)
    rc = lib().vjo_resize_linear(_p(src, C.c_uint8), sw, sh, src.strides[0], _p(dst, C.c_uint8), dw, dh, dw)
    if rc:
        raise ValueError(lib().vjo_last_error().decode())
    return dst


def integral(img: np.ndarray, tilted: bool = False):
    """-> (sum int32 [(h+1),(w+1)], sqsum float64, tilted int32 | None)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    s = np.zeros((h + 1, w + 1), np.int32)
    q = np.zeros((h + 1, w + 1), np.float64)
    t = np.zeros((h + 1, w + 1), np.int32) if tilted else None
    lib().vjo_integral(_p(img, C.c_uint8), w, h, img.strides[0], _p(s, C.c_int32), _p(q, C.c_double),
                       _p(t, C.c_int32))
    return s, q, t


def group_rectangles(rects: np.ndarray, group_threshold: int, eps: float = 0.2):
    """AgroupRectangles -> (rects[m,4], weights[m])"""
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4).copy()
    w = np.zeros(max(len(r), 1), np.int32)
    m = lib().vjo_group_rectangles(_p(r, C.c_int32), len(r), group_threshold, eps, _p(w, C.c_int32))
    return r[:m].copy(), w[:m].copy()


def group_rectangles_roc(rects: np.ndarray, reject_levels: np.ndarray, level_weights: np.ndarray,
                         group_threshold: int, eps: float = 0.2):
    """AgroupRectangles, ROC variant (tempcv.cpp:255-258) -> (rects[m,4], levels[m], weights[m])"""
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4).copy()
    lv = np.ascontiguousarray(reject_levels, np.int32).copy()
    wt = np.ascontiguousarray(level_weights, np.float64).copy()
    if len(r) == 0:
        return r, lv, wt
    m = lib().vjo_group_rectangles_roc(_p(r, C.c_int32), len(r), group_threshold, eps, _p(lv, C.c_int32), _p(wt, C.c_double))
    return r[:m].copy(), lv[:m].copy(), wt[:m].copy()
