#!/usr/bin/env python
"""CPU model of k_cascade_tiles' shared-memory traffic (tools/, not product code).

Replays the tile kernel's schedule -- fixed-geometry phase, hand-over by bank class, warp-autonomous
compacted phase, window x group mode -- on the exit codes the oracle computes for one frame, and
counts warp-level LDS instructions and 128-byte wavefronts (bank conflicts from the real tile
addresses).  Used to rank scheduling variants before spending GPU time; the `current` variant is
checked against ncu's l1tex__data_pipe_lsu_wavefronts_mem_shared (profiles/).

  python tools/tile_model.py [--cascade frontalface_alt] [--w 1920 --h 1080] [--variant current,...]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from clfacedetection_b200.frames import octave_frame  # noqa: E402

TW = 64
NWARP = 8


def dense_half(win_w, ystep):
    cols = ((TW - 1) * ystep + win_w + 1 + 3) & ~3
    return 0 if ystep == 1 else (cols // 2 + 3) & ~3


def dense_stride(win_w, ystep):
    cols = ((TW - 1) * ystep + win_w + 1 + 3) & ~3
    s = cols if ystep == 1 else 2 * dense_half(win_w, ystep)
    s = (s + 3) & ~3
    while (ystep * s) % 32 != 8:
        s += 4
    return s


def stump_tables(cas, ystep):
    """per stump: word offsets of its corners in the tile (list), number of loads in the resident
    (six-load) form, stage index"""
    f = cas.flat
    _, nr, _, _, _ = cas.hidden()
    S = dense_stride(f.win_w, ystep)

    def off(y, x):
        return y * S + x if ystep == 1 else y * S + (x & 1) * dense_half(f.win_w, ystep) + (x >> 1)
    offs, loads6 = [], []
    for n in range(f.n_nodes):
        pts = []
        rects = []
        for k in range(int(nr[n])):
            x, y, w, h = (int(v) for v in f.nd_rect[n, k])
            c = [(y, x), (y, x + w), (y + h, x), (y + h, x + w)]
            rects.append(c)
            pts += c
        offs.append(np.array([off(y, x) for (y, x) in pts], np.int64))
        shared = len(rects) == 2 and len(set(rects[0]) & set(rects[1])) == 2
        loads6.append(6 if shared else len(pts))
    first = np.concatenate([[0], np.cumsum(f.st_ntrees)])
    return offs, np.array(loads6), first, S


def wavefronts_rows(classes, valid):
    """classes [R,32] bank class per lane, valid [R,32] -> wavefronts per load for each row
    (all lanes read the same corner of the same stump: bank = class + const, distinct windows =
    distinct addresses)"""
    out = np.zeros(classes.shape[0], np.int64)
    for r in range(classes.shape[0]):
        v = classes[r][valid[r]]
        out[r] = np.bincount(v, minlength=32).max() if v.size else 0
    return out


class Counter:
    def __init__(self, n_stages):
        self.instr = np.zeros(n_stages, np.float64)   # warp-level LDS instructions
        self.wave = np.zeros(n_stages, np.float64)    # wavefronts
        self.useful = np.zeros(n_stages, np.float64)  # corner loads of live windows (thread level)


def simulate_tile(depth, wxs, wys, ystep, S, offs, loads6, first, n_stages, variant, cnt, rng):
    """depth: stages passed per window of the tile (== n_stages: accepted)"""
    nwin = depth.size
    cls = (wxs + 8 * wys) & 31
    base = ystep * wys * S + wxs
    TH = wys.max() + 1
    nfix = variant.get("n_fixed", 3)
    alive = np.ones(nwin, bool)
    # ---- phase 1: thread t owns windows (wx = t & 63, wy = t//64 + 4k); chunks of 4 (3 for 24 rows) slots
    slots = (TH + 3) // 4
    chunk = 4 if slots % 4 == 0 else 3
    s = 0
    grid = np.zeros((TH_pad(TH), TW), bool)
    while s < nfix:
        if variant.get("adaptive") and s >= 2 and alive.mean() < variant["adaptive"]:
            break
        grid[:] = False
        grid[wys, wxs] = alive
        nloads = loads6[first[s]:first[s + 1]].sum()
        # a chunk = slots k0..k0+chunk-1 of the 8 warps: warp w covers wx half (w & 1), wy0 = w >> 1
        for k0 in range(0, slots, chunk):
            for w in range(NWARP):
                x0, wy0 = 32 * (w & 1), w >> 1
                rows = [wy0 + 4 * k for k in range(k0, k0 + chunk) if wy0 + 4 * k < grid.shape[0]]
                blk = grid[rows, x0:x0 + 32]
                if not blk.any():
                    continue
                cnt.instr[s] += nloads * chunk
                cnt.wave[s] += nloads * chunk
        cnt.useful[s] += alive.sum() * nloads
        alive &= depth > s
        s += 1
    idx = np.flatnonzero(alive)
    if idx.size == 0:
        return
    # ---- hand-over
    n_alive = idx.size
    order = idx[np.argsort(cls[idx], kind="stable")]   # counting sort by class (rank i)
    lists = [[] for _ in range(NWARP)]
    if variant.get("full_rows"):
        # rows of one window per class while every class still has one: conflict free
        counts = np.bincount(cls[idx], minlength=32)
        m = counts.min()
        rank_in_class = np.zeros(n_alive, np.int64)
        pos = {}
        for j, wdw in enumerate(order):
            c = cls[wdw]
            rank_in_class[j] = pos.get(c, 0)
            pos[c] = rank_in_class[j] + 1
        full = order[rank_in_class < m]
        rest = order[rank_in_class >= m]
        # full rows: row r = windows with rank r, lane = class
        fr = [[] for _ in range(m)]
        for j, wdw in enumerate(order):
            if rank_in_class[j] < m:
                fr[rank_in_class[j]].append(wdw)
        warp_rows = [[] for _ in range(NWARP)]
        for r in range(m):
            warp_rows[r % NWARP].append(np.array(fr[r]))
        rest_lists = [rest[w::NWARP] for w in range(NWARP)]
    else:
        warp_rows = [[] for _ in range(NWARP)]
        rest_lists = [order[w::NWARP] for w in range(NWARP)]
    g1 = variant.get("g1_min", 12)
    if variant.get("pool"):
        # tile-wide pool: at every stage the tile's survivors are class-sorted and dealt into
        # ceil(n/32) balanced rows (warps pull rows); window x group mode once n <= g1
        surv = idx
        st = s
        T = variant.get("T", 0)
        while st < n_stages and surv.size and (surv.size > T or not T):
            nloads = loads6[first[st]:first[st + 1]].sum()
            cnt.useful[st] += surv.size * nloads
            if surv.size > g1:
                o = surv[np.argsort(cls[surv], kind="stable")]
                R0 = (o.size + 31) >> 5
                hmax = np.bincount(cls[surv], minlength=32).max()
                best = None
                if variant.get("heur"):
                    h = np.bincount(cls[surv], minlength=32)
                    bb = variant.get("beta", 0.8)
                    ests = [(R * (1 + bb) + min(R, np.maximum(h - R, 0).sum()), R) for R in range(R0, hmax + 1)]
                    Rh = min(ests)[1]
                    cand = [Rh]
                else:
                    cand = (range(R0, hmax + 1) if variant.get("optR") else [R0])
                for R in cand:
                    wv = sum(np.bincount(cls[o[r::R]], minlength=32).max() for r in range(R))
                    cost = wv + variant.get("beta", 0.8) * R
                    if best is None or cost < best[0]:
                        best = (cost, R, wv)
                cnt.instr[st] += nloads * best[1]
                cnt.wave[st] += nloads * best[2]
            else:
                n1 = surv.size
                lw = 4
                while lw > 0 and (1 << (lw - 1)) >= n1:
                    lw -= 1
                G = 32 >> lw
                ids = list(range(first[st], first[st + 1]))
                for it in range((len(ids) + G - 1) // G):
                    js = ids[it * G:(it + 1) * G]
                    ncorner = max(offs[j].size for j in js)
                    for cidx in range(ncorner):
                        banks = {}
                        for j in js:
                            if cidx >= offs[j].size:
                                continue
                            for wdw in surv:
                                a = int(base[wdw] + offs[j][cidx])
                                banks.setdefault(a & 31, set()).add(a)
                        cnt.instr[st] += 1
                        cnt.wave[st] += max(len(v) for v in banks.values()) if banks else 0
            surv = surv[depth[surv] > st]
            st += 1
        if not T or not surv.size:
            return
        # the rest: dealt to the warps (sorted by class, round robin), warp-autonomous as before
        order = surv[np.argsort(cls[surv], kind="stable")]
        warp_rows = [[] for _ in range(NWARP)]
        rest_lists = [order[w::NWARP] for w in range(NWARP)]
        s = st
    for w in range(NWARP):
        # first compacted stage rows: the full rows, then the dealt share column-major over R rows
        rest = rest_lists[w]
        n = rest.size
        rows = list(warp_rows[w])
        if n:
            R = (n + 31) >> 5
            for r in range(R):
                rows.append(rest[r::R])   # column-major dealing: row r holds entries r, r+R, ...
        cur_rows = rows
        st = s
        while st < n_stages and cur_rows:
            nloads = loads6[first[st]:first[st + 1]].sum() if st < 99 else 0
            stump_ids = range(first[st], first[st + 1])
            nwin_w = sum(r.size for r in cur_rows)
            cnt.useful[st] += nwin_w * nloads
            if nwin_w > g1:
                # thread per window, two rows per pass (instructions per row all the same)
                for r in cur_rows:
                    mult = np.bincount(cls[r], minlength=32).max()
                    cnt.instr[st] += nloads
                    cnt.wave[st] += nloads * mult
            else:
                allw = np.concatenate(cur_rows)
                n1 = allw.size
                lw = 4
                while lw > 0 and (1 << (lw - 1)) >= n1:
                    lw -= 1
                wslots, G = 1 << lw, 32 >> lw
                ids = list(stump_ids)
                # lane (slot, grp) evaluates stumps grp, grp+G, ...
                for it in range((len(ids) + G - 1) // G):
                    js = ids[it * G:(it + 1) * G]
                    ncorner = max(offs[j].size for j in js)
                    for cidx in range(ncorner):
                        banks = {}
                        for g, j in enumerate(js):
                            if cidx >= offs[j].size:
                                continue
                            for wdw in allw:
                                a = int(base[wdw] + offs[j][cidx])
                                banks.setdefault(a & 31, set()).add(a)
                        cnt.instr[st] += 1
                        cnt.wave[st] += max(len(v) for v in banks.values()) if banks else 0
                cur_rows = [allw]
            # survivors, re-compacted contiguously: rows of 32
            surv = np.concatenate(cur_rows)
            surv = surv[depth[surv] > st]
            if variant.get("resort") and surv.size:
                # per-warp counting sort by class, dealt column-major over R balanced rows
                o = surv[np.argsort(cls[surv], kind="stable")]
                R = (o.size + 31) >> 5
                cur_rows = [o[r::R] for r in range(R)]
            else:
                cur_rows = [surv[i:i + 32] for i in range(0, surv.size, 32)]
            st += 1


def TH_pad(th):
    return ((th + 3) // 4) * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cascade", default="frontalface_alt")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--sf", type=float, default=1.2)
    ap.add_argument("--variants", default="current,full_rows,adaptive,both")
    args = ap.parse_args()
    cas = oracle.Cascade(os.path.join(ROOT, "data", "haarcascades", f"haarcascade_{args.cascade}.xml"))
    frame = octave_frame(args.w, args.h, 0)
    _, codes, _, st, levels = cas.detect(frame, args.sf)
    n_stages = cas.flat.n_stages
    variants = {"current": {}, "full_rows": {"full_rows": True}, "adaptive": {"adaptive": 0.40},
                "both": {"full_rows": True, "adaptive": 0.40}, "nfix2": {"n_fixed": 2},
                "nfix2_full": {"n_fixed": 2, "full_rows": True}, "resort": {"resort": True},
                "resort_nfix2": {"resort": True, "n_fixed": 2}, "resort_g8": {"resort": True, "g1_min": 8},
                "resort_g6": {"resort": True, "g1_min": 6}, "g8": {"g1_min": 8}, "g16": {"g1_min": 16},
                "pool": {"pool": True}, "pool_nfix2": {"pool": True, "n_fixed": 2}, "pool_g16": {"pool": True, "g1_min": 16},
                "pool_nfix1": {"pool": True, "n_fixed": 1},
                "opt_nfix2": {"pool": True, "n_fixed": 2, "optR": True}, "opt_nfix1": {"pool": True, "n_fixed": 1, "optR": True},
                "opt_nfix3": {"pool": True, "n_fixed": 3, "optR": True},
                "hyb2_64": {"pool": True, "n_fixed": 2, "optR": True, "T": 64}, "hyb2_128": {"pool": True, "n_fixed": 2, "optR": True, "T": 128},
                "hyb2_256": {"pool": True, "n_fixed": 2, "optR": True, "T": 256}, "hyb1_128": {"pool": True, "n_fixed": 1, "optR": True, "T": 128},
                "hyb2_128_g16": {"pool": True, "n_fixed": 2, "optR": True, "T": 128, "g1_min": 16},
                "cl2_128": {"pool": True, "n_fixed": 2, "optR": True, "beta": 0.0, "T": 128},
                "heur2_128": {"pool": True, "n_fixed": 2, "heur": True, "T": 128}, "heur2_96": {"pool": True, "n_fixed": 2, "heur": True, "T": 96}}
    tabs = {ys: stump_tables(cas, ys) for ys in (1, 2)}
    rng = np.random.default_rng(0)
    for vn in args.variants.split(","):
        var = variants[vn]
        cnt = Counter(n_stages)
        off = 0
        for lv in levels:
            n = lv.nx * lv.ny
            depth = codes[off:off + n].astype(np.int64).reshape(lv.ny, lv.nx)
            off += n
            offs, loads6, first, S = tabs[lv.ystep]
            TH = 24 if lv.ystep == 2 else 32
            for ty in range(0, lv.ny, TH):
                for tx in range(0, lv.nx, TW):
                    d = depth[ty:ty + TH, tx:tx + TW]
                    wy, wx = np.mgrid[0:d.shape[0], 0:d.shape[1]]
                    simulate_tile(d.ravel(), wx.ravel(), wy.ravel(), lv.ystep, S, offs, loads6, first, n_stages, var, cnt, rng)
        print(f"== {vn}: LDS warp-instr {cnt.instr.sum() / 1e6:.2f} M, wavefronts {cnt.wave.sum() / 1e6:.2f} M, "
              f"useful thread loads/32 {cnt.useful.sum() / 32e6:.2f} M  (per frame, stump corner loads only)")
        print("   stage  instr(M)  wave(M)  useful/32(M)")
        for s_ in range(min(n_stages, 12)):
            print(f"   {s_:5d} {cnt.instr[s_] / 1e6:9.2f} {cnt.wave[s_] / 1e6:8.2f} {cnt.useful[s_] / 32e6:10.2f}")
        print(f"   rest  {cnt.instr[12:].sum() / 1e6:9.2f} {cnt.wave[12:].sum() / 1e6:8.2f} {cnt.useful[12:].sum() / 32e6:10.2f}")


if __name__ == "__main__":
    main()
