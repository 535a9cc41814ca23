#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ (run in the build container:
needs cv2 4.x for the OpenCV pins and the oracle for the detection fixtures).

  * opencv_pins.npz : cv2.resize(INTER_LINEAR), cv2.integral3 and cv2.groupRectangles outputs
    on small fixed-seed inputs -- the external (OpenCV) arithmetic the reference calls at
    tempcv.cpp:1301-1302,160 pinned independently of our own restatement.
  * refsc.npz : REF-SC (scale-cascade mode) oracle outputs for four cascade kinds, same layout.
  * refsi_<cascade>.npz : REF-SI oracle outputs (raw rects, exit-code histogram, CRC of the
    exit-code map, stats) for two 320x240 frames per cascade -- regression pins for the oracle
    and golden vectors for the CUDA path.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

import oracle  # noqa: E402
from clfacedetection_b200.frames import octave_frame, uniform_frame  # noqa: E402

CASCADES = ["frontalface_alt", "frontalface_default", "frontalface_alt_tree", "eye", "profileface", "fullbody",
            "frontalface_alt2", "eye_tree_eyeglasses", "mcs_nose"]


def main():
    pins = {}
    src = octave_frame(160, 120, 11)
    pins["resize_src"] = src
    for i, (dw, dh) in enumerate([(133, 100), (111, 83), (77, 58), (160, 120), (20, 15)]):
        pins[f"resize_{i}_{dw}x{dh}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    img = uniform_frame(61, 47, 12)
    s, q, t = cv2.integral3(img)
    pins.update(integral_src=img, integral_sum=s.astype(np.int32), integral_sq=q.astype(np.float64),
                integral_tilted=t.astype(np.int32))
    rng = np.random.default_rng(5)
    base = rng.integers(0, 200, size=(4, 2))
    rects = np.array([[b[0] + rng.integers(-4, 5), b[1] + rng.integers(-4, 5), 40 + rng.integers(-2, 3), 40 + rng.integers(-2, 3)]
                      for b in base[rng.integers(0, 4, size=40)]], np.int32)
    g, w = cv2.groupRectangles(rects.tolist(), 2, 0.2)
    pins.update(group_in=rects, group_out=np.array(g, np.int32).reshape(-1, 4), group_w=np.array(w, np.int32).reshape(-1))
    np.savez_compressed(os.path.join(HERE, "opencv_pins.npz"), **pins)

    for name in CASCADES:
        cas = oracle.Cascade(os.path.join(ROOT, "data", "haarcascades", f"haarcascade_{name}.xml"))
        out = {}
        for fi, frame in enumerate([octave_frame(320, 240, 21), uniform_frame(320, 240, 22)]):
            r, codes, near, st, levels = cas.detect(frame, 1.2)
            out[f"rects_{fi}"] = r
            out[f"hist_{fi}"] = np.bincount(codes.astype(np.int64), minlength=128).astype(np.int64)
            out[f"crc_{fi}"] = np.array([zlib.crc32(codes.tobytes())], np.uint32)
            out[f"stats_{fi}"] = np.array([st.windows, st.weak_evals, st.node_evals, st.accepted, st.near_stage_thr], np.int64)
            out[f"levels_{fi}"] = np.array([[l.img_w, l.img_h, l.win_w, l.win_h, l.ystep, l.nx, l.ny] for l in levels], np.int32)
        np.savez_compressed(os.path.join(HERE, f"refsi_{name}.npz"), **out)
        print(name, {k: out[k].tolist() for k in ("stats_0", "stats_1")})

    # REF-SC (scale-cascade path, tempcv.cpp:1330-1456): one file for four cascade kinds
    out = {}
    for name in ("frontalface_alt", "frontalface_alt2", "frontalface_alt_tree", "fullbody"):
        cas = oracle.Cascade(os.path.join(ROOT, "data", "haarcascades", f"haarcascade_{name}.xml"))
        for fi, frame in enumerate([octave_frame(320, 240, 21), uniform_frame(320, 240, 22)]):
            r, codes, st, levels = cas.detect_sc(frame, 1.2)
            out[f"{name}_rects_{fi}"] = r
            out[f"{name}_crc_{fi}"] = np.array([zlib.crc32(codes.tobytes())], np.uint32)
            out[f"{name}_stats_{fi}"] = np.array([st.windows, st.weak_evals, st.accepted, int((codes == -32768).sum()),
                                                  int((codes == -32767).sum())], np.int64)
            out[f"{name}_levels_{fi}"] = np.array([[l.win_w, l.win_h, l.nx, l.ny] for l in levels], np.int32)
    np.savez_compressed(os.path.join(HERE, "refsc.npz"), **out)


if __name__ == "__main__":
    main()
