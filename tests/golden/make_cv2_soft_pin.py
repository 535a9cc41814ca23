#!/usr/bin/env python
"""cv2_detector_soft_pin.npz: raw (ungrouped, minNeighbors=0) candidates of OpenCV 4.x's own
cv2.CascadeClassifier on fixed-seed frames, for the soft pin in tests/test_oracle_pins.py.

OpenCV 4.x converts the old-format cascade and runs a different pipeline than the 2.4-era
cvHaarDetectObjects the reference calls (INTER_LINEAR_EXACT resize, float evaluator), so its
output is NOT bit-comparable -- but it is an independent implementation of the same detector,
and the oracle's raw candidate set must agree with it on the large majority of windows."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from clfacedetection_b200.frames import octave_frame  # noqa: E402

# OpenCV 4.x flattens the alt_tree stage tree when it converts the file (left out), and its evaluator
# drops the 0.5 weight correction the 2.4-era code applies to tilted features (tempcv.cpp:733): for the
# tilted cascades the test compares with the oracle's correction switched off as well.
CASCADES = ["frontalface_alt", "frontalface_default", "eye", "profileface", "frontalface_alt2",
            "fullbody", "mcs_nose", "eye_tree_eyeglasses"]
FRAMES = [(960, 540, 0), (960, 540, 7), (1280, 720, 3), (800, 600, 5)]


def main():
    out = {"cv2_version": np.array([int(v) for v in cv2.__version__.split(".")[:3]], np.int32)}
    for name in CASCADES:
        cc = cv2.CascadeClassifier(os.path.join(ROOT, "data", "haarcascades", f"haarcascade_{name}.xml"))
        assert not cc.empty(), name
        for fi, (w, h, seed) in enumerate(FRAMES):
            r = cc.detectMultiScale(octave_frame(w, h, seed), scaleFactor=1.2, minNeighbors=0)
            out[f"{name}_{fi}"] = np.asarray(r, np.int32).reshape(-1, 4)
            print(name, fi, len(out[f"{name}_{fi}"]))
            # accepted windows with the stage sum of the last stage (levelWeights of detectMultiScale3)
            r3, l3, w3 = cc.detectMultiScale3(octave_frame(w, h, seed), scaleFactor=1.2, minNeighbors=0, outputRejectLevels=True)
            out[f"{name}_{fi}_roc_rects"] = np.asarray(r3, np.int32).reshape(-1, 4)
            out[f"{name}_{fi}_roc_levels"] = np.asarray(l3, np.int32).reshape(-1)
            out[f"{name}_{fi}_roc_weights"] = np.asarray(w3, np.float64).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "cv2_detector_soft_pin.npz"), **out)


if __name__ == "__main__":
    main()
