"""Where the tile kernel's warp slots go (run on the GPU box with a library built with -DCLFD_TILE_TIMING:
   make -C clfacedetection_b200/csrc EXTRA=-DCLFD_TILE_TIMING OUT=../../ab/libclfd_b200_timing.so ../../ab/libclfd_b200_timing.so
   CLFD_LIB=ab/libclfd_b200_timing.so python tools/tile_timing.py [cascade] [n_frames]).
That build replaces three counters of clfd_run_stats by cycle sums of k_cascade_tiles (both launches):
   exact_stage_evals     = sum over CTAs and warps of the cycles until the warp left the kernel   (busy)
   near_threshold_events = sum over CTAs of 8 x the cycles of the CTA's last warp                 (slots held)
   deep_windows          = sum over CTAs and warps of the cycles until the hand-over              (staging + phase 1)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clfacedetection_b200 as clfd  # noqa: E402
from clfacedetection_b200.frames import octave_frame  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "frontalface_alt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
frames = np.stack([octave_frame(1920, 1080, i) for i in range(n)])
ctx = clfd.Context(0)
cas = clfd.Cascade(os.path.join(here, "data", "haarcascades", f"haarcascade_{name}.xml"))
det = clfd.Detector(ctx, cas, 1920, 1080, max_batch=n, scale_factor=1.2)
for _ in range(3):
    res = det.detect(frames)
busy, held, p1 = (res.stats[k] for k in ("exact_stage_evals", "near_threshold_events", "deep_windows"))
print(f"{name}: warp-slot cycles held {held:.4g}, busy {busy:.4g} = {busy / held:.3f} of held "
      f"(idle inside CTAs {1 - busy / held:.3f}); staging + phase 1 {p1 / held:.3f} of held, {p1 / busy:.3f} of busy")
det.close()
