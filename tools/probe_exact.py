import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import clfacedetection_b200 as clfd
from clfacedetection_b200.frames import octave_frame
ctx = clfd.Context(0)
for name in sys.argv[1:]:
    cas = clfd.Cascade(f"data/haarcascades/haarcascade_{name}.xml")
    det = clfd.Detector(ctx, cas, 1920, 1080, max_batch=1, scale_factor=1.2, want_codes=True)
    r = det.detect(octave_frame(1920, 1080, 0)[None])
    print(name, {k: r.stats[k] for k in ("windows", "rects", "exact_stage_evals", "near_threshold_events", "deep_windows")})
    det.close()
