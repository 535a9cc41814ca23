"""Time the detector on frames with large flat regions (black letterbox bars, a saturated half, an all-black frame):
flat windows have sigma = 0 and feature sums of exactly 0, the corner case of the FP32 filters."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import clfacedetection_b200 as clfd
from clfacedetection_b200.frames import octave_frame
ctx = clfd.Context(0)
name = sys.argv[1] if len(sys.argv) > 1 else "frontalface_alt"
cas = clfd.Cascade(f"data/haarcascades/haarcascade_{name}.xml")
B = 16
det = clfd.Detector(ctx, cas, 1920, 1080, max_batch=B, scale_factor=1.2)
base = np.stack([octave_frame(1920, 1080, i) for i in range(B)])
kinds = {"noise": base}
lb = base.copy(); lb[:, :140] = 0; lb[:, -140:] = 0; kinds["letterbox (26 % black)"] = lb
sat = base.copy(); sat[:, :, 960:] = 255; kinds["right half 255"] = sat
kinds["all black"] = np.zeros_like(base)
kinds["all 128"] = np.full_like(base, 128)
for k, fr in kinds.items():
    pinned = torch.from_numpy(fr).pin_memory()
    for _ in range(2):
        r = det.detect(pinned)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        r = det.detect(pinned)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:20s} {k:24s} {B / dt:8.1f} frames/s  rects {len(r.rects)}")
