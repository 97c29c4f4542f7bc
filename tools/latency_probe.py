#!/usr/bin/env python
"""Single-frame latency of orbx_extract (host in, host out, synchronous) for a few frame sizes.

  [ORBX_PDL=0|1|2] python tools/latency_probe.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vo_slam_test_b200 as vo  # noqa: E402
from vo_slam_test_b200 import synth  # noqa: E402

for H, W, nf in [(480, 640, 1000), (1080, 1920, 2000)]:
    img = synth.make_frame(42, H, W)
    ex = vo.ORBextractor(nf, 1.2, 8, 20, 7)
    for _ in range(20):
        ex(img)
    t0 = time.perf_counter()
    reps = 300
    for _ in range(reps):
        kps, desc = ex(img)
    dt = (time.perf_counter() - t0) / reps
    print("%dx%d: %.3f ms per frame (%d keypoints)" % (W, H, dt * 1e3, len(kps)))
    ex.close()
