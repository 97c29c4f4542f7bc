"""Throughput-mode chunk (12 frames in one chunk: one stream, octree_kernel<false>) for compute-sanitizer, beside tools/sanitize_smoke.py
(which runs the latency-mode paths):  compute-sanitizer --tool racecheck python tools/sanitize_batch.py"""
import os, sys
sys.path.insert(0, os.getcwd())
os.environ["ORBX_CHUNK"] = "12"
import numpy as np, oracle
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth
imgs = np.stack([synth.make_frame(40 + f, 240, 320) for f in range(12)])
ex = vo.ORBextractor(300)
kk, dd, cc = ex.extract_batch(imgs)
P = oracle.Port(300)
rk, rd = P.extract(imgs[5])
assert cc[5] == len(rk) and np.array_equal(kk[5, :cc[5]], rk) and np.array_equal(dd[5, :cc[5]], rd)
print("batch-mode ok", ex.launch_count())
