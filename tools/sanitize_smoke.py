"""Small end-to-end invocation of every kernel, meant to be run under compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth

P = oracle.Port(300)
img = synth.make_frame(7, 240, 320)
ex = vo.ORBextractor(300)
k, d = ex(img)
rk, rd = P.extract(img)
assert np.array_equal(k, rk) and np.array_equal(d, rd)
img2 = synth.make_frame(8, 240, 320)
k2, d2 = ex(img2)
imgs = np.stack([img, img2, synth.make_frame(9, 240, 320)])
kk, dd, cc = ex.extract_batch(imgs)
assert cc[0] == len(rk)
M = vo.Matcher(0.7)
got = M.knn2(d, d2)
want = oracle.Port().knn2(d, d2, 50, 0.7)
assert all(np.array_equal(a, b) for a, b in zip(got, want))
t = synth.make_descriptors(5000, 3)
got = M.knn2(d[:100], t)
want = oracle.Port().knn2(d[:100], t, 50, 0.7)
assert all(np.array_equal(a, b) for a, b in zip(got, want))
sf = P.tables()[0]
frame, pts = synth.make_projection_case(k, d, sf, 800, seed=1, W=320, H=240)
a1, c1 = M.searchByProjection(frame, pts, 15.0)
a0, c0 = oracle.Port().sbp_frame(frame, pts, 15.0)
assert c1 == c0 and np.array_equal(a1, a0)
frame, pts = synth.make_projection_case(k, d, sf, 800, seed=2, W=320, H=240, local=True)
a1, c1 = M.searchByProjectionLocal(frame, pts, 3.0)
a0, c0 = oracle.Port().sbp_local(frame, pts, 3.0, 0.7)
assert c1 == c0 and np.array_equal(a1, a0)
A = synth.make_bow_side(d, k["angle"], None, 5, 1)
B = synth.make_bow_side(d2, k2["angle"], None, 5, 2)
m1, n1 = M.searchByBoW(A, B, mode=0)
m0, n0 = oracle.Port().search_by_bow(A, B, 0, 0.7, 50, True)
assert n1 == n0 and np.array_equal(m1, m0)
ex.close()
print("sanitize smoke ok")
