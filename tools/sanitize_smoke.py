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
# round 2: resident frame (frame_finish_kernel with feat / angle outputs, patch_uright_kernel, packed search I/O), uploaded frame,
# frame-to-frame pairs kernel and the opt-in IMMA Hamming kernels (pairs + split map scan)
import torch
from vo_slam_test_b200 import api
cam = vo.camera(300.0, 300.0, 160.0, 120.0, [0.1, -0.2, 0.001, 0.002, 0.05], 40.0, (0.0, 320.0, 0.0, 240.0))
depth = np.random.default_rng(1).uniform(0.5, 6.0, (240, 320)).astype(np.float32)
fr = vo.Frame(ex, cam, img, depth)
want_un = P.frame_finish(rk, dict(fx=300.0, fy=300.0, cx=160.0, cy=120.0, dist=[0.1, -0.2, 0.001, 0.002, 0.05], bf=40.0, bounds=(0.0, 320.0, 0.0, 240.0)), depth)
assert fr.unkps.tobytes() == want_un[0].tobytes() and fr.uright.tobytes() == want_un[1].tobytes()
frame, pts = synth.make_projection_case(fr.unkps, fr.desc, sf, 600, seed=4, W=320, H=240)
frame["uright"] = fr.uright
a1, c1 = M.searchByProjectionH(fr, frame["occupied0"], pts, 15.0)
a0, c0 = oracle.Port().sbp_frame(frame, pts, 15.0)
assert c1 == c0 and np.array_equal(a1, a0)
Bf = synth.make_bow_side(fr.desc, fr.unkps["angle"], None, 5, 2)
m1, n1 = M.searchByBoWH(A, fr, Bf)
m0, n0 = oracle.Port().search_by_bow(A, Bf, 0, 0.7, 50, True)
assert n1 == n0 and np.array_equal(m1, m0)
up = vo.UploadedFrame(frame)
a1, c1 = M.searchByProjectionH(up, frame["occupied0"], pts, 15.0)
assert c1 == c0 and np.array_equal(a1, a0)
up.close(); fr.close()
dev = torch.device("cuda", 0)
cap = kk.shape[1]
d_desc = torch.from_numpy(dd).to(dev); d_cnt = torch.from_numpy(cc.astype(np.int32)).to(dev)
qf = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev); tf = torch.tensor([1, 2, 0], dtype=torch.int32, device=dev)
for variant in (0, 1):
    api.set_hamming_variant(variant)
    o = [torch.zeros((3, cap), dtype=torch.int32, device=dev) for _ in range(3)] + [torch.zeros((3, cap), dtype=torch.uint8, device=dev)]
    api.knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), cap, qf.data_ptr(), tf.data_ptr(), 3, 50, 0.7, o[0].data_ptr(), o[1].data_ptr(),
                          o[2].data_ptr(), o[3].data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for p_ in range(3):
        a, b = int(qf[p_]), int(tf[p_])
        w = oracle.Port().knn2(dd[a, :cc[a]], dd[b, :cc[b]], 50, 0.7)
        assert all(np.array_equal(x[p_, :cc[a]].cpu().numpy(), y) for x, y in zip(o, w)), (variant, p_)
    got = M.knn2(d[:100], t)                    # split scan + merge
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
api.set_hamming_variant(0)
ex.close()
print("sanitize smoke ok")
