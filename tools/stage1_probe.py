"""Per-stage device time of ONE frame per call (the latency chain): events around every stage, single stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
imgs = torch.from_numpy(synth.make_sequence(n, seed=3)).cuda()
ex = vo.ORBextractor()
cap = ex.max_keypoints
kps = torch.zeros((n, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
acc = []
for r in range(30):
    ms = ex.profile_stages(imgs.data_ptr(), n, 640, 480, 640, 640 * 480, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if r >= 10: acc.append(ms.copy())
m = np.median(np.stack(acc), axis=0) * 1e3
print("us per stage (pyramid, fast, quadtree, blur, orient_desc, total):", np.round(m, 1))
