"""Per-stage device time of a resident batch at a given resolution / feature count (CUDA events inside the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth
W, H, NF, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
imgs = torch.from_numpy(np.stack([synth.make_frame(10 + f, H, W) for f in range(min(n, 8))])).cuda()
imgs = imgs.repeat((n + imgs.shape[0] - 1) // imgs.shape[0], 1, 1)[:n].contiguous()
ex = vo.ORBextractor(NF)
cap = ex.max_keypoints
kps = torch.zeros((n, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
acc = []
for r in range(8):
    ms = ex.profile_stages(imgs.data_ptr(), n, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if r >= 3: acc.append(ms.copy())
m = np.median(np.stack(acc), axis=0)
print("%dx%d nfeat %d, %d frames: ms (pyramid, fast, quadtree, blur, orient_desc, total):" % (W, H, NF, n), np.round(m, 3), "kp/frame", float(cnt.float().mean()))
