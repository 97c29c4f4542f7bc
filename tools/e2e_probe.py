"""Where the host-pipelined step spends its time: the same call with kernels and/or downloads switched off (ORBX_E2E_DEBUG)."""
import os, sys, time
import numpy as np
import torch
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
imgs = synth.make_sequence(B, seed=1)
pin = torch.from_numpy(imgs).pin_memory()
ex = vo.ORBextractor()
cap = ex.max_keypoints
def pinned(shape, dt): return torch.zeros(shape, dtype=dt).pin_memory()
kps = pinned((B, cap, 28), torch.uint8); desc = pinned((B, cap, 32), torch.uint8); cnt = pinned((B,), torch.int32)
midx = pinned((B - 1, cap), torch.int32); md1 = pinned((B - 1, cap), torch.int32); md2 = pinned((B - 1, cap), torch.int32)
mok = pinned((B - 1, cap), torch.uint8)
def run():
    ex.extract_match_batch(pin.data_ptr(), B, 640, 480, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr(), 50, 0.7,
                           midx.data_ptr(), md1.data_ptr(), md2.data_ptr(), mok.data_ptr())
for dbg in sys.argv[2:] or ["0", "1", "2", "3", "0"]:
    for kv in dbg.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
        else:
            os.environ["ORBX_E2E_DEBUG"] = kv
    for _ in range(2): run()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    print(dbg, "ms/step min %.2f median %.2f" % (min(ts) * 1e3, sorted(ts)[2] * 1e3), flush=True)
