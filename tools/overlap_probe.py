"""Probe: frame-to-frame matching of chunk c on a third stream while the two extraction lanes work on later chunks.
One handle per chunk (fixed level-0 base pointer, so no tensor-map re-encoding) emulates what a fused device entry point
would do.  usage: python tools/overlap_probe.py"""
import os
import sys

os.environ["ORBX_LANES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import api, synth

F, CH = 4096, 512
W, H = 640, 480
dev = torch.device("cuda", 0)
h = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
synth.make_sequence(F, seed=0, out=h.numpy())
d_imgs = h.to(dev)
NC = F // CH
exs = [vo.ORBextractor(1000, 1.2, 8, 20, 7, device=0) for _ in range(NC)]
cap = exs[0].max_keypoints
d_kps = torch.zeros((F, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.zeros((F, cap, 32), dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(F, dtype=torch.int32, device=dev)
npairs = F - 1
d_qf = torch.arange(0, F, dtype=torch.int32, device=dev)
d_midx = torch.empty((npairs, cap), dtype=torch.int32, device=dev)
d_md1 = torch.empty_like(d_midx); d_md2 = torch.empty_like(d_midx)
d_mok = torch.zeros((npairs, cap), dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream()
A, B, M = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def pairs(p0, p1, stream):
    if p1 <= p0:
        return
    api.knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), cap, d_qf.data_ptr() + 4 * p0, d_qf.data_ptr() + 4 * (p0 + 1), p1 - p0, 50, 0.7,
                          d_midx.data_ptr() + 4 * p0 * cap, d_md1.data_ptr() + 4 * p0 * cap, d_md2.data_ptr() + 4 * p0 * cap,
                          d_mok.data_ptr() + p0 * cap, stream.cuda_stream)


def chunk(c, stream):
    a = c * CH
    exs[c].extract_batch_device(d_imgs.data_ptr() + a * W * H, CH, W, H, W, W * H, d_kps.data_ptr() + a * cap * 28,
                                d_desc.data_ptr() + a * cap * 32, cap, d_cnt.data_ptr() + a * 4, stream.cuda_stream)
    e = torch.cuda.Event(); e.record(stream)
    return e


def step(mode):
    fork = torch.cuda.Event(); fork.record(main)
    half = NC // 2
    if mode == "serial":
        for c in range(NC):
            chunk(c, main)
        pairs(0, npairs, main)
        return
    A.wait_event(fork); B.wait_event(fork); M.wait_event(fork)
    evs = {}
    for i in range(half):
        evs[i] = chunk(i, A)
        evs[half + i] = chunk(half + i, B)
        if mode == "overlap":
            for c in (i, half + i):
                M.wait_event(evs[c])
                a = c * CH
                p0 = a if (c == 0 or c == half) else a - 1     # the pair across the lane boundary comes last
                pairs(p0, a + CH - 1, M)
    if mode == "overlap":
        pairs(half * CH - 1, half * CH, M)
        j = torch.cuda.Event(); j.record(M); main.wait_event(j)
    else:                                                   # "lanes": join, then all pairs
        for c in (half - 1, NC - 1):
            main.wait_event(evs[c])
        pairs(0, npairs, main)


def timeit(mode, reps=5):
    for _ in range(3):
        step(mode)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps):
        step(mode)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref = None
for mode in ("serial", "lanes", "overlap"):
    d_mok.zero_(); d_midx.zero_()
    t = timeit(mode)
    sig = (int(d_cnt.sum().item()), int(d_mok.sum().item()), int(d_midx.long().sum().item()))
    ref = ref or sig
    print("%-8s %.3f ms -> %.1f k frames/s   %s" % (mode, t, F / t, "same result" if sig == ref else "RESULT DIFFERS %s vs %s" % (sig, ref)))
