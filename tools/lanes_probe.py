"""Probe: does running the extractor as two concurrent lanes (two handles = two workspaces, two streams, half of the batch
each) beat one lane?  Kernel tails and the latency-bound stages (quadtree, orientation/descriptor) of one lane can then
fill under the throughput-bound stages of the other.  Prints ms per 4096-frame step for 1 lane and 2 lanes, with the
frame-to-frame pairs kernel after the join, and checks that the outputs are identical.
usage: python tools/lanes_probe.py [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import api, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W, H = 640, 480
dev = torch.device("cuda", 0)
h = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
synth.make_sequence(F, seed=0, out=h.numpy())
d_imgs = h.to(dev)
exs = [vo.ORBextractor(1000, 1.2, 8, 20, 7, device=0) for _ in range(3)]
cap = exs[0].max_keypoints


def outputs():
    return (torch.zeros((F, cap, 7), dtype=torch.float32, device=dev), torch.zeros((F, cap, 32), dtype=torch.uint8, device=dev),
            torch.zeros(F, dtype=torch.int32, device=dev))


npairs = F - 1
d_qf = torch.arange(0, npairs, dtype=torch.int32, device=dev); d_tf = d_qf + 1
d_midx = torch.empty((npairs, cap), dtype=torch.int32, device=dev)
d_md1 = torch.empty_like(d_midx); d_md2 = torch.empty_like(d_midx)
d_mok = torch.zeros((npairs, cap), dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream()
lanes = [torch.cuda.Stream(), torch.cuda.Stream()]


def pairs(o, stream):
    api.knn2_pairs_device(o[1].data_ptr(), o[2].data_ptr(), cap, d_qf.data_ptr(), d_tf.data_ptr(), npairs, 50, 0.7,
                          d_midx.data_ptr(), d_md1.data_ptr(), d_md2.data_ptr(), d_mok.data_ptr(), stream)


def one_lane(o, match=True):
    exs[0].extract_batch_device(d_imgs.data_ptr(), F, W, H, W, W * H, o[0].data_ptr(), o[1].data_ptr(), cap, o[2].data_ptr(), main.cuda_stream)
    if match:
        pairs(o, main.cuda_stream)


def two_lanes(o, match=True, split=None):
    split = split or F // 2
    fork = torch.cuda.Event(); fork.record(main)
    bounds = [(0, split), (split, F)]
    for k, (a, b) in enumerate(bounds):
        lanes[k].wait_event(fork)
        exs[1 + k].extract_batch_device(d_imgs.data_ptr() + a * W * H, b - a, W, H, W, W * H, o[0].data_ptr() + a * cap * 28,
                                        o[1].data_ptr() + a * cap * 32, cap, o[2].data_ptr() + a * 4, lanes[k].cuda_stream)
        j = torch.cuda.Event(); j.record(lanes[k]); main.wait_event(j)
    if match:
        pairs(o, main.cuda_stream)


def timeit(fn, o, reps=5, **kw):
    for _ in range(3):
        fn(o, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps):
        fn(o, **kw)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


o1, o2 = outputs(), outputs()
t1 = timeit(one_lane, o1); t1x = timeit(one_lane, o1, match=False)
t2 = timeit(two_lanes, o2); t2x = timeit(two_lanes, o2, match=False)
# rows at or beyond counts[f] are unspecified (include/orb_b200.h): compare the counts and the valid rows only
valid = (torch.arange(cap, device=dev)[None, :] < o1[2][:, None])
same = torch.equal(o1[2], o2[2]) and torch.equal(o1[0][valid], o2[0][valid]) and torch.equal(o1[1][valid], o2[1][valid])
print("frames %d  chunk %s" % (F, os.environ.get("ORBX_CHUNK", "default")))
print("1 lane : %.3f ms (extract only %.3f)  -> %.1f k frames/s" % (t1, t1x, F / t1))
print("2 lanes: %.3f ms (extract only %.3f)  -> %.1f k frames/s   outputs identical: %s" % (t2, t2x, F / t2, same))
