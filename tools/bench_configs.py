#!/usr/bin/env python
"""Side benchmark of the BASELINE.json configs that bench.py does not headline (configs 1, 3, 4; config 5 shape is in
bench.py's `hamming_map`).  One GPU.  Prints one JSON object; results are copied into profiles/.

  python tools/bench_configs.py > gpurun_out/configs.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import oracle  # noqa: E402
import vo_slam_test_b200 as vo  # noqa: E402
from vo_slam_test_b200 import synth  # noqa: E402


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def resident_fps(W, H, nfeat, B):
    ex = vo.ORBextractor(nfeat)
    cap = ex.max_keypoints
    imgs = np.stack([synth.make_frame(100 + i, H, W) for i in range(min(B, 8))])
    imgs = np.concatenate([imgs] * ((B + len(imgs) - 1) // len(imgs)))[:B]
    d = torch.from_numpy(imgs).cuda()
    k = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda"); de = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
    c = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    t = timed(lambda: ex.extract_batch_device(d.data_ptr(), B, W, H, W, W * H, k.data_ptr(), de.data_ptr(), cap, c.data_ptr(), st), 5)
    ms = ex.profile_stages(d.data_ptr(), B, W, H, W, W * H, k.data_ptr(), de.data_ptr(), cap, c.data_ptr(), st)
    one = imgs[0]
    t1 = timed(lambda: ex(one), 20)
    cores = os.cpu_count() or 1
    R = oracle.Ref(nfeat, parity=False)
    n_cpu = max(cores, 8)
    t0 = time.perf_counter(); R.extract_batch(imgs[:n_cpu] if B >= n_cpu else np.concatenate([imgs] * n_cpu)[:n_cpu], cores, keep_outputs=False)
    tc = (time.perf_counter() - t0) / n_cpu
    ex.close()
    return {"frames_per_s_resident": B / t, "batch": B, "single_frame_ms": t1 * 1e3, "mean_keypoints": float(c.float().mean().item()),
            "stage_ms_per_batch": {n: float(v) for n, v in zip(["pyramid", "fast", "quadtree", "blur", "orient_desc", "total"], ms)},
            "cpu_reference_frames_per_s_all_cores": 1.0 / tc, "cpu_cores": cores}


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    out["config1_640x480_1000"] = resident_fps(640, 480, 1000, 512)
    out["config3_1920x1080_2000"] = resident_fps(1920, 1080, 2000, 128)
    out["config3_3840x2160_5000"] = resident_fps(3840, 2160, 5000, 32)
    # config 4: 10k map points projected into one 640x480 frame, radius 15
    P = oracle.Port()
    kps, desc = P.extract(synth.make_frame(42))
    sf = P.tables()[0]
    M = vo.Matcher(0.9)
    res = {}
    for name, kw in [("frame", {}), ("frame_stereo", {"stereo": True})]:
        frame, pts = synth.make_projection_case(kps, desc, sf, 10000, seed=1, **kw)
        t = timed(lambda: M.searchByProjection(frame, pts, 15.0), 20)
        t0 = time.perf_counter()
        for _ in range(20):
            P.sbp_frame(frame, pts, 15.0)
        tc = (time.perf_counter() - t0) / 20
        res[name] = {"gpu_ms_per_search_host_api": t * 1e3, "points_per_s": 10000 / t, "cpu_port_ms_1core": tc * 1e3}
    frame, pts = synth.make_projection_case(kps, desc, sf, 10000, seed=1, local=True)
    ML = vo.Matcher(0.8)
    t = timed(lambda: ML.searchByProjectionLocal(frame, pts, 3.0), 20)
    t0 = time.perf_counter()
    for _ in range(20):
        P.sbp_local(frame, pts, 3.0, 0.8)
    tc = (time.perf_counter() - t0) / 20
    res["local"] = {"gpu_ms_per_search_host_api": t * 1e3, "points_per_s": 10000 / t, "cpu_port_ms_1core": tc * 1e3}
    out["config4_projection_10k_points"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
