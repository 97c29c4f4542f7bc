"""Randomised parity sweep over the latency-mode paths (two-stream chain, CUDA-graph capture / replay / invalidation, resident frames,
zero-copy searches): random image sizes and extractor parameters, four frames per geometry through orbx_extract and through
orbx_frame_create, every result against the CPU oracle.   python tools/fuzz_latency_paths.py [configs] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = 0
for c in range(N):
    W = int(rng.integers(160, 900)); H = int(rng.integers(120, 700))
    nf = int(rng.choice([100, 300, 500, 1000, 1500])); nl = int(rng.integers(2, 9)); sf = float(rng.choice([1.1, 1.2, 1.3, 1.5]))
    ini = int(rng.choice([12, 20, 30])); mn = int(rng.choice([5, 7]))
    try:
        ex = vo.ORBextractor(nf, sf, nl, ini, mn)
    except Exception as e:
        print("skip", W, H, nf, nl, sf, str(e)[:60]); continue
    P = oracle.Port(nf, sf, nl, ini, mn)
    cam = dict(fx=0.8 * W, fy=0.8 * W, cx=W / 2 - 1.5, cy=H / 2 + 2.0, dist=[float(x) for x in rng.normal(0, [0.2, 0.3, 0.002, 0.002, 0.1])],
               bf=40.0, bounds=(0.0, float(W), 0.0, float(H)))
    camv = vo.camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["dist"], cam["bf"], cam["bounds"])
    ok = True
    try:
        for k in range(4):
            img = synth.make_frame(1000 * c + k, H, W)
            rk, rd = P.extract(img)
            gk, gd = ex(img)
            ok &= len(gk) == len(rk) and gk.tobytes() == rk.tobytes() and np.array_equal(gd, rd)
            depth = rng.uniform(0.3, 9.0, (H, W)).astype(np.float32)
            fr = vo.Frame(ex, camv, img, depth)
            wun, wur, wdp, wstart, wids = P.frame_finish(rk, cam, depth)
            ok &= fr.n == len(rk) and fr.kps.tobytes() == rk.tobytes() and np.array_equal(fr.desc, rd)
            ok &= fr.unkps.tobytes() == wun.tobytes() and fr.uright.tobytes() == wur.tobytes() and fr.depth.tobytes() == wdp.tobytes()
            if fr.n > 20:
                sfv = np.asarray(ex.GetScaleFactors(), np.float32)
                frame, pts = synth.make_projection_case(fr.unkps, fr.desc, sfv, 400, seed=k, W=W, H=H)
                frame["uright"] = fr.uright
                a1, c1 = vo.Matcher(0.9).searchByProjectionH(fr, frame["occupied0"], pts, 15.0)
                a0, c0 = oracle.Port().sbp_frame(frame, pts, 15.0)
                ok &= c1 == c0 and np.array_equal(a1, a0)
            fr.close()
    except Exception as e:           # noqa: BLE001
        if "error -4" in str(e):     # ORBX_ERR_SHAPE: a documented limit, not a parity failure
            print("skip (shape)", W, H, nf, nl, sf, str(e)[-60:]); ex.close(); continue
        print("error", W, H, nf, nl, sf, ini, mn, str(e)[:120]); ok = False
    print("ok  " if ok else "FAIL", W, H, nf, nl, sf, ini, mn, flush=True)
    bad += 0 if ok else 1
    ex.close()
print("configs failed:", bad)
sys.exit(1 if bad else 0)
