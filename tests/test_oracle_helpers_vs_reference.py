"""Oracle pinning, layer 3: the small Frame / KeyFrame / MapPoint / Camera helpers.

frame.cpp, keyframe.cpp, mappoint.cpp and camera.cpp do not compile as a whole here (Eigen / Sophus / DBoW3), but the helpers
the matcher path depends on are self-contained member functions.  oracle/Makefile pulls those line ranges out of the read-only
reference tree at build time and compiles them, unmodified, inside oracle/ref_helpers_wrap.cpp
(oracle/_ref/librefhelpers.so).  Here the oracle port's restatements are held against that code on random inputs:
assignFeaturesToGrid + getFeaturesInArea (Frame and KeyFrame), findDepth, undistortKeyPoints, MapPoint::computeDescriptor."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "librefhelpers.so")

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) or os.path.exists("/root/reference/src/frame.cpp")),
                                reason="oracle/_ref/librefhelpers.so not built and /root/reference absent")


@pytest.fixture(scope="module")
def ref():
    oracle.build()
    L = C.CDLL(LIB)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.refh_grid_build.argtypes = [vp, i32, f32, f32, f32, f32, vp, vp]
    L.refh_features_in_area.argtypes = [vp, i32, f32, f32, f32, f32, vp, vp, vp, vp, vp, i32, vp, vp, i32]
    L.refh_find_depth.argtypes = [vp, vp, i32, vp, i32, i32, f32, vp, vp]
    L.refh_medoid.argtypes = [vp, vp, i32, vp]
    L.refh_undistort.argtypes = [vp, i32, f32, f32, f32, f32, vp, i32, vp]
    return L


def _keypoints(rng, n, W, H, margin):
    k = np.zeros(n, oracle.KP_DTYPE)
    k["x"] = rng.uniform(-margin, W + margin, n).astype(np.float32)      # some outside the image: dropped by postionInGrad
    k["y"] = rng.uniform(-margin, H + margin, n).astype(np.float32)
    k["x"][: n // 10] = np.round(k["x"][: n // 10])                      # exact pixel centres and cell borders
    k["y"][: n // 10] = np.round(k["y"][: n // 10] / 10) * 10
    k["octave"] = rng.integers(0, 8, n)
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    return k


@pytest.mark.parametrize("seed,W,H,n", [(1, 640, 480, 1000), (2, 640, 480, 3000), (3, 1241, 376, 2000), (4, 752, 480, 50), (5, 640, 480, 0)])
def test_grid_and_window_queries_vs_reference(port, ref, seed, W, H, n):
    rng = np.random.default_rng(seed)
    kps = _keypoints(rng, n, W, H, 12.0)
    start = np.empty(64 * 48 + 1, np.int32); ids = np.empty(max(n, 1), np.int32)
    m = ref.refh_grid_build(kps.ctypes.data, n, 0.0, float(W), 0.0, float(H), start.ctypes.data, ids.ctypes.data)
    ps, pi = port.grid_build(kps, 0.0, float(W), 0.0, float(H))
    assert np.array_equal(ps, start) and np.array_equal(pi, ids[:m])
    nq = 400
    u = rng.uniform(-30, W + 30, nq).astype(np.float32); v = rng.uniform(-30, H + 30, nq).astype(np.float32)
    r = rng.choice([3.0, 7.5, 15.0, 31.1, 60.0, 200.0], nq).astype(np.float32)
    lo = rng.integers(-1, 7, nq).astype(np.int32); hi = (lo + rng.integers(0, 3, nq)).astype(np.int32)
    kf = rng.random(nq) < 0.3
    lo_ref = np.where(kf, -1, np.maximum(lo, 0)).astype(np.int32)
    out = np.empty(max(n, 1) * nq, np.int32); ostart = np.empty(nq + 1, np.int32)
    ref.refh_features_in_area(kps.ctypes.data, n, 0.0, float(W), 0.0, float(H), u.ctypes.data, v.ctypes.data, r.ctypes.data,
                              lo_ref.ctypes.data, hi.ctypes.data, nq, out.ctypes.data, ostart.ctypes.data, len(out))
    hits = 0
    for q in range(nq):
        want = out[ostart[q]:ostart[q + 1]]
        got = port.features_in_area(kps, (0.0, float(W), 0.0, float(H)), float(u[q]), float(v[q]), float(r[q]),
                                    0 if kf[q] else int(max(lo[q], 0)), 100 if kf[q] else int(hi[q]))   # KeyFrame version: no level filter
        assert np.array_equal(got, want), (q, u[q], v[q], r[q], lo[q], hi[q], kf[q])
        hits += len(want)
    assert n == 0 or hits > nq


def test_find_depth_vs_reference(port, ref):
    rng = np.random.default_rng(7)
    W, H, n = 640, 480, 1500
    kps = _keypoints(rng, n, W, H, 0.0)
    kps["x"] = np.minimum(kps["x"], W - 1); kps["y"] = np.minimum(kps["y"], H - 1)
    depth = np.where(rng.random((H, W)) < 0.2, 0.0, rng.uniform(0.3, 9.0, (H, W))).astype(np.float32)
    depth[rng.random((H, W)) < 0.02] = -1.0
    cam = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3, dist=[0.0, 0.0, 0.0, 0.0], bf=40.0, bounds=(0.0, float(W), 0.0, float(H)))
    un, ur, dp, _, _ = port.frame_finish(kps, cam, depth)
    assert np.array_equal(un["x"], kps["x"])                      # k1 == 0: undistortKeyPoints copies (frame.cpp:41-45)
    d_ref = np.empty(n, np.float32); ur_ref = np.empty(n, np.float32)
    unx = np.ascontiguousarray(un["x"])
    ref.refh_find_depth(kps.ctypes.data, unx.ctypes.data, n, depth.ctypes.data, W, H, 40.0, d_ref.ctypes.data, ur_ref.ctypes.data)
    assert np.array_equal(dp, d_ref) and np.array_equal(ur, ur_ref)
    assert (d_ref > 0).sum() > n // 2 and (d_ref == -1).sum() > n // 10


@pytest.mark.parametrize("dist", [[0.2624, -0.9531, -0.0054, 0.0026, 1.1633], [0.2312, -0.7849, -0.0033, -0.0001], [-0.28, 0.07, 0.0002, 0.00002],
                                  [0.0, 0.3, 0.001, 0.001]])
def test_undistort_keypoints_vs_reference(port, ref, dist):
    """Frame::undistortKeyPoints (frame.cpp:36-70) itself -- the k1 == 0 shortcut, the N x 2 <-> N x 1 x 2 reshapes, the
    float round trip -- around the cv2-pinned cv::undistortPoints arithmetic, against the port's frame_finish."""
    rng = np.random.default_rng(3)
    kps = _keypoints(rng, 2000, 640, 480, 0.0)
    cam = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3, dist=dist, bf=40.0, bounds=(0.0, 640.0, 0.0, 480.0))
    un, _, _, _, _ = port.frame_finish(kps, cam, None)
    out = np.empty_like(kps)
    d = np.asarray(dist, np.float32)
    ref.refh_undistort(kps.ctypes.data, len(kps), 517.3, 516.5, 318.6, 255.3, d.ctypes.data, len(d), out.ctypes.data)
    assert np.array_equal(un, out)
    moved = int((out["x"] != kps["x"]).sum())
    assert moved == 0 if dist[0] == 0.0 else moved > 1500


def test_compute_descriptor_vs_reference(port, ref):
    """MapPoint::computeDescriptor (mappoint.cpp:118-179) itself, one call per map point, against the batched port."""
    rng = np.random.default_rng(11)
    counts = np.concatenate([rng.integers(1, 25, 300), [0, 1, 2, 3, 40]])
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    base = rng.integers(0, 256, (len(counts), 32), np.uint8)
    desc = np.repeat(base, counts, axis=0)
    flips = rng.random((len(desc), 256)) < 0.12                   # observations of one point: noisy copies of its descriptor
    desc ^= np.packbits(flips, axis=1, bitorder="little")
    desc[start[5]:start[5] + 2] = desc[start[5]]                  # exact duplicates: first wins
    best_ref = np.empty(len(counts), np.int32)
    ref.refh_medoid(desc.ctypes.data, start.ctypes.data, len(counts), best_ref.ctypes.data)
    best = port.medoid(desc, start)
    # the reference hands back a descriptor, not an index: compare descriptors (twins are interchangeable)
    for p in range(len(counts)):
        if counts[p] == 0:
            assert best[p] == -1 and best_ref[p] == -1
        else:
            assert np.array_equal(desc[start[p] + best[p]], desc[start[p] + best_ref[p]]), p


def test_stand_in_types_vs_reference_helpers(ref, tmp_path):
    """The helpers that oracle/compat_myslam/myslam_stub.hpp restates (the reference's matcher.cpp, compiled in place, runs on top
    of them): assignFeaturesToGrid, both getFeaturesInArea, isInImg, both predictScale, camera2pixel -- against the reference's own
    code for each, on random inputs (tests/tools/helpers_check.cpp)."""
    import subprocess
    exe = str(tmp_path / "hcheck")
    rdir = os.path.join(ROOT, "oracle", "_ref")
    r = subprocess.run(["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "oracle", "compat_myslam"), "-I" + os.path.join(ROOT, "oracle", "compat"),
                        os.path.join(ROOT, "tests", "tools", "helpers_check.cpp"), "-L" + rdir, "-lrefhelpers", "-Wl,-rpath," + rdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "agree with the reference's own code" in r.stdout, r.stdout + r.stderr
