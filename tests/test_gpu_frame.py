"""GPU parity of the Frame post-processing (frame.cpp:36-133: undistortKeyPoints, findDepth, assignFeaturesToGrid;
SURVEY section 8f rank 3) against the oracle port, which is itself pinned to cv2.undistortPoints
(tests/golden/cv2_undistort.npz)."""
import os

import numpy as np
import pytest

import oracle
from vo_slam_test_b200 import synth

TUM1 = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3, dist=[0.2624, -0.9531, -0.0054, 0.0026, 1.1633], bf=40.0,
            bounds=(0.0, 640.0, 0.0, 480.0))


def _vo_cam(vo, cam):
    return vo.camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["dist"], cam["bf"], cam["bounds"])


def _same(a, b):
    return a.tobytes() == b.tobytes() or np.array_equal(a, b, equal_nan=True)


def _compare(got, want, n, f=0):
    un, ur, dp, start, ids = got
    wun, wur, wdp, wstart, wids = want
    for name in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert _same(un[f, :n][name], wun[name]), name
    assert _same(ur[f, :n], wur) and _same(dp[f, :n], wdp)
    assert np.array_equal(start[f], wstart)
    assert np.array_equal(ids[f, :wstart[-1]], wids)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tum1", "four", "rational", "strong"])
def test_undistort_golden_gpu(name):
    """The CUDA kernel against the committed cv2 4.13.0 outputs directly (no oracle in between)."""
    import vo_slam_test_b200 as vo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_undistort.npz"))
    pts = g["pts"]; K = g["K"]; n = len(pts)
    kps = np.zeros((1, n), vo.KP_DTYPE); kps["x"][0] = pts[:, 0]; kps["y"][0] = pts[:, 1]
    cam = vo.camera(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), [float(x) for x in g["D_" + name]], 40.0,
                    (0.0, 640.0, 0.0, 480.0))
    un = vo.frame_finish(kps, [n], cam)[0]
    out = np.stack([un["x"][0], un["y"][0]], 1)
    assert np.array_equal(out, g["out_" + name], equal_nan=True)


@pytest.mark.gpu
def test_frame_finish_batch_matches_oracle():
    import vo_slam_test_b200 as vo
    P = oracle.Port()
    ex = vo.ORBextractor()
    imgs = np.stack([synth.make_frame(300 + i) for i in range(6)])
    imgs[4] = 90                                                   # a frame without keypoints inside the batch
    kps, desc, counts = ex.extract_batch(imgs)
    rng = np.random.default_rng(3)
    depth = rng.uniform(0.3, 9.0, (6, 480, 640)).astype(np.float32)
    depth[rng.random(depth.shape) < 0.25] = 0.0
    depth[rng.random(depth.shape) < 0.05] = -1.0
    for cam in (TUM1, dict(TUM1, dist=[0.0, 0.0, 0.0, 0.0]), dict(TUM1, dist=[-0.28, 0.07, 0.0002, 0.00002], bounds=(-8.0, 652.0, -5.0, 489.0))):
        got = vo.frame_finish(kps, counts, _vo_cam(vo, cam), depth)
        for f in range(6):
            n = int(counts[f])
            want = P.frame_finish(kps[f, :n], cam, depth[f])
            _compare(got, want, n, f)
        assert counts[4] == 0 and got[3][4, -1] == 0
    # no depth image: findDepth leaves -1 everywhere
    got = vo.frame_finish(kps, counts, _vo_cam(vo, TUM1), None)
    n0 = int(counts[0])
    assert np.all(got[1][0, :n0] == -1) and np.all(got[2][0, :n0] == -1)
    _compare(got, P.frame_finish(kps[0, :n0], TUM1, None), n0, 0)
    ex.close()


@pytest.mark.gpu
def test_frame_finish_grid_equals_grid_build_and_feeds_projection():
    """The CSR produced on the undistorted points is the one orbx_grid_build makes from them (frame.cpp:72-89)."""
    import vo_slam_test_b200 as vo
    ex = vo.ORBextractor()
    k, d = ex(synth.make_frame(17))
    un, ur, dp, start, ids = vo.frame_finish(k[None], [len(k)], _vo_cam(vo, TUM1), None)
    s2, i2 = vo.grid_build(un[0, :len(k)], TUM1["bounds"])
    assert np.array_equal(start[0], s2) and np.array_equal(ids[0, :s2[-1]], i2)
    assert s2[-1] <= len(k) and s2[-1] > 0.9 * len(k)
    ex.close()


@pytest.mark.gpu
def test_frame_finish_errors():
    import ctypes as C
    import vo_slam_test_b200 as vo
    L = vo.lib()
    cam = _vo_cam(vo, TUM1)
    kps = np.zeros((1, 8), vo.KP_DTYPE); cnt = np.zeros(1, np.int32)
    o = [np.zeros((1, 8), vo.KP_DTYPE), np.zeros(8, np.float32), np.zeros(8, np.float32), np.zeros(3073, np.int32), np.zeros(8, np.int32)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    args = lambda c, k: (c, k, p(cnt), 1, 8, None, 0, 0, 0, 0, p(o[0]), p(o[1]), p(o[2]), p(o[3]), p(o[4]), 0)
    assert L.orbx_frame_finish(*args(C.byref(cam), p(kps))) == 0
    assert L.orbx_frame_finish(*args(None, p(kps))) == -1
    assert L.orbx_frame_finish(*args(C.byref(cam), None)) == -1
    bad = _vo_cam(vo, dict(TUM1, bounds=(0.0, 0.0, 0.0, 480.0)))
    assert L.orbx_frame_finish(*args(C.byref(bad), p(kps))) == -1
    depth = np.zeros((4, 4), np.float32)
    assert L.orbx_frame_finish(C.byref(cam), p(kps), p(cnt), 1, 8, p(depth), 4, 4, 6, 0, p(o[0]), p(o[1]), p(o[2]), p(o[3]), p(o[4]), 0) == -1
