"""Oracle pinning, layer 1: the C++ restatements of the OpenCV / libm primitives (oracle/cvprims.h)
against (a) fixtures produced by the REAL cv2 4.13.0 / glibc 2.39 (tests/golden, always run) and
(b) the live cv2 wheel when importable (wider sweep)."""
import os

import numpy as np
import pytest

import oracle
from vo_slam_test_b200 import synth


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "cv2_primitives.npz"))


def test_resize_golden(port, g):
    assert np.array_equal(port.resize(g["resize_src"], 109, 81), g["resize_dst"])
    assert np.array_equal(port.resize(g["frame"], 133, 100), g["frame_resize"])


def test_blur_golden(port, g):
    assert np.array_equal(port.blur(g["resize_src"]), g["blur_noise"])
    assert np.array_equal(port.blur(g["frame"]), g["blur_frame"])


def test_fast_golden(port, g):
    f = g["frame"]
    assert np.array_equal(port.fast(f, 20), g["fast20_frame"])
    assert np.array_equal(port.fast(f, 7), g["fast7_frame"])
    roi = f[30:73, 40:83]  # non-contiguous view, like the reference's rowRange().colRange()
    assert np.array_equal(port.fast(roi, 20), g["fast20_roi"])
    assert np.array_equal(port.fast(roi, 7), g["fast7_roi"])
    assert np.array_equal(port.fast(g["resize_src"], 7), g["fast7_noise"])
    assert len(g["fast7_frame"]) > 50


def test_fast_atan2_golden(port, g):
    got = np.array([port.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got.view(np.uint32), g["atan_deg"].view(np.uint32))
    assert port.fast_atan2(0, 0) == 0.0
    assert abs(port.fast_atan2(5, 5) - 44.990456) < 1e-5


def test_libm_sincos_golden(port, golden_dir):
    """The box's libm must be the glibc the fixtures were made with (descriptor rotation, ORBextractor.cpp:115)."""
    d = np.load(os.path.join(golden_dir, "libm_sincos.npz"))
    sc = np.array([port.sincosf(float(a)) for a in d["rad"]], np.float32)
    assert np.array_equal(sc[:, 0].view(np.uint32), d["sin"].view(np.uint32))
    assert np.array_equal(sc[:, 1].view(np.uint32), d["cos"].view(np.uint32))


def test_tables_match_survey(port):
    sc, inv, nf, um = port.tables()
    assert list(nf) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(um) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [port.level_size(640, 480, l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278),
                                                               (309, 231), (257, 193), (214, 161), (179, 134)]
    assert abs(float(sc[7]) - 3.5831816196) < 1e-6


# ---- live cv2 sweep (authoring container; skipped where cv2 is missing) -----------------------
cv2 = pytest.importorskip("cv2")


def _fast_cv2(img, th):
    k = cv2.FastFeatureDetector_create(th, True).detect(img)
    return np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in k], np.int32).reshape(-1, 3)


@pytest.mark.parametrize("shape", [(480, 640), (1080, 1920), (231, 309)])
def test_pyramid_chain_vs_cv2(port, shape):
    cv2.setNumThreads(1)
    H, W = shape
    prev = synth.make_frame(11, H, W)
    for l in range(1, 8):
        w, h = port.level_size(W, H, l)
        want = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(port.resize(prev, w, h), want), l
        assert np.array_equal(port.blur(want), cv2.GaussianBlur(want, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        prev = want


def test_fast_cells_vs_cv2(port):
    """Every cell ROI of one level, both thresholds, exactly as ORBextractor.cpp:796-836 slices them."""
    img = synth.make_frame(12)
    rng = np.random.default_rng(0)
    for _ in range(60):
        x0 = int(rng.integers(0, 600)); y0 = int(rng.integers(0, 440))
        w = int(rng.integers(7, 44)); h = int(rng.integers(7, 44))
        roi = img[y0:y0 + h, x0:x0 + w]
        for th in (20, 7):
            assert np.array_equal(port.fast(roi, th), _fast_cv2(roi, th))


def test_fast_tiny_rois(port):
    img = synth.make_frame(13)
    for (h, w) in [(6, 40), (40, 6), (7, 7), (3, 3), (8, 9)]:
        roi = img[100:100 + h, 100:100 + w]
        assert np.array_equal(port.fast(roi, 7), _fast_cv2(roi, 7))


def _cam(K, D, bf=40.0, bounds=(0.0, 640.0, 0.0, 480.0)):
    return dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]), dist=[float(x) for x in D], bf=bf,
                bounds=bounds)


@pytest.mark.parametrize("name", ["tum1", "four", "rational", "strong"])
def test_undistort_golden(port, golden_dir, name):
    """cv::undistortPoints(.., K, D, noArray(), K) of frame.cpp:58, pinned to cv2 4.13.0 (bit-exact, NaNs alike)."""
    g = np.load(os.path.join(golden_dir, "cv2_undistort.npz"))
    kps = np.zeros((len(g["pts"]), 7), np.float32); kps[:, :2] = g["pts"]
    un = port.frame_finish(kps, _cam(g["K"], g["D_" + name]))[0]
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), g["out_" + name], equal_nan=True)
    assert np.array_equal(un["size"], kps[:, 2]) and np.array_equal(un["octave"].view(np.float32), kps[:, 5])


def test_undistort_live_cv2(port):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(-20, 1300, 50000), rng.uniform(-20, 740, 50000)], 1).astype(np.float32)
    K = np.array([[718.856, 0, 607.1928], [0, 718.2, 185.2157], [0, 0, 1]], np.float32)
    D = np.array([-0.3, 0.12, 0.0007, -0.0003, -0.02], np.float32)
    kps = np.zeros((len(pts), 7), np.float32); kps[:, :2] = pts
    un = port.frame_finish(kps, _cam(K, D, bounds=(0.0, 1280.0, 0.0, 720.0)))[0]
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), ref, equal_nan=True)


def test_frame_finish_depth_and_grid(port):
    """findDepth (frame.cpp:108-133) and the grid on the UNDISTORTED points (frame.cpp:72-89), restated in numpy."""
    rng = np.random.default_rng(9)
    kps, _ = port.extract(synth.make_frame(11))
    depth = rng.uniform(0.3, 8.0, (480, 640)).astype(np.float32)
    depth[rng.random((480, 640)) < 0.2] = 0.0
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    for D in ([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], [0.0, 0.0, 0.0, 0.0]):
        cam = _cam(K, np.array(D, np.float32))
        un, ur, dp, start, ids = port.frame_finish(kps, cam, depth[:, :])
        if D[0] == 0.0:
            assert np.array_equal(un, kps)                                          # frame.cpp:41-45
        d = depth[kps["y"].astype(np.int32), kps["x"].astype(np.int32)]
        pos = d > 0
        assert np.array_equal(dp, np.where(pos, d, np.float32(-1)))
        exp_ur = np.where(pos, un["x"] - np.float32(cam["bf"]) / np.where(pos, d, np.float32(1)), np.float32(-1)).astype(np.float32)
        assert np.array_equal(ur, exp_ur)
        s2, i2 = port.grid_build(un, *cam["bounds"])
        assert np.array_equal(start, s2) and np.array_equal(ids, i2)
