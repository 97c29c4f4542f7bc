"""GPU parity of the device-resident frame (orbx_frame_t): Frame::Frame (frame.cpp:22-32) in one call and the tracking
thread's searches (visualOdometry.cpp:240,265,329,354) against the handle.  Everything is compared with the CPU oracle:
extractor output, undistorted keypoints, depth / uRight, and the assignments of the four searches -- which must also equal the
host-array entry points bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from vo_slam_test_b200 import synth

TUM1 = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3, dist=[0.2624, -0.9531, -0.0054, 0.0026, 1.1633], bf=40.0,
            bounds=(0.0, 640.0, 0.0, 480.0))


@pytest.fixture(scope="module")
def vo():
    import vo_slam_test_b200 as v
    assert v.device_count() > 0
    return v


def _cam(vo, cam):
    return vo.camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["dist"], cam["bf"], cam["bounds"])


def _depth(seed, H=480, W=640):
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.3, 9.0, (H, W)).astype(np.float32)
    d[rng.random(d.shape) < 0.25] = 0.0
    d[rng.random(d.shape) < 0.05] = -1.0
    return d


@pytest.mark.parametrize("with_depth,cam", [(True, TUM1), (False, TUM1), (True, dict(TUM1, dist=[0.0, 0.0, 0.0, 0.0])),
                                            (True, dict(TUM1, dist=[-0.28, 0.07, 0.0002, 0.00002], bounds=(-8.0, 652.0, -5.0, 489.0)))])
def test_frame_create_equals_oracle(vo, with_depth, cam):
    P = oracle.Port()
    ex = vo.ORBextractor()
    for seed in (501, 502, 503):                       # three frames through the same extractor: the block pool is reused
        img = synth.make_frame(seed)
        depth = _depth(seed) if with_depth else None
        fr = vo.Frame(ex, _cam(vo, cam), img, depth)
        rk, rd = P.extract(img)
        assert fr.n == len(rk)
        assert fr.kps.tobytes() == rk.tobytes() and np.array_equal(fr.desc, rd)
        wun, wur, wdp, wstart, wids = P.frame_finish(rk, cam, depth)
        assert fr.unkps.tobytes() == wun.tobytes()
        assert fr.uright.tobytes() == wur.tobytes() and fr.depth.tobytes() == wdp.tobytes()
        start, ids = fr.grid()
        assert np.array_equal(start, wstart) and np.array_equal(ids, wids[:wstart[-1]])
        fr.close()
    ex.close()


def test_frame_create_flat_image(vo):
    """No keypoints at all: n = 0, the searches return empty assignments."""
    ex = vo.ORBextractor()
    fr = vo.Frame(ex, _cam(vo, TUM1), np.full((480, 640), 90, np.uint8), None)
    assert fr.n == 0
    fr.close(); ex.close()


@pytest.mark.parametrize("stereo", [False, True])
def test_searches_on_the_handle_equal_oracle(vo, stereo):
    P = oracle.Port()
    ex = vo.ORBextractor()
    sf = np.asarray(ex.GetScaleFactors(), np.float32)
    img = synth.make_frame(611)
    fr = vo.Frame(ex, _cam(vo, TUM1), img, _depth(7) if stereo else None)
    if stereo:
        assert (fr.uright > 0).sum() > 100
    n = fr.n
    # --- searchByProjection(Frame*, Frame*) : visualOdometry.cpp:240 ---------------------------------------------------------
    frame, pts = synth.make_projection_case(fr.unkps, fr.desc, sf, 1500, seed=3)
    frame["uright"] = fr.uright                      # the handle's own uRight_ (depth-based), as the reference's Frame holds it
    M = vo.Matcher(0.9)
    for kw in ({}, {"forward": True}, {"backward": True}, {"checkRot": False}):
        okw = {("check_rot" if k == "checkRot" else k): v for k, v in kw.items()}
        want, wc = P.sbp_frame(frame, pts, 15.0, **okw)
        got, gc = M.searchByProjectionH(fr, frame["occupied0"], pts, 15.0, **kw)
        host, hc = M.searchByProjection(frame, pts, 15.0, **kw)
        assert gc == wc == hc and np.array_equal(got, want) and np.array_equal(host, want), kw
    assert wc > 200
    # --- searchByProjection(Frame*, local map) : visualOdometry.cpp:354 -----------------------------------------------------
    frameL, ptsL = synth.make_projection_case(fr.unkps, fr.desc, sf, 4000, seed=4, local=True)
    frameL["uright"] = fr.uright
    ML = vo.Matcher(0.8)
    want, wc = P.sbp_local(frameL, ptsL, 3.0, 0.8)
    got, gc = ML.searchByProjectionLocalH(fr, frameL["occupied0"], ptsL, 3.0)
    assert gc == wc and np.array_equal(got, want) and wc > 200
    # --- searchByProjection(Frame*, KeyFrame*) : relocalisation, visualOdometry.cpp:329 -----------------------------------------
    frameR, ptsR = synth.make_projection_case(fr.unkps, fr.desc, sf, 900, seed=5)
    frameR["uright"] = fr.uright
    for rot in (True, False):
        want, wc = P.sbp_reloc(frameR, ptsR, 15.0, 64.0, rot)
        got, gc = M.searchByProjectionKeyFrameH(fr, frameR["occupied0"], ptsR, 15.0, 64.0, rot)
        assert gc == wc and np.array_equal(got, want)
    assert wc > 100
    # --- searchByBoW(KeyFrame*, Frame*) : visualOdometry.cpp:265 ------------------------------------------------------------------
    rng = np.random.default_rng(9)
    m = 900
    src = rng.integers(0, n, m)
    kd = synth.flip_bits(fr.desc[src], rng.integers(0, 40, m), rng)
    kd[:, 0] = fr.desc[src][:, 0]; kd[:, 1] = (kd[:, 1] & 0x7F) | (fr.desc[src][:, 1] & 0x80)
    kang = ((fr.unkps["angle"][src] + rng.normal(0, 8, m)) % 360).astype(np.float32)
    A = synth.make_bow_side(kd, kang, (rng.random(m) > 0.1).astype(np.uint8), 9, 1)
    B = synth.make_bow_side(fr.desc, fr.unkps["angle"], None, 9, 2)
    for rot in (True, False):
        want, wc = P.search_by_bow(A, B, 0, 0.7, 50, rot)
        got, gc = vo.Matcher(0.7).searchByBoWH(A, fr, B, checkRot=rot)
        assert gc == wc and np.array_equal(got, want)
    assert wc > 100
    fr.close(); ex.close()


def test_uploaded_frame_searches_equal_oracle(vo):
    """orbx_frame_upload: a host-side frame (BASELINE config 4 shape: 10 k points, stereo on) made resident once, searched
    repeatedly."""
    P = oracle.Port()
    kps, desc = P.extract(synth.make_frame(42))
    sf = P.tables()[0]
    frame, pts = synth.make_projection_case(kps, desc, sf, 10000, seed=1, stereo=True)
    up = vo.UploadedFrame(frame)
    M = vo.Matcher(0.9)
    for rep in range(2):
        want, wc = P.sbp_frame(frame, pts, 15.0)
        got, gc = M.searchByProjectionH(up, frame["occupied0"], pts, 15.0)
        assert gc == wc and np.array_equal(got, want) and wc > 500
    frameL, ptsL = synth.make_projection_case(kps, desc, sf, 10000, seed=2, stereo=True, local=True)
    frameL["uright"] = frame["uright"]
    want, wc = P.sbp_local(frameL, ptsL, 3.0, 0.8)
    got, gc = vo.Matcher(0.8).searchByProjectionLocalH(up, frameL["occupied0"], ptsL, 3.0)
    assert gc == wc and np.array_equal(got, want)
    B = synth.make_bow_side(desc, kps["angle"], None, 9, 2)
    rng = np.random.default_rng(3)
    src = rng.integers(0, len(kps), 700)
    kd = synth.flip_bits(desc[src], rng.integers(0, 40, 700), rng)
    A = synth.make_bow_side(kd, kps["angle"][src], None, 9, 1)
    want, wc = P.search_by_bow(A, B, 0, 0.7, 50, True)
    got, gc = vo.Matcher(0.7).searchByBoWH(A, up, B)
    assert gc == wc and np.array_equal(got, want)
    up.close()


def test_two_live_frames_and_batch_use_of_the_same_extractor(vo):
    """The tracking thread keeps the last and the current frame alive; a batch call on the same extractor in between must not
    disturb them (the frames own their device blocks)."""
    P = oracle.Port()
    ex = vo.ORBextractor()
    sf = np.asarray(ex.GetScaleFactors(), np.float32)
    a = vo.Frame(ex, _cam(vo, TUM1), synth.make_frame(21), None)
    b = vo.Frame(ex, _cam(vo, TUM1), synth.make_frame(22), None)
    ex.extract_batch(np.stack([synth.make_frame(30 + i) for i in range(4)]))
    for fr in (a, b):
        frame, pts = synth.make_projection_case(fr.unkps, fr.desc, sf, 800, seed=6)
        want, wc = P.sbp_frame(frame, pts, 15.0)
        got, gc = vo.Matcher(0.9).searchByProjectionH(fr, frame["occupied0"], pts, 15.0)
        assert gc == wc and np.array_equal(got, want)
    a.close(); b.close(); ex.close()


def test_frame_handle_errors(vo):
    import ctypes as C
    L = vo.lib()
    ex = vo.ORBextractor()
    cam = _cam(vo, TUM1)
    img = synth.make_frame(1)
    h = C.c_void_p(); n = C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)     # noqa: E731
    assert L.orbx_frame_create(ex._h, None, p(img), 640, 480, 640, None, 0, C.byref(h), C.byref(n)) == -1
    assert L.orbx_frame_create(ex._h, C.byref(cam), p(img), 640, 480, 100, None, 0, C.byref(h), C.byref(n)) == -1      # stride < width
    assert L.orbx_frame_create(None, C.byref(cam), p(img), 640, 480, 640, None, 0, C.byref(h), C.byref(n)) == -1
    assert L.orbx_frame_create(ex._h, C.byref(cam), p(img), 640, 480, 640, None, 0, C.byref(h), C.byref(n)) == 0 and n.value > 900
    small = np.zeros(10, vo.KP_DTYPE)
    assert L.orbx_frame_get(h, p(small), None, None, None, None, 10) == -3                  # ORBX_ERR_CAPACITY
    assert L.orbx_search_by_projection_frame_h(h, None, None, 15.0, 40.0, 0, 0, 1, None, None) == -1
    assert L.orbx_frame_destroy(h) == 0 and L.orbx_frame_destroy(None) == 0
    ex.close()


def test_frame_create_graph_replay_and_invalidation(vo):
    """orbx_frame_create captures its device chain into a CUDA graph on the second frame built in a block and replays it from
    the third on.  The replay must give the oracle's bits, and a different camera, a different image size (new workspace) and
    the way back must each re-capture instead of replaying a stale graph."""
    P = oracle.Port()
    ex = vo.ORBextractor()
    camB = dict(TUM1, dist=[-0.28, 0.07, 0.0002, 0.00002], bounds=(-8.0, 652.0, -5.0, 489.0))
    small = dict(fx=258.6, fy=258.2, cx=159.3, cy=127.6, dist=TUM1["dist"], bf=20.0, bounds=(0.0, 320.0, 0.0, 240.0))
    plan = [(TUM1, 601, 480, 640), (TUM1, 602, 480, 640), (TUM1, 603, 480, 640), (TUM1, 604, 480, 640),     # eager, capture, replay x2
            (camB, 601, 480, 640), (camB, 605, 480, 640),                                                     # other camera: re-capture, replay
            (small, 606, 240, 320), (small, 607, 240, 320), (small, 608, 240, 320),                          # other size: new workspace
            (TUM1, 609, 480, 640), (TUM1, 610, 480, 640), (TUM1, 611, 480, 640)]                             # and back
    for cam, seed, H, W in plan:
        img = synth.make_frame(seed, H, W)
        depth = _depth(seed, H, W)
        fr = vo.Frame(ex, _cam(vo, cam), img, depth)
        rk, rd = P.extract(img)
        assert fr.n == len(rk) and fr.kps.tobytes() == rk.tobytes() and np.array_equal(fr.desc, rd), (seed, H)
        wun, wur, wdp, wstart, wids = P.frame_finish(rk, cam, depth)
        assert fr.unkps.tobytes() == wun.tobytes() and fr.uright.tobytes() == wur.tobytes() and fr.depth.tobytes() == wdp.tobytes(), seed
        start, ids = fr.grid()
        assert np.array_equal(start, wstart) and np.array_equal(ids, wids[:wstart[-1]])
        fr.close()
    ex.close()
