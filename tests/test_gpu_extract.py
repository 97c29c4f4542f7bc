"""GPU parity, extractor: the CUDA path (through the C ABI) against the CPU oracle, bit-exact, stage by stage and
end to end, plus the committed reference fixtures.  Run with -m gpu on the B200 box."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from vo_slam_test_b200 import synth


@pytest.fixture(scope="module")
def vo():
    import vo_slam_test_b200 as v
    assert v.device_count() > 0, "no CUDA device"
    return v


def _stage_check(ex, P, img, nfeat):
    """pyramid, blur, FAST candidates, quadtree selection of frame 0 against the oracle port."""
    H, W = img.shape
    _, _, levels = P.extract(img, want_levels=True)
    for l in range(P.nlevels):
        got = ex.debug_level(0, l)
        assert np.array_equal(got, levels[l]), "pyramid level %d" % l
        assert np.array_equal(ex.debug_level(0, l, blurred=True), P.blur(levels[l])), "blur level %d" % l
        cand = P.fast_cells(levels[l])
        gc = ex.debug_candidates(0, l)
        assert np.array_equal(gc, cand), "FAST candidates level %d (%d vs %d)" % (l, len(gc), len(cand))
        w, h = levels[l].shape[1], levels[l].shape[0]
        sel = P.octree(cand, w - 32, h - 32, int(nfeat[l]))
        gs = ex.debug_selected(0, l)
        assert np.array_equal(gs, cand[sel]), "quadtree level %d" % l


@pytest.mark.parametrize("seed,H,W,nf", [(42, 480, 640, 1000), (7, 240, 320, 300), (3, 1080, 1920, 2000),
                                         (5, 479, 641, 500), (11, 2160, 3840, 5000)])
def test_extract_matches_oracle(vo, seed, H, W, nf):
    img = synth.make_frame(seed, H, W)
    P = oracle.Port(nf)
    ex = vo.ORBextractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    assert np.array_equal(ex.features_per_level(), P.tables()[2])
    assert np.array_equal(ex.GetScaleFactors().view(np.uint32), P.tables()[0].view(np.uint32))
    _stage_check(ex, P, img, P.tables()[2])
    rk, rd = P.extract(img)
    assert len(kps) == len(rk)
    for name in KP_FIELDS:
        assert np.array_equal(kps[name].view(np.uint32), rk[name].view(np.uint32)), name
    assert np.array_equal(desc, rd)
    ex.close()


KP_FIELDS = ["x", "y", "size", "angle", "response", "octave", "class_id"]


@pytest.mark.parametrize("ini,mn", [(130, 128), (140, 20), (20, 20), (200, 150), (40, 3)])
def test_unusual_fast_thresholds(vo, ini, mn):
    """Thresholds >= 128 take the unfiltered path of the FAST kernel (the packed-byte rejection test covers t < 128 only);
    minTh >= iniTh disables the retry.  Same bit-exact contract."""
    img = synth.make_frame(21, 240, 320)
    img[60:120, 40:200] = np.where((np.indices((60, 160)).sum(0) // 9) % 2 == 0, 0, 255).astype(np.uint8)   # contrast > 128
    P = oracle.Port(300, 1.2, 8, ini, mn)
    ex = vo.ORBextractor(300, 1.2, 8, ini, mn)
    kps, desc = ex(img)
    rk, rd = P.extract(img)
    assert len(kps) == len(rk) and len(rk) > 0
    for name in KP_FIELDS:
        assert np.array_equal(kps[name].view(np.uint32), rk[name].view(np.uint32)), name
    assert np.array_equal(desc, rd)
    ex.close()


def test_extract_matches_reference_fixtures(vo):
    """Committed outputs of the reference's own ORBextractor.cpp (tests/golden/orb_*.npz)."""
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "orb_*.npz"))):
        g = np.load(path)
        img = synth.make_frame(int(g["seed"]), int(g["H"]), int(g["W"]))
        ex = vo.ORBextractor(int(g["nfeatures"]))
        kps, desc = ex(img)
        assert np.array_equal(kps, g["kps"]), path
        assert np.array_equal(desc, g["desc"]), path
        ex.close()


def test_batch_equals_single_and_oracle(vo):
    """Batch path (chunked, many frames per launch) == per-frame path == oracle; strided input rows."""
    B = 70   # > one chunk of 64
    imgs = synth.make_sequence(B, seed=3)
    ex = vo.ORBextractor()
    kps, desc, cnt = ex.extract_batch(imgs)
    P = oracle.Port()
    for f in [0, 1, 33, 63, 64, 69]:
        rk, rd = P.extract(imgs[f])
        assert cnt[f] == len(rk)
        assert np.array_equal(kps[f, :cnt[f]], rk) and np.array_equal(desc[f, :cnt[f]], rd)
    # non-contiguous rows (stride != width) through the single-frame entry point
    big = np.zeros((480, 700), np.uint8)
    big[:, 30:670] = imgs[5]
    k2, d2 = ex(big[:, 30:670])
    rk, rd = P.extract(imgs[5])
    assert np.array_equal(k2, rk) and np.array_equal(d2, rd)
    ex.close()


def test_edge_cases(vo):
    ex = vo.ORBextractor()
    k, d = ex(np.full((480, 640), 100, np.uint8))          # flat image: zero keypoints, descriptors released (:1073-1074)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.zeros((0, 0), np.uint8))                  # empty image: silent no-op (:1054)
    assert len(k) == 0
    low = (synth.make_frame(40).astype(np.int32) // 8 + 100).astype(np.uint8)   # minThFAST retry cells
    rk, rd = oracle.Port().extract(low)
    k, d = ex(low)
    assert len(rk) > 50 and np.array_equal(k, rk) and np.array_equal(d, rd)
    with pytest.raises(vo.OrbError):
        ex(np.zeros((40, 40), np.uint8))                   # too small for an 8-level pyramid
    with pytest.raises(vo.OrbError):
        ex(np.zeros((600, 250), np.uint8))                 # aspect ratio with zero quadtree roots (reference divides by zero)
    ex.close()


@pytest.mark.parametrize("kind", ["noise", "checker3", "checker2", "saltpepper", "stripes", "blocks_equal_scores"])
@pytest.mark.parametrize("H,W,nf", [(240, 320, 500), (480, 640, 1000)])
def test_adversarial_images(vo, kind, H, W, nf):
    """Corner-dense inputs: every queue of the FAST kernel near its worst case, thousands of equal scores for the NMS and the
    quadtree's strongest-key ties, cells with no corner at all next to saturated ones."""
    rng = np.random.default_rng(H + len(kind))
    yy, xx = np.indices((H, W))
    img = {
        "noise": lambda: rng.integers(0, 256, (H, W), dtype=np.uint8),
        "checker3": lambda: (((yy // 3 + xx // 3) % 2) * 255).astype(np.uint8),
        "checker2": lambda: (((yy // 2 + xx // 2) % 2) * 200 + 20).astype(np.uint8),
        "saltpepper": lambda: np.where(rng.random((H, W)) < 0.1, 255, 0).astype(np.uint8),
        "stripes": lambda: ((xx % 4 < 2) * 255).astype(np.uint8),
        "blocks_equal_scores": lambda: (((yy // 9 + xx // 11) % 2) * 90 + 60).astype(np.uint8),
    }[kind]()
    P = oracle.Port(nf)
    ex = vo.ORBextractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    _stage_check(ex, P, img, P.tables()[2])
    rk, rd = P.extract(img)
    assert len(kps) == len(rk)
    for name in KP_FIELDS:
        assert np.array_equal(kps[name].view(np.uint32), rk[name].view(np.uint32)), name
    assert np.array_equal(desc, rd)
    ex.close()


def test_other_parameters(vo):
    img = synth.make_frame(77, 600, 800)
    for (nf, sf, nl, ini, mn) in [(500, 1.2, 4, 20, 7), (1500, 1.1, 8, 30, 10), (800, 1.5, 5, 12, 5)]:
        P = oracle.Port(nf, sf, nl, ini, mn)
        ex = vo.ORBextractor(nf, sf, nl, ini, mn)
        k, d = ex(img)
        rk, rd = P.extract(img)
        assert np.array_equal(k, rk) and np.array_equal(d, rd), (nf, sf, nl)
        ex.close()


def test_error_convention(vo):
    """The C ABI never throws: bad arguments, too small buffers and unsupported shapes come back as status codes."""
    import ctypes as C
    L = vo.lib()
    ex = vo.ORBextractor()
    img = synth.make_frame(3)
    kps = np.zeros(10, vo.KP_DTYPE); desc = np.zeros((10, 32), np.uint8); n = C.c_int(-1)
    rc = L.orbx_extract(ex._h, C.c_void_p(img.ctypes.data), 640, 480, 640, C.c_void_p(kps.ctypes.data), C.c_void_p(desc.ctypes.data), 10,
                        C.byref(n))
    assert rc == -3 and n.value > 900                      # ORBX_ERR_CAPACITY, the needed count is still reported
    rk, rd = oracle.Port().extract(img)
    assert np.array_equal(kps, rk[:10]) and np.array_equal(desc, rd[:10])     # the part that fits is correct
    assert L.orbx_extract(None, C.c_void_p(img.ctypes.data), 640, 480, 640, None, None, 10, C.byref(n)) == -1
    assert L.orbx_extract(ex._h, C.c_void_p(img.ctypes.data), 640, 480, 100, C.c_void_p(kps.ctypes.data), C.c_void_p(desc.ctypes.data), 10,
                          C.byref(n)) == -1               # stride < width
    wide = np.zeros((300, 4300), np.uint8)
    with pytest.raises(vo.OrbError, match="too large|aspect"):
        ex(wide)
    assert b"" != L.orbx_last_error()
    p = vo.api._Params(1000, 0.9, 8, 20, 7, 0); h = C.c_void_p()
    assert L.orbx_create(C.byref(p), C.byref(h)) == -1     # scale factor <= 1
    p = vo.api._Params(1000, 1.2, 8, 20, 7, 99)
    assert L.orbx_create(C.byref(p), C.byref(h)) == -2     # no such device
    ex.close()


def test_random_configurations(vo):
    """The generator of tests/test_oracle_vs_reference.py::test_port_equals_compiled_reference_random_configurations on the
    GPU: random frame sizes, feature budgets, scale factors, level counts and thresholds.  Shapes the library declines
     must come back as an error, never as different bits."""
    rng = np.random.default_rng(20261017)
    done = declined = 0
    while done < 24:
        H = int(rng.integers(140, 700)); W = int(rng.integers(H, 1000))
        nf = int(rng.choice([50, 200, 500, 1000, 2000])); sc = float(rng.choice([1.2, 1.1, 1.5, 2.0])); nl = int(rng.integers(1, 9))
        ini = int(rng.choice([20, 12, 40, 7])); mn = min(int(rng.choice([7, 3, ini])), ini)
        if min(W, H) / (sc ** (nl - 1)) < 80:
            continue
        img = synth.make_frame(int(rng.integers(0, 10 ** 6)), H, W)
        if done % 5 == 0:
            img = (img.astype(np.int32) // 6 + 90).astype(np.uint8)
        done += 1
        rk, rd = oracle.Port(nf, sc, nl, ini, mn).extract(img)
        ex = vo.ORBextractor(nf, sc, nl, ini, mn)
        try:
            k, d = ex(img)
        except vo.OrbError:
            declined += 1
            continue
        finally:
            ex.close()
        assert np.array_equal(k, rk) and np.array_equal(d, rd), (H, W, nf, sc, nl, ini, mn)
    assert declined <= 4
