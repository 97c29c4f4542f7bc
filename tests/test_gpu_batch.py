"""GPU parity of the batch / device / pipelined entry points and size-independent properties at BASELINE scale."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from vo_slam_test_b200 import synth


@pytest.fixture(scope="module")
def vo():
    import vo_slam_test_b200 as v
    assert v.device_count() > 0
    return v


def test_extract_match_batch_equals_oracle(vo):
    """orbx_extract_match_batch (host ABI, three-stream pipeline, chunked) == oracle extractor + oracle top-2."""
    import os
    os.environ["ORBX_CHUNK"] = "5"          # force several chunks and pairs that straddle chunk boundaries
    try:
        B = 13
        imgs = synth.make_sequence(B, seed=11)
        ex = vo.ORBextractor()
        cap = ex.max_keypoints
        kps = np.zeros((B, cap), vo.KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); cnt = np.zeros(B, np.int32)
        midx = np.full((B - 1, cap), -7, np.int32); md1 = np.zeros((B - 1, cap), np.int32); md2 = np.zeros((B - 1, cap), np.int32)
        mok = np.zeros((B - 1, cap), np.uint8)
        ex.extract_match_batch(imgs.ctypes.data, B, 640, 480, kps.ctypes.data, desc.ctypes.data, cap, cnt.ctypes.data, 50, 0.7,
                               midx.ctypes.data, md1.ctypes.data, md2.ctypes.data, mok.ctypes.data)
        P = oracle.Port()
        ref = [P.extract(imgs[f]) for f in range(B)]
        for f in range(B):
            assert cnt[f] == len(ref[f][0])
            assert np.array_equal(kps[f, :cnt[f]], ref[f][0]) and np.array_equal(desc[f, :cnt[f]], ref[f][1])
        for p in range(B - 1):
            want = P.knn2(ref[p][1], ref[p + 1][1], 50, 0.7)
            n = cnt[p]
            assert np.array_equal(midx[p, :n], want[0]) and np.array_equal(md1[p, :n], want[1])
            assert np.array_equal(md2[p, :n], want[2]) and np.array_equal(mok[p, :n], want[3])
            assert want[3].sum() > 100          # consecutive frames of the panning sequence really match
        ex.close()
    finally:
        del os.environ["ORBX_CHUNK"]


def test_device_api_with_unaligned_strides(vo):
    """Resident frames whose row stride is not a multiple of 4/16: the byte-granular fallback paths of the pyramid, FAST,
    blur and patch staging must give the same bits as the aligned paths."""
    torch = pytest.importorskip("torch")
    B, H, W = 3, 480, 640
    imgs = synth.make_sequence(B, seed=5)
    P = oracle.Port()
    ex = vo.ORBextractor()
    cap = ex.max_keypoints
    for pad, off in [(0, 0), (3, 1), (7, 2), (16, 0)]:
        big = torch.zeros((B, H, W + pad), dtype=torch.uint8, device="cuda")
        big[:, :, off:off + W] = torch.from_numpy(imgs).cuda() if pad else torch.from_numpy(imgs).cuda()[:, :, :W]
        d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        ex.extract_batch_device(big.data_ptr() + off, B, W, H, W + pad, (W + pad) * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                d_cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        cnt = d_cnt.cpu().numpy(); kp = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28); de = d_desc.cpu().numpy()
        for f in range(B):
            rk, rd = P.extract(imgs[f])
            assert cnt[f] == len(rk), (pad, off)
            assert np.array_equal(kp[f, :cnt[f]].reshape(-1).view(vo.KP_DTYPE), rk), (pad, off)
            assert np.array_equal(de[f, :cnt[f]], rd), (pad, off)
    ex.close()


def test_properties_at_baseline_scale(vo):
    """Size-independent properties on a 1024-frame resident batch (too big for the CPU oracle in a test):
    determinism across runs and chunkings, duplicate frames give identical output, self-matching is the identity,
    keypoint counts respect the quadtree's N .. N+3 per level bound, spot frames equal the oracle."""
    torch = pytest.importorskip("torch")
    import os
    B = 1024
    imgs = synth.make_sequence(B, seed=21)
    imgs[700] = imgs[100]                                  # planted duplicate
    ex = vo.ORBextractor()
    cap = ex.max_keypoints
    d_imgs = torch.from_numpy(imgs).cuda()
    outs = []
    for chunk in ("256", "37"):
        os.environ["ORBX_CHUNK"] = chunk
        ex2 = vo.ORBextractor()
        d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        ex2.extract_batch_device(d_imgs.data_ptr(), B, 640, 480, 640, 640 * 480, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                 d_cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append((d_kps.cpu().numpy(), d_desc.cpu().numpy(), d_cnt.cpu().numpy()))
        ex2.close()
    del os.environ["ORBX_CHUNK"]
    (k0, d0, c0), (k1, d1, c1) = outs
    assert np.array_equal(c0, c1)
    nf = ex.features_per_level()
    assert c0.min() > 900 and c0.max() <= int(nf.sum()) + 3 * len(nf)
    for f in range(B):
        n = c0[f]
        assert np.array_equal(k0[f, :n].view(np.uint32), k1[f, :n].view(np.uint32)) and np.array_equal(d0[f, :n], d1[f, :n])
    n = c0[100]
    assert c0[700] == n and np.array_equal(d0[700, :n], d0[100, :n]) and np.array_equal(k0[700, :n].view(np.uint32), k0[100, :n].view(np.uint32))
    P = oracle.Port()
    for f in (0, 511, 1023):
        rk, rd = P.extract(imgs[f])
        assert c0[f] == len(rk) and np.array_equal(d0[f, :c0[f]], rd)
    # self-matching: every descriptor's nearest neighbour in its own frame is the first identical row (itself unless duplicated)
    M = vo.Matcher(0.7)
    idx, dd1, dd2, ok = M.knn2(d0[5, :c0[5]], d0[5, :c0[5]])
    assert (dd1 == 0).all()
    first = np.array([np.flatnonzero((d0[5, :c0[5]] == row).all(1))[0] for row in d0[5, :c0[5]]])
    assert np.array_equal(idx, first)
    ex.close()


def test_two_lane_resident_batch_equals_oracle(vo):
    """orbx_extract_batch_device splits a batch of >= 2 chunks over two lanes (this handle on the caller's stream, a sibling
    handle with its own workspace on a second stream, joined by an event).  Both lane counts must give the oracle's bits,
    for an odd number of chunks and a ragged last chunk, and later work on the caller's stream must see every frame."""
    import os
    torch = pytest.importorskip("torch")
    B, H, W = 19, 240, 320
    imgs = np.stack([synth.make_frame(100 + f, H, W) for f in range(B)])
    P = oracle.Port(300)
    ref = [P.extract(imgs[f]) for f in range(B)]
    d_imgs = torch.from_numpy(imgs).cuda()
    try:
        os.environ["ORBX_CHUNK"] = "4"       # 5 chunks: 3 on the first lane, 2 (the last one ragged) on the second
        for lanes in ("2", "1", "2"):
            os.environ["ORBX_LANES"] = lanes
            ex = vo.ORBextractor(300, 1.2, 8, 20, 7)
            cap = ex.max_keypoints
            for rep in range(2):             # second pass reuses both workspaces and the level-0 tensor maps
                d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
                d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
                d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
                st = torch.cuda.current_stream()
                ex.extract_batch_device(d_imgs.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                        d_cnt.data_ptr(), st.cuda_stream)
                total = d_cnt.sum()          # ordered after the join on the caller's stream, no explicit synchronisation
                cnt = d_cnt.cpu().numpy()
                assert int(total.item()) == sum(len(r[0]) for r in ref)
                kps = d_kps.cpu().numpy().view(vo.KP_DTYPE).reshape(B, cap)
                desc = d_desc.cpu().numpy()
                for f in range(B):
                    assert cnt[f] == len(ref[f][0]), (lanes, f)
                    assert np.array_equal(kps[f, :cnt[f]], ref[f][0]) and np.array_equal(desc[f, :cnt[f]], ref[f][1]), (lanes, f)
            # 7 resize + FAST + quadtree + blur + orient/desc per chunk, both lanes counted; chunks of <= 8 frames run in latency
            # mode, where level 0 has its own FAST and quadtree launches on a second stream (+2)
            assert ex.launch_count() == 2 * 5 * 13
            ex.close()
    finally:
        os.environ.pop("ORBX_CHUNK", None)
        os.environ.pop("ORBX_LANES", None)


def test_batch_mode_equals_latency_mode_and_oracle(vo):
    """One chunk of 12 frames (throughput mode: one stream, quadtree keys in global memory) against the same frames one call
    at a time (latency mode: <= 8 frames per call run level 0 and the blur on a second stream and keep the quadtree's keys in
    shared memory) and against the oracle."""
    import os
    B = 12
    imgs = synth.make_sequence(B, seed=23)
    P = oracle.Port()
    os.environ["ORBX_CHUNK"] = "12"
    try:
        ex = vo.ORBextractor()
        kk, dd, cc = ex.extract_batch(imgs)
        launches_batch = ex.launch_count()
        for f in range(B):
            k1, d1 = ex(imgs[f])
            rk, rd = P.extract(imgs[f])
            assert cc[f] == len(rk) == len(k1)
            assert np.array_equal(kk[f, :cc[f]], rk) and np.array_equal(dd[f, :cc[f]], rd)
            assert np.array_equal(k1, rk) and np.array_equal(d1, rd)
        assert launches_batch == 11                      # 7 resize + FAST + quadtree + blur + orient/desc: one chunk, one stream
        assert ex.launch_count() == 11 + B * 13          # latency mode: level 0 has its own FAST and quadtree launches
        ex.close()
    finally:
        del os.environ["ORBX_CHUNK"]


def test_resident_extract_match_equals_two_step_path_and_oracle(vo):
    """orbx_extract_match_batch_device (matching of every chunk launched on its lane right behind the extraction; the pairs that
    straddle a lane boundary after the join) against orbx_extract_batch_device + hamm_knn2_pairs_device, and against the oracle
    on a few pairs: several chunks per lane, a ragged last chunk, two and one lanes."""
    import os
    torch = pytest.importorskip("torch")
    from vo_slam_test_b200 import api
    B, H, W = 23, 240, 320
    imgs = np.stack([synth.make_frame(300 + f // 2, H, W) if f % 5 else synth.make_frame(300 + f, H, W) for f in range(B)])
    d_imgs = torch.from_numpy(imgs).cuda()
    P = oracle.Port(300)
    try:
        os.environ["ORBX_CHUNK"] = "4"
        for lanes in ("2", "1"):
            os.environ["ORBX_LANES"] = lanes
            ex = vo.ORBextractor(300, 1.2, 8, 20, 7)
            cap = ex.max_keypoints
            outs = []
            for fused in (True, False):
                d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
                d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
                d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
                m = [torch.full((B - 1, cap), -9, dtype=torch.int32, device="cuda") for _ in range(3)]
                ok = torch.full((B - 1, cap), 7, dtype=torch.uint8, device="cuda")
                st = torch.cuda.current_stream().cuda_stream
                if fused:
                    ex.extract_match_batch_device(d_imgs.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                                  d_cnt.data_ptr(), 50, 0.7, m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(),
                                                  ok.data_ptr(), st)
                else:
                    ex.extract_batch_device(d_imgs.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                            d_cnt.data_ptr(), st)
                    qf = torch.arange(0, B - 1, dtype=torch.int32, device="cuda"); tf = qf + 1
                    api.knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), cap, qf.data_ptr(), tf.data_ptr(), B - 1, 50, 0.7,
                                          m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(), ok.data_ptr(), st)
                torch.cuda.synchronize()
                cnt = d_cnt.cpu().numpy()
                outs.append((cnt, d_desc.cpu().numpy(), [x.cpu().numpy() for x in m], ok.cpu().numpy()))
            (c0, de0, m0, ok0), (c1, de1, m1, ok1) = outs
            assert np.array_equal(c0, c1) and np.array_equal(de0, de1)
            for p in range(B - 1):
                n = c0[p]
                for a, b in zip(m0, m1):
                    assert np.array_equal(a[p, :n], b[p, :n]), (lanes, p)
                assert np.array_equal(ok0[p, :n], ok1[p, :n]), (lanes, p)
            for p in (0, 3, 11, 12, 21):       # chunk boundaries (3|4, 11|12 = the lane boundary with two lanes) and interior pairs
                want = oracle.Port().knn2(de0[p, :c0[p]], de0[p + 1, :c0[p + 1]], 50, 0.7)
                n = c0[p]
                assert np.array_equal(m0[0][p, :n], want[0]) and np.array_equal(m0[1][p, :n], want[1])
                assert np.array_equal(m0[2][p, :n], want[2]) and np.array_equal(ok0[p, :n], want[3])
            rk, rd = P.extract(imgs[12])
            assert c0[12] == len(rk) and np.array_equal(de0[12, :c0[12]], rd)
            ex.close()
    finally:
        os.environ.pop("ORBX_CHUNK", None)
        os.environ.pop("ORBX_LANES", None)
