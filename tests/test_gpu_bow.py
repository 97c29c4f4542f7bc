"""GPU parity of the BoW-guided matcher (matcher.cpp:449-559 and :561-677) against the CPU oracle port, plus an
independent pure-Python restatement of the reference loop on a small case (the reference has no compilable matcher TU)."""
import numpy as np
import pytest

import oracle
from vo_slam_test_b200 import synth


def _sides(seed, n=1000, m=1000, flips=40, node_bits=9, invalid=0.1):
    rng = np.random.default_rng(seed)
    db = synth.make_descriptors(m, seed=seed)
    src = rng.integers(0, m, n)
    da = synth.flip_bits(db[src], rng.integers(0, flips, n), rng)
    da[:, 0] = db[src][:, 0]; da[:, 1] = (da[:, 1] & 0x7F) | (db[src][:, 1] & 0x80)   # keep the node bits of most pairs
    ang_b = rng.uniform(0, 360, m).astype(np.float32)
    ang_a = ((ang_b[src] + rng.normal(0, 8, n)) % 360).astype(np.float32)
    ang_a[rng.random(n) < 0.15] = rng.uniform(0, 360, int((rng.random(n) < 0.15).sum()) or 1)[0]
    A = synth.make_bow_side(da, ang_a, (rng.random(n) > invalid).astype(np.uint8), node_bits, seed)
    B = synth.make_bow_side(db, ang_b, None, node_bits, seed + 1)
    return A, B


def _python_bow(a, b, mode, ratio, th_low, check_rot):
    """Line-by-line Python restatement of the two reference loops (small inputs only)."""
    def ham(x, y):
        return int(np.unpackbits(x ^ y).sum())
    n_out = len(b["desc"]) if mode == 0 else len(a["desc"])
    match = np.full(n_out, -1, np.int32); taken = np.zeros(len(b["desc"]), bool)
    hist = [[] for _ in range(30)]; cnt = 0
    bnodes = {int(v): g for g, v in enumerate(b["node_ids"])}
    for ga, node in enumerate(a["node_ids"]):
        gb = bnodes.get(int(node))
        if gb is None:
            continue
        for i1 in a["feat_idx"][a["group_start"][ga]:a["group_start"][ga + 1]]:
            if not a["valid"][i1]:
                continue
            b1 = b2 = 256; bi = -1
            for i2 in b["feat_idx"][b["group_start"][gb]:b["group_start"][gb + 1]]:
                if taken[i2] or not b["valid"][i2]:
                    continue
                d = ham(a["desc"][i1], b["desc"][i2])
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if b1 <= th_low and np.float32(b1) < np.float32(ratio) * np.float32(b2):
                taken[bi] = True
                out = bi if mode == 0 else i1
                match[out] = i1 if mode == 0 else bi
                if check_rot:
                    rot = np.float32(a["angle"][i1]) - np.float32(b["angle"][bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    v = np.float32(rot * np.float32(np.float32(30) / np.float32(360.0)))
                    bn = int(np.rint(v)) if mode == 0 else int(np.floor(v + np.float32(0.5)))
                    if bn == 30:
                        bn = 0
                    hist[bn].append(out)
                cnt += 1
    if check_rot:
        sizes = [len(h) for h in hist]
        m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, i3, m2, i2, m1, i1 = m2, i2, m1, i1, s, i
            elif s > m2:
                m3, i3, m2, i2 = m2, i2, s, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        for k in range(30):
            if k not in (i1, i2, i3):
                for o in hist[k]:
                    match[o] = -2; cnt -= 1
    return match, cnt


@pytest.mark.parametrize("mode", [0, 1])
def test_port_matches_python_restatement(mode):
    A, B = _sides(3, n=300, m=300, node_bits=6)
    if mode == 1:
        B["valid"] = (np.random.default_rng(1).random(300) > 0.1).astype(np.uint8)
    want = _python_bow(A, B, mode, 0.75, 50, True)
    got = oracle.Port().search_by_bow(A, B, mode, 0.75, 50, True)
    assert got[1] == want[1] and np.array_equal(got[0], want[0])
    assert want[1] > 20


@pytest.mark.gpu
@pytest.mark.parametrize("mode,rot,bits,n", [(0, True, 9, 1000), (1, True, 9, 1000), (0, False, 9, 1000), (0, True, 4, 1200),
                                             (1, True, 2, 800), (0, True, 12, 50), (1, False, 9, 3)])
def test_search_by_bow_gpu(mode, rot, bits, n):
    import vo_slam_test_b200 as vo
    A, B = _sides(mode * 10 + bits, n=n, m=n, node_bits=bits)
    if mode == 1:
        B["valid"] = (np.random.default_rng(2).random(n) > 0.1).astype(np.uint8)
    ratio = 0.7 if mode == 0 else 0.75
    want = oracle.Port().search_by_bow(A, B, mode, ratio, 50, rot)
    got = vo.Matcher(ratio).searchByBoW(A, B, mode=mode, checkRot=rot)
    assert got[1] == want[1]
    assert np.array_equal(got[0], want[0])
    if n >= 800:
        assert want[1] > 50
        if rot:
            assert (want[0] == -2).sum() > 0


@pytest.mark.gpu
def test_search_by_bow_empty():
    import vo_slam_test_b200 as vo
    A, B = _sides(5, n=20, m=20)
    E = synth.make_bow_side(np.zeros((0, 32), np.uint8), np.zeros(0, np.float32))
    m, c = vo.Matcher(0.7).searchByBoW(E, B, mode=0)
    assert c == 0 and (m == -1).all() and len(m) == 20
    m, c = vo.Matcher(0.7).searchByBoW(A, E, mode=1)
    assert c == 0 and (m == -1).all() and len(m) == 20


@pytest.mark.gpu
@pytest.mark.parametrize("rot,stereo", [(True, False), (False, True), (True, True)])
def test_search_for_triangulation_gpu(rot, stereo):
    """Matcher::searchForTriangulation (matcher.cpp:867-1010) on two frames of the panning sequence; the fundamental
    matrix of a pure x-translation makes the epipolar test select same-row pairs."""
    import vo_slam_test_b200 as vo
    P = oracle.Port()
    seq = synth.make_sequence(2, seed=31)
    k1, d1 = P.extract(seq[0]); k2, d2 = P.extract(seq[1])
    sf = P.tables()[0]
    rng = np.random.default_rng(8)
    def side(k, d, seed):
        s = synth.make_bow_side(d, k["angle"], (rng.random(len(k)) > 0.3).astype(np.uint8), 6, seed)   # valid = no map point yet
        s["kps"] = k
        s["uright"] = np.where(rng.random(len(k)) < 0.5, k["x"] - 10.0, -1.0).astype(np.float32) if stereo else np.full(len(k), -1.0, np.float32)
        return s
    A, B = side(k1, d1, 1), side(k2, d2, 2)
    F12 = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])          # p1^T F p2 = y2 - y1  (epipolar lines are image rows)
    want, wc = P.search_for_triangulation(A, B, F12, (600.0, 240.0), sf, 50, rot)
    got, gc = vo.Matcher(0.6).searchForTriangulation(A, B, F12, (600.0, 240.0), sf, checkRot=rot)
    assert gc == wc and np.array_equal(got, want)
    assert wc > 30
