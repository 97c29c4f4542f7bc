"""GPU parity of MapPoint::computeDescriptor (mappoint.cpp:118-179) against the oracle port."""
import numpy as np
import pytest

import oracle
from vo_slam_test_b200 import synth


def _case(seed, P=2000, maxobs=40):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, maxobs, P)
    counts[:5] = [0, 1, 2, 3, 64]
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    base = synth.make_descriptors(P, seed)
    rows = []
    for p in range(P):
        n = counts[p]
        if n:
            d = synth.flip_bits(np.repeat(base[p:p + 1], n, 0), rng.integers(0, 50, n), rng)
            if n > 3 and p % 7 == 0:
                d[2] = d[0]                       # exact duplicates: median ties, first row must win
            rows.append(d)
    return np.concatenate(rows), start


def test_port_medoid_small_numpy():
    desc, start = _case(1, P=50, maxobs=12)
    best = oracle.Port().medoid(desc, start)
    for p in range(50):
        D = desc[start[p]:start[p + 1]]
        if len(D) == 0:
            assert best[p] == -1
            continue
        dist = np.unpackbits(D[:, None, :] ^ D[None, :, :], axis=2).sum(2)
        mids = np.sort(dist, 1)[:, int(0.5 * (len(D) - 1))]
        assert best[p] == int(np.argmin(mids))    # argmin returns the first minimum, like the strict '<' loop


@pytest.mark.gpu
def test_medoid_gpu():
    import vo_slam_test_b200 as vo
    desc, start = _case(2)
    want = oracle.Port().medoid(desc, start)
    got = vo.medoid_descriptors(desc, start)
    assert np.array_equal(got, want)
    assert len(vo.medoid_descriptors(np.zeros((0, 32), np.uint8), np.zeros(1, np.int32))) == 0
