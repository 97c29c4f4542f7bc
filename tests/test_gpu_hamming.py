"""GPU parity, Hamming top-2 (matcher.cpp:481-507,1240-1256) against the CPU oracle port, bit-exact."""
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


@pytest.mark.parametrize("nq,nt", [(1000, 1000), (1, 1), (130, 7), (257, 100000), (1000, 300001), (5, 0)])
def test_knn2_matches_oracle(vo, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    t = synth.make_descriptors(max(nt, 1), seed=nt)[:nt]
    q = synth.make_descriptors(nq, seed=nq + 1)
    if nt > 10:
        # plant near duplicates so the ratio test fires, and exact duplicates so index ties matter
        src = rng.integers(0, nt, nq)
        q = synth.flip_bits(t[src], rng.integers(0, 60, nq), rng)
        t[rng.integers(0, nt, 20)] = t[rng.integers(0, nt, 20)]
    P = oracle.Port()
    want = P.knn2(q, t, 50, 0.7, nthreads=8)
    got = vo.Matcher(0.7).knn2(q, t, th=50)
    for a, b, name in zip(got, want, ["idx", "d1", "d2", "ok"]):
        assert np.array_equal(a, b), name
    if nt > 10:
        assert got[3].sum() > 0


def test_tie_breaking_first_index_wins(vo):
    """All-equal train rows: every distance ties; best index must be 0 and d2 == d1 (duplicates count)."""
    q = synth.make_descriptors(64, seed=1)
    t = np.repeat(synth.make_descriptors(1, seed=2), 5000, axis=0)
    idx, d1, d2, ok = vo.Matcher(0.7).knn2(q, t)
    assert (idx == 0).all() and np.array_equal(d1, d2)
    want = oracle.Port().knn2(q, t)
    assert np.array_equal(idx, want[0]) and np.array_equal(d1, want[1])


def test_compute_distance(vo):
    rng = np.random.default_rng(3)
    P = oracle.Port()
    for _ in range(5):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert vo.Matcher.computeDistance(a, b) == P.hamming(a, b) == int(np.unpackbits(a ^ b).sum())
    z = np.zeros(32, np.uint8)
    assert vo.Matcher.computeDistance(z, z) == 0


def test_fused_peer_exchange_two_virtual_ranks(vo):
    """hamm_knn2_sharded_device with two 'ranks' living on one GPU in one process: both exchange buffers are plain device
    allocations, and the phase-split entry point lets the single stream enqueue the scan + scatter of BOTH ranks before
    either merge (in a real job every rank has its own GPU and the merge's flag wait overlaps the peers' scatters).
    Exercises the record layout, the flag protocol, the parity switch between calls and the merge rule without CUDA IPC.
    Result == one kNN over the whole train set == the CPU oracle."""
    import torch
    from vo_slam_test_b200 import api
    rng = np.random.default_rng(5)
    Q, M = 700, 9000
    q = rng.integers(0, 256, (Q, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (M, 32), dtype=np.uint8)
    t[rng.integers(0, M, 200)] = q[rng.integers(0, Q, 200)]           # exact duplicates: ties across the shard boundary
    want = oracle.Port().knn2(q, t, 50, 0.7)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    world, split = 2, 4000
    shards = [torch.from_numpy(t[:split]).to(dev), torch.from_numpy(t[split:]).to(dev)]
    los = [0, split]
    bufs = [api.exchange_alloc(0, world, 1024)[0] for _ in range(world)]
    st = torch.cuda.current_stream().cuda_stream
    try:
        for epoch in (1, 2, 3):                                        # three calls: both parities, buffer reuse
            outs = []
            for r in range(world):
                idx = torch.empty(Q, dtype=torch.int32, device=dev); d1 = torch.empty_like(idx); d2 = torch.empty_like(idx)
                ok = torch.empty(Q, dtype=torch.uint8, device=dev)
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                wsb = api.knn2_workspace_bytes(Q, shards[r].shape[0])
                ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
                outs.append((idx, d1, d2, ok, status, ws, wsb))
            for phases in (1, 2):
                for r in range(world):
                    idx, d1, d2, ok, status, ws, wsb = outs[r]
                    api.knn2_sharded_device(d_q.data_ptr(), Q, shards[r].data_ptr(), shards[r].shape[0], los[r], 50, 0.7, r, world,
                                            bufs, 1024, epoch, idx.data_ptr(), d1.data_ptr(), d2.data_ptr(), ok.data_ptr(),
                                            status.data_ptr(), ws.data_ptr(), wsb, st, phases=phases)
            torch.cuda.synchronize()
            for r in range(world):
                idx, d1, d2, ok, status = outs[r][:5]
                assert int(status.item()) == 0, "exchange timed out"
                got = (idx.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy(), ok.cpu().numpy())
                assert all(np.array_equal(a, b) for a, b in zip(got, want)), "rank %d epoch %d" % (r, epoch)
    finally:
        for b in bufs:
            api.exchange_free(b)


def test_dynamic_block_scan_equals_static_and_oracle(vo):
    """Map-scale scans hand the rows out in blocks from one counter per query tile (knn2_dyn_kernel): a CTA's blocks are not
    contiguous, so the merge breaks distance ties by row index.  Forced on a problem the oracle finishes in seconds, with exact
    duplicates of the queries planted far apart (ties at distance 0 across blocks and CTAs: the lowest row must win and the
    second-best must count the duplicate), all-equal rows, and a ragged last block."""
    from vo_slam_test_b200 import api
    rng = np.random.default_rng(77)
    nq, nt = 300, 200003
    t = synth.make_descriptors(nt, seed=5)
    src = rng.integers(0, nt, nq)
    q = synth.flip_bits(t[src], rng.integers(0, 40, nq), rng)
    for k in range(0, nq, 3):                       # every third query: exact copies at three far-apart rows
        rows = rng.choice(nt, 3, replace=False)
        t[rows] = q[k]
    t[150000:150600] = t[150000]                    # a run of identical rows spanning a block boundary
    want = oracle.Port().knn2(q, t, 50, 0.7, nthreads=8)
    prev = api.set_hamming_dynamic(2)
    try:
        got_dyn = vo.Matcher(0.7).knn2(q, t, th=50)
        api.set_hamming_dynamic(0)
        got_sta = vo.Matcher(0.7).knn2(q, t, th=50)
    finally:
        api.set_hamming_dynamic(prev)
    for a, b, c, name in zip(got_dyn, got_sta, want, ["idx", "d1", "d2", "ok"]):
        assert np.array_equal(a, c), "dynamic " + name
        assert np.array_equal(b, c), "static " + name
    assert (want[1][::3] == 0).all() and (want[2][::3] == 0).all()      # the planted ties are really there
