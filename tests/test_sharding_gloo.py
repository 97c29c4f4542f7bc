"""Multi-GPU host logic without GPUs: frame/pair/train-shard partitioning and the top-2 all-gather + merge rule,
run with torch.distributed gloo, world_size 2 (and 3), on CPU.  The local top-2 of each shard comes from the oracle
(standing in for the CUDA kernel, which the -m gpu tests check separately); the merge under test is the rule the
GPU merge kernel implements (smallest d1, lowest shard wins ties, second = 2nd smallest of {d1_s, d2_s})."""
import os
import socket

import numpy as np
import pytest

import oracle
from vo_slam_test_b200 import sharded, synth


def test_frame_and_pair_blocks_cover_everything():
    for n in (1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharded.frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            pairs = []
            for r in range(world):
                lo, hi, halo = sharded.pair_block(n, r, world)
                pairs += list(range(lo, hi))
                if halo is not None:
                    assert halo == blocks[r][1] and halo < n       # the replicated boundary frame
            assert pairs == list(range(n - 1))                       # every consecutive pair exactly once
            for r in range(world):                                   # the strong-scaling view of the same partition
                lo, hi, hi_ext, p_hi = sharded.strong_block(n, r, world)
                assert (lo, hi) == blocks[r] and hi_ext in (hi, hi + 1) and hi_ext <= n
                assert all(p + 1 < hi_ext for p in range(lo, p_hi))  # every owned pair has both frames on the rank


def test_train_shards():
    for M in (0, 1, 10, 16777216, 1000003):
        for world in (1, 2, 8):
            sh = [sharded.train_shard(M, r, world) for r in range(world)]
            assert sh[0][0] == 0 and sh[-1][1] == M and all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))


def test_merge_rule_equals_sequential_scan():
    rng = np.random.default_rng(0)
    P = oracle.Port()
    t = synth.make_descriptors(6000, seed=5)
    q = synth.flip_bits(t[rng.integers(0, 6000, 300)], rng.integers(0, 70, 300), rng)
    t[100:110] = t[4000:4010]            # exact duplicates across shards: index ties
    want = P.knn2(q, t, 50, 0.7)
    for world in (2, 3, 8):
        idx, d1, d2 = [], [], []
        for r in range(world):
            lo, hi = sharded.train_shard(len(t), r, world)
            i, a, b, _ = P.knn2(q, t[lo:hi], 50, 0.7)
            idx.append(np.where(i >= 0, i + lo, i)); d1.append(a); d2.append(b)
        mi, m1, m2 = sharded.merge_top2_numpy(np.stack(idx), np.stack(d1), np.stack(d2))
        assert np.array_equal(mi, want[0]) and np.array_equal(m1, want[1]) and np.array_equal(m2, want[2])


def test_merge_rule_with_empty_shards():
    """M = 9 rows over 8 ranks: train_shard gives trailing ranks an EMPTY shard; its record is the sentinel (-1, 256, 256)
    and the merge must ignore it."""
    rng = np.random.default_rng(4)
    P = oracle.Port()
    t = synth.make_descriptors(9, seed=6)
    q = synth.flip_bits(t[rng.integers(0, 9, 40)], rng.integers(0, 40, 40), rng)
    want = P.knn2(q, t, 50, 0.7)
    idx, d1, d2 = [], [], []
    empties = 0
    for r in range(8):
        lo, hi = sharded.train_shard(9, r, 8)
        empties += hi == lo
        i, a, b, _ = P.knn2(q, t[lo:hi], 50, 0.7)
        idx.append(np.where(i >= 0, i + lo, i)); d1.append(a); d2.append(b)
    assert empties >= 3
    mi, m1, m2 = sharded.merge_top2_numpy(np.stack(idx), np.stack(d1), np.stack(d2))
    assert np.array_equal(mi, want[0]) and np.array_equal(m1, want[1]) and np.array_equal(m2, want[2])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, t, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharded.train_shard(len(t), rank, world)
        i, a, b, _ = oracle.Port().knn2(q, t[lo:hi], 50, 0.7)
        rec = torch.from_numpy(np.stack([np.where(i >= 0, i + lo, i), a, b], 1).astype(np.int32))
        allrec = sharded.allgather_records(rec, dist).numpy()          # [world, Q, 3]
        mi, m1, m2 = sharded.merge_top2_numpy(allrec[:, :, 0], allrec[:, :, 1], allrec[:, :, 2])
        # frames: every rank extracts only its block; gather the keypoint counts to rank 0
        flo, fhi = sharded.frame_block(8, rank, world)
        cnt = torch.zeros(8, dtype=torch.int32)
        for f in range(flo, fhi):
            cnt[f] = len(oracle.Port(300).extract(synth.make_frame(f, 240, 320))[0])
        dist.all_reduce(cnt)
        out.put((rank, mi, m1, m2, cnt.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_knn_and_frame_blocks_gloo(world):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(1)
    t = synth.make_descriptors(3001, seed=9)
    q = synth.flip_bits(t[rng.integers(0, 3001, 64)], rng.integers(0, 70, 64), rng)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, t, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.Port().knn2(q, t, 50, 0.7)
    wantcnt = np.array([len(oracle.Port(300).extract(synth.make_frame(f, 240, 320))[0]) for f in range(8)], np.int32)
    for rank, mi, m1, m2, cnt in res:
        assert np.array_equal(mi, want[0]) and np.array_equal(m1, want[1]) and np.array_equal(m2, want[2])
        assert np.array_equal(cnt, wantcnt)       # every frame extracted exactly once across ranks


def _peer_setup_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            sharded.PeerExchange(dist, 0, 16)
            out.put((rank, "constructed"))
        except RuntimeError as e:
            out.put((rank, "raised: %s" % e))
        t = __import__("torch").ones(1)
        dist.all_reduce(t)                    # the ranks are still in step after the failed set-up
        out.put((rank, "in step %d" % int(t.item())))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_setup_fails_on_every_rank_together():
    """Without a CUDA device the exchange buffers cannot be allocated.  The set-up must then fail on EVERY rank (no rank may
    be left waiting in a collective) and leave the process group usable -- the property bench.py's fallback to the NCCL
    all-gather relies on."""
    import torch.multiprocessing as mp
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA: with a GPU the set-up succeeds (covered by the -m gpu tests / bench --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_setup_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        msgs = [m for r, m in res if r == rank]
        assert any(m.startswith("raised: peer exchange set-up failed") for m in msgs), msgs
        assert "in step 2" in msgs
