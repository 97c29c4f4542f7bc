"""Multi-rank parity on hardware (SURVEY section 4 layer 5): several processes, one rank each, results compared with the CPU
oracle -- not with another path of this repo.

* the map-scale sharded Hamming top-2 through BOTH exchange paths (NCCL all-gather + merge kernel; fused peer-memory
  stores over CUDA IPC + flag wait), on data with planted exact duplicates straddling the shard boundaries (index ties
  across shards: the lowest global index must win, matcher.cpp:494-498) and on a map so small that a trailing rank's
  shard is EMPTY;
* BASELINE config 2 as written: one frame sequence block-partitioned over the ranks with the boundary frame replicated,
  extraction + frame-to-frame top-2 per rank, stitched and compared with the oracle run over the whole sequence.

With >= 2 GPUs every rank owns a GPU and the host collectives run over NCCL; on a one-GPU box the same ranks share GPU 0
and the host collectives run over gloo (the fused path is unchanged: the exchange buffers are still mapped across
processes through CUDA IPC)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from vo_slam_test_b200 import sharded, synth

NF, FW, FH, NFEAT = 13, 320, 240, 300


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ndev, q, t, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("ORBX_PEER_TIMEOUT_MS", "15000")
    import torch
    import torch.distributed as dist
    import vo_slam_test_b200 as vo
    from vo_slam_test_b200 import api
    devi = rank % ndev
    torch.cuda.set_device(devi)
    dev = torch.device("cuda", devi)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {"rank": rank, "backend": dist.get_backend()}
    try:
        def host(o):
            return tuple(x.cpu().numpy() for x in o[:4])
        d_q = torch.from_numpy(q).to(dev)
        for name, tt in (("map", t), ("tiny", t[:1])):       # "tiny": M = 1 -> every rank but 0 holds an empty shard
            lo, hi = sharded.train_shard(len(tt), rank, world)
            d_t = torch.from_numpy(np.ascontiguousarray(tt[lo:hi])).to(dev)
            res[name + "_nccl"] = host(sharded.sharded_knn2_cuda(d_q, d_t, lo, 50, 0.7, dist, check=True))
            xchg = sharded.PeerExchange(dist, devi, len(q))
            try:
                for call in range(3):                        # both parities of the exchange buffer + reuse
                    o = sharded.sharded_knn2_peer(d_q, d_t, lo, 50, 0.7, xchg, check=True)
                    res["%s_peer%d" % (name, call)] = host(o)
            finally:
                xchg.close()
        # ---- config 2, strong scaling: this rank's frame block + halo ------------------------------------------------
        lo, hi, hi_ext, p_hi = sharded.strong_block(NF, rank, world)
        imgs = np.stack([synth.make_frame(100 + f, FH, FW) for f in range(lo, hi_ext)]) if hi_ext > lo else np.zeros((0, FH, FW), np.uint8)
        ex = vo.ORBextractor(NFEAT, 1.2, 8, 20, 7, device=devi)
        cap = ex.max_keypoints
        n = len(imgs)
        d_imgs = torch.from_numpy(imgs).to(dev)
        d_kps = torch.empty((max(n, 1), cap, 7), dtype=torch.float32, device=dev)
        d_desc = torch.empty((max(n, 1), cap, 32), dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        if n:
            ex.extract_batch_device(d_imgs.data_ptr(), n, FW, FH, FW, FW * FH, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr(), st)
        npairs = p_hi - lo
        mi = torch.full((max(npairs, 1), cap), -7, dtype=torch.int32, device=dev); m1 = torch.zeros_like(mi); m2 = torch.zeros_like(mi)
        mo = torch.zeros((max(npairs, 1), cap), dtype=torch.uint8, device=dev)
        if npairs > 0:
            qf = torch.arange(0, npairs, dtype=torch.int32, device=dev); tf = qf + 1
            api.knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), cap, qf.data_ptr(), tf.data_ptr(), npairs, 50, 0.7,
                                  mi.data_ptr(), m1.data_ptr(), m2.data_ptr(), mo.data_ptr(), st)
        torch.cuda.synchronize()
        res["frames"] = (lo, hi, hi_ext, p_hi, d_cnt.cpu().numpy()[:n], d_kps.cpu().numpy()[:n], d_desc.cpu().numpy()[:n],
                         mi.cpu().numpy()[:max(npairs, 0)], m1.cpu().numpy()[:max(npairs, 0)], m2.cpu().numpy()[:max(npairs, 0)],
                         mo.cpu().numpy()[:max(npairs, 0)])
        ex.close()
        dist.barrier()
    except Exception as e:      # noqa: BLE001
        import traceback
        res["error"] = traceback.format_exc() + repr(e)
    finally:
        out.put(res)
        dist.destroy_process_group()


def _planted(world):
    rng = np.random.default_rng(77)
    Q, M = 600, 24001
    t = synth.make_descriptors(M, seed=3)
    q = synth.flip_bits(t[rng.integers(0, M, Q)], rng.integers(0, 60, Q), rng)
    # exact duplicates of train rows on BOTH sides of every shard boundary: distance ties between shards
    for r in range(1, world):
        b, _ = sharded.train_shard(M, r, world)
        for k in range(40):
            src = int(rng.integers(0, M))
            t[b - 1 - k] = t[src]; t[b + k] = t[src]
            q[(r * 80 + 2 * k) % Q] = synth.flip_bits(t[src:src + 1], [int(rng.integers(0, 30))], rng)[0]
    q[:16] = t[:16]                                              # distance 0, duplicated below in the last shard
    t[M - 16:] = t[:16]
    return q, t


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_results_equal_oracle(world):
    import torch
    import torch.multiprocessing as mp
    ndev = torch.cuda.device_count()
    assert ndev >= 1
    if 1 < ndev < world:
        pytest.skip("%d ranks need either one shared GPU or >= %d GPUs" % (world, world))
    q, t = _planted(world)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndev, q, t, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert "error" not in r, r["error"]
    P = oracle.Port()
    want = {"map": P.knn2(q, t, 50, 0.7, nthreads=8), "tiny": P.knn2(q, t[:1], 50, 0.7)}
    assert want["map"][3].sum() > 50                              # the ratio test fires
    for r in res:
        for name in ("map", "tiny"):
            for key in [name + "_nccl"] + ["%s_peer%d" % (name, c) for c in range(3)]:
                for a, b, nm in zip(r[key], want[name], ["idx", "d1", "d2", "ok"]):
                    assert np.array_equal(a, b), (r["rank"], key, nm)
    # index ties across the boundary really occurred: best distance == second best with the best in a lower shard
    idx, d1, d2, _ = want["map"]
    ties = int(((d1 == d2) & (idx >= 0)).sum())
    assert ties >= 16
    # ---- config 2: stitched per-rank blocks == the oracle over the whole sequence ------------------------------------
    PX = oracle.Port(NFEAT)
    seq = [synth.make_frame(100 + f, FH, FW) for f in range(NF)]
    ref = [PX.extract(im) for im in seq]
    covered_frames, covered_pairs = set(), set()
    for r in res:
        lo, hi, hi_ext, p_hi, cnt, kps, desc, mi, m1, m2, mo = r["frames"]
        for j, f in enumerate(range(lo, hi_ext)):
            rk, rd = ref[f]
            assert cnt[j] == len(rk), (r["rank"], f)
            assert np.array_equal(np.ascontiguousarray(kps[j, :cnt[j]]).view(np.uint8).reshape(-1), rk.view(np.uint8).reshape(-1)), \
                (r["rank"], f, "keypoints")
            assert np.array_equal(desc[j, :cnt[j]], rd), (r["rank"], f, "descriptors")
            if f < hi:
                covered_frames.add(f)
        for j, p in enumerate(range(lo, p_hi)):
            w = P.knn2(ref[p][1], ref[p + 1][1], 50, 0.7)
            nq = len(ref[p][1])
            assert np.array_equal(mi[j, :nq], w[0]) and np.array_equal(m1[j, :nq], w[1]) and np.array_equal(m2[j, :nq], w[2]) \
                and np.array_equal(mo[j, :nq], w[3]), (r["rank"], p, "pair")
            assert p not in covered_pairs
            covered_pairs.add(p)
    assert covered_frames == set(range(NF)) and covered_pairs == set(range(NF - 1))
