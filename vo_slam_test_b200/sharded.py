"""Multi-GPU host logic: frame sharding (no collective) and map-scale sharded Hamming top-2 (one all-gather).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests).

* Frames (BASELINE config 2) are independent units: rank r owns a contiguous block; nothing is exchanged.
* Map-scale brute-force kNN (BASELINE config 5): the train set is split into `world` contiguous index ranges
  (rank r holds rows [r*per, min((r+1)*per, M))); queries are replicated.  Every rank computes its local top-2 with
  global indices, the [Q,3] int32 records (idx, d1, d2) are all-gathered (12 B per query per rank) and every rank
  runs the same merge: smallest d1 wins, lowest index wins ties (== the sequential scan of matcher.cpp:494-498),
  second best = 2nd smallest of the multiset {d1_s, d2_s}.
"""
import numpy as np


def frame_block(nframes, rank, world):
    """Contiguous block [lo, hi) of `nframes` frames owned by `rank` (block sizes differ by at most one)."""
    base, rem = divmod(nframes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pair_block(nframes, rank, world):
    """Frame-to-frame pairs (p, p+1) owned by `rank`: those whose FIRST frame is in the rank's block.  The block's
    last pair needs frame hi, which belongs to the next rank: that boundary frame is replicated (returned as `halo`)."""
    lo, hi = frame_block(nframes, rank, world)
    p_hi = min(hi, nframes - 1)
    halo = hi if (hi < nframes and hi > lo) else None
    return lo, max(p_hi, lo), halo


def train_shard(M, rank, world):
    """Rows [lo, hi) of the train/map descriptor set held by `rank`."""
    per = (M + world - 1) // world
    lo = min(rank * per, M)
    return lo, min(lo + per, M)


def merge_top2_numpy(idx, d1, d2):
    """Merge rule on host arrays [S, Q] (used by the CPU tests of the sharding logic; the GPU path runs
    hamm_knn2_merge_device).  Shards must be in ascending train-index order."""
    S, Q = d1.shape
    b1 = np.full(Q, 256, np.int32); b2 = np.full(Q, 256, np.int32); bi = np.full(Q, -1, np.int32)
    for s in range(S):
        better = d1[s] < b1
        b2 = np.where(better, np.minimum(b1, d2[s]), np.minimum(b2, d1[s]))
        bi = np.where(better, idx[s], bi)
        b1 = np.where(better, d1[s], b1)
    return bi, b1, b2


def allgather_records(rec, dist):
    """All-gather a [Q,3] int32 tensor (idx, d1, d2) from every rank -> [world, Q, 3].  NCCL gathers device tensors in
    place; gloo (CPU tests, or several ranks sharing one GPU) has no device all-gather, so device records take the host
    route there."""
    import torch
    world = dist.get_world_size()
    rec = rec.contiguous()
    if rec.is_cuda and dist.get_backend() != "gloo":
        out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec)
        return out
    host = rec.cpu()
    out = torch.empty((world,) + tuple(host.shape), dtype=host.dtype)
    dist.all_gather(list(out.unbind(0)), host)
    return out.to(rec.device)


def strong_block(nframes, rank, world):
    """BASELINE config 2 as SURVEY section 8(d,e) states it: ONE batch of `nframes` frames block-partitioned over `world`
    ranks.  Rank r extracts frames [lo, hi_ext) = its block plus the replicated boundary frame (the halo: first frame of the
    next rank's block) and matches the pairs (p, p+1) for p in [lo, p_hi); with local frame numbering (frame lo -> 0) that
    is pairs 0 .. p_hi-lo-1.  Returns (lo, hi, hi_ext, p_hi).  No data-path collective: every pair is owned by exactly one
    rank (matcher.cpp:481-507 is a loop over independent query/train sets)."""
    lo, hi = frame_block(nframes, rank, world)
    plo, p_hi, halo = pair_block(nframes, rank, world)
    hi_ext = hi + 1 if halo is not None else hi
    return lo, hi, hi_ext, p_hi


def sharded_knn2_cuda(d_q, d_t_local, shard_lo, th, ratio, dist, stream=None, check=False):
    """Local shard top-2 on this rank's GPU, NCCL all-gather of the records, merge kernel.  torch CUDA tensors in,
    (idx, d1, d2, ok) torch CUDA tensors out.  d_t_local holds train rows [shard_lo, shard_lo + len); an EMPTY shard (a
    trailing rank when the map is small relative to the world, train_shard(9, 7, 8)) is legal.

    A rank whose local kernel call fails must not leave its peers waiting in the collective: it contributes records with
    the error marker idx = -2 and raises only AFTER the all-gather; with check=True every rank looks for the marker (one
    sync) and all ranks raise together."""
    import torch
    from . import api
    Q, Ml = d_q.shape[0], d_t_local.shape[0]
    dev = d_q.device
    idx = torch.empty(Q, dtype=torch.int32, device=dev); d1 = torch.empty_like(idx); d2 = torch.empty_like(idx)
    ok = torch.empty(Q, dtype=torch.uint8, device=dev)
    wsb = api.knn2_workspace_bytes(Q, Ml)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    err = None
    try:
        api.knn2_device(d_q.data_ptr(), Q, d_t_local.data_ptr() if Ml > 0 else 0, Ml, th, ratio, idx.data_ptr(), d1.data_ptr(),
                        d2.data_ptr(), ok.data_ptr(), ws.data_ptr(), wsb, st)
    except Exception as e:      # noqa: BLE001
        if world == 1:
            raise
        err = e
        idx.fill_(-2); d1.fill_(256); d2.fill_(256)
    else:
        idx = torch.where(idx >= 0, idx + shard_lo, idx)      # local -> global row index
    rec = torch.stack([idx, d1, d2], dim=1).contiguous()
    if world == 1:
        return idx, d1, d2, ok
    allrec = allgather_records(rec, dist)                      # [world, Q, 3]
    if err is not None:
        raise err
    if check and Q > 0 and bool((allrec[:, :, 0] == -2).any().item()):
        raise RuntimeError("sharded top-2: the local kernel of a peer rank failed")
    parts = allrec.permute(2, 0, 1).contiguous()               # [3, world, Q]
    api.knn2_merge_device(parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), world, Q, th, ratio,
                          idx.data_ptr(), d1.data_ptr(), d2.data_ptr(), ok.data_ptr(), st)
    return idx, d1, d2, ok


class PeerExchange:
    """Exchange buffers of the fused sharded top-2 (hamm_knn2_sharded_device): every rank allocates one buffer in its HBM,
    the 64-byte CUDA IPC handles are all-gathered once on the host side, and every rank maps all peers' buffers.  After
    that the per-call data path has no collective: the merge kernel stores its records into the peers' buffers over
    NVLink and the final merge waits on per-source flags in its own buffer."""

    def __init__(self, dist, device, max_queries):
        from . import api
        self.dist, self.device, self.max_queries = dist, int(device), int(max_queries)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.bufs, self.own, self.epoch, self.broken = None, None, 0, False
        # The flag protocol orders call n's gather before call n+1's scatter only inside ONE stream: the stream that is
        # current at construction is remembered and every call must be issued on it (sharded_knn2_peer checks).
        self.stream = None
        try:
            import torch
            if torch.cuda.is_available():
                self.stream = torch.cuda.current_stream(self.device).cuda_stream
        except Exception:       # noqa: BLE001
            pass
        # Every rank runs the SAME sequence of collectives whatever fails locally (a rank that raised early would leave
        # its peers waiting in a collective): local failures are recorded and agreed on afterwards.
        handle, err = None, None
        try:
            self.own, handle = api.exchange_alloc(self.device, self.world, self.max_queries)
        except Exception as e:      # noqa: BLE001
            err = repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        bufs = [None] * self.world
        if err is None and all(h is not None for h in handles):
            try:
                for r in range(self.world):
                    bufs[r] = self.own if r == self.rank else api.exchange_open(self.device, handles[r])
            except Exception as e:  # noqa: BLE001
                err = repr(e)
        elif err is None:
            err = "a peer could not allocate its exchange buffer"
        errs = [None] * self.world
        dist.all_gather_object(errs, err)          # also the barrier: nobody scatters before every rank has mapped every buffer
        if any(e is not None for e in errs):
            for r, b in enumerate(bufs):
                if b is not None and r != self.rank:
                    api.exchange_close(b)
            if self.own is not None:
                api.exchange_free(self.own)
            self.own = None
            raise RuntimeError("peer exchange set-up failed: %s" % next(e for e in errs if e is not None))
        self.bufs = bufs

    def close(self):
        from . import api
        if self.bufs is None:
            return
        self.dist.barrier()                  # nobody unmaps while a peer may still store into it
        for r, b in enumerate(self.bufs):
            if r != self.rank:
                api.exchange_close(b)
        self.dist.barrier()
        api.exchange_free(self.own)
        self.bufs = None


def sharded_knn2_peer(d_q, d_t_local, shard_lo, th, ratio, xchg, check=True):
    """Same contract as sharded_knn2_cuda, but the exchange is fused into the kernels (xchg = PeerExchange).

    check=True reads the status word back (one sync), agrees on it across ranks (one small all-reduce) and raises on EVERY
    rank if any rank's gather timed out; the exchange is then marked broken and must be closed (epochs of the ranks may no
    longer be in step).  check=False is the pure data path: a timed-out gather writes sentinel records (idx -1, ok 0), never
    stale ones, and leaves the device status word in xchg.last_status for the caller to look at."""
    import torch
    from . import api
    if xchg.broken:
        raise RuntimeError("peer exchange is broken after a timeout: close it and build a new one")
    cur = torch.cuda.current_stream().cuda_stream
    if xchg.stream is not None and cur != xchg.stream:
        raise RuntimeError("every call on a PeerExchange must be issued on the stream it was created on")
    Q, Ml = d_q.shape[0], d_t_local.shape[0]
    dev = d_q.device
    idx = torch.empty(Q, dtype=torch.int32, device=dev); d1 = torch.empty_like(idx); d2 = torch.empty_like(idx)
    ok = torch.empty(Q, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = api.knn2_workspace_bytes(Q, Ml)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    xchg.epoch += 1
    api.knn2_sharded_device(d_q.data_ptr(), Q, d_t_local.data_ptr() if Ml > 0 else 0, Ml, int(shard_lo), th, ratio, xchg.rank,
                            xchg.world, xchg.bufs, xchg.max_queries, xchg.epoch, idx.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                            ok.data_ptr(), status.data_ptr(), ws.data_ptr(), wsb, cur)
    if check:
        if xchg.world > 1:
            xchg.dist.all_reduce(status, op=xchg.dist.ReduceOp.MAX)
        if int(status.item()) != 0:
            xchg.broken = True
            raise RuntimeError("sharded top-2: a peer's records did not arrive (exchange timed out)")
    xchg.last_status = status
    return idx, d1, d2, ok
