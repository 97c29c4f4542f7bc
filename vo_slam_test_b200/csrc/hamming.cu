// hamming.cu -- brute-force 256-bit Hamming top-2 ("kNN k=2 + ratio test") for sm_100a.
//
// Replaces Matcher::computeDistance (matcher.cpp:1240-1256) inside the best/second-best loop of
// matcher.cpp:481-507, generalised to all pairs.  Pure integer work: XOR + POPC on eight
// 32-bit words per pair, no tensor cores.  One thread owns one query (8 registers); train descriptors are
// staged through shared memory and read as warp-wide broadcasts.
//
// Ordering contract (bit-exact with the sequential scan): train rows ascending, strict '<', first index wins.
// Realised with packed keys  key = dist << 23 | local_index : min() over keys prefers the smaller distance and,
// for equal distances, the smaller index; second best = min(best2, max(key, old best1)).
#include <atomic>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace orbx {

#ifndef ORBX_HAMM_QT
#define ORBX_HAMM_QT 128
#endif
#ifndef ORBX_HAMM_TT
#define ORBX_HAMM_TT 256
#endif
#ifndef ORBX_HAMM_STAGES
#define ORBX_HAMM_STAGES 3
#endif
constexpr int kQT = ORBX_HAMM_QT;          // queries per CTA (threads)
constexpr int kTT = ORBX_HAMM_TT;          // train descriptors per shared-memory tile (8 KB at 256)
constexpr int kStages = ORBX_HAMM_STAGES;  // tiles in flight (TMA bulk copies, mbarrier-tracked)
constexpr uint32_t kKeyInit = (256u << 23) | 0x7FFFFFu;

static std::atomic<long long> g_hamm_launches{0};

// ---- mbarrier + 1-D bulk copy (TMA engine: cp.async.bulk -> UBLKCP) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 256-bit Hamming distance with a carry-save adder tree: 7 of the 8 XOR words are compressed 3:2 (LOP3 0x96 / 0xE8)
// so that only 5 POPC are needed instead of 8 -- POPC is the scarce pipe (16/clk/SM), see DESIGN.md.
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
__device__ __forceinline__ int hamming8(const uint32_t (&q)[8], const uint4 a, const uint4 b) {
  const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
  const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
  const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
  const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
  const uint32_t s3 = s1 ^ s2 ^ x6, c3 = maj3(s1, s2, x6);
  return (__popc(s3) + __popc(x7)) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__device__ __forceinline__ void top2_update(uint32_t& k1, uint32_t& k2, uint32_t key) {
  k2 = min(k2, max(key, k1));
  k1 = min(k1, key);
}

struct TileRing {
  uint4* tile;        // kStages * kTT * 2 uint4
  uint64_t* full;     // kStages barriers
};

// Scan train rows [0, nt) (nt < 2^23) for the CTA's queries.  Tiles of kTT rows are fetched by the TMA engine
// (one elected thread issues cp.async.bulk, completion is tracked by an mbarrier per stage), kStages deep.
__device__ __forceinline__ void scan_train(const uint32_t (&q)[8], const uint8_t* __restrict__ train, int nt, const TileRing& R,
                                           uint32_t& k1, uint32_t& k2) {
  const int tid = threadIdx.x;
  const int ntiles = (nt + kTT - 1) / kTT;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&R.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kStages;
    const uint32_t bytes = (uint32_t)min(kTT, nt - t * kTT) * 32u;
    mbar_expect_tx(&R.full[s], bytes);
    bulk_g2s(R.tile + (size_t)s * kTT * 2, train + (size_t)t * kTT * 32, bytes, &R.full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < min(kStages, ntiles); ++t) issue(t);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&R.full[s], (uint32_t)((t / kStages) & 1));
    const uint4* tile = R.tile + (size_t)s * kTT * 2;
    const int cnt = min(kTT, nt - t * kTT);
    const int base = t * kTT;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const uint4 a = tile[2 * j], b = tile[2 * j + 1];
      const uint32_t d = (uint32_t)hamming8(q, a, b);
      top2_update(k1, k2, (d << 23) | (uint32_t)(base + j));
    }
    __syncthreads();                                   // everyone is done reading stage s
    if (tid == 0 && t + kStages < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
      issue(t + kStages);
    }
  }
}

__device__ __forceinline__ void load_query(uint32_t (&q)[8], const uint8_t* __restrict__ qs, int i, int nq) {
  if (i < nq) {
    const uint4* p = reinterpret_cast<const uint4*>(qs + (size_t)i * 32);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = 0;
  }
}

__device__ __forceinline__ uint8_t accept(int d1, int d2, int th, float ratio) {
  return (d1 <= th && (float)d1 < __fmul_rn(ratio, (float)d2)) ? 1 : 0;   // matcher.cpp:504-506
}

// grid (ceil(nq/kQT), nsplit).  Split s scans train rows [s*per, min((s+1)*per, nt)).
// nsplit == 1: final outputs are written directly; otherwise partial (idx, d1, d2) go to the workspace.
__global__ void __launch_bounds__(kQT) knn2_kernel(const uint8_t* __restrict__ qs, int nq, const uint8_t* __restrict__ train,
                                                   long long nt, long long per, int th, float ratio, int32_t* idx,
                                                   int32_t* d1, int32_t* d2, uint8_t* ok, int32_t* part) {
  __shared__ __align__(128) uint4 tile[kStages * kTT * 2];
  __shared__ __align__(8) uint64_t full[kStages];
  const TileRing ring{tile, full};
  const int i = blockIdx.x * kQT + threadIdx.x;
  const int s = blockIdx.y;
  const long long t0 = (long long)s * per;
  const int cnt = (int)min(per, nt - t0);
  uint32_t q[8];
  load_query(q, qs, i, nq);
  uint32_t k1 = kKeyInit, k2 = kKeyInit;
  scan_train(q, train + (size_t)t0 * 32, cnt, ring, k1, k2);
  if (i >= nq) return;
  const int bd1 = (int)(k1 >> 23), bd2 = (int)(k2 >> 23);
  const int bi = bd1 < 256 ? (int)(t0 + (long long)(k1 & 0x7FFFFFu)) : -1;
  if (gridDim.y == 1) {
    idx[i] = bi; d1[i] = bd1; d2[i] = bd2;
    ok[i] = accept(bd1, bd2, th, ratio);
  } else {
    int32_t* p = part + ((size_t)s * nq + i) * 3;
    p[0] = bi; p[1] = bd1; p[2] = bd2;
  }
}

// Merge per-split / per-shard partial results (ascending train ranges): best = smallest d1, lowest range
// wins ties (== lowest index); second = 2nd smallest of the multiset {d1_s, d2_s}.
__global__ void knn2_merge_kernel(const int32_t* __restrict__ pidx, const int32_t* __restrict__ pd1,
                                  const int32_t* __restrict__ pd2, int stride, int nparts, int nq, int th, float ratio,
                                  int32_t* idx, int32_t* d1, int32_t* d2, uint8_t* ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int b1 = 256, b2 = 256, bi = -1;
  for (int s = 0; s < nparts; ++s) {
    const size_t o = ((size_t)s * nq + i) * stride;
    const int a1 = pd1[o], a2 = pd2[o], ai = pidx[o];
    if (a1 < b1) { b2 = min(b1, a2); b1 = a1; bi = ai; }
    else { b2 = min(b2, a1); }
  }
  idx[i] = bi; d1[i] = b1; d2[i] = b2;
  ok[i] = accept(b1, b2, th, ratio);
}

// Frame-to-frame matching over extractor output: pair p = (query frame qf[p], train frame tf[p]).
__global__ void __launch_bounds__(kQT) knn2_pairs_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ counts,
                                                         int cap, const int32_t* __restrict__ qf, const int32_t* __restrict__ tf,
                                                         int th, float ratio, int32_t* idx, int32_t* d1, int32_t* d2,
                                                         uint8_t* ok) {
  __shared__ __align__(128) uint4 tile[kStages * kTT * 2];
  __shared__ __align__(8) uint64_t full[kStages];
  const TileRing ring{tile, full};
  const int p = blockIdx.y;
  const int fq = qf[p], ft = tf[p];
  const int nq = min(counts[fq], cap), nt = min(counts[ft], cap);
  if (blockIdx.x * kQT >= nq) return;
  const int i = blockIdx.x * kQT + threadIdx.x;
  uint32_t q[8];
  load_query(q, desc + (size_t)fq * cap * 32, i, nq);
  uint32_t k1 = kKeyInit, k2 = kKeyInit;
  scan_train(q, desc + (size_t)ft * cap * 32, nt, ring, k1, k2);
  if (i >= nq) return;
  const int bd1 = (int)(k1 >> 23), bd2 = (int)(k2 >> 23);
  const size_t o = (size_t)p * cap + i;
  idx[o] = bd1 < 256 ? (int)(k1 & 0x7FFFFFu) : -1;
  d1[o] = bd1; d2[o] = bd2;
  ok[o] = accept(bd1, bd2, th, ratio);
}

// ---- sharded top-2 with the exchange fused into the kernels (no NCCL on the data path) -------------------------------
// Every rank owns an exchange buffer in its HBM that all peers have mapped (CUDA IPC over NVLink / NVSwitch):
//   int32 [0, 2*kMaxPeers)   flags[parity][source rank] = epoch of the last completed scatter of that source
//   int32 [32]               CTA counter of the local scatter kernel
//   int32 [kXchgHdr ...)     records[parity][source rank][max_queries][3] = (global idx, d1, d2)
// knn2_merge_scatter_kernel merges the per-split partials of the local shard and STORES each record straight into all
// peers' buffers (P2P stores), fences at system scope; the last CTA then releases flags[parity][rank] = epoch on every peer.
// knn2_gather_merge_kernel acquires the flags of all sources in its own buffer (bounded spin) and merges the records.
// Two parities alternate per call: a rank that is one call ahead writes the other half, and it cannot get two calls
// ahead because its gather of call n+1 waits for every peer's scatter of call n+1.
constexpr int kMaxPeers = 16;
constexpr int kXchgHdr = 64;
struct PeerSet { int32_t* p[kMaxPeers]; };
__host__ __device__ inline size_t xchg_rec_off(int parity, int src, int world, int maxq) {
  return (size_t)kXchgHdr + ((size_t)(parity * world + src) * maxq) * 3;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void knn2_merge_scatter_kernel(const int32_t* __restrict__ pidx, const int32_t* __restrict__ pd1,
                                          const int32_t* __restrict__ pd2, int stride, int nparts, int nq, long long shard_lo,
                                          PeerSet peers, int rank, int world, int maxq, int parity, int epoch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) {
    int b1 = 256, b2 = 256, bi = -1;
    for (int s = 0; s < nparts; ++s) {
      const size_t o = ((size_t)s * nq + i) * stride;
      const int a1 = pd1[o], a2 = pd2[o], ai = pidx[o];
      if (a1 < b1) { b2 = min(b1, a2); b1 = a1; bi = ai; }
      else { b2 = min(b2, a1); }
    }
    const int gi = bi >= 0 ? (int)(shard_lo + bi) : -1;       // local -> global row index
    const size_t o = xchg_rec_off(parity, rank, world, maxq) + (size_t)3 * i;
    for (int p = 0; p < world; ++p) {
      int32_t* r = peers.p[p] + o;
      r[0] = gi; r[1] = b1; r[2] = b2;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t* done = peers.p[rank] + 32;
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {           // last CTA: every record of this rank is written and fenced
      atomicExch(done, 0);
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys(peers.p[p] + parity * kMaxPeers + rank, epoch);
    }
  }
}

__global__ void knn2_gather_merge_kernel(const int32_t* __restrict__ own, int world, int nq, int maxq, int parity, int epoch, int th,
                                         float ratio, int32_t* idx, int32_t* d1, int32_t* d2, uint8_t* ok, int32_t* status, long long spin_ticks) {
  // one thread per source rank; a bounded spin (spin_ticks of clock64, set by the host: ORBX_PEER_TIMEOUT_MS, default 20 s --
  // long enough for a peer's first-call module load) instead of a hang.  On a timeout NOTHING stale is merged: every query of
  // this CTA gets the sentinel record (idx -1, d1 = d2 = 256, ok 0) and *status = 1; the exchange must then be torn down
  // (sharded.py raises on every rank together).
  __shared__ int timedOut;
  if (threadIdx.x == 0) timedOut = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int32_t* flag = own + parity * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != epoch) {
      if (clock64() - t0 > spin_ticks) { atomicExch(status, 1); timedOut = 1; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  if (timedOut) { idx[i] = -1; d1[i] = 256; d2[i] = 256; ok[i] = 0; return; }
  int b1 = 256, b2 = 256, bi = -1;
  for (int s = 0; s < world; ++s) {                           // sources in ascending train-index order
    const int32_t* r = own + xchg_rec_off(parity, s, world, maxq) + (size_t)3 * i;
    const int ai = __ldcg(r), a1 = __ldcg(r + 1), a2 = __ldcg(r + 2);   // L2 only: peers wrote these lines
    if (a1 < b1) { b2 = min(b1, a2); b1 = a1; bi = ai; }
    else { b2 = min(b2, a1); }
  }
  idx[i] = bi; d1[i] = b1; d2[i] = b2;
  ok[i] = accept(b1, b2, th, ratio);
}

static int pick_splits(int nq, long long nt, long long* per_out) {
  if (nt <= 0 || nq <= 0) { *per_out = kTT; return 1; }
  const int qtiles = (nq + kQT - 1) / kQT;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // aim at ~8 resident CTAs per SM, each split at least 4 tiles and below 2^23 rows
  long long want = ((long long)sms * 8 + qtiles - 1) / qtiles;
  long long maxSplits = (nt + 4LL * kTT - 1) / (4LL * kTT);
  long long ns = std::max(1LL, std::min(want, maxSplits));
  long long per = (nt + ns - 1) / ns;
  per = (per + kTT - 1) / kTT * kTT;
  const long long kMaxPer = (1LL << 23) - kTT;
  if (per > kMaxPer) per = kMaxPer;
  ns = (nt + per - 1) / per;
  if (ns < 1) ns = 1;
  *per_out = per;
  return (int)ns;
}

}  // namespace orbx

using namespace orbx;

extern "C" {

long long hamm_launch_count(void) { return g_hamm_launches.load(); }

size_t hamm_knn2_workspace_bytes(int nq, long long nt) {
  long long per;
  const int ns = pick_splits(nq, nt, &per);
  return ns > 1 ? (size_t)ns * nq * 3 * sizeof(int32_t) : 0;
}

int hamm_knn2_device(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, int th, float ratio, int32_t* d_idx,
                     int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, void* d_workspace, size_t workspace_bytes, void* stream) {
  // an empty train set (nt == 0: a trailing rank of a small sharded map) is legal and may come with a null pointer
  if (!d_q || (!d_t && nt > 0) || !d_idx || !d_d1 || !d_d2 || !d_ok || nq < 0 || nt < 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (nq == 0) return ORBX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long per;
  const int ns = pick_splits(nq, nt, &per);
  if (ns > 1 && (!d_workspace || workspace_bytes < (size_t)ns * nq * 3 * sizeof(int32_t))) {
    set_error("workspace too small (see hamm_knn2_workspace_bytes)");
    return ORBX_ERR_CAPACITY;
  }
  dim3 grid((nq + kQT - 1) / kQT, ns);
  knn2_kernel<<<grid, kQT, 0, st>>>(d_q, nq, d_t, nt, per, th, ratio, d_idx, d_d1, d_d2, d_ok, (int32_t*)d_workspace);
  g_hamm_launches++;
  if (ns > 1) {
    const int32_t* part = (const int32_t*)d_workspace;
    knn2_merge_kernel<<<(nq + 255) / 256, 256, 0, st>>>(part, part + 1, part + 2, 3, ns, nq, th, ratio, d_idx, d_d1, d_d2, d_ok);
    g_hamm_launches++;
  }
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int hamm_knn2(const uint8_t* q, int nq, const uint8_t* t, long long nt, int th, float ratio, int32_t* idx, int32_t* d1,
              int32_t* d2, uint8_t* ok, int device) {
  if (!q || (!t && nt > 0) || !idx || !d1 || !d2 || !ok || nq < 0 || nt < 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (nq == 0) return ORBX_OK;
  // grow-only per-thread arenas (one packed upload, one packed download) instead of 7 cudaMalloc/cudaFree + 6 copies per call:
  // this entry point is what Matcher::computeDistance / matchTop2 of the adapter call per descriptor set
  static thread_local DevArena arena;
  static thread_local HostArena host;
  const size_t wsb = hamm_knn2_workspace_bytes(nq, nt);
  const size_t qB = align_up_sz((size_t)nq * 32, 256), tB = align_up_sz((size_t)nt * 32, 256);
  const size_t outB = align_up_sz((size_t)nq * 13, 256);                 // idx, d1, d2 (int32) + ok (u8), one block
  if (arena.reserve(qB + tB + outB + wsb + 1024, device) || host.reserve(qB + tB + outB)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  uint8_t* db = arena.take<uint8_t>(qB + tB);
  uint8_t* dout = arena.take<uint8_t>(outB);
  void* ws = wsb ? (void*)arena.take<uint8_t>(wsb) : nullptr;
  memcpy(host.base, q, (size_t)nq * 32);
  if (nt) memcpy(host.base + qB, t, (size_t)nt * 32);
  cudaStream_t st = nullptr;
  ORBX_CUDA(cudaMemcpyAsync(db, host.base, qB + (nt ? (size_t)nt * 32 : 0), cudaMemcpyHostToDevice, st));
  int32_t *di = (int32_t*)dout, *dd1 = di + nq, *dd2 = dd1 + nq;
  uint8_t* dok = (uint8_t*)(dd2 + nq);
  const int rc = hamm_knn2_device(db, nq, db + qB, nt, th, ratio, di, dd1, dd2, dok, ws, wsb, st);
  if (rc != ORBX_OK) return rc;
  uint8_t* hout = host.base + qB + tB;
  ORBX_CUDA(cudaMemcpyAsync(hout, dout, (size_t)nq * 13, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  memcpy(idx, hout, sizeof(int32_t) * nq); memcpy(d1, hout + 4 * (size_t)nq, sizeof(int32_t) * nq);
  memcpy(d2, hout + 8 * (size_t)nq, sizeof(int32_t) * nq); memcpy(ok, hout + 12 * (size_t)nq, nq);
  return ORBX_OK;
}

int hamm_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int cap, const int32_t* d_qf, const int32_t* d_tf,
                           int npairs, int th, float ratio, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok,
                           void* stream) {
  if (!d_desc || !d_counts || !d_qf || !d_tf || !d_idx || !d_d1 || !d_d2 || !d_ok || cap <= 0 || npairs < 0) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  if (npairs == 0) return ORBX_OK;
  dim3 grid((cap + kQT - 1) / kQT, npairs);
  nvtxRangePushA("orbx:hamming_pairs");
  knn2_pairs_kernel<<<grid, kQT, 0, (cudaStream_t)stream>>>(d_desc, d_counts, cap, d_qf, d_tf, th, ratio, d_idx, d_d1, d_d2, d_ok);
  nvtxRangePop();
  g_hamm_launches++;
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int hamm_knn2_merge_device(const int32_t* d_idx_in, const int32_t* d_d1_in, const int32_t* d_d2_in, int nshards, int nq, int th,
                           float ratio, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, void* stream) {
  if (!d_idx_in || !d_d1_in || !d_d2_in || !d_idx || !d_d1 || !d_d2 || !d_ok || nshards < 1 || nq < 0) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  if (nq == 0) return ORBX_OK;
  knn2_merge_kernel<<<(nq + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_idx_in, d_d1_in, d_d2_in, 1, nshards, nq, th, ratio,
                                                                       d_idx, d_d1, d_d2, d_ok);
  g_hamm_launches++;
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

size_t hamm_exchange_bytes(int world, int max_queries) {
  if (world < 1 || world > kMaxPeers || max_queries < 1) return 0;
  return sizeof(int32_t) * (xchg_rec_off(2, 0, world, max_queries));
}

int hamm_exchange_alloc(int device, int world, int max_queries, void** buf, unsigned char ipc_handle[64]) {
  const size_t bytes = hamm_exchange_bytes(world, max_queries);
  if (!buf || !ipc_handle || bytes == 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  ORBX_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  ORBX_CUDA(cudaMalloc(&p, bytes));
  ORBX_CUDA(cudaMemset(p, 0, bytes));
  ORBX_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t hdl;
  cudaError_t e = cudaIpcGetMemHandle(&hdl, p);
  if (e != cudaSuccess) { cudaFree(p); set_error(cudaGetErrorString(e)); return ORBX_ERR_CUDA; }
  memcpy(ipc_handle, &hdl, 64);
  *buf = p;
  return ORBX_OK;
}

int hamm_exchange_open(int device, const unsigned char ipc_handle[64], void** peer_buf) {
  if (!ipc_handle || !peer_buf) { set_error("bad argument"); return ORBX_ERR_ARG; }
  ORBX_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t hdl;
  memcpy(&hdl, ipc_handle, 64);
  ORBX_CUDA(cudaIpcOpenMemHandle(peer_buf, hdl, cudaIpcMemLazyEnablePeerAccess));
  return ORBX_OK;
}

int hamm_exchange_close(void* peer_buf) {
  if (peer_buf) ORBX_CUDA(cudaIpcCloseMemHandle(peer_buf));
  return ORBX_OK;
}

int hamm_exchange_free(void* buf) {
  if (buf) ORBX_CUDA(cudaFree(buf));
  return ORBX_OK;
}

int hamm_knn2_sharded_device(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, long long shard_lo, int th, float ratio,
                             int rank, int world, void* const* bufs, int max_queries, int epoch, int32_t* d_idx, int32_t* d_d1,
                             int32_t* d_d2, uint8_t* d_ok, int32_t* d_status, void* d_workspace, size_t workspace_bytes,
                             void* stream) {
  return hamm_knn2_sharded_phases_device(d_q, nq, d_t, nt, shard_lo, th, ratio, rank, world, bufs, max_queries, epoch, d_idx, d_d1,
                                         d_d2, d_ok, d_status, d_workspace, workspace_bytes, stream, 3);
}

int hamm_knn2_sharded_phases_device(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, long long shard_lo, int th,
                                    float ratio, int rank, int world, void* const* bufs, int max_queries, int epoch, int32_t* d_idx,
                                    int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, int32_t* d_status, void* d_workspace,
                                    size_t workspace_bytes, void* stream, int phases) {
  if (!d_q || (!d_t && nt > 0) || !d_idx || !d_d1 || !d_d2 || !d_ok || !d_status || !bufs || nq < 0 || nt < 0 || world < 1 ||
      world > kMaxPeers || rank < 0 || rank >= world || nq > max_queries || epoch < 1) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  if (nq == 0) return ORBX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PeerSet peers{};
  for (int p = 0; p < world; ++p) {
    if (!bufs[p]) { set_error("null exchange buffer"); return ORBX_ERR_ARG; }
    peers.p[p] = (int32_t*)bufs[p];
  }
  long long per;
  const int ns = pick_splits(nq, nt, &per);
  if (ns > 1 && (!d_workspace || workspace_bytes < (size_t)ns * nq * 3 * sizeof(int32_t))) {
    set_error("workspace too small (see hamm_knn2_workspace_bytes)");
    return ORBX_ERR_CAPACITY;
  }
  const int parity = epoch & 1;
  if (phases & 1) {
    // local shard: per-split partials (ns > 1) or the shard result in the output arrays (ns == 1), local row indices
    dim3 grid((nq + kQT - 1) / kQT, ns);
    knn2_kernel<<<grid, kQT, 0, st>>>(d_q, nq, d_t, nt, per, th, ratio, d_idx, d_d1, d_d2, d_ok, (int32_t*)d_workspace);
    const int32_t* part = (const int32_t*)d_workspace;
    if (ns > 1)
      knn2_merge_scatter_kernel<<<(nq + 127) / 128, 128, 0, st>>>(part, part + 1, part + 2, 3, ns, nq, shard_lo, peers, rank, world,
                                                                 max_queries, parity, epoch);
    else
      knn2_merge_scatter_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_idx, d_d1, d_d2, 1, 1, nq, shard_lo, peers, rank, world,
                                                                 max_queries, parity, epoch);
    g_hamm_launches += 2;
  }
  if (phases & 2) {
    static const long long spinTicks = []() {
      const char* e = getenv("ORBX_PEER_TIMEOUT_MS");
      const long long ms = e && atoll(e) > 0 ? atoll(e) : 20000;
      return ms * 2000000LL;                                 // clock64 ticks at <= 2 GHz: a lower bound of the wait
    }();
    knn2_gather_merge_kernel<<<(nq + 127) / 128, 128, 0, st>>>(peers.p[rank], world, nq, max_queries, parity, epoch, th, ratio, d_idx,
                                                              d_d1, d_d2, d_ok, d_status, spinTicks);
    g_hamm_launches += 1;
  }
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

}  // extern "C"
