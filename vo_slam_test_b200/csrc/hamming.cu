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
#ifndef ORBX_HAMM_MINB
#define ORBX_HAMM_MINB 9          // min CTAs per SM of the POPC kernels (an explicit 1 lets ptxas take more registers: 4.91 -> 5.01 ms)
#endif
#ifndef ORBX_HAMM_TRACE
#define ORBX_HAMM_TRACE 0
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
                                           uint32_t& k1, uint32_t& k2, int idx0 = 0) {
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
    const int base = idx0 + t * kTT;
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
__global__ void __launch_bounds__(kQT, ORBX_HAMM_MINB) knn2_kernel(const uint8_t* __restrict__ qs, int nq, const uint8_t* __restrict__ train,
                                                   long long nt, long long per, int th, float ratio, int32_t* idx,
                                                   int32_t* d1, int32_t* d2, uint8_t* ok, int32_t* part) {
  __shared__ __align__(128) uint4 tile[kStages * kTT * 2];
  __shared__ __align__(8) uint64_t full[kStages];
  const TileRing ring{tile, full};
  const int i = blockIdx.x * kQT + threadIdx.x;
  const int s = blockIdx.y;
  const long long t0 = (long long)s * per;
  const int cnt = (int)min(per, nt - t0);
#if ORBX_HAMM_TRACE
  unsigned long long tr0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr0));
#endif
  uint32_t q[8];
  load_query(q, qs, i, nq);
  uint32_t k1 = kKeyInit, k2 = kKeyInit;
  scan_train(q, train + (size_t)t0 * 32, cnt, ring, k1, k2);
#if ORBX_HAMM_TRACE
  if (threadIdx.x == 0 && gridDim.y > 64) {      // debugging aid: one line per CTA of a large map scan (SM, start, end in ns)
    unsigned long long tr1; unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    printf("HT %d %d %u %llu %llu\n", (int)blockIdx.x, (int)blockIdx.y, sm, tr0, tr1);
  }
#endif
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

// Map-scale scans: the same work for every CTA does NOT mean the same finishing time.  The warp schedulers favour some warps
// (cycle stamps of the static-split kernel on 1k x 2M rows: all 1176 CTAs start within 0.4 us, the first one is done after
// 0.79 ms, the median after 1.80 ms, the last after 2.89 ms), so SMs run out of resident warps long before the kernel ends: 21
// resident warps per SM on average against 32 for the frame-pair kernel, whose grid keeps refilling the SMs.  Here the rows
// are handed out in blocks of kDynTiles tiles (2 or 4: 793 ... 800 Gmatch/s, inside the run-to-run spread) from one counter per query tile: a CTA that gets ahead simply scans more
// blocks, and all CTAs finish within one block of each other.  A CTA's blocks are not contiguous, so its keys carry the row
// index relative to the whole set (nt < 2^23) and the merge breaks distance ties by index.
#ifndef ORBX_HAMM_DYN_TILES
#define ORBX_HAMM_DYN_TILES 2
#endif
constexpr int kDynTiles = ORBX_HAMM_DYN_TILES;
__global__ void __launch_bounds__(kQT, ORBX_HAMM_MINB) knn2_dyn_kernel(const uint8_t* __restrict__ qs, int nq, const uint8_t* __restrict__ train, int nt,
                                                       int nblocks, int* __restrict__ counters, int32_t* __restrict__ part) {
  __shared__ __align__(128) uint4 tile[kStages * kTT * 2];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int s_base[kStages], s_cnt[kStages];     // first row / rows of the tile in each ring slot (0 rows: no more work)
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kQT + tid;
  uint32_t q[8];
  load_query(q, qs, i, nq);
  uint32_t k1 = kKeyInit, k2 = kKeyInit;
  constexpr int kRows = kDynTiles * kTT;
  // The tile ring never drains between blocks: the producer (thread 0) takes the next block from the counter as soon as the
  // current one has no tile left to request, so the stream of tiles is continuous until the counter runs out.
  int curBlk = -1, tileInBlk = kDynTiles;
  bool done = false;
  auto produce = [&](int slot) {                      // thread 0 only
    for (;;) {
      if (done) { s_cnt[slot] = 0; return; }
      if (tileInBlk == kDynTiles) {
        curBlk = atomicAdd(&counters[blockIdx.x], 1);
        tileInBlk = 0;
        if (curBlk >= nblocks) { done = true; continue; }
      }
      const int r0 = curBlk * kRows + tileInBlk * kTT, c = min(kTT, nt - r0);
      if (c <= 0) { tileInBlk = kDynTiles; continue; }   // ragged last block
      ++tileInBlk;
      s_base[slot] = r0; s_cnt[slot] = c;
      mbar_expect_tx(&full[slot], (uint32_t)c * 32u);
      bulk_g2s(tile + (size_t)slot * kTT * 2, train + (size_t)r0 * 32, (uint32_t)c * 32u, &full[slot]);
      return;
    }
  };
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < kStages; ++s) produce(s);
  }
  __syncthreads();
  for (int t = 0;; ++t) {
    const int s = t % kStages;
    const int cnt = s_cnt[s];
    if (cnt == 0) break;                               // uniform: the slots are consumed in the order they were filled
    mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
    const uint4* tl = tile + (size_t)s * kTT * 2;
    const int base = s_base[s];
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const uint4 a = tl[2 * j], b = tl[2 * j + 1];
      const uint32_t d = (uint32_t)hamming8(q, a, b);
      top2_update(k1, k2, (d << 23) | (uint32_t)(base + j));
    }
    __syncthreads();                                   // everyone is done reading slot s (tile and its two words)
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
      produce(s);
    }
  }
  if (i >= nq) return;
  const int bd1 = (int)(k1 >> 23), bd2 = (int)(k2 >> 23);
  int32_t* p = part + ((size_t)blockIdx.y * nq + i) * 3;
  p[0] = bd1 < 256 ? (int)(k1 & 0x7FFFFFu) : -1; p[1] = bd1; p[2] = bd2;
}

// Merge per-split / per-shard partial results: best = smallest d1, the lowest row index wins ties (parts usually come in
// ascending ranges, but the dynamic scan's do not); second = 2nd smallest of the multiset {d1_s, d2_s}.
__global__ void knn2_merge_kernel(const int32_t* __restrict__ pidx, const int32_t* __restrict__ pd1,
                                  const int32_t* __restrict__ pd2, int stride, int nparts, int nq, int th, float ratio,
                                  int32_t* idx, int32_t* d1, int32_t* d2, uint8_t* ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int b1 = 256, b2 = 256, bi = -1;
  for (int s = 0; s < nparts; ++s) {
    const size_t o = ((size_t)s * nq + i) * stride;
    const int a1 = pd1[o], a2 = pd2[o], ai = pidx[o];
    if (a1 < b1 || (a1 == b1 && a1 < 256 && ai < bi)) { b2 = min(b1, a2); b1 = a1; bi = ai; }
    else { b2 = min(b2, a1); }
  }
  idx[i] = bi; d1[i] = b1; d2[i] = b2;
  ok[i] = accept(b1, b2, th, ratio);
}

// Frame-to-frame matching over extractor output: pair p = (query frame qf[p], train frame tf[p]).
__global__ void __launch_bounds__(kQT, ORBX_HAMM_MINB) knn2_pairs_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ counts,
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

// ---- opt-in variant: the distance matrix on the integer tensor pipe (ORBX_HAMM_MMA=1; default stays the POPC kernel) ----------
// BASELINE's north_star keeps the Hamming stage off the tensor cores; the round-1 review asked for an A/B of exactly that
// alternative, so it is built as a switchable variant with identical results.  Descriptor bits become +-1 bytes (bit 1 -> +1,
// bit 0 -> -1), so dot(a, b) = 256 - 2 * hamming(a, b) is an exact integer and the packed key is ONE multiply-add of the
// accumulator:  key = hamming << 23 | col = dot * (-2^22) + (2^30 + col).  On sm_100a ptxas lowers the one-bit mma.sync
// (xor.popc / and.popc) to this same u8/s8 IMMA.16832 plus per-instruction LOP3 bit-plane masks, so expanding once per
// tile in shared memory is the cheaper route to the same pipe (measured IMMA peak: tools/probes/imma_peak.cu).
//   CTA = 128 queries (8 warps x 16 rows), A fragments for all 256 bits live in 32 registers per thread; train tiles of 64
//   descriptors are expanded through a 256-entry byte -> 8-byte table into a double-buffered shared-memory tile (row pitch
//   272 B: ldmatrix rows land in distinct banks); per 16 x 8 output tile 4 ldmatrix.x4 + 8 IMMA + the top-2 update of the four
//   accumulators each thread owns; the 4 lanes that share a query row merge their candidates at the very end.
#ifndef ORBX_MMA_UNROLL
#define ORBX_MMA_UNROLL 2
#endif
constexpr int kMmaUnroll = ORBX_MMA_UNROLL;
constexpr int kMQ = 256, kMT = 64, kMRow = 272;
// the query tile (kMQ rows) is dead once the A fragments are in registers: the two train buffers live in the same bytes
constexpr size_t kMmaSmem = 2048 + (size_t)kMQ * kMRow;
static_assert(2 * kMT <= kMQ, "train double buffer aliases the query tile");

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void imma_s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 8 descriptor bytes (two words) -> 64 expanded bytes at dst (8-byte aligned)
__device__ __forceinline__ void expand8(const uint2* lut, uint32_t lo, uint32_t hi, uint8_t* dst) {
  uint2* d = reinterpret_cast<uint2*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = lut[(lo >> (8 * i)) & 0xFFu];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[4 + i] = lut[(hi >> (8 * i)) & 0xFFu];
}

// Every warp owns 32 query rows (two m16 tiles): a B fragment fetched from shared memory feeds two IMMAs, which halves the
// ldmatrix traffic per MMA -- with 16 rows per warp the shared-memory data pipe (85 % busy) bound the kernel, not the tensor
// pipe (50 %) (profiles/r2_hamming_mma.md).
// mma_scan: top-2 of the CTA's kMQ query rows (rows q0.. of qsrc, valid below nq) over train rows [0, nt) of tsrc (nt < 2^23);
// a1/a2[i] = packed keys (distance << 23 | local train index) of row  q0 + warp*32 + (i>>1)*16 + lane/4 + 8*(i&1),  merged over
// the 4 lanes of the row (valid in every lane).
__device__ __forceinline__ void mma_scan(const uint8_t* __restrict__ qsrc, int nq, int q0, const uint8_t* __restrict__ tsrc, int nt,
                                         uint8_t* msm, uint32_t (&a1)[4], uint32_t (&a2)[4]) {
  uint2* lut = reinterpret_cast<uint2*>(msm);
  uint8_t* qa = msm + 2048;
  uint8_t* tb = qa;                                         // reused after the A fragments are loaded
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {  // byte -> 8 bytes: bit j set -> 0x01, clear -> 0xFF
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j >> 2] |= (((uint32_t)tid >> j) & 1u ? 0x01u : 0xFFu) << (8 * (j & 3));
    lut[tid] = make_uint2(w[0], w[1]);
  }
  __syncthreads();
  {  // queries: 256 rows x 32 B, one descriptor per thread
    uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
    if (q0 + tid < nq) {
      const uint4* src = reinterpret_cast<const uint4*>(qsrc + (size_t)(q0 + tid) * 32);
      v0 = __ldg(src); v1 = __ldg(src + 1);
    }
    uint8_t* dst = qa + (size_t)tid * kMRow;
    expand8(lut, v0.x, v0.y, dst); expand8(lut, v0.z, v0.w, dst + 64);
    expand8(lut, v1.x, v1.y, dst + 128); expand8(lut, v1.z, v1.w, dst + 192);
  }
  __syncthreads();
  uint32_t A[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const uint8_t* base = qa + (size_t)(warp * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kMRow + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldsm_x4(A[mt][ks], base + ks * 32);
  }
  __syncthreads();                                          // every warp holds its fragments: the tile becomes the train double buffer
  const int ntiles = (nt + kMT - 1) / kMT;
  // train tile: 64 rows x 32 B = 256 threads x 8 B; thread = (row, quarter)
  const int tr = tid >> 2, tq = tid & 3;
  auto fetch = [&](int t) {
    const int row = t * kMT + tr;
    return row < nt ? __ldg(reinterpret_cast<const uint2*>(tsrc + (size_t)row * 32) + tq) : make_uint2(0u, 0u);
  };
  if (ntiles > 0) { const uint2 v = fetch(0); expand8(lut, v.x, v.y, tb + (size_t)tr * kMRow + tq * 64); }
  __syncthreads();
  uint32_t k1[2][2], k2[2][2];                              // [m tile][row lane/4 | lane/4 + 8]
#pragma unroll
  for (int i = 0; i < 4; ++i) { k1[i >> 1][i & 1] = kKeyInit; k2[i >> 1][i & 1] = kKeyInit; }
  const uint32_t colLane = 0x40000000u + (uint32_t)((lane & 3) * 2);
  for (int t = 0; t < ntiles; ++t) {
    const bool more = t + 1 < ntiles;
    uint2 nxt = make_uint2(0u, 0u);
    if (more) nxt = fetch(t + 1);                           // in flight under this tile's MMAs
    const uint8_t* tile = tb + (size_t)(t & 1) * kMT * kMRow;
    const bool tail = (t + 1) * kMT > nt;                   // only the last tile can hold columns >= nt
#pragma unroll kMmaUnroll
    for (int n8 = 0; n8 < kMT / 8; ++n8) {
      const uint8_t* brow = tile + (size_t)(n8 * 8 + (lane & 7)) * kMRow + (lane >> 3) * 16;
      uint32_t Bf[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x) ldsm_x4(Bf[x], brow + x * 64);
      int c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        imma_s8(c[0], A[0][ks], Bf[ks >> 1][(ks & 1) * 2], Bf[ks >> 1][(ks & 1) * 2 + 1]);
        imma_s8(c[1], A[1][ks], Bf[ks >> 1][(ks & 1) * 2], Bf[ks >> 1][(ks & 1) * 2 + 1]);
      }
      const uint32_t col = (uint32_t)(t * kMT + n8 * 8) + colLane;      // 2^30 + column of c[.][0] / c[.][2]
      uint32_t kill0 = 0u, kill1 = 0u;
      if (tail) {
        const int cc = t * kMT + n8 * 8 + (lane & 3) * 2;
        kill0 = cc >= nt ? 0xFFFFFFFFu : 0u; kill1 = cc + 1 >= nt ? 0xFFFFFFFFu : 0u;
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t key0 = ((uint32_t)c[mt][0] * 0xFFC00000u + col) | kill0, key1 = ((uint32_t)c[mt][1] * 0xFFC00000u + col + 1u) | kill1;
        const uint32_t key2 = ((uint32_t)c[mt][2] * 0xFFC00000u + col) | kill0, key3 = ((uint32_t)c[mt][3] * 0xFFC00000u + col + 1u) | kill1;
        top2_update(k1[mt][0], k2[mt][0], key0); top2_update(k1[mt][0], k2[mt][0], key1);
        top2_update(k1[mt][1], k2[mt][1], key2); top2_update(k1[mt][1], k2[mt][1], key3);
      }
    }
    if (more) expand8(lut, nxt.x, nxt.y, tb + (size_t)((t + 1) & 1) * kMT * kMRow + (size_t)tr * kMRow + tq * 64);
    __syncthreads();
  }
  // the 4 lanes of a row hold disjoint column subsets: two smallest keys of the union
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t x1 = k1[i >> 1][i & 1], x2 = k2[i >> 1][i & 1];
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const uint32_t b1 = __shfl_xor_sync(0xffffffffu, x1, o), b2 = __shfl_xor_sync(0xffffffffu, x2, o);
      x2 = min(max(x1, b1), min(x2, b2)); x1 = min(x1, b1);
    }
    a1[i] = x1; a2[i] = x2;
  }
}

__global__ void __launch_bounds__(256, 2) knn2_pairs_mma_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ counts, int cap,
                                                                const int32_t* __restrict__ qf, const int32_t* __restrict__ tf, int th,
                                                                float ratio, int32_t* idx, int32_t* d1, int32_t* d2, uint8_t* ok) {
  extern __shared__ __align__(128) uint8_t msm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.y;
  const int fq = qf[p], ft = tf[p];
  const int nq = min(counts[fq], cap), nt = min(counts[ft], cap);
  const int q0 = blockIdx.x * kMQ;
  if (q0 >= nq) return;
  uint32_t a1[4], a2[4];
  mma_scan(desc + (size_t)fq * cap * 32, nq, q0, desc + (size_t)ft * cap * 32, nt, msm, a1, a2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + warp * 32 + (i >> 1) * 16 + (lane >> 2) + 8 * (i & 1);
    if ((lane & 3) == 0 && r < nq) {
      const int bd1 = (int)min(a1[i] >> 23, 256u), bd2 = (int)min(a2[i] >> 23, 256u);
      const size_t o = (size_t)p * cap + r;
      idx[o] = bd1 < 256 ? (int)(a1[i] & 0x7FFFFFu) : -1;
      d1[o] = bd1; d2[o] = bd2;
      ok[o] = accept(bd1, bd2, th, ratio);
    }
  }
}

// Same scan for one query set against a long train set (map-scale search): grid (ceil(nq/kMQ), nsplit) like knn2_kernel, same
// outputs (final when there is one split, per-split partials (local idx + split offset, d1, d2) otherwise).
__global__ void __launch_bounds__(256, 2) knn2_mma_kernel(const uint8_t* __restrict__ qs, int nq, const uint8_t* __restrict__ train, long long nt,
                                                          long long per, int th, float ratio, int32_t* idx, int32_t* d1, int32_t* d2, uint8_t* ok,
                                                          int32_t* part) {
  extern __shared__ __align__(128) uint8_t msm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const long long t0 = (long long)s * per;
  const int cnt = (int)min(per, nt - t0);
  const int q0 = blockIdx.x * kMQ;
  uint32_t a1[4], a2[4];
  mma_scan(qs, nq, q0, train + (size_t)t0 * 32, cnt, msm, a1, a2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + warp * 32 + (i >> 1) * 16 + (lane >> 2) + 8 * (i & 1);
    if ((lane & 3) == 0 && r < nq) {
      const int bd1 = (int)min(a1[i] >> 23, 256u), bd2 = (int)min(a2[i] >> 23, 256u);
      const int bi = bd1 < 256 ? (int)(t0 + (long long)(a1[i] & 0x7FFFFFu)) : -1;
      if (gridDim.y == 1) {
        idx[r] = bi; d1[r] = bd1; d2[r] = bd2;
        ok[r] = accept(bd1, bd2, th, ratio);
      } else {
        int32_t* pp = part + ((size_t)s * nq + r) * 3;
        pp[0] = bi; pp[1] = bd1; pp[2] = bd2;
      }
    }
  }
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
      if (a1 < b1 || (a1 == b1 && a1 < 256 && ai < bi)) { b2 = min(b1, a2); b1 = a1; bi = ai; }   // ties: lowest row (dynamic scan parts are unordered)
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
  static const int perSm = [] { const char* e = getenv("ORBX_HAMM_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();   // A/B knob
  long long want = ((long long)sms * perSm + qtiles - 1) / qtiles;
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

// which kernels serve the Hamming entry points: 0 = POPC (default, BASELINE's stated design), 1 = the integer-tensor-pipe
// variant above (same results).  Initial value from ORBX_HAMM_MMA, changed at run time by hamm_set_variant().
static std::atomic<int> g_variant{-1};
static bool use_mma() {
  int v = g_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    v = (getenv("ORBX_HAMM_MMA") && atoi(getenv("ORBX_HAMM_MMA")) != 0) ? 1 : 0;
    g_variant.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}
static int mma_prepare() {
  static thread_local int attrDev = -1;
  int dev = 0;
  ORBX_CUDA(cudaGetDevice(&dev));
  if (attrDev != dev) {
    ORBX_CUDA(cudaFuncSetAttribute(knn2_pairs_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    ORBX_CUDA(cudaFuncSetAttribute(knn2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    attrDev = dev;
  }
  return ORBX_OK;
}
// the local-shard scan of hamm_knn2_device / hamm_knn2_sharded*: grid (query tiles, splits)
static int launch_knn2(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, long long per, int ns, int th, float ratio,
                       int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, int32_t* part, cudaStream_t st) {
  if (use_mma()) {
    if (int rc = mma_prepare()) return rc;
    knn2_mma_kernel<<<dim3((nq + kMQ - 1) / kMQ, ns), 256, kMmaSmem, st>>>(d_q, nq, d_t, nt, per, th, ratio, d_idx, d_d1, d_d2, d_ok, part);
  } else {
    knn2_kernel<<<dim3((nq + kQT - 1) / kQT, ns), kQT, 0, st>>>(d_q, nq, d_t, nt, per, th, ratio, d_idx, d_d1, d_d2, d_ok, part);
  }
  return ORBX_OK;
}

}  // namespace orbx

using namespace orbx;

extern "C" {

long long hamm_launch_count(void) { return g_hamm_launches.load(); }

int hamm_set_variant(int variant) {
  const int prev = use_mma() ? 1 : 0;
  if (variant == 0 || variant == 1) g_variant.store(variant, std::memory_order_relaxed);
  return prev;
}

// dynamic block hand-out (knn2_dyn_kernel) when the scan is long enough to matter: >= 8 blocks per worker CTA, rows < 2^23
static std::atomic<int> g_dynamic{-1};      // 0 = static splits only, 1 = automatic (default), 2 = whenever there is more than one split
static bool use_dynamic(int nq, long long nt, int ns) {
  int mode = g_dynamic.load(std::memory_order_relaxed);
  if (mode < 0) {
    mode = getenv("ORBX_HAMM_DYNAMIC") ? atoi(getenv("ORBX_HAMM_DYNAMIC")) : 1;
    g_dynamic.store(mode, std::memory_order_relaxed);
  }
  if (mode <= 0 || use_mma() || ns <= 1 || nt >= (1LL << 23)) return false;
  const long long nblocks = (nt + (long long)kDynTiles * kTT - 1) / ((long long)kDynTiles * kTT);
  return mode >= 2 ? nblocks >= 1 : nblocks >= 8LL * ns;
}
int hamm_set_dynamic(int mode) {
  const int prev = g_dynamic.load(std::memory_order_relaxed);
  if (mode >= 0 && mode <= 2) g_dynamic.store(mode, std::memory_order_relaxed);
  return prev < 0 ? 1 : prev;
}

constexpr size_t kDynCounterBytes = 256;          // one counter per query tile (<= 64 tiles = 8192 queries in dynamic mode)

size_t hamm_knn2_workspace_bytes(int nq, long long nt) {
  long long per;
  const int ns = pick_splits(nq, nt, &per);
  return ns > 1 ? (size_t)ns * nq * 3 * sizeof(int32_t) + kDynCounterBytes : 0;
}

// The local scan of both entry points: per-split partials into the workspace (ns > 1; dynamic block hand-out for long scans) or
// the final result into the output arrays (ns == 1).
static int scan_local(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, long long per, int ns, int th, float ratio,
                      int32_t* d_idx, int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, void* d_workspace, cudaStream_t st) {
  const int qtiles = (nq + kQT - 1) / kQT;
  if (use_dynamic(nq, nt, ns) && (size_t)qtiles * sizeof(int) <= kDynCounterBytes) {
    int* counters = (int*)((uint8_t*)d_workspace + (size_t)ns * nq * 3 * sizeof(int32_t));
    ORBX_CUDA(cudaMemsetAsync(counters, 0, kDynCounterBytes, st));
    const int nblocks = (int)((nt + (long long)kDynTiles * kTT - 1) / ((long long)kDynTiles * kTT));
    knn2_dyn_kernel<<<dim3(qtiles, ns), kQT, 0, st>>>(d_q, nq, d_t, (int)nt, nblocks, counters, (int32_t*)d_workspace);
    return ORBX_OK;
  }
  return launch_knn2(d_q, nq, d_t, nt, per, ns, th, ratio, d_idx, d_d1, d_d2, d_ok, (int32_t*)d_workspace, st);
}

int hamm_knn2_device(const uint8_t* d_q, int nq, const uint8_t* d_t, long long nt, int th, float ratio, int32_t* d_idx,
                     int32_t* d_d1, int32_t* d_d2, uint8_t* d_ok, void* d_workspace, size_t workspace_bytes, void* stream) {
  // an empty train set (nt == 0: a trailing rank of a small sharded map) is legal and may come with a null pointer
  if (!d_q || (!d_t && nt > 0) || !d_idx || !d_d1 || !d_d2 || !d_ok || nq < 0 || nt < 0) { set_error("bad argument"); return ORBX_ERR_ARG; }
  if (nq == 0) return ORBX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long per;
  const int ns = pick_splits(nq, nt, &per);
  if (ns > 1 && (!d_workspace || workspace_bytes < (size_t)ns * nq * 3 * sizeof(int32_t) + kDynCounterBytes)) {
    set_error("workspace too small (see hamm_knn2_workspace_bytes)");
    return ORBX_ERR_CAPACITY;
  }
  if (int rc = scan_local(d_q, nq, d_t, nt, per, ns, th, ratio, d_idx, d_d1, d_d2, d_ok, d_workspace, st)) return rc;
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
  nvtxRangePushA("orbx:hamming_pairs");
  if (use_mma()) {
    if (int rc = mma_prepare()) { nvtxRangePop(); return rc; }
    dim3 grid((cap + kMQ - 1) / kMQ, npairs);
    knn2_pairs_mma_kernel<<<grid, 256, kMmaSmem, (cudaStream_t)stream>>>(d_desc, d_counts, cap, d_qf, d_tf, th, ratio, d_idx, d_d1, d_d2, d_ok);
  } else {
    dim3 grid((cap + kQT - 1) / kQT, npairs);
    knn2_pairs_kernel<<<grid, kQT, 0, (cudaStream_t)stream>>>(d_desc, d_counts, cap, d_qf, d_tf, th, ratio, d_idx, d_d1, d_d2, d_ok);
  }
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
  if (ns > 1 && (!d_workspace || workspace_bytes < (size_t)ns * nq * 3 * sizeof(int32_t) + kDynCounterBytes)) {
    set_error("workspace too small (see hamm_knn2_workspace_bytes)");
    return ORBX_ERR_CAPACITY;
  }
  const int parity = epoch & 1;
  if (phases & 1) {
    // local shard: per-split partials (ns > 1) or the shard result in the output arrays (ns == 1), local row indices
    if (int rc = scan_local(d_q, nq, d_t, nt, per, ns, th, ratio, d_idx, d_d1, d_d2, d_ok, d_workspace, st)) return rc;
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
