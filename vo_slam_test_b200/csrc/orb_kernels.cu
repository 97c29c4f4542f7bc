// orb_kernels.cu -- sm_100a kernels of the ORB extractor (pyramid, FAST, quadtree, blur, orientation+rBRIEF).
//
// Built with -fmad=false: the reference binary has no FMA contraction (CMakeLists.txt:4-5), and the float
// paths below (fastAtan2, descriptor rotation, keypoint scaling) must round exactly like it.
#include <stdlib.h>

#include "orb_extract.cuh"

// grid order switches (A/B builds): 1 = frames are the fast grid dimension
#ifndef ORBX_FAST_FF
#define ORBX_FAST_FF 1
#endif
#ifndef ORBX_OCT_FF
#define ORBX_OCT_FF 1
#endif
#ifndef ORBX_OD_FF
#define ORBX_OD_FF 0
#endif

namespace orbx {

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {   // a * b + c as one IMAD (FMA pipe), never LEA/SHF
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

__device__ __forceinline__ const uint8_t* level_ptr(const Geom& G, const Bufs& B, int l, int f, int& pitch) {
  if (l == 0) {
    pitch = (int)B.rowStride0;
    return B.img0 + (size_t)f * B.frameStride0;
  }
  const LevelGeom& L = G.L[l];
  pitch = L.pitch;
  return B.pyr + L.pyrOff + (size_t)f * L.h * L.pitch;
}

// ======================================================================================================
// K1  pyramid level l from level l-1: cv::resize INTER_LINEAR, 11-bit fixed point (ORBextractor.cpp:1129;
// arithmetic: SURVEY App. A.1).  One thread = 4 horizontally adjacent destination pixels, one 32-bit store.
// ======================================================================================================
__global__ void __launch_bounds__(256) resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                     size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                     int dpitch, size_t dframe, ResizeTaps T) {
  pdl_prologue();
  const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int dy = blockIdx.y * 8 + threadIdx.y;
  if (dx0 >= dw || dy >= dh) return;
  const int f = blockIdx.z;
  const int sy0 = __ldg(T.yofs + dy);
  const int sy1 = min(sy0 + 1, sh - 1);
  const int b0 = __ldg(T.yb0 + dy), b1 = __ldg(T.yb1 + dy);
  const uint8_t* r0 = src + f * sframe + (size_t)sy0 * spitch;
  const uint8_t* r1 = src + f * sframe + (size_t)sy1 * spitch;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dx = min(dx0 + i, dw - 1);
    const int sx = __ldg(T.xofs + dx);
    const int sx1 = min(sx + 1, sw - 1);
    const int a0 = __ldg(T.xa0 + dx), a1 = __ldg(T.xa1 + dx);
    const int h0 = __ldg(r0 + sx) * a0 + __ldg(r0 + sx1) * a1;
    const int h1 = __ldg(r1 + sx) * a0 + __ldg(r1 + sx1) * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    out |= (uint32_t)v << (8 * i);
  }
  uint8_t* d = dst + f * dframe + (size_t)dy * dpitch + dx0;
  if (dx0 + 3 < dw) {
    *reinterpret_cast<uint32_t*>(d) = out;   // dpitch and dx0 are multiples of 4, dst base is 256-B aligned
  } else {
    for (int i = 0; dx0 + i < dw; ++i) d[i] = (uint8_t)(out >> (8 * i));
  }
}

// Word-granular variant (the one normally used): one thread = 4 destination pixels; per source row three aligned
// 32-bit loads + two funnel shifts give the 8 bytes starting at the quad's first tap, PRMT picks each pixel's byte
// pair and one IDP (dp2a) applies the two 11-bit weights.  Needs 4-byte aligned source rows and scale <= 2.
__global__ void __launch_bounds__(256) resize4_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                      size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                      int dpitch, size_t dframe, ResizeTaps T) {
  pdl_prologue();
  const int q = blockIdx.x * 32 + threadIdx.x;
  const int dy = blockIdx.y * 8 + threadIdx.y;
  const int dx0 = q * 4;
  if (dx0 >= dw || dy >= dh) return;
  const int f = blockIdx.z;
  const int4 Q = __ldg(T.quad + q);
  const uint4 Wt = __ldg(T.xw + q);
  const int sy0 = __ldg(T.yofs + dy);
  const int sy1 = min(sy0 + 1, sh - 1);
  const int b0 = __ldg(T.yb0 + dy), b1 = __ldg(T.yb1 + dy);
  const int lastw = (sw - 1) & ~3;                      // words past it hold no tap with a non-zero weight
  const int o0 = min(Q.x, lastw), o1 = min(Q.x + 4, lastw), o2 = min(Q.x + 8, lastw);
  const uint8_t* r0 = src + f * sframe + (size_t)sy0 * spitch;
  const uint8_t* r1 = src + f * sframe + (size_t)sy1 * spitch;
  const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(r0 + o0)), a1 = __ldg(reinterpret_cast<const uint32_t*>(r0 + o1)),
                 a2 = __ldg(reinterpret_cast<const uint32_t*>(r0 + o2));
  const uint32_t c0 = __ldg(reinterpret_cast<const uint32_t*>(r1 + o0)), c1 = __ldg(reinterpret_cast<const uint32_t*>(r1 + o1)),
                 c2 = __ldg(reinterpret_cast<const uint32_t*>(r1 + o2));
  const uint32_t lo0 = __funnelshift_r(a0, a1, Q.y), hi0 = __funnelshift_r(a1, a2, Q.y);
  const uint32_t lo1 = __funnelshift_r(c0, c1, Q.y), hi1 = __funnelshift_r(c1, c2, Q.y);
  const uint32_t wts[4] = {Wt.x, Wt.y, Wt.z, Wt.w};
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t sel = ((uint32_t)Q.z >> (8 * i)) & 0xFFu;
    const int h0 = (int)__dp2a_lo(wts[i], __byte_perm(lo0, hi0, sel), 0u);
    const int h1 = (int)__dp2a_lo(wts[i], __byte_perm(lo1, hi1, sel), 0u);
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    out |= (uint32_t)v << (8 * i);
  }
  uint8_t* d = dst + f * dframe + (size_t)dy * dpitch + dx0;
  if (dx0 + 3 < dw) {
    *reinterpret_cast<uint32_t*>(d) = out;
  } else {
    for (int i = 0; dx0 + i < dw; ++i) d[i] = (uint8_t)(out >> (8 * i));
  }
}

// Register-blocked variant (the default when resize4_kernel's preconditions hold): one thread owns 4 adjacent destination
// columns and walks down kRwRows destination rows.  The per-quad tap tables are loaded once per thread, and the
// horizontal interpolation of a source row is reused by the next destination row whenever its first tap row is the
// previous row's second one (4 rows out of 5 at scale 1.2): 1.25 instead of 2 horizontal passes per destination row.
// A warp = 32 quads of one row segment, so the reuse test is warp-uniform.  Same integer arithmetic.

__global__ void __launch_bounds__(128) resize_walk_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                          size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                          int dpitch, size_t dframe, ResizeTaps T) {
  pdl_prologue();
  const int q = blockIdx.x * 32 + threadIdx.x;
  const int dy0 = (blockIdx.y * 4 + threadIdx.y) * kRwRows;
  const int dx0 = q * 4;
  if (dx0 >= dw || dy0 >= dh) return;
  const int f = blockIdx.z;
  const int4 Q = __ldg(T.quad + q);
  const uint4 Wt = __ldg(T.xw + q);
  const int lastw = (sw - 1) & ~3;                      // words past it hold no tap with a non-zero weight
  const uint8_t* p0 = src + f * sframe + min(Q.x, lastw);
  const int d1 = min(Q.x + 4, lastw) - min(Q.x, lastw), d2 = min(Q.x + 8, lastw) - min(Q.x, lastw);
  const uint32_t sel0 = (uint32_t)Q.z & 0xFFu, sel1 = ((uint32_t)Q.z >> 8) & 0xFFu, sel2 = ((uint32_t)Q.z >> 16) & 0xFFu,
                 sel3 = ((uint32_t)Q.z >> 24) & 0xFFu;
  auto hrow = [&](int sy, int (&h)[4]) {                 // (a0*p[sx] + a1*p[sx+1]) >> 4 for the four columns
    const uint8_t* r = p0 + (uint32_t)(sy * spitch);     // one level of one frame is far below 4 GiB
    const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(r)), a1 = __ldg(reinterpret_cast<const uint32_t*>(r + d1)),
                   a2 = __ldg(reinterpret_cast<const uint32_t*>(r + d2));
    const uint32_t lo = __funnelshift_r(a0, a1, Q.y), hi = __funnelshift_r(a1, a2, Q.y);
    h[0] = (int)(__dp2a_lo(Wt.x, __byte_perm(lo, hi, sel0), 0u) >> 4);
    h[1] = (int)(__dp2a_lo(Wt.y, __byte_perm(lo, hi, sel1), 0u) >> 4);
    h[2] = (int)(__dp2a_lo(Wt.z, __byte_perm(lo, hi, sel2), 0u) >> 4);
    h[3] = (int)(__dp2a_lo(Wt.w, __byte_perm(lo, hi, sel3), 0u) >> 4);
  };
  uint8_t* d = dst + f * dframe + (size_t)dy0 * dpitch + dx0;
  const bool whole = dx0 + 3 < dw;
  const int dyEnd = min(dy0 + kRwRows, dh);
  int h0[4], h1[4];
  int r1 = -1;                                          // source row held in h1
  for (int dy = dy0; dy < dyEnd; ++dy) {
    const int sy0 = __ldg(T.yofs + dy);
    const int sy1 = min(sy0 + 1, sh - 1);
    const int b0 = __ldg(T.yb0 + dy), b1 = __ldg(T.yb1 + dy);
    if (sy0 == r1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) h0[i] = h1[i];
    } else {
      hrow(sy0, h0);
    }
    if (sy1 != sy0) {
      hrow(sy1, h1);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) h1[i] = h0[i];
    }
    r1 = sy1;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (uint32_t)((((b0 * h0[i]) >> 16) + ((b1 * h1[i]) >> 16) + 2) >> 2);
    const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    if (whole) {
      *reinterpret_cast<uint32_t*>(d) = out;
    } else {
      for (int i = 0; dx0 + i < dw; ++i) d[i] = (uint8_t)(out >> (8 * i));
    }
    d += dpitch;
  }
}

// TMA-staged variant (the default when the source level can be described): the 128 x 64 destination tile of a CTA needs
// a source box of about 170 x 80 bytes, fetched by ONE cp.async.bulk.tensor (UTMALDG) into shared memory while the
// threads load their tap tables; the walk then reads its three words per source row with LDS, so the dependent
// global-load latency that bounded resize_walk_kernel (long-scoreboard stalls 9 per issue) is paid once per CTA and
// overlapped by the other resident CTAs.  Row taps of the 16 rows of a warp live in lanes 0..15 and are broadcast by
// shuffles.  Same integer arithmetic as resize_walk_kernel.
// One destination tile (tx, ty) of frame f: shared by the per-level kernel and the fused pyramid kernel.  `bar` is a fresh
// (uninitialised) mbarrier of this CTA.  No early exit: every thread returns from here (the fused kernel signals afterwards).
__device__ __forceinline__ void resize_tma_tile(const CUtensorMap* __restrict__ map, int z, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                                                int dpitch, size_t dframe, const ResizeTaps& T, const ResizeTma& R, int tx, int ty, int f,
                                                uint8_t* rz_tile, uint64_t* bar) {
  const int lane = threadIdx.x, wy = threadIdx.y;
  const int tx0 = R.x0[tx], ty0 = R.y0[ty];
  const uint32_t barA = (uint32_t)__cvta_generic_to_shared(bar);
  if (lane == 0 && wy == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"((uint32_t)(R.boxW * R.boxH)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(rz_tile)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(tx0), "r"(ty0), "r"(z), "r"(barA)
        : "memory");
  }
  const int q = tx * (kRzTileW / 4) + lane;
  const int dy0 = (ty * 4 + wy) * R.rows;
  const int dx0 = q * 4;
  const bool live = dx0 < dw && dy0 < dh;
  int4 Q = make_int4(0, 0, 0, 0);
  uint4 Wt = make_uint4(0, 0, 0, 0);
  if (live) { Q = __ldg(T.quad + q); Wt = __ldg(T.xw + q); }
  // lane i < 16 holds the taps of destination row dy0 + i
  int tsy = 0, tb = 0;
  if (lane < R.rows && dy0 + lane < dh) {
    tsy = __ldg(T.yofs + dy0 + lane);
    tb = (int)(uint16_t)__ldg(T.yb0 + dy0 + lane) | ((int)(uint16_t)__ldg(T.yb1 + dy0 + lane) << 16);
  }
  __syncthreads();                               // barrier initialised before anyone polls it
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "RW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra RD_%=;\n\t"
      "bra RW_%=;\n\t"
      "RD_%=:\n\t}" ::"r"(barA) : "memory");
  if (dy0 >= dh) return;                         // warp-uniform
  const uint32_t sel0 = (uint32_t)Q.z & 0xFFu, sel1 = ((uint32_t)Q.z >> 8) & 0xFFu, sel2 = ((uint32_t)Q.z >> 16) & 0xFFu,
                 sel3 = ((uint32_t)Q.z >> 24) & 0xFFu;
  const uint8_t* p0 = rz_tile + (live ? Q.x - tx0 : 0);
  const int bw = R.boxW;
  auto hrow = [&](int sy, int (&h)[4]) {         // (a0*p[sx] + a1*p[sx+1]) >> 4 for the four columns
    const uint32_t* r = reinterpret_cast<const uint32_t*>(p0 + (sy - ty0) * bw);
    const uint32_t a0 = r[0], a1 = r[1], a2 = r[2];
    const uint32_t lo = __funnelshift_r(a0, a1, Q.y), hi = __funnelshift_r(a1, a2, Q.y);
    h[0] = (int)(__dp2a_lo(Wt.x, __byte_perm(lo, hi, sel0), 0u) >> 4);
    h[1] = (int)(__dp2a_lo(Wt.y, __byte_perm(lo, hi, sel1), 0u) >> 4);
    h[2] = (int)(__dp2a_lo(Wt.z, __byte_perm(lo, hi, sel2), 0u) >> 4);
    h[3] = (int)(__dp2a_lo(Wt.w, __byte_perm(lo, hi, sel3), 0u) >> 4);
  };
  uint8_t* d = dst + f * dframe + (size_t)dy0 * dpitch + dx0;
  const bool whole = dx0 + 3 < dw;
  const int rows = min(R.rows, dh - dy0);
  int h0[4], h1[4];
  int r1 = -1;                                   // source row held in h1
  for (int i = 0; i < rows; ++i) {
    const int sy0 = __shfl_sync(0xffffffffu, tsy, i);
    const int bb = __shfl_sync(0xffffffffu, tb, i);
    const int b0 = (int)(short)(bb & 0xFFFF), b1 = bb >> 16;
    const int sy1 = min(sy0 + 1, sh - 1);
    if (sy0 == r1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) h0[k] = h1[k];
    } else {
      hrow(sy0, h0);
    }
    if (sy1 != sy0) {
      hrow(sy1, h1);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) h1[k] = h0[k];
    }
    r1 = sy1;
    if (live) {
      uint32_t v[4];
#pragma unroll
      // "+ 2" rides on the second product (a multiple of 2^16 added before the floor shift) and the four results (<= 255)
      // are packed by multiply-adds: both moves take work off the ALU pipe, the bound of this kernel, onto the FMA pipe
      for (int k = 0; k < 4; ++k) v[k] = (uint32_t)((((b0 * h0[k]) >> 16) + ((b1 * h1[k] + 0x20000) >> 16)) >> 2);
      const uint32_t out = mad_u32(mad_u32(mad_u32(v[3], 256u, v[2]), 256u, v[1]), 256u, v[0]);
      if (whole) {
        *reinterpret_cast<uint32_t*>(d) = out;
      } else {
        for (int k = 0; dx0 + k < dw; ++k) d[k] = (uint8_t)(out >> (8 * k));
      }
    }
    d += dpitch;
  }
}

#ifndef ORBX_RZ_MINB
#define ORBX_RZ_MINB 1
#endif
__global__ void __launch_bounds__(128, ORBX_RZ_MINB) resize_tma_kernel(const CUtensorMap* __restrict__ map, int z0, int sh, uint8_t* __restrict__ dst,
                                                         int dw, int dh, int dpitch, size_t dframe, ResizeTaps T,
                                                         const __grid_constant__ ResizeTma R) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t rz_tile[];
  __shared__ __align__(8) uint64_t bar;
  resize_tma_tile(map, z0 + (int)blockIdx.z, sh, dst, dw, dh, dpitch, dframe, T, R, blockIdx.x, blockIdx.y, blockIdx.z, rz_tile, &bar);
}

// ONE launch for the whole pyramid (north_star: "one kernel builds the scale pyramid").  Tickets are handed out in level-major
// order, so a CTA only ever waits for CTAs that hold earlier tickets and are therefore already running: no deadlock whatever the
// number of resident CTAs.  A tile of level l needs source rows [y0, y0 + boxH) of level l-1 = a few destination tile rows of
// that level: it spins (one thread, ld.acquire.gpu) until their completion counters reach the tiles-per-row of level l-1, orders
// the generic-proxy acquire in front of its async-proxy (TMA) read with fence.proxy.async, and after its own stores signals
// its tile row (stores -> __syncthreads -> __threadfence -> atomicAdd).  sync[0] = ticket counter, sync[1 + ((f * nlev + i) *
// kPyrSyncStride) + ty] = finished tiles of tile row ty of level i + 1 of frame f; zeroed by the host before every launch.
__global__ void __launch_bounds__(128) pyramid_fused_kernel(const PyrLevel* __restrict__ plan, const __grid_constant__ PyrLaunch P,
                                                            int* __restrict__ sync) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t rz_tile[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_ticket;
  __shared__ PyrLevel sL;
  const bool t0 = threadIdx.x == 0 && threadIdx.y == 0;
  if (t0) s_ticket = atomicAdd(&sync[0], 1);
  __syncthreads();
  const int t = s_ticket;
  int li = 0;
  while (li + 1 < P.nlev && t >= P.start[li + 1]) ++li;
  {   // this level's record -> shared memory (the taps / tile origins are indexed at run time)
    const int* src = reinterpret_cast<const int*>(plan + li);
    int* dstw = reinterpret_cast<int*>(&sL);
    for (int i = threadIdx.y * 32 + threadIdx.x; i < (int)(sizeof(PyrLevel) / sizeof(int)); i += 128) dstw[i] = __ldg(src + i);
  }
  __syncthreads();
  const PyrLevel& L = sL;
  const int r = t - P.start[li], per = L.ntx * L.nty;
  const int f = r / per, tt = r - f * per, ty = tt / L.ntx, tx = tt - ty * L.ntx;
  int* mine = sync + 1 + ((size_t)f * P.nlev + li) * kPyrSyncStride;
  if (li > 0 && t0) {
    const int y0 = L.R.y0[ty], y1 = min(y0 + L.R.boxH, L.sh) - 1;
    const int* cnt = mine - kPyrSyncStride;
    for (int sy = y0 / L.srcTileH; sy <= y1 / L.srcTileH; ++sy) {
      int v;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt + sy) : "memory");
      } while (v < L.srcNtx);
    }
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  resize_tma_tile(L.map, (li == 0 ? P.z0 : 0) + f, L.sh, L.dst, L.dw, L.dh, L.dpitch, L.dframe, L.T, L.R, tx, ty, f, rz_tile, &bar);
  __threadfence();                               // every thread's stores are visible GPU-wide before the tile row is signalled
  __syncthreads();
  if (t0 && li + 1 < P.nlev) atomicAdd(mine + ty, 1);
}

// ======================================================================================================
// K2  per-cell FAST-9/16 + per-cell 3x3 NMS + iniTh/minTh retry + ordered compaction
// (ORBextractor.cpp:796-836 around cv::FAST; arithmetic: SURVEY App. A.2).
//
// One CTA = one "slot" = up to kCellsPerCta neighbouring cells of one cell row of one level of one frame.
// The corner strength  best(p) = max over the 16 arcs of 9 of max(min d, min -d)  does not depend on the
// threshold: p is a corner at threshold t iff best > t, and its score is best-1.  So best is computed once,
// the two NMS passes (t = iniTh, then t = minTh for cells that came back empty) reuse it.
// ======================================================================================================
__device__ __forceinline__ int fast_best(const uint8_t* p, int pitch, int c) {
  // ring order of cv::FAST, k = 0..15
  int d[16];
  d[0] = c - p[3 * pitch];       d[1] = c - p[3 * pitch + 1];   d[2] = c - p[2 * pitch + 2];
  d[3] = c - p[pitch + 3];       d[4] = c - p[3];               d[5] = c - p[-pitch + 3];
  d[6] = c - p[-2 * pitch + 2];  d[7] = c - p[-3 * pitch + 1];  d[8] = c - p[-3 * pitch];
  d[9] = c - p[-3 * pitch - 1];  d[10] = c - p[-2 * pitch - 2]; d[11] = c - p[-pitch - 3];
  d[12] = c - p[-3];             d[13] = c - p[pitch - 3];      d[14] = c - p[2 * pitch - 2];
  d[15] = c - p[3 * pitch - 1];
  int mn2[16], mx2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
  int mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
  int a = -255, b = 255;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);   // min over ring k..k+8
    const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
    a = max(a, mn9);
    b = min(b, mx9);
  }
  return max(a, -b);
}

// ---- packed-byte helpers for the rejection test: one thread tests 4 horizontally adjacent centres at once ----------
// W[0..3] = 16 bytes of one tile row starting at the 4-byte-aligned address at or below (xi-3); SH = misalignment.
template <int SH, int DX>
__device__ __forceinline__ uint32_t pick4(const uint32_t (&W)[4]) {
  constexpr int off = SH + 3 + DX, wi = off >> 2, sft = (off & 3) * 8;
  if (sft == 0) return W[wi];
  return __funnelshift_r(W[wi], W[wi + 1], sft);
}
__device__ __forceinline__ void load_window(uint32_t (&W)[4], const uint8_t* p) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
  W[0] = q[0]; W[1] = q[1]; W[2] = q[2]; W[3] = q[3];
}
// per-byte |s - c| > t  ->  bit 7 of each byte.  kc = 0x7F7F7F7F - (t & 0x7F) * 0x01010101; hiT = (t >= 128).
// The flags only feed a REJECTION filter (every survivor is re-decided by the exact arc score), so for t < 128 the
// bytes are added without masking bit 7 first: a byte's own flag stays exact (a carry out of it means a >= 129 + t,
// whose bit 7 is set anyway) and a carry INTO a byte can only turn its "a > t" into "a >= t" -- a superset.
__device__ __forceinline__ uint32_t far4(uint32_t s, uint32_t c, uint32_t kc, bool hiT) {
  const uint32_t a = __vabsdiffu4(s, c);
  if (!hiT) return (a | (a + kc)) & 0x80808080u;
  const uint32_t low = (a & 0x7F7F7F7Fu) + kc;          // bit 7 set iff (a & 0x7F) > (t & 0x7F)
  return (a & low) & 0x80808080u;
}
// bit 7 of the four bytes of m -> bits 0..3, and back (one IMAD on the otherwise idle FMA pipe instead of shift/mask chains)
__device__ __forceinline__ uint32_t gather_flags(uint32_t m) { return (((m >> 7) & 0x01010101u) * 0x00204081u) >> 21 & 0xFu; }
__device__ __forceinline__ uint32_t scatter_flags(uint32_t bits) { return (bits * 0x10204080u) & 0x80808080u; }

// Build switch ORBX_FAST_A2 (default 0, A/B only, tools/build_variants.sh): stage a tests only the two axis pairs
// (0|8, 4|12) and the two pairs next to the vertical axis move to stage b -- half the stage-a work on every pixel against
// more survivors in stage b.  Both stages are filters in front of the exact arc score, so the result cannot change.
#ifndef ORBX_FAST_A2
#define ORBX_FAST_A2 0
#endif
// Necessary condition of a FAST-9 corner (same shape as OpenCV's quick test, sign-agnostic): every opposite ring pair
// has a member that differs from the centre by more than t.  Per-byte flags (bit 7) for the 4 centres, in two stages so
// that the second one can run densely over the items that survived the first.
template <int SH>
__device__ __forceinline__ uint32_t reject4_a(const uint8_t* rowm3, int sp, uint32_t kc, bool hiT) {   // rows -3, 0, +3
  uint32_t Wp3[4], Wm3[4], W0[4];
  load_window(Wm3, rowm3);
  load_window(W0, rowm3 + 3 * sp);
  load_window(Wp3, rowm3 + 6 * sp);
  const uint32_t c = pick4<SH, 0>(W0);
  uint32_t m = far4(pick4<SH, 0>(Wp3), c, kc, hiT) | far4(pick4<SH, 0>(Wm3), c, kc, hiT);     // ring 0 | 8
  m &= far4(pick4<SH, 3>(W0), c, kc, hiT) | far4(pick4<SH, -3>(W0), c, kc, hiT);               // ring 4 | 12
#if !ORBX_FAST_A2
  m &= far4(pick4<SH, 1>(Wp3), c, kc, hiT) | far4(pick4<SH, -1>(Wm3), c, kc, hiT);             // ring 1 | 9
  m &= far4(pick4<SH, 1>(Wm3), c, kc, hiT) | far4(pick4<SH, -1>(Wp3), c, kc, hiT);             // ring 7 | 15
#endif
  return m;
}
template <int SH>
__device__ __forceinline__ uint32_t reject4_b(const uint8_t* rowm3, int sp, uint32_t kc, bool hiT, uint32_t m) {   // rows +-1, +-2
  uint32_t W0[4], Wa[4], Wb[4];
  load_window(W0, rowm3 + 3 * sp);
  const uint32_t c = pick4<SH, 0>(W0);
  load_window(Wa, rowm3 + 5 * sp);   // row +2
  load_window(Wb, rowm3 + 1 * sp);   // row -2
  m &= far4(pick4<SH, 2>(Wa), c, kc, hiT) | far4(pick4<SH, -2>(Wb), c, kc, hiT);               // ring 2 | 10
  m &= far4(pick4<SH, 2>(Wb), c, kc, hiT) | far4(pick4<SH, -2>(Wa), c, kc, hiT);               // ring 6 | 14
  load_window(Wa, rowm3 + 4 * sp);   // row +1
  load_window(Wb, rowm3 + 2 * sp);   // row -1
  m &= far4(pick4<SH, 3>(Wa), c, kc, hiT) | far4(pick4<SH, -3>(Wb), c, kc, hiT);               // ring 3 | 11
  m &= far4(pick4<SH, 3>(Wb), c, kc, hiT) | far4(pick4<SH, -3>(Wa), c, kc, hiT);               // ring 5 | 13
#if ORBX_FAST_A2
  load_window(Wa, rowm3 + 6 * sp);   // row +3
  load_window(Wb, rowm3);            // row -3
  m &= far4(pick4<SH, 1>(Wa), c, kc, hiT) | far4(pick4<SH, -1>(Wb), c, kc, hiT);               // ring 1 | 9
  m &= far4(pick4<SH, 1>(Wb), c, kc, hiT) | far4(pick4<SH, -1>(Wa), c, kc, hiT);               // ring 7 | 15
#endif
  return m;
}

__device__ __forceinline__ uint32_t pack_lo_hi(uint32_t lo, uint32_t hi) {   // lo | hi << 16 for lo, hi < 2^16, as one IMAD
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
  return r;
}
// Corner strength of TWO pixels at once in packed signed 16-bit lanes (VIMNMX.S16x2).
__device__ __forceinline__ void fast_best2(const uint8_t* pa, const uint8_t* pb, int sp, int& bestA, int& bestB) {
  // d = centre - ring pixel, both pixels' differences in one register.  The centres carry a bias of 256 per lane so that ONE
  // 32-bit subtraction serves both lanes (the low lane never borrows); min/max commute with the bias, removed at the end.
  const uint32_t cc = ((uint32_t)pa[0] + 256u) | (((uint32_t)pb[0] + 256u) << 16);
  uint32_t d[16];
  // the two ring bytes are packed by a multiply-add (FMA pipe) instead of PRMT: the ALU pipe is the bound of this kernel
#define D2(k, off) d[k] = cc - pack_lo_hi((uint32_t)pa[off], (uint32_t)pb[off])
  D2(0, 3 * sp);       D2(1, 3 * sp + 1);   D2(2, 2 * sp + 2);    D2(3, sp + 3);
  D2(4, 3);            D2(5, -sp + 3);      D2(6, -2 * sp + 2);   D2(7, -3 * sp + 1);
  D2(8, -3 * sp);      D2(9, -3 * sp - 1);  D2(10, -2 * sp - 2);  D2(11, -sp - 3);
  D2(12, -3);          D2(13, sp - 3);      D2(14, 2 * sp - 2);   D2(15, 3 * sp - 1);
#undef D2
  // min / max over every arc of 9 consecutive ring positions as 3 x 3 windows of three-input packed min/max (VIMNMX3):
  // 32 + 32 instructions for the 16 arcs, then a three-input reduction tree over the arcs
  uint32_t mn3[16], mx3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    mx3[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  }
  uint32_t mn9[16], mx9[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn9[k] = __vimin3_s16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);   // min over ring k..k+8
    mx9[k] = __vimax3_s16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
  }
  uint32_t a5[5], b5[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    a5[k] = __vimax3_s16x2(mn9[3 * k], mn9[3 * k + 1], mn9[3 * k + 2]);
    b5[k] = __vimin3_s16x2(mx9[3 * k], mx9[3 * k + 1], mx9[3 * k + 2]);
  }
  const uint32_t a = __vimax3_s16x2(__vimax3_s16x2(a5[0], a5[1], a5[2]), a5[3], __vmaxs2(a5[4], mn9[15]));
  const uint32_t b = __vimin3_s16x2(__vimin3_s16x2(b5[0], b5[1], b5[2]), b5[3], __vmins2(b5[4], mx9[15]));
  bestA = max((int)(a & 0xFFFF) - 256, 256 - (int)(b & 0xFFFF));
  bestB = max((int)(a >> 16) - 256, 256 - (int)(b >> 16));
}

__global__ void __launch_bounds__(kFastThreads) fast_kernel(const __grid_constant__ Geom G, const Bufs B, const __grid_constant__ TmaSet TM) {
  pdl_prologue();
  // Work-efficient layout: (1) tile staged with 16-byte loads, (2) a cheap 16-pixel-ring rejection test over all
  // pixels that pushes the few survivors into a shared-memory queue, (3) the full arc score and the per-cell NMS run
  // densely over that queue only, (4) survivors are ranked by (cell, row, column) to emit them in the reference order.
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int cellCnt[kCellsPerCta];
  __shared__ int qn, sn, in, cn, retryMask;
  __shared__ __align__(8) uint64_t tmaBar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = blockIdx.x + B.slotOff, f = blockIdx.y;
  int l = 0;
  while (l + 1 < G.nlevels && slot >= G.L[l + 1].slot0) ++l;
  const LevelGeom& L = G.L[l];
  const int ls = slot - L.slot0;
  const int ci = ls / L.groups, grp = ls - ci * L.groups;
  const int j0 = grp * kCellsPerCta, j1 = min(j0 + kCellsPerCta, L.nCols);
  const int tx0 = j0 * L.wCell, tx1 = min(j1 * L.wCell + 6, L.regW);
  const int ty0 = ci * L.hCell, ty1 = min(ty0 + L.hCell + 6, L.regH);
  const int tw = tx1 - tx0, th = ty1 - ty0;
  const int iw = tw - 6, ih = th - 6;
  int* out_count = B.slotCount + (size_t)f * G.totalSlots + slot;
  if (iw <= 0 || ih <= 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  const int sp = G.fastTileW;                                  // smem row pitch (multiple of 16)
  uint8_t* img = smem;                                         // [fastTileH][sp]
  uint8_t* sc = smem + G.fastTileH * sp;                       // corner strength (0 = not a corner at the pass threshold)
  uint16_t* queue = (uint16_t*)(sc + G.fastTileH * sp);        // yi<<8 | xi of pixels passing the rejection test
  uint32_t* surv = (uint32_t*)(queue + G.fastTileH * sp);      // cell<<24 | yi<<12 | cx of NMS survivors
  uint16_t* itemq = (uint16_t*)(surv + G.fastSurvCap);         // quads surviving rejection stage a: item | flags<<12
  uint16_t* cornerq = itemq + ((G.fastTileH * sp) >> 2);        // pixels whose strength exceeds the pass threshold

  int pitch;
  const uint8_t* lvl = level_ptr(G, B, l, f, pitch);
  const int gx0 = kMinBorder + tx0;
  int ox = 0;
  const bool viaTma = TM.use[l] != 0;
  if (viaTma) {
    // ---- one TMA tile load (UTMALDG): box = fastTileW x fastTileH bytes at (gx0 & ~15, 16+ty0, frame); rows land with
    //      pitch sp, out-of-image parts are zero-filled and never read.  Completion is tracked by an mbarrier.
    //      The innermost TMA coordinate must be 16-byte aligned (an unaligned one raises "illegal instruction" on
    //      B200, measured with tools/scratch/tma_test*.cu), hence the same `ox` shift as the vector-load path.
    ox = gx0 & 15;
    if (tid == 0) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tmaBar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(sp * G.fastTileH)) : "memory");
      const int z = (l == 0 ? TM.frame0 : 0) + f;
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              (uint32_t)__cvta_generic_to_shared(img)),
          "l"(reinterpret_cast<uint64_t>(TM.map + l)), "r"(gx0 - ox), "r"(kMinBorder + ty0), "r"(z), "r"(bar)
          : "memory");
    }
  } else if ((((uintptr_t)lvl | (uintptr_t)pitch) & 15) == 0) {
    ox = gx0 & 15;
    const int wpr = (ox + tw + 15) >> 4;                       // 16-byte words per tile row (<= 32)
    const uint8_t* base = lvl + (size_t)(kMinBorder + ty0) * pitch + (gx0 - ox);
    const int lg = wpr <= 16 ? 4 : 5;                          // slots per row = 1 << lg
    const int w = tid & ((1 << lg) - 1);
    if (w < wpr)
      for (int y = tid >> lg; y < th; y += kFastThreads >> lg)
        *reinterpret_cast<uint4*>(img + y * sp + 16 * w) = __ldg(reinterpret_cast<const uint4*>(base + (size_t)y * pitch) + w);
  } else {
    const uint8_t* base = lvl + (size_t)(kMinBorder + ty0) * pitch + gx0;
    for (int y = warp; y < th; y += kFastThreads / 32)
      for (int x = lane; x < tw; x += 32) img[y * sp + x] = __ldg(base + (size_t)y * pitch + x);
  }
  {
    uint4* z = reinterpret_cast<uint4*>(sc);
    const int nz = (th * sp) >> 4;
    for (int i = tid; i < nz; i += kFastThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid < kCellsPerCta) cellCnt[tid] = 0;
  if (tid == 0) { qn = 0; sn = 0; in = 0; cn = 0; retryMask = 0; }
  __syncthreads();
  if (viaTma) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tmaBar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra TD_%=;\n\t"
        "bra TW_%=;\n\t"
        "TD_%=:\n\t}" ::"r"(bar) : "memory");
  }

  const int w1 = L.wCell, w2 = 2 * L.wCell, w3 = 3 * L.wCell;
  const uint8_t* img0 = img + 3 * sp + 3 + ox;                 // interior origin
  uint8_t* sc0 = sc + 3 * sp + 3 + ox;

  int t = G.iniTh;
  int mask = 0xF;                                              // cells taking part in this pass
  for (int pass = 0; pass < 2; ++pass) {
    // ---- rejection test, 4 adjacent centres per thread in packed bytes.  Stage a (rows -3,0,+3) runs over every
    //      quad; quads with a surviving centre are queued and stage b (rows +-1,+-2) runs densely over that queue;
    //      surviving centres go to the pixel queue.
    {
      const uint32_t kc = 0x7F7F7F7Fu - (uint32_t)(t & 0x7F) * 0x01010101u;
      const bool hiT = t >= 128;
      const int nq4 = (iw + 3) >> 2;
      const uint32_t magic = 0xFFFFFFFFu / (uint32_t)nq4 + 1;            // item / nq4 == umulhi(item, magic) for item < 2^16
      const int sh = ox & 3;
      const int aox = ox & ~3;
      for (int item = tid; item < ih * nq4; item += kFastThreads) {
        const int yi = nq4 == 1 ? item : (int)__umulhi((uint32_t)item, magic), xi = 4 * (item - yi * nq4);
        const uint8_t* rowm3 = img + yi * sp + aox + xi;                // tile row yi (= interior row yi - 3), aligned window
        uint32_t m;
        switch (sh) {
          case 0: m = reject4_a<0>(rowm3, sp, kc, hiT); break;
          case 1: m = reject4_a<1>(rowm3, sp, kc, hiT); break;
          case 2: m = reject4_a<2>(rowm3, sp, kc, hiT); break;
          default: m = reject4_a<3>(rowm3, sp, kc, hiT); break;
        }
        m &= 0x80808080u;
        if (m) {
          const uint32_t bits = ((m >> 7) & 1) | ((m >> 14) & 2) | ((m >> 21) & 4) | ((m >> 28) & 8);
          itemq[atomicAdd(&in, 1)] = (uint16_t)(item | (bits << 12));
        }
      }
      __syncthreads();
      const int ni = in;
      for (int k = tid; k < ni; k += kFastThreads) {
        const int e = itemq[k], item = e & 0xFFF;
        const uint32_t bits = e >> 12;
        const int yi = nq4 == 1 ? item : (int)__umulhi((uint32_t)item, magic), xi = 4 * (item - yi * nq4);
        const uint8_t* rowm3 = img + yi * sp + aox + xi;
        uint32_t m = ((bits & 1) << 7) | ((bits & 2) << 14) | ((bits & 4) << 21) | ((bits & 8) << 28);
        switch (sh) {
          case 0: m = reject4_b<0>(rowm3, sp, kc, hiT, m); break;
          case 1: m = reject4_b<1>(rowm3, sp, kc, hiT, m); break;
          case 2: m = reject4_b<2>(rowm3, sp, kc, hiT, m); break;
          default: m = reject4_b<3>(rowm3, sp, kc, hiT, m); break;
        }
        // keep centres that exist (x < iw) and whose cell takes part in this pass
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = xi + j;
          const int jj = (x >= w1) + (x >= w2) + (x >= w3);
          if (((m >> (8 * j + 7)) & 1) && x < iw && ((mask >> jj) & 1)) keep |= 1u << j;
        }
        if (keep) {
          int pos = atomicAdd(&qn, __popc(keep));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((keep >> j) & 1) queue[pos++] = (uint16_t)((yi << 8) | (xi + j));
        }
      }
    }
    __syncthreads();
    const int nq = qn;
    // ---- full arc score, densely over the queue, two entries per thread (packed 16-bit min/max network)
    for (int k = tid; 2 * k < nq; k += kFastThreads) {
      const int e0 = queue[2 * k], e1 = queue[min(2 * k + 1, nq - 1)];
      const int y0 = e0 >> 8, x0 = e0 & 255, y1 = e1 >> 8, x1 = e1 & 255;
      int b0, b1;
      fast_best2(img0 + y0 * sp + x0, img0 + y1 * sp + x1, sp, b0, b1);
      const bool c0 = b0 > t, c1 = b1 > t && (2 * k + 1 < nq);
      sc0[y0 * sp + x0] = (uint8_t)(c0 ? b0 : 0);
      if (2 * k + 1 < nq) sc0[y1 * sp + x1] = (uint8_t)(b1 > t ? b1 : 0);
      if (c0 | c1) {
        int pos = atomicAdd(&cn, (int)c0 + (int)c1);
        if (c0) cornerq[pos++] = (uint16_t)e0;
        if (c1) cornerq[pos] = (uint16_t)e1;
      }
    }
    __syncthreads();
    // ---- 3x3 NMS inside each cell's own candidate rectangle (outside counts as score 0, like the zeroed border of
    //      cv::FAST on the cell ROI); strict '>' against all eight neighbours
    const int nc = cn;
    for (int k = tid; k < nc; k += kFastThreads) {
      const int e = cornerq[k], yi = e >> 8, xi = e & 255;
      const uint8_t* s = sc0 + yi * sp + xi;
      const int v = s[0];
      const int jj = (xi >= w1) + (xi >= w2) + (xi >= w3);
      const int cx = xi - jj * L.wCell;
      const int cw = min(L.wCell, iw - jj * L.wCell);
      const bool hasL = cx > 0, hasR = cx + 1 < cw, hasU = yi > 0, hasD = yi + 1 < ih;
      bool keep = true;
      keep &= !(hasU && hasL && s[-sp - 1] >= v);
      keep &= !(hasU && s[-sp] >= v);
      keep &= !(hasU && hasR && s[-sp + 1] >= v);
      keep &= !(hasL && s[-1] >= v);
      keep &= !(hasR && s[1] >= v);
      keep &= !(hasD && hasL && s[sp - 1] >= v);
      keep &= !(hasD && s[sp] >= v);
      keep &= !(hasD && hasR && s[sp + 1] >= v);
      if (keep) {
        atomicAdd(&cellCnt[jj], 1);
        surv[atomicAdd(&sn, 1)] = ((uint32_t)jj << 24) | ((uint32_t)yi << 12) | (uint32_t)cx;
      }
    }
    __syncthreads();
    // ---- cells that came back empty are retried at minThFAST (ORBextractor.cpp:820-824)
    if (pass == 1 || G.minTh >= G.iniTh) break;
    if (tid == 0) {
      int r = 0;
      for (int jj = 0; jj < j1 - j0; ++jj)
        if (cellCnt[jj] == 0 && jj * L.wCell < iw) r |= 1 << jj;
      retryMask = r;
      qn = 0;
      in = 0;
      cn = 0;
    }
    __syncthreads();
    mask = retryMask;
    if (mask == 0) break;
    t = G.minTh;
  }

  // ---- ordered emission: rank of a survivor = number of survivors with a smaller (cell, row, column) key
  const int S = sn;
  uint32_t* out = B.slotKeys + (size_t)f * G.slotKeysPerFrame + __ldg(B.slotKeyBase + slot);
  for (int k = tid; k < S; k += kFastThreads) {
    const uint32_t key = surv[k];
    int rank = 0;
    for (int j = 0; j < S; ++j) rank += surv[j] < key ? 1 : 0;
    const int jj = key >> 24, yi = (key >> 12) & 0xFFF, cx = key & 0xFFF;
    const int xi = jj * L.wCell + cx;
    out[rank] = pack_key(tx0 + xi + 3, ty0 + yi + 3, sc0[yi * sp + xi] - 1);
  }
  if (tid == 0) *out_count = S;
}

// ---- warp-synchronous variant (the default): one warp owns one cell end to end, so the phases are separated by
// __syncwarp instead of CTA barriers (barrier stalls were the top stall reason of fast_kernel) and queue appends use
// ballots / warp scans instead of shared-memory atomics.  Same arithmetic and the same output order as fast_kernel.
constexpr int kWarpCells = kCellsPerCta;          // cells (= warps) per CTA of the warp-synchronous variant

template <int SH>
__device__ __forceinline__ uint32_t fw_stage_a(const uint8_t* rowm3, int sp, uint32_t kc, bool hiT) { return reject4_a<SH>(rowm3, sp, kc, hiT); }

// The two rejection stages of one cell as functions of the window misalignment SH (a warp-uniform runtime value) for
// t < 128: dispatching ONCE per stage keeps the per-item code free of the 4-way switch and, more importantly, of the
// predicated double implementation of far4 (ptxas issued both threshold forms of every test, predicated on hiT: 6 issue
// slots per ring pixel instead of 3).  FILTER = false: pass everything (see FW_DISPATCH).
template <int SH, bool FILTER>
__device__ __forceinline__ int fw_stage_a(const uint8_t* img, int sp, int aox, int nq4, uint32_t magic, int nItems, uint32_t kc,
                                          uint16_t* itemq, int lane) {
  int ni = 0;
#pragma unroll 1
  for (int base = 0; base < nItems; base += 32) {
    const int item = base + lane;
    uint32_t m = 0;
    if (item < nItems) {
      const int yi = nq4 == 1 ? item : (int)__umulhi((uint32_t)item, magic), xq = 4 * (item - yi * nq4);
      m = FILTER ? reject4_a<SH>(img + yi * sp + aox + xq, sp, kc, false) & 0x80808080u : 0x80808080u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m != 0);
    if (m) itemq[ni + __popc(bal & ((1u << lane) - 1))] = (uint16_t)(item | (gather_flags(m) << 12));
    ni += __popc(bal);
  }
  return ni;
}
template <int SH, bool FILTER>
__device__ __forceinline__ int fw_stage_b(const uint8_t* img, int sp, int aox, int nq4, uint32_t magic, int cw, uint32_t kc,
                                          const uint16_t* itemq, int ni, uint16_t* queue, int lane) {
  int nq = 0;
#pragma unroll 1
  for (int base = 0; base < ni; base += 32) {
    const int k = base + lane;
    uint32_t keep = 0;
    int yi = 0, xq = 0;
    if (k < ni) {
      const int e = itemq[k], item = e & 0xFFF;
      yi = nq4 == 1 ? item : (int)__umulhi((uint32_t)item, magic); xq = 4 * (item - yi * nq4);
      const uint32_t m0 = scatter_flags((uint32_t)e >> 12);
      const uint32_t m = FILTER ? reject4_b<SH>(img + yi * sp + aox + xq, sp, kc, false, m0) : m0;
      // centres xq + j >= cw lie outside the cell: keep the flags of the first cw - xq bytes only
      keep = gather_flags(m & (0x80808080u >> (8 * max(0, 4 - (cw - xq)))));
    }
    // warp-ordered append of up to 4 pixels per lane: position = pixels kept by lower lanes (four independent ballots
    // instead of a five-step shuffle scan) + pixels kept by this lane before j
    const unsigned lt = (1u << lane) - 1;
    const unsigned b0 = __ballot_sync(0xffffffffu, keep & 1), b1 = __ballot_sync(0xffffffffu, keep & 2),
                   b2 = __ballot_sync(0xffffffffu, keep & 4), b3 = __ballot_sync(0xffffffffu, keep & 8);
    int pos = nq + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((keep >> j) & 1) queue[pos++] = (uint16_t)((yi << 8) | (xq + j));
    nq += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
  }
  return nq;
}
// t >= 128 is legal but meaningless in practice.  The rejection stages are only filters in front of the exact arc score,
// so that case simply runs without them (every pixel of the cell is scored): correct, slow, and no second set of four
// variants in the kernel (with all eight inlined, instruction-fetch stalls ate the whole gain of the dispatch).
#define FW_DISPATCH(RES, FN, ...)                                                                   \
  do {                                                                                              \
    if (hiT || FW_NOFILTER_##FN) RES = FN<0, false>(__VA_ARGS__);                                   \
    else switch (sh) {                                                                              \
        case 0: RES = FN<0, true>(__VA_ARGS__); break;                                              \
        case 1: RES = FN<1, true>(__VA_ARGS__); break;                                              \
        case 2: RES = FN<2, true>(__VA_ARGS__); break;                                              \
        default: RES = FN<3, true>(__VA_ARGS__); break;                                             \
      }                                                                                             \
  } while (0)

// A/B switch: ORBX_FAST_NOB=1 runs stage b without its tests (stage-a survivors go straight to the arc score)
#ifndef ORBX_FAST_NOB
#define ORBX_FAST_NOB 0
#endif
#define FW_NOFILTER_fw_stage_a 0
#define FW_NOFILTER_fw_stage_b ORBX_FAST_NOB
#ifndef ORBX_FAST_MINB
#define ORBX_FAST_MINB 16
#endif
// kDivMagic[n] = 0xFFFFFFFF / n + 1 (n = 1..32; cells are at most 63 px = 16 quads wide): division by a warp-uniform small
// divisor without the ~20-instruction integer division sequence
__constant__ uint32_t kDivMagic[33] = {0u, 0u /* n == 1 is special-cased by the callers */, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u,
    0x2AAAAAABu, 0x24924925u, 0x20000000u, 0x1C71C71Du, 0x1999999Au, 0x1745D175u, 0x15555556u, 0x13B13B14u, 0x12492493u, 0x11111112u,
    0x10000000u, 0x0F0F0F10u, 0x0E38E38Fu, 0x0D79435Fu, 0x0CCCCCCDu, 0x0C30C30Du, 0x0BA2E8BBu, 0x0B21642Du, 0x0AAAAAABu, 0x0A3D70A4u,
    0x09D89D8Au, 0x097B425Fu, 0x0924924Au, 0x08D3DCB1u, 0x08888889u, 0x08421085u, 0x08000000u};

__global__ void __launch_bounds__(32 * kWarpCells, ORBX_FAST_MINB) fast_warp_kernel(const __grid_constant__ Geom G, const Bufs B, const __grid_constant__ TmaSet TM) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int cellOut[kWarpCells];
  __shared__ __align__(8) uint64_t tmaBar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if ORBX_FAST_FF
  const int slot = blockIdx.y + B.slotOff, f = blockIdx.x;
#else
  const int slot = blockIdx.x + B.slotOff, f = blockIdx.y;
#endif
  const int keyBase = __ldg(B.slotKeyBase + slot);        // needed only by the emission: fetched now, off the critical path
  int l = 0;
  while (l + 1 < G.nlevels && slot >= G.L[l + 1].slot0) ++l;
  const LevelGeom& L = G.L[l];
  const int ls = slot - L.slot0;
  const int ci = L.groups == 1 ? ls : (int)__umulhi((uint32_t)ls, L.groupsMagic), grp = ls - ci * L.groups;   // no integer division
  const int j0 = grp * kCellsPerCta, j1 = min(j0 + kCellsPerCta, L.nCols);
  const int tx0 = j0 * L.wCell, tx1 = min(j1 * L.wCell + 6, L.regW);
  const int ty0 = ci * L.hCell, ty1 = min(ty0 + L.hCell + 6, L.regH);
  const int tw = tx1 - tx0, th = ty1 - ty0;
  const int iw = tw - 6, ih = th - 6;
  int* out_count = B.slotCount + (size_t)f * G.totalSlots + slot;
  if (iw <= 0 || ih <= 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  const int sp = G.fastTileW;
  const int tileBytes = G.fastTileH * sp;
  uint8_t* img = smem;
  uint8_t* sc = smem + tileBytes;                        // corner strengths of the cell interiors, each cell framed by a zero border
  const int scp = G.fastScW, scBytes = G.fastScW * G.fastScH;
  // per-warp scratch: queue (u16 per cell pixel) and one region shared by itemq (u16 per quad, dead after stage b) and
  // surv (u16 per NMS survivor, written after stage b; a retry pass only happens when no survivor was written)
  const int cellPix = G.fastCellPix;
  const int shared2 = (max(2 * G.fastCellQuads, 2 * G.fastCellSurv) + 15) & ~15;
  uint8_t* wbase = sc + scBytes + (size_t)warp * (2 * cellPix + shared2);
  uint16_t* queue = (uint16_t*)wbase;
  uint16_t* itemq = queue + cellPix;
  uint16_t* surv = itemq;

  int pitch;
  const uint8_t* lvl = level_ptr(G, B, l, f, pitch);
  const int gx0 = kMinBorder + tx0;
  int ox = 0;
  const bool viaTma = TM.use[l] != 0;
  if (viaTma) {
    ox = gx0 & 15;
    if (tid == 0) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tmaBar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)tileBytes) : "memory");
      const int z = (l == 0 ? TM.frame0 : 0) + f;
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              (uint32_t)__cvta_generic_to_shared(img)),
          "l"(reinterpret_cast<uint64_t>(TM.map + l)), "r"(gx0 - ox), "r"(kMinBorder + ty0), "r"(z), "r"(bar)
          : "memory");
    }
  } else if ((((uintptr_t)lvl | (uintptr_t)pitch) & 15) == 0) {
    ox = gx0 & 15;
    const int wpr = (ox + tw + 15) >> 4;
    const uint8_t* base = lvl + (size_t)(kMinBorder + ty0) * pitch + (gx0 - ox);
    for (int i = tid; i < th * wpr; i += 32 * kWarpCells) {
      const int y = i / wpr, w = i - y * wpr;
      *reinterpret_cast<uint4*>(img + y * sp + 16 * w) = __ldg(reinterpret_cast<const uint4*>(base + (size_t)y * pitch) + w);
    }
  } else {
    const uint8_t* base = lvl + (size_t)(kMinBorder + ty0) * pitch + gx0;
    for (int y = warp; y < th; y += kWarpCells)
      for (int x = lane; x < tw; x += 32) img[y * sp + x] = __ldg(base + (size_t)y * pitch + x);
  }
  {
    uint4* z = reinterpret_cast<uint4*>(sc);
    for (int i = tid; i < ((ih + 2) * scp) >> 4; i += 32 * kWarpCells) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if (viaTma) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tmaBar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra TD_%=;\n\t"
        "bra TW_%=;\n\t"
        "TD_%=:\n\t}" ::"r"(bar) : "memory");
  }

  // ---- this warp's cell -----------------------------------------------------------------------------------------
  const int jj = warp;
  const int cx0 = jj * L.wCell;                           // first interior column of the cell (tile-interior coordinates)
  const int cw = min(L.wCell, iw - cx0);                  // cell width (<= 0: the cell does not exist in this slot)
  int S = 0;
  const uint8_t* img0 = img + 3 * sp + 3 + ox;
  // this cell's score map: sc0[yi * scp + cx], yi in [-1, ih], cx in [-1, cw]; the frame stays 0 = "outside the cell counts
  // as 0" (the zeroed border of cv::FAST on the cell ROI), so the NMS needs no edge flags
  uint8_t* sc0 = sc + scp + jj * (L.wCell + 2) + 1;
  if (jj < j1 - j0 && cw > 0) {
    const int nq4 = (cw + 3) >> 2;
    const uint32_t magic = kDivMagic[min(nq4, 32)];       // item / nq4 == umulhi(item, magic) for item < 2^16
    const int sh = (ox + cx0) & 3, aox = (ox + cx0) & ~3;  // quads start at the cell's first pixel
    int t = G.iniTh;
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t kc = 0x7F7F7F7Fu - (uint32_t)(t & 0x7F) * 0x01010101u;
      const bool hiT = t >= 128;
      // stage a over every quad of the cell -> quads with a surviving centre
      int ni = 0;
      FW_DISPATCH(ni, fw_stage_a, img, sp, aox, nq4, magic, ih * nq4, kc, itemq, lane);
      __syncwarp();
      // stage b densely over the surviving quads -> pixel queue
      int nq = 0;
      FW_DISPATCH(nq, fw_stage_b, img, sp, aox, nq4, magic, cw, kc, itemq, ni, queue, lane);
      __syncwarp();
      // score two queue entries per lane; corners are compacted in place at the front of the same queue
      int nc = 0;
      for (int base = 0; base < nq; base += 64) {
        const int k0 = base + 2 * lane, k1 = k0 + 1;
        int e0 = 0, e1 = 0, b0 = 0, b1 = 0;
        if (k0 < nq) {
          e0 = queue[k0]; e1 = queue[min(k1, nq - 1)];
          const int y0 = e0 >> 8, x0 = e0 & 255, y1 = e1 >> 8, x1 = e1 & 255;
          fast_best2(img0 + y0 * sp + cx0 + x0, img0 + y1 * sp + cx0 + x1, sp, b0, b1);
          sc0[y0 * scp + x0] = (uint8_t)(b0 > t ? b0 : 0);
          if (k1 < nq) sc0[y1 * scp + x1] = (uint8_t)(b1 > t ? b1 : 0);
        }
        const bool c0 = k0 < nq && b0 > t, c1 = k1 < nq && b1 > t;
        const unsigned lt = (1u << lane) - 1;
        const unsigned bal0 = __ballot_sync(0xffffffffu, c0), bal1 = __ballot_sync(0xffffffffu, c1);
        __syncwarp();                                  // every lane has read its entries before anyone overwrites the front
        int pos = nc + __popc(bal0 & lt) + __popc(bal1 & lt);
        if (c0) queue[pos++] = (uint16_t)e0;
        if (c1) queue[pos] = (uint16_t)e1;
        nc += __popc(bal0) + __popc(bal1);
        __syncwarp();
      }
      // NMS inside the cell (the zero frame of the score map = "outside the cell counts as 0"); the ballot compaction keeps
      // the queue's (row, column) order, which is the emission order
      for (int base = 0; base < nc; base += 32) {
        const int k = base + lane;
        bool keep = false;
        int e = 0;
        if (k < nc) {
          e = queue[k];
          const int yi = e >> 8, cx = e & 255;
          const uint8_t* s = sc0 + yi * scp + cx;
          const int v = s[0];
          keep = (v > s[-scp - 1]) & (v > s[-scp]) & (v > s[-scp + 1]) & (v > s[-1]) & (v > s[1]) & (v > s[scp - 1]) & (v > s[scp]) & (v > s[scp + 1]);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) surv[S + __popc(bal & ((1u << lane) - 1))] = (uint16_t)e;      // yi<<8 | cx : already the (row, col) order key
        S += __popc(bal);
      }
      __syncwarp();
      if (S > 0 || pass == 1 || G.minTh >= G.iniTh) break;     // empty cell: retry at minThFAST (ORBextractor.cpp:820-824)
      t = G.minTh;
    }
  }
  // ---- ordered emission: cells left to right, inside a cell by (row, column) ----------------------------------------
  if (lane == 0) cellOut[warp] = S;
  __syncthreads();
  int off = 0, total = 0;
  for (int w = 0; w < kWarpCells; ++w) { const int c = cellOut[w]; if (w < warp) off += c; total += c; }
  uint32_t* out = B.slotKeys + (size_t)f * G.slotKeysPerFrame + keyBase;
  // Every queue of this warp is filled in increasing (row, column) order and every compaction (ballot / warp scan) keeps
  // that order, so surv[] is already sorted by its key: position = rank, no ranking pass.
  for (int k = lane; k < S; k += 32) {
    const uint32_t key = surv[k];
    const int yi = key >> 8, cx = key & 255;
    out[off + k] = pack_key(tx0 + cx0 + cx + 3, ty0 + yi + 3, sc0[yi * scp + cx] - 1);
  }
  if (tid == 0) *out_count = total;
}

// ======================================================================================================
// K3  DistributeOctTree (ORBextractor.cpp:545-769) -- one CTA per (level, frame).
//
// Keys are never moved: each key carries the list position of its node; a pass counts keys per child
// quadrant, rebuilds the (<= N+3 entry) node list in closed form and relabels the keys.  The new list is
// [children of the last processed parent as n4,n3,n2,n1] ... [children of the first processed parent],
// then the untouched nodes in old order.  The "largest first" phase processes expandable nodes in a stable
// descending-size order of the list (== the reference's sort of (size, node*) read from the back when node
// addresses grow in creation order), stopping right after the split that reaches N nodes.
// ======================================================================================================
struct OctSmem {
  short4* box[2];   // x0,y0,x1,y1
  int* cnt[2];
  int* cnt4;        // [cap*4]
  int* rank;        // [cap]
  int* childR;      // [cap] children count by rank -> exclusive scan
  int* keepPos;     // [cap]  new position of an unsplit node (or -1)
  int* childPos;    // [cap*4]
  unsigned long long* best;  // [cap]
  int* slotPre;     // [maxSlots+1]
  int* ws;          // [40]
};
// (node labels in shared memory instead of global scratch were measured and dropped: 12 KB more per CTA costs a third of the
// resident CTAs, 1.51 -> 1.73 ms per 4096 VGA frames)
#ifndef ORBX_OCT_BATCH
#define ORBX_OCT_BATCH 4
#endif
constexpr int kOctBatch = ORBX_OCT_BATCH;   // independent (label, key) fetches in flight per thread in the per-key loops

__device__ __forceinline__ int quadrant_of(const short4 b, int x, int y) {
  const int mx = b.x + ((b.z - b.x + 1) >> 1);   // ceil(float(x1-x0)/2)   (ORBextractor.cpp:489-490)
  const int my = b.y + ((b.w - b.y + 1) >> 1);
  return (x < mx ? 0 : 1) | (y < my ? 0 : 2);     // 0:n1 1:n2 2:n3 3:n4   (:521-531)
}

// (warp-aggregated shared atomics -- __match_any_sync groups electing one leader per node -- were measured and dropped: the
// quadtree went 1.47 -> 1.80 ms per 4096 frames and 36 -> 39 us for one frame; same-address shared atomics are not its bound)
// Exclusive scan of `n` ints in shared memory, in place, by ONE warp (all 32 lanes call it; the data must be visible to the
// warp: __syncwarp before).  Returns the total.
__device__ __forceinline__ int warp_exclusive_scan(int* data, int n) {
  const int lane = threadIdx.x & 31;
  const int per = (n + 31) >> 5;
  const int lo = min(lane * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += data[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int run = incl - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = data[i];
    data[i] = run;
    run += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  return total;
}

#ifndef ORBX_OCT_TRACE
#define ORBX_OCT_TRACE 0            // debugging aid: level-0 CTA of frame 0 prints its phase timeline (cycles)
#endif
#if ORBX_OCT_TRACE
#define OCT_STAMP(name) do { if (tid == 0 && l == 0 && f == 0 && nst < 48) { st_t[nst] = clock64(); st_n[nst++] = name; } } while (0)
#else
#define OCT_STAMP(name) do { } while (0)
#endif
template <bool KS>   // KS: latency mode, keys and node labels in shared memory (Bufs::octKeySmem)
__global__ void __launch_bounds__(kOctMaxThreads, 1) octree_kernel(const __grid_constant__ Geom G, const Bufs B) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, T = blockDim.x;            // 128 ... 1024 threads, chosen by the host from the frame size
#if ORBX_OCT_FF
  const int l = blockIdx.y + B.levelOff, f = blockIdx.x;
#else
  const int l = blockIdx.x + B.levelOff, f = blockIdx.y;
#endif
  const LevelGeom& L = G.L[l];
  const int cap = G.nodeCap;
#if ORBX_OCT_TRACE
  long long st_t[48]; const char* st_n[48]; int nst = 0;
#endif
  OCT_STAMP("start");
  OctSmem S;
  {
    uint8_t* p = smem_raw;
    S.best = (unsigned long long*)p; p += sizeof(unsigned long long) * cap;
    S.box[0] = (short4*)p; p += sizeof(short4) * cap;
    S.box[1] = (short4*)p; p += sizeof(short4) * cap;
    S.cnt[0] = (int*)p; p += 4 * cap;
    S.cnt[1] = (int*)p; p += 4 * cap;
    S.cnt4 = (int*)p; p += 16 * cap;
    S.rank = (int*)p; p += 4 * cap;
    S.childR = (int*)p; p += 4 * (cap + 1);
    S.keepPos = (int*)p; p += 4 * cap;
    S.childPos = (int*)p; p += 16 * cap;
    S.slotPre = (int*)p; p += 4 * (G.maxSlotsPerLevel + 1);
    S.ws = (int*)p;
  }
  __shared__ int sh_stop, sh_res[3];
  __shared__ int rootCnt[kMaxRoots], rootPos[kMaxRoots];

  uint32_t* keys = B.flatKeys + (size_t)f * G.keysPerFrame + L.keyBase;
  uint16_t* nodeOf = B.nodeOf + (size_t)f * G.keysPerFrame + L.keyBase;
  int* selCount = B.selCount + (size_t)f * G.nlevels + l;
  uint32_t* sel = B.sel + (size_t)f * G.selPerFrame + L.selBase;

  // ---- gather the level's candidates from the FAST slots, in slot order -------------------------------
  const int* slotCount = B.slotCount + (size_t)f * G.totalSlots + L.slot0;
  for (int s = tid; s < L.nSlots; s += T) S.slotPre[s] = slotCount[s];
  __syncthreads();
  const int n = block_exclusive_scan(S.slotPre, L.nSlots, S.ws);
  if (tid == 0) B.candCount[(size_t)f * G.nlevels + l] = n;
  if (n == 0) {
    if (tid == 0) *selCount = 0;
    return;
  }
  // Latency mode (a handful of CTAs per launch, occupancy is irrelevant): keys and node labels of this level stay in shared
  // memory, so the two key sweeps of every pass cost shared-memory instead of L2 latency (the sweeps were ~80 % of the
  // single-frame kernel).  The flat key list still goes to global memory once (stage taps read it).
  const bool inS = KS && n <= B.octKeySmem;
  uint32_t* gkeys = keys;
  if (KS && inS) {
    keys = reinterpret_cast<uint32_t*>(S.ws + 40);                       // behind the scan scratch: [octKeySmem] keys, then labels
    nodeOf = reinterpret_cast<uint16_t*>(keys + B.octKeySmem);
  }
  {
    // one THREAD per KEY: the slot of flat position k is found by a binary search in the scanned slot counts (shared memory),
    // then its key is one independent load -- four keys in flight per thread.  (One thread per slot walked a slot's few dozen keys
    // as dependent rounds of four: 8400 of the 45000 cycles of a single-frame level-0 CTA.)
    const uint32_t* slotKeys = B.slotKeys + (size_t)f * G.slotKeysPerFrame;
    const int* keyBase = B.slotKeyBase + L.slot0;
    const int nS = L.nSlots;
    for (int k0 = tid; k0 < n; k0 += 4 * T) {
      uint32_t v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + j * T;
        v[j] = 0u;
        if (k < n) {
          int lo = 0, hi = nS - 1;                   // largest s with slotPre[s] <= k (skips empty slots)
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (S.slotPre[mid] <= k) lo = mid; else hi = mid - 1;
          }
          v[j] = slotKeys[__ldg(keyBase + lo) + (k - S.slotPre[lo])];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + j * T;
        if (k < n) { gkeys[k] = v[j]; if (KS && inS) keys[k] = v[j]; }
      }
    }
  }
  if (tid < kMaxRoots) rootCnt[tid] = 0;
  __syncthreads();
  OCT_STAMP("gather");

  // ---- roots (ORBextractor.cpp:549-590) ----------------------------------------------------------------
  int cur = 0;
  if (L.nIni == 1) {
    for (int k = tid; k < n; k += T) nodeOf[k] = 0;
    if (tid == 0) { rootCnt[0] = n; }
  } else {
    for (int k = tid; k < n; k += T) {
      int r = (int)__fdiv_rn((float)key_x(keys[k]), L.hX);
      r = min(r, L.nIni - 1);
      nodeOf[k] = (uint16_t)r;
      atomicAdd(&rootCnt[r], 1);
    }
  }
  __syncthreads();
  int Scount;
  if (tid == 0) {
    int s = 0;
    for (int r = 0; r < L.nIni; ++r) {
      if (rootCnt[r] > 0) {
        rootPos[r] = s;
        S.box[0][s] = make_short4((short)L.rootX[r], 0, (short)L.rootX[r + 1], (short)L.regH);
        S.cnt[0][s] = rootCnt[r];
        ++s;
      } else rootPos[r] = -1;
    }
    sh_stop = s;
  }
  __syncthreads();
  Scount = sh_stop;
  if (L.nIni > 1) {
    for (int k = tid; k < n; k += T) nodeOf[k] = (uint16_t)rootPos[nodeOf[k]];
  }
  __syncthreads();

  const int N = L.nFeat;
  int phase = 1;
  bool finish = false;
  const int lane = tid & 31, wid = tid >> 5;
  // One pass = [count sweep, all threads] barrier [node list rebuild, WARP 0 ALONE with warp scans and __syncwarp] barrier
  // [relabel sweep, all threads].  The rebuild works on at most N + 3 nodes: spread over the whole CTA it was ~20 CTA barriers
  // per pass (barrier stalls were half of the kernel's samples); one warp does it in a few hundred instructions while the
  // other warps of the SM's CTAs run.  Only the quadratic phase-2 ranking stays CTA-wide.  Every thread owns the keys
  // k = tid (mod T) in all sweeps, so the node labels need no barrier between the relabel sweep and the next count sweep.
  for (int i = tid; i < Scount * 4; i += T) S.cnt4[i] = 0;
  __syncthreads();
  OCT_STAMP("roots");
  while (!finish) {
    // selects instead of S.box[cur]: a runtime index would put the pointer struct into local memory
    const short4* box = cur ? S.box[1] : S.box[0];
    const int* cnt = cur ? S.cnt[1] : S.cnt[0];
    short4* nbox = cur ? S.box[0] : S.box[1];
    int* ncnt = cur ? S.cnt[0] : S.cnt[1];
    const int Sn = Scount;

    // keys and node labels live in global memory (L2): four independent (label, key) pairs are fetched per thread before
    // anything is done with them, instead of label -> test -> key -> atomic one key at a time (two dependent L2 latencies each)
    for (int k0 = tid; k0 < n; k0 += kOctBatch * T) {
      int pp[kOctBatch]; uint32_t kk[kOctBatch];
#pragma unroll
      for (int j = 0; j < kOctBatch; ++j) {
        const int k = k0 + j * T;
        pp[j] = k < n ? (int)nodeOf[k] : -1;
        kk[j] = k < n ? keys[k] : 0u;
      }
#pragma unroll
      for (int j = 0; j < kOctBatch; ++j)
        if (pp[j] >= 0 && cnt[pp[j]] > 1) atomicAdd(&S.cnt4[pp[j] * 4 + quadrant_of(box[pp[j]], key_x(kk[j]), key_y(kk[j]))], 1);
    }
    __syncthreads();
    OCT_STAMP("count");

    if (phase == 2) {
      // stable descending-size order of the list (reference: sort of (size, node*) walked from the back)
      for (int p = tid; p < Sn; p += T) {
        const int c = cnt[p];
        if (c > 1) {
          int r = 0;
          for (int q = 0; q < Sn; ++q) {
            const int cq = cnt[q];
            r += (cq > 1 && (cq > c || (cq == c && q < p))) ? 1 : 0;
          }
          S.rank[p] = r;
        }
      }
      __syncthreads();
    }

    if (wid == 0 && phase == 1) {
      // Phase 1 splits EVERY expandable node in list order, so the new position of a node's children is the number of
      // children of the expandable nodes behind it, and a kept node follows all children in list order: one packed warp scan
      // (children << 16 | kept) over contiguous chunks of the list replaces the rank / per-rank child count / kept scans.
      const int per = (Sn + 31) >> 5, p0 = min(lane * per, Sn), p1 = min(p0 + per, Sn);
      int mine = 0;
      for (int p = p0; p < p1; ++p) {
        if (cnt[p] > 1) {
          const int* c4 = S.cnt4 + p * 4;
          mine += ((c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0)) << 16;
        } else ++mine;
      }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      const int totalChildren = tot >> 16, nKept = tot & 0xFFFF;
      int chBefore = (incl - mine) >> 16, keptBefore = (incl - mine) & 0xFFFF, nexp = 0;
      for (int p = p0; p < p1; ++p) {
        int* cp = S.childPos + p * 4;
        if (cnt[p] > 1) {
          const int* c4 = S.cnt4 + p * 4;
          const int nch = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
          int o = totalChildren - chBefore - nch;         // blocks of later-processed parents come first
          chBefore += nch;
          const short4 b = box[p];
          const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
#pragma unroll
          for (int q = 3; q >= 0; --q) {
            if (c4[q] > 0) {
              cp[q] = o;
              nbox[o] = make_short4((q & 1) ? mx : b.x, (q & 2) ? my : b.y, (q & 1) ? b.z : mx, (q & 2) ? b.w : my);
              ncnt[o] = c4[q];
              nexp += c4[q] > 1 ? 1 : 0;
              ++o;
            } else cp[q] = -1;
          }
          S.keepPos[p] = -1;
        } else {
          const int o = totalChildren + keptBefore;
          nbox[o] = box[p];
          ncnt[o] = cnt[p];
          S.keepPos[p] = keptBefore++;
        }
      }
      nexp = __reduce_add_sync(0xffffffffu, nexp);
      __syncwarp();                                    // every lane has read its nodes' cnt4 before the zeroing below
      const int newS = totalChildren + nKept;
      for (int i = lane; i < newS * 4; i += 32) S.cnt4[i] = 0;      // quadrant counters of the next pass
      if (lane == 0) { sh_res[0] = totalChildren; sh_res[1] = newS; sh_res[2] = nexp; }
    } else if (wid == 0) {
      // phase 2: processing rank = stable descending-size order (S.rank, above); E expandable nodes
      int E;
      {
        int c = 0;
        for (int p = lane; p < Sn; p += 32) c += cnt[p] > 1 ? 1 : 0;
        E = __reduce_add_sync(0xffffffffu, c);
      }
      // children per processing rank
      for (int p = lane; p < Sn; p += 32) {
        if (cnt[p] > 1) {
          const int* c4 = S.cnt4 + p * 4;
          S.childR[S.rank[p]] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
        }
      }
      __syncwarp();
      // keep a copy of per-rank child counts in keepPos (borrowed) before the scan overwrites them
      for (int t = lane; t < E; t += 32) S.keepPos[t] = S.childR[t];
      __syncwarp();
      const int allChildren = warp_exclusive_scan(S.childR, E);
      if (lane == 0) S.childR[E] = allChildren;
      __syncwarp();
      int Pn = E;
      if (phase == 2) {
        // first rank t whose split makes the list reach N:  Sn + (children up to and incl. t) - (t+1) >= N
        int stop = 0x7fffffff;
        for (int t = lane; t < E; t += 32) {
          const int incl = S.childR[t] + S.keepPos[t];
          if (Sn + incl - (t + 1) >= N) stop = min(stop, t);
        }
        stop = __reduce_min_sync(0xffffffffu, stop);
        if (stop != 0x7fffffff) Pn = stop + 1;
      }
      const int totalChildren = S.childR[Pn];
      __syncwarp();   // every lane has read childR[Pn] / keepPos before keepPos is rewritten below

      // positions of the children blocks and of the surviving old nodes
      int nexp = 0;
      for (int p = lane; p < Sn; p += 32) {
        const bool split = cnt[p] > 1 && S.rank[p] < Pn;
        int* cp = S.childPos + p * 4;
        if (split) {
          const int t = S.rank[p];
          const int* c4 = S.cnt4 + p * 4;
          const int nch = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
          int o = totalChildren - S.childR[t] - nch;      // blocks of later-processed parents come first
          const short4 b = box[p];
          const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
#pragma unroll
          for (int q = 3; q >= 0; --q) {
            if (c4[q] > 0) {
              cp[q] = o;
              nbox[o] = make_short4((q & 1) ? mx : b.x, (q & 2) ? my : b.y, (q & 1) ? b.z : mx, (q & 2) ? b.w : my);
              ncnt[o] = c4[q];
              nexp += c4[q] > 1 ? 1 : 0;
              ++o;
            } else cp[q] = -1;
          }
        }
      }
      nexp = __reduce_add_sync(0xffffffffu, nexp);
      // surviving nodes keep their relative order behind the children
      for (int p = lane; p < Sn; p += 32) S.keepPos[p] = (cnt[p] > 1 && S.rank[p] < Pn) ? 0 : 1;
      __syncwarp();
      const int nKept = warp_exclusive_scan(S.keepPos, Sn);
      for (int p = lane; p < Sn; p += 32) {
        const bool split = cnt[p] > 1 && S.rank[p] < Pn;
        if (!split) {
          const int o = totalChildren + S.keepPos[p];
          nbox[o] = box[p];
          ncnt[o] = cnt[p];
        } else {
          S.keepPos[p] = -1;
        }
      }
      const int newS = totalChildren + nKept;
      for (int i = lane; i < newS * 4; i += 32) S.cnt4[i] = 0;      // quadrant counters of the next pass
      if (lane == 0) { sh_res[0] = totalChildren; sh_res[1] = newS; sh_res[2] = nexp; }
    }
    __syncthreads();
    OCT_STAMP("nodes");
    const int totalChildren = sh_res[0], newS = sh_res[1], nToExpand = sh_res[2];
    for (int k0 = tid; k0 < n; k0 += kOctBatch * T) {
      int pp[kOctBatch]; uint32_t kk[kOctBatch];
#pragma unroll
      for (int j = 0; j < kOctBatch; ++j) {
        const int k = k0 + j * T;
        pp[j] = k < n ? (int)nodeOf[k] : -1;
        kk[j] = k < n ? keys[k] : 0u;
      }
#pragma unroll
      for (int j = 0; j < kOctBatch; ++j) {
        if (pp[j] < 0) continue;
        const int p = pp[j], kp = S.keepPos[p];
        nodeOf[k0 + j * T] = (uint16_t)(kp >= 0 ? totalChildren + kp : S.childPos[p * 4 + quadrant_of(box[p], key_x(kk[j]), key_y(kk[j]))]);
      }
    }
    OCT_STAMP("relabel");
    cur ^= 1;
    Scount = newS;
    // termination (ORBextractor.cpp:672-744)
    if (newS >= N || newS == Sn) finish = true;
    else if (phase == 1 && newS + 3 * nToExpand > N) phase = 2;
  }
  __syncthreads();      // the last relabel sweep read keepPos / childPos / box; the strongest-key pass below reuses nothing of them,
                        // but Scount-sized arrays are initialised by all threads

  // ---- strongest key per node, first wins ties (ORBextractor.cpp:747-766) ------------------------------
  // max (response, then FIRST index): response + 1 (9 bits) above the complemented key index.  Below 2^23 candidates the pair fits
  // 32 bits and the shared-memory atomicMax is one native instruction; the 64-bit form is a compare-and-swap loop (it was 12 %
  // of the kernel's stall samples) and remains only for levels with more candidates than that
  const bool narrow = n < (1 << 23);
  uint32_t* best32 = reinterpret_cast<uint32_t*>(S.best);
  for (int p = tid; p < Scount; p += T) { if (narrow) best32[p] = 0u; else S.best[p] = 0ull; }
  __syncthreads();
  for (int k0 = tid; k0 < n; k0 += kOctBatch * T) {
    int pp[kOctBatch]; uint32_t kk[kOctBatch];
#pragma unroll
    for (int j = 0; j < kOctBatch; ++j) {
      const int k = k0 + j * T;
      pp[j] = k < n ? (int)nodeOf[k] : -1;
      kk[j] = k < n ? keys[k] : 0u;
    }
#pragma unroll
    for (int j = 0; j < kOctBatch; ++j) {
      if (pp[j] < 0) continue;
      const int k = k0 + j * T;
      if (narrow) atomicMax(&best32[pp[j]], ((uint32_t)(key_s(kk[j]) + 1) << 23) | (0x7FFFFFu - (uint32_t)k));   // native 32-bit shared atomic
      else atomicMax(&S.best[pp[j]], ((unsigned long long)(key_s(kk[j]) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k));
    }
  }
  __syncthreads();
  for (int p = tid; p < Scount; p += T) {
    const uint32_t k = narrow ? 0x7FFFFFu - (best32[p] & 0x7FFFFFu) : 0xFFFFFFFFu - (uint32_t)(S.best[p] & 0xFFFFFFFFull);
    sel[p] = keys[k];
  }
  if (tid == 0) *selCount = Scount;
#if ORBX_OCT_TRACE
  OCT_STAMP("best");
  if (tid == 0 && l == 0 && f == 0) {
    printf("[octree] n=%d nodes=%d T=%d:", n, Scount, T);
    for (int i = 1; i < nst; ++i) printf(" %s %lld", st_n[i], st_t[i] - st_t[i - 1]);
    printf(" | total %lld cycles\n", st_t[nst - 1] - st_t[0]);
  }
#endif
}

// ======================================================================================================
// K5  cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101, 8.8 fixed point (ORBextractor.cpp:1093-1094;
// arithmetic: SURVEY App. A.3).  All levels of all frames of the chunk in one launch; one CTA = one
// 64x32 output tile, horizontal pass into shared u16, vertical pass out of it.
// ======================================================================================================
constexpr int kBlurTW = 64, kBlurTH = 26;        // output tile; TH+6 = 32 input rows = 16 row pairs x 16 column quads = 256 threads
constexpr int kBlurInW = kBlurTW + 8;            // input tile columns: image x0-4 .. x0+TW+3 (whole 32-bit words)

__device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

// 8.8 fixed-point kernel {18,34,48,56,48,34,18} packed for the integer dot-product instructions
constexpr uint32_t kBlurK0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);   // taps 0..3 (dp4a)
constexpr uint32_t kBlurK1 = 48u | (34u << 8) | (18u << 16);                 // taps 4..6 (dp4a)

__device__ __forceinline__ uint32_t blur_h(uint32_t w0, uint32_t w1) {       // w0 = p[x-3..x], w1 = p[x+1..x+4]
  return __dp4a(w1, kBlurK1, __dp4a(w0, kBlurK0, 0u));
}

__global__ void __launch_bounds__(256) blur_kernel(const __grid_constant__ Geom G, const Bufs B) {
  pdl_prologue();
  // One CTA = one 64-pixel-wide column strip of one level of one frame; it walks down the strip in 26-row tiles so that
  // the level lookup, pointers and alignment checks are paid once per strip instead of once per tile.
  __shared__ __align__(16) uint8_t tin[2][kBlurTH + 6][kBlurInW];           // double buffered: tile ty+1 streams in (cp.async) under tile ty's math
  __shared__ __align__(16) uint32_t hpair[(kBlurTH + 6) / 2][kBlurTW];      // rows 2p (low half) and 2p+1 (high half) of the horizontal pass
  const int tid = threadIdx.x, f = blockIdx.y;
  int l = 0;
  while (l + 1 < G.nlevels && (int)blockIdx.x >= G.L[l + 1].blurTile0) ++l;
  const LevelGeom& L = G.L[l];
  const int tx = blockIdx.x - L.blurTile0;
  const int x0 = tx * kBlurTW;
  const int W = L.w, H = L.h, bpitch = L.bpitch, tilesY = L.blurTilesY;
  int pitch;
  const uint8_t* src = level_ptr(G, B, l, f, pitch);
  const bool aligned = (((uintptr_t)src | (uintptr_t)pitch) & 3) == 0;
  uint8_t* dst = B.blur + L.blurOff + (size_t)f * H * bpitch;
  constexpr int WPR = kBlurInW / 4;   // 18 words per row
  const int lr = tid >> 3, wb = tid & 7;                    // load role: row, first word
  const int hq = tid & 15, hrp = tid >> 4;                  // horizontal/vertical role: column quad, row pair
  // per-thread column classification for the load (fixed for the whole strip)
  int gxw[3]; bool fastw[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int w = wb + 8 * k;
    gxw[k] = x0 - 4 + 4 * w;
    fastw[k] = aligned && w < WPR && gxw[k] >= 0 && gxw[k] + 3 < W;
  }
  // stage 32 x 72 input bytes of tile ty as 32-bit words (rows resolved with BORDER_REFLECT_101 once per thread): interior
  // words by cp.async straight into shared memory, words that touch the left/right border assembled from reflected bytes
  auto stage = [&](int buf, int ty) {
    const int sy = reflect101(min(ty * kBlurTH + lr - 3, H + 2), H);
    const uint8_t* row = src + (size_t)sy * pitch;
    uint32_t* trow = reinterpret_cast<uint32_t*>(&tin[buf][lr][0]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int w = wb + 8 * k;
      if (w >= WPR) continue;
      if (fastw[k]) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(trow + w)), "l"(row + gxw[k]) : "memory");
      } else {
        uint32_t v = 0;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int sx = reflect101(min(max(gxw[k] + bb, -3), W + 2), W);
          v |= (uint32_t)__ldg(row + sx) << (8 * bb);
        }
        trow[w] = v;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  for (int ty = 0; ty < tilesY; ++ty) {
    const int y0 = ty * kBlurTH, buf = ty & 1;
    if (ty + 1 < tilesY) {
      stage(buf ^ 1, ty + 1);     // its previous contents were last read before the second barrier of iteration ty-1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // ---- horizontal pass: one thread = 4 adjacent columns of 2 adjacent rows; exact in u16 (max 255*256)
    {
      uint32_t o[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(&tin[buf][2 * hrp + rr][4 * hq]);
        const uint32_t a = wp[0], b = wp[1], c = wp[2];   // bytes j = 0..11 <-> image x0-4+4q+j; output k uses bytes k+1..k+7
        o[rr][0] = blur_h(__funnelshift_r(a, b, 8), __funnelshift_r(b, c, 8));
        o[rr][1] = blur_h(__funnelshift_r(a, b, 16), __funnelshift_r(b, c, 16));
        o[rr][2] = blur_h(__funnelshift_r(a, b, 24), __funnelshift_r(b, c, 24));
        o[rr][3] = blur_h(b, c);
      }
      *reinterpret_cast<uint4*>(&hpair[hrp][4 * hq]) = make_uint4(o[0][0] | (o[1][0] << 16), o[0][1] | (o[1][1] << 16),
                                                                 o[0][2] | (o[1][2] << 16), o[0][3] | (o[1][3] << 16));
    }
    __syncthreads();
    // ---- vertical pass: one thread = 4 columns x 2 rows (2g, 2g+1); both rows read the same four row pairs
    if (hrp < kBlurTH / 2) {
      const int g = hrp;
      uint4 P[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) P[p] = *reinterpret_cast<const uint4*>(&hpair[g + p][4 * hq]);
      // even output row 2g: taps on rows 2g..2g+6 -> pairs (k0,k1)(k2,k3)(k4,k5)(k6,-); odd row 2g+1: (-,k0)(k1,k2)(k3,k4)(k5,k6)
      constexpr uint32_t E0 = 18u | (34u << 8), E1 = 48u | (56u << 8), E2 = 48u | (34u << 8), E3 = 18u;
      constexpr uint32_t O0 = 18u << 8, O1 = 34u | (48u << 8), O2 = 56u | (48u << 8), O3 = 34u | (18u << 8);
      uint32_t ev = 0, od = 0;
      const uint32_t* c0 = reinterpret_cast<const uint32_t*>(&P[0]);
      const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&P[1]);
      const uint32_t* c2 = reinterpret_cast<const uint32_t*>(&P[2]);
      const uint32_t* c3 = reinterpret_cast<const uint32_t*>(&P[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t e = __dp2a_lo(c0[k], E0, 32768u);
        e = __dp2a_lo(c1[k], E1, e); e = __dp2a_lo(c2[k], E2, e); e = __dp2a_lo(c3[k], E3, e);
        uint32_t o = __dp2a_lo(c0[k], O0, 32768u);
        o = __dp2a_lo(c1[k], O1, o); o = __dp2a_lo(c2[k], O2, o); o = __dp2a_lo(c3[k], O3, o);
        ev |= (e >> 16) << (8 * k);
        od |= (o >> 16) << (8 * k);
      }
      const int x = x0 + 4 * hq;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int y = y0 + 2 * g + rr;
        const uint32_t packed = rr ? od : ev;
        if (y >= H || x >= W) continue;
        uint8_t* d = dst + (size_t)y * bpitch + x;
        if (x + 3 < W) *reinterpret_cast<uint32_t*>(d) = packed;     // bpitch is a multiple of 64, x of 4
        else for (int k = 0; x + k < W; ++k) d[k] = (uint8_t)(packed >> (8 * k));
      }
    }
    // the next iteration's horizontal pass writes hpair only after its first barrier, which every thread reaches after
    // finishing this vertical pass
  }
}

// ---- register-blocked variant (the default): no shared memory, no barriers.  One thread owns 4 adjacent columns and
// walks down kBwRows output rows: per input row pair it does the horizontal pass once (3 aligned words -> 4 pixels x 2
// rows, dp4a) and keeps the last four pair results (u16x2 per column) in registers; every step then emits two output
// rows with dp2a.  All index arithmetic, table lookups and barriers of the tiled kernel are paid once per 26 rows
// instead of once per row pair (31 -> ~12 thread instructions per pixel).  Same arithmetic, bit for bit.
constexpr int kBwRows = 26;                        // output rows per thread (+6 halo rows = 16 input row pairs)
constexpr int kBwPairs = (kBwRows + 6) / 2;
constexpr int kBwSegs = 8;                         // vertical segments per CTA: 4 warps x 2 half-warps
constexpr int kBwTileH = kBwRows * kBwSegs;        // 208 output rows per CTA, 64 columns

// Row fetch of the register-blocked blur: the 12 bytes x = 4q-4 .. 4q+7 of one (already row-reflected) image row as three
// words.  MODE 0: interior quads, three aligned loads.  MODE 1: quads whose window crosses the left/right image border:
// three loads at clamped word indices + PRMT gathers with per-thread selectors (computed once per thread, the border
// pattern is the same for every row) that apply BORDER_REFLECT_101.  MODE 2: rows that are not 4-byte aligned (a caller
// image with an odd stride): reflected byte loads, out of line.
struct BlurEdge { uint32_t s1[3], s2[3], m[3]; int w0, w2; };

__device__ __forceinline__ BlurEdge blur_edge_selectors(int q, int W) {
  BlurEdge E;
  const int lastw = (W - 1) >> 2;
  E.w0 = max(q - 1, 0); E.w2 = min(q + 1, lastw);
  const int wk[3] = {E.w0, q, E.w2};
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    uint32_t s1 = 0, s2 = 0, m = 0;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int x = reflect101(min(max(4 * q - 4 + 4 * v + bb, -3), W + 2), W);
      int idx = 0;                                   // position in the 12-byte pool (words w0, q, w2); 0 if never needed
      if ((x >> 2) == wk[0]) idx = (x & 3);
      else if ((x >> 2) == wk[1]) idx = 4 + (x & 3);
      else if ((x >> 2) == wk[2]) idx = 8 + (x & 3);
      if (idx < 8) { s1 |= (uint32_t)idx << (4 * bb); m |= 0xFFu << (8 * bb); }
      else s2 |= (uint32_t)(idx - 4) << (4 * bb);
    }
    E.s1[v] = s1; E.s2[v] = s2; E.m[v] = m;
  }
  return E;
}

__device__ __noinline__ void blur_fetch_bytes(const uint8_t* __restrict__ row, int q, int W, uint32_t& a, uint32_t& b, uint32_t& c) {
  uint32_t v[3] = {0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int sx = reflect101(min(max(4 * q - 4 + 4 * k + bb, -3), W + 2), W);
      v[k] |= (uint32_t)__ldg(row + sx) << (8 * bb);
    }
  a = v[0]; b = v[1]; c = v[2];
}

template <int MODE>
__device__ __forceinline__ void blur_fetch(const uint8_t* __restrict__ row, int q, int W, const BlurEdge& E, uint32_t& a, uint32_t& b, uint32_t& c) {
  if (MODE == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + q;
    a = __ldg(w - 1); b = __ldg(w); c = __ldg(w + 1);
  } else if (MODE == 1) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row);
    const uint32_t p0 = __ldg(w + E.w0), p1 = __ldg(w + q), p2 = __ldg(w + E.w2);
    a = (__byte_perm(p0, p1, E.s1[0]) & E.m[0]) | (__byte_perm(p1, p2, E.s2[0]) & ~E.m[0]);
    b = (__byte_perm(p0, p1, E.s1[1]) & E.m[1]) | (__byte_perm(p1, p2, E.s2[1]) & ~E.m[1]);
    c = (__byte_perm(p0, p1, E.s1[2]) & E.m[2]) | (__byte_perm(p1, p2, E.s2[2]) & ~E.m[2]);
  } else {
    blur_fetch_bytes(row, q, W, a, b, c);
  }
}

// VEDGE = false: every input row y0-3 .. y0+kBwRows+2 of the thread lies inside the image, rows are reached by pointer
// increments; VEDGE = true: rows are resolved with BORDER_REFLECT_101 one by one (top/bottom segments).
template <int MODE, bool VEDGE>
__device__ __forceinline__ void blur_walk_body(const uint8_t* __restrict__ src, int pitch, int W, int H, int q, int y0,
                                               uint8_t* __restrict__ dst, int bpitch) {
  constexpr uint32_t E0 = 18u | (34u << 8), E1 = 48u | (56u << 8), E2 = 48u | (34u << 8), E3 = 18u;
  constexpr uint32_t O0 = 18u << 8, O1 = 34u | (48u << 8), O2 = 56u | (48u << 8), O3 = 34u | (18u << 8);
  BlurEdge E;
  if (MODE == 1) E = blur_edge_selectors(q, W);
  const uint8_t* rowp = src + (ptrdiff_t)(y0 - 3) * pitch;          // VEDGE == false only
  const int rowsLeft = H - y0;                                      // output row y0 + k exists iff k < rowsLeft
  uint32_t ring[4][4];                                             // [pair & 3][column]: row 2p low half, row 2p+1 high half
  static_assert(kBwPairs % 4 == 0, "ring period");
#pragma unroll 1
  for (int pg = 0; pg < kBwPairs / 4; ++pg) {
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const int p = pg * 4 + pi;
      uint32_t o[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const uint8_t* row;
        if (VEDGE) row = src + (size_t)reflect101(min(y0 - 3 + 2 * p + rr, H + 2), H) * pitch;
        else row = rowp + (2 * pi + rr) * pitch;
        uint32_t a, b, c;
        blur_fetch<MODE>(row, q, W, E, a, b, c);
        o[rr][0] = blur_h(__funnelshift_r(a, b, 8), __funnelshift_r(b, c, 8));
        o[rr][1] = blur_h(__funnelshift_r(a, b, 16), __funnelshift_r(b, c, 16));
        o[rr][2] = blur_h(__funnelshift_r(a, b, 24), __funnelshift_r(b, c, 24));
        o[rr][3] = blur_h(b, c);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) ring[pi][k] = __byte_perm(o[0][k], o[1][k], 0x5410);   // both fit 16 bits
      if (pg > 0 || pi == 3) {
        // output rows 2g (even taps) and 2g+1 (odd taps) of the thread, g = p - 3; dst already points at row 2*(4*pg - 3)
        uint32_t e[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c0 = ring[(pi + 1) & 3][k], c1 = ring[(pi + 2) & 3][k], c2 = ring[(pi + 3) & 3][k], c3 = ring[pi][k];
          e[k] = __dp2a_lo(c3, E3, __dp2a_lo(c2, E2, __dp2a_lo(c1, E1, __dp2a_lo(c0, E0, 32768u))));
          d[k] = __dp2a_lo(c3, O3, __dp2a_lo(c2, O2, __dp2a_lo(c1, O1, __dp2a_lo(c0, O0, 32768u))));
        }
        // result = bits 16..23 of each accumulator (<= 255): gather byte 2 of four registers into one word
        const uint32_t ev = __byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410);
        const uint32_t od = __byte_perm(__byte_perm(d[0], d[1], 0x0062), __byte_perm(d[2], d[3], 0x0062), 0x5410);
        const int k0 = 2 * (p - 3);
        uint8_t* d0 = dst + (ptrdiff_t)(2 * pi) * bpitch;           // bpitch is a multiple of 64: columns >= W are padding
        if (VEDGE) {
          if (k0 < rowsLeft) *reinterpret_cast<uint32_t*>(d0) = ev;
          if (k0 + 1 < rowsLeft) *reinterpret_cast<uint32_t*>(d0 + bpitch) = od;
        } else {
          *reinterpret_cast<uint32_t*>(d0) = ev;
          *reinterpret_cast<uint32_t*>(d0 + bpitch) = od;
        }
      }
    }
    rowp += 8 * pitch;
    dst += (ptrdiff_t)8 * bpitch;
  }
}

__global__ void __launch_bounds__(128, 8) blur_walk_kernel(const __grid_constant__ Geom G, const Bufs B) {
  pdl_prologue();
  // frames are the FAST grid dimension: CTAs that are resident together work on the same tile of different frames, i.e. run
  // the same specialisation of the walk (interior / border variants), which keeps the instruction cache warm
  const int tid = threadIdx.x, f = blockIdx.x;
  int l = 0;
  while (l + 1 < G.nlevels && (int)blockIdx.y >= G.L[l + 1].bwTile0) ++l;
  const LevelGeom& L = G.L[l];
  const int t = blockIdx.y - L.bwTile0;
  const int ty = t / L.bwTilesX, tx = t - ty * L.bwTilesX;
  const int W = L.w, H = L.h, bpitch = L.bpitch;
  const int q = tx * 16 + (tid & 15);                              // column quad
  const int y0 = ty * kBwTileH + (tid >> 4) * kBwRows;             // first output row of this thread
  const bool live = 4 * q < W && y0 < H;
  int pitch;
  const uint8_t* src = level_ptr(G, B, l, f, pitch);
  // dst points at output row y0 - 6: the walk advances it by 8 rows per group of 4 pairs and the first group only emits pair 3
  uint8_t* dst = B.blur + L.blurOff + (size_t)f * H * bpitch + (ptrdiff_t)(y0 - 6) * bpitch + 4 * q;
  const bool unaligned = (((uintptr_t)src | (uintptr_t)pitch) & 3) != 0;                 // uniform per level
  const bool edge = __any_sync(0xffffffffu, live && (q == 0 || 4 * q + 7 >= W));          // uniform per warp
  const bool vedge = __any_sync(0xffffffffu, live && (y0 < 3 || y0 + kBwRows + 3 > H));   // uniform per warp
  if (!live) return;
  if (unaligned) blur_walk_body<2, true>(src, pitch, W, H, q, y0, dst, bpitch);
  else if (edge) { if (vedge) blur_walk_body<1, true>(src, pitch, W, H, q, y0, dst, bpitch); else blur_walk_body<1, false>(src, pitch, W, H, q, y0, dst, bpitch); }
  else { if (vedge) blur_walk_body<0, true>(src, pitch, W, H, q, y0, dst, bpitch); else blur_walk_body<0, false>(src, pitch, W, H, q, y0, dst, bpitch); }
}

// ---- TMA-staged variant (the default when every level can be described) ------------------------------------------
// Same register-blocked walk, but the 64 x 208 tile of a CTA (+3 halo rows/columns, box 96 x 214 bytes) arrives by ONE
// cp.async.bulk.tensor.  TMA zero-fills outside the image, so BORDER_REFLECT_101 is applied IN shared memory afterwards
// (3 halo rows, then 3 halo columns per row: a few hundred byte copies, border tiles only).  Every thread then runs the
// interior code path on LDS: one specialisation instead of five (instruction-cache misses were the top stall of
// blur_walk_kernel) and no dependent global-load latency in the walk.
static_assert(kBtBoxH == kBwTileH + 6 && kBtBoxW == 64 + 32, "blur box");
#ifndef ORBX_BLUR_MINB
#define ORBX_BLUR_MINB 8
#endif
__global__ void __launch_bounds__(128, ORBX_BLUR_MINB) blur_tma_kernel(const __grid_constant__ Geom G, const Bufs B, const __grid_constant__ TmaSet TM) {
  pdl_prologue();
  __shared__ __align__(128) uint8_t tile[kBtBoxH * kBtBoxW];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, f = blockIdx.x;
  int l = 0;
  while (l + 1 < G.nlevels && (int)blockIdx.y >= G.L[l + 1].bwTile0) ++l;
  const LevelGeom& L = G.L[l];
  const int t = blockIdx.y - L.bwTile0;
  const int ty = t / L.bwTilesX, tx = t - ty * L.bwTilesX;
  const int W = L.w, H = L.h, bpitch = L.bpitch;
  const int bx0 = tx * 64 - 16, by0 = ty * kBwTileH - 3;           // image coordinates of tile[0][0]
  const uint32_t barA = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"((uint32_t)(kBtBoxW * kBtBoxH)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(tile)),
        "l"(reinterpret_cast<uint64_t>(TM.map + 4 * kMaxLevels + l)), "r"(bx0), "r"(by0), "r"((l == 0 ? TM.frame0 : 0) + f), "r"(barA)
        : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "BW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra BD_%=;\n\t"
      "bra BW_%=;\n\t"
      "BD_%=:\n\t}" ::"r"(barA) : "memory");
  // ---- BORDER_REFLECT_101 inside the tile (CTA-uniform conditions) ----------------------------------------------
  const bool top = by0 < 0, bottom = by0 + kBtBoxH > H, left = bx0 < 0, right = bx0 + kBtBoxW > W;
  if (top | bottom) {
    if (top)                                                       // image rows -3..-1 <- rows 3..1
      for (int i = tid; i < 3 * kBtBoxW; i += 128) {
        const int k = i / kBtBoxW + 1, x = i - (k - 1) * kBtBoxW;
        tile[(-k - by0) * kBtBoxW + x] = tile[(k - by0) * kBtBoxW + x];
      }
    if (bottom)                                                    // image rows H..H+2 <- rows H-2..H-4
      for (int i = tid; i < 3 * kBtBoxW; i += 128) {
        const int k = i / kBtBoxW, x = i - k * kBtBoxW;
        const int r = H + k - by0;
        if (r < kBtBoxH) tile[r * kBtBoxW + x] = tile[(H - 2 - k - by0) * kBtBoxW + x];
      }
    __syncthreads();
  }
  if (left | right) {
    for (int r = tid; r < kBtBoxH; r += 128) {
      uint8_t* row = tile + r * kBtBoxW;
      if (left) { row[-1 - bx0] = row[1 - bx0]; row[-2 - bx0] = row[2 - bx0]; row[-3 - bx0] = row[3 - bx0]; }
      if (right) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (W + k - bx0 < kBtBoxW) row[W + k - bx0] = row[W - 2 - k - bx0];
      }
    }
    __syncthreads();
  }

  const int q = tx * 16 + (tid & 15);                              // column quad
  const int seg = tid >> 4;
  const int y0 = ty * kBwTileH + seg * kBwRows;                    // first output row of this thread
  if (!(4 * q < W && y0 < H)) return;
  constexpr uint32_t E0 = 18u | (34u << 8), E1 = 48u | (56u << 8), E2 = 48u | (34u << 8), E3 = 18u;
  constexpr uint32_t O0 = 18u << 8, O1 = 34u | (48u << 8), O2 = 56u | (48u << 8), O3 = 34u | (18u << 8);
  // tile row of image row y0-3 is seg*kBwRows; word (tid&15) + 4 of a tile row holds image bytes 4q..4q+3
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(tile + seg * kBwRows * kBtBoxW) + 4 + (tid & 15);
  uint8_t* dst = B.blur + L.blurOff + (size_t)f * H * bpitch + (ptrdiff_t)(y0 - 6) * bpitch + 4 * q;
  const int rowsLeft = H - y0;                                     // output row y0 + k exists iff k < rowsLeft
  uint32_t ring[4][4];                                             // [pair & 3][column]: row 2p low half, row 2p+1 high half
#pragma unroll 1
  for (int pg = 0; pg < kBwPairs / 4; ++pg) {
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const int p = pg * 4 + pi;
      uint32_t o[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const uint32_t* w = wp + (2 * pi + rr) * (kBtBoxW / 4);
        const uint32_t a = w[-1], b = w[0], c = w[1];
        o[rr][0] = blur_h(__funnelshift_r(a, b, 8), __funnelshift_r(b, c, 8));
        o[rr][1] = blur_h(__funnelshift_r(a, b, 16), __funnelshift_r(b, c, 16));
        o[rr][2] = blur_h(__funnelshift_r(a, b, 24), __funnelshift_r(b, c, 24));
        o[rr][3] = blur_h(b, c);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) ring[pi][k] = __byte_perm(o[0][k], o[1][k], 0x5410);   // both fit 16 bits
      if (pg > 0 || pi == 3) {
        uint32_t e[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c0 = ring[(pi + 1) & 3][k], c1 = ring[(pi + 2) & 3][k], c2 = ring[(pi + 3) & 3][k], c3 = ring[pi][k];
          e[k] = __dp2a_lo(c3, E3, __dp2a_lo(c2, E2, __dp2a_lo(c1, E1, __dp2a_lo(c0, E0, 32768u))));
          d[k] = __dp2a_lo(c3, O3, __dp2a_lo(c2, O2, __dp2a_lo(c1, O1, __dp2a_lo(c0, O0, 32768u))));
        }
        const uint32_t ev = __byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410);
        const uint32_t od = __byte_perm(__byte_perm(d[0], d[1], 0x0062), __byte_perm(d[2], d[3], 0x0062), 0x5410);
        const int k0 = 2 * (p - 3);
        uint8_t* d0 = dst + (ptrdiff_t)(2 * pi) * bpitch;           // bpitch is a multiple of 64: columns >= W are padding
        if (k0 < rowsLeft) *reinterpret_cast<uint32_t*>(d0) = ev;
        if (k0 + 1 < rowsLeft) *reinterpret_cast<uint32_t*>(d0 + bpitch) = od;
      }
    }
    wp += 8 * (kBtBoxW / 4);
    dst += (ptrdiff_t)8 * bpitch;
  }
}

// ======================================================================================================
// K4+K6  IC_Angle (ORBextractor.cpp:79-107) + rotated BRIEF (ORBextractor.cpp:110-151) + output assembly
// (ORBextractor.cpp:845-855, 1085-1111).  One warp per selected keypoint: lanes = patch rows for the
// moments, lanes = descriptor bytes for the 256 comparisons.
// ======================================================================================================
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {   // cv::fastAtan2, SURVEY App. A.4
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// glibc 2.39 sinf/cosf (the ARM optimized-routines sincosf algorithm) evaluated in IEEE fp64 without FMA and
// rounded once to float: bit-identical to libm on [0, 2*pi] (checked exhaustively by
// tests/tools/check_sincos.cpp).  CUDA's own sinf/cosf must NOT be used here.
__device__ __forceinline__ void glibc_sincosf(float y, float& s_out, float& c_out) {
  const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
               c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  double x = (double)y;
  int n = 0;
  double k = 1.0;                 // the n&2 half of glibc's table negates the cosine coefficients
  const float ay = fabsf(y);
  if (ay < 0x1.921FB6p-1f) {
    if (ay < 0x1p-12f) { s_out = y; c_out = 1.0f; return; }
  } else {
    const double r = __dmul_rn(x, hpi_inv);
    n = ((int)r + 0x800000) >> 24;
    x = __dadd_rn(__dmul_rn(-(double)n, hpi), x);
    if (n & 2) k = -1.0;
  }
  const double x2 = __dmul_rn(x, x);
  const int q = n & 3;
  if (q == 1 || q == 2) x = -x;   // sign table {+1,-1,-1,+1}
  const double x4 = __dmul_rn(x2, x2), x3 = __dmul_rn(x2, x);
  const double cc2 = __dadd_rn(__dmul_rn(x2, k * c4), k * c3);
  const double ss1 = __dadd_rn(__dmul_rn(x2, s3), s2);
  const double cc1 = __dadd_rn(__dmul_rn(x2, k * c1), k * c0);
  const double x5 = __dmul_rn(x3, x2), x6 = __dmul_rn(x4, x2);
  const double sv = __dadd_rn(__dmul_rn(x3, s1), x);
  const double cv = __dadd_rn(__dmul_rn(x4, k * c2), cc1);
  const float sinv = (float)__dadd_rn(__dmul_rn(x5, ss1), sv);
  const float cosv = (float)__dadd_rn(__dmul_rn(x6, cc2), cv);
  if (n & 1) { s_out = cosv; c_out = sinv; } else { s_out = sinv; c_out = cosv; }
}

constexpr int kPatchRowsU = 31, kPatchPitchU = 44;   // unblurred 31x31 patch rows, 44-byte pitch (11 words: conflict-free rows)
constexpr int kPatchRowsB = 37, kPatchPitchB = 56;   // blurred 37x37 patch (+-18), 56-byte pitch (14 words, 8-byte aligned rows)
constexpr int kPatchBytes = kPatchRowsU * kPatchPitchU + kPatchRowsB * kPatchPitchB;   // 3436 B per warp

// Stage rows [y-R, y+R] x columns [x-R, x+R] of an image into shared memory (one warp).  Returns the column offset `ox`
// such that patch(r, c) = dst[r*P + ox + c], c = 0 <-> image column x-R.  When the image is CW-byte aligned (CW = 4 or
// 8) the rows are fetched as aligned CW-byte pieces with cp.async (LDGSTS): no register round trip, so all copies of
// both patches are in flight at once and the warp waits once (cp_async_wait_all) instead of once per load.  Lanes
// map to (row-in-group, piece) and step down by whole row groups: no per-copy index arithmetic.
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int R, int P, int CW>
__device__ __forceinline__ int stage_patch(uint8_t* dst, const uint8_t* img, int pitch, int x, int y, int lane) {
  constexpr int ROWS = 2 * R + 1;
  const uint8_t* base = img + (size_t)(y - R) * pitch + (x - R);
  if ((((uintptr_t)img | (uintptr_t)pitch) & (CW - 1)) == 0) {
    const int ox = (int)((uintptr_t)base & (CW - 1));
    constexpr int PPR = (ROWS + CW - 1 + CW - 1) / CW;          // pieces per row covering ox + ROWS bytes
    constexpr int RPG = 32 / PPR;                               // rows per group of lanes
    static_assert(PPR * CW <= P && P % CW == 0, "patch pitch");
    const int rs = lane / PPR, w = lane - rs * PPR;
    if (rs < RPG) {
      const uint8_t* src = base - ox + (size_t)rs * pitch + CW * w;
      uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + rs * P + CW * w);
#pragma unroll
      for (int r = rs; r < ROWS + RPG - 1; r += RPG) {
        if (r < ROWS) asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(CW) : "memory");
        src += (size_t)RPG * pitch;
        d += RPG * P;
      }
    }
    return ox;
  }
  for (int i = lane; i < ROWS * ROWS; i += 32) {
    const int r = i / ROWS, c = i - r * ROWS;
    dst[r * P + c] = __ldg(base + (size_t)r * pitch + c);
  }
  return 0;
}

__global__ void __launch_bounds__(256) orient_desc_kernel(const __grid_constant__ Geom G, const Bufs B, orbx_keypoint* __restrict__ kps_out,
                                                          uint8_t* __restrict__ desc_out, int cap,
                                                          int32_t* __restrict__ counts_out, int frame0) {
  pdl_prologue();
  __shared__ __align__(16) uint8_t patches[8][kPatchBytes + 4];   // per warp: blurred patch (8-byte aligned rows) then the unblurred one
  static_assert((kPatchBytes + 4) % 8 == 0 && (kPatchRowsB * kPatchPitchB) % 8 == 0, "patch alignment");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slotIdx = blockIdx.x * 8 + warp;   // index into the per-frame selected array
  const int f = blockIdx.y;
  if (slotIdx >= G.selPerFrame) return;
  int l = 0;
  while (l + 1 < G.nlevels && slotIdx >= G.L[l + 1].selBase) ++l;
  const LevelGeom& L = G.L[l];
  const int i = slotIdx - L.selBase;
  // lane q fetches the count of level q while the (possibly unused) key of this slot is already on its way: one global
  // round trip instead of three dependent ones before the patches can be requested
  const int cl = lane < G.nlevels ? __ldg(B.selCount + (size_t)f * G.nlevels + lane) : 0;
  const uint32_t key = __ldg(B.sel + (size_t)f * G.selPerFrame + slotIdx);
  const int total = __reduce_add_sync(0xffffffffu, cl);
  const int before = __reduce_add_sync(0xffffffffu, lane < l ? cl : 0);
  const int cntL = __shfl_sync(0xffffffffu, cl, l);
  if (slotIdx == 0 && lane == 0) counts_out[frame0 + f] = total;
  if (i >= cntL) return;
  const int o = before + i;
  if (o >= cap) return;   // caller buffer smaller than the keypoint count: the count is still reported

  const int x = key_x(key) + kMinBorder, y = key_y(key) + kMinBorder;   // (ORBextractor.cpp:851-852)
  int pitch;
  const uint8_t* img = level_ptr(G, B, l, f, pitch);
  uint8_t* pb = patches[warp];
  uint8_t* pu = pb + kPatchRowsB * kPatchPitchB;
  const int oxu = stage_patch<15, kPatchPitchU, 4>(pu, img, pitch, x, y, lane);
  const int oxb = stage_patch<18, kPatchPitchB, 8>(pb, B.blur + L.blurOff + (size_t)f * L.h * L.bpitch, L.bpitch, x, y, lane);
  cp_async_wait_all();
  __syncwarp();

  // ---- moments over the radius-15 disc: lane <-> row v = lane-15 --------------------------------------
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int v = lane - 15;
    const int d = __ldg(B.umax + abs(v));
    const uint8_t* row = pu + lane * kPatchPitchU + oxu + 15;
    int sum = 0;
    for (int u = -d; u <= d; ++u) {
      const int p = row[u];
      m10 += u * p;
      sum += p;
    }
    m01 = v * sum;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- rBRIEF on the blurred patch ---------------------------------------------------------------------
  const float factorPI = (float)(3.1415926535897932384626433832795 / (double)180.f);
  float a, b;
  glibc_sincosf(__fmul_rn(angle, factorPI), b, a);
  const uint8_t* bl = pb + 18 * kPatchPitchB + oxb + 18;
  // pattern is stored transposed, [point-in-byte 0..15][byte/lane 0..31], so a warp reads 256 contiguous bytes
  const float2* pat = B.pattern + lane;
  int val = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 p0 = __ldg(pat + 32 * (2 * j)), p1 = __ldg(pat + 32 * (2 * j + 1));
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b)));
    const int t0 = bl[r0 * kPatchPitchB + c0], t1 = bl[r1 * kPatchPitchB + c1];
    val |= (t0 < t1) << j;
  }
  desc_out[((size_t)(frame0 + f) * cap + o) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    orbx_keypoint k;
    k.x = (float)x; k.y = (float)y;
    if (l != 0) { k.x = __fmul_rn(k.x, L.scale); k.y = __fmul_rn(k.y, L.scale); }   // (:1102-1108)
    k.size = L.kpSize; k.angle = angle; k.response = (float)key_s(key); k.octave = l; k.class_id = -1;
    kps_out[(size_t)(frame0 + f) * cap + o] = k;
  }
}

// ---- TMA-fed, software-pipelined variant (the default when every level has patch descriptors) -----------------------
// The kernel above is bound by the LSU data pipe (77 % of peak: byte-wise smem reads for the moments, LDGSTS staging,
// pattern loads) with the DRAM latency of the patches exposed behind two dependent global loads.  Here
//  * both patches of a keypoint arrive by ONE cp.async.bulk.tensor each (UTMALDG; no LSU wavefronts, no register
//    traffic) into a two-stage per-warp ring guarded by mbarriers, and a warp walks kOdK consecutive slots so the
//    patches of keypoint k+1 are in flight while keypoint k is computed;
//  * the moments read a row as three 16-byte words per lane and use IDP (dp4a, u8 x s8) against the column weights
//    after masking the bytes outside the radius-15 disc: 12 LDS wavefronts + 16 IDP instead of 31 LDS.U8 + 62 IMAD/IADD;
//  * the 16 pattern points of a lane stay in registers across the warp's keypoints;
//  * cvRound = add-magic-number rounding on the FMA pipe (exact round-half-even for |v| < 2^22) instead of F2I.
// Arithmetic and output are identical to orient_desc_kernel.
#ifndef ORBX_OD_K
#define ORBX_OD_K 8
#endif
#ifndef ORBX_OD_WARPS
#define ORBX_OD_WARPS 4
#endif
constexpr int kOdK = ORBX_OD_K;         // consecutive selected-keypoint slots per warp
constexpr int kOdWarps = ORBX_OD_WARPS; // warps per CTA
constexpr int kOdStage = 3968;          // bytes per ring stage: 48x31 patch at 0, 64x37 patch at 1536 (both 128-byte aligned)
constexpr int kOdBOff = 1536;
constexpr uint32_t kOdTx = kOdUW * kOdUH + kOdBW * kOdBH;

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // sum of (unsigned byte of a) * (signed byte of b) + c
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__host__ __device__ constexpr uint32_t od_wu(int j) {   // column weights u = -15 + 4j + b of aligned patch word j as packed s8
  return (uint32_t)((-15 + 4 * j) & 0xFF) | ((uint32_t)((-14 + 4 * j) & 0xFF) << 8) | ((uint32_t)((-13 + 4 * j) & 0xFF) << 16) |
         ((uint32_t)((-12 + 4 * j) & 0xFF) << 24);
}
template <int WS>
__device__ __forceinline__ void od_row_moments(const uint4* row, int bsh, const uint32_t* mk, int& m10, int& sum) {
  const uint4 q0 = row[0], q1 = row[1], q2 = row[2];
  const uint32_t W[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = __funnelshift_r(W[WS + j], W[WS + j + 1], bsh) & mk[j];
    sum = dp4a_us(a, 0x01010101u, sum);
    m10 = dp4a_us(a, od_wu(j), m10);
  }
}
__device__ __forceinline__ int od_round(float v) {                        // cvRound(v) + 0x4B400000
  return __float_as_int(__fadd_rn(v, 12582912.f));
}

#ifndef ORBX_OD_MINB
#define ORBX_OD_MINB 7          // 72 registers, 7 CTAs per SM: 2.53 -> 2.47 ms per 4096 frames (6: 2.50, 8: 2.52)
#endif
__global__ void __launch_bounds__(32 * kOdWarps, ORBX_OD_MINB) orient_desc_tma_kernel(const __grid_constant__ Geom G, const Bufs B, const __grid_constant__ TmaSet TM,
                                                                        orbx_keypoint* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
                                                                        int cap, int32_t* __restrict__ counts_out, int frame0, int K) {
  pdl_prologue();
  __shared__ __align__(128) uint8_t ring[kOdWarps][2][kOdStage];
  __shared__ __align__(8) uint64_t bars[kOdWarps][2];
  __shared__ uint32_t momMask[16 * 9];       // [|v|][word j] byte mask of |u| <= umax[|v|]; row stride 9: conflict-free across lanes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if ORBX_OD_FF
  const int f = blockIdx.x, bx = blockIdx.y;
#else
  const int f = blockIdx.y, bx = blockIdx.x;
#endif
  for (int i = tid; i < 16 * 9; i += 32 * kOdWarps) {
    const int av = i / 9, j = i - av * 9, d = __ldg(B.umax + av);
    uint32_t m = 0;
    for (int b = 0; b < 4; ++b) {
      const int u = -15 + 4 * j + b;
      if (j < 8 && u >= -d && u <= d) m |= 0xFFu << (8 * b);
    }
    momMask[i] = m;
  }
  if (tid < 2 * kOdWarps) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[tid >> 1][tid & 1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                            // the only CTA-wide barrier: the warps are independent from here on

  const int slot0 = (bx * kOdWarps + warp) * K;             // K consecutive slots per warp (kOdK; 2 for latency-bound launches)
  if (slot0 >= G.selPerFrame) return;
  // lane q < nlevels: count of level q; lane k < kOdK: key / level / output position of slot slot0 + k (one global round trip)
  const int cl = lane < G.nlevels ? __ldg(B.selCount + (size_t)f * G.nlevels + lane) : 0;
  const int myslot = min(slot0 + lane, G.selPerFrame - 1);
  const uint32_t mykey = lane < K ? __ldg(B.sel + (size_t)f * G.selPerFrame + myslot) : 0u;
  int myl = 0;
  while (myl + 1 < G.nlevels && myslot >= G.L[myl + 1].selBase) ++myl;
  int before = 0, cntL = 0, total = 0;
  for (int q = 0; q < G.nlevels; ++q) {
    const int c = __shfl_sync(0xffffffffu, cl, q);
    total += c;
    if (q < myl) before += c;
    if (q == myl) cntL = c;
  }
  if (bx == 0 && tid == 0) counts_out[frame0 + f] = total;
  const int myi = myslot - G.L[myl].selBase, myo = before + myi;
  // caller buffer smaller than the keypoint count: the count is still reported
  unsigned todo = __ballot_sync(0xffffffffu, lane < K && slot0 + lane < G.selPerFrame && myi < cntL && myo < cap);
  if (todo == 0) return;

  float2 pat[16];                             // pattern stored transposed [point-in-byte][lane]: coalesced, loaded once per warp
#pragma unroll
  for (int k = 0; k < 16; ++k) pat[k] = __ldg(B.pattern + 32 * k + lane);

  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(&ring[warp][0][0]);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[warp][0]);
  auto issue = [&](int k, int stage) {        // lane 0 only: request both patches of slot slot0 + k into `stage`
    const uint32_t key = __shfl_sync(0xffffffffu, mykey, k);
    const int l = __shfl_sync(0xffffffffu, myl, k);
    if (lane == 0) {
      const int x = key_x(key) + kMinBorder, y = key_y(key) + kMinBorder;
      const int xu = x - 15, xb = x - 18;
      const uint32_t bar = bar0 + 8 * stage, dst = ring0 + kOdStage * stage;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kOdTx) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
          "l"(reinterpret_cast<uint64_t>(TM.map + kMaxLevels + l)), "r"(xu & ~15), "r"(y - 15), "r"((l == 0 ? TM.frame0 : 0) + f), "r"(bar)
          : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst + kOdBOff),
          "l"(reinterpret_cast<uint64_t>(TM.map + 2 * kMaxLevels + l)), "r"(xb & ~15), "r"(y - 18), "r"(f), "r"(bar)
          : "memory");
    }
  };

  int cur = __ffs(todo) - 1;
  todo &= todo - 1;
  issue(cur, 0);
  unsigned n = 0;                             // keypoints done by this warp: stage = n & 1, barrier parity = (n >> 1) & 1
  while (cur >= 0) {
    const int nxt = todo ? __ffs(todo) - 1 : -1;
    todo &= todo - 1;
    const int stage = n & 1;
    if (nxt >= 0) issue(nxt, stage ^ 1);      // the other stage was released by the __syncwarp that ended the previous iteration
    {
      const uint32_t bar = bar0 + 8 * stage, par = (n >> 1) & 1;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "OW_%=:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra OD_%=;\n\t"
          "bra OW_%=;\n\t"
          "OD_%=:\n\t}" ::"r"(bar), "r"(par) : "memory");
    }
    const uint32_t key = __shfl_sync(0xffffffffu, mykey, cur);
    const int l = __shfl_sync(0xffffffffu, myl, cur);
    const int o = __shfl_sync(0xffffffffu, myo, cur);
    const LevelGeom& L = G.L[l];
    const int x = key_x(key) + kMinBorder, y = key_y(key) + kMinBorder;   // (ORBextractor.cpp:851-852)
    const uint8_t* pu = &ring[warp][stage][0];
    const uint8_t* pb = pu + kOdBOff;
    const int oxu = (x - 15) & 15, oxb = (x - 18) & 15;

    // ---- moments over the radius-15 disc: lane <-> row v = lane-15 (ORBextractor.cpp:79-107) --------------
    int m10 = 0, m01 = 0;
    if (lane < 31) {
      const int v = lane - 15;
      const uint4* row = reinterpret_cast<const uint4*>(pu + lane * kOdUW);
      const uint32_t* mk = momMask + abs(v) * 9;
      const int bsh = 8 * (oxu & 3);
      int sum = 0;
      switch (oxu >> 2) {
        case 0: od_row_moments<0>(row, bsh, mk, m10, sum); break;
        case 1: od_row_moments<1>(row, bsh, mk, m10, sum); break;
        case 2: od_row_moments<2>(row, bsh, mk, m10, sum); break;
        default: od_row_moments<3>(row, bsh, mk, m10, sum); break;
      }
      m01 = v * sum;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rBRIEF on the blurred patch (ORBextractor.cpp:110-151) ---------------------------------------------
    const float factorPI = (float)(3.1415926535897932384626433832795 / (double)180.f);
    float a, b;
    glibc_sincosf(__fmul_rn(angle, factorPI), b, a);
    // sample address = centre + row * 64 + col with row, col carrying the rounding bias 0x4B400000 each
    const uint32_t bl = (uint32_t)__cvta_generic_to_shared(pb) + 18 * kOdBW + oxb + 18 - 65u * 0x4B400000u;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 p0 = pat[2 * j], p1 = pat[2 * j + 1];
      const uint32_t r0 = (uint32_t)od_round(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)));
      const uint32_t c0 = (uint32_t)od_round(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)));
      const uint32_t r1 = (uint32_t)od_round(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a)));
      const uint32_t c1 = (uint32_t)od_round(__fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b)));
      uint32_t t0, t1;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(bl + r0 * kOdBW + c0));
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(bl + r1 * kOdBW + c1));
      val |= (t0 < t1) << j;
    }
    desc_out[((size_t)(frame0 + f) * cap + o) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
      orbx_keypoint k;
      k.x = (float)x; k.y = (float)y;
      if (l != 0) { k.x = __fmul_rn(k.x, L.scale); k.y = __fmul_rn(k.y, L.scale); }   // (:1102-1108)
      k.size = L.kpSize; k.angle = angle; k.response = (float)key_s(key); k.octave = l; k.class_id = -1;
      kps_out[(size_t)(frame0 + f) * cap + o] = k;
    }
    __syncwarp();                             // every lane is done with this stage before it is requested again
    cur = nxt;
    ++n;
  }
}

// Latency mode: with a handful of frames per launch (orbx_extract: one) the chunk chain is bound by launch + CTA start-up
// latency, so every kernel is launched as a programmatic dependent launch; with full chunks that costs 2 % (ORBX_PDL).
// (measured, one frame per call: VGA 0.242 -> 0.208 ms; at 1080p the early-resident CTAs of the next kernel take slots
// from the running one and it gets slower, hence the pixel bound)
constexpr int kLatencyFrames = 8;
constexpr long long kLatencyPixels = 1000000;
static inline bool pdl_long(const Geom& G, int nframes) {
  return pdl_enabled() >= 2 || (pdl_enabled() >= 1 && (long long)nframes * G.W * G.H <= kLatencyPixels);
}

// ---- launch wrappers (called from the C ABI in orb_capi.cu) ---------------------------------------------
void launch_resize(const uint8_t* src, int sw, int sh, int spitch, size_t sframe, uint8_t* dst, int dw, int dh,
                   int dpitch, size_t dframe, const ResizeTaps& T, const ResizeTma& R, const CUtensorMap* map, int z0, int nframes,
                   cudaStream_t st) {
  dim3 block(32, 8), grid((dw + 127) / 128, (dh + 7) / 8, nframes);
  const bool aligned = ((((uintptr_t)src) | (uintptr_t)spitch | (uintptr_t)sframe) & 3) == 0;
  // default: TMA-staged tiles; ORBX_RESIZE_TMA=0 selects the register-blocked global-load walk, ORBX_RESIZE_WALK=0 the
  // one-quad-per-thread kernel (A/B runs)
  static const bool tma = !(getenv("ORBX_RESIZE_TMA") && atoi(getenv("ORBX_RESIZE_TMA")) == 0);
  static const bool walk = !(getenv("ORBX_RESIZE_WALK") && atoi(getenv("ORBX_RESIZE_WALK")) == 0);
  if (T.quadOk && R.use && tma && map) {
    const dim3 tblock(32, 4), tgrid((dw + kRzTileW - 1) / kRzTileW, (dh + 4 * R.rows - 1) / (4 * R.rows), nframes);
    launch_chain(pdl_enabled() >= 1, resize_tma_kernel, tgrid, tblock, (size_t)R.boxW * R.boxH, st, map, z0, sh, dst, dw, dh, dpitch, dframe, T, R);
  } else if (T.quadOk && aligned && walk) {
    const dim3 wblock(32, 4), wgrid((dw + 127) / 128, (dh + 4 * kRwRows - 1) / (4 * kRwRows), nframes);
    launch_chain(pdl_enabled() >= 1, resize_walk_kernel, wgrid, wblock, 0, st, src, sw, sh, spitch, sframe, dst, dw, dh, dpitch, dframe, T);
  } else if (T.quadOk && aligned)
    launch_chain(pdl_enabled() >= 1, resize4_kernel, grid, block, 0, st, src, sw, sh, spitch, sframe, dst, dw, dh, dpitch, dframe, T);
  else
    launch_chain(pdl_enabled() >= 1, resize_kernel, grid, block, 0, st, src, sw, sh, spitch, sframe, dst, dw, dh, dpitch, dframe, T);
}

void launch_pyramid_fused(const PyrLevel* d_plan, const PyrLaunch& P, int* d_sync, size_t smem, bool pdl, cudaStream_t st) {
  const dim3 tblock(32, 4), tgrid(P.start[P.nlev]);
  launch_chain(pdl, pyramid_fused_kernel, tgrid, tblock, smem, st, d_plan, P, d_sync);
}
bool pyramid_pdl() { return pdl_enabled() >= 1; }

size_t fast_warp_smem_bytes(const Geom& G);

size_t fast_smem_bytes(const Geom& G) {
  return (size_t)4 * G.fastTileW * G.fastTileH + 4 * (size_t)G.fastSurvCap + (size_t)G.fastTileW * G.fastTileH / 2 +
         (size_t)2 * G.fastTileW * G.fastTileH + 64;
}

size_t octree_smem_bytes(const Geom& G) {
  const size_t cap = G.nodeCap;
  return 8 * cap + 8 * cap * 2 + 4 * cap * 2 + 16 * cap + 4 * cap + 4 * (cap + 1) + 4 * cap + 16 * cap +
         4 * (size_t)(G.maxSlotsPerLevel + 1) + 4 * 40 + 64;
}
// latency mode: (key, label) pairs kept in shared memory behind the node arrays -- as many as a 200 KB CTA allows, at most the
// largest level's candidate capacity
int octree_key_smem(const Geom& G) {
  int want = 0;
  for (int l = 0; l < G.nlevels; ++l) want = std::max(want, G.L[l].keyCap);
  const size_t base = octree_smem_bytes(G), room = base < ((size_t)200 << 10) ? (((size_t)200 << 10) - base) / 6 : 0;
  return (int)std::min<size_t>((size_t)want, room) & ~7;
}

cudaError_t configure_kernels(const Geom& G) {
  cudaError_t e = cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(G));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fast_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_warp_smem_bytes(G));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(octree_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(octree_smem_bytes(G) + 6 * (size_t)octree_key_smem(G)));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(octree_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem_bytes(G));
}

size_t fast_warp_smem_bytes(const Geom& G) {
  return (size_t)G.fastTileW * G.fastTileH + (size_t)G.fastScW * G.fastScH +
         (size_t)kWarpCells * (2 * G.fastCellPix + ((std::max(2 * G.fastCellQuads, 2 * G.fastCellSurv) + 15) & ~15)) + 64;
}

static bool use_fast_warp() {
  // default: the warp-synchronous kernel (measured 9 % faster); ORBX_FAST_WARP=0 selects the CTA-wide one for A/B runs
  static const bool v = !(getenv("ORBX_FAST_WARP") && atoi(getenv("ORBX_FAST_WARP")) == 0) && kCellsPerCta == kWarpCells;
  return v;
}

// level0 / nlev: the pyramid levels of this launch (the whole pyramid by default; latency mode splits it over two streams)
void launch_fast(const Geom& G, const Bufs& B0, const TmaSet& TM, int nframes, cudaStream_t st, int level0, int nlev) {
  if (nlev < 0) nlev = G.nlevels - level0;
  Bufs B = B0;
  B.slotOff = G.L[level0].slot0;
  const int nslots = (level0 + nlev < G.nlevels ? G.L[level0 + nlev].slot0 : G.totalSlots) - B.slotOff;
  if (use_fast_warp()) {
    launch_chain(pdl_long(G, nframes), fast_warp_kernel, ORBX_FAST_FF ? dim3(nframes, nslots) : dim3(nslots, nframes), dim3(32 * kWarpCells), fast_warp_smem_bytes(G), st, G, B, TM);
    return;
  }
  launch_chain(pdl_long(G, nframes), fast_kernel, dim3(nslots, nframes), dim3(kFastThreads), fast_smem_bytes(G), st, G, B, TM);
}
void launch_octree(const Geom& G, const Bufs& B0, int nframes, cudaStream_t st, int level0, int nlev) {
  if (nlev < 0) nlev = G.nlevels - level0;
  Bufs B = B0;
  B.levelOff = level0;
  // CTA size: barriers dominate small levels (128 threads: 1.74 ms vs 1.96 at 256 for 640x480), the quadratic ranking pass
  // and the key loops dominate large ones (4K level 0: ~150 k candidates, ~1100 nodes); ORBX_OCT_THREADS overrides
  static const int forced = getenv("ORBX_OCT_THREADS") ? atoi(getenv("ORBX_OCT_THREADS")) : 0;
  const long long area = (long long)G.W * G.H;
  int threads = area < 600000 ? 128 : area < 3000000 ? 512 : 1024;
  if (nframes <= kLatencyFrames) threads = std::max(threads, 512);   // a handful of CTAs: the level-0 CTA is the critical path    // measured: 1080p 0.67 / 0.42 / 0.35 ms per 128 frames at 128 / 256 / 512
  if (forced == 128 || forced == 256 || forced == 512 || forced == 1024) threads = forced;
  static const bool keySmem = !(getenv("ORBX_OCT_KEYSMEM") && atoi(getenv("ORBX_OCT_KEYSMEM")) == 0);
  B.octKeySmem = (keySmem && nframes <= kLatencyFrames) ? octree_key_smem(G) : 0;
  if (B.octKeySmem > 0)
    launch_chain(pdl_long(G, nframes), octree_kernel<true>, ORBX_OCT_FF ? dim3(nframes, nlev) : dim3(nlev, nframes), dim3(threads),
                 octree_smem_bytes(G) + 6 * (size_t)B.octKeySmem, st, G, B);
  else
    launch_chain(pdl_long(G, nframes), octree_kernel<false>, ORBX_OCT_FF ? dim3(nframes, nlev) : dim3(nlev, nframes), dim3(threads),
                 octree_smem_bytes(G), st, G, B);
}
void launch_blur(const Geom& G, const Bufs& B, const TmaSet& TM, int nframes, cudaStream_t st) {
  // default: TMA-staged tiles; ORBX_BLUR_TMA=0 selects the global-load walk for A/B runs
  static const bool tma = !(getenv("ORBX_BLUR_TMA") && atoi(getenv("ORBX_BLUR_TMA")) == 0);
  if (tma && TM.useBlur) {
    launch_chain(pdl_long(G, nframes), blur_tma_kernel, dim3(nframes, G.bwTiles), dim3(128), 0, st, G, B, TM);
    return;
  }
  // default: the register-blocked kernel; ORBX_BLUR_WALK=0 selects the shared-memory tile kernel for A/B runs
  static const bool walk = !(getenv("ORBX_BLUR_WALK") && atoi(getenv("ORBX_BLUR_WALK")) == 0);
  if (walk) launch_chain(pdl_long(G, nframes), blur_walk_kernel, dim3(nframes, G.bwTiles), dim3(128), 0, st, G, B);
  else launch_chain(pdl_long(G, nframes), blur_kernel, dim3(G.blurTiles, nframes), dim3(256), 0, st, G, B);
}
void launch_orient_desc(const Geom& G, const Bufs& B, const TmaSet& TM, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts,
                        int frame0, int nframes, cudaStream_t st) {
  // default: the TMA-fed pipelined kernel; ORBX_OD_TMA=0 (or a level that TMA cannot describe) selects the LDGSTS one
  static const bool tma = !(getenv("ORBX_OD_TMA") && atoi(getenv("ORBX_OD_TMA")) == 0);
  if (tma && TM.usePatch) {
    const int K = nframes <= kLatencyFrames ? 2 : kOdK;      // few frames: more, shorter warps (12.4 -> ~5 us for one VGA frame)
    const int per = kOdWarps * K;
    launch_chain(pdl_long(G, nframes), orient_desc_tma_kernel, ORBX_OD_FF ? dim3(nframes, (G.selPerFrame + per - 1) / per) : dim3((G.selPerFrame + per - 1) / per, nframes), dim3(32 * kOdWarps), 0, st,
                 G, B, TM, kps, desc, cap, counts, frame0, K);
    return;
  }
  launch_chain(pdl_long(G, nframes), orient_desc_kernel, dim3((G.selPerFrame + 7) / 8, nframes), dim3(256), 0, st, G, B, kps, desc, cap, counts, frame0);
}

}  // namespace orbx
