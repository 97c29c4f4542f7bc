// sbp.cu -- Frame grid index and the grid-window projection searches for sm_100a.
//
//   grid_build      Frame::assignFeaturesToGrid                         (frame.cpp:72-97, camera.h:8-9)
//   window query    Frame::getFeaturesInArea                            (frame.cpp:199-247)
//   sbp_frame       Matcher::searchByProjection(Frame*,Frame*,r,rot)     (matcher.cpp:18-148)
//   sbp_local       Matcher::searchByProjection(Frame*,vector<MapPoint*>)(matcher.cpp:274-353)
//
// The reference loops are greedy: a feature claimed by an earlier map point whose observation count is > 0 is
// skipped by later points (matcher.cpp:87,314), later points overwrite earlier claims (:110,:347).  Once a
// feature is claimed by a point with observations nobody can claim it again, so "blocked" is monotone and
//   blocked(feature c, seen by point i)  <=>  occupied0[c]  or  exists j < i with choice[j] == c and has_obs[j].
// That turns the sequential loop into a fixed point.  Points are resolved in index order in chunks of one CTA
// (1024 points): inside a chunk all choices are recomputed in parallel from the previous round's block times until
// nothing changes (a point only depends on earlier points, so the fixed point is unique and equals the sequential
// result); the chunk's block times are then frozen for the following chunks.
//
// Candidates are produced once by one warp per map point.  The grid is a CSR with cell index ix*48+iy, so for a fixed
// ix the cells iy0..iy1 of a window are ONE contiguous CSR range: the warp strides over it with coalesced loads and
// compacts the survivors in order with ballots, which is exactly the ix-outer / iy-inner / in-cell order of
// frame.cpp:223-243 (so distance ties resolve identically).
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "frame_handle.cuh"

namespace orbx {

constexpr int GC = ORBX_GRID_COLS, GR = ORBX_GRID_ROWS, NCELL = GC * GR;
constexpr int TH_HIGH = 100;      // matcher.cpp:11
constexpr int HISTO = 30;         // matcher.cpp:13
constexpr int kResolveThreads = 1024;

struct FrameDev {
  const orbx_keypoint* kps; const uint8_t* desc; int n;
  float xmin, xmax, ymin, ymax, gw, gh;
  const float* scale; int nlevels; const uint8_t* occupied0;
  const int* cellStart; const int* ids;
  const float4* feat;               // per feature: x, y, octave (as float bits), uRight
};

struct PointsDev {   // union of the two variants' per-point inputs
  int m;
  const uint8_t* valid; const float* u; const float* v; const float* aux;   // aux: invz (frame) / ur (local)
  const int32_t* level; const float* angle_or_cos; const uint8_t* desc; const uint8_t* has_obs;
};

struct SearchParams {
  float radius;      // frame: radius; local: thRadius
  float bf; int forward, backward, check_rot;
  float ratio;
  float th;          // accept iff (float)bestDist <= th   (TH_HIGH, distThreshold or TH_LOW depending on the overload)
  int block_all;     // every accepted claim blocks the feature (overloads that test `mappoints_[idx]` without the obs count)
  int pos_block;     // the "already matched" test is indexed by the candidate's POSITION in the window list instead of its
                     // feature index: the reference bug `matchMapPoints[j]` at matcher.cpp:422, reproduced for parity
  int level_at_select;   // window query has no level filter (KeyFrame::getFeaturesInArea); levels gate at selection (:425-427)
  int no_block;      // points are independent (no claimed-feature test at all): searchBySim3 / fuse overloads
  int chi2;          // fuseMapPoints' reprojection gate (matcher.cpp:1077-1095); per-point aux carries ur = u - bf/z
  int host_gates;    // depth-sign and image-bounds gates were already applied by the caller (folded into valid[])
  int level_span_lo, level_span_hi;   // non-LOCAL level window = [lvl + lo, lvl + hi] when neither forward nor backward
};

struct Window { bool ok; float u, v, r, aux; int minL, maxL; };

template <bool LOCAL>
__device__ __forceinline__ Window make_window(const FrameDev& F, const PointsDev& P, const SearchParams& S, int i) {
  Window w;
  w.ok = false;
  if (!P.valid[i]) return w;
  w.u = P.u[i]; w.v = P.v[i]; w.aux = P.aux[i];
  const int lvl = P.level[i];
  if (LOCAL) {
    float radius = ((double)P.angle_or_cos[i] > 0.998) ? 2.5f : 4.0f;    // matcher.cpp:287-293
    radius = __fmul_rn(radius, S.radius);
    w.r = __fmul_rn(radius, F.scale[lvl]);                               // :296
    w.minL = lvl - 1; w.maxL = lvl;                                      // :297-298
  } else {
    if (!S.host_gates) {
      if (w.aux < 0.0f) return w;                                        // z < 0 (:51)
      const int xMin = (int)F.xmin, xMax = (int)F.xmax, yMin = (int)F.ymin, yMax = (int)F.ymax;   // :27-30
      if (w.u < xMin || w.u > xMax) return w;                            // :60-63
      if (w.v < yMin || w.v > yMax) return w;
    }
    w.r = __fmul_rn(S.radius, F.scale[lvl]);                             // :67
    if (S.forward) { w.minL = lvl; w.maxL = F.nlevels; }                 // :70-75
    else if (S.backward) { w.minL = 0; w.maxL = lvl; }
    else { w.minL = lvl + S.level_span_lo; w.maxL = lvl + S.level_span_hi; }
  }
  w.ok = true;
  return w;
}

// cell range of a window (frame.cpp:205-221): floor for both ends
__device__ __forceinline__ bool window_cells(const FrameDev& F, const Window& w, int& x0, int& x1, int& y0, int& y1) {
  const float du = __fsub_rn(w.u, F.xmin), dv = __fsub_rn(w.v, F.ymin);
  x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(du, w.r), F.gw)));
  if (x0 >= GC) return false;
  x1 = min(GC - 1, (int)floorf(__fmul_rn(__fadd_rn(du, w.r), F.gw)));
  if (x1 < 0) return false;
  y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dv, w.r), F.gh)));
  if (y0 >= GR) return false;
  y1 = min(GR - 1, (int)floorf(__fmul_rn(__fadd_rn(dv, w.r), F.gh)));
  if (y1 < 0) return false;
  return true;
}

// level / distance / stereo gates of one candidate (frame.cpp:234-241, matcher.cpp:90-96 / :317-322)
template <bool LOCAL>
__device__ __forceinline__ bool gate(const FrameDev& F, const SearchParams& S, const Window& w, const float4 ft) {
  const int oct = __float_as_int(ft.z);
  if (!S.level_at_select && (oct < w.minL || oct > w.maxL)) return false;
  if (!(fabsf(__fsub_rn(ft.x, w.u)) < w.r && fabsf(__fsub_rn(ft.y, w.v)) < w.r)) return false;
  if (S.chi2) {                                                          // matcher.cpp:1073-1095
    const float ex = __fsub_rn(w.u, ft.x), ey = __fsub_rn(w.v, ft.y);
    const float invSigma = __fdiv_rn(1.0f, F.scale[oct]);
    float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    float lim = 5.991f;
    if (ft.w >= 0) {
      const float er = __fsub_rn(w.aux, ft.w);
      e2 = __fadd_rn(e2, __fmul_rn(er, er));
      lim = 7.815f;
    }
    return !(__fmul_rn(__fmul_rn(e2, invSigma), invSigma) > lim);
  }
  if (ft.w > 0 && (LOCAL || !S.host_gates)) {      // stereo consistency exists only in the two tracking overloads
    float err;
    if (LOCAL) err = fabsf(__fsub_rn(w.aux, ft.w));
    else err = fabsf(__fsub_rn(__fsub_rn(w.u, __fmul_rn(S.bf, w.aux)), ft.w));
    if (err > w.r) return false;
  }
  return true;
}

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint8_t* b) {
  const uint4* q = reinterpret_cast<const uint4*>(b);
  const uint4 b0 = __ldg(q), b1 = __ldg(q + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- grid build (one CTA) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) grid_build_kernel(const orbx_keypoint* __restrict__ kps, const float* __restrict__ uright,
                                                          int n, float xmin, float ymin, float gw, float gh, int* cellOf,
                                                          int* cellStart, int* ids, float4* feat, float* angle = nullptr) {
  __shared__ int cnt[NCELL];
  __shared__ int ws[40];
  const int tid = threadIdx.x, T = blockDim.x;
  for (int c = tid; c < NCELL; c += T) cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const orbx_keypoint k = kps[i];
    const int gx = (int)roundf(__fmul_rn(__fsub_rn(k.x, xmin), gw));      // frame.cpp:83-84
    const int gy = (int)roundf(__fmul_rn(__fsub_rn(k.y, ymin), gh));
    int c = -1;
    if (gx >= 0 && gx < GC && gy >= 0 && gy < GR) { c = gx * GR + gy; atomicAdd(&cnt[c], 1); }   // :91-97
    cellOf[i] = c;
    if (feat) feat[i] = make_float4(k.x, k.y, __int_as_float(k.octave), uright ? uright[i] : -1.f);
    if (angle) angle[i] = k.angle;
  }
  __syncthreads();
  const int total = block_exclusive_scan(cnt, NCELL, ws);
  for (int c = tid; c < NCELL; c += T) cellStart[c] = cnt[c];
  if (tid == 0) cellStart[NCELL] = total;
  __syncthreads();
  // ids ascending inside a cell (push_back order of the reference loop)
  for (int i = tid; i < n; i += T) {
    const int c = cellOf[i];
    if (c < 0) continue;
    int r = 0;
    for (int j = 0; j < i; ++j) r += (cellOf[j] == c) ? 1 : 0;
    ids[cnt[c] + r] = i;
  }
}

// ---- candidate generation: one warp per map point ---------------------------------------------------------
// Candidate word: feature index (20 bits) | distance << 20 (9 bits, 511 = "never selectable") | octave << 29 (3 bits, only read
// when the frame has at most 8 levels).
// slotCap > 0: ONE pass, point i owns cand[i * slotCap .. + slotCap) (the common case: a window holds a few dozen features);
//              a point with more candidates raises misc[3] and the host re-runs with slotCap == 0.
// slotCap == 0: two passes, exact allocation: misc[0] = running total of candidates, misc[3] = overflow of `cap`.
constexpr int kSlotCap = 64;
constexpr int kCandCache = 12;      // candidates a resolve thread keeps in registers across its fixed-point rounds
template <bool LOCAL>
__global__ void __launch_bounds__(256) sbp_walk_kernel(FrameDev F, PointsDev P, SearchParams S, int* offs, int* cnts, uint32_t* cand,
                                                       int cap, int slotCap, int* misc) {
  pdl_prologue();      // lets the resolve kernel (a programmatic dependent launch) become resident while this grid drains
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= P.m) return;
  const Window w = make_window<LOCAL>(F, P, S, i);
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  const bool any = w.ok && window_cells(F, w, x0, x1, y0, y1);
  int base = 0, limit = 0;
  if (slotCap > 0) {
    base = i * slotCap; limit = base + slotCap;
    if (!any) { if (lane == 0) { offs[i] = base; cnts[i] = 0; } return; }
  } else {
    int cnt = 0;
    if (any) {
      for (int ix = x0; ix <= x1; ++ix) {
        const int lo = __ldg(F.cellStart + ix * GR + y0), hi = __ldg(F.cellStart + ix * GR + y1 + 1);
        for (int e = lo + lane; e < ((hi - lo + 31) & ~31) + lo; e += 32) {
          bool pass = false;
          if (e < hi) pass = gate<LOCAL>(F, S, w, __ldg(F.feat + __ldg(F.ids + e)));
          cnt += __popc(__ballot_sync(0xffffffffu, pass));
        }
      }
    }
    if (lane == 0) {
      if (cnt > 0) base = atomicAdd(&misc[0], cnt);
      offs[i] = base; cnts[i] = cnt;
      if (base + cnt > cap) { misc[3] = 1; cnts[i] = 0; }
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cnt == 0 || base + cnt > cap) return;
    limit = base + cnt;
  }
  const uint4* dq = reinterpret_cast<const uint4*>(P.desc + (size_t)i * 32);
  const uint4 d0 = __ldg(dq), d1 = __ldg(dq + 1);
  int run = base;
  auto emit = [&](bool live, int e) {        // one CSR entry per lane, survivors appended in lane order
    bool pass = false;
    int idx = 0;
    float4 ft = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) { idx = __ldg(F.ids + e); ft = __ldg(F.feat + idx); pass = gate<LOCAL>(F, S, w, ft); }
    const unsigned b = __ballot_sync(0xffffffffu, pass);
    const int pos = run + __popc(b & ((1u << lane) - 1));
    if (pass && pos < limit) {
      int dist = hamming256(d0, d1, F.desc + (size_t)idx * 32);
      const int oct = __float_as_int(ft.z);
      if (S.level_at_select && (oct < w.minL || oct > w.maxL)) dist = 511;   // keeps its position in the list but can never be selected
      cand[pos] = (uint32_t)idx | ((uint32_t)dist << 20) | ((uint32_t)(oct & 7) << 29);
    }
    run += __popc(b);
  };
  const int ncol = x1 - x0 + 1;
  if (ncol <= 32) {
    // The window's columns are walked as ONE flat list: lane c fetches the CSR range of column x0 + c (all columns at once), a warp
    // scan gives every column its offset, and entry t of the list belongs to the column whose range covers it.  Same order as
    // the column loop below (ix outer, CSR position inner), but the dependent chain cellStart -> ids -> feat -> descriptor is
    // paid once per 32 entries instead of once per column (a tracking window holds a handful of features in 4..7 columns).
    int lo = 0, len = 0;
    if (lane < ncol) {
      lo = __ldg(F.cellStart + (x0 + lane) * GR + y0);
      len = __ldg(F.cellStart + (x0 + lane) * GR + y1 + 1) - lo;
    }
    int inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int E = __shfl_sync(0xffffffffu, inc, 31), pre = inc - len;
    for (int t0 = 0; t0 < E; t0 += 32) {
      const int t = t0 + lane;
      int c = 0;
      for (int j = 0; j < ncol - 1; ++j) c += (__shfl_sync(0xffffffffu, inc, j) <= t) ? 1 : 0;
      const int lo_c = __shfl_sync(0xffffffffu, lo, c), pre_c = __shfl_sync(0xffffffffu, pre, c);
      emit(t < E, lo_c + (t - pre_c));
    }
  } else {
    for (int ix = x0; ix <= x1; ++ix) {
      const int lo = __ldg(F.cellStart + ix * GR + y0), hi = __ldg(F.cellStart + ix * GR + y1 + 1);
      for (int e = lo + lane; e < ((hi - lo + 31) & ~31) + lo; e += 32) emit(e < hi, e);
    }
  }
  if (slotCap > 0 && lane == 0) {
    offs[i] = base; cnts[i] = min(run - base, slotCap);
    if (run > limit) misc[3] = 1;
    atomicAdd(&misc[0], run - base);
  }
}

// ---- ordered resolution (one CTA) -------------------------------------------------------------------------
// One candidate of point i against the running best / second best (matcher.cpp:87-105 / :314-339 / :218 / :422-427).
template <bool LOCAL>
__device__ __forceinline__ void consider(const FrameDev& F, const SearchParams& S, uint32_t c, int pos, int i, const int* blockTime,
                                         bool octInCand, int& bestD, int& bestI, int& bestD2, int& bestL, int& bestL2) {
  const int idx = (int)(c & 0xFFFFFu), d = (int)((c >> 20) & 0x1FFu);
  if (blockTime[S.pos_block ? pos : idx] < i) return;                  // matcher.cpp:87 / :314 / :218 / :422
  if (d == 511) return;                                                // level gate applied at selection (:425-427)
  if (LOCAL) {
    if (d < bestD) { bestD2 = bestD; bestD = d; bestL2 = bestL; bestL = octInCand ? (int)(c >> 29) : F.kps[idx].octave; bestI = idx; }   // :327-339
    else if (d < bestD2) { bestL2 = octInCand ? (int)(c >> 29) : F.kps[idx].octave; bestD2 = d; }
  } else {
    if (d < bestD) { bestD = d; bestI = idx; }                         // :101-105
  }
}
// cache[j] = cand[b + j] for j < min(e - b, kCandCache): fetched once per point, reused by every round of the fixed point
template <bool LOCAL>
__device__ __forceinline__ int select_choice(const FrameDev& F, const SearchParams& S, const uint32_t* cand, const uint32_t (&cache)[kCandCache],
                                             int b, int e, int i, const int* blockTime, bool octInCand) {
  int bestD = 256, bestI = -1, bestD2 = 256, bestL = -1, bestL2 = -1;
#pragma unroll
  for (int j = 0; j < kCandCache; ++j)
    if (b + j < e) consider<LOCAL>(F, S, cache[j], j, i, blockTime, octInCand, bestD, bestI, bestD2, bestL, bestL2);
  for (int k = b + kCandCache; k < e; ++k) consider<LOCAL>(F, S, cand[k], k - b, i, blockTime, octInCand, bestD, bestI, bestD2, bestL, bestL2);
  if (!((float)bestD <= S.th)) return -1;                              // :108 / :342 / :233 / :439
  if (LOCAL && bestL == bestL2 && (float)bestD > __fmul_rn(S.ratio, (float)bestD2)) return -1;   // :344
  return bestI;
}

// blockF / blockT: block times (global scratch, or shared memory when n fits)
template <bool LOCAL>
__global__ void __launch_bounds__(kResolveThreads) sbp_resolve_kernel(FrameDev F, PointsDev P, SearchParams S, const int* offs,
                                                                     const int* cnts, const uint32_t* cand, int* choice,
                                                                     int* gBlockF, int* gBlockT, int smemN, int32_t* assign,
                                                                     int* misc, int* hostOut, int wantChoice) {
  pdl_prologue();      // blocks until the walk kernel's candidates are complete and visible
  extern __shared__ int sblock[];
  __shared__ int cnt, hist[HISTO], keepBin[3];
  const int tid = threadIdx.x, T = blockDim.x, m = P.m, n = F.n;
  const bool octInCand = F.nlevels <= 8;                   // the 3-bit octave field of the candidate word is complete
  int* blockF = (n <= smemN) ? sblock : gBlockF;
  int* blockT = (n <= smemN) ? sblock + smemN : gBlockT;
  for (int c = tid; c < n; c += T) blockF[c] = (!S.no_block && F.occupied0[c]) ? -1 : INT_MAX;
  __syncthreads();
  int rounds = 0;
  for (int c0 = 0; c0 < m; c0 += T) {
    const int i = c0 + tid;
    const bool valid = i < m;
    const int b = valid ? offs[i] : 0, e = valid ? b + cnts[i] : 0;
    const bool obs = valid && !S.no_block && (S.block_all || P.has_obs[i]);
    uint32_t cache[kCandCache];
#pragma unroll
    for (int j = 0; j < kCandCache; ++j) cache[j] = b + j < e ? cand[b + j] : 0u;     // independent loads: one L2 latency per point
    int ch = -1;
    while (true) {
      for (int c = tid; c < n; c += T) blockT[c] = blockF[c];
      __syncthreads();
      if (obs && ch >= 0) atomicMin(&blockT[ch], i);
      __syncthreads();
      const int nc = (e > b) ? select_choice<LOCAL>(F, S, cand, cache, b, e, i, blockT, octInCand) : -1;
      const int changed = nc != ch;
      ch = nc;
      ++rounds;
      if (!__syncthreads_or(changed)) break;
    }
    // blockT holds the block times implied by the (now stable) choices of this chunk: freeze them
    for (int c = tid; c < n; c += T) blockF[c] = blockT[c];
    if (valid) choice[i] = ch;
    __syncthreads();
  }
  // final holders: the last accepted writer of each feature (matcher.cpp:110 / :347)
  if (tid == 0) cnt = 0;
  if (tid < HISTO) hist[tid] = 0;
  for (int c = tid; c < n; c += T) assign[c] = -1;
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < m; i += T) {
    const int c = choice[i];
    if (c < 0) continue;
    ++mine;
    atomicMax(&assign[c], i);
    if (!LOCAL && S.check_rot) {
      float rot = __fsub_rn(P.angle_or_cos[i], F.kps[c].angle);       // :115-122
      if (rot < 0) rot = __fadd_rn(rot, 360.0f);
      int bin = __float2int_rn(__fmul_rn(rot, (float)HISTO / 360.0f));
      if (bin == HISTO) bin = 0;
      if (bin >= 0 && bin < HISTO) atomicAdd(&hist[bin], 1);
    }
  }
  if (mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (!LOCAL && S.check_rot) {
    if (tid == 0) {   // computeThreeMax, matcher.cpp:1258-1304
      int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
      for (int i = 0; i < HISTO; ++i) {
        const int s = hist[i];
        if (s > m1) { m3 = m2; i3 = i2; m2 = m1; i2 = i1; m1 = s; i1 = i; }
        else if (s > m2) { m3 = m2; i3 = i2; m2 = s; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
      }
      if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
      else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
      keepBin[0] = i1; keepBin[1] = i2; keepBin[2] = i3;
    }
    __syncthreads();
    int dropped = 0;
    for (int i = tid; i < m; i += T) {
      const int c = choice[i];
      if (c < 0) continue;
      float rot = __fsub_rn(P.angle_or_cos[i], F.kps[c].angle);
      if (rot < 0) rot = __fadd_rn(rot, 360.0f);
      int bin = __float2int_rn(__fmul_rn(rot, (float)HISTO / 360.0f));
      if (bin == HISTO) bin = 0;
      if (bin != keepBin[0] && bin != keepBin[1] && bin != keepBin[2]) { assign[c] = -2; ++dropped; }   // :136-144
    }
    if (dropped) atomicSub(&cnt, dropped);
    __syncthreads();
  }
  if (tid == 0) { misc[1] = cnt; misc[2] = rounds; }
  if (hostOut) {
    // zero-copy result: [counters (8) | assign (n) | choice (m)] written straight into the caller's pinned (mapped) block, so
    // the search ends with one stream synchronisation instead of a device->host copy; the counters are then cleared for the
    // next search (they live in a persistent device block in this mode: no upload zeroes them)
    __syncthreads();
    if (tid < 8) hostOut[tid] = misc[tid];
    for (int c = tid; c < n; c += T) hostOut[8 + c] = assign[c];
    if (wantChoice) for (int i = tid; i < m; i += T) hostOut[8 + n + i] = choice[i];
    __syncthreads();
    if (tid < 8) misc[tid] = 0;
  }
}

// ---- host orchestration -----------------------------------------------------------------------------------
static thread_local DevArena g_arena, g_cand_arena;
static thread_local HostArena g_host;
// zero-copy searches: 8 counters per device, zero between searches (cleared by the resolve kernel that read them)
struct ZcMisc {
  int* p[64] = {};
  int* get(int dev) {
    if (dev < 0 || dev >= 64) return nullptr;
    if (!p[dev]) {
      if (cudaSetDevice(dev) != cudaSuccess || cudaMalloc(&p[dev], sizeof(int) * 8) != cudaSuccess) { p[dev] = nullptr; return nullptr; }
      if (cudaMemset(p[dev], 0, sizeof(int) * 8) != cudaSuccess) { cudaFree(p[dev]); p[dev] = nullptr; return nullptr; }
    }
    return p[dev];
  }
};
static thread_local ZcMisc g_zcMisc;
static bool env_flag(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) != 0 : dflt != 0; }

// one packed upload: the same layout on the host staging buffer and on the device
struct Packer {
  size_t used = 0;
  size_t add(size_t bytes) { used = align_up_sz(used, 256); size_t o = used; used += bytes; return o; }
};

static int check_frame(const orbx_frame_view* f) {
  if (!f || f->n < 0 || f->n >= (1 << 20) || f->nlevels < 1 || !f->scale_factors || !(f->xmax > f->xmin) || !(f->ymax > f->ymin) ||
      (f->n > 0 && (!f->kps || !f->desc || !f->uright || !f->occupied0))) {
    set_error("bad frame view");
    return ORBX_ERR_ARG;
  }
  return ORBX_OK;
}

// `rf` != nullptr: the frame side is a device-resident orbx_frame (keypoints, descriptors, uRight, grid and per-feature records
// already in HBM, built by orbx_frame_create); `frame` then only carries n / bounds / occupied0 and nothing of it but
// occupied0 is uploaded, and no grid is built.
template <bool LOCAL>
static int run_search(const orbx_frame_view* frame, int m, const uint8_t* valid, const float* u, const float* v, const float* aux,
                      const int32_t* level, const float* angle_or_cos, const uint8_t* desc, const uint8_t* has_obs,
                      const SearchParams& S, int32_t* assign, int* match_cnt, int device, int32_t* choice_out = nullptr,
                      const orbx_frame* rf = nullptr) {
  if (!rf && check_frame(frame)) return ORBX_ERR_ARG;
  if (rf && (!frame || frame->n != rf->n || (frame->n > 0 && !frame->occupied0))) { set_error("bad resident frame view"); return ORBX_ERR_ARG; }
  if (!assign || !match_cnt || m < 0 || (m > 0 && (!valid || !u || !v || !aux || !level || !angle_or_cos || !desc || !has_obs))) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  const int n = frame->n;
  if (m == 0 || n == 0) {
    for (int i = 0; i < n; ++i) assign[i] = -1;
    if (choice_out) for (int i = 0; i < m; ++i) choice_out[i] = -1;
    *match_cnt = 0;
    return ORBX_OK;
  }
  // ---- pack every input into one pinned staging block ------------------------------------------------------
  Packer pk;
  const size_t o_kps = pk.add(rf ? 0 : sizeof(orbx_keypoint) * n), o_desc = pk.add(rf ? 0 : (size_t)n * 32), o_ur = pk.add(rf ? 0 : sizeof(float) * n);
  const size_t o_sc = pk.add(rf ? 0 : sizeof(float) * frame->nlevels), o_occ = pk.add(n);
  const size_t o_valid = pk.add(m), o_obs = pk.add(m), o_u = pk.add(sizeof(float) * m), o_v = pk.add(sizeof(float) * m);
  const size_t o_aux = pk.add(sizeof(float) * m), o_ac = pk.add(sizeof(float) * m), o_lvl = pk.add(sizeof(int32_t) * m);
  const size_t o_pd = pk.add((size_t)m * 32);
  // counters = the LAST 32 bytes of the upload (zeroed by it, no separate memset); the results follow directly, so that
  // [counters | assign (n) | choice (m)] come back in ONE copy into the pinned staging block
  const size_t inBytes = align_up_sz(pk.used + sizeof(int) * 8, 256);
  const size_t o_misc = inBytes - sizeof(int) * 8;
  pk.used = inBytes;
  const size_t o_assign = pk.used; pk.used += sizeof(int32_t) * n;
  const size_t o_choice = pk.used; pk.used += sizeof(int) * m;
  const size_t outBytes = pk.used - o_misc;
  // device-only scratch behind the inputs
  const size_t o_cellOf = pk.add(rf ? 0 : sizeof(int) * n), o_cellStart = pk.add(rf ? 0 : sizeof(int) * (NCELL + 1)), o_ids = pk.add(rf ? 0 : sizeof(int) * n);
  const size_t o_feat = pk.add(rf ? 0 : sizeof(float4) * n), o_offs = pk.add(sizeof(int) * m), o_cnts = pk.add(sizeof(int) * m);
  const size_t o_bF = pk.add(sizeof(int) * n), o_bT = pk.add(sizeof(int) * n);
  if (g_host.reserve(inBytes + outBytes + 256) || g_arena.reserve(pk.used + 256, device)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  uint8_t* hb = g_host.base;
  uint8_t* hout = hb + inBytes;
  memset(hb + o_misc, 0, sizeof(int) * 8);
  if (!rf) {
    memcpy(hb + o_kps, frame->kps, sizeof(orbx_keypoint) * n); memcpy(hb + o_desc, frame->desc, (size_t)n * 32);
    memcpy(hb + o_ur, frame->uright, sizeof(float) * n); memcpy(hb + o_sc, frame->scale_factors, sizeof(float) * frame->nlevels);
  }
  memcpy(hb + o_occ, frame->occupied0, n);
  memcpy(hb + o_valid, valid, m); memcpy(hb + o_obs, has_obs, m); memcpy(hb + o_u, u, sizeof(float) * m);
  memcpy(hb + o_v, v, sizeof(float) * m); memcpy(hb + o_aux, aux, sizeof(float) * m); memcpy(hb + o_ac, angle_or_cos, sizeof(float) * m);
  memcpy(hb + o_lvl, level, sizeof(int32_t) * m); memcpy(hb + o_pd, desc, (size_t)m * 32);
  cudaStream_t st = nullptr;
  uint8_t* db = g_arena.take<uint8_t>(pk.used);
  static const bool phaseTiming = getenv("ORBX_SEARCH_TIMING") != nullptr;      // debugging aid: host clock with a sync after every phase
  auto tnow = []() { return std::chrono::steady_clock::now(); };
  auto t0 = tnow();
  // Zero-copy mode (resident frame, small point sets -- the tracking thread's calls): the kernels read the points straight
  // from the pinned staging block over PCIe (each field once, coalesced) and the resolve kernel writes the result into it, so
  // the search is two launches + one synchronisation, no copies.  The counters then live in a persistent device block that the
  // resolve kernel clears at its end.
  static const bool zcOn = env_flag("ORBX_ZEROCOPY", 1);
  uint8_t* hdev = nullptr;      // device-side address of the staging block
  const bool zc = zcOn && rf && inBytes <= ((size_t)256 << 10) && cudaHostGetDevicePointer((void**)&hdev, hb, 0) == cudaSuccess && hdev &&
                  g_zcMisc.get(device) != nullptr;
  uint8_t* pb = zc ? hdev : db;   // where the kernels find the points
  if (!zc) ORBX_CUDA(cudaMemcpyAsync(db, hb, inBytes, cudaMemcpyHostToDevice, st));
  if (phaseTiming) { cudaStreamSynchronize(st); fprintf(stderr, "[sbp] m=%d n=%d upload %zu B: %.1f us\n", m, n, inBytes, std::chrono::duration<double, std::micro>(tnow() - t0).count()); t0 = tnow(); }
  int* d_misc = zc ? g_zcMisc.get(device) : (int*)(db + o_misc);

  FrameDev F;
  F.kps = rf ? rf->d_unkps : (const orbx_keypoint*)(db + o_kps); F.desc = rf ? rf->d_desc : db + o_desc; F.n = n;
  F.xmin = frame->xmin; F.xmax = frame->xmax; F.ymin = frame->ymin; F.ymax = frame->ymax;
  F.gw = (float)GC / (frame->xmax - frame->xmin);           // camera.cpp:47-48
  F.gh = (float)GR / (frame->ymax - frame->ymin);
  F.scale = rf ? rf->d_scale : (const float*)(db + o_sc); F.nlevels = frame->nlevels; F.occupied0 = pb + o_occ;
  F.cellStart = rf ? rf->d_cellStart : (const int*)(db + o_cellStart); F.ids = rf ? rf->d_ids : (const int*)(db + o_ids);
  F.feat = rf ? rf->d_feat : (const float4*)(db + o_feat);
  PointsDev P;
  P.m = m; P.valid = pb + o_valid; P.has_obs = pb + o_obs; P.u = (const float*)(pb + o_u); P.v = (const float*)(pb + o_v);
  P.aux = (const float*)(pb + o_aux); P.angle_or_cos = (const float*)(pb + o_ac); P.level = (const int32_t*)(pb + o_lvl);
  P.desc = pb + o_pd;

  if (!rf)
    grid_build_kernel<<<1, 1024, 0, st>>>(F.kps, (const float*)(db + o_ur), n, F.xmin, F.ymin, F.gw, F.gh, (int*)(db + o_cellOf),
                                          (int*)(db + o_cellStart), (int*)(db + o_ids), (float4*)(db + o_feat));
  // candidate buffer: first attempt = one-pass walk with kSlotCap slots per point; a point that needs more makes the search
  // re-run with the two-pass exact allocation (size known from the first attempt's counter)
  size_t capCand = std::max<size_t>((size_t)m * kSlotCap, 4096);
  int slotCap = kSlotCap;
  const int smemN = 8192;
  // the attribute is per device and every entry point takes a device argument: set it on every call (cheap, as frame.cu does)
  ORBX_CUDA(cudaFuncSetAttribute(sbp_resolve_kernel<LOCAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * 2 * smemN)));
  for (int attempt = 0; attempt < 3; ++attempt) {
    if (g_cand_arena.reserve(sizeof(uint32_t) * capCand + 256, device)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
    uint32_t* d_cand = g_cand_arena.take<uint32_t>(capCand);
    if (attempt && !zc) ORBX_CUDA(cudaMemsetAsync(d_misc, 0, sizeof(int) * 8, st));
    sbp_walk_kernel<LOCAL><<<(m + 7) / 8, 256, 0, st>>>(F, P, S, (int*)(db + o_offs), (int*)(db + o_cnts), d_cand, (int)std::min<size_t>(capCand, INT_MAX), slotCap, d_misc);
    if (phaseTiming) { cudaStreamSynchronize(st); fprintf(stderr, "[sbp] walk: %.1f us\n", std::chrono::duration<double, std::micro>(tnow() - t0).count()); t0 = tnow(); }
    launch_chain(!phaseTiming, sbp_resolve_kernel<LOCAL>, dim3(1), dim3(kResolveThreads), sizeof(int) * 2 * smemN, st, F, P, S,
                 (const int*)(db + o_offs), (const int*)(db + o_cnts), (const uint32_t*)d_cand, (int*)(db + o_choice), (int*)(db + o_bF),
                 (int*)(db + o_bT), smemN, (int32_t*)(db + o_assign), d_misc, zc ? (int*)(hdev + inBytes) : (int*)nullptr, choice_out ? 1 : 0);
    if (phaseTiming) { cudaStreamSynchronize(st); fprintf(stderr, "[sbp] resolve: %.1f us\n", std::chrono::duration<double, std::micro>(tnow() - t0).count()); t0 = tnow(); }
    if (!zc) ORBX_CUDA(cudaMemcpyAsync(hout, db + o_misc, choice_out ? outBytes : outBytes - sizeof(int) * m, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    ORBX_CUDA(cudaGetLastError());
    if (phaseTiming) fprintf(stderr, "[sbp] download %zu B: %.1f us (rounds %d)\n", outBytes, std::chrono::duration<double, std::micro>(tnow() - t0).count(), ((const int*)hout)[2]);
    const int* res = (const int*)hout;
    if (!res[3]) {
      memcpy(assign, hout + sizeof(int) * 8, sizeof(int32_t) * n);
      if (choice_out) memcpy(choice_out, hout + sizeof(int) * 8 + sizeof(int32_t) * n, sizeof(int32_t) * m);
      *match_cnt = res[1];
      return ORBX_OK;
    }
    capCand = (size_t)res[0] + 1024;      // exact need (the allocation counter keeps counting past the capacity)
    slotCap = 0;
  }
  set_error("candidate buffer overflow");
  return ORBX_ERR_CAPACITY;
}

}  // namespace orbx

using namespace orbx;

extern "C" {

int orbx_grid_build(const orbx_keypoint* kps, int n, float xmin, float xmax, float ymin, float ymax, int32_t* cell_start,
                    int32_t* ids, int device) {
  if (n < 0 || (n > 0 && !kps) || !cell_start || !ids || !(xmax > xmin) || !(ymax > ymin)) { set_error("bad argument"); return ORBX_ERR_ARG; }
  const size_t nn = std::max(n, 1);
  if (g_arena.reserve(sizeof(orbx_keypoint) * nn + sizeof(int) * (2 * nn + NCELL + 1) + 2048, device)) {
    set_error("device scratch allocation failed");
    return ORBX_ERR_CUDA;
  }
  orbx_keypoint* d_k = g_arena.take<orbx_keypoint>(nn);
  int* d_cellOf = g_arena.take<int>(nn); int* d_start = g_arena.take<int>(NCELL + 1); int* d_ids = g_arena.take<int>(nn);
  if (n) ORBX_CUDA(cudaMemcpy(d_k, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice));
  grid_build_kernel<<<1, 1024>>>(d_k, nullptr, n, xmin, ymin, (float)GC / (xmax - xmin), (float)GR / (ymax - ymin), d_cellOf, d_start, d_ids,
                                 nullptr);
  ORBX_CUDA(cudaMemcpy(cell_start, d_start, sizeof(int) * (NCELL + 1), cudaMemcpyDeviceToHost));
  if (cell_start[NCELL] > 0) ORBX_CUDA(cudaMemcpy(ids, d_ids, sizeof(int) * cell_start[NCELL], cudaMemcpyDeviceToHost));
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int orbx_search_by_projection_frame(const orbx_frame_view* frame, const orbx_sbp_frame_points* pts, float radius, float bf,
                                    int forward, int backward, int check_rot, int32_t* assign, int* match_cnt, int device) {
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = radius; S.bf = bf; S.forward = forward; S.backward = backward; S.check_rot = check_rot; S.ratio = 0.f;
  S.th = (float)TH_HIGH; S.level_span_lo = -1; S.level_span_hi = 1;
  return run_search<false>(frame, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, device);
}

int orbx_search_by_projection_local(const orbx_frame_view* frame, const orbx_sbp_local_points* pts, float th_radius, float ratio,
                                    int32_t* assign, int* match_cnt, int device) {
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = th_radius; S.ratio = ratio; S.th = (float)TH_HIGH;
  return run_search<true>(frame, pts->m, pts->valid, pts->u, pts->v, pts->ur, pts->level, pts->view_cos, pts->desc, pts->has_obs, S,
                          assign, match_cnt, device);
}


// Make an existing host-side Frame / KeyFrame resident: unKeypoints_, descriptors_, uRight_, scaleFactors_ go up once, the grid
// and the per-feature records are built on the device.  The handle has no extractor (owner == nullptr) and no host mirror.
int orbx_frame_upload(const orbx_frame_view* view, int device, orbx_frame_t* out) {
  if (!out) { set_error("null argument"); return ORBX_ERR_ARG; }
  *out = nullptr;
  orbx_frame_view v;
  if (view) { v = *view; static const uint8_t kZero = 0; if (!v.occupied0) v.occupied0 = &kZero; }
  if (check_frame(view ? &v : nullptr)) return ORBX_ERR_ARG;
  ORBX_CUDA(cudaSetDevice(device));
  const int n = v.n, cap = std::max(n, 1), nl = v.nlevels;
  Packer pk;
  const size_t o_kps = pk.add(sizeof(orbx_keypoint) * cap), o_desc = pk.add((size_t)32 * cap), o_ur = pk.add(sizeof(float) * cap);
  const size_t o_sc = pk.add(sizeof(float) * nl);
  const size_t inBytes = align_up_sz(pk.used, 256);
  const size_t o_ang = pk.add(sizeof(float) * cap), o_cs = pk.add(sizeof(int) * (NCELL + 1)), o_ids = pk.add(sizeof(int) * cap);
  const size_t o_feat = pk.add(sizeof(float4) * cap), o_cellOf = pk.add(sizeof(int) * cap);
  if (g_host.reserve(inBytes)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  orbx_frame* f = new orbx_frame();
  if (cudaMalloc(&f->d_block, pk.used + 256) != cudaSuccess) { delete f; set_error("frame block allocation failed"); return ORBX_ERR_CUDA; }
  uint8_t* hb = g_host.base;
  if (n) {
    memcpy(hb + o_kps, v.kps, sizeof(orbx_keypoint) * n); memcpy(hb + o_desc, v.desc, (size_t)32 * n);
    memcpy(hb + o_ur, v.uright, sizeof(float) * n);
  }
  memcpy(hb + o_sc, v.scale_factors, sizeof(float) * nl);
  uint8_t* b = f->d_block;
  f->device = device; f->cap = cap; f->n = n; f->nlevels = nl;
  f->xmin = v.xmin; f->xmax = v.xmax; f->ymin = v.ymin; f->ymax = v.ymax;
  f->d_unkps = (orbx_keypoint*)(b + o_kps); f->d_kps = f->d_unkps; f->d_desc = b + o_desc; f->d_uright = (float*)(b + o_ur);
  f->d_scale = (float*)(b + o_sc); f->d_angle = (float*)(b + o_ang); f->d_cellStart = (int32_t*)(b + o_cs); f->d_ids = (int32_t*)(b + o_ids);
  f->d_feat = (float4*)(b + o_feat);
  cudaStream_t st = nullptr;
  cudaError_t e = cudaMemcpyAsync(b, hb, inBytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    grid_build_kernel<<<1, 1024, 0, st>>>(f->d_unkps, f->d_uright, n, f->xmin, f->ymin, (float)GC / (f->xmax - f->xmin),
                                          (float)GR / (f->ymax - f->ymin), (int*)(b + o_cellOf), f->d_cellStart, f->d_ids, f->d_feat, f->d_angle);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the staging block is reused by the next call
  if (e != cudaSuccess) { cudaFree(f->d_block); delete f; set_error(cudaGetErrorString(e)); return ORBX_ERR_CUDA; }
  *out = f;
  return ORBX_OK;
}

// ---- the same three tracking-thread searches against a device-resident frame (orbx_frame_t) ---------------------------------
static int resident_view(const orbx_frame* f, const uint8_t* occupied0, orbx_frame_view* v) {
  if (!f || (f->n > 0 && !occupied0)) { set_error("null frame handle / occupied0"); return ORBX_ERR_ARG; }
  memset(v, 0, sizeof(*v));
  v->n = f->n; v->xmin = f->xmin; v->xmax = f->xmax; v->ymin = f->ymin; v->ymax = f->ymax; v->nlevels = f->nlevels; v->occupied0 = occupied0;
  return ORBX_OK;
}

int orbx_search_by_projection_frame_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_frame_points* pts, float radius, float bf,
                                      int forward, int backward, int check_rot, int32_t* assign, int* match_cnt) {
  orbx_frame_view v;
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  if (resident_view(frame, occupied0, &v)) return ORBX_ERR_ARG;
  SearchParams S{};
  S.radius = radius; S.bf = bf; S.forward = forward; S.backward = backward; S.check_rot = check_rot; S.ratio = 0.f;
  S.th = (float)TH_HIGH; S.level_span_lo = -1; S.level_span_hi = 1;
  return run_search<false>(&v, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, frame->device, nullptr, frame);
}

int orbx_search_by_projection_local_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_local_points* pts, float th_radius,
                                      float ratio, int32_t* assign, int* match_cnt) {
  orbx_frame_view v;
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  if (resident_view(frame, occupied0, &v)) return ORBX_ERR_ARG;
  SearchParams S{};
  S.radius = th_radius; S.ratio = ratio; S.th = (float)TH_HIGH;
  return run_search<true>(&v, pts->m, pts->valid, pts->u, pts->v, pts->ur, pts->level, pts->view_cos, pts->desc, pts->has_obs, S,
                          assign, match_cnt, frame->device, nullptr, frame);
}

int orbx_search_by_projection_reloc_h(orbx_frame_t frame, const uint8_t* occupied0, const orbx_sbp_frame_points* pts, float radius,
                                      float dist_threshold, int check_rot, int32_t* assign, int* match_cnt) {
  orbx_frame_view v;
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  if (resident_view(frame, occupied0, &v)) return ORBX_ERR_ARG;
  SearchParams S{};
  S.radius = radius; S.check_rot = check_rot; S.th = dist_threshold; S.block_all = 1; S.host_gates = 1; S.level_span_lo = -1; S.level_span_hi = 1;
  return run_search<false>(&v, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, frame->device, nullptr, frame);
}

// Matcher::searchByProjection(Frame*, KeyFrame*, radius, distThreshold, found, checkRot)  (matcher.cpp:150-272).
// pts->valid folds the host-side gates (:173-200: null/bad/found, z <= 0, image bounds, distance range); pts->octave is
// mp->predictScale(); any feature that already holds a map point is skipped (:218), every accepted claim blocks.
int orbx_search_by_projection_reloc(const orbx_frame_view* frame, const orbx_sbp_frame_points* pts, float radius, float dist_threshold,
                                    int check_rot, int32_t* assign, int* match_cnt, int device) {
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = radius; S.check_rot = check_rot; S.th = dist_threshold; S.block_all = 1; S.host_gates = 1; S.level_span_lo = -1; S.level_span_hi = 1;
  return run_search<false>(frame, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, device);
}

// Matcher::searchByProjection(KeyFrame*, Sim3&, loopMapPoints, matchMapPoints, th)  (matcher.cpp:356-447).
// The window comes from KeyFrame::getFeaturesInArea (no level filter, keyframe.cpp:268-312); levels [l-1, l] gate at
// selection (:425-427); accept at TH_LOW (:439); the "already matched" test reads matchMapPoints[j] with the window
// POSITION j (:422) -- reproduced as is.  frame->occupied0 = matchMapPoints[i] != nullptr on entry.
int orbx_search_by_projection_sim3(const orbx_frame_view* keyframe, const orbx_sbp_frame_points* pts, int th, int32_t* assign,
                                   int* match_cnt, int device) {
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = (float)th; S.check_rot = 0; S.th = 50.0f; S.block_all = 1; S.host_gates = 1; S.pos_block = 1; S.level_at_select = 1;
  S.level_span_lo = -1; S.level_span_hi = 0;
  return run_search<false>(keyframe, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, device);
}


// Independent windowed best match per point: the common core of searchBySim3 (matcher.cpp:717-775, 777-836),
// fuseMapPoints (:1026-1100) and fuseByPose (:1157-1224): window from KeyFrame::getFeaturesInArea, levels
// [level_predict-1, level_predict], strict '<' argmin in window order, accepted iff bestDist <= dist_threshold.
// chi2_gate != 0 adds fuseMapPoints' reprojection test (pts->invz then carries ur = u - bf/z).  Points do not interact.
int orbx_window_argmin(const orbx_frame_view* keyframe, const orbx_sbp_frame_points* pts, float th_radius, float dist_threshold,
                       int chi2_gate, int32_t* best_idx, int device) {
  if (!pts || !best_idx) { set_error("null argument"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = th_radius; S.th = dist_threshold; S.no_block = 1; S.host_gates = 1; S.chi2 = chi2_gate; S.level_span_lo = -1; S.level_span_hi = 0;
  std::vector<int32_t> assign(std::max(keyframe ? keyframe->n : 0, 1));
  int cnt = 0;
  return run_search<false>(keyframe, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign.data(), &cnt, device, best_idx);
}

// Matcher::searchBySim3 (matcher.cpp:679-865): both directed searches + the mutual-consistency check (:848-861).
// pts12: kf1's map points projected into kf2 (m = kf1 features); pts21: kf2's projected into kf1.  match12[i] = matched
// kf2 feature of kf1 feature i, or -1.
int orbx_search_by_sim3(const orbx_frame_view* kf1, const orbx_sbp_frame_points* pts12, const orbx_frame_view* kf2,
                        const orbx_sbp_frame_points* pts21, float th, int32_t* match12, int* found, int device) {
  if (!kf1 || !kf2 || !pts12 || !pts21 || !match12 || !found) { set_error("null argument"); return ORBX_ERR_ARG; }
  if (pts12->m != kf1->n || pts21->m != kf2->n) { set_error("one projected point per key-frame feature expected"); return ORBX_ERR_ARG; }
  std::vector<int32_t> m1(std::max(pts12->m, 1)), m2(std::max(pts21->m, 1));
  int rc = orbx_window_argmin(kf2, pts12, th, (float)TH_HIGH, 0, m1.data(), device);
  if (rc) return rc;
  rc = orbx_window_argmin(kf1, pts21, th, (float)TH_HIGH, 0, m2.data(), device);
  if (rc) return rc;
  int f = 0;
  for (int i = 0; i < pts12->m; ++i) {
    const int idx2 = m1[i];
    match12[i] = -1;
    if (idx2 >= 0 && m2[idx2] == i) { match12[i] = idx2; ++f; }
  }
  *found = f;
  return ORBX_OK;
}

}  // extern "C"
