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
// That turns the sequential loop into a fixed point: every round recomputes all choices in parallel from the
// previous round's block times; point i is final after at most i+1 rounds and in practice after a handful.
// Candidates are produced once (parallel over points, window walked ix-outer / iy-inner like frame.cpp:223-243
// so distance ties resolve identically) and only the cheap selection is iterated.
#include <limits.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace orbx {

constexpr int GC = ORBX_GRID_COLS, GR = ORBX_GRID_ROWS, NCELL = GC * GR;
constexpr int TH_HIGH = 100;      // matcher.cpp:11
constexpr int HISTO = 30;         // matcher.cpp:13

struct FrameDev {
  const orbx_keypoint* kps; const uint8_t* desc; const float* uright; int n;
  float xmin, xmax, ymin, ymax, gw, gh;
  const float* scale; int nlevels; const uint8_t* occupied0;
  const int* cellStart; const int* ids;
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
    if (w.aux < 0.0f) return w;                                          // z < 0 (:51)
    const int xMin = (int)F.xmin, xMax = (int)F.xmax, yMin = (int)F.ymin, yMax = (int)F.ymax;   // :27-30
    if (w.u < xMin || w.u > xMax) return w;                              // :60-63
    if (w.v < yMin || w.v > yMax) return w;
    w.r = __fmul_rn(S.radius, F.scale[lvl]);                             // :67
    if (S.forward) { w.minL = lvl; w.maxL = F.nlevels; }                 // :70-75
    else if (S.backward) { w.minL = 0; w.maxL = lvl; }
    else { w.minL = lvl - 1; w.maxL = lvl + 1; }
  }
  w.ok = true;
  return w;
}

// Walk the window exactly like Frame::getFeaturesInArea and apply the per-candidate gates of the matcher loop
// that do not depend on earlier assignments (stereo consistency).  fn(idx) is called in traversal order.
template <bool LOCAL, typename Fn>
__device__ __forceinline__ void walk_window(const FrameDev& F, const SearchParams& S, const Window& w, Fn fn) {
  const float du = __fsub_rn(w.u, F.xmin), dv = __fsub_rn(w.v, F.ymin);
  const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(du, w.r), F.gw)));          // frame.cpp:205-221
  if (x0 >= GC) return;
  const int x1 = min(GC - 1, (int)floorf(__fmul_rn(__fadd_rn(du, w.r), F.gw)));
  if (x1 < 0) return;
  const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dv, w.r), F.gh)));
  if (y0 >= GR) return;
  const int y1 = min(GR - 1, (int)floorf(__fmul_rn(__fadd_rn(dv, w.r), F.gh)));
  if (y1 < 0) return;
  for (int ix = x0; ix <= x1; ++ix)
    for (int iy = y0; iy <= y1; ++iy) {
      const int c = ix * GR + iy;
      for (int e = F.cellStart[c]; e < F.cellStart[c + 1]; ++e) {
        const int idx = F.ids[e];
        const orbx_keypoint k = F.kps[idx];
        if (k.octave < w.minL || k.octave > w.maxL) continue;
        if (!(fabsf(__fsub_rn(k.x, w.u)) < w.r && fabsf(__fsub_rn(k.y, w.v)) < w.r)) continue;
        const float urt = F.uright[idx];
        if (urt > 0) {
          float err;
          if (LOCAL) err = fabsf(__fsub_rn(w.aux, urt));                                        // matcher.cpp:317-322
          else err = fabsf(__fsub_rn(__fsub_rn(w.u, __fmul_rn(S.bf, w.aux)), urt));             // :90-96
          if (err > w.r) continue;
        }
        fn(idx);
      }
    }
}

__device__ __forceinline__ int hamming256(const uint8_t* a, const uint8_t* b) {
  const uint4* p = reinterpret_cast<const uint4*>(a);
  const uint4* q = reinterpret_cast<const uint4*>(b);
  const uint4 a0 = p[0], a1 = p[1], b0 = q[0], b1 = q[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- grid build (one CTA) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) grid_build_kernel(const orbx_keypoint* __restrict__ kps, int n, float xmin, float ymin,
                                                          float gw, float gh, int* cellOf, int* cellStart, int* ids) {
  __shared__ int cnt[NCELL];
  __shared__ int ws[40];
  const int tid = threadIdx.x, T = blockDim.x;
  for (int c = tid; c < NCELL; c += T) cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const int gx = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, xmin), gw));      // frame.cpp:83-84
    const int gy = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, ymin), gh));
    int c = -1;
    if (gx >= 0 && gx < GC && gy >= 0 && gy < GR) { c = gx * GR + gy; atomicAdd(&cnt[c], 1); }   // :91-97
    cellOf[i] = c;
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

// ---- candidate generation ---------------------------------------------------------------------------------
template <bool LOCAL>
__global__ void sbp_count_kernel(FrameDev F, PointsDev P, SearchParams S, int* counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.m) return;
  int c = 0;
  const Window w = make_window<LOCAL>(F, P, S, i);
  if (w.ok) walk_window<LOCAL>(F, S, w, [&](int) { ++c; });
  counts[i] = c;
}

__global__ void __launch_bounds__(1024) sbp_scan_kernel(int* counts, int m, int* total) {
  // in-place exclusive scan of counts[0..m) by one CTA, chunked through shared memory
  __shared__ int buf[4096];
  __shared__ int ws[40];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 4096) {
    const int len = min(4096, m - base);
    for (int i = threadIdx.x; i < len; i += blockDim.x) buf[i] = counts[base + i];
    __syncthreads();
    const int t = block_exclusive_scan(buf, len, ws);
    const int c = carry;
    for (int i = threadIdx.x; i < len; i += blockDim.x) counts[base + i] = buf[i] + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + t;
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[m] = carry; *total = carry; }
}

template <bool LOCAL>
__global__ void sbp_fill_kernel(FrameDev F, PointsDev P, SearchParams S, const int* offs, uint32_t* cand) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.m) return;
  const Window w = make_window<LOCAL>(F, P, S, i);
  if (!w.ok) return;
  int o = offs[i];
  const uint8_t* d = P.desc + (size_t)i * 32;
  walk_window<LOCAL>(F, S, w, [&](int idx) {
    const int dist = hamming256(d, F.desc + (size_t)idx * 32);
    cand[o++] = (uint32_t)idx | ((uint32_t)dist << 20);
  });
}

// ---- ordered resolution (one CTA) -------------------------------------------------------------------------
template <bool LOCAL>
__device__ __forceinline__ int select_choice(const FrameDev& F, const SearchParams& S, const uint32_t* cand, int b, int e, int i,
                                             const int* blockTime) {
  int bestD = 256, bestI = -1, bestD2 = 256, bestL = -1, bestL2 = -1;
  for (int k = b; k < e; ++k) {
    const uint32_t c = cand[k];
    const int idx = (int)(c & 0xFFFFFu), d = (int)(c >> 20);
    if (blockTime[idx] < i) continue;                                  // matcher.cpp:87 / :314
    if (LOCAL) {
      if (d < bestD) { bestD2 = bestD; bestD = d; bestL2 = bestL; bestL = F.kps[idx].octave; bestI = idx; }   // :327-339
      else if (d < bestD2) { bestL2 = F.kps[idx].octave; bestD2 = d; }
    } else {
      if (d < bestD) { bestD = d; bestI = idx; }                       // :101-105
    }
  }
  if (bestD > TH_HIGH) return -1;                                      // :108 / :342
  if (LOCAL && bestL == bestL2 && (float)bestD > __fmul_rn(S.ratio, (float)bestD2)) return -1;   // :344
  return bestI;
}

template <bool LOCAL>
__global__ void __launch_bounds__(1024) sbp_resolve_kernel(FrameDev F, PointsDev P, SearchParams S, const int* offs,
                                                           const uint32_t* cand, int* choice, int* blockTime, int32_t* assign,
                                                           int* match_cnt, int* rounds_out) {
  __shared__ int changed, cnt, hist[HISTO], keepBin[3];
  const int tid = threadIdx.x, T = blockDim.x, m = P.m, n = F.n;
  for (int i = tid; i < m; i += T) choice[i] = -1;
  int rounds = 0;
  while (true) {
    if (tid == 0) changed = 0;
    for (int c = tid; c < n; c += T) blockTime[c] = F.occupied0[c] ? -1 : INT_MAX;
    __syncthreads();
    for (int i = tid; i < m; i += T) {
      const int c = choice[i];
      if (c >= 0 && P.has_obs[i]) atomicMin(&blockTime[c], i);
    }
    __syncthreads();
    for (int i = tid; i < m; i += T) {
      const int b = offs[i], e = offs[i + 1];
      const int nc = (e > b) ? select_choice<LOCAL>(F, S, cand, b, e, i, blockTime) : -1;
      if (nc != choice[i]) { choice[i] = nc; changed = 1; }
    }
    __syncthreads();
    ++rounds;
    const bool again = changed != 0;
    __syncthreads();
    if (!again || rounds > m + 2) break;
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
  if (tid == 0) { *match_cnt = cnt; *rounds_out = rounds; }
}

// ---- host orchestration -----------------------------------------------------------------------------------
struct DevArena {      // grow-only per-thread device scratch so repeated searches do not pay cudaMalloc
  uint8_t* base = nullptr; size_t cap = 0, used = 0; int device = -1;
  int reserve(size_t bytes, int dev) {
    if (dev != device || bytes > cap) {
      if (base) { cudaSetDevice(device < 0 ? dev : device); cudaFree(base); base = nullptr; cap = 0; }
      if (cudaSetDevice(dev) != cudaSuccess) return ORBX_ERR_CUDA;
      size_t want = std::max(bytes, (size_t)8 << 20);
      if (cudaMalloc(&base, want) != cudaSuccess) { base = nullptr; return ORBX_ERR_CUDA; }
      cap = want; device = dev;
    }
    used = 0;
    return ORBX_OK;
  }
  template <typename T> T* take(size_t count) {
    used = align_up_sz(used, 256);
    T* p = reinterpret_cast<T*>(base + used);
    used += sizeof(T) * count;
    return p;
  }
};
static thread_local DevArena g_arena, g_cand_arena;

template <typename T> static size_t padded(size_t count) { return align_up_sz(sizeof(T) * count, 256) + 256; }

struct FrameUpload { FrameDev F; int* cellOf; int* cellStart; int* ids; };

static size_t frame_bytes(const orbx_frame_view* f) {
  const size_t n = std::max(f->n, 1);
  return padded<orbx_keypoint>(n) + padded<uint8_t>(n * 32) + padded<float>(n) + padded<float>(f->nlevels + 1) + padded<uint8_t>(n) +
         padded<int>(n) + padded<int>(NCELL + 1) + padded<int>(n);
}

static int upload_frame(DevArena& A, const orbx_frame_view* f, FrameUpload& U, cudaStream_t st) {
  const int n = f->n;
  orbx_keypoint* kps = A.take<orbx_keypoint>(std::max(n, 1));
  uint8_t* desc = A.take<uint8_t>((size_t)std::max(n, 1) * 32);
  float* ur = A.take<float>(std::max(n, 1));
  float* sc = A.take<float>(f->nlevels + 1);
  uint8_t* occ = A.take<uint8_t>(std::max(n, 1));
  U.cellOf = A.take<int>(std::max(n, 1)); U.cellStart = A.take<int>(NCELL + 1); U.ids = A.take<int>(std::max(n, 1));
  if (n > 0) {
    ORBX_CUDA(cudaMemcpyAsync(kps, f->kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice, st));
    ORBX_CUDA(cudaMemcpyAsync(desc, f->desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    ORBX_CUDA(cudaMemcpyAsync(ur, f->uright, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    ORBX_CUDA(cudaMemcpyAsync(occ, f->occupied0, n, cudaMemcpyHostToDevice, st));
  }
  ORBX_CUDA(cudaMemcpyAsync(sc, f->scale_factors, sizeof(float) * f->nlevels, cudaMemcpyHostToDevice, st));
  FrameDev& F = U.F;
  F.kps = kps; F.desc = desc; F.uright = ur; F.n = n; F.xmin = f->xmin; F.xmax = f->xmax; F.ymin = f->ymin; F.ymax = f->ymax;
  F.gw = (float)GC / (f->xmax - f->xmin);           // camera.cpp:47-48
  F.gh = (float)GR / (f->ymax - f->ymin);
  F.scale = sc; F.nlevels = f->nlevels; F.occupied0 = occ; F.cellStart = U.cellStart; F.ids = U.ids;
  grid_build_kernel<<<1, 1024, 0, st>>>(kps, n, F.xmin, F.ymin, F.gw, F.gh, U.cellOf, U.cellStart, U.ids);
  return ORBX_OK;
}

static int check_frame(const orbx_frame_view* f) {
  if (!f || f->n < 0 || f->n >= (1 << 20) || f->nlevels < 1 || !f->scale_factors || !(f->xmax > f->xmin) || !(f->ymax > f->ymin) ||
      (f->n > 0 && (!f->kps || !f->desc || !f->uright || !f->occupied0))) {
    set_error("bad frame view");
    return ORBX_ERR_ARG;
  }
  return ORBX_OK;
}

template <bool LOCAL>
static int run_search(const orbx_frame_view* frame, int m, const uint8_t* valid, const float* u, const float* v, const float* aux,
                      const int32_t* level, const float* angle_or_cos, const uint8_t* desc, const uint8_t* has_obs,
                      const SearchParams& S, int32_t* assign, int* match_cnt, int device) {
  if (check_frame(frame)) return ORBX_ERR_ARG;
  if (!assign || !match_cnt || m < 0 || (m > 0 && (!valid || !u || !v || !aux || !level || !angle_or_cos || !desc || !has_obs))) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  const int n = frame->n;
  if (m == 0 || n == 0) {
    for (int i = 0; i < n; ++i) assign[i] = -1;
    *match_cnt = 0;
    return ORBX_OK;
  }
  const size_t bytes = frame_bytes(frame) + padded<uint8_t>(m) * 2 + padded<float>(m) * 4 + padded<int32_t>(m) + padded<uint8_t>((size_t)m * 32) +
                       padded<int>(m + 1) + padded<int>(m) + padded<int>(n) + padded<int32_t>(n) + padded<int>(4);
  if (g_arena.reserve(bytes, device)) { set_error("device scratch allocation failed"); return ORBX_ERR_CUDA; }
  cudaStream_t st = nullptr;
  DevArena& A = g_arena;
  FrameUpload U;
  int rc = upload_frame(A, frame, U, st);
  if (rc) return rc;
  PointsDev P;
  P.m = m;
  uint8_t* d_valid = A.take<uint8_t>(m); uint8_t* d_obs = A.take<uint8_t>(m);
  float* d_u = A.take<float>(m); float* d_v = A.take<float>(m); float* d_aux = A.take<float>(m); float* d_ac = A.take<float>(m);
  int32_t* d_level = A.take<int32_t>(m); uint8_t* d_desc = A.take<uint8_t>((size_t)m * 32);
  int* d_offs = A.take<int>(m + 1); int* d_choice = A.take<int>(m); int* d_block = A.take<int>(n);
  int32_t* d_assign = A.take<int32_t>(n); int* d_misc = A.take<int>(4);
  ORBX_CUDA(cudaMemcpyAsync(d_valid, valid, m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_obs, has_obs, m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_u, u, sizeof(float) * m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_v, v, sizeof(float) * m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_aux, aux, sizeof(float) * m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_ac, angle_or_cos, sizeof(float) * m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_level, level, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(d_desc, desc, (size_t)m * 32, cudaMemcpyHostToDevice, st));
  P.valid = d_valid; P.u = d_u; P.v = d_v; P.aux = d_aux; P.level = d_level; P.angle_or_cos = d_ac; P.desc = d_desc; P.has_obs = d_obs;
  const int blocks = (m + 127) / 128;
  sbp_count_kernel<LOCAL><<<blocks, 128, 0, st>>>(U.F, P, S, d_offs);
  sbp_scan_kernel<<<1, 1024, 0, st>>>(d_offs, m, d_misc);
  int total = 0;
  ORBX_CUDA(cudaMemcpyAsync(&total, d_misc, sizeof(int), cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  if (g_cand_arena.reserve(padded<uint32_t>(std::max(total, 1)), device)) { set_error("device scratch allocation failed"); return ORBX_ERR_CUDA; }
  uint32_t* d_cand = g_cand_arena.take<uint32_t>(std::max(total, 1));
  sbp_fill_kernel<LOCAL><<<blocks, 128, 0, st>>>(U.F, P, S, d_offs, d_cand);
  sbp_resolve_kernel<LOCAL><<<1, 1024, 0, st>>>(U.F, P, S, d_offs, d_cand, d_choice, d_block, d_assign, d_misc + 1, d_misc + 2);
  int res[2] = {0, 0};
  ORBX_CUDA(cudaMemcpyAsync(assign, d_assign, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(res, d_misc + 1, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  ORBX_CUDA(cudaGetLastError());
  *match_cnt = res[0];
  return ORBX_OK;
}

}  // namespace orbx

using namespace orbx;

extern "C" {

int orbx_grid_build(const orbx_keypoint* kps, int n, float xmin, float xmax, float ymin, float ymax, int32_t* cell_start,
                    int32_t* ids, int device) {
  if (n < 0 || (n > 0 && !kps) || !cell_start || !ids || !(xmax > xmin) || !(ymax > ymin)) { set_error("bad argument"); return ORBX_ERR_ARG; }
  const size_t nn = std::max(n, 1);
  if (g_arena.reserve(padded<orbx_keypoint>(nn) + padded<int>(nn) * 2 + padded<int>(NCELL + 1), device)) {
    set_error("device scratch allocation failed");
    return ORBX_ERR_CUDA;
  }
  orbx_keypoint* d_k = g_arena.take<orbx_keypoint>(nn);
  int* d_cellOf = g_arena.take<int>(nn); int* d_start = g_arena.take<int>(NCELL + 1); int* d_ids = g_arena.take<int>(nn);
  if (n) ORBX_CUDA(cudaMemcpy(d_k, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice));
  grid_build_kernel<<<1, 1024>>>(d_k, n, xmin, ymin, (float)GC / (xmax - xmin), (float)GR / (ymax - ymin), d_cellOf, d_start, d_ids);
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
  return run_search<false>(frame, pts->m, pts->valid, pts->u, pts->v, pts->invz, pts->octave, pts->angle, pts->desc, pts->has_obs, S,
                           assign, match_cnt, device);
}

int orbx_search_by_projection_local(const orbx_frame_view* frame, const orbx_sbp_local_points* pts, float th_radius, float ratio,
                                    int32_t* assign, int* match_cnt, int device) {
  if (!pts) { set_error("null points"); return ORBX_ERR_ARG; }
  SearchParams S{};
  S.radius = th_radius; S.ratio = ratio;
  return run_search<true>(frame, pts->m, pts->valid, pts->u, pts->v, pts->ur, pts->level, pts->view_cos, pts->desc, pts->has_obs, S,
                          assign, match_cnt, device);
}

}  // extern "C"
