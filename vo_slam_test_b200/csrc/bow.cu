// bow.cu -- BoW-guided matching: Matcher::searchByBoW(KeyFrame*, Frame*, ...) (matcher.cpp:449-559) and
// Matcher::searchByBoW(KeyFrame*, KeyFrame*, ...) (matcher.cpp:561-677) for sm_100a.
//
// Both overloads merge-walk two DBoW3 FeatureVectors (vocabulary node -> feature indices) and, inside every node the
// two sides share, run the greedy best/second-best + ratio loop: side-A features in order, each scanning the side-B
// features of the node that are not yet taken (matcher.cpp:488 / :606).  A feature belongs to exactly one node, so
// the greedy dependency never crosses nodes: one warp owns one shared node and replays the loop exactly (A features
// sequentially, B features across the lanes, lexicographic (distance, position) top-2 by shuffles = strict '<' in
// scan order), nodes run in parallel.  The rotation-histogram check (computeThreeMax, :537-556 / :655-671) runs in a
// second one-CTA kernel.
#include <string.h>

#include <algorithm>

#include "frame_handle.cuh"

namespace orbx {

constexpr int HISTO_B = 30;

struct BowSideDev {
  int n; const uint8_t* desc; const float* angle; const uint8_t* valid;
  int ngroups; const uint32_t* node_ids; const int32_t* group_start; const int32_t* feat_idx;
};

__device__ __forceinline__ int hamm(const uint4 a0, const uint4 a1, const uint8_t* b) {
  const uint4* q = reinterpret_cast<const uint4*>(b);
  const uint4 b0 = __ldg(q), b1 = __ldg(q + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// mode 0: KeyFrame -> Frame   : match[b_idx] = a_idx   (mappointMatches is indexed by the frame feature, :508)
// mode 1: KeyFrame -> KeyFrame: match[a_idx] = b_idx   (mappointMatches is indexed by the kf1 feature, :629)
__global__ void __launch_bounds__(256) bow_match_kernel(BowSideDev A, BowSideDev B, int mode, float ratio, int th_low, int check_rot,
                                                        int* takenB, int32_t* match, int8_t* binOf, int* hist) {
  pdl_prologue();      // lets bow_finish_kernel (a programmatic dependent launch) become resident while this grid drains
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= A.ngroups) return;
  const uint32_t node = __ldg(A.node_ids + g);
  int lo = 0, hi = B.ngroups;                       // lower_bound of `node` in B's sorted node ids (:531-534)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(B.node_ids + mid) < node) lo = mid + 1; else hi = mid;
  }
  if (lo >= B.ngroups || __ldg(B.node_ids + lo) != node) return;
  const int as = __ldg(A.group_start + g), ae = __ldg(A.group_start + g + 1);
  const int bs = __ldg(B.group_start + lo), be = __ldg(B.group_start + lo + 1);
  const uint32_t kInit = (256u << 20) | 0xFFFFFu;
  for (int ia = as; ia < ae; ++ia) {
    const int idx1 = __ldg(A.feat_idx + ia);
    if (!A.valid[idx1]) continue;                   // `!mpk || mpk->isBad()` (:476 / :594)
    const uint4* dq = reinterpret_cast<const uint4*>(A.desc + (size_t)idx1 * 32);
    const uint4 d0 = __ldg(dq), d1 = __ldg(dq + 1);
    uint32_t k1 = kInit, k2 = kInit;
    for (int ib = bs + lane; ib < be; ib += 32) {
      const int idx2 = __ldg(B.feat_idx + ib);
      if (*((volatile int*)takenB + idx2) || !B.valid[idx2]) continue;    // (:488 / :606-609)
      const uint32_t key = ((uint32_t)hamm(d0, d1, B.desc + (size_t)idx2 * 32) << 20) | (uint32_t)(ib - bs);
      k2 = min(k2, max(key, k1));
      k1 = min(k1, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
      const uint32_t n1 = min(k1, o1);
      k2 = min(max(k1, o1), min(k2, o2));
      k1 = n1;
    }
    const int best1 = (int)(k1 >> 20), best2 = (int)(k2 >> 20);
    if (best1 <= th_low && (float)best1 < __fmul_rn(ratio, (float)best2)) {          // (:504-506 / :625-627)
      const int idx2 = __ldg(B.feat_idx + bs + (int)(k1 & 0xFFFFFu));
      if (lane == 0) {
        takenB[idx2] = 1;
        const int outIdx = mode == 0 ? idx2 : idx1;
        match[outIdx] = mode == 0 ? idx1 : idx2;
        if (check_rot) {
          float rot = __fsub_rn(A.angle[idx1], B.angle[idx2]);                       // (:513-521 / :634-642)
          if (rot < 0) rot = __fadd_rn(rot, 360.0f);
          const float v = __fmul_rn(rot, (float)HISTO_B / 360.0f);
          int bin = mode == 0 ? __float2int_rn(v) : (int)roundf(v);                  // cvRound (:517) vs round (:637)
          if (bin == HISTO_B) bin = 0;
          binOf[outIdx] = (int8_t)bin;
          if (bin >= 0 && bin < HISTO_B) atomicAdd(&hist[bin], 1);
        }
      }
    }
    __syncwarp();
  }
}

// ---- searchForTriangulation (matcher.cpp:867-1010) -----------------------------------------------------------
struct TriSideDev {
  BowSideDev s;                 // valid[i] = the feature has NO map point yet (:902, :921)
  const orbx_keypoint* kps;     // unKeypoints_
  const float* uright;          // uRight_ (>= 0: stereo)
};
struct TriParams {
  double F[9];                  // F12, row major
  float ex, ey;                 // epipole of camera 1 in image 2 (:886-890)
  const float* scale2;          // keyframe2->scaleFactors_
  int check_rot;
};

// Matcher::checkEpipolarConstrain (matcher.cpp:1306-1324): doubles up to the two casts to float, no FMA.
__device__ __forceinline__ bool epipolar_ok(const TriParams& T, const orbx_keypoint& k1, const orbx_keypoint& k2) {
  const double x1 = (double)k1.x, y1 = (double)k1.y, x2 = (double)k2.x, y2 = (double)k2.y;
  double l[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) l[j] = __dadd_rn(__dadd_rn(__dmul_rn(x1, T.F[j]), __dmul_rn(y1, T.F[3 + j])), T.F[6 + j]);
  const float num = (float)__dadd_rn(__dadd_rn(__dmul_rn(l[0], x2), __dmul_rn(l[1], y2)), l[2]);
  const float den = (float)__dadd_rn(__dmul_rn(l[0], l[0]), __dmul_rn(l[1], l[1]));
  if (den == 0) return false;
  const float d2 = __fdiv_rn(__fmul_rn(num, num), den);
  const float sigma = T.scale2[k2.octave];
  return d2 < __fmul_rn(__fmul_rn(3.84f, sigma), sigma);
}

// One warp per shared vocabulary node, like bow_match_kernel, but with the triangulation rules: candidates need
// dist <= TH_LOW and dist <= running best (so among equal distances the LAST one in scan order wins, :925-926,:941-945),
// mono-mono pairs close to the epipole are dropped (:930-938) and the epipolar test must pass.  match[idx1] = idx2.
__global__ void __launch_bounds__(256) tri_match_kernel(TriSideDev A, TriSideDev B, TriParams T, int th_low, int* takenB, int32_t* match,
                                                        int8_t* binOf, int* hist) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= A.s.ngroups) return;
  const uint32_t node = __ldg(A.s.node_ids + g);
  int lo = 0, hi = B.s.ngroups;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(B.s.node_ids + mid) < node) lo = mid + 1; else hi = mid;
  }
  if (lo >= B.s.ngroups || __ldg(B.s.node_ids + lo) != node) return;
  const int as = __ldg(A.s.group_start + g), ae = __ldg(A.s.group_start + g + 1);
  const int bs = __ldg(B.s.group_start + lo), be = __ldg(B.s.group_start + lo + 1);
  const uint32_t kNone = 0xFFFFFFFFu;
  for (int ia = as; ia < ae; ++ia) {
    const int idx1 = __ldg(A.s.feat_idx + ia);
    if (!A.s.valid[idx1]) continue;
    const bool stereo1 = A.uright[idx1] >= 0;
    const orbx_keypoint k1 = A.kps[idx1];
    const uint4* dq = reinterpret_cast<const uint4*>(A.s.desc + (size_t)idx1 * 32);
    const uint4 d0 = __ldg(dq), d1 = __ldg(dq + 1);
    uint32_t best = kNone;                            // dist << 20 | (0xFFFFF - position): min = smallest dist, then LAST position
    for (int ib = bs + lane; ib < be; ib += 32) {
      const int idx2 = __ldg(B.s.feat_idx + ib);
      if (*((volatile int*)takenB + idx2) || !B.s.valid[idx2]) continue;
      const int dist = hamm(d0, d1, B.s.desc + (size_t)idx2 * 32);
      if (dist > th_low) continue;
      const orbx_keypoint k2 = B.kps[idx2];
      if (!stereo1 && !(B.uright[idx2] >= 0)) {
        const float dx = __fsub_rn(T.ex, k2.x), dy = __fsub_rn(T.ey, k2.y);
        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, T.scale2[k2.octave])) continue;
      }
      if (!epipolar_ok(T, k1, k2)) continue;
      best = min(best, ((uint32_t)dist << 20) | (0xFFFFFu - (uint32_t)(ib - bs)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (best != kNone) {
      const int idx2 = __ldg(B.s.feat_idx + bs + (int)(0xFFFFFu - (best & 0xFFFFFu)));
      if (lane == 0) {
        takenB[idx2] = 1;
        match[idx1] = idx2;
        if (T.check_rot) {
          float rot = __fsub_rn(k1.angle, B.kps[idx2].angle);
          if (rot < 0) rot = __fadd_rn(rot, 360.0f);
          int bin = (int)roundf(__fmul_rn(rot, (float)HISTO_B / 360.0f));      // round(), :961
          if (bin == HISTO_B) bin = 0;
          binOf[idx1] = (int8_t)bin;
          if (bin >= 0 && bin < HISTO_B) atomicAdd(&hist[bin], 1);
        }
      }
    }
    __syncwarp();
  }
}

// hostOut != nullptr: [match (nOut) | count] also go straight into the caller's pinned (mapped) block -- the search then ends
// with one stream synchronisation instead of a device->host copy.
__global__ void __launch_bounds__(1024) bow_finish_kernel(int nOut, int check_rot, int32_t* match, const int8_t* binOf, const int* hist,
                                                          int* result, int32_t* hostOut = nullptr) {
  pdl_prologue();      // blocks until the match kernel's results are complete and visible
  __shared__ int keep[3], cnt;
  const int tid = threadIdx.x;
  if (tid == 0) {
    cnt = 0;
    int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;          // computeThreeMax, matcher.cpp:1258-1304
    for (int i = 0; i < HISTO_B; ++i) {
      const int s = hist[i];
      if (s > m1) { m3 = m2; i3 = i2; m2 = m1; i2 = i1; m1 = s; i1 = i; }
      else if (s > m2) { m3 = m2; i3 = i2; m2 = s; i2 = i; }
      else if (s > m3) { m3 = s; i3 = i; }
    }
    if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
    else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
    keep[0] = i1; keep[1] = i2; keep[2] = i3;
  }
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < nOut; i += blockDim.x) {
    int mi = match[i];
    if (mi >= 0) {
      if (check_rot) {
        const int b = binOf[i];
        if (b != keep[0] && b != keep[1] && b != keep[2]) { mi = -2; match[i] = -2; }   // cleared by the rotation check
      }
      if (mi >= 0) ++mine;
    }
    if (hostOut) hostOut[i] = mi;
  }
  if (mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (tid == 0) { *result = cnt; if (hostOut) hostOut[nOut] = cnt; }
}

struct Arena {
  uint8_t* d = nullptr; uint8_t* h = nullptr; size_t dcap = 0, hcap = 0; int device = -1;
  int reserve(size_t dbytes, size_t hbytes, int dev) {
    if (cudaSetDevice(dev) != cudaSuccess) return ORBX_ERR_CUDA;
    if (dev != device || dbytes > dcap) {
      if (d) cudaFree(d);
      d = nullptr; dcap = 0;
      const size_t want = std::max(dbytes + dbytes / 2, (size_t)4 << 20);
      if (cudaMalloc(&d, want) != cudaSuccess) { d = nullptr; return ORBX_ERR_CUDA; }
      dcap = want; device = dev;
    }
    if (hbytes > hcap) {
      if (h) cudaFreeHost(h);
      h = nullptr; hcap = 0;
      const size_t want = std::max(hbytes + hbytes / 2, (size_t)4 << 20);
      if (cudaMallocHost(&h, want) != cudaSuccess) { h = nullptr; return ORBX_ERR_CUDA; }
      hcap = want;
    }
    return ORBX_OK;
  }
};
static thread_local Arena g_bow;

static int check_side(const orbx_bow_side* s) {
  if (!s || s->n < 0 || s->ngroups < 0 || (s->n > 0 && (!s->desc || !s->angle || !s->valid)) ||
      (s->ngroups > 0 && (!s->node_ids || !s->group_start || !s->feat_idx))) {
    set_error("bad BoW side");
    return ORBX_ERR_ARG;
  }
  return ORBX_OK;
}

}  // namespace orbx

using namespace orbx;

// rfB != nullptr: side b is a device-resident frame (orbx_frame_t): its descriptors and angles are read in place from the
// handle, b->desc / b->angle are ignored and only b's validity flags and FeatureVector CSR are uploaded.
static int bow_run(const orbx_bow_side* a, const orbx_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                   int32_t* match, int* match_cnt, int device, const orbx_frame* rfB) {
  orbx_bow_side bfix;
  static const uint8_t kDummy = 0;
  if (rfB && b) {          // let the side checks pass without host descriptor / angle arrays
    bfix = *b;
    if (bfix.n != rfB->n) { set_error("side b: n differs from the frame handle"); return ORBX_ERR_ARG; }
    if (!bfix.desc) bfix.desc = &kDummy;
    if (!bfix.angle) bfix.angle = (const float*)&kDummy;
    b = &bfix;
  }
  if (check_side(a) || check_side(b)) return ORBX_ERR_ARG;
  if (!match || !match_cnt || (mode != 0 && mode != 1)) { set_error("bad argument"); return ORBX_ERR_ARG; }
  const int nOut = mode == 0 ? b->n : a->n;
  if (a->ngroups == 0 || b->ngroups == 0 || a->n == 0 || b->n == 0) {
    for (int i = 0; i < nOut; ++i) match[i] = -1;
    *match_cnt = 0;
    return ORBX_OK;
  }
  size_t used = 0;
  auto add = [&](size_t bytes) { used = align_up_sz(used, 256); size_t o = used; used += bytes; return o; };
  size_t oa[6], ob[6];
  const orbx_bow_side* sides[2] = {a, b};
  size_t* offs[2] = {oa, ob};
  for (int s = 0; s < 2; ++s) {
    const orbx_bow_side* S = sides[s];
    const int nfi = S->group_start[S->ngroups];
    const bool res = s == 1 && rfB;
    offs[s][0] = add(res ? 0 : (size_t)S->n * 32); offs[s][1] = add(res ? 0 : sizeof(float) * S->n); offs[s][2] = add(S->n);
    offs[s][3] = add(sizeof(uint32_t) * S->ngroups); offs[s][4] = add(sizeof(int32_t) * (S->ngroups + 1)); offs[s][5] = add(sizeof(int32_t) * nfi);
  }
  // the zero / -1 initialised work arrays ride in the same upload (no memsets); [match | res] come back in one copy
  const size_t o_taken = add(sizeof(int) * b->n), o_hist = add(sizeof(int) * 32 + 256);
  const size_t o_match = add(sizeof(int32_t) * nOut + sizeof(int) * 4), o_res = o_match + sizeof(int32_t) * nOut;
  const size_t inBytes = align_up_sz(used, 256);
  const size_t outBytes = sizeof(int32_t) * nOut + sizeof(int) * 4;
  const size_t o_bin = add(nOut);
  if (g_bow.reserve(used + 256, inBytes + outBytes + 256, device)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  memset(g_bow.h + o_taken, 0, sizeof(int) * b->n); memset(g_bow.h + o_hist, 0, sizeof(int) * 32 + 256);
  memset(g_bow.h + o_match, 0xFF, sizeof(int32_t) * nOut); memset(g_bow.h + o_res, 0, sizeof(int) * 4);
  for (int s = 0; s < 2; ++s) {
    const orbx_bow_side* S = sides[s];
    const int nfi = S->group_start[S->ngroups];
    if (!(s == 1 && rfB)) { memcpy(g_bow.h + offs[s][0], S->desc, (size_t)S->n * 32); memcpy(g_bow.h + offs[s][1], S->angle, sizeof(float) * S->n); }
    memcpy(g_bow.h + offs[s][2], S->valid, S->n); memcpy(g_bow.h + offs[s][3], S->node_ids, sizeof(uint32_t) * S->ngroups);
    memcpy(g_bow.h + offs[s][4], S->group_start, sizeof(int32_t) * (S->ngroups + 1)); memcpy(g_bow.h + offs[s][5], S->feat_idx, sizeof(int32_t) * nfi);
  }
  cudaStream_t st = nullptr;
  uint8_t* db = g_bow.d;
  ORBX_CUDA(cudaMemcpyAsync(db, g_bow.h, inBytes, cudaMemcpyHostToDevice, st));
  BowSideDev D[2];
  for (int s = 0; s < 2; ++s) {
    D[s].n = sides[s]->n; D[s].desc = db + offs[s][0]; D[s].angle = (const float*)(db + offs[s][1]); D[s].valid = db + offs[s][2];
    D[s].ngroups = sides[s]->ngroups; D[s].node_ids = (const uint32_t*)(db + offs[s][3]);
    D[s].group_start = (const int32_t*)(db + offs[s][4]); D[s].feat_idx = (const int32_t*)(db + offs[s][5]);
  }
  if (rfB) { D[1].desc = rfB->d_desc; D[1].angle = rfB->d_angle; }
  bow_match_kernel<<<(a->ngroups + 7) / 8, 256, 0, st>>>(D[0], D[1], mode, ratio, th_low, check_rot, (int*)(db + o_taken),
                                                         (int32_t*)(db + o_match), (int8_t*)(db + o_bin), (int*)(db + o_hist));
  uint8_t* hout = g_bow.h + inBytes;
  static const bool zcOn = [] { const char* e = getenv("ORBX_ZEROCOPY"); return e ? atoi(e) != 0 : true; }();
  uint8_t* hdev = nullptr;
  const bool zc = zcOn && nOut <= (1 << 16) && cudaHostGetDevicePointer((void**)&hdev, hout, 0) == cudaSuccess && hdev;
  launch_chain(true, bow_finish_kernel, dim3(1), dim3(1024), 0, st, nOut, check_rot, (int32_t*)(db + o_match), (const int8_t*)(db + o_bin),
               (const int*)(db + o_hist), (int*)(db + o_res), zc ? (int32_t*)hdev : (int32_t*)nullptr);
  if (!zc) ORBX_CUDA(cudaMemcpyAsync(hout, db + o_match, outBytes, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  ORBX_CUDA(cudaGetLastError());
  memcpy(match, hout, sizeof(int32_t) * nOut);
  *match_cnt = *(const int*)(hout + sizeof(int32_t) * nOut);
  return ORBX_OK;
}

extern "C" int orbx_search_by_bow(const orbx_bow_side* a, const orbx_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                                  int32_t* match, int* match_cnt, int device) {
  return bow_run(a, b, mode, ratio, th_low, check_rot, match, match_cnt, device, nullptr);
}

// Matcher::searchByBoW(KeyFrame*, Frame*, ...) (matcher.cpp:449-559, mode 0) with the frame resident on the device.
extern "C" int orbx_search_by_bow_h(const orbx_bow_side* keyframe, orbx_frame_t frame, const orbx_bow_side* frame_groups, float ratio,
                                    int th_low, int check_rot, int32_t* match, int* match_cnt) {
  if (!frame || !frame_groups) { set_error("null frame handle"); return ORBX_ERR_ARG; }
  return bow_run(keyframe, frame_groups, 0, ratio, th_low, check_rot, match, match_cnt, frame->device, frame);
}

// Matcher::searchForTriangulation(KeyFrame*, KeyFrame*, matchIdxs, F12, checkRot)  (matcher.cpp:867-1010).
// match[i] (i < a->side.n) = matched keyframe-2 feature, -1 none, -2 cleared by the rotation check; the caller builds
// matchIdxs from the entries >= 0 in ascending i (:1000-1007).
extern "C" int orbx_search_for_triangulation(const orbx_tri_side* a, const orbx_tri_side* b, const double* F12, float ex, float ey,
                                             const float* scale_factors2, int nlevels, int th_low, int check_rot, int32_t* match,
                                             int* match_cnt, int device) {
  if (!a || !b || !F12 || !scale_factors2 || nlevels < 1) { set_error("null argument"); return ORBX_ERR_ARG; }
  if (check_side(&a->side) || check_side(&b->side)) return ORBX_ERR_ARG;
  if (!match || !match_cnt || (a->side.n > 0 && (!a->kps || !a->uright)) || (b->side.n > 0 && (!b->kps || !b->uright))) {
    set_error("bad argument");
    return ORBX_ERR_ARG;
  }
  const int nOut = a->side.n;
  if (a->side.ngroups == 0 || b->side.ngroups == 0 || a->side.n == 0 || b->side.n == 0) {
    for (int i = 0; i < nOut; ++i) match[i] = -1;
    *match_cnt = 0;
    return ORBX_OK;
  }
  size_t used = 0;
  auto add = [&](size_t bytes) { used = align_up_sz(used, 256); size_t o = used; used += bytes; return o; };
  size_t off[2][8];
  const orbx_tri_side* sides[2] = {a, b};
  for (int s = 0; s < 2; ++s) {
    const orbx_bow_side* S = &sides[s]->side;
    const int nfi = S->group_start[S->ngroups];
    off[s][0] = add((size_t)S->n * 32); off[s][1] = add(sizeof(float) * S->n); off[s][2] = add(S->n);
    off[s][3] = add(sizeof(uint32_t) * S->ngroups); off[s][4] = add(sizeof(int32_t) * (S->ngroups + 1)); off[s][5] = add(sizeof(int32_t) * nfi);
    off[s][6] = add(sizeof(orbx_keypoint) * S->n); off[s][7] = add(sizeof(float) * S->n);
  }
  const size_t o_scale = add(sizeof(float) * nlevels);
  const size_t inBytes = align_up_sz(used, 256);
  const size_t o_taken = add(sizeof(int) * b->side.n), o_match = add(sizeof(int32_t) * nOut), o_bin = add(nOut), o_hist = add(sizeof(int) * 32),
               o_res = add(sizeof(int) * 4);
  if (g_bow.reserve(used + 256, inBytes, device)) { set_error("scratch allocation failed"); return ORBX_ERR_CUDA; }
  uint8_t* hb = g_bow.h;
  for (int s = 0; s < 2; ++s) {
    const orbx_bow_side* S = &sides[s]->side;
    const int nfi = S->group_start[S->ngroups];
    memcpy(hb + off[s][0], S->desc, (size_t)S->n * 32); memcpy(hb + off[s][1], S->angle, sizeof(float) * S->n);
    memcpy(hb + off[s][2], S->valid, S->n); memcpy(hb + off[s][3], S->node_ids, sizeof(uint32_t) * S->ngroups);
    memcpy(hb + off[s][4], S->group_start, sizeof(int32_t) * (S->ngroups + 1)); memcpy(hb + off[s][5], S->feat_idx, sizeof(int32_t) * nfi);
    memcpy(hb + off[s][6], sides[s]->kps, sizeof(orbx_keypoint) * S->n); memcpy(hb + off[s][7], sides[s]->uright, sizeof(float) * S->n);
  }
  memcpy(hb + o_scale, scale_factors2, sizeof(float) * nlevels);
  cudaStream_t st = nullptr;
  uint8_t* db = g_bow.d;
  ORBX_CUDA(cudaMemcpyAsync(db, hb, inBytes, cudaMemcpyHostToDevice, st));
  ORBX_CUDA(cudaMemsetAsync(db + o_taken, 0, sizeof(int) * b->side.n, st));
  ORBX_CUDA(cudaMemsetAsync(db + o_match, 0xFF, sizeof(int32_t) * nOut, st));
  ORBX_CUDA(cudaMemsetAsync(db + o_hist, 0, sizeof(int) * 32 + 256, st));
  TriSideDev D[2];
  for (int s = 0; s < 2; ++s) {
    const orbx_bow_side* S = &sides[s]->side;
    D[s].s.n = S->n; D[s].s.desc = db + off[s][0]; D[s].s.angle = (const float*)(db + off[s][1]); D[s].s.valid = db + off[s][2];
    D[s].s.ngroups = S->ngroups; D[s].s.node_ids = (const uint32_t*)(db + off[s][3]);
    D[s].s.group_start = (const int32_t*)(db + off[s][4]); D[s].s.feat_idx = (const int32_t*)(db + off[s][5]);
    D[s].kps = (const orbx_keypoint*)(db + off[s][6]); D[s].uright = (const float*)(db + off[s][7]);
  }
  TriParams T;
  for (int i = 0; i < 9; ++i) T.F[i] = F12[i];
  T.ex = ex; T.ey = ey; T.scale2 = (const float*)(db + o_scale); T.check_rot = check_rot;
  tri_match_kernel<<<(a->side.ngroups + 7) / 8, 256, 0, st>>>(D[0], D[1], T, th_low, (int*)(db + o_taken), (int32_t*)(db + o_match),
                                                              (int8_t*)(db + o_bin), (int*)(db + o_hist));
  bow_finish_kernel<<<1, 1024, 0, st>>>(nOut, check_rot, (int32_t*)(db + o_match), (const int8_t*)(db + o_bin), (const int*)(db + o_hist),
                                        (int*)(db + o_res));
  int res = 0;
  ORBX_CUDA(cudaMemcpyAsync(match, db + o_match, sizeof(int32_t) * nOut, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(&res, db + o_res, sizeof(int), cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  ORBX_CUDA(cudaGetLastError());
  *match_cnt = res;
  return ORBX_OK;
}
