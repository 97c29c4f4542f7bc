// orb_kernels.cu -- sm_100a kernels of the ORB extractor (pyramid, FAST, quadtree, blur, orientation+rBRIEF).
//
// Built with -fmad=false: the reference binary has no FMA contraction (CMakeLists.txt:4-5), and the float
// paths below (fastAtan2, descriptor rotation, keypoint scaling) must round exactly like it.
#include "orb_extract.cuh"

namespace orbx {

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

__global__ void __launch_bounds__(kFastThreads) fast_kernel(const Geom G, const Bufs B) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int cnt20[kCellsPerCta];
  __shared__ int scan_ws[40];
  const int tid = threadIdx.x;
  const int slot = blockIdx.x, f = blockIdx.y;
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
  const int sp = G.fastTileW;                   // smem pitch
  uint8_t* img = smem;                          // [fastTileH][sp]
  uint8_t* sc = smem + G.fastTileH * sp;        // corner strength, same layout
  uint8_t* fl = sc + G.fastTileH * sp;          // keep flags: bit0 = survives NMS at iniTh, bit1 = at minTh

  int pitch;
  const uint8_t* lvl = level_ptr(G, B, l, f, pitch);
  const uint8_t* base = lvl + (size_t)(kMinBorder + ty0) * pitch + kMinBorder + tx0;
  for (int i = tid; i < th * tw; i += kFastThreads) {
    const int y = i / tw, x = i - y * tw;
    img[y * sp + x] = __ldg(base + (size_t)y * pitch + x);
    sc[y * sp + x] = 0;
    fl[y * sp + x] = 0;
  }
  if (tid < kCellsPerCta) cnt20[tid] = 0;
  __syncthreads();

  const int tmin = min(G.iniTh, G.minTh);
  for (int i = tid; i < ih * iw; i += kFastThreads) {
    const int y = i / iw + 3, x = i - (y - 3) * iw + 3;
    const uint8_t* p = img + y * sp + x;
    const int c = p[0];
    const int lo = c - tmin, hi = c + tmin;
#define FCODE(v) ((int)((v) < lo) | ((int)((v) > hi) << 1))
    int m = FCODE(p[3 * sp]) | FCODE(p[-3 * sp]);
    if (!m) continue;
    m &= FCODE(p[3]) | FCODE(p[-3]);
    if (!m) continue;
    m &= FCODE(p[2 * sp + 2]) | FCODE(p[-2 * sp - 2]);
    m &= FCODE(p[-2 * sp + 2]) | FCODE(p[2 * sp - 2]);
    if (!m) continue;
    m &= FCODE(p[3 * sp + 1]) | FCODE(p[-3 * sp - 1]);
    m &= FCODE(p[sp + 3]) | FCODE(p[-sp - 3]);
    m &= FCODE(p[-sp + 3]) | FCODE(p[sp - 3]);
    m &= FCODE(p[-3 * sp + 1]) | FCODE(p[3 * sp - 1]);
    if (!m) continue;
#undef FCODE
    const int best = fast_best(p, sp, c);
    if (best > tmin) sc[y * sp + x] = (uint8_t)best;
  }
  __syncthreads();

  // NMS inside each cell's own candidate rectangle (pixels outside it count as score 0, like the zeroed
  // border rows/columns of cv::FAST on the cell ROI)
  for (int i = tid; i < ih * iw; i += kFastThreads) {
    const int yi = i / iw, xi = i - yi * iw;
    const int v = sc[(yi + 3) * sp + xi + 3];
    if (v == 0) continue;
    const int jj = xi / L.wCell, cx = xi - jj * L.wCell;
    const int cw = min(L.wCell, iw - jj * L.wCell);
    const bool hasL = cx > 0, hasR = cx + 1 < cw, hasU = yi > 0, hasD = yi + 1 < ih;
    const uint8_t* s = sc + (yi + 3) * sp + xi + 3;
    int nb[8];
    nb[0] = (hasU && hasL) ? s[-sp - 1] : 0;  nb[1] = hasU ? s[-sp] : 0;  nb[2] = (hasU && hasR) ? s[-sp + 1] : 0;
    nb[3] = hasL ? s[-1] : 0;                 nb[4] = hasR ? s[1] : 0;
    nb[5] = (hasD && hasL) ? s[sp - 1] : 0;   nb[6] = hasD ? s[sp] : 0;   nb[7] = (hasD && hasR) ? s[sp + 1] : 0;
    int flags = 0;
    if (v > G.iniTh) {
      bool keep = true;
#pragma unroll
      for (int k = 0; k < 8; ++k) keep &= !(nb[k] > G.iniTh && nb[k] >= v);
      if (keep) { flags |= 1; atomicAdd(&cnt20[jj], 1); }
    }
    if (v > G.minTh) {
      bool keep = true;
#pragma unroll
      for (int k = 0; k < 8; ++k) keep &= !(nb[k] > G.minTh && nb[k] >= v);
      if (keep) flags |= 2;
    }
    fl[(yi + 3) * sp + xi + 3] = (uint8_t)flags;
  }
  __syncthreads();

  // ordered compaction: cells left to right, inside a cell row-major (the order cv::FAST emits and the
  // reference appends, ORBextractor.cpp:826-834)
  const int ncell = j1 - j0;
  const int cellSeq = ih * L.wCell;
  const int total = ncell * cellSeq;
  const int per = (total + kFastThreads - 1) / kFastThreads;
  const int s0 = min(tid * per, total), s1 = min(s0 + per, total);
  auto kept = [&](int s, int& xi, int& yi) -> bool {
    const int jj = s / cellSeq, r = s - jj * cellSeq;
    yi = r / L.wCell;
    const int cx = r - yi * L.wCell;
    xi = jj * L.wCell + cx;
    if (xi >= iw) return false;
    const int need = (cnt20[jj] > 0) ? 1 : 2;
    return (fl[(yi + 3) * sp + xi + 3] & need) != 0;
  };
  int mine = 0, xi, yi;
  for (int s = s0; s < s1; ++s) mine += kept(s, xi, yi) ? 1 : 0;
  // block exclusive scan of `mine`
  __shared__ int offs[kFastThreads];
  offs[tid] = mine;
  __syncthreads();
  const int totalKept = block_exclusive_scan(offs, kFastThreads, scan_ws);
  int pos = offs[tid];
  uint32_t* out = B.slotKeys + (size_t)f * G.slotKeysPerFrame + __ldg(B.slotKeyBase + slot);
  for (int s = s0; s < s1; ++s) {
    if (kept(s, xi, yi)) {
      const int best = sc[(yi + 3) * sp + xi + 3];
      out[pos++] = pack_key(tx0 + xi + 3, ty0 + yi + 3, best - 1);
    }
  }
  if (tid == 0) *out_count = totalKept;
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

__device__ __forceinline__ int quadrant_of(const short4 b, int x, int y) {
  const int mx = b.x + ((b.z - b.x + 1) >> 1);   // ceil(float(x1-x0)/2)   (ORBextractor.cpp:489-490)
  const int my = b.y + ((b.w - b.y + 1) >> 1);
  return (x < mx ? 0 : 1) | (y < my ? 0 : 2);     // 0:n1 1:n2 2:n3 3:n4   (:521-531)
}

__global__ void __launch_bounds__(kOctThreads) octree_kernel(const Geom G, const Bufs B) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, T = kOctThreads;
  const int l = blockIdx.x, f = blockIdx.y;
  const LevelGeom& L = G.L[l];
  const int cap = G.nodeCap;
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
  __shared__ int sh_stop, sh_nexp;
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
  {
    const uint32_t* slotKeys = B.slotKeys + (size_t)f * G.slotKeysPerFrame;
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = warp; s < L.nSlots; s += T / 32) {
      const int c = slotCount[s], o = S.slotPre[s];
      const uint32_t* src = slotKeys + __ldg(B.slotKeyBase + L.slot0 + s);
      for (int k = lane; k < c; k += 32) keys[o + k] = src[k];
    }
  }
  if (tid < kMaxRoots) rootCnt[tid] = 0;
  __syncthreads();

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
  while (!finish) {
    const short4* box = S.box[cur];
    const int* cnt = S.cnt[cur];
    short4* nbox = S.box[cur ^ 1];
    int* ncnt = S.cnt[cur ^ 1];
    const int Sn = Scount;

    for (int i = tid; i < Sn * 4; i += T) S.cnt4[i] = 0;
    if (tid == 0) { sh_stop = 0x7fffffff; sh_nexp = 0; }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
      const int p = nodeOf[k];
      if (cnt[p] > 1) {
        const uint32_t key = keys[k];
        atomicAdd(&S.cnt4[p * 4 + quadrant_of(box[p], key_x(key), key_y(key))], 1);
      }
    }
    __syncthreads();

    // processing rank of every expandable node
    for (int p = tid; p < Sn; p += T) S.rank[p] = cnt[p] > 1 ? 1 : 0;
    __syncthreads();
    const int E = block_exclusive_scan(S.rank, Sn, S.ws);   // phase 1: list order
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
    // children per processing rank
    for (int p = tid; p < Sn; p += T) {
      if (cnt[p] > 1) {
        const int* c4 = S.cnt4 + p * 4;
        S.childR[S.rank[p]] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
      }
    }
    __syncthreads();
    // keep a copy of per-rank child counts in keepPos (borrowed) before the scan overwrites them
    for (int t = tid; t < E; t += T) S.keepPos[t] = S.childR[t];
    __syncthreads();
    const int allChildren = block_exclusive_scan(S.childR, E, S.ws);
    if (tid == 0) S.childR[E] = allChildren;
    int Pn = E;
    if (phase == 2) {
      // first rank t whose split makes the list reach N:  Sn + (children up to and incl. t) - (t+1) >= N
      for (int t = tid; t < E; t += T) {
        const int incl = S.childR[t] + S.keepPos[t];
        if (Sn + incl - (t + 1) >= N) atomicMin(&sh_stop, t);
      }
    }
    __syncthreads();
    if (phase == 2 && sh_stop != 0x7fffffff) Pn = sh_stop + 1;
    const int totalChildren = S.childR[Pn];
    __syncthreads();   // everyone has read childR[Pn] / keepPos before keepPos is rewritten below

    // positions of the children blocks and of the surviving old nodes
    for (int p = tid; p < Sn; p += T) {
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
            if (c4[q] > 1) atomicAdd(&sh_nexp, 1);
            ++o;
          } else cp[q] = -1;
        }
      }
    }
    __syncthreads();
    // surviving nodes keep their relative order behind the children
    for (int p = tid; p < Sn; p += T) S.keepPos[p] = (cnt[p] > 1 && S.rank[p] < Pn) ? 0 : 1;
    __syncthreads();
    const int nKept = block_exclusive_scan(S.keepPos, Sn, S.ws);
    for (int p = tid; p < Sn; p += T) {
      const bool split = cnt[p] > 1 && S.rank[p] < Pn;
      if (!split) {
        const int o = totalChildren + S.keepPos[p];
        nbox[o] = box[p];
        ncnt[o] = cnt[p];
      } else {
        S.keepPos[p] = -1;
      }
    }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
      const int p = nodeOf[k];
      const int kp = S.keepPos[p];
      if (kp >= 0) {
        nodeOf[k] = (uint16_t)(totalChildren + kp);
      } else {
        const uint32_t key = keys[k];
        nodeOf[k] = (uint16_t)S.childPos[p * 4 + quadrant_of(box[p], key_x(key), key_y(key))];
      }
    }
    const int newS = totalChildren + nKept;
    const int nToExpand = sh_nexp;
    __syncthreads();
    cur ^= 1;
    Scount = newS;
    // termination (ORBextractor.cpp:672-744)
    if (newS >= N || newS == Sn) finish = true;
    else if (phase == 1 && newS + 3 * nToExpand > N) phase = 2;
  }

  // ---- strongest key per node, first wins ties (ORBextractor.cpp:747-766) ------------------------------
  for (int p = tid; p < Scount; p += T) S.best[p] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    const unsigned long long v = ((unsigned long long)(key_s(keys[k]) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
    atomicMax(&S.best[nodeOf[k]], v);
  }
  __syncthreads();
  for (int p = tid; p < Scount; p += T) {
    const uint32_t k = 0xFFFFFFFFu - (uint32_t)(S.best[p] & 0xFFFFFFFFull);
    sel[p] = keys[k];
  }
  if (tid == 0) *selCount = Scount;
}

// ======================================================================================================
// K5  cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101, 8.8 fixed point (ORBextractor.cpp:1093-1094;
// arithmetic: SURVEY App. A.3).  All levels of all frames of the chunk in one launch; one CTA = one
// 64x32 output tile, horizontal pass into shared u16, vertical pass out of it.
// ======================================================================================================
constexpr int kBlurTW = 64, kBlurTH = 32;

__device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) blur_kernel(const Geom G, const Bufs B) {
  __shared__ uint8_t tin[kBlurTH + 6][kBlurTW + 8];
  __shared__ uint16_t hbuf[kBlurTH + 6][kBlurTW];
  const int tid = threadIdx.x, f = blockIdx.y;
  int l = 0;
  while (l + 1 < G.nlevels && (int)blockIdx.x >= G.L[l + 1].blurTile0) ++l;
  const LevelGeom& L = G.L[l];
  const int t = blockIdx.x - L.blurTile0;
  const int ty = t / L.blurTilesX, tx = t - ty * L.blurTilesX;
  const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
  int pitch;
  const uint8_t* src = level_ptr(G, B, l, f, pitch);
  for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
    const int r = i / (kBlurTW + 6), c = i - r * (kBlurTW + 6);
    const int sy = reflect101(min(y0 + r - 3, L.h + 2), L.h);
    const int sx = reflect101(min(x0 + c - 3, L.w + 2), L.w);
    tin[r][c] = __ldg(src + (size_t)sy * pitch + sx);
  }
  __syncthreads();
  for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
    const int r = i / kBlurTW, c = i - r * kBlurTW;
    const uint8_t* p = &tin[r][c];
    hbuf[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
  }
  __syncthreads();
  uint8_t* dst = B.blur + L.blurOff + (size_t)f * L.h * L.bpitch;
  for (int i = tid; i < kBlurTH * kBlurTW; i += 256) {
    const int r = i / kBlurTW, c = i - r * kBlurTW;
    const int y = y0 + r, x = x0 + c;
    if (y < L.h && x < L.w) {
      const uint32_t acc = 18u * (hbuf[r][c] + hbuf[r + 6][c]) + 34u * (hbuf[r + 1][c] + hbuf[r + 5][c]) +
                           48u * (hbuf[r + 2][c] + hbuf[r + 4][c]) + 56u * hbuf[r + 3][c] + 32768u;
      dst[(size_t)y * L.bpitch + x] = (uint8_t)(acc >> 16);
    }
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

__global__ void __launch_bounds__(256) orient_desc_kernel(const Geom G, const Bufs B, orbx_keypoint* __restrict__ kps_out,
                                                          uint8_t* __restrict__ desc_out, int cap,
                                                          int32_t* __restrict__ counts_out, int frame0) {
  const int lane = threadIdx.x & 31;
  const int slotIdx = blockIdx.x * 8 + (threadIdx.x >> 5);   // index into the per-frame selected array
  const int f = blockIdx.y;
  if (slotIdx >= G.selPerFrame) return;
  int l = 0;
  while (l + 1 < G.nlevels && slotIdx >= G.L[l + 1].selBase) ++l;
  const LevelGeom& L = G.L[l];
  const int i = slotIdx - L.selBase;
  const int* selCount = B.selCount + (size_t)f * G.nlevels;
  int before = 0, total = 0;
  for (int q = 0; q < G.nlevels; ++q) {
    const int c = selCount[q];
    if (q < l) before += c;
    total += c;
  }
  if (slotIdx == 0 && lane == 0) counts_out[frame0 + f] = total;
  if (i >= selCount[l]) return;
  const int o = before + i;
  if (o >= cap) return;   // caller buffer smaller than the keypoint count: count is still reported

  const uint32_t key = B.sel[(size_t)f * G.selPerFrame + slotIdx];
  const int x = key_x(key) + kMinBorder, y = key_y(key) + kMinBorder;   // (ORBextractor.cpp:851-852)
  int pitch;
  const uint8_t* img = level_ptr(G, B, l, f, pitch);

  // ---- moments over the radius-15 disc: lane <-> row v = lane-15 --------------------------------------
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int v = lane - 15;
    const int d = __ldg(B.umax + abs(v));
    const uint8_t* row = img + (size_t)(y + v) * pitch + x;
    int sum = 0;
    for (int u = -d; u <= d; ++u) {
      const int p = __ldg(row + u);
      m10 += u * p;
      sum += p;
    }
    m01 = v * sum;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- rBRIEF on the blurred level ---------------------------------------------------------------------
  const float factorPI = (float)(3.1415926535897932384626433832795 / (double)180.f);
  float a, b;
  glibc_sincosf(__fmul_rn(angle, factorPI), b, a);
  const uint8_t* bl = B.blur + L.blurOff + (size_t)f * L.h * L.bpitch + (size_t)y * L.bpitch + x;
  const float2* pat = B.pattern + lane * 16;
  int val = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 p0 = __ldg(pat + 2 * j), p1 = __ldg(pat + 2 * j + 1);
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b)));
    const int t0 = __ldg(bl + r0 * L.bpitch + c0), t1 = __ldg(bl + r1 * L.bpitch + c1);
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

// ---- launch wrappers (called from the C ABI in orb_capi.cu) ---------------------------------------------
void launch_resize(const uint8_t* src, int sw, int sh, int spitch, size_t sframe, uint8_t* dst, int dw, int dh,
                   int dpitch, size_t dframe, const ResizeTaps& T, int nframes, cudaStream_t st) {
  dim3 block(32, 8), grid((dw + 127) / 128, (dh + 7) / 8, nframes);
  resize_kernel<<<grid, block, 0, st>>>(src, sw, sh, spitch, sframe, dst, dw, dh, dpitch, dframe, T);
}

size_t fast_smem_bytes(const Geom& G) { return (size_t)3 * G.fastTileW * G.fastTileH; }

size_t octree_smem_bytes(const Geom& G) {
  const size_t cap = G.nodeCap;
  return 8 * cap + 8 * cap * 2 + 4 * cap * 2 + 16 * cap + 4 * cap + 4 * (cap + 1) + 4 * cap + 16 * cap +
         4 * (size_t)(G.maxSlotsPerLevel + 1) + 4 * 40 + 64;
}

cudaError_t configure_kernels(const Geom& G) {
  cudaError_t e = cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(G));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem_bytes(G));
}

void launch_fast(const Geom& G, const Bufs& B, int nframes, cudaStream_t st) {
  fast_kernel<<<dim3(G.totalSlots, nframes), kFastThreads, fast_smem_bytes(G), st>>>(G, B);
}
void launch_octree(const Geom& G, const Bufs& B, int nframes, cudaStream_t st) {
  octree_kernel<<<dim3(G.nlevels, nframes), kOctThreads, octree_smem_bytes(G), st>>>(G, B);
}
void launch_blur(const Geom& G, const Bufs& B, int nframes, cudaStream_t st) {
  blur_kernel<<<dim3(G.blurTiles, nframes), 256, 0, st>>>(G, B);
}
void launch_orient_desc(const Geom& G, const Bufs& B, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts,
                        int frame0, int nframes, cudaStream_t st) {
  orient_desc_kernel<<<dim3((G.selPerFrame + 7) / 8, nframes), 256, 0, st>>>(G, B, kps, desc, cap, counts, frame0);
}

}  // namespace orbx
