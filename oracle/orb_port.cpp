// oracle/orb_port.cpp -- TEST INFRASTRUCTURE (CPU oracle "port"), not product code.
//
// Independent CPU restatement of the reference's ORB front-end hot path, written from the
// behaviour of /root/reference/src/ORBextractor.cpp and src/matcher.cpp / src/frame.cpp (never
// copied from them).  It is deliberately phrased the way the CUDA kernels are (flat arrays, node
// ids, order-independent reductions), so every GPU stage has a stage-level CPU twin.
//
// Pinning (see DESIGN.md "Oracle"): the primitives in cvprims.h are asserted bit-equal to the real
// cv2 4.13.0; this port is asserted bit-equal to the reference's own ORBextractor.cpp compiled in
// place (oracle/_ref/liborbref_parity.so) in tests/test_oracle_vs_reference.py, and both are
// asserted against the committed fixtures in tests/golden/.
// Matcher part: the reference's OWN src/matcher.cpp is compiled in place too (oracle/_ref/libmatcherref.so: its own
// include/myslam/matcher.h, with Frame / KeyFrame / MapPoint / SE3 / Sim3 / FeatureVector replaced by the stand-in types of
// oracle/compat_myslam because Eigen, Sophus and DBoW3 are not installed); tests/test_matcher_adapter.py asserts that all
// eleven Matcher entry points, run on identical object graphs, leave the same pointers and counts as this port behind the
// C++ adapters (164 scenes in the CPU suite).  The small helpers underneath -- Frame::assignFeaturesToGrid / getFeaturesInArea /
// findDepth, KeyFrame::getFeaturesInArea / isInImg, MapPoint::predictScale / computeDescriptor, Camera::camera2pixel -- are
// pinned the same way: their line ranges are compiled in place (oracle/ref_helpers_wrap.cpp -> _ref/librefhelpers.so) and
// held against this port and against the stand-in types (tests/test_oracle_helpers_vs_reference.py), Frame::undistortKeyPoints
// included (its loop is the reference's; the cv::undistortPoints underneath is the restatement pinned against cv2 4.13.0,
// tests/golden/cv2_undistort.npz).
//
// Reference lines followed:
//   ctor tables            ORBextractor.cpp:414-476          -> port_tables()
//   ComputePyramid         ORBextractor.cpp:1115-1142        -> port_pyramid()
//   cell loop + FAST retry ORBextractor.cpp:771-837          -> port_fast_cells()
//   DistributeOctTree      ORBextractor.cpp:487-769          -> port_octree()
//   IC_Angle               ORBextractor.cpp:79-107           -> port_ic_angle()
//   blur + descriptors     ORBextractor.cpp:110-151,1085-1111-> port_blur(), port_descriptor()
//   operator()             ORBextractor.cpp:1051-1112        -> port_extract()
//   computeDistance        matcher.cpp:1240-1256             -> port_hamming()
//   top-2 + ratio loop     matcher.cpp:481-507               -> port_knn2()
//   grid build / query     frame.cpp:72-97,199-247           -> port_grid_build(), port_features_in_area()
//   searchByProjection(F,F)      matcher.cpp:18-148          -> port_sbp_frame()
//   searchByProjection(F,local)  matcher.cpp:274-353         -> port_sbp_local()
//   computeThreeMax        matcher.cpp:1258-1304             -> three_max()
//   searchByProjection(F,KF)     matcher.cpp:150-272         -> port_sbp_reloc()
//   searchByProjection(KF,Sim3)  matcher.cpp:356-447         -> port_sbp_sim3()
//   searchBySim3 / fuse cores    matcher.cpp:679-865,1012-1238 -> port_window_argmin(), port_search_by_sim3()
//   searchForTriangulation       matcher.cpp:867-1010,1306-1324 -> port_search_for_triangulation()
//   MapPoint::computeDescriptor  mappoint.cpp:118-179        -> port_medoid()
//   searchByBoW (both)     matcher.cpp:449-559, 561-677      -> port_search_by_bow()
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <thread>

#include "cvprims.h"
#include "orb_pattern.h"

#define CV_PI_PORT 3.1415926535897932384626433832795

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

struct Tables {
  int nlevels;
  std::vector<float> scale, inv_scale;
  std::vector<int> nfeat;
  int umax[16];
};

// ORBextractor.cpp:414-476.  `scaleFactor` is a double member holding the float argument.
void make_tables(int nfeatures, float scaleFactorF, int nlevels, Tables& t) {
  const double scaleFactor = scaleFactorF;
  t.nlevels = nlevels;
  t.scale.assign(nlevels, 1.f);
  t.inv_scale.assign(nlevels, 1.f);
  for (int i = 1; i < nlevels; ++i) t.scale[i] = (float)(t.scale[i - 1] * scaleFactor);
  for (int i = 0; i < nlevels; ++i) t.inv_scale[i] = 1.0f / t.scale[i];
  t.nfeat.assign(nlevels, 0);
  float factor = (float)(1.0f / scaleFactor);
  float want = (float)(nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels)));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; ++l) {
    t.nfeat[l] = cvp::cv_round(want);
    sum += t.nfeat[l];
    want *= factor;
  }
  t.nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
  // quarter-circle extents of the radius-15 patch
  int vmax = cvp::cv_floor(kHalfPatch * sqrt(2.f) / 2 + 1);
  int vmin = cvp::cv_ceil(kHalfPatch * sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (int v = 0; v <= vmax; ++v) t.umax[v] = cvp::cv_round(sqrt(hp2 - v * v));
  for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
    t.umax[v] = v0;
    ++v0;
  }
}

struct Level {
  int w, h;
  std::vector<uint8_t> px;  // tight, no border (the 19-px border is never read: SURVEY App. B.7)
};

void level_size(const Tables& t, int W, int H, int l, int& w, int& h) {
  float s = t.inv_scale[l];
  w = cvp::cv_round((float)W * s);
  h = cvp::cv_round((float)H * s);
}

struct Cand { int x, y, score; };  // coordinates relative to (minBorderX, minBorderY) = (16,16)

// ORBextractor.cpp:771-837: per-cell FAST with the ini/min threshold retry.
void fast_cells(const uint8_t* img, int w, int h, size_t step, int iniTh, int minTh, std::vector<Cand>& out) {
  out.clear();
  const int minB = kEdge - 3;
  const int maxBX = w - kEdge + 3, maxBY = h - kEdge + 3;
  const int width = maxBX - minB, height = maxBY - minB;
  const int nCols = width / 30, nRows = height / 30;
  if (nCols <= 0 || nRows <= 0) return;
  const int wCell = (width + nCols - 1) / nCols;
  const int hCell = (height + nRows - 1) / nRows;
  std::vector<cvp::FastKp> cell;
  for (int i = 0; i < nRows; ++i) {
    const int iniY = minB + i * hCell;
    int maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = maxBY;
    for (int j = 0; j < nCols; ++j) {
      const int iniX = minB + j * wCell;
      int maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = maxBX;
      const uint8_t* roi = img + (size_t)iniY * step + iniX;
      cvp::fast9_16(roi, maxX - iniX, maxY - iniY, step, iniTh, true, cell);
      if (cell.empty()) cvp::fast9_16(roi, maxX - iniX, maxY - iniY, step, minTh, true, cell);
      for (const auto& k : cell) out.push_back(Cand{k.x + j * wCell, k.y + i * hCell, k.score});
    }
  }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree (ORBextractor.cpp:545-769) as flat arrays.
//
// Facts used (each checked against the compiled reference in tests):
//  * a node is expandable iff it holds more than one key;
//  * every pass pushes the children of the processed parents to the list front, so the new list is
//    [children of the last processed parent as n4,n3,n2,n1] ... [children of the first] followed by
//    the untouched nodes in their old order;
//  * the "largest first" pass sorts by (size, node address); with addresses increasing in creation
//    order (the canonical rule, realised in the reference by a monotonic allocator) processing from
//    the back of that sort is a STABLE sort of the current list by descending size;
//  * the retained key of a node is its first maximum-response key in candidate order, which is an
//    order-independent reduction, so keys never have to be physically partitioned.
// ---------------------------------------------------------------------------------------------
struct QNode { int x0, y0, x1, y1, cnt; };

void octree(const std::vector<Cand>& keys, int regionW, int regionH, int N, std::vector<int>& selected) {
  selected.clear();
  const int n = (int)keys.size();
  const int nIni = (int)roundf((float)regionW / (float)regionH);
  if (nIni < 1) return;  // reference divides by zero here; the product rejects such shapes
  const float hX = (float)regionW / nIni;
  std::vector<QNode> list;
  std::vector<int> nodeOf(n);
  {
    std::vector<QNode> roots(nIni);
    for (int i = 0; i < nIni; ++i) roots[i] = QNode{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), regionH, 0};
    std::vector<int> rootOf(n);
    for (int k = 0; k < n; ++k) {
      int r = (int)((float)keys[k].x / hX);
      rootOf[k] = r;
      roots[r].cnt++;
    }
    std::vector<int> pos(nIni, -1);
    for (int i = 0; i < nIni; ++i)
      if (roots[i].cnt > 0) { pos[i] = (int)list.size(); list.push_back(roots[i]); }
    for (int k = 0; k < n; ++k) nodeOf[k] = pos[rootOf[k]];
  }

  auto quadrant = [&](const QNode& nd, const Cand& k) {
    const int mx = nd.x0 + (nd.x1 - nd.x0 + 1) / 2;  // ceil(float(x1-x0)/2)
    const int my = nd.y0 + (nd.y1 - nd.y0 + 1) / 2;
    return (k.x < mx) ? (k.y < my ? 0 : 2) : (k.y < my ? 1 : 3);
  };

  // One pass: split parents `order[0..)` (list positions) in that order; if stopAtN, stop right after
  // the split that makes the list reach N nodes.
  auto pass = [&](const std::vector<int>& order, bool stopAtN) {
    const int S = (int)list.size();
    std::vector<int> cnt4((size_t)S * 4, 0);
    for (int k = 0; k < n; ++k) {
      const QNode& nd = list[nodeOf[k]];
      if (nd.cnt > 1) cnt4[(size_t)nodeOf[k] * 4 + quadrant(nd, keys[k])]++;
    }
    int size = S, processed = 0;
    for (size_t t = 0; t < order.size(); ++t) {
      int p = order[t], nonempty = 0;
      for (int q = 0; q < 4; ++q) nonempty += cnt4[(size_t)p * 4 + q] > 0;
      size += nonempty - 1;
      processed = (int)t + 1;
      if (stopAtN && size >= N) break;
    }
    std::vector<QNode> nl;
    nl.reserve(size);
    std::vector<int> childPos((size_t)S * 4, -1), keepPos(S, -1);
    std::vector<char> split(S, 0);
    for (int t = processed - 1; t >= 0; --t) {
      int p = order[t];
      split[p] = 1;
      const QNode& nd = list[p];
      const int mx = nd.x0 + (nd.x1 - nd.x0 + 1) / 2, my = nd.y0 + (nd.y1 - nd.y0 + 1) / 2;
      for (int q = 3; q >= 0; --q) {
        int c = cnt4[(size_t)p * 4 + q];
        if (!c) continue;
        QNode ch;
        ch.x0 = (q & 1) ? mx : nd.x0; ch.x1 = (q & 1) ? nd.x1 : mx;
        ch.y0 = (q & 2) ? my : nd.y0; ch.y1 = (q & 2) ? nd.y1 : my;
        ch.cnt = c;
        childPos[(size_t)p * 4 + q] = (int)nl.size();
        nl.push_back(ch);
      }
    }
    for (int p = 0; p < S; ++p)
      if (!split[p]) { keepPos[p] = (int)nl.size(); nl.push_back(list[p]); }
    for (int k = 0; k < n; ++k) {
      int p = nodeOf[k];
      nodeOf[k] = split[p] ? childPos[(size_t)p * 4 + quadrant(list[p], keys[k])] : keepPos[p];
    }
    list.swap(nl);
  };

  auto expandable = [&]() {
    std::vector<int> e;
    for (int p = 0; p < (int)list.size(); ++p) if (list[p].cnt > 1) e.push_back(p);
    return e;
  };

  bool finish = list.empty();
  while (!finish) {
    int prev = (int)list.size();
    pass(expandable(), false);
    int S = (int)list.size();
    int nToExpand = (int)expandable().size();
    if (S >= N || S == prev) {
      finish = true;
    } else if (S + nToExpand * 3 > N) {
      while (!finish) {
        prev = (int)list.size();
        std::vector<int> e = expandable();
        std::stable_sort(e.begin(), e.end(), [&](int a, int b) { return list[a].cnt > list[b].cnt; });
        pass(e, true);
        S = (int)list.size();
        if (S >= N || S == prev) finish = true;
      }
    }
  }

  const int S = (int)list.size();
  std::vector<int> best(S, -1);
  for (int k = 0; k < n; ++k) {
    int p = nodeOf[k];
    if (best[p] < 0 || keys[k].score > keys[best[p]].score) best[p] = k;
  }
  selected.assign(best.begin(), best.end());
}

// ORBextractor.cpp:79-107
float ic_angle(const uint8_t* img, size_t step, int x, int y, const int* umax) {
  const uint8_t* c = img + (size_t)y * step + x;
  int m01 = 0, m10 = 0;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vs = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int p = c[u + (ptrdiff_t)v * (ptrdiff_t)step], m = c[u - (ptrdiff_t)v * (ptrdiff_t)step];
      vs += p - m;
      m10 += u * (p + m);
    }
    m01 += v * vs;
  }
  return cvp::fast_atan2((float)m01, (float)m10);
}

// ORBextractor.cpp:110-151 (glibc cosf/sinf, round-half-even, no FMA)
void descriptor(const uint8_t* img, size_t step, int x, int y, float angleDeg, uint8_t* out) {
  const float factorPI = (float)(CV_PI_PORT / 180.f);
  float ang = angleDeg * factorPI;
  float a = cosf(ang), b = sinf(ang);
  const uint8_t* c = img + (size_t)y * step + x;
  const int* pat = orb_bit_pattern_31;
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int j = 0; j < 8; ++j) {
      const float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1];
      const float x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
      int r0 = cvp::cv_round(x0 * b + y0 * a), c0 = cvp::cv_round(x0 * a - y0 * b);
      int r1 = cvp::cv_round(x1 * b + y1 * a), c1 = cvp::cv_round(x1 * a - y1 * b);
      int t0 = c[(ptrdiff_t)r0 * (ptrdiff_t)step + c0], t1 = c[(ptrdiff_t)r1 * (ptrdiff_t)step + c1];
      val |= (t0 < t1) << j;
    }
    out[i] = (uint8_t)val;
  }
}

struct PortKp { float x, y, size, angle, response; int octave, class_id; };
static_assert(sizeof(PortKp) == 28, "keypoint layout");

// matcher.cpp:1240-1256
inline int hamming256(const uint8_t* a, const uint8_t* b) {
  const uint32_t* p = (const uint32_t*)a;
  const uint32_t* q = (const uint32_t*)b;
  int d = 0;
  for (int i = 0; i < 8; ++i) d += __builtin_popcount(p[i] ^ q[i]);
  return d;
}

// matcher.cpp:1258-1304
void three_max(const int* hist, int L, int& i1, int& i2, int& i3) {
  int m1 = 0, m2 = 0, m3 = 0;
  i1 = i2 = i3 = -1;
  for (int i = 0; i < L; ++i) {
    const int s = hist[i];
    if (s > m1) { m3 = m2; i3 = i2; m2 = m1; i2 = i1; m1 = s; i1 = i; }
    else if (s > m2) { m3 = m2; i3 = i2; m2 = s; i2 = i; }
    else if (s > m3) { m3 = s; i3 = i; }
  }
  if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
  else if (m3 < 0.1f * (float)m1) { i3 = -1; }
}

struct Grid {
  // frame.cpp:72-97; camera.h:8-9; camera.cpp:42-48
  static const int COLS = 64, ROWS = 48;
  float xMin, yMin, gw, gh;
  std::vector<std::vector<int>> cell;  // [ix*ROWS + iy]
  void build(const PortKp* kps, int n, float xmin, float xmax, float ymin, float ymax) {
    xMin = xmin; yMin = ymin;
    gw = (float)COLS / (xmax - xmin);
    gh = (float)ROWS / (ymax - ymin);
    cell.assign(COLS * ROWS, std::vector<int>());
    for (int i = 0; i < n; ++i) {
      const int gx = (int)roundf((kps[i].x - xMin) * gw);
      const int gy = (int)roundf((kps[i].y - yMin) * gh);
      if (gx < 0 || gx >= COLS || gy < 0 || gy >= ROWS) continue;
      cell[gx * ROWS + gy].push_back(i);
    }
  }
  // frame.cpp:199-247
  void area(const PortKp* kps, float u, float v, float r, int minL, int maxL, std::vector<int>& out) const {
    out.clear();
    const int x0 = std::max(0, (int)floorf((u - xMin - r) * gw));
    if (x0 >= COLS) return;
    const int x1 = std::min(COLS - 1, (int)floorf((u - xMin + r) * gw));
    if (x1 < 0) return;
    const int y0 = std::max(0, (int)floorf((v - yMin - r) * gh));
    if (y0 >= ROWS) return;
    const int y1 = std::min(ROWS - 1, (int)floorf((v - yMin + r) * gh));
    if (y1 < 0) return;
    for (int ix = x0; ix <= x1; ++ix)
      for (int iy = y0; iy <= y1; ++iy)
        for (int id : cell[ix * ROWS + iy]) {
          const PortKp& k = kps[id];
          if (k.octave < minL || k.octave > maxL) continue;
          if (fabsf(k.x - u) < r && fabsf(k.y - v) < r) out.push_back(id);
        }
  }
};

}  // namespace

extern "C" {

struct port_params { int nfeatures; float scale_factor; int nlevels; int ini_th; int min_th; };

// Host tables exactly as the reference constructor builds them.
int port_tables(const port_params* p, float* scale, float* inv_scale, int* nfeat, int* umax) {
  Tables t;
  make_tables(p->nfeatures, p->scale_factor, p->nlevels, t);
  for (int i = 0; i < p->nlevels; ++i) { scale[i] = t.scale[i]; inv_scale[i] = t.inv_scale[i]; nfeat[i] = t.nfeat[i]; }
  for (int i = 0; i < 16; ++i) umax[i] = t.umax[i];
  return 0;
}

int port_level_size(const port_params* p, int W, int H, int level, int* w, int* h) {
  Tables t;
  make_tables(p->nfeatures, p->scale_factor, p->nlevels, t);
  level_size(t, W, H, level, *w, *h);
  return 0;
}

void port_resize(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep) {
  cvp::resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void port_blur(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
  cvp::gaussian_blur7_s2_u8(src, w, h, sstep, dst, dstep);
}
float port_fast_atan2(float y, float x) { return cvp::fast_atan2(y, x); }
void port_sincosf(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }

// cv::FAST on one image/ROI.  out: triples (x, y, score); returns count (may exceed cap; only cap written).
int port_fast(const uint8_t* img, int w, int h, size_t step, int th, int nms, int* out, int cap) {
  std::vector<cvp::FastKp> k;
  cvp::fast9_16(img, w, h, step, th, nms != 0, k);
  for (int i = 0; i < (int)k.size() && i < cap; ++i) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].score; }
  return (int)k.size();
}

// Per-cell FAST over one level.  out: triples (x, y, score) relative to (16,16).
int port_fast_cells(const uint8_t* img, int w, int h, size_t step, int iniTh, int minTh, int* out, int cap) {
  std::vector<Cand> c;
  fast_cells(img, w, h, step, iniTh, minTh, c);
  for (int i = 0; i < (int)c.size() && i < cap; ++i) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].score; }
  return (int)c.size();
}

// Quadtree selection.  cand: n triples; sel: indices into cand in output (list) order.
int port_octree(const int* cand, int n, int regionW, int regionH, int N, int* sel, int cap) {
  std::vector<Cand> keys(n);
  for (int i = 0; i < n; ++i) keys[i] = Cand{cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]};
  std::vector<int> s;
  octree(keys, regionW, regionH, N, s);
  for (int i = 0; i < (int)s.size() && i < cap; ++i) sel[i] = s[i];
  return (int)s.size();
}

float port_ic_angle(const uint8_t* img, size_t step, int x, int y) {
  Tables t;
  make_tables(1000, 1.2f, 8, t);
  return ic_angle(img, step, x, y, t.umax);
}

void port_descriptor(const uint8_t* blurred, size_t step, int x, int y, float angle, uint8_t* out32) {
  descriptor(blurred, step, x, y, angle, out32);
}

// Full extractor.  kps: 28-byte cv::KeyPoint layout; desc: n x 32.  Returns keypoint count, or <0 on error.
// If `levels_out` is non-null it receives the tight (border-less) pyramid levels back to back.
int port_extract(const port_params* p, const uint8_t* img, int W, int H, size_t stride, void* kps_out,
                 uint8_t* desc_out, int cap, uint8_t* levels_out) {
  if (!img || W <= 0 || H <= 0) return 0;
  Tables t;
  make_tables(p->nfeatures, p->scale_factor, p->nlevels, t);
  std::vector<Level> pyr(t.nlevels);
  for (int l = 0; l < t.nlevels; ++l) {
    level_size(t, W, H, l, pyr[l].w, pyr[l].h);
    if (pyr[l].w <= 0 || pyr[l].h <= 0) return -2;
    pyr[l].px.resize((size_t)pyr[l].w * pyr[l].h);
    if (l == 0) {
      for (int y = 0; y < H; ++y) memcpy(&pyr[0].px[(size_t)y * W], img + (size_t)y * stride, W);
    } else {
      cvp::resize_linear_u8(pyr[l - 1].px.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, pyr[l].px.data(),
                            pyr[l].w, pyr[l].h, pyr[l].w);
    }
  }
  if (levels_out) {
    size_t off = 0;
    for (int l = 0; l < t.nlevels; ++l) { memcpy(levels_out + off, pyr[l].px.data(), pyr[l].px.size()); off += pyr[l].px.size(); }
  }
  PortKp* kps = (PortKp*)kps_out;
  int total = 0;
  std::vector<Cand> cand;
  std::vector<int> sel;
  std::vector<uint8_t> blurred;
  for (int l = 0; l < t.nlevels; ++l) {
    const Level& L = pyr[l];
    fast_cells(L.px.data(), L.w, L.h, L.w, p->ini_th, p->min_th, cand);
    const int minB = kEdge - 3;
    octree(cand, (L.w - kEdge + 3) - minB, (L.h - kEdge + 3) - minB, t.nfeat[l], sel);
    if (sel.empty()) continue;
    blurred.resize(L.px.size());
    cvp::gaussian_blur7_s2_u8(L.px.data(), L.w, L.h, L.w, blurred.data(), L.w);
    const float size = (float)(int)(kPatch * t.scale[l]);
    for (int id : sel) {
      const int x = cand[id].x + minB, y = cand[id].y + minB;
      const float angle = ic_angle(L.px.data(), L.w, x, y, t.umax);
      if (total < cap) {
        PortKp k;
        k.x = (float)x; k.y = (float)y;
        if (l != 0) { k.x *= t.scale[l]; k.y *= t.scale[l]; }
        k.size = size; k.angle = angle; k.response = (float)cand[id].score; k.octave = l; k.class_id = -1;
        kps[total] = k;
        descriptor(blurred.data(), L.w, x, y, angle, desc_out + (size_t)total * 32);
      }
      ++total;
    }
  }
  return total;
}

// Many frames on `nthreads` host threads, one extractor state per thread (bench CPU baseline "port").
int port_extract_batch(const port_params* p, const uint8_t* imgs, int nframes, int W, int H, void* kps_out,
                       uint8_t* desc_out, int cap, int* counts, int nthreads) {
  std::vector<std::thread> th;
  for (int tIdx = 0; tIdx < nthreads; ++tIdx)
    th.emplace_back([=]() {
      for (int f = tIdx; f < nframes; f += nthreads)
        counts[f] = port_extract(p, imgs + (size_t)f * W * H, W, H, W, (char*)kps_out + (size_t)f * cap * 28,
                                 desc_out + (size_t)f * cap * 32, cap, nullptr);
    });
  for (auto& x : th) x.join();
  return 0;
}

int port_hamming(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

// matcher.cpp:481-507 generalised to all pairs: train scanned in ascending index, strict <,
// first index wins; accepted iff d1 <= th && (float)d1 < ratio*(float)d2.
void port_knn2(const uint8_t* q, int Q, const uint8_t* t, long long M, int th, float ratio, int32_t* idx,
               int32_t* d1, int32_t* d2, uint8_t* ok, int nthreads) {
  auto work = [=](int q0, int q1) {
    for (int i = q0; i < q1; ++i) {
      int b1 = 256, b2 = 256, bi = -1;
      for (long long j = 0; j < M; ++j) {
        int d = hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
        if (d < b1) { b2 = b1; b1 = d; bi = (int)j; }
        else if (d < b2) b2 = d;
      }
      idx[i] = bi; d1[i] = b1; d2[i] = b2;
      ok[i] = (b1 <= th && (float)b1 < ratio * (float)b2) ? 1 : 0;
    }
  };
  if (nthreads <= 1) { work(0, Q); return; }
  std::vector<std::thread> thr;
  for (int k = 0; k < nthreads; ++k) thr.emplace_back(work, (int)((long long)Q * k / nthreads), (int)((long long)Q * (k + 1) / nthreads));
  for (auto& x : thr) x.join();
}

// Grid build: returns CSR (cell_start[64*48+1], ids[]) with cells indexed ix*48+iy.
int port_grid_build(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, int* cell_start, int* ids) {
  Grid g;
  g.build((const PortKp*)kps, n, xmin, xmax, ymin, ymax);
  int o = 0;
  for (int c = 0; c < Grid::COLS * Grid::ROWS; ++c) {
    cell_start[c] = o;
    for (int id : g.cell[c]) ids[o++] = id;
  }
  cell_start[Grid::COLS * Grid::ROWS] = o;
  return o;
}

// Frame post-processing after extraction (frame.cpp:22-32): undistortKeyPoints (:36-70), findDepth (:108-133),
// assignFeaturesToGrid (:72-97).  Camera constants as floats like camera.cpp:10-48; dist = k1 k2 p1 p2 [k3 [k4 k5 k6]].
struct port_camera { float fx, fy, cx, cy; float dist[8]; int ndist; float bf; float xmin, xmax, ymin, ymax; };

int port_frame_finish(const port_camera* cam, const void* kps_in, int n, const float* depth, int W, int H, size_t depth_step,
                      void* unkps_out, float* uright, float* depth_out, int* cell_start, int* ids) {
  const PortKp* kps = (const PortKp*)kps_in;
  PortKp* un = (PortKp*)unkps_out;
  double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < cam->ndist && i < 8; ++i) k[i] = (double)cam->dist[i];
  for (int i = 0; i < n; ++i) {
    un[i] = kps[i];
    if (cam->dist[0] != 0.0f)                                                        // frame.cpp:41-45
      cvp::undistort_point_k(kps[i].x, kps[i].y, cam->fx, cam->fy, cam->cx, cam->cy, k, &un[i].x, &un[i].y);
    uright[i] = -1.f; depth_out[i] = -1.f;                                           // :113-114
    if (depth) {
      const int c = (int)kps[i].x, r = (int)kps[i].y;                                // at<float>(v,u): float -> int truncation (:123)
      (void)W; (void)H;
      const float d = *(const float*)((const char*)depth + (size_t)r * depth_step + (size_t)c * sizeof(float));
      if (d > 0) { depth_out[i] = d; uright[i] = un[i].x - cam->bf / d; }            // :126-130
    }
  }
  return port_grid_build(un, n, cam->xmin, cam->xmax, cam->ymin, cam->ymax, cell_start, ids);
}

int port_features_in_area(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, float u, float v,
                          float r, int minL, int maxL, int* out, int cap) {
  Grid g;
  g.build((const PortKp*)kps, n, xmin, xmax, ymin, ymax);
  std::vector<int> o;
  g.area((const PortKp*)kps, u, v, r, minL, maxL, o);
  for (int i = 0; i < (int)o.size() && i < cap; ++i) out[i] = o[i];
  return (int)o.size();
}

// Inputs of the two projection searches, already projected (no Sophus arithmetic).
struct port_sbp_frame_in {
  // current frame
  const void* kps; const uint8_t* desc; const float* uright; int n;
  float xmin, xmax, ymin, ymax;
  const float* scale_factors; int nlevels;
  const uint8_t* occupied0;   // per current feature: already holds a map point with observe_cnt_>0
  // map points of the last frame, in index order; valid[i]==0 models `!mp || outlier`
  int m; const uint8_t* valid; const float* u; const float* v; const float* invz; const int32_t* octave;
  const float* angle; const uint8_t* mp_desc; const uint8_t* has_obs;
  float radius; float bf; int forward; int backward; int check_rot;
};

// matcher.cpp:18-148.  assign[idx] = index of the map point finally held by current feature idx, -1 = none
// written, -2 = written then cleared by the rotation check.  Returns match_cnt.
int port_sbp_frame(const port_sbp_frame_in* in, int32_t* assign) {
  const PortKp* kps = (const PortKp*)in->kps;
  Grid g;
  g.build(kps, in->n, in->xmin, in->xmax, in->ymin, in->ymax);
  std::vector<uint8_t> blocked(in->occupied0, in->occupied0 + in->n);
  for (int i = 0; i < in->n; ++i) assign[i] = -1;
  std::vector<std::vector<int>> hist(30);
  const float pdf = 30 / 360.0f;
  int cnt = 0;
  std::vector<int> win;
  for (int i = 0; i < in->m; ++i) {
    if (!in->valid[i]) continue;
    const float invz = in->invz[i];
    if (invz < 0.0f) continue;  // z < 0  (input contract: invz = 1/z)
    const float u = in->u[i], v = in->v[i];
    if (u < in->xmin || u > in->xmax) continue;
    if (v < in->ymin || v > in->ymax) continue;
    const int oct = in->octave[i];
    const float rs = in->radius * in->scale_factors[oct];
    if (in->forward) g.area(kps, u, v, rs, oct, in->nlevels, win);
    else if (in->backward) g.area(kps, u, v, rs, 0, oct, win);
    else g.area(kps, u, v, rs, oct - 1, oct + 1, win);
    if (win.empty()) continue;
    int bestD = 256, bestI = -1;
    for (int idx : win) {
      if (blocked[idx]) continue;
      if (in->uright[idx] > 0) {
        const float ur = u - in->bf * invz;
        if (fabsf(ur - in->uright[idx]) > rs) continue;
      }
      int d = hamming256(in->mp_desc + (size_t)i * 32, in->desc + (size_t)idx * 32);
      if (d < bestD) { bestD = d; bestI = idx; }
    }
    if (bestD <= 100) {
      assign[bestI] = i;
      blocked[bestI] = in->has_obs[i];
      ++cnt;
      if (in->check_rot) {
        float rot = in->angle[i] - kps[bestI].angle;
        if (rot < 0) rot += 360.0f;
        int bin = cvp::cv_round(rot * pdf);
        if (bin == 30) bin = 0;
        hist[bin].push_back(bestI);
      }
    }
  }
  if (in->check_rot) {
    int sizes[30], i1, i2, i3;
    for (int b = 0; b < 30; ++b) sizes[b] = (int)hist[b].size();
    three_max(sizes, 30, i1, i2, i3);
    for (int b = 0; b < 30; ++b)
      if (b != i1 && b != i2 && b != i3)
        for (int idx : hist[b]) { assign[idx] = -2; --cnt; }
  }
  return cnt;
}

struct port_sbp_local_in {
  const void* kps; const uint8_t* desc; const float* uright; int n;
  float xmin, xmax, ymin, ymax;
  const float* scale_factors; int nlevels;
  const uint8_t* occupied0;
  // local map points; valid[i]==0 models isBad() || !trackInLocalMap_
  int m; const uint8_t* valid; const float* u; const float* v; const float* ur; const int32_t* level;
  const float* view_cos; const uint8_t* mp_desc; const uint8_t* has_obs;
  float th_radius; float ratio;
};

// matcher.cpp:274-353
int port_sbp_local(const port_sbp_local_in* in, int32_t* assign) {
  const PortKp* kps = (const PortKp*)in->kps;
  Grid g;
  g.build(kps, in->n, in->xmin, in->xmax, in->ymin, in->ymax);
  std::vector<uint8_t> blocked(in->occupied0, in->occupied0 + in->n);
  for (int i = 0; i < in->n; ++i) assign[i] = -1;
  int cnt = 0;
  std::vector<int> win;
  for (int i = 0; i < in->m; ++i) {
    if (!in->valid[i]) continue;
    float radius = (in->view_cos[i] > 0.998) ? 2.5f : 4.0f;
    radius *= in->th_radius;
    const int lvl = in->level[i];
    const float rs = radius * in->scale_factors[lvl];
    g.area(kps, in->u[i], in->v[i], rs, lvl - 1, lvl, win);
    if (win.empty()) continue;
    int bestD = 256, bestL = -1, bestD2 = 256, bestL2 = -1, bestI = -1;
    for (int idx : win) {
      if (blocked[idx]) continue;
      if (in->uright[idx] > 0) {
        if (fabsf(in->ur[i] - in->uright[idx]) > rs) continue;
      }
      int d = hamming256(in->mp_desc + (size_t)i * 32, in->desc + (size_t)idx * 32);
      if (d < bestD) { bestD2 = bestD; bestD = d; bestL2 = bestL; bestL = kps[idx].octave; bestI = idx; }
      else if (d < bestD2) { bestL2 = kps[idx].octave; bestD2 = d; }
    }
    if (bestD <= 100) {
      if (bestL == bestL2 && (float)bestD > in->ratio * (float)bestD2) continue;
      assign[bestI] = i;
      blocked[bestI] = in->has_obs[i];
      ++cnt;
    }
  }
  return cnt;
}


struct port_bow_side {
  int n; const uint8_t* desc; const float* angle; const uint8_t* valid;
  int ngroups; const uint32_t* node_ids; const int32_t* group_start; const int32_t* feat_idx;
};

// matcher.cpp:449-559 (mode 0, KeyFrame -> Frame) and :561-677 (mode 1, KeyFrame -> KeyFrame).
// The FeatureVectors arrive as CSR sorted by node id; the merge walk (:465-535) visits the nodes both sides share.
int port_search_by_bow(const port_bow_side* a, const port_bow_side* b, int mode, float ratio, int th_low, int check_rot,
                       int32_t* match) {
  const int nOut = mode == 0 ? b->n : a->n;
  for (int i = 0; i < nOut; ++i) match[i] = -1;
  std::vector<uint8_t> taken(b->n, 0);
  std::vector<std::vector<int>> hist(30);
  const float pdf = 30 / 360.0f;
  int cnt = 0;
  int ga = 0, gb = 0;
  while (ga < a->ngroups && gb < b->ngroups) {
    if (a->node_ids[ga] == b->node_ids[gb]) {
      for (int ia = a->group_start[ga]; ia < a->group_start[ga + 1]; ++ia) {
        const int idx1 = a->feat_idx[ia];
        if (!a->valid[idx1]) continue;
        int best1 = 256, best2 = 256, bestIdx = -1;
        for (int ib = b->group_start[gb]; ib < b->group_start[gb + 1]; ++ib) {
          const int idx2 = b->feat_idx[ib];
          if (taken[idx2] || !b->valid[idx2]) continue;
          const int d = hamming256(a->desc + (size_t)idx1 * 32, b->desc + (size_t)idx2 * 32);
          if (d < best1) { best2 = best1; best1 = d; bestIdx = idx2; }
          else if (d < best2) best2 = d;
        }
        if (best1 <= th_low && (float)best1 < ratio * (float)best2) {
          taken[bestIdx] = 1;
          const int outIdx = mode == 0 ? bestIdx : idx1;
          match[outIdx] = mode == 0 ? idx1 : bestIdx;
          if (check_rot) {
            float rot = a->angle[idx1] - b->angle[bestIdx];
            if (rot < 0) rot += 360.0f;
            int bin = mode == 0 ? cvp::cv_round(rot * pdf) : (int)roundf(rot * pdf);
            if (bin == 30) bin = 0;
            hist[bin].push_back(outIdx);
          }
          ++cnt;
        }
      }
      ++ga; ++gb;
    } else if (a->node_ids[ga] < b->node_ids[gb]) {
      while (ga < a->ngroups && a->node_ids[ga] < b->node_ids[gb]) ++ga;      // lower_bound
    } else {
      while (gb < b->ngroups && b->node_ids[gb] < a->node_ids[ga]) ++gb;
    }
  }
  if (check_rot) {
    int sizes[30], i1, i2, i3;
    for (int k = 0; k < 30; ++k) sizes[k] = (int)hist[k].size();
    three_max(sizes, 30, i1, i2, i3);
    for (int k = 0; k < 30; ++k)
      if (k != i1 && k != i2 && k != i3)
        for (int idx : hist[k]) { match[idx] = -2; --cnt; }
  }
  return cnt;
}


// KeyFrame::getFeaturesInArea (keyframe.cpp:268-312): the Frame version without the level filter.
static void area_nolevel(const Grid& g, const PortKp* kps, float u, float v, float r, std::vector<int>& out) {
  g.area(kps, u, v, r, -1000000, 1000000, out);
}

// matcher.cpp:150-272.  Reuses port_sbp_frame_in: valid[i] = all host-side gates passed, octave[i] = predicted level,
// occupied0[i] = frame_curr->mappoints_[i] != nullptr; invz / has_obs / uright unused.
int port_sbp_reloc(const port_sbp_frame_in* in, float dist_threshold, int32_t* assign) {
  const PortKp* kps = (const PortKp*)in->kps;
  Grid g;
  g.build(kps, in->n, in->xmin, in->xmax, in->ymin, in->ymax);
  std::vector<uint8_t> taken(in->occupied0, in->occupied0 + in->n);
  for (int i = 0; i < in->n; ++i) assign[i] = -1;
  std::vector<std::vector<int>> hist(30);
  const float pdf = 30 / 360.0f;
  int cnt = 0;
  std::vector<int> win;
  for (int i = 0; i < in->m; ++i) {
    if (!in->valid[i]) continue;
    const int lp = in->octave[i];
    const float rs = in->radius * in->scale_factors[lp];
    g.area(kps, in->u[i], in->v[i], rs, lp - 1, lp + 1, win);
    if (win.empty()) continue;
    int bestD = 256, bestI = -1;
    for (int idx : win) {
      if (taken[idx]) continue;                                            // :218
      int d = hamming256(in->mp_desc + (size_t)i * 32, in->desc + (size_t)idx * 32);
      if (d < bestD) { bestD = d; bestI = idx; }
    }
    if (bestD <= dist_threshold) {                                         // :231
      assign[bestI] = i;
      taken[bestI] = 1;
      ++cnt;
      if (in->check_rot) {
        float rot = in->angle[i] - kps[bestI].angle;
        if (rot < 0) rot += 360.0f;
        int bin = cvp::cv_round(rot * pdf);
        if (bin == 30) bin = 0;
        hist[bin].push_back(bestI);
      }
    }
  }
  if (in->check_rot) {
    int sizes[30], i1, i2, i3;
    for (int b = 0; b < 30; ++b) sizes[b] = (int)hist[b].size();
    three_max(sizes, 30, i1, i2, i3);
    for (int b = 0; b < 30; ++b)
      if (b != i1 && b != i2 && b != i3)
        for (int idx : hist[b]) { assign[idx] = -2; --cnt; }
  }
  return cnt;
}

// matcher.cpp:356-447, including the `matchMapPoints[j]` indexing by window position (:422).
int port_sbp_sim3(const port_sbp_frame_in* in, int th, int32_t* assign) {
  const PortKp* kps = (const PortKp*)in->kps;
  Grid g;
  g.build(kps, in->n, in->xmin, in->xmax, in->ymin, in->ymax);
  std::vector<uint8_t> matched(in->occupied0, in->occupied0 + in->n);       // matchMapPoints[i] != nullptr
  for (int i = 0; i < in->n; ++i) assign[i] = -1;
  int cnt = 0;
  std::vector<int> win;
  for (int i = 0; i < in->m; ++i) {
    if (!in->valid[i]) continue;
    const int lp = in->octave[i];
    const float radius = th * in->scale_factors[lp];
    area_nolevel(g, kps, in->u[i], in->v[i], radius, win);
    if (win.empty()) continue;
    int bestD = 256, bestI = -1;
    for (int j = 0; j < (int)win.size(); ++j) {
      const int idx = win[j];
      if (matched[j]) continue;                                            // sic: [j], not [idx]
      const int level = kps[idx].octave;
      if (level < lp - 1 || level > lp) continue;
      int d = hamming256(in->mp_desc + (size_t)i * 32, in->desc + (size_t)idx * 32);
      if (d < bestD) { bestD = d; bestI = idx; }
    }
    if (bestD <= 50) {                                                     // TH_LOW
      assign[bestI] = i;
      matched[bestI] = 1;
      ++cnt;
    }
  }
  return cnt;
}


// mappoint.cpp:118-179 for many map points (CSR of their observed descriptors).
void port_medoid(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best) {
  for (int p = 0; p < npoints; ++p) {
    const int N = start[p + 1] - start[p];
    if (N <= 0) { best[p] = -1; continue; }
    const uint8_t* D = desc + (size_t)start[p] * 32;
    std::vector<std::vector<float>> distances(N, std::vector<float>(N, 0));
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        const int dij = hamming256(D + (size_t)i * 32, D + (size_t)j * 32);
        distances[i][j] = (float)dij;
        distances[j][i] = (float)dij;
      }
    int bestMid = 256, bestIdx = 0;
    for (int i = 0; i < N; ++i) {
      std::vector<int> dist(distances[i].begin(), distances[i].end());
      std::sort(dist.begin(), dist.end());
      const int mid = dist[int(0.5 * (N - 1))];
      if (mid < bestMid) { bestMid = mid; bestIdx = i; }
    }
    best[p] = bestIdx;
  }
}


// Common search core of searchBySim3 (matcher.cpp:745-775 / :806-836), fuseMapPoints (:1057-1100, chi2 != 0) and
// fuseByPose (:1190-1224).  in->invz carries ur when chi2 != 0; in->uright uses the `>= 0` convention of :1077.
void port_window_argmin(const port_sbp_frame_in* in, float th_radius, float dist_threshold, int chi2, int32_t* best) {
  const PortKp* kps = (const PortKp*)in->kps;
  Grid g;
  g.build(kps, in->n, in->xmin, in->xmax, in->ymin, in->ymax);
  std::vector<int> win;
  for (int i = 0; i < in->m; ++i) {
    best[i] = -1;
    if (!in->valid[i]) continue;
    const int lp = in->octave[i];
    const float u = in->u[i], v = in->v[i];
    const float radius = th_radius * in->scale_factors[lp];
    area_nolevel(g, kps, u, v, radius, win);
    if (win.empty()) continue;
    int bestD = 256, bestI = -1;
    for (int idx : win) {
      const PortKp& kp = kps[idx];
      if (kp.octave < lp - 1 || kp.octave > lp) continue;
      if (chi2) {
        const float ex = u - kp.x, ey = v - kp.y;
        const float invSigma = 1.0f / in->scale_factors[kp.octave];
        if (in->uright[idx] >= 0) {
          const float er = in->invz[i] - in->uright[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * invSigma * invSigma > 7.815f) continue;
        } else {
          const float e2 = ex * ex + ey * ey;
          if (e2 * invSigma * invSigma > 5.991f) continue;
        }
      }
      const int d = hamming256(in->mp_desc + (size_t)i * 32, in->desc + (size_t)idx * 32);
      if (d < bestD) { bestD = d; bestI = idx; }
    }
    if (bestD <= dist_threshold) best[i] = bestI;
  }
}

// matcher.cpp:679-865
int port_search_by_sim3(const port_sbp_frame_in* in12 /* frame = kf2, points = kf1's */, const port_sbp_frame_in* in21, float th,
                        int32_t* match12) {
  std::vector<int32_t> m1(std::max(in12->m, 1)), m2(std::max(in21->m, 1));
  port_window_argmin(in12, th, 100.f, 0, m1.data());
  port_window_argmin(in21, th, 100.f, 0, m2.data());
  int found = 0;
  for (int i = 0; i < in12->m; ++i) {
    match12[i] = -1;
    const int idx2 = m1[i];
    if (idx2 >= 0 && m2[idx2] == i) { match12[i] = idx2; ++found; }
  }
  return found;
}


struct port_tri_side { port_bow_side side; const void* kps; const float* uright; };

// matcher.cpp:1306-1324 (Eigen 3-term products evaluated left to right in double, then the two casts to float)
static bool epipolar_ok(const double* F, const PortKp& k1, const PortKp& k2, const float* scale2) {
  const double p1[3] = {(double)k1.x, (double)k1.y, 1.0};
  double l[3];
  for (int j = 0; j < 3; ++j) l[j] = p1[0] * F[j] + p1[1] * F[3 + j] + p1[2] * F[6 + j];
  const float numerator = (float)(l[0] * (double)k2.x + l[1] * (double)k2.y + l[2] * 1.0);
  const float denominator = (float)(l[0] * l[0] + l[1] * l[1]);
  if (denominator == 0) return false;
  const float d_square = numerator * numerator / denominator;
  const float sigma = scale2[k2.octave];
  return d_square < 3.84f * sigma * sigma;
}

// matcher.cpp:867-1010
int port_search_for_triangulation(const port_tri_side* a, const port_tri_side* b, const double* F12, float ex, float ey,
                                  const float* scale2, int th_low, int check_rot, int32_t* match) {
  const PortKp* k1s = (const PortKp*)a->kps;
  const PortKp* k2s = (const PortKp*)b->kps;
  for (int i = 0; i < a->side.n; ++i) match[i] = -1;
  std::vector<uint8_t> matched2(b->side.n, 0);
  std::vector<std::vector<int>> hist(30);
  const float pdf = 30 / 360.0f;
  int cnt = 0, ga = 0, gb = 0;
  while (ga < a->side.ngroups && gb < b->side.ngroups) {
    if (a->side.node_ids[ga] == b->side.node_ids[gb]) {
      for (int ia = a->side.group_start[ga]; ia < a->side.group_start[ga + 1]; ++ia) {
        const int idx1 = a->side.feat_idx[ia];
        if (!a->side.valid[idx1]) continue;                      // `if (mpk) continue;`
        const bool stereo1 = a->uright[idx1] >= 0;
        int bestDist = th_low, bestIdx2 = -1;
        for (int ib = b->side.group_start[gb]; ib < b->side.group_start[gb + 1]; ++ib) {
          const int idx2 = b->side.feat_idx[ib];
          if (matched2[idx2] || !b->side.valid[idx2]) continue;
          const bool stereo2 = b->uright[idx2] >= 0;
          const int dist = hamming256(a->side.desc + (size_t)idx1 * 32, b->side.desc + (size_t)idx2 * 32);
          if (dist > th_low || dist > bestDist) continue;
          const PortKp& kpt2 = k2s[idx2];
          if (!stereo1 && !stereo2) {
            const float distex = ex - kpt2.x, distey = ey - kpt2.y;
            if (distex * distex + distey * distey < 100 * scale2[kpt2.octave]) continue;
          }
          if (epipolar_ok(F12, k1s[idx1], kpt2, scale2)) { bestDist = dist; bestIdx2 = idx2; }
        }
        if (bestIdx2 >= 0) {
          match[idx1] = bestIdx2;
          matched2[bestIdx2] = 1;
          if (check_rot) {
            float rot = k1s[idx1].angle - k2s[bestIdx2].angle;
            if (rot < 0) rot += 360.0f;
            int bin = (int)roundf(rot * pdf);
            if (bin == 30) bin = 0;
            hist[bin].push_back(idx1);
          }
          ++cnt;
        }
      }
      ++ga; ++gb;
    } else if (a->side.node_ids[ga] < b->side.node_ids[gb]) {
      while (ga < a->side.ngroups && a->side.node_ids[ga] < b->side.node_ids[gb]) ++ga;
    } else {
      while (gb < b->side.ngroups && b->side.node_ids[gb] < a->side.node_ids[ga]) ++gb;
    }
  }
  if (check_rot) {
    int sizes[30], i1, i2, i3;
    for (int k = 0; k < 30; ++k) sizes[k] = (int)hist[k].size();
    three_max(sizes, 30, i1, i2, i3);
    for (int k = 0; k < 30; ++k)
      if (k != i1 && k != i2 && k != i3)
        for (int idx : hist[k]) { match[idx] = -2; --cnt; }
  }
  return cnt;
}

}  // extern "C"
