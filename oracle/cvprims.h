// oracle/cvprims.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Plain C++ restatements of the third-party (OpenCV 4.13.0) primitives that the
// reference's ORB extractor calls.  OpenCV is NOT vendored in /root/reference
// (CMakeLists.txt:12 `find_package(OpenCV REQUIRED)`) and has no C++ headers in
// this image, so the arithmetic is restated here from its published algorithm
// and pinned bit-for-bit against the real `cv2` 4.13.0 wheel by
// tests/test_oracle_primitives.py and the fixtures in tests/golden/.
//
// Call sites in the reference these stand in for:
//   cv::resize(INTER_LINEAR)              ORBextractor.cpp:1129
//   cv::copyMakeBorder(REFLECT_101)       ORBextractor.cpp:1131,1137
//   cv::FAST(roi, kps, th, true)          ORBextractor.cpp:817,822
//   cv::GaussianBlur(7x7, 2, 2, R101)     ORBextractor.cpp:1094
//   cv::fastAtan2                         ORBextractor.cpp:106
//   cvRound / cvFloor / cvCeil            ORBextractor.cpp:83,118,123,447,459-465,1120
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
// may use anything under oracle/.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace cvp {

// cvRound: round-half-to-even in the default FP environment (SSE cvtss2si / lrint).
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

static inline int reflect101(int p, int len) {
  // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

// ---------------------------------------------------------------------------
// cv::resize, CV_8UC1, INTER_LINEAR (classic 11-bit fixed-point path).
// ---------------------------------------------------------------------------
struct ResizeTaps {
  std::vector<int> ofs;        // left/top source index
  std::vector<short> a0, a1;   // 11-bit weights (sum 2048)
};

static inline void resize_taps(int ssize, int dsize, ResizeTaps& t) {
  t.ofs.resize(dsize); t.a0.resize(dsize); t.a1.resize(dsize);
  const double inv_scale = (double)dsize / ssize;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cv_floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    t.ofs[d] = s;
    // saturate_cast<short>(float) rounds half-to-even
    t.a0[d] = (short)cv_round((1.f - f) * 2048.f);
    t.a1[d] = (short)cv_round(f * 2048.f);
  }
}

static inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                                    uint8_t* dst, int dw, int dh, size_t dstep) {
  ResizeTaps tx, ty;
  resize_taps(sw, dw, tx);
  resize_taps(sh, dh, ty);
  std::vector<int> row0(dw), row1(dw);
  int cached0 = -1, cached1 = -1;
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uint8_t* s = src + (size_t)sy * sstep;
    for (int d = 0; d < dw; ++d) {
      int sx = tx.ofs[d];
      int sx1 = std::min(sx + 1, sw - 1);
      out[d] = s[sx] * tx.a0[d] + s[sx1] * tx.a1[d];
    }
  };
  for (int dy = 0; dy < dh; ++dy) {
    int sy0 = ty.ofs[dy];
    int sy1 = std::min(sy0 + 1, sh - 1);
    if (cached1 == sy0) { row0.swap(row1); std::swap(cached0, cached1); }
    if (cached0 != sy0) { hrow(sy0, row0); cached0 = sy0; }
    if (cached1 != sy1) {
      if (sy1 == sy0) { row1 = row0; } else { hrow(sy1, row1); }
      cached1 = sy1;
    }
    const int b0 = ty.a0[dy], b1 = ty.a1[dy];
    uint8_t* d = dst + (size_t)dy * dstep;
    for (int x = 0; x < dw; ++x) {
      int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
      d[x] = (uint8_t)v;  // always within [0,255]
    }
  }
}

// ---------------------------------------------------------------------------
// cv::GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101), CV_8UC1.
// OpenCV >= 3.4.1 fixed-point path: 8.8 kernel {18,34,48,56,48,34,18}.
// dst may alias src.
// ---------------------------------------------------------------------------
static inline void gaussian_blur7_s2_u8(const uint8_t* src, int w, int h, size_t sstep,
                                        uint8_t* dst, size_t dstep) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> hbuf((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = src + (size_t)y * sstep;
    uint16_t* o = hbuf.data() + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      int acc = 0;
      if (x >= 3 && x + 3 < w) {
        for (int i = 0; i < 7; ++i) acc += k[i] * s[x + i - 3];
      } else {
        for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, w)];
      }
      o[x] = (uint16_t)acc;
    }
  }
  for (int y = 0; y < h; ++y) {
    const uint16_t* r[7];
    for (int j = 0; j < 7; ++j) r[j] = hbuf.data() + (size_t)reflect101(y + j - 3, h) * w;
    uint8_t* d = dst + (size_t)y * dstep;
    for (int x = 0; x < w; ++x) {
      uint32_t acc = 32768u;
      for (int j = 0; j < 7; ++j) acc += (uint32_t)k[j] * r[j][x];
      d[x] = (uint8_t)(acc >> 16);
    }
  }
}

// ---------------------------------------------------------------------------
// cv::fastAtan2(float y, float x) scalar path -> degrees in [0,360).
// Non-FMA float32 evaluation (compile with -ffp-contract=off).
// ---------------------------------------------------------------------------
static inline float fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression), TYPE_9_16.
// Output: row-major list of (x, y, score) relative to the image/ROI origin.
// ---------------------------------------------------------------------------
struct FastKp { int x, y, score; };

// Corner strength of one pixel: max over the 16 arcs of 9 contiguous ring pixels of
// max(min(d), min(-d)), d_k = I(p) - I(ring_k).  A pixel is a corner at threshold t iff
// strength > t and its stored score is strength - 1.
static inline int fast_strength(const uint8_t* p, const int* ring) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; ++k) d[k] = v - p[ring[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = 0;
  for (int s = 0; s < 16; ++s) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; ++k) { mn = std::min(mn, d[s + k]); mx = std::max(mx, d[s + k]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

static inline void fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms,
                            std::vector<FastKp>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int ring[16];
  for (int k = 0; k < 16; ++k) ring[k] = dy[k] * (int)step + dx[k];
  threshold = std::min(std::max(threshold, 0), 255);
  // threshold table: 1 = ring darker than centre by more than t, 2 = brighter.
  uint8_t tab[512];
  for (int i = -255; i <= 255; ++i) tab[i + 255] = (uint8_t)(i < -threshold ? 1 : i > threshold ? 2 : 0);

  // three rolling rows of scores (zero = not a corner) and corner x positions
  std::vector<uint8_t> sbuf((size_t)3 * w, 0);
  std::vector<int> cbuf((size_t)3 * (w + 1), 0);
  for (int i = 3; i < h - 2; ++i) {
    uint8_t* curr = &sbuf[(size_t)((i - 3) % 3) * w];
    int* cpos = &cbuf[(size_t)((i - 3) % 3) * (w + 1)];
    std::memset(curr, 0, w);
    int ncorners = 0;
    if (i < h - 3) {
      const uint8_t* row = img + (size_t)i * step;
      for (int j = 3; j < w - 3; ++j) {
        const uint8_t* p = row + j;
        const uint8_t* t = &tab[255 - p[0]];
        int m = t[p[ring[0]]] | t[p[ring[8]]];
        if (!m) continue;
        m &= t[p[ring[2]]] | t[p[ring[10]]];
        m &= t[p[ring[4]]] | t[p[ring[12]]];
        m &= t[p[ring[6]]] | t[p[ring[14]]];
        if (!m) continue;
        m &= t[p[ring[1]]] | t[p[ring[9]]];
        m &= t[p[ring[3]]] | t[p[ring[11]]];
        m &= t[p[ring[5]]] | t[p[ring[13]]];
        m &= t[p[ring[7]]] | t[p[ring[15]]];
        if (!m) continue;
        int s = fast_strength(p, ring);
        if (s > threshold) {
          cpos[ncorners++] = j;
          curr[j] = (uint8_t)(s - 1);
        }
      }
    }
    cpos[w] = ncorners;  // count stored past the positions
    if (i == 3) continue;
    // emit row i-1 (its neighbours i-2 and i are now known)
    const uint8_t* prev = &sbuf[(size_t)((i - 4 + 3) % 3) * w];
    const uint8_t* pprev = &sbuf[(size_t)((i - 5 + 3) % 3) * w];
    const int* ppos = &cbuf[(size_t)((i - 4 + 3) % 3) * (w + 1)];
    int n = ppos[w];
    for (int k = 0; k < n; ++k) {
      int j = ppos[k];
      int sc = prev[j];
      if (!nms || (sc > prev[j + 1] && sc > prev[j - 1] && sc > pprev[j - 1] && sc > pprev[j] &&
                   sc > pprev[j + 1] && sc > curr[j - 1] && sc > curr[j] && sc > curr[j + 1])) {
        out.push_back(FastKp{j, i - 1, sc});
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::undistortPoints(src, dst, K, D, noArray(), P = K), CV_32FC2 (call site frame.cpp:58).  OpenCV >= 3.x runs the
// fixed-point inversion of the distortion model for exactly 5 iterations (TermCriteria(MAX_ITER, 5, 0.01): the
// epsilon is unused because the type has no EPS bit), everything in double without FMA contraction, then applies
// P * R (R = I) and narrows to float.  k = {k1,k2,p1,p2,k3,k4,k5,k6} widened from float; the thin-prism and tilt
// terms the reference never passes (camera.cpp:27-38 builds 4 or 5 coefficients) are zero and drop out exactly.
// Pinned bit-exact against cv2 4.13.0 (tests/golden/cv2_undistort.npz).
static inline void undistort_point_k(float u_in, float v_in, double fx, double fy, double cx, double cy, const double* k,
                                     float* u_out, float* v_out) {
  const double ifx = 1. / fx, ify = 1. / fy;
  const double u = u_in, v = v_in;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) {          // OpenCV regression test 14583: give up and return the linear back-projection
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  *u_out = (float)(fx * x + cx);
  *v_out = (float)(fy * y + cy);
}

}  // namespace cvp
