// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's OWN extractor class.
//
// Compiled together with /root/reference/src/ORBextractor.cpp (unmodified, in place; see Makefile) against
// oracle/compat.  Nothing here restates the algorithm; it only marshals POD buffers in and out of
// ORB_SLAM2::ORBextractor (ORBextractor.h:45-108) so Python (ctypes) and bench.py can call it.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "myslam/ORBextractor.h"

namespace {
// exposes the protected stages for stage-level parity (no source change in the reference)
struct Probe : public ORB_SLAM2::ORBextractor {
  using ORB_SLAM2::ORBextractor::ORBextractor;
  using ORB_SLAM2::ORBextractor::ComputePyramid;
  using ORB_SLAM2::ORBextractor::ComputeKeyPointsOctTree;
  using ORB_SLAM2::ORBextractor::DistributeOctTree;
  using ORB_SLAM2::ORBextractor::mnFeaturesPerLevel;
  using ORB_SLAM2::ORBextractor::umax;
};
}  // namespace

extern "C" {

void* orbref_create(int nfeatures, float scale, int nlevels, int iniTh, int minTh) {
  return new Probe(nfeatures, scale, nlevels, iniTh, minTh);
}
void orbref_destroy(void* h) { delete (Probe*)h; }

void orbref_tables(void* h, float* scale, float* inv_scale, int* nfeat, int* umax) {
  Probe* p = (Probe*)h;
  std::vector<float> s = p->GetScaleFactors(), is = p->GetInverseScaleFactors();
  for (size_t i = 0; i < s.size(); ++i) { scale[i] = s[i]; inv_scale[i] = is[i]; nfeat[i] = p->mnFeaturesPerLevel[i]; }
  for (int i = 0; i < 16; ++i) umax[i] = p->umax[i];
}

// ORBextractor::operator()  (ORBextractor.cpp:1051).  kps: 28-byte cv::KeyPoint records.
int orbref_extract(void* h, const uint8_t* img, int W, int H, size_t stride, void* kps, uint8_t* desc, int cap) {
  Probe* p = (Probe*)h;
  cv::Mat image(H, W, CV_8UC1, (void*)img, stride);
  std::vector<cv::KeyPoint> k;
  k.reserve(1 << 14);
  cv::Mat d;
  (*p)(image, cv::Mat(), k, d);
  int n = (int)k.size();
  for (int i = 0; i < n && i < cap; ++i) {
    std::memcpy((char*)kps + (size_t)i * 28, &k[i], 28);
    std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
  }
  return n;
}

// copy of mvImagePyramid[level] (ORBextractor.h:85) after the last extract; returns w*h
int orbref_pyramid_level(void* h, int level, uint8_t* out, int* w, int* hh) {
  Probe* p = (Probe*)h;
  const cv::Mat& m = p->mvImagePyramid[level];
  *w = m.cols; *hh = m.rows;
  if (out) for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
  return m.cols * m.rows;
}

// DistributeOctTree (ORBextractor.cpp:545) on a caller-supplied candidate list.
// cand: n triples (x, y, score) relative to the FAST region; out: triples of the kept keys in list order.
int orbref_octree(void* h, const int* cand, int n, int regionW, int regionH, int N, int* out, int cap) {
  Probe* p = (Probe*)h;
  std::vector<cv::KeyPoint> in(n);
  for (int i = 0; i < n; ++i) in[i] = cv::KeyPoint((float)cand[3 * i], (float)cand[3 * i + 1], 7.f, -1.f, (float)cand[3 * i + 2]);
  const int minX = 16, maxX = 16 + regionW, minY = 16, maxY = 16 + regionH, level = 0;
  std::vector<cv::KeyPoint> r = p->DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
  for (int i = 0; i < (int)r.size() && i < cap; ++i) {
    out[3 * i] = (int)r[i].pt.x; out[3 * i + 1] = (int)r[i].pt.y; out[3 * i + 2] = (int)r[i].response;
  }
  return (int)r.size();
}

// One extractor per host thread over disjoint frames (the extractor is stateful: ORBextractor.h:85).
// counts[f] receives the keypoint count of frame f; outputs beyond `cap` per frame are dropped.
int orbref_extract_batch(int nfeatures, float scale, int nlevels, int iniTh, int minTh, const uint8_t* imgs,
                         int nframes, int W, int H, void* kps, uint8_t* desc, int cap, int* counts, int nthreads) {
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=]() {
      Probe ex(nfeatures, scale, nlevels, iniTh, minTh);
      for (int f = t; f < nframes; f += nthreads)
        counts[f] = orbref_extract(&ex, imgs + (size_t)f * W * H, W, H, W,
                                   kps ? (char*)kps + (size_t)f * cap * 28 : nullptr,
                                   desc ? desc + (size_t)f * cap * 32 : nullptr, kps ? cap : 0);
    });
  for (auto& x : th) x.join();
  return 0;
}

}  // extern "C"
