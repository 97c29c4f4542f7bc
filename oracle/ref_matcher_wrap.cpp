// oracle/ref_matcher_wrap.cpp -- TEST INFRASTRUCTURE: entry points around the reference's OWN myslam::Matcher.
//
// Compiled together with /root/reference/src/matcher.cpp (unmodified, in place; see Makefile) against the reference's own
// include/myslam/matcher.h and the stand-in object types of oracle/compat_myslam.  Nothing here restates the algorithm: each
// function forwards to the Matcher method of the same name, so that a program built WITHOUT the reference's headers (the GPU
// box has no /root/reference) can run the reference's matcher on its stand-in objects.
#include "myslam/matcher.h"

using namespace myslam;

int refm_searchByProjection_FF(float ratio, Frame* c, Frame* l, float radius, bool checkRot) {
  Matcher m(ratio); return m.searchByProjection(c, l, radius, checkRot);
}
int refm_searchByProjection_FK(float ratio, Frame* c, KeyFrame* kf, float radius, float distTh, const std::set<MapPoint*>& found, bool checkRot) {
  Matcher m(ratio); return m.searchByProjection(c, kf, radius, distTh, found, checkRot);
}
int refm_searchByProjection_local(float ratio, Frame* f, const std::vector<MapPoint*>& mps, float th) {
  Matcher m(ratio); return m.searchByProjection(f, mps, th);
}
int refm_searchByProjection_sim3(float ratio, KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& matchPts, int th) {
  Matcher m(ratio); return m.searchByProjection(kf, Scw, loopPts, matchPts, th);
}
int refm_searchByBoW_KF(float ratio, KeyFrame* kf, Frame* f, std::vector<MapPoint*>& out, bool checkRot) {
  Matcher m(ratio); return m.searchByBoW(kf, f, out, checkRot);
}
int refm_searchByBoW_KK(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& out, bool checkRot) {
  Matcher m(ratio); return m.searchByBoW(k1, k2, out, checkRot);
}
int refm_searchBySim3(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& m12, Sim3& S12, float th) {
  Matcher m(ratio); return m.searchBySim3(k1, k2, m12, S12, th);
}
int refm_searchForTriangulation(float ratio, KeyFrame* k1, KeyFrame* k2, std::vector<std::pair<int, int> >& idxs, Matrix3d& F12, bool checkRot) {
  Matcher m(ratio); return m.searchForTriangulation(k1, k2, idxs, F12, checkRot);
}
int refm_fuseMapPoints(float ratio, KeyFrame* kf, std::vector<MapPoint*>& mps, const float& th) {
  Matcher m(ratio); return m.fuseMapPoints(kf, mps, th);
}
int refm_fuseByPose(float ratio, KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& replacePts, float th) {
  Matcher m(ratio); return m.fuseByPose(kf, Scw, loopPts, replacePts, th);
}
int refm_computeDistance(const cv::Mat& a, const cv::Mat& b) { return Matcher::computeDistance(a, b); }

// Frame-to-frame top-2 + ratio test the way the reference's matching loops do it (matcher.cpp:481-507: a Mat row header per
// descriptor, Matcher::computeDistance per pair, strict '<' so the first index wins), threaded over queries for the CPU
// timing baseline of bench.py.  The distance function is the reference's own (matcher.cpp:1240-1256).
#include <thread>
extern "C" void refm_knn2(const uint8_t* q, int Q, const uint8_t* t, int M, int th, float ratio, int32_t* idx, int32_t* d1, int32_t* d2,
                          uint8_t* ok, int nthreads) {
  const cv::Mat qm(Q, 32, CV_8UC1, (void*)q, 32), tm(M, 32, CV_8UC1, (void*)t, 32);
  auto work = [&](int q0, int q1) {
    for (int i = q0; i < q1; ++i) {
      const cv::Mat qd = qm.row(i);
      int best1 = 256, best2 = 256, bi = -1;
      for (int j = 0; j < M; ++j) {
        const cv::Mat td = tm.row(j);
        const int dist = Matcher::computeDistance(qd, td);
        if (dist < best1) { best2 = best1; best1 = dist; bi = j; }
        else if (dist < best2) best2 = dist;
      }
      idx[i] = bi; d1[i] = best1; d2[i] = best2;
      ok[i] = (best1 <= th && static_cast<float>(best1) < ratio * static_cast<float>(best2)) ? 1 : 0;
    }
  };
  if (nthreads <= 1) { work(0, Q); return; }
  std::vector<std::thread> pool;
  for (int k = 0; k < nthreads; ++k) pool.emplace_back(work, (int)((long long)Q * k / nthreads), (int)((long long)Q * (k + 1) / nthreads));
  for (auto& x : pool) x.join();
}
