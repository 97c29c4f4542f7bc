// oracle/compat_myslam/myslam_stub.hpp -- TEST INFRASTRUCTURE.  Stand-ins for the reference's Frame / KeyFrame / MapPoint / Camera /
// SE3 types with exactly the members Matcher's loops touch (include/myslam/frame.h:16-71, keyframe.h, mappoint.h:20-90,
// camera.h:13-45), so that include/orb_b200_matcher.hpp can be compiled and run without Sophus / Eigen / DBoW3, plus a
// loop-for-loop CPU statement of the reference's search functions over these objects (`RefMatcher`), which is what the
// adapter's results are compared with.  The same types stand in for the real ones when the reference's OWN src/matcher.cpp is
// compiled in place (oracle/Makefile, _ref/libmatcherref.so; myslam/keyframe.h and myslam/mappoint.h in this directory
// shadow the reference's headers), which is the stronger checker of the two.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <vector>

#include "opencv/cv.h"

namespace myslam {

struct Vector3d {
  double d[3];
  Vector3d() { d[0] = d[1] = d[2] = 0; }
  Vector3d(double x, double y, double z) { d[0] = x; d[1] = y; d[2] = z; }
  double operator[](int i) const { return d[i]; }
  Vector3d operator-(const Vector3d& o) const { return Vector3d(d[0] - o.d[0], d[1] - o.d[1], d[2] - o.d[2]); }
  Vector3d operator+(const Vector3d& o) const { return Vector3d(d[0] + o.d[0], d[1] + o.d[1], d[2] + o.d[2]); }
  Vector3d operator/(double s) const { return Vector3d(d[0] / s, d[1] / s, d[2] / s); }
  double dot(const Vector3d& o) const { return d[0] * o.d[0] + d[1] * o.d[1] + d[2] * o.d[2]; }
  double norm() const { return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); }
  inline struct RowVector3d transpose() const;
};
struct Matrix3d;
struct RowVector3d {                                  // p.transpose(): only what (p1.transpose() * F12).transpose() needs
  double d[3];
  inline RowVector3d operator*(const Matrix3d& M) const;
  Vector3d transpose() const { return Vector3d(d[0], d[1], d[2]); }
};
inline RowVector3d transposeOf(const Vector3d& v) { RowVector3d r; r.d[0] = v[0]; r.d[1] = v[1]; r.d[2] = v[2]; return r; }
struct Matrix3d {                                     // the slice of Eigen::Matrix3d the Sim3 searches use
  double m[9];
  double operator()(int r, int c) const { return m[r * 3 + c]; }
  Matrix3d operator/(double s) const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] / s; return r; }
  Matrix3d operator-() const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = -m[i]; return r; }
  Matrix3d transpose() const { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = m[j * 3 + i]; return r; }
  Vector3d operator*(const Vector3d& p) const {
    return Vector3d(m[0] * p[0] + m[1] * p[1] + m[2] * p[2], m[3] * p[0] + m[4] * p[1] + m[5] * p[2], m[6] * p[0] + m[7] * p[1] + m[8] * p[2]);
  }
};
struct Vector2d {
  double d[2];
  Vector2d(double x, double y) { d[0] = x; d[1] = y; }
  double operator[](int i) const { return d[i]; }
};

inline RowVector3d RowVector3d::operator*(const Matrix3d& M) const {      // Eigen evaluates the 3-term sums left to right
  RowVector3d r;
  for (int j = 0; j < 3; ++j) r.d[j] = d[0] * M(0, j) + d[1] * M(1, j) + d[2] * M(2, j);
  return r;
}
inline RowVector3d Vector3d::transpose() const { return transposeOf(*this); }

// rigid transform x -> R x + t (what Sophus::SE3 provides to the matcher: operator*, inverse, translation)
struct SE3 {
  double R[9], t[3];
  SE3() { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0); t[0] = t[1] = t[2] = 0; }
  SE3(const Matrix3d& Rm, const Vector3d& tv) { for (int i = 0; i < 9; ++i) R[i] = Rm.m[i]; for (int i = 0; i < 3; ++i) t[i] = tv[i]; }
  Matrix3d rotation_matrix() const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = R[i]; return r; }
  static SE3 rotY(double a, double tx, double ty, double tz) {
    SE3 s; s.R[0] = std::cos(a); s.R[2] = std::sin(a); s.R[6] = -std::sin(a); s.R[8] = std::cos(a);
    s.t[0] = tx; s.t[1] = ty; s.t[2] = tz; return s;
  }
  Vector3d operator*(const Vector3d& p) const {
    return Vector3d(R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0], R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1],
                    R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2]);
  }
  SE3 operator*(const SE3& o) const {
    SE3 s;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) s.R[i * 3 + j] = R[i * 3] * o.R[j] + R[i * 3 + 1] * o.R[3 + j] + R[i * 3 + 2] * o.R[6 + j];
      s.t[i] = R[i * 3] * o.t[0] + R[i * 3 + 1] * o.t[1] + R[i * 3 + 2] * o.t[2] + t[i];
    }
    return s;
  }
  SE3 inverse() const {
    SE3 s;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.R[i * 3 + j] = R[j * 3 + i];
    for (int i = 0; i < 3; ++i) s.t[i] = -(s.R[i * 3] * t[0] + s.R[i * 3 + 1] * t[1] + s.R[i * 3 + 2] * t[2]);
    return s;
  }
  Vector3d translation() const { return Vector3d(t[0], t[1], t[2]); }
};

// similarity x -> s R x + t (Sophus::Sim3: rotation_matrix() returns s*R, like the non-templated Sophus the reference uses)
struct Sim3 {
  SE3 rt; double s;
  Sim3(const SE3& T, double scale) : rt(T), s(scale) {}
  double scale() const { return s; }
  Matrix3d rotation_matrix() const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = s * rt.R[i]; return r; }
  Vector3d translation() const { return rt.translation(); }
  Vector3d operator*(const Vector3d& p) const {
    const Vector3d r = rt * p;                          // R p + t ...
    return Vector3d(s * (r[0] - rt.t[0]) + rt.t[0], s * (r[1] - rt.t[1]) + rt.t[1], s * (r[2] - rt.t[2]) + rt.t[2]);   // ... as s R p + t
  }
  Sim3 inverse() const {                                // x = (1/s) R^T (y - t)
    SE3 inv = rt.inverse();
    for (int k = 0; k < 3; ++k) inv.t[k] /= s;
    return Sim3(inv, 1.0 / s);
  }
};

struct Camera {
  float fx_, fy_, cx_, cy_, bf_, b_;
  cv::Mat K_, distCoef_;                                  // CV_32F 3x3 and 4x1 / 5x1, stored in byte matrices of 4x the width
  void setIntrinsics(const float* dist, int ndist) {      // camera.cpp:19-38
    K_ = cv::Mat(3, 3 * 4, CV_8UC1); distCoef_ = cv::Mat(ndist, 4, CV_8UC1);
    const float k[9] = {fx_, 0, cx_, 0, fy_, cy_, 0, 0, 1};
    for (int i = 0; i < 9; ++i) K_.at<float>(i / 3, i % 3) = k[i];
    for (int i = 0; i < ndist; ++i) distCoef_.at<float>(i, 0) = dist[i];
  }
  Vector2d camera2pixel(const Vector3d& p) { return Vector2d(fx_ * p[0] / p[2] + cx_, fy_ * p[1] / p[2] + cy_); }   // camera.cpp:72-75
};

struct Frame;
struct MapPoint {
  Vector3d pos_;
  float minDistance_ = 0, maxDistance_ = 0;
  Vector3d normalVector_;
  Vector3d getNormalVector() { return normalVector_; }
  std::map<struct KeyFrame*, size_t> observedKFs_;   // mappoint.h:49
  struct NoMutex { void lock() {} void unlock() {} } mutexFeature_;   // single-threaded stand-in for std::mutex
  int getIndexInKeyFrame(struct KeyFrame* kf) { return observedKFs_.count(kf) ? (int)observedKFs_[kf] : -1; }
  bool beObserved(struct KeyFrame* kf) { return observedKFs_.count(kf) != 0; }
  void addObservation(struct KeyFrame* kf, int idx) { if (!observedKFs_.count(kf)) { observedKFs_[kf] = idx; ++observe_cnt_; } }
  inline void replaceMapPoint(MapPoint* mp);            // stand-in for mappoint.cpp:214-262: hand the observations over, go bad
  inline int predictScale(const float& currDist, struct KeyFrame* kf);  // mappoint.cpp:198-212
  float getMinDistanceThreshold() { return 0.8f * minDistance_; }      // mappoint.cpp:391-401
  float getMaxDistanceThreshold() { return 1.2f * maxDistance_; }
  inline int predictScale(const float& currDist, Frame* frame);        // mappoint.cpp:182-196
  cv::Mat descriptor_;
  int observe_cnt_ = 0;
  bool badFlag_ = false;
  bool trackInLocalMap_ = false;
  int trackScaleLevel_ = 0;
  float trackProj_u_ = 0, trackProj_uR_ = 0, trackProj_v_ = 0, viewCos_ = 0;
  Vector3d getPose() { return pos_; }
  cv::Mat getDescriptor() { return descriptor_.clone(); }
  int getObsCnt() { return observe_cnt_; }
  bool isBad() { return badFlag_; }
};

typedef std::map<unsigned, std::vector<unsigned> > FeatureVector;   // DBoW3::FeatureVector

struct Frame {
  Camera* camera_ = nullptr;
  SE3 Tcw_;
  std::vector<cv::KeyPoint> keypoints_;
  std::vector<cv::KeyPoint> unKeypoints_;
  std::vector<float> depth_;
  std::vector<float> uRight_;
  cv::Mat descriptors_;
  std::vector<MapPoint*> mappoints_;
  std::vector<float> scaleFactors_;
  size_t N_ = 0;
  float xMin_ = 0, xMax_ = 0, yMin_ = 0, yMax_ = 0, gridPerPixelWidth_ = 0, gridPerPixelHeight_ = 0;
  std::vector<int> gridKeypoints_[64][48];
  std::vector<bool> outliers_;
  FeatureVector featVec_;

  // frame.cpp:36-70 with cv::undistortPoints(mat, mat, K, distCoef, Mat(), K) stated by oracle/cvprims.h (pinned to cv2 4.13)
  void undistortKeyPoints() {
    const cv::Mat& D = camera_->distCoef_;
    unKeypoints_ = keypoints_;
    if (D.at<float>(0, 0) == 0.0f) return;
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < D.rows && i < 8; ++i) k[i] = (double)D.at<float>(i, 0);
    const cv::Mat& K = camera_->K_;
    for (int i = 0; i < (int)N_; ++i)
      cvp::undistort_point_k(keypoints_[i].pt.x, keypoints_[i].pt.y, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2),
                             K.at<float>(1, 2), k, &unKeypoints_[i].pt.x, &unKeypoints_[i].pt.y);
  }
  void findDepth(cv::Mat& depthImg) {                                  // frame.cpp:108-133
    if (keypoints_.empty()) return;
    uRight_ = std::vector<float>(N_, -1);
    depth_ = std::vector<float>(N_, -1);
    for (int i = 0; i < (int)N_; ++i) {
      const float u = keypoints_[i].pt.x, v = keypoints_[i].pt.y;
      const float d = depthImg.at<float>(v, u);
      if (d > 0) { depth_[i] = d; uRight_[i] = unKeypoints_[i].pt.x - camera_->bf_ / d; }
    }
  }
  void assignFeaturesToGrid() {                                       // frame.cpp:72-97
    for (int ix = 0; ix < 64; ++ix) for (int iy = 0; iy < 48; ++iy) gridKeypoints_[ix][iy].clear();
    for (int i = 0; i < (int)N_; ++i) {
      const int gx = (int)round((unKeypoints_[i].pt.x - xMin_) * gridPerPixelWidth_);
      const int gy = (int)round((unKeypoints_[i].pt.y - yMin_) * gridPerPixelHeight_);
      if (gx >= 0 && gx < 64 && gy >= 0 && gy < 48) gridKeypoints_[gx][gy].push_back(i);
    }
  }
  std::vector<int> getFeaturesInArea(const float& u, const float& v, const float& radius, int min_level, int max_level) {   // frame.cpp:199-247
    std::vector<int> out;
    const int x0 = std::max(0, (int)floor((u - xMin_ - radius) * gridPerPixelWidth_));
    if (x0 >= 64) return out;
    const int x1 = std::min(63, (int)floor((u - xMin_ + radius) * gridPerPixelWidth_));
    if (x1 < 0) return out;
    const int y0 = std::max(0, (int)floor((v - yMin_ - radius) * gridPerPixelHeight_));
    if (y0 >= 48) return out;
    const int y1 = std::min(47, (int)floor((v - yMin_ + radius) * gridPerPixelHeight_));
    if (y1 < 0) return out;
    for (int ix = x0; ix <= x1; ++ix)
      for (int iy = y0; iy <= y1; ++iy)
        for (size_t k = 0; k < gridKeypoints_[ix][iy].size(); ++k) {
          const int id = gridKeypoints_[ix][iy][k];
          const cv::KeyPoint& kp = unKeypoints_[id];
          if (kp.octave < min_level || kp.octave > max_level) continue;
          if (fabs(kp.pt.x - u) < radius && fabs(kp.pt.y - v) < radius) out.push_back(id);
        }
    return out;
  }
};

inline int MapPoint::predictScale(const float& currDist, Frame* frame) {
  const float ratio = maxDistance_ / currDist;
  int scale = (int)ceil(log(ratio) / log(frame->scaleFactors_[1]));
  if (scale < 0) scale = 0;
  else if (scale >= (int)frame->scaleFactors_.size()) scale = (int)frame->scaleFactors_.size() - 1;
  return scale;
}

struct KeyFrame {
  bool bad_kf_ = false;
  bool isBad() { return bad_kf_; }
  Camera* camera_ = nullptr;
  SE3 Tcw_;
  SE3 getPose() { return Tcw_; }
  Vector3d getCamCenter() { return Tcw_.inverse().translation(); }
  void addMapPoint(MapPoint* mp, int idx) { mappoints_[idx] = mp; }
  std::vector<float> uRight_;
  float xMin_ = 0, xMax_ = 0, yMin_ = 0, yMax_ = 0, gridPerPixelWidth_ = 0, gridPerPixelHeight_ = 0;
  std::vector<std::vector<std::vector<int> > > gridKeypoints_;
  bool isInImg(const float& u, const float& v) { return u >= xMin_ && v >= yMin_ && u < xMax_ && v < yMax_; }   // keyframe.cpp:64-67
  void assignFeaturesToGrid() {                                        // same rule as Frame (keyframe.cpp copies the frame's grid)
    gridKeypoints_.assign(64, std::vector<std::vector<int> >(48));
    for (int i = 0; i < (int)N_; ++i) {
      const int gx = (int)round((unKeypoints_[i].pt.x - xMin_) * gridPerPixelWidth_);
      const int gy = (int)round((unKeypoints_[i].pt.y - yMin_) * gridPerPixelHeight_);
      if (gx >= 0 && gx < 64 && gy >= 0 && gy < 48) gridKeypoints_[gx][gy].push_back(i);
    }
  }
  std::vector<int> getFeaturesInArea(const float& u, const float& v, const float& radius) {   // keyframe.cpp:268-312
    std::vector<int> out;
    const int x0 = std::max(0, (int)floor((u - xMin_ - radius) * gridPerPixelWidth_));
    if (x0 >= 64) return out;
    const int x1 = std::min(63, (int)floor((u - xMin_ + radius) * gridPerPixelWidth_));
    if (x1 < 0) return out;
    const int y0 = std::max(0, (int)floor((v - yMin_ - radius) * gridPerPixelHeight_));
    if (y0 >= 48) return out;
    const int y1 = std::min(47, (int)floor((v - yMin_ + radius) * gridPerPixelHeight_));
    if (y1 < 0) return out;
    for (int ix = x0; ix <= x1; ++ix)
      for (int iy = y0; iy <= y1; ++iy)
        for (size_t k = 0; k < gridKeypoints_[ix][iy].size(); ++k) {
          const int id = gridKeypoints_[ix][iy][k];
          if (fabs(unKeypoints_[id].pt.x - u) < radius && fabs(unKeypoints_[id].pt.y - v) < radius) out.push_back(id);
        }
    return out;
  }
  std::vector<float> scaleFactors_;
  std::vector<cv::KeyPoint> unKeypoints_;
  cv::Mat descriptors_;
  size_t N_ = 0;
  FeatureVector featVec_;
  std::vector<MapPoint*> mappoints_;
  std::vector<MapPoint*> getMapPoints() { return mappoints_; }
};

inline int MapPoint::predictScale(const float& currDist, KeyFrame* kf) {
  const float ratio = maxDistance_ / currDist;
  int scale = (int)ceil(log(ratio) / log(kf->scaleFactors_[1]));
  if (scale < 0) scale = 0;
  else if (scale >= (int)kf->scaleFactors_.size()) scale = (int)kf->scaleFactors_.size() - 1;
  return scale;
}

inline void MapPoint::replaceMapPoint(MapPoint* mp) {
  if (mp == this) return;
  std::map<KeyFrame*, size_t> obs = observedKFs_;
  observedKFs_.clear();
  badFlag_ = true;
  for (std::map<KeyFrame*, size_t>::iterator it = obs.begin(); it != obs.end(); ++it) {
    if (!mp->beObserved(it->first)) { it->first->mappoints_[it->second] = mp; mp->addObservation(it->first, it->second); }
    else it->first->mappoints_[it->second] = nullptr;
  }
}

// MapPoint::computeDescriptor (mappoint.cpp:118-179) over the stand-in objects
inline void refComputeDescriptor(MapPoint* mp) {
  if (mp->badFlag_ || mp->observedKFs_.empty()) return;
  std::vector<cv::Mat> desp;
  for (std::map<KeyFrame*, size_t>::iterator it = mp->observedKFs_.begin(); it != mp->observedKFs_.end(); ++it)
    if (!it->first->isBad()) desp.push_back(it->first->descriptors_.row((int)it->second));
  if (desp.empty()) return;
  const size_t N = desp.size();
  std::vector<std::vector<float> > dm(N, std::vector<float>(N, 0));
  for (size_t i = 0; i < N; ++i)
    for (size_t j = i + 1; j < N; ++j) {
      int d = 0;
      for (int b = 0; b < 32; ++b) d += __builtin_popcount((unsigned)(desp[i].data[b] ^ desp[j].data[b]));
      dm[i][j] = d; dm[j][i] = d;
    }
  int bestMid = 256, bestIdx = 0;
  for (size_t i = 0; i < N; ++i) {
    std::vector<int> row(dm[i].begin(), dm[i].end());
    std::sort(row.begin(), row.end());
    const int mid = row[int(0.5 * (N - 1))];
    if (mid < bestMid) { bestMid = mid; bestIdx = (int)i; }
  }
  mp->descriptor_ = desp[bestIdx].clone();
}

// ---- CPU statement of the reference's loops over the objects above (checker only) ------------------------------------
struct RefMatcher {
  float ratio_;
  explicit RefMatcher(float r) : ratio_(r) {}

  static int computeDistance(const cv::Mat& a, const cv::Mat& b) {   // matcher.cpp:1240-1256 (value: 256-bit Hamming distance)
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a.data[i] ^ b.data[i]));
    return d;
  }
  static void threeMax(std::vector<int>* h, int L, int& i1, int& i2, int& i3) {   // matcher.cpp:1258-1304
    int m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < L; ++i) {
      const int s = (int)h[i].size();
      if (s > m1) { m3 = m2; i3 = i2; m2 = m1; i2 = i1; m1 = s; i1 = i; }
      else if (s > m2) { m3 = m2; i3 = i2; m2 = s; i2 = i; }
      else if (s > m3) { m3 = s; i3 = i; }
    }
    if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
    else if (m3 < 0.1f * (float)m1) i3 = -1;
  }
  static int histBin(float rot, bool half_even) {
    if (rot < 0) rot += 360.0f;
    const float x = rot * (30 / 360.0f);
    int bin = half_even ? cvRound(x) : (int)round(x);     // :118 uses cvRound, :637 uses round
    return bin == 30 ? 0 : bin;
  }

  int searchByProjection(Frame* cur, Frame* last, const float radius, bool checkRot) {   // matcher.cpp:18-148
    int cnt = 0;
    std::vector<int> hist[30];
    Camera* cam = cur->camera_;
    const int xMax = cur->xMax_, xMin = cur->xMin_, yMax = cur->yMax_, yMin = cur->yMin_;
    const int levels = (int)cur->scaleFactors_.size();
    SE3 Tcw = cur->Tcw_;
    SE3 Tlc = last->Tcw_ * Tcw.inverse();
    const bool forward = (float)Tlc.translation()[2] > cam->b_;
    const bool backward = -(float)Tlc.translation()[2] > cam->b_;
    for (int i = 0; i < (int)last->mappoints_.size(); ++i) {
      MapPoint* mp = last->mappoints_[i];
      if (!mp || last->outliers_[i]) continue;
      Vector3d pc = Tcw * mp->getPose();
      const float z = (float)pc[2];
      if (z < 0.0f) continue;
      const float invz = 1.0f / z;
      Vector2d px = cam->camera2pixel(pc);
      const float u = px[0], v = px[1];
      if (u < xMin || u > xMax || v < yMin || v > yMax) continue;
      const int oct = last->unKeypoints_[i].octave;
      const float rs = radius * cur->scaleFactors_[oct];
      std::vector<int> ids = forward ? cur->getFeaturesInArea(u, v, rs, oct, levels)
                           : backward ? cur->getFeaturesInArea(u, v, rs, 0, oct)
                                      : cur->getFeaturesInArea(u, v, rs, oct - 1, oct + 1);
      if (ids.empty()) continue;
      int best = 256, bestIdx = -1;
      const cv::Mat dl = mp->getDescriptor();
      for (size_t j = 0; j < ids.size(); ++j) {
        const int idx = ids[j];
        if (cur->mappoints_[idx] && cur->mappoints_[idx]->observe_cnt_ > 0) continue;
        if (cur->uRight_[idx] > 0) {
          const float ur = u - cam->bf_ * invz;
          if (fabs(ur - cur->uRight_[idx]) > rs) continue;
        }
        const int d = computeDistance(dl, cur->descriptors_.row(idx));
        if (d < best) { best = d; bestIdx = idx; }
      }
      if (best <= 100) {
        cur->mappoints_[bestIdx] = mp;
        ++cnt;
        if (checkRot) hist[histBin(last->unKeypoints_[i].angle - cur->unKeypoints_[bestIdx].angle, true)].push_back(bestIdx);
      }
    }
    if (checkRot) {
      int i1 = -1, i2 = -1, i3 = -1;
      threeMax(hist, 30, i1, i2, i3);
      for (int i = 0; i < 30; ++i)
        if (i != i1 && i != i2 && i != i3)
          for (size_t j = 0; j < hist[i].size(); ++j) { cur->mappoints_[hist[i][j]] = nullptr; --cnt; }
    }
    return cnt;
  }

  int searchByProjection(Frame* cur, KeyFrame* kf, const float radius, const float distThreshold,
                         const std::set<MapPoint*>& found, bool checkRot) {                     // matcher.cpp:150-272
    int cnt = 0;
    std::vector<int> hist[30];
    const int xMax = cur->xMax_, xMin = cur->xMin_, yMax = cur->yMax_, yMin = cur->yMin_;
    const SE3 Tcw = cur->Tcw_;
    Vector3d Ow = Tcw.inverse().translation();
    const std::vector<MapPoint*> mps = kf->getMapPoints();
    for (int i = 0; i < (int)mps.size(); ++i) {
      MapPoint* mp = mps[i];
      if (!mp || mp->isBad() || found.count(mp)) continue;
      Vector3d pc = Tcw * mp->getPose();
      const float z = pc[2];
      if (z <= 0) continue;
      Vector2d px = cur->camera_->camera2pixel(pc);
      const float u = px[0], v = px[1];
      if (u > xMax || u < xMin || v > yMax || v < yMin) continue;
      const float dist3 = (mp->getPose() - Ow).norm();
      if (dist3 < mp->getMinDistanceThreshold() || dist3 > mp->getMaxDistanceThreshold()) continue;
      const int lp = mp->predictScale(dist3, cur);
      const float rs = radius * kf->scaleFactors_[lp];
      std::vector<int> ids = cur->getFeaturesInArea(u, v, rs, lp - 1, lp + 1);
      if (ids.empty()) continue;
      int best = 256, bestIdx = -1;
      const cv::Mat dl = mp->getDescriptor();
      for (size_t j = 0; j < ids.size(); ++j) {
        const int idx = ids[j];
        if (cur->mappoints_[idx]) continue;
        const int d = computeDistance(dl, cur->descriptors_.row(idx));
        if (d < best) { best = d; bestIdx = idx; }
      }
      if (best <= distThreshold) {
        cur->mappoints_[bestIdx] = mp;
        ++cnt;
        if (checkRot) hist[histBin(kf->unKeypoints_[i].angle - cur->unKeypoints_[bestIdx].angle, true)].push_back(bestIdx);
      }
    }
    if (checkRot) {
      int i1 = -1, i2 = -1, i3 = -1;
      threeMax(hist, 30, i1, i2, i3);
      for (int i = 0; i < 30; ++i)
        if (i != i1 && i != i2 && i != i3)
          for (size_t j = 0; j < hist[i].size(); ++j) { cur->mappoints_[hist[i][j]] = nullptr; --cnt; }
    }
    return cnt;
  }

  int searchByProjection(KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& matchPts, int th) {
    Camera* cam = kf->camera_;                                                                  // matcher.cpp:356-447
    const double sc = Scw.scale();
    Matrix3d Rcw = Scw.rotation_matrix() / sc;
    Vector3d tcw = Scw.translation() / sc;
    Vector3d Ow = -Rcw.transpose() * tcw;
    std::set<MapPoint*> already(matchPts.begin(), matchPts.end());
    already.erase(nullptr);
    int cnt = 0;
    for (size_t i = 0; i < loopPts.size(); ++i) {
      MapPoint* mp = loopPts[i];
      if (!mp || mp->isBad() || already.count(mp)) continue;
      Vector3d pc = Rcw * mp->getPose() + tcw;
      const float z = (float)pc[2];
      if (z < 0) continue;
      const float invz = 1.0f / z;
      const float x = (float)pc[0] * invz, y = (float)pc[1] * invz;
      const float u = cam->fx_ * x + cam->cx_, v = cam->fy_ * y + cam->cy_;
      if (!kf->isInImg(u, v)) continue;
      Vector3d pl = mp->getPose() - Ow;
      const float dist3 = pl.norm();
      if (dist3 < mp->getMinDistanceThreshold() || dist3 > mp->getMaxDistanceThreshold()) continue;
      if (pl.dot(mp->getNormalVector()) < 0.5 * dist3) continue;
      const int lp = mp->predictScale(dist3, kf);
      const float radius = th * kf->scaleFactors_[lp];
      const std::vector<int> ids = kf->getFeaturesInArea(u, v, radius);
      if (ids.empty()) continue;
      cv::Mat dm = mp->getDescriptor();
      int best = 256, bestIdx = -1;
      for (int j = 0; j < (int)ids.size(); ++j) {
        const int idx = ids[j];
        if (matchPts[j]) continue;                          // the reference indexes by the loop counter here (:422)
        const int level = kf->unKeypoints_[idx].octave;
        if (level < lp - 1 || level > lp) continue;
        const int d = computeDistance(dm, kf->descriptors_.row(idx));
        if (d < best) { best = d; bestIdx = idx; }
      }
      if (best <= 50) { matchPts[bestIdx] = mp; ++cnt; }
    }
    return cnt;
  }

  // windowed argmin shared by the two fuse functions (matcher.cpp:1064-1100 with chi2, :1197-1214 without)
  static int fuseArgmin(KeyFrame* kf, MapPoint* mp, float u, float v, float ur, float radius, int lp, bool chi2) {
    const std::vector<int> ids = kf->getFeaturesInArea(u, v, radius);
    int best = 256, bestIdx = -1;
    const cv::Mat dm = mp->getDescriptor();
    for (size_t k = 0; k < ids.size(); ++k) {
      const int idx = ids[k];
      const cv::KeyPoint kp = kf->unKeypoints_[idx];
      if (kp.octave < lp - 1 || kp.octave > lp) continue;
      if (chi2) {
        const float ex = u - kp.pt.x, ey = v - kp.pt.y;
        const float invSigma = 1.0f / kf->scaleFactors_[kp.octave];
        if (kf->uRight_[idx] >= 0) {
          const float er = ur - kf->uRight_[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * invSigma * invSigma > 7.815f) continue;
        } else {
          const float e2 = ex * ex + ey * ey;
          if (e2 * invSigma * invSigma > 5.991f) continue;
        }
      }
      const int d = computeDistance(dm, kf->descriptors_.row(idx));
      if (d < best) { best = d; bestIdx = idx; }
    }
    return best <= 50 ? bestIdx : -1;
  }
  int fuseMapPoints(KeyFrame* kf, std::vector<MapPoint*>& mps, const float& threshold) {   // matcher.cpp:1012-1133
    int cnt = 0;
    Camera* cam = kf->camera_;
    SE3 Tcw = kf->getPose();
    Vector3d Ow = kf->getCamCenter();
    for (size_t i = 0; i < mps.size(); ++i) {
      MapPoint* mp = mps[i];
      if (!mp || mp->isBad() || mp->beObserved(kf)) continue;
      Vector3d pw = mp->getPose();
      Vector3d pc = Tcw * pw;
      const float z = (float)pc[2];
      if (z < 0.0f) continue;
      const float invz = 1.0f / z;
      const float x = (float)pc[0] * invz, y = (float)pc[1] * invz;
      const float u = cam->fx_ * x + cam->cx_, v = cam->fy_ * y + cam->cy_;
      if (!kf->isInImg(u, v)) continue;
      const float ur = u - cam->bf_ * invz;
      Vector3d line = pw - Ow;
      const float dist = line.norm();
      if (dist < mp->getMinDistanceThreshold() || dist > mp->getMaxDistanceThreshold()) continue;
      if (line.dot(mp->getNormalVector()) < 0.5 * dist) continue;
      const int lp = mp->predictScale(dist, kf);
      const int bestIdx = fuseArgmin(kf, mp, u, v, ur, threshold * kf->scaleFactors_[lp], lp, true);
      if (bestIdx < 0) continue;
      MapPoint* org = kf->mappoints_[bestIdx];
      if (org) {
        if (!org->isBad()) { if (org->getObsCnt() > mp->getObsCnt()) mp->replaceMapPoint(org); else org->replaceMapPoint(mp); }
      } else { mp->addObservation(kf, bestIdx); kf->addMapPoint(mp, bestIdx); }
      ++cnt;
    }
    return cnt;
  }
  int fuseByPose(KeyFrame* kf, Sim3& Scw, std::vector<MapPoint*>& loopPts, std::vector<MapPoint*>& replacePts, const float th) {
    Camera* cam = kf->camera_;                                                              // matcher.cpp:1135-1238
    SE3 Tcw(Scw.rotation_matrix(), Scw.translation());
    Vector3d Ow = -Tcw.rotation_matrix().transpose() * Tcw.translation();
    std::set<MapPoint*> already;
    for (size_t i = 0; i < kf->mappoints_.size(); ++i)
      if (kf->mappoints_[i] && !kf->mappoints_[i]->isBad()) already.insert(kf->mappoints_[i]);
    int fused = 0;
    for (size_t i = 0; i < loopPts.size(); ++i) {
      MapPoint* mp = loopPts[i];
      if (!mp || mp->isBad() || already.count(mp)) continue;
      Vector3d pc = Tcw * mp->getPose();
      const float z = (float)pc[2];
      if (z < 0) continue;
      const float invz = 1.0f / z;
      const float x = (float)pc[0] * invz, y = (float)pc[1] * invz;
      const float u = cam->fx_ * x + cam->cx_, v = cam->fy_ * y + cam->cy_;
      if (!kf->isInImg(u, v)) continue;
      Vector3d pl = mp->getPose() - Ow;
      const float dist = pl.norm();
      if (dist < mp->getMinDistanceThreshold() || dist > mp->getMaxDistanceThreshold()) continue;
      if (pl.dot(mp->getNormalVector()) < 0.5 * dist) continue;
      const int lp = mp->predictScale(dist, kf);
      const int bestIdx = fuseArgmin(kf, mp, u, v, 0.f, th * kf->scaleFactors_[lp], lp, false);
      if (bestIdx < 0) continue;
      MapPoint* mk = kf->mappoints_[bestIdx];
      if (mk) { if (!mk->isBad()) replacePts[i] = mk; }
      else { mp->addObservation(kf, bestIdx); kf->addMapPoint(mp, bestIdx); }
      ++fused;
    }
    return fused;
  }

  static bool epipolarOk(const cv::KeyPoint& k1, const cv::KeyPoint& k2, const Matrix3d& F, KeyFrame* kf2) {   // matcher.cpp:1306-1324
    const double p1[3] = {k1.pt.x, k1.pt.y, 1}, p2[3] = {k2.pt.x, k2.pt.y, 1};
    double l2[3];
    for (int j = 0; j < 3; ++j) l2[j] = p1[0] * F(0, j) + p1[1] * F(1, j) + p1[2] * F(2, j);
    const float num = l2[0] * p2[0] + l2[1] * p2[1] + l2[2] * p2[2];
    const float den = l2[0] * l2[0] + l2[1] * l2[1];
    if (den == 0) return false;
    const float sigma = kf2->scaleFactors_[k2.octave];
    return num * num / den < 3.84f * sigma * sigma;
  }
  int searchForTriangulation(KeyFrame* kf1, KeyFrame* kf2, std::vector<std::pair<int, int> >& out, Matrix3d& F12, bool checkRot) {
    int cnt = 0;                                                                             // matcher.cpp:867-1010
    std::vector<int> m12(kf1->N_, -1);
    std::vector<bool> matched2(kf2->N_, false);
    std::vector<int> hist[30];
    const std::vector<MapPoint*> mps1 = kf1->getMapPoints(), mps2 = kf2->getMapPoints();
    const Vector3d C2 = kf2->getPose() * kf1->getCamCenter();
    const Vector2d e = kf2->camera_->camera2pixel(C2);
    const float ex = e[0], ey = e[1];
    FeatureVector::const_iterator a = kf1->featVec_.begin(), ae = kf1->featVec_.end(), b = kf2->featVec_.begin(), be = kf2->featVec_.end();
    while (a != ae && b != be) {
      if (a->first == b->first) {
        for (size_t ik = 0; ik < a->second.size(); ++ik) {
          const unsigned i1 = a->second[ik];
          if (mps1[i1]) continue;
          const bool stereo1 = kf1->uRight_[i1] >= 0;
          int best = 50, bestIdx2 = -1;
          for (size_t ir = 0; ir < b->second.size(); ++ir) {
            const unsigned i2 = b->second[ir];
            if (matched2[i2] || mps2[i2]) continue;
            const bool stereo2 = kf2->uRight_[i2] >= 0;
            const int d = computeDistance(kf1->descriptors_.row(i1), kf2->descriptors_.row(i2));
            if (d > 50 || d > best) continue;
            const cv::KeyPoint k2 = kf2->unKeypoints_[i2];
            if (!stereo1 && !stereo2) {
              const float dx = ex - k2.pt.x, dy = ey - k2.pt.y;
              if (dx * dx + dy * dy < 100 * kf2->scaleFactors_[k2.octave]) continue;
            }
            if (epipolarOk(kf1->unKeypoints_[i1], k2, F12, kf2)) { best = d; bestIdx2 = (int)i2; }
          }
          if (bestIdx2 >= 0) {
            m12[i1] = bestIdx2;
            matched2[bestIdx2] = true;
            if (checkRot) hist[histBin(kf1->unKeypoints_[i1].angle - kf2->unKeypoints_[bestIdx2].angle, false)].push_back((int)i1);
            ++cnt;
          }
        }
        ++a; ++b;
      } else if (a->first < b->first) a = kf1->featVec_.lower_bound(b->first);
      else b = kf2->featVec_.lower_bound(a->first);
    }
    if (checkRot) {
      int i1 = -1, i2 = -1, i3 = -1;
      threeMax(hist, 30, i1, i2, i3);
      for (int i = 0; i < 30; ++i)
        if (i != i1 && i != i2 && i != i3)
          for (size_t j = 0; j < hist[i].size(); ++j) { m12[hist[i][j]] = -1; --cnt; }
    }
    out.clear();
    for (int i = 0; i < (int)m12.size(); ++i) if (m12[i] >= 0) out.push_back(std::make_pair(i, m12[i]));
    return cnt;
  }

  // one direction of searchBySim3 (matcher.cpp:711-780 with badTest/strict = true, :782-846 with false)
  static void sim3Direction(const std::vector<MapPoint*>& mps, const std::vector<bool>& matched, bool badTest, bool strict,
                            const SE3& Tcw, Sim3& S, KeyFrame* target, Camera* cam, float th, std::vector<int>& match) {
    for (int i = 0; i < (int)mps.size(); ++i) {
      MapPoint* mp = mps[i];
      if (!mp || matched[i]) continue;
      if (badTest && mp->isBad()) continue;
      Vector3d po = S * (Tcw * mp->pos_);
      const float z = (float)po[2];
      if (strict ? z < 0 : z <= 0) continue;
      const float invz = 1.0f / z;
      const float x = (float)po[0] * invz, y = (float)po[1] * invz;
      const float u = cam->fx_ * x + cam->cx_, v = cam->fy_ * y + cam->cy_;
      if (!target->isInImg(u, v)) continue;
      const float dist3 = po.norm();
      if (dist3 < mp->getMinDistanceThreshold() || dist3 > mp->getMaxDistanceThreshold()) continue;
      const int lp = mp->predictScale(dist3, target);
      const float radius = th * target->scaleFactors_[lp];
      const std::vector<int> ids = target->getFeaturesInArea(u, v, radius);
      if (ids.empty()) continue;
      const cv::Mat dm = mp->getDescriptor();
      int best = 256, bestIdx = -1;
      for (size_t k = 0; k < ids.size(); ++k) {
        const cv::KeyPoint& kp = target->unKeypoints_[ids[k]];
        if (kp.octave < lp - 1 || kp.octave > lp) continue;
        const int d = computeDistance(dm, target->descriptors_.row(ids[k]));
        if (d < best) { best = d; bestIdx = ids[k]; }
      }
      if (best <= 100) match[i] = bestIdx;
    }
  }
  int searchBySim3(KeyFrame* kf1, KeyFrame* kf2, std::vector<MapPoint*>& matches12, Sim3& S12, const float th) {   // matcher.cpp:679-865
    std::vector<MapPoint*> mps1 = kf1->getMapPoints(), mps2 = kf2->getMapPoints();
    const int N1 = (int)mps1.size(), N2 = (int)mps2.size();
    std::vector<bool> matched1(N1, false), matched2(N2, false);
    Sim3 S21 = S12.inverse();
    for (int i = 0; i < N1; ++i)
      if (matches12[i]) {
        matched1[i] = true;
        const int idx2 = matches12[i]->getIndexInKeyFrame(kf2);
        if (idx2 >= 0 && idx2 < N2) matched2[idx2] = true;
      }
    std::vector<int> match1(N1, -1), match2(N2, -1);
    sim3Direction(mps1, matched1, true, true, kf1->getPose(), S21, kf2, kf1->camera_, th, match1);
    sim3Direction(mps2, matched2, false, false, kf2->getPose(), S12, kf1, kf1->camera_, th, match2);
    int found = 0;
    for (int i = 0; i < N1; ++i)
      if (match1[i] >= 0 && match2[match1[i]] == i) { matches12[i] = mps2[match1[i]]; ++found; }
    return found;
  }

  int searchByProjection(Frame* f, const std::vector<MapPoint*>& mps, const float thRadius) {   // matcher.cpp:274-353
    int cnt = 0;
    for (size_t im = 0; im < mps.size(); ++im) {
      MapPoint* mp = mps[im];
      if (mp->isBad() || !mp->trackInLocalMap_) continue;
      float radius = mp->viewCos_ > 0.998 ? 2.5 : 4.0;
      radius *= thRadius;
      const int lp = mp->trackScaleLevel_;
      const float rs = radius * f->scaleFactors_[lp];
      const std::vector<int> ids = f->getFeaturesInArea(mp->trackProj_u_, mp->trackProj_v_, rs, lp - 1, lp);
      if (ids.empty()) continue;
      int best = 256, bestLevel = -1, best2 = 256, bestLevel2 = -1, bestIdx = -1;
      const cv::Mat dl = mp->getDescriptor();
      for (size_t j = 0; j < ids.size(); ++j) {
        const int idx = ids[j];
        if (f->mappoints_[idx] && f->mappoints_[idx]->getObsCnt() > 0) continue;
        if (f->uRight_[idx] > 0 && fabs(mp->trackProj_uR_ - f->uRight_[idx]) > rs) continue;
        const int d = computeDistance(dl, f->descriptors_.row(idx));
        if (d < best) { best2 = best; best = d; bestLevel2 = bestLevel; bestLevel = f->unKeypoints_[idx].octave; bestIdx = idx; }
        else if (d < best2) { bestLevel2 = f->unKeypoints_[idx].octave; best2 = d; }
      }
      if (best <= 100) {
        if (bestLevel == bestLevel2 && float(best) > ratio_ * float(best2)) continue;
        f->mappoints_[bestIdx] = mp;
        ++cnt;
      }
    }
    return cnt;
  }

  // matcher.cpp:449-559 (kf2 == nullptr) and :561-677 (frame == nullptr): the two BoW searches share the merge walk over the
  // two node-sorted feature vectors; they differ in the exclusion rule, the output indexing and the rounding of the bin.
  int searchByBoW(KeyFrame* kf, Frame* frame, KeyFrame* kf2, std::vector<MapPoint*>& out, bool checkRot) {
    int cnt = 0;
    const bool kk = kf2 != nullptr;
    out.assign(kk ? kf->N_ : frame->N_, nullptr);
    std::vector<bool> matched2(kk ? kf2->N_ : 0, false);
    std::vector<MapPoint*> mps = kf->getMapPoints(), mps2;
    if (kk) mps2 = kf2->getMapPoints();
    const FeatureVector& fb = kk ? kf2->featVec_ : frame->featVec_;
    const cv::Mat& descB = kk ? kf2->descriptors_ : frame->descriptors_;
    const std::vector<cv::KeyPoint>& kpsB = kk ? kf2->unKeypoints_ : frame->unKeypoints_;
    std::vector<int> hist[30];
    FeatureVector::const_iterator a = kf->featVec_.begin(), ae = kf->featVec_.end(), b = fb.begin(), be = fb.end();
    while (a != ae && b != be) {
      if (a->first == b->first) {
        for (size_t ik = 0; ik < a->second.size(); ++ik) {
          const unsigned ia = a->second[ik];
          MapPoint* mpk = mps[ia];
          if (!mpk || mpk->isBad()) continue;
          int best = 256, bestIdx = -1, best2 = 256;
          for (size_t ir = 0; ir < b->second.size(); ++ir) {
            const unsigned ib = b->second[ir];
            if (kk) { if (matched2[ib] || !mps2[ib] || mps2[ib]->isBad()) continue; }
            else if (out[ib]) continue;
            const int d = computeDistance(kf->descriptors_.row(ia), descB.row(ib));
            if (d < best) { best2 = best; best = d; bestIdx = (int)ib; }
            else if (d < best2) best2 = d;
          }
          if (best <= 50 && (float)best < ratio_ * (float)best2) {
            if (kk) { out[ia] = mps2[bestIdx]; matched2[bestIdx] = true; }
            else out[bestIdx] = mpk;
            if (checkRot) hist[histBin(kf->unKeypoints_[ia].angle - kpsB[bestIdx].angle, !kk)].push_back(kk ? (int)ia : bestIdx);
            ++cnt;
          }
        }
        ++a; ++b;
      } else if (a->first < b->first) a = kf->featVec_.lower_bound(b->first);
      else b = fb.lower_bound(a->first);
    }
    if (checkRot) {
      int i1 = -1, i2 = -1, i3 = -1;
      threeMax(hist, 30, i1, i2, i3);
      for (int i = 0; i < 30; ++i)
        if (i != i1 && i != i2 && i != i3)
          for (size_t j = 0; j < hist[i].size(); ++j) { out[hist[i][j]] = nullptr; --cnt; }
    }
    return cnt;
  }
};

}  // namespace myslam
