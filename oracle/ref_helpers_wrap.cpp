// oracle/ref_helpers_wrap.cpp -- TEST INFRASTRUCTURE: the reference's OWN small Frame / KeyFrame / MapPoint / Camera helpers.
//
// frame.cpp, keyframe.cpp, mappoint.cpp and camera.cpp cannot be compiled as a whole (Eigen, Sophus, DBoW3, cv::FileStorage),
// but the helpers the matcher path depends on are self-contained member functions.  The Makefile pulls exactly these line
// ranges out of the read-only reference tree at build time into a temporary file (REFSRC_GENERATED, never stored in the
// repo) and compiles them here, unmodified, as members of the minimal classes below (the reference's member names):
//   frame.cpp:36-70     Frame::undistortKeyPoints (separately, REFSRC_GENERATED_U: it needs a float Mat, see below)
//   frame.cpp:72-97     Frame::assignFeaturesToGrid, Frame::postionInGrad
//   frame.cpp:108-133   Frame::findDepth
//   frame.cpp:199-247   Frame::getFeaturesInArea
//   keyframe.cpp:64-67  KeyFrame::isInImg
//   keyframe.cpp:268-312 KeyFrame::getFeaturesInArea
//   mappoint.cpp:118-212 MapPoint::computeDescriptor, MapPoint::predictScale (both overloads)
//   camera.cpp:72-75    Camera::camera2pixel
// The extern "C" functions at the bottom only marshal flat arrays in and out, so that tests/test_oracle_vs_reference.py can
// hold the oracle port's restatements (and the stand-in types of compat_myslam) against the reference's own code.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "opencv/cv.h"
#include "myslam_stub.hpp"      // Vector2d / Vector3d only

#define FRAME_GRID_ROWS 48      // camera.h:8-9
#define FRAME_GRID_COLS 64

namespace Eigen { typedef myslam::Vector3d Vector3d; }

namespace refsrc {
using namespace std;
using cv::Mat;
using myslam::Vector2d;
using myslam::Vector3d;

struct Matcher { static int computeDistance(const Mat& a, const Mat& b); };   // matcher.cpp:1240, linked from libmatcherref

struct Camera {
  float fx_, fy_, cx_, cy_, bf_;
  Vector2d camera2pixel(const Vector3d& p3d);
};
struct KeyFrame;
struct Frame {
  Camera* camera_;
  vector<cv::KeyPoint> keypoints_, unKeypoints_;
  vector<float> depth_, uRight_, scaleFactors_;
  size_t N_;
  float xMin_, xMax_, yMin_, yMax_, gridPerPixelWidth_, gridPerPixelHeight_;
  vector<int> gridKeypoints_[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  void findDepth(Mat& depthImg);
  vector<int> getFeaturesInArea(const float& u, const float& v, const float& radius, const int min_level, const int max_level);
  void assignFeaturesToGrid();
  bool postionInGrad(const int& gradNumX, const int& gradNumY);
};
struct KeyFrame {
  vector<cv::KeyPoint> unKeypoints_;
  vector<float> scaleFactors_;
  Mat descriptors_;
  size_t N_;
  float xMin_, xMax_, yMin_, yMax_, gridPerPixelWidth_, gridPerPixelHeight_;
  vector<vector<vector<int> > > gridKeypoints_;
  bool bad_;
  bool isBad() { return bad_; }
  bool isInImg(const float& u, const float& v);
  vector<int> getFeaturesInArea(const float& u, const float& v, const float& radius);
};
struct MapPoint {
  Mat descriptor_;
  float maxDistance_;
  map<KeyFrame*, size_t> observedKFs_;
  mutex mutexFeature_, mutexPose_;
  bool badFlag_;
  void computeDescriptor();
  int predictScale(const float& currDist, Frame* frame);
  int predictScale(const float& currDist, KeyFrame* kf);
};

#include REFSRC_GENERATED

}  // namespace refsrc

// mappoint.cpp:154 calls Matcher::computeDistance: forwarded to the reference's own (matcher.cpp:1240-1256, in libmatcherref.so;
// only the static member's symbol is used, which is why a one-line declaration of the class is enough here)
namespace myslam { class Matcher { public: static int computeDistance(const cv::Mat& desp1, const cv::Mat& desp2); }; }
int refsrc::Matcher::computeDistance(const Mat& a, const Mat& b) { return myslam::Matcher::computeDistance(a, b); }

// ---- Frame::undistortKeyPoints (frame.cpp:36-70), the reference's own loop around cv::undistortPoints -------------------------
// It needs a CV_32F matrix with reshape(), which the byte-only cv::Mat of oracle/compat does not have, so this block brings its
// own minimal float Mat and a cv::undistortPoints built on cvprims.h (the arithmetic pinned against cv2 4.13.0).
#include <memory>
namespace refsrc_u {
using namespace std;
struct Mat {
  int rows, cols, cn;
  std::shared_ptr<std::vector<float> > buf;
  Mat() : rows(0), cols(0), cn(1) {}
  Mat(int r, int c, int /*CV_32F*/) : rows(r), cols(c), cn(1), buf(new std::vector<float>((size_t)r * c, 0.f)) {}
  template <typename T> T& at(int i) { return (*buf)[i]; }
  template <typename T> const T& at(int i) const { return (*buf)[i]; }
  template <typename T> T& at(int i, int j) { return (*buf)[((size_t)i * cols + j) * cn]; }
  Mat reshape(int newCn) const {                       // same data, rows kept: N x 2 x 1 <-> N x 1 x 2 (frame.cpp:55-59)
    Mat m = *this;
    m.cols = cols * cn / newCn; m.cn = newCn;
    return m;
  }
};
namespace cv {
using ::cv::KeyPoint;
inline void undistortPoints(const Mat& src, Mat dst, const Mat& K, const Mat& D, const Mat& /*R*/, const Mat& P) {
  double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int nd = D.rows * D.cols;
  for (int i = 0; i < nd && i < 8; ++i) k[i] = (double)D.at<float>(i);
  for (int i = 0; i < src.rows; ++i) {
    float u, v;
    cvp::undistort_point_k((*src.buf)[2 * i], (*src.buf)[2 * i + 1], K.at<float>(0), K.at<float>(4), K.at<float>(2), K.at<float>(5), k, &u, &v);
    (void)P;                                           // the call site passes P = K (frame.cpp:58), which undistort_point_k applies
    (*dst.buf)[2 * i] = u; (*dst.buf)[2 * i + 1] = v;
  }
}
}  // namespace cv
struct Camera { Mat distCoef_, K_; };
struct Frame {
  Camera* camera_;
  vector<cv::KeyPoint> keypoints_, unKeypoints_;
  size_t N_;
  void undistortKeyPoints();
};
#include REFSRC_GENERATED_U
}  // namespace refsrc_u

extern "C" void refh_undistort(const void* kps, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist, void* unkps) {
  refsrc_u::Camera cam;
  cam.K_ = refsrc_u::Mat(3, 3, CV_32F);
  const float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  for (int i = 0; i < 9; ++i) cam.K_.at<float>(i) = K[i];
  cam.distCoef_ = refsrc_u::Mat(ndist, 1, CV_32F);
  for (int i = 0; i < ndist; ++i) cam.distCoef_.at<float>(i) = dist[i];
  refsrc_u::Frame f;
  f.camera_ = &cam;
  f.keypoints_.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.N_ = n;
  f.undistortKeyPoints();
  std::memcpy(unkps, f.unKeypoints_.data(), sizeof(cv::KeyPoint) * (size_t)n);
}

using namespace refsrc;

static void fill_frame(Frame& f, const cv::KeyPoint* kps, int n, float xmin, float xmax, float ymin, float ymax) {
  f.unKeypoints_.assign(kps, kps + n);
  f.N_ = n;
  f.xMin_ = xmin; f.xMax_ = xmax; f.yMin_ = ymin; f.yMax_ = ymax;
  f.gridPerPixelWidth_ = (float)FRAME_GRID_COLS / (float)(xmax - xmin);       // camera.cpp:47-48
  f.gridPerPixelHeight_ = (float)FRAME_GRID_ROWS / (float)(ymax - ymin);
}

extern "C" {

// Frame::assignFeaturesToGrid -> CSR with cell = ix * 48 + iy (the order the buckets are walked by getFeaturesInArea)
int refh_grid_build(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, int* cell_start, int* ids) {
  Frame f;
  fill_frame(f, (const cv::KeyPoint*)kps, n, xmin, xmax, ymin, ymax);
  f.assignFeaturesToGrid();
  int o = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
    for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
      cell_start[ix * FRAME_GRID_ROWS + iy] = o;
      for (size_t k = 0; k < f.gridKeypoints_[ix][iy].size(); ++k) ids[o++] = f.gridKeypoints_[ix][iy][k];
    }
  cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = o;
  return o;
}

// Frame::getFeaturesInArea (levels >= 0) or KeyFrame::getFeaturesInArea (min_level < 0: no level filter) for nq queries
int refh_features_in_area(const void* kps, int n, float xmin, float xmax, float ymin, float ymax, const float* u, const float* v,
                          const float* r, const int* min_level, const int* max_level, int nq, int* out, int* out_start, int cap) {
  Frame f;
  fill_frame(f, (const cv::KeyPoint*)kps, n, xmin, xmax, ymin, ymax);
  f.assignFeaturesToGrid();
  KeyFrame kf;
  kf.unKeypoints_ = f.unKeypoints_; kf.N_ = n;
  kf.xMin_ = xmin; kf.xMax_ = xmax; kf.yMin_ = ymin; kf.yMax_ = ymax;
  kf.gridPerPixelWidth_ = f.gridPerPixelWidth_; kf.gridPerPixelHeight_ = f.gridPerPixelHeight_;
  kf.gridKeypoints_.assign(FRAME_GRID_COLS, vector<vector<int> >(FRAME_GRID_ROWS));
  for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
    for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) kf.gridKeypoints_[ix][iy] = f.gridKeypoints_[ix][iy];
  int o = 0;
  for (int q = 0; q < nq; ++q) {
    out_start[q] = o;
    const vector<int> ids = min_level[q] < 0 ? kf.getFeaturesInArea(u[q], v[q], r[q]) : f.getFeaturesInArea(u[q], v[q], r[q], min_level[q], max_level[q]);
    for (size_t k = 0; k < ids.size() && o < cap; ++k) out[o++] = ids[k];
  }
  out_start[nq] = o;
  return o;
}

int refh_is_in_img(float xmin, float xmax, float ymin, float ymax, float u, float v) {
  KeyFrame kf; kf.xMin_ = xmin; kf.xMax_ = xmax; kf.yMin_ = ymin; kf.yMax_ = ymax;
  return kf.isInImg(u, v) ? 1 : 0;
}

// Frame::findDepth on an n-keypoint frame: kps = keypoints_, unx = unKeypoints_[i].pt.x
void refh_find_depth(const void* kps, const float* unx, int n, const float* depth, int W, int H, float bf, float* depth_out, float* uright) {
  Frame f; Camera cam; cam.bf_ = bf; f.camera_ = &cam;
  f.keypoints_.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.unKeypoints_ = f.keypoints_;
  for (int i = 0; i < n; ++i) f.unKeypoints_[i].pt.x = unx[i];
  f.N_ = n;
  Mat d(H, W, CV_32F, (void*)depth, (size_t)W * sizeof(float));
  f.findDepth(d);
  for (int i = 0; i < n; ++i) { depth_out[i] = f.depth_[i]; uright[i] = f.uRight_[i]; }
}

int refh_predict_scale(float maxDistance, float currDist, const float* scaleFactors, int nlevels, int keyframe) {
  MapPoint mp; mp.maxDistance_ = maxDistance;
  if (keyframe) { KeyFrame kf; kf.scaleFactors_.assign(scaleFactors, scaleFactors + nlevels); return mp.predictScale(currDist, &kf); }
  Frame f; f.scaleFactors_.assign(scaleFactors, scaleFactors + nlevels);
  return mp.predictScale(currDist, &f);
}

void refh_camera2pixel(float fx, float fy, float cx, float cy, double x, double y, double z, double* u, double* v) {
  Camera c; c.fx_ = fx; c.fy_ = fy; c.cx_ = cx; c.cy_ = cy;
  const Vector2d p = c.camera2pixel(Vector3d(x, y, z));
  *u = p[0]; *v = p[1];
}

// MapPoint::computeDescriptor for npoints map points: CSR of observed descriptors like orbx_medoid_descriptors;
// best[p] = index (relative to start[p]) of the row that became descriptor_, -1 if the function returned early.
void refh_medoid(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best) {
  for (int p = 0; p < npoints; ++p) {
    const int N = start[p + 1] - start[p];
    best[p] = -1;
    if (N <= 0) continue;
    vector<KeyFrame> kfs(N);
    MapPoint mp; mp.badFlag_ = false;
    for (int i = 0; i < N; ++i) {                    // one key frame per observation, ascending addresses = std::map order
      kfs[i].bad_ = false;
      kfs[i].descriptors_ = Mat(1, 32, CV_8UC1);
      std::memcpy(kfs[i].descriptors_.data, desc + (size_t)(start[p] + i) * 32, 32);
      mp.observedKFs_[&kfs[i]] = 0;
    }
    mp.computeDescriptor();
    for (int i = 0; i < N && best[p] < 0; ++i)
      if (std::memcmp(mp.descriptor_.data, kfs[i].descriptors_.data, 32) == 0) best[p] = i;   // first equal row = the winner or its twin
  }
}

}  // extern "C"
