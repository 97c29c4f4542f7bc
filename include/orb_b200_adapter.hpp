// orb_b200_adapter.hpp -- header-only C++ adapter: the reference's operator API on top of the C ABI (orb_b200.h).
//
// A maintainer of guisongchen/vo_slam_test replaces
//     #include "myslam/ORBextractor.h"            with   #include "orb_b200_adapter.hpp"
// and links libvoslam_b200.so; `ORB_SLAM2::ORBextractor` keeps its constructor, call operator and getters
// (include/myslam/ORBextractor.h:45-108), so Frame::Frame (src/frame.cpp:22) and VisualOdometry
// (src/visualOdometry.cpp:27-31) compile unchanged.  Needs OpenCV's core types (cv::Mat, cv::KeyPoint,
// cv::InputArray, cv::OutputArray); in this repo it is type-checked against oracle/compat (tests/test_abi_and_adapter.py).
//
// Error convention: the reference returns void and asserts; the adapter throws std::runtime_error with
// orbx_last_error() when the CUDA path fails (there is no CPU fallback to fall back to).
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_b200.h"

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
      : h_(nullptr), nlevels_(nlevels), scaleFactor_(scaleFactor) {
    orbx_params p = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device};
    check(orbx_create(&p, &h_));
    check(orbx_max_keypoints(h_, &cap_));
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    check(orbx_scale_factors(h_, mvScaleFactor.data(), nlevels));
    check(orbx_inv_scale_factors(h_, mvInvScaleFactor.data(), nlevels));
    kps_.resize(cap_);
    desc_.resize((size_t)cap_ * 32);
  }
  ~ORBextractor() { orbx_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // ORBextractor::operator()  (ORBextractor.cpp:1051-1112).  Mask is ignored, like in the reference (ORBextractor.h:58).
  void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    // the reference asserts image.type() == CV_8UC1 (ORBextractor.cpp:1058); a multi-channel or non-8-bit Mat must not be
    // read as raw bytes
    if (image.type() != CV_8UC1) throw std::runtime_error("libvoslam_b200: ORBextractor needs a CV_8UC1 image");
    int n = 0;
    check(orbx_extract(h_, image.data, image.cols, image.rows, (size_t)image.step, kps_.data(), desc_.data(), cap_, &n));
    if (n == 0) {
      _descriptors.release();
    } else {
      _descriptors.create(n, 32, CV_8U);
      cv::Mat d = _descriptors.getMat();
      for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
    }
    keypoints.clear();
    keypoints.reserve(n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
    for (int i = 0; i < n; ++i) {
      const orbx_keypoint& k = kps_[i];
      keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id));
    }
  }

  orbx_handle handle() const { return h_; }      // for myslam_b200::constructFrame (orb_b200_frame.hpp)
  int GetLevels() { return nlevels_; }
  float GetScaleFactor() { return scaleFactor_; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }

 protected:
  static void check(int rc) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
  }
  orbx_handle h_;
  int nlevels_, cap_ = 0;
  float scaleFactor_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor;
  std::vector<orbx_keypoint> kps_;
  std::vector<uint8_t> desc_;
};

}  // namespace ORB_SLAM2

namespace myslam_b200 {

// Matcher::computeDistance (matcher.cpp:1240-1256) and the best/second-best loop (matcher.cpp:481-507) on
// descriptor matrices (n x 32, CV_8U, continuous rows).
struct Top2 { std::vector<int32_t> idx, d1, d2; std::vector<uint8_t> ok; };

inline int computeDistance(const cv::Mat& a, const cv::Mat& b, int device = 0) {
  int32_t idx, d1, d2; uint8_t ok;
  if (hamm_knn2(a.data, 1, b.data, 1, 256, 1.0f, &idx, &d1, &d2, &ok, device) != ORBX_OK)
    throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
  return d1;
}

inline Top2 matchTop2(const cv::Mat& queries, const cv::Mat& train, int th_low, float ratio, int device = 0) {
  Top2 r;
  const int nq = queries.rows;
  r.idx.resize(nq); r.d1.resize(nq); r.d2.resize(nq); r.ok.resize(nq);
  if (hamm_knn2(queries.data, nq, train.data, train.rows, th_low, ratio, r.idx.data(), r.d1.data(), r.d2.data(), r.ok.data(),
                device) != ORBX_OK)
    throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
  return r;
}

}  // namespace myslam_b200
