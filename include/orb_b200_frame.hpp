// orb_b200_frame.hpp -- header-only C++ adapter for what Frame::Frame does between the extractor call and the matcher
// (src/frame.cpp:29-31): undistortKeyPoints (:36-70), findDepth (:108-133) and assignFeaturesToGrid (:72-97), as ONE device
// call (orbx_frame_finish) instead of three host loops + cv::undistortPoints.  A maintainer replaces those three calls in the
// constructor by
//
//     myslam_b200::finishFrame(this, depthImg);
//
// Template over the reference's own Frame type (include/myslam/frame.h:16-71); touches keypoints_, unKeypoints_, depth_,
// uRight_, gridKeypoints_, xMin_.., camera_->{K_, distCoef_, bf_}.  Results are bit-identical to the reference built against
// OpenCV 4.13 (cv::undistortPoints' 5 fixed-point iterations in double; tests/golden/cv2_undistort.npz pins the arithmetic).
// Checked in tests/test_matcher_adapter.py next to the Matcher adapter.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_b200.h"
#include "orb_b200_matcher.hpp"      // adoptResident / dropResident: the registry the Matcher adapter looks frames up in

namespace myslam_b200 {

// Frame::Frame (src/frame.cpp:22-32) as ONE device call: the extractor, undistortKeyPoints, findDepth and assignFeaturesToGrid
// run back to back on the GPU (orbx_frame_create), the members come back in one packed copy, and the frame STAYS RESIDENT:
// the handle is registered for `frame`, so the Matcher adapter's searches against this frame (visualOdometry.cpp:240,265,329,
// 354) read it in HBM instead of re-uploading it.  A maintainer replaces lines 22-31 of the constructor by
//
//     myslam_b200::constructFrame(this, extractor->handle(), imgGray, imgDepth);      // extractor: ORB_SLAM2::ORBextractor adapter
//
// and adds `myslam_b200::dropResident(this);` to ~Frame().  Fills keypoints_, descriptors_, N_, unKeypoints_, uRight_, depth_,
// gridKeypoints_; reads camera_->{K_, distCoef_, bf_} and xMin_ .. yMax_.
template <class FrameT, class MatT>
orbx_frame_t constructFrame(FrameT* frame, orbx_handle extractor, const MatT& gray, const MatT& depthImg) {
  orbx_camera cam;
  std::memset(&cam, 0, sizeof(cam));
  const MatT& K = frame->camera_->K_;
  const MatT& D = frame->camera_->distCoef_;
  cam.fx = K.template at<float>(0, 0); cam.fy = K.template at<float>(1, 1);
  cam.cx = K.template at<float>(0, 2); cam.cy = K.template at<float>(1, 2);
  cam.ndist = D.rows < 8 ? D.rows : 8;
  for (int i = 0; i < cam.ndist; ++i) cam.dist[i] = D.template at<float>(i, 0);
  cam.bf = frame->camera_->bf_;
  cam.xmin = frame->xMin_; cam.xmax = frame->xMax_; cam.ymin = frame->yMin_; cam.ymax = frame->yMax_;
  const bool haveDepth = depthImg.data != nullptr && depthImg.rows > 0 && depthImg.cols > 0;
  orbx_frame_t h = nullptr;
  int n = 0;
  if (orbx_frame_create(extractor, &cam, gray.data, gray.cols, gray.rows, (size_t)gray.step,
                        haveDepth ? reinterpret_cast<const float*>(depthImg.data) : nullptr, haveDepth ? (size_t)depthImg.step : 0, &h,
                        &n) != ORBX_OK)
    throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
  static_assert(sizeof(frame->keypoints_[0]) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
  frame->N_ = n;
  frame->keypoints_.resize(n); frame->unKeypoints_.resize(n);
  frame->uRight_.assign(n, -1.f); frame->depth_.assign(n, -1.f);
  frame->descriptors_ = n > 0 ? MatT(n, 32, 0 /* CV_8UC1 */) : MatT();
  std::vector<int32_t> cell_start(ORBX_GRID_COLS * ORBX_GRID_ROWS + 1, 0), ids(n > 0 ? n : 1);
  if (n > 0) {
    int rc = orbx_frame_get(h, reinterpret_cast<orbx_keypoint*>(frame->keypoints_.data()), frame->descriptors_.data,
                            reinterpret_cast<orbx_keypoint*>(frame->unKeypoints_.data()), frame->uRight_.data(), frame->depth_.data(), n);
    if (rc == ORBX_OK) rc = orbx_frame_grid(h, cell_start.data(), ids.data(), n);
    if (rc != ORBX_OK) { orbx_frame_destroy(h); throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error()); }
  }
  for (int ix = 0; ix < ORBX_GRID_COLS; ++ix)
    for (int iy = 0; iy < ORBX_GRID_ROWS; ++iy) {
      const int c = ix * ORBX_GRID_ROWS + iy;
      frame->gridKeypoints_[ix][iy].assign(ids.begin() + cell_start[c], ids.begin() + cell_start[c + 1]);
    }
  adoptResident(frame, h);
  return h;
}

template <class FrameT, class MatT>
void finishFrame(FrameT* frame, const MatT& depthImg, int device = 0) {
  const int N = (int)frame->keypoints_.size();
  if (N == 0) return;                                                   // frame.cpp:26-27
  static_assert(sizeof(frame->keypoints_[0]) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
  orbx_camera cam;
  std::memset(&cam, 0, sizeof(cam));
  const MatT& K = frame->camera_->K_;                                   // CV_32F 3x3 (camera.cpp:19-21)
  const MatT& D = frame->camera_->distCoef_;                            // CV_32F 4x1 or 5x1 (camera.cpp:27-38)
  cam.fx = K.template at<float>(0, 0); cam.fy = K.template at<float>(1, 1);
  cam.cx = K.template at<float>(0, 2); cam.cy = K.template at<float>(1, 2);
  cam.ndist = D.rows < 8 ? D.rows : 8;
  for (int i = 0; i < cam.ndist; ++i) cam.dist[i] = D.template at<float>(i, 0);
  cam.bf = frame->camera_->bf_;
  cam.xmin = frame->xMin_; cam.xmax = frame->xMax_; cam.ymin = frame->yMin_; cam.ymax = frame->yMax_;

  const int32_t count = N;
  std::vector<int32_t> cell_start(ORBX_GRID_COLS * ORBX_GRID_ROWS + 1), ids(N);
  frame->unKeypoints_.resize(N);
  frame->uRight_.assign(N, -1.f);                                       // :113-114
  frame->depth_.assign(N, -1.f);
  const bool haveDepth = depthImg.data != nullptr && depthImg.rows > 0 && depthImg.cols > 0;
  const int rc = orbx_frame_finish(&cam, reinterpret_cast<const orbx_keypoint*>(frame->keypoints_.data()), &count, 1, N,
                                   haveDepth ? reinterpret_cast<const float*>(depthImg.data) : nullptr, depthImg.cols, depthImg.rows,
                                   (size_t)depthImg.step, (size_t)depthImg.step * (size_t)depthImg.rows,
                                   reinterpret_cast<orbx_keypoint*>(frame->unKeypoints_.data()), frame->uRight_.data(),
                                   frame->depth_.data(), cell_start.data(), ids.data(), device);
  if (rc != ORBX_OK) throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
  for (int ix = 0; ix < ORBX_GRID_COLS; ++ix)                           // CSR (cell = ix * 48 + iy, ids ascending) -> the reference's buckets
    for (int iy = 0; iy < ORBX_GRID_ROWS; ++iy) {
      const int c = ix * ORBX_GRID_ROWS + iy;
      frame->gridKeypoints_[ix][iy].assign(ids.begin() + cell_start[c], ids.begin() + cell_start[c + 1]);
    }
}

}  // namespace myslam_b200
