// orb_b200_matcher.hpp -- header-only C++ adapter: the reference's Matcher search signatures on top of the C ABI.
//
// The reference's `myslam::Matcher` (include/myslam/matcher.h:16-37) walks Frame / KeyFrame / MapPoint objects through
// raw pointers.  The object walk (null / bad / outlier gates, the Sophus projection, the writes into `mappoints_`) stays
// on the host, verbatim in meaning; the data-parallel inner loops (grid window, Hamming distances, best / second best,
// greedy exclusion, rotation histogram) run in libvoslam_b200.so.  A maintainer keeps matcher.h and replaces the bodies
// in src/matcher.cpp by one-liners:
//
//     #include "orb_b200_matcher.hpp"
//     int Matcher::searchByProjection(Frame* c, Frame* l, const float radius, bool checkRot)
//     { return myslam_b200::searchByProjection(c, l, radius, checkRot); }
//     int Matcher::searchByProjection(Frame* f, const vector<MapPoint*>& mps, const float th)
//     { return myslam_b200::searchByProjection(f, mps, th, ratio_); }
//     int Matcher::searchByBoW(KeyFrame* kf, Frame* f, vector<MapPoint*>& m, bool checkRot)
//     { return myslam_b200::searchByBoW(kf, f, m, checkRot, ratio_); }
//     int Matcher::searchByBoW(KeyFrame* k1, KeyFrame* k2, vector<MapPoint*>& m, bool checkRot)
//     { return myslam_b200::searchByBoWKeyFrames(k1, k2, m, checkRot, ratio_); }
//
// or uses `myslam_b200::MatcherT<Frame, KeyFrame, MapPoint>` as the class itself.  Everything is a template over the
// reference's own types and touches only members the reference's loops touch (named at each use), so the header needs
// neither Sophus, Eigen nor DBoW3 itself.  In this repo it is compiled and run against stand-in types
// (oracle/compat_myslam/myslam_stub.hpp, tests/tools/matcher_adapter_check.cpp), next to a loop-for-loop CPU statement of the
// reference functions over the same objects.
//
// Error convention: the reference returns match counts and cannot fail; the adapter throws std::runtime_error with
// orbx_last_error() when the CUDA path fails (there is no CPU fallback to fall back to).
#pragma once
#include <cstdint>
#include <cstring>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "orb_b200.h"

namespace myslam_b200 {

namespace detail {

inline void check(int rc) {
  if (rc != ORBX_OK) throw std::runtime_error(std::string("libvoslam_b200: ") + orbx_last_error());
}

static_assert(sizeof(orbx_keypoint) == 28, "orbx_keypoint mirrors cv::KeyPoint");

// Flat view of the feature side of a Frame / KeyFrame: unKeypoints_, descriptors_ (n x 32 CV_8U), uRight_, scaleFactors_
// and the image bounds (frame.h:24-43).  cv::KeyPoint is layout-compatible with orbx_keypoint (28 bytes).
template <class FrameT>
struct FrameArrays {
  std::vector<uint8_t> desc, occupied;
  orbx_frame_view view;
  explicit FrameArrays(FrameT* f) {
    const int n = (int)f->unKeypoints_.size();
    desc.resize((size_t)n * 32);
    for (int i = 0; i < n; ++i) std::memcpy(&desc[(size_t)i * 32], f->descriptors_.ptr(i), 32);
    occupied.assign(n > 0 ? n : 1, 0);
    view.kps = reinterpret_cast<const orbx_keypoint*>(f->unKeypoints_.data());
    view.desc = desc.data();
    view.uright = f->uRight_.data();
    view.n = n;
    view.xmin = f->xMin_; view.xmax = f->xMax_; view.ymin = f->yMin_; view.ymax = f->yMax_;
    view.scale_factors = f->scaleFactors_.data();
    view.nlevels = (int)f->scaleFactors_.size();
    view.occupied0 = occupied.data();
  }
};

// Device-resident frames (orbx_frame_t, include/orb_b200.h): one handle per Frame* / KeyFrame*, looked up by the searches below.
// A frame registered here is searched in place in HBM (only the projected points and the "already holds a point" flags go up
// per call) instead of being flattened and re-uploaded.  The handle must describe the frame's CURRENT unKeypoints_ /
// descriptors_ / uRight_ (a Frame's features never change after its constructor; drop the handle in the destructor).
inline std::unordered_map<const void*, orbx_frame_t>& residentMap() { static std::unordered_map<const void*, orbx_frame_t> m; return m; }
inline std::mutex& residentMutex() { static std::mutex m; return m; }
inline orbx_frame_t resident(const void* frame, int n) {
  std::lock_guard<std::mutex> lock(residentMutex());
  std::unordered_map<const void*, orbx_frame_t>::const_iterator it = residentMap().find(frame);
  if (it == residentMap().end()) return nullptr;
  int hn = -1;
  return (orbx_frame_size(it->second, &hn) == ORBX_OK && hn == n) ? it->second : nullptr;
}

// A DBoW3::FeatureVector (std::map<NodeId, std::vector<unsigned>>, iterated in node-id order) as the CSR of orbx_bow_side.
template <class FeatVecT>
struct FeatCsr {
  std::vector<uint32_t> node_ids;
  std::vector<int32_t> group_start, feat_idx;
  explicit FeatCsr(const FeatVecT& fv) {
    group_start.push_back(0);
    for (typename FeatVecT::const_iterator it = fv.begin(); it != fv.end(); ++it) {
      node_ids.push_back((uint32_t)it->first);
      for (size_t k = 0; k < it->second.size(); ++k) feat_idx.push_back((int32_t)it->second[k]);
      group_start.push_back((int32_t)feat_idx.size());
    }
    if (feat_idx.empty()) feat_idx.push_back(0);      // keep data() non-null
    if (node_ids.empty()) node_ids.push_back(0);
  }
  int ngroups() const { return (int)group_start.size() - 1; }
};

template <class FrameLike>
struct BowSide {
  std::vector<uint8_t> desc, valid;
  std::vector<float> angle;
  orbx_bow_side side;
  template <class Csr>
  BowSide(FrameLike* f, const Csr& csr) {
    const int n = (int)f->unKeypoints_.size();
    desc.resize((size_t)(n > 0 ? n : 1) * 32);
    angle.resize(n > 0 ? n : 1);
    valid.assign(n > 0 ? n : 1, 1);
    for (int i = 0; i < n; ++i) {
      std::memcpy(&desc[(size_t)i * 32], f->descriptors_.ptr(i), 32);
      angle[i] = f->unKeypoints_[i].angle;
    }
    side.n = n; side.desc = desc.data(); side.angle = angle.data(); side.valid = valid.data();
    side.ngroups = csr.ngroups(); side.node_ids = csr.node_ids.data(); side.group_start = csr.group_start.data();
    side.feat_idx = csr.feat_idx.data();
  }
};

}  // namespace detail

// Register `handle` (from orbx_frame_create, see orb_b200_frame.hpp constructFrame) as the resident copy of `frame`; an older
// handle of the same object is destroyed.
inline void adoptResident(const void* frame, orbx_frame_t handle) {
  std::lock_guard<std::mutex> lock(detail::residentMutex());
  orbx_frame_t& slot = detail::residentMap()[frame];
  if (slot && slot != handle) orbx_frame_destroy(slot);
  slot = handle;
}
// Upload a host-side Frame / KeyFrame once (key frames that relocalisation / local-map tracking search repeatedly).
template <class FrameT>
orbx_frame_t makeResident(FrameT* frame, int device = 0) {
  detail::FrameArrays<FrameT> a(frame);
  orbx_frame_t h = nullptr;
  detail::check(orbx_frame_upload(&a.view, device, &h));
  adoptResident(frame, h);
  return h;
}
// Call from the frame's destructor (or when its features are rebuilt).
inline void dropResident(const void* frame) {
  std::lock_guard<std::mutex> lock(detail::residentMutex());
  std::unordered_map<const void*, orbx_frame_t>::iterator it = detail::residentMap().find(frame);
  if (it == detail::residentMap().end()) return;
  orbx_frame_destroy(it->second);
  detail::residentMap().erase(it);
}

// Matcher::searchByProjection(Frame* frame_curr, Frame* frame_last, radius, checkRot)          matcher.cpp:18-148
// Host: the per-point gates and the projection of :33-66 (same types, same float casts); device: the window search, the
// "already holds an observed point" exclusion (:90), the stereo gate (:93-99), the strict-< argmin (:104-108), TH_HIGH
// and the rotation histogram (:111-145).  The writes of :113 and :141 are replayed from `assign`.
template <class FrameT>
int searchByProjection(FrameT* frame_curr, FrameT* frame_last, const float radius, bool checkRot = true, int device = 0) {
  const int xMax = frame_curr->xMax_, xMin = frame_curr->xMin_;     // int on purpose: the reference truncates (:28-31)
  const int yMax = frame_curr->yMax_, yMin = frame_curr->yMin_;
  const float b = frame_curr->camera_->b_, bf = frame_curr->camera_->bf_;

  auto Tcw = frame_curr->Tcw_;
  auto Tlc = frame_last->Tcw_ * Tcw.inverse();
  auto tlc = Tlc.translation();
  const bool forward = static_cast<float>(tlc[2]) > b;
  const bool backward = -static_cast<float>(tlc[2]) > b;

  const int M = (int)frame_last->mappoints_.size();
  if (M == 0 || frame_curr->unKeypoints_.empty()) return 0;           // nothing to project / nothing to match: the loops are no-ops
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), has_obs(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), invz(M1, 0.f), angle(M1, 0.f);
  std::vector<int32_t> octave(M1, 0);
  for (int i = 0; i < M; ++i) {
    auto* mp = frame_last->mappoints_[i];
    if (!mp || frame_last->outliers_[i]) continue;
    auto p_camera = Tcw * mp->getPose();
    const float z = static_cast<float>(p_camera[2]);
    if (z < 0.0f) continue;
    auto pixel = frame_curr->camera_->camera2pixel(p_camera);
    const float pu = pixel[0], pv = pixel[1];
    if (pu < xMin || pu > xMax) continue;
    if (pv < yMin || pv > yMax) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv; invz[i] = 1.0f / z;
    octave[i] = frame_last->unKeypoints_[i].octave;
    angle[i] = frame_last->unKeypoints_[i].angle;
    has_obs[i] = mp->observe_cnt_ > 0;
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }

  const int n = (int)frame_curr->unKeypoints_.size();
  std::vector<uint8_t> occupied(n, 0);
  for (int i = 0; i < n; ++i)
    occupied[i] = frame_curr->mappoints_[i] && frame_curr->mappoints_[i]->observe_cnt_ > 0;
  orbx_sbp_frame_points pts = {M, valid.data(), u.data(), v.data(), invz.data(), octave.data(), angle.data(), desc.data(),
                               has_obs.data()};
  std::vector<int32_t> assign(n, -1);
  int match_cnt = 0;
  if (orbx_frame_t rh = detail::resident(frame_curr, n)) {              // the frame is resident in HBM: search it in place
    detail::check(orbx_search_by_projection_frame_h(rh, occupied.data(), &pts, radius, bf, forward, backward, checkRot, assign.data(),
                                                    &match_cnt));
  } else {
    detail::FrameArrays<FrameT> cur(frame_curr);
    cur.view.occupied0 = occupied.data();
    detail::check(orbx_search_by_projection_frame(&cur.view, &pts, radius, bf, forward, backward, checkRot, assign.data(),
                                                  &match_cnt, device));
  }
  for (int i = 0; i < n; ++i) {
    if (assign[i] >= 0) frame_curr->mappoints_[i] = frame_last->mappoints_[assign[i]];
    else if (assign[i] == -2) frame_curr->mappoints_[i] = nullptr;
  }
  return match_cnt;
}

// Matcher::searchByProjection(Frame* frame, const vector<MapPoint*>& mappoints, thRadius)        matcher.cpp:274-353
// `ratio` is the Matcher's ratio_ member (:344).
template <class FrameT, class MapPointT>
int searchByProjection(FrameT* frame, const std::vector<MapPointT*>& mappoints, const float thRadius, float ratio,
                       int device = 0) {
  const int M = (int)mappoints.size();
  if (M == 0 || frame->unKeypoints_.empty()) return 0;
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), has_obs(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), ur(M1, 0.f), view_cos(M1, 0.f);
  std::vector<int32_t> level(M1, 0);
  for (int i = 0; i < M; ++i) {
    MapPointT* mp = mappoints[i];
    if (mp->isBad() || !mp->trackInLocalMap_) continue;
    valid[i] = 1;
    u[i] = mp->trackProj_u_; v[i] = mp->trackProj_v_; ur[i] = mp->trackProj_uR_;
    level[i] = mp->trackScaleLevel_;
    view_cos[i] = mp->viewCos_;
    has_obs[i] = mp->getObsCnt() > 0;
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
  const int n = (int)frame->unKeypoints_.size();
  std::vector<uint8_t> occupied(n, 0);
  for (int i = 0; i < n; ++i) occupied[i] = frame->mappoints_[i] && frame->mappoints_[i]->getObsCnt() > 0;
  orbx_sbp_local_points pts = {M, valid.data(), u.data(), v.data(), ur.data(), level.data(), view_cos.data(), desc.data(),
                               has_obs.data()};
  std::vector<int32_t> assign(n, -1);
  int match_cnt = 0;
  if (orbx_frame_t rh = detail::resident(frame, n)) {
    detail::check(orbx_search_by_projection_local_h(rh, occupied.data(), &pts, thRadius, ratio, assign.data(), &match_cnt));
  } else {
    detail::FrameArrays<FrameT> cur(frame);
    cur.view.occupied0 = occupied.data();
    detail::check(orbx_search_by_projection_local(&cur.view, &pts, thRadius, ratio, assign.data(), &match_cnt, device));
  }
  for (int i = 0; i < n; ++i)
    if (assign[i] >= 0) frame->mappoints_[i] = mappoints[assign[i]];
  return match_cnt;
}

// Matcher::searchByProjection(Frame* frame_curr, KeyFrame* keyframe, radius, distThreshold, found, checkRot)
//                                                                                                matcher.cpp:150-272
// Host: null / bad / `found` gates, projection, depth sign, image bounds, distance range, predictScale (:170-202);
// device: window over [level_predict-1, level_predict+1], "feature already holds a point" exclusion (:218), argmin,
// distThreshold, rotation histogram.  The window radius uses the KEY FRAME's scale factors (:203), like the reference.
template <class FrameT, class KeyFrameT, class FoundSetT>
int searchByProjection(FrameT* frame_curr, KeyFrameT* keyframe, const float radius, const float distThreshold,
                       const FoundSetT& found, bool checkRot = true, int device = 0) {
  const int xMax = frame_curr->xMax_, xMin = frame_curr->xMin_;
  const int yMax = frame_curr->yMax_, yMin = frame_curr->yMin_;
  const auto Tcw = frame_curr->Tcw_;
  auto Ow = Tcw.inverse().translation();
  const auto mappoints = keyframe->getMapPoints();
  const int M = (int)mappoints.size();
  if (M == 0 || frame_curr->unKeypoints_.empty()) return 0;
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), zeros8(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), zerosf(M1, 0.f), angle(M1, 0.f);
  std::vector<int32_t> octave(M1, 0);
  for (int i = 0; i < M; ++i) {
    auto* mp = mappoints[i];
    if (!mp) continue;
    if (mp->isBad() || found.count(mp)) continue;
    auto p_camera = Tcw * mp->getPose();
    const float z = p_camera[2];
    if (z <= 0) continue;
    auto pixel = frame_curr->camera_->camera2pixel(p_camera);
    const float pu = pixel[0], pv = pixel[1];
    if (pu > xMax || pu < xMin) continue;
    if (pv > yMax || pv < yMin) continue;
    auto line = mp->getPose() - Ow;
    const float distance = line.norm();
    const float maxDistance = mp->getMaxDistanceThreshold();
    const float minDistance = mp->getMinDistanceThreshold();
    if (distance < minDistance || distance > maxDistance) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv;
    octave[i] = mp->predictScale(distance, frame_curr);
    angle[i] = keyframe->unKeypoints_[i].angle;
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
  const int n = (int)frame_curr->unKeypoints_.size();
  std::vector<uint8_t> occupied(n, 0);
  for (int i = 0; i < n; ++i) occupied[i] = frame_curr->mappoints_[i] != nullptr;               // :218
  orbx_sbp_frame_points pts = {M, valid.data(), u.data(), v.data(), zerosf.data(), octave.data(), angle.data(), desc.data(),
                               zeros8.data()};
  std::vector<int32_t> assign(n, -1);
  int match_cnt = 0;
  // the window radius uses the KEY FRAME's scale factors (:203); the resident frame carries its own, so the handle is used
  // only when they are the same table (one extractor: always, in the reference)
  orbx_frame_t rh = keyframe->scaleFactors_ == frame_curr->scaleFactors_ ? detail::resident(frame_curr, n) : nullptr;
  if (rh) {
    detail::check(orbx_search_by_projection_reloc_h(rh, occupied.data(), &pts, radius, distThreshold, checkRot, assign.data(), &match_cnt));
  } else {
    detail::FrameArrays<FrameT> cur(frame_curr);
    cur.view.scale_factors = keyframe->scaleFactors_.data();                                   // :203
    cur.view.nlevels = (int)keyframe->scaleFactors_.size();
    cur.view.occupied0 = occupied.data();
    detail::check(orbx_search_by_projection_reloc(&cur.view, &pts, radius, distThreshold, checkRot, assign.data(), &match_cnt,
                                                  device));
  }
  for (int i = 0; i < n; ++i) {
    if (assign[i] >= 0) frame_curr->mappoints_[i] = mappoints[assign[i]];
    else if (assign[i] == -2) frame_curr->mappoints_[i] = nullptr;
  }
  return match_cnt;
}

// Matcher::searchByProjection(KeyFrame* keyframe, Sim3& Scw, loopMapPoints, matchMapPoints, th)       matcher.cpp:356-447
// Host: Sim3 decomposition (:365-368), null / bad / already-found gates (:370-377), projection, isInImg, distance range,
// viewing angle (:379-404), predictScale; device: KeyFrame::getFeaturesInArea window (no level filter), levels
// [level_predict-1, level_predict], TH_LOW, and the reference's `matchMapPoints[j]` test by WINDOW POSITION (:422).
template <class KeyFrameT, class Sim3T, class MapPointT>
int searchByProjection(KeyFrameT* keyframe, Sim3T& Scw, std::vector<MapPointT*>& loopMapPoints,
                       std::vector<MapPointT*>& matchMapPoints, int th, int device = 0) {
  const float fx = keyframe->camera_->fx_, fy = keyframe->camera_->fy_;
  const float cx = keyframe->camera_->cx_, cy = keyframe->camera_->cy_;
  const double scale = Scw.scale();
  auto Rcw = Scw.rotation_matrix() / scale;
  auto tcw = Scw.translation() / scale;
  auto Ow = -Rcw.transpose() * tcw;
  const int M = (int)loopMapPoints.size();
  if (M == 0 || keyframe->unKeypoints_.empty()) return 0;
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), zeros8(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), zerosf(M1, 0.f);
  std::vector<int32_t> octave(M1, 0);
  std::set<MapPointT*> alreadyFound(matchMapPoints.begin(), matchMapPoints.end());          // :370-371
  alreadyFound.erase(static_cast<MapPointT*>(nullptr));
  for (int i = 0; i < M; ++i) {
    MapPointT* mp = loopMapPoints[i];
    if (!mp || mp->isBad() || alreadyFound.count(mp)) continue;
    auto p_cam = Rcw * mp->getPose() + tcw;
    const float z = static_cast<float>(p_cam[2]);
    if (z < 0) continue;
    const float invz = 1.0f / z;
    const float x = static_cast<float>(p_cam[0]) * invz;
    const float y = static_cast<float>(p_cam[1]) * invz;
    const float pu = fx * x + cx, pv = fy * y + cy;
    if (!keyframe->isInImg(pu, pv)) continue;
    const float max_distance = mp->getMaxDistanceThreshold();
    const float min_distance = mp->getMinDistanceThreshold();
    auto pline = mp->getPose() - Ow;
    const float distance = pline.norm();
    if (distance < min_distance || distance > max_distance) continue;
    auto pNormal = mp->getNormalVector();
    if (pline.dot(pNormal) < 0.5 * distance) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv;
    octave[i] = mp->predictScale(distance, keyframe);
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
  detail::FrameArrays<KeyFrameT> kf(keyframe);
  for (int i = 0; i < kf.view.n; ++i) kf.occupied[i] = matchMapPoints[i] != nullptr;
  orbx_sbp_frame_points pts = {M, valid.data(), u.data(), v.data(), zerosf.data(), octave.data(), zerosf.data(), desc.data(),
                               zeros8.data()};
  std::vector<int32_t> assign(kf.view.n > 0 ? kf.view.n : 1, -1);
  int inlier_cnt = 0;
  detail::check(orbx_search_by_projection_sim3(&kf.view, &pts, th, assign.data(), &inlier_cnt, device));
  for (int i = 0; i < kf.view.n; ++i)
    if (assign[i] >= 0) matchMapPoints[i] = loopMapPoints[assign[i]];                         // :439
  return inlier_cnt;
}

// Matcher::searchBySim3(KeyFrame* keyframe1, KeyFrame* keyframe2, matches12, Sim3& S12, th)             matcher.cpp:679-865
// Host: the "already matched" marks (:697-709), both projections with their gates (:715-743, :781-804; note z < 0 vs
// z <= 0 and the isBad() test only in the first direction, as in the reference); device: both windowed TH_HIGH searches
// and the mutual-consistency check (:838-852).
namespace detail {
template <class KeyFrameT, class MapPointT, class TransformT, class PoseT>
void sim3Project(const std::vector<MapPointT*>& mappoints, const std::vector<bool>& matched, bool testBad, bool zStrict,
                 const PoseT& Tcw, TransformT& S, KeyFrameT* target, float fx, float fy, float cx, float cy,
                 std::vector<uint8_t>& valid, std::vector<float>& u, std::vector<float>& v, std::vector<int32_t>& octave,
                 std::vector<uint8_t>& desc) {
  const int N = (int)mappoints.size();
  const int N1 = N > 0 ? N : 1;
  valid.assign(N1, 0); u.assign(N1, 0.f); v.assign(N1, 0.f); octave.assign(N1, 0); desc.assign((size_t)N1 * 32, 0);
  for (int i = 0; i < N; ++i) {
    MapPointT* mp = mappoints[i];
    if (!mp || matched[i]) continue;
    if (testBad && mp->isBad()) continue;
    auto p_world = mp->pos_;
    auto p_own = Tcw * p_world;
    auto p_other = S * p_own;
    const float z = static_cast<float>(p_other[2]);
    if (zStrict ? (z < 0) : (z <= 0)) continue;
    const float invz = 1.0f / z;
    const float x = static_cast<float>(p_other[0]) * invz;
    const float y = static_cast<float>(p_other[1]) * invz;
    const float pu = fx * x + cx, pv = fy * y + cy;
    if (!target->isInImg(pu, pv)) continue;
    const float maxDistance = mp->getMaxDistanceThreshold();
    const float minDistance = mp->getMinDistanceThreshold();
    const float distance = p_other.norm();
    if (distance < minDistance || distance > maxDistance) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv;
    octave[i] = mp->predictScale(distance, target);
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
}
}  // namespace detail

template <class KeyFrameT, class MapPointT, class Sim3T>
int searchBySim3(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<MapPointT*>& matches12, Sim3T& S12, const float th,
                 int device = 0) {
  const float fx = keyframe1->camera_->fx_, fy = keyframe1->camera_->fy_;
  const float cx = keyframe1->camera_->cx_, cy = keyframe1->camera_->cy_;
  std::vector<MapPointT*> mappoints1 = keyframe1->getMapPoints();
  std::vector<MapPointT*> mappoints2 = keyframe2->getMapPoints();
  const int N1 = (int)mappoints1.size(), N2 = (int)mappoints2.size();
  if (N1 == 0 || N2 == 0) return 0;
  std::vector<bool> matched1(N1, false), matched2(N2, false);
  auto Tcw1 = keyframe1->getPose();
  auto Tcw2 = keyframe2->getPose();
  auto S21 = S12.inverse();
  for (int i = 0; i < N1; ++i) {
    MapPointT* mp = matches12[i];
    if (mp) {
      matched1[i] = true;
      const int idx2 = mp->getIndexInKeyFrame(keyframe2);
      if (idx2 >= 0 && idx2 < N2) matched2[idx2] = true;
    }
  }
  std::vector<uint8_t> valid1, valid2, desc1, desc2;
  std::vector<float> u1, v1, u2, v2;
  std::vector<int32_t> oct1, oct2;
  detail::sim3Project(mappoints1, matched1, true, true, Tcw1, S21, keyframe2, fx, fy, cx, cy, valid1, u1, v1, oct1, desc1);
  detail::sim3Project(mappoints2, matched2, false, false, Tcw2, S12, keyframe1, fx, fy, cx, cy, valid2, u2, v2, oct2, desc2);
  std::vector<float> zf((size_t)(N1 > N2 ? N1 : N2) + 1, 0.f);
  std::vector<uint8_t> z8((size_t)(N1 > N2 ? N1 : N2) + 1, 0);
  detail::FrameArrays<KeyFrameT> a1(keyframe1), a2(keyframe2);
  orbx_sbp_frame_points p12 = {N1, valid1.data(), u1.data(), v1.data(), zf.data(), oct1.data(), zf.data(), desc1.data(), z8.data()};
  orbx_sbp_frame_points p21 = {N2, valid2.data(), u2.data(), v2.data(), zf.data(), oct2.data(), zf.data(), desc2.data(), z8.data()};
  std::vector<int32_t> match12(N1 > 0 ? N1 : 1, -1);
  int found = 0;
  detail::check(orbx_search_by_sim3(&a1.view, &p12, &a2.view, &p21, th, match12.data(), &found, device));
  for (int i = 0; i < N1; ++i)
    if (match12[i] >= 0) matches12[i] = mappoints2[match12[i]];                                 // :848
  return found;
}

// Matcher::fuseMapPoints(KeyFrame* keyframe, vector<MapPoint*>& mappoints, threshold)                  matcher.cpp:1012-1133
// The windowed search of point i (:1057-1100, with the chi-square reprojection gate :1073-1095) depends only on the point's
// own geometry and descriptor, so all searches run in one device call.  The gates that the fusing itself can change
// (isBad(), beObserved(keyframe): :1029) are evaluated in the sequential replay, exactly where the reference evaluates them:
// a point whose state an earlier fusion changed is, afterwards, either bad or observed by `keyframe`, i.e. skipped either
// way -- so searching it up front with stale data cannot leak into the result.  replaceMapPoint / addObservation /
// addMapPoint (:1104-1121) stay the reference's own methods, called in point order.
template <class KeyFrameT, class MapPointT>
int fuseMapPoints(KeyFrameT* keyframe, std::vector<MapPointT*>& mappoints, const float& threshold, int device = 0) {
  const int TH_LOW = 50;
  const float fx = keyframe->camera_->fx_, fy = keyframe->camera_->fy_;
  const float cx = keyframe->camera_->cx_, cy = keyframe->camera_->cy_, bf = keyframe->camera_->bf_;
  auto Tcw = keyframe->getPose();
  auto Ow = keyframe->getCamCenter();
  const int M = (int)mappoints.size();
  if (M == 0 || keyframe->unKeypoints_.empty()) return 0;
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), zeros8(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), ur(M1, 0.f), zerosf(M1, 0.f);
  std::vector<int32_t> octave(M1, 0);
  for (int i = 0; i < M; ++i) {
    MapPointT* mp = mappoints[i];
    if (!mp) continue;
    auto p_world = mp->getPose();
    auto pcam = Tcw * p_world;
    const float z = static_cast<float>(pcam[2]);
    if (z < 0.0f) continue;
    const float invz = 1.0f / z;
    const float x = static_cast<float>(pcam[0]) * invz;
    const float y = static_cast<float>(pcam[1]) * invz;
    const float pu = fx * x + cx, pv = fy * y + cy;
    if (!keyframe->isInImg(pu, pv)) continue;
    auto line = p_world - Ow;
    const float dist = line.norm();
    const float minDistance = mp->getMinDistanceThreshold();
    const float maxDistance = mp->getMaxDistanceThreshold();
    if (dist < minDistance || dist > maxDistance) continue;
    auto pn = mp->getNormalVector();
    if (line.dot(pn) < 0.5 * dist) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv; ur[i] = pu - bf * invz;
    octave[i] = mp->predictScale(dist, keyframe);
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
  detail::FrameArrays<KeyFrameT> kf(keyframe);
  orbx_sbp_frame_points pts = {M, valid.data(), u.data(), v.data(), ur.data(), octave.data(), zerosf.data(), desc.data(),
                               zeros8.data()};
  std::vector<int32_t> best(M1, -1);
  detail::check(orbx_window_argmin(&kf.view, &pts, threshold, (float)TH_LOW, 1, best.data(), device));
  int cnt = 0;
  for (int i = 0; i < M; ++i) {
    MapPointT* mp = mappoints[i];
    if (!mp || mp->isBad() || mp->beObserved(keyframe)) continue;                              // :1029, at replay time
    if (best[i] < 0) continue;
    const int bestIdx = best[i];
    MapPointT* mpOrg = keyframe->mappoints_[bestIdx];
    if (mpOrg) {
      if (!mpOrg->isBad()) {
        if (mpOrg->getObsCnt() > mp->getObsCnt()) mp->replaceMapPoint(mpOrg);
        else mpOrg->replaceMapPoint(mp);
      }
    } else {
      mp->addObservation(keyframe, bestIdx);
      keyframe->addMapPoint(mp, bestIdx);
    }
    ++cnt;
  }
  return cnt;
}

// Matcher::fuseByPose(KeyFrame* keyframe, Sim3& Scw, loopMapPoints, replaceMapPoints, th)              matcher.cpp:1135-1238
// Every gate is fixed before the loop (alreadyFound is built once, :1147-1154); only keyframe->mappoints_[bestIdx] is read
// live in the replay (:1217-1228), because addMapPoint of an earlier point can fill it.
template <class KeyFrameT, class Sim3T, class MapPointT>
int fuseByPose(KeyFrameT* keyframe, Sim3T& Scw, std::vector<MapPointT*>& loopMapPoints,
               std::vector<MapPointT*>& replaceMapPoints, const float th, int device = 0) {
  const int TH_LOW = 50;
  const float fx = keyframe->camera_->fx_, fy = keyframe->camera_->fy_;
  const float cx = keyframe->camera_->cx_, cy = keyframe->camera_->cy_;
  typedef decltype(keyframe->getPose()) SE3T;
  SE3T Tcw(Scw.rotation_matrix(), Scw.translation());
  auto Ow = -Tcw.rotation_matrix().transpose() * Tcw.translation();
  const int M = (int)loopMapPoints.size();
  if (M == 0 || keyframe->unKeypoints_.empty()) return 0;
  const int M1 = M;
  std::vector<uint8_t> valid(M1, 0), zeros8(M1, 0), desc((size_t)M1 * 32, 0);
  std::vector<float> u(M1, 0.f), v(M1, 0.f), zerosf(M1, 0.f);
  std::vector<int32_t> octave(M1, 0);
  std::set<MapPointT*> alreadyFound;                                 // the key frame's own good points (:1147-1154)
  for (size_t k = 0; k < keyframe->mappoints_.size(); ++k) {
    MapPointT* mp = keyframe->mappoints_[k];
    if (mp && !mp->isBad()) alreadyFound.insert(mp);
  }
  for (int i = 0; i < M; ++i) {
    MapPointT* mp = loopMapPoints[i];
    if (!mp || mp->isBad() || alreadyFound.count(mp)) continue;
    auto p_cam = Tcw * mp->getPose();
    const float z = static_cast<float>(p_cam[2]);
    if (z < 0) continue;
    const float invz = 1.0f / z;
    const float x = static_cast<float>(p_cam[0]) * invz;
    const float y = static_cast<float>(p_cam[1]) * invz;
    const float pu = fx * x + cx, pv = fy * y + cy;
    if (!keyframe->isInImg(pu, pv)) continue;
    const float max_distance = mp->getMaxDistanceThreshold();
    const float min_distance = mp->getMinDistanceThreshold();
    auto pline = mp->getPose() - Ow;
    const float distance = pline.norm();
    if (distance < min_distance || distance > max_distance) continue;
    auto pNormal = mp->getNormalVector();
    if (pline.dot(pNormal) < 0.5 * distance) continue;
    valid[i] = 1;
    u[i] = pu; v[i] = pv;
    octave[i] = mp->predictScale(distance, keyframe);
    std::memcpy(&desc[(size_t)i * 32], mp->getDescriptor().data, 32);
  }
  detail::FrameArrays<KeyFrameT> kf(keyframe);
  orbx_sbp_frame_points pts = {M, valid.data(), u.data(), v.data(), zerosf.data(), octave.data(), zerosf.data(), desc.data(),
                               zeros8.data()};
  std::vector<int32_t> best(M1, -1);
  detail::check(orbx_window_argmin(&kf.view, &pts, th, (float)TH_LOW, 0, best.data(), device));
  int fused = 0;
  for (int i = 0; i < M; ++i) {
    if (!valid[i] || best[i] < 0) continue;
    MapPointT* mp = loopMapPoints[i];
    MapPointT* mpKF = keyframe->mappoints_[best[i]];
    if (mpKF) {
      if (!mpKF->isBad()) replaceMapPoints[i] = mpKF;
    } else {
      mp->addObservation(keyframe, best[i]);
      keyframe->addMapPoint(mp, best[i]);
    }
    ++fused;
  }
  return fused;
}

// Matcher::searchForTriangulation(KeyFrame* keyframe1, KeyFrame* keyframe2, matchIdxs, Matrix3d& F12, checkRot)
//                                                                                     matcher.cpp:867-1010, 1306-1324
// Host: the epipole of camera 1 in image 2 (:886-890) and the flattening; device: the BoW-guided greedy search between
// features WITHOUT map points, the epipole distance test, the epipolar constraint in double and the rotation histogram.
template <class KeyFrameT, class Matrix3T>
int searchForTriangulation(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<std::pair<int, int> >& matchIdxs,
                           Matrix3T& F12, bool checkRot = true, int device = 0) {
  const int TH_LOW = 50;
  matchIdxs.clear();
  if (keyframe1->unKeypoints_.empty() || keyframe2->unKeypoints_.empty() || keyframe1->featVec_.empty() || keyframe2->featVec_.empty()) return 0;
  const auto mappoints1 = keyframe1->getMapPoints();
  const auto mappoints2 = keyframe2->getMapPoints();
  const auto Cw = keyframe1->getCamCenter();
  const auto C2 = keyframe2->getPose() * Cw;
  const auto C2_pixel = keyframe2->camera_->camera2pixel(C2);
  const float ex = C2_pixel[0], ey = C2_pixel[1];
  typedef typename std::remove_reference<decltype(keyframe1->featVec_)>::type FV;
  detail::FeatCsr<FV> c1(keyframe1->featVec_), c2(keyframe2->featVec_);
  detail::BowSide<KeyFrameT> a(keyframe1, c1), b(keyframe2, c2);
  for (int i = 0; i < a.side.n; ++i) a.valid[i] = mappoints1[i] == nullptr;                    // :902 looks for features without points
  for (int i = 0; i < b.side.n; ++i) b.valid[i] = mappoints2[i] == nullptr;                    // :921
  orbx_tri_side ta = {a.side, reinterpret_cast<const orbx_keypoint*>(keyframe1->unKeypoints_.data()), keyframe1->uRight_.data()};
  orbx_tri_side tb = {b.side, reinterpret_cast<const orbx_keypoint*>(keyframe2->unKeypoints_.data()), keyframe2->uRight_.data()};
  double F[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[r * 3 + c] = F12(r, c);
  std::vector<int32_t> match(a.side.n > 0 ? a.side.n : 1, -1);
  int match_cnt = 0;
  detail::check(orbx_search_for_triangulation(&ta, &tb, F, ex, ey, keyframe2->scaleFactors_.data(),
                                              (int)keyframe2->scaleFactors_.size(), TH_LOW, checkRot, match.data(), &match_cnt,
                                              device));
  matchIdxs.clear();
  matchIdxs.reserve(match_cnt > 0 ? match_cnt : 0);
  for (int i = 0; i < a.side.n; ++i)
    if (match[i] >= 0) matchIdxs.push_back(std::make_pair(i, (int)match[i]));                  // :1001-1007
  return match_cnt;
}

// Matcher::searchByBoW(KeyFrame* keyframe, Frame* frame, mappointMatches, checkRot)              matcher.cpp:449-559
template <class KeyFrameT, class FrameT, class MapPointT>
int searchByBoW(KeyFrameT* keyframe, FrameT* frame, std::vector<MapPointT*>& mappointMatches, bool checkRot, float ratio,
                int device = 0) {
  const int TH_LOW = 50;                                                                       // matcher.cpp:12
  mappointMatches = std::vector<MapPointT*>(frame->N_, static_cast<MapPointT*>(nullptr));
  if (frame->unKeypoints_.empty() || keyframe->unKeypoints_.empty() || frame->featVec_.empty() || keyframe->featVec_.empty()) return 0;
  const std::vector<MapPointT*> mappoints = keyframe->getMapPoints();
  typedef typename std::remove_reference<decltype(keyframe->featVec_)>::type FVa;
  typedef typename std::remove_reference<decltype(frame->featVec_)>::type FVb;
  detail::FeatCsr<FVa> ca(keyframe->featVec_);
  detail::FeatCsr<FVb> cb(frame->featVec_);
  detail::BowSide<KeyFrameT> a(keyframe, ca);
  detail::BowSide<FrameT> bside(frame, cb);
  for (int i = 0; i < a.side.n; ++i) a.valid[i] = mappoints[i] && !mappoints[i]->isBad();      // :476-478
  std::vector<int32_t> match(bside.side.n > 0 ? bside.side.n : 1, -1);
  int match_cnt = 0;
  if (orbx_frame_t rh = detail::resident(frame, bside.side.n))      // descriptors / angles of the frame are read in place
    detail::check(orbx_search_by_bow_h(&a.side, rh, &bside.side, ratio, TH_LOW, checkRot, match.data(), &match_cnt));
  else
    detail::check(orbx_search_by_bow(&a.side, &bside.side, 0, ratio, TH_LOW, checkRot, match.data(), &match_cnt, device));
  for (int i = 0; i < bside.side.n; ++i)
    if (match[i] >= 0) mappointMatches[i] = mappoints[match[i]];                               // :508 (cleared ones stay null, :552)
  return match_cnt;
}

// Matcher::searchByBoW(KeyFrame* keyframe1, KeyFrame* keyframe2, mappointMatches, checkRot)      matcher.cpp:561-677
template <class KeyFrameT, class MapPointT>
int searchByBoWKeyFrames(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<MapPointT*>& mappointMatches, bool checkRot,
                         float ratio, int device = 0) {
  const int TH_LOW = 50;
  mappointMatches = std::vector<MapPointT*>(keyframe1->N_, static_cast<MapPointT*>(nullptr));
  if (keyframe1->unKeypoints_.empty() || keyframe2->unKeypoints_.empty() || keyframe1->featVec_.empty() || keyframe2->featVec_.empty()) return 0;
  const std::vector<MapPointT*> mappoints1 = keyframe1->getMapPoints();
  const std::vector<MapPointT*> mappoints2 = keyframe2->getMapPoints();
  typedef typename std::remove_reference<decltype(keyframe1->featVec_)>::type FV;
  detail::FeatCsr<FV> c1(keyframe1->featVec_), c2(keyframe2->featVec_);
  detail::BowSide<KeyFrameT> a(keyframe1, c1), bside(keyframe2, c2);
  for (int i = 0; i < a.side.n; ++i) a.valid[i] = mappoints1[i] && !mappoints1[i]->isBad();            // :590-592
  for (int i = 0; i < bside.side.n; ++i) bside.valid[i] = mappoints2[i] && !mappoints2[i]->isBad();    // :603-608
  std::vector<int32_t> match(a.side.n > 0 ? a.side.n : 1, -1);
  int match_cnt = 0;
  detail::check(orbx_search_by_bow(&a.side, &bside.side, 1, ratio, TH_LOW, checkRot, match.data(), &match_cnt, device));
  for (int i = 0; i < a.side.n; ++i)
    if (match[i] >= 0) mappointMatches[i] = mappoints2[match[i]];                                      // :629
  return match_cnt;
}

// MapPoint::computeDescriptor()                                                                   mappoint.cpp:118-179
// for MANY map points in one device call (the reference calls it point by point: localMapping.cpp / tracking after new
// observations).  Per point: the descriptors of its good observing key frames in std::map order
// (:131-137), all-pairs Hamming distances, the row with the smallest median (first wins) becomes descriptor_ (:173-176).
// Bad points, points without observations or without a good key frame keep their descriptor, like the early returns.
template <class MapPointT>
void computeDescriptors(const std::vector<MapPointT*>& points, int device = 0) {
  const int P = (int)points.size();
  if (P == 0) return;
  typedef typename std::remove_reference<decltype(points[0]->observedKFs_)>::type ObsMap;
  typedef typename std::remove_reference<decltype(points[0]->descriptor_)>::type MatT;
  std::vector<uint8_t> desc;
  std::vector<int32_t> start(1, 0);
  std::vector<MatT> rows;                                   // the Mat rows behind `desc`, to clone the winner from
  for (int p = 0; p < P; ++p) {
    ObsMap observedKFs;
    bool bad = true;
    if (points[p]) {
      std::unique_lock<decltype(points[p]->mutexFeature_)> lock(points[p]->mutexFeature_);
      bad = points[p]->badFlag_;
      if (!bad) observedKFs = points[p]->observedKFs_;
    }
    if (!bad)
      for (typename ObsMap::iterator it = observedKFs.begin(); it != observedKFs.end(); ++it)
        if (!it->first->isBad()) {
          rows.push_back(it->first->descriptors_.row((int)it->second));
          desc.insert(desc.end(), rows.back().data, rows.back().data + 32);
        }
    start.push_back((int32_t)rows.size());
  }
  if (rows.empty()) return;
  std::vector<int32_t> best(P, -1);
  detail::check(orbx_medoid_descriptors(desc.data(), start.data(), P, best.data(), device));
  for (int p = 0; p < P; ++p)
    if (best[p] >= 0) {
      std::unique_lock<decltype(points[p]->mutexFeature_)> lock(points[p]->mutexFeature_);
      points[p]->descriptor_ = rows[start[p] + best[p]].clone();
    }
}

// Matcher::computeDistance(const Mat&, const Mat&)                                               matcher.cpp:1240-1256
// For parity and completeness: a single 32-byte pair costs a host<->device round trip here, so scalar call sites
// (mappoint.cpp:154) should keep the reference's inline version or be batched (orbx_medoid_descriptors, hamm_knn2).
template <class MatT>
int computeDistance(const MatT& desp1, const MatT& desp2, int device = 0) {
  int32_t idx, d1, d2; uint8_t ok;
  detail::check(hamm_knn2(desp1.data, 1, desp2.data, 1, 256, 1.0f, &idx, &d1, &d2, &ok, device));
  return d1;
}

// The class itself, for callers that prefer to swap the type: `typedef myslam_b200::MatcherT<Frame, KeyFrame, MapPoint>
// Matcher;` gives the call sites of visualOdometry.cpp:238,263,354 / localMapping.cpp:134 the same constructor and the
// same overload set for the searches above.
template <class FrameT, class KeyFrameT, class MapPointT>
class MatcherT {
 public:
  MatcherT() : ratio_(0.f), device_(0) {}
  explicit MatcherT(float ratio, int device = 0) : ratio_(ratio), device_(device) {}
  int searchByProjection(FrameT* frame_curr, FrameT* frame_last, const float radius, bool checkRot = true) {
    return myslam_b200::searchByProjection(frame_curr, frame_last, radius, checkRot, device_);
  }
  int searchByProjection(FrameT* frame, const std::vector<MapPointT*>& mappoints, const float thRadius) {
    return myslam_b200::searchByProjection(frame, mappoints, thRadius, ratio_, device_);
  }
  template <class FoundSetT>
  int searchByProjection(FrameT* frame_curr, KeyFrameT* keyframe, const float radius, const float distThreshold,
                         const FoundSetT& found, bool checkRot = true) {
    return myslam_b200::searchByProjection(frame_curr, keyframe, radius, distThreshold, found, checkRot, device_);
  }
  template <class Sim3T>
  int searchByProjection(KeyFrameT* keyframe, Sim3T& Scw, std::vector<MapPointT*>& loopMapPoints,
                         std::vector<MapPointT*>& matchMapPoints, int th) {
    return myslam_b200::searchByProjection(keyframe, Scw, loopMapPoints, matchMapPoints, th, device_);
  }
  int searchByBoW(KeyFrameT* keyframe, FrameT* frame, std::vector<MapPointT*>& mappointMatches, bool checkRot = true) {
    return myslam_b200::searchByBoW(keyframe, frame, mappointMatches, checkRot, ratio_, device_);
  }
  int searchByBoW(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<MapPointT*>& mappointMatches, bool checkRot) {
    return myslam_b200::searchByBoWKeyFrames(keyframe1, keyframe2, mappointMatches, checkRot, ratio_, device_);
  }
  template <class Sim3T>
  int searchBySim3(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<MapPointT*>& matches12, Sim3T& S12, const float th) {
    return myslam_b200::searchBySim3(keyframe1, keyframe2, matches12, S12, th, device_);
  }
  int fuseMapPoints(KeyFrameT* keyframe, std::vector<MapPointT*>& mappoints, const float& threshold) {
    return myslam_b200::fuseMapPoints(keyframe, mappoints, threshold, device_);
  }
  template <class Sim3T>
  int fuseByPose(KeyFrameT* keyframe, Sim3T& Scw, std::vector<MapPointT*>& loopMapPoints,
                 std::vector<MapPointT*>& replaceMapPoints, const float th) {
    return myslam_b200::fuseByPose(keyframe, Scw, loopMapPoints, replaceMapPoints, th, device_);
  }
  template <class Matrix3T>
  int searchForTriangulation(KeyFrameT* keyframe1, KeyFrameT* keyframe2, std::vector<std::pair<int, int> >& matchIdxs,
                             Matrix3T& F12, bool checkRot = true) {
    return myslam_b200::searchForTriangulation(keyframe1, keyframe2, matchIdxs, F12, checkRot, device_);
  }
  template <class MatT>
  static int computeDistance(const MatT& desp1, const MatT& desp2) { return myslam_b200::computeDistance(desp1, desp2); }

 private:
  float ratio_;
  int device_;
};

}  // namespace myslam_b200
