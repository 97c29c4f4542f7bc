// tests/tools/matcher_oneliners.cpp -- the replacement bodies of src/matcher.cpp that INTEGRATION.md section 2 gives a maintainer,
// compiled against the reference's OWN include/myslam/matcher.h (only where /root/reference exists; the object types are the
// stand-ins of oracle/compat_myslam because Eigen / Sophus / DBoW3 are not installed).  Proves that the adapter templates bind
// to the reference's exact member signatures: this file defines every Matcher member the reference's matcher.cpp defines.
#include "myslam/matcher.h"
#include "myslam/mappoint.h"

#include "orb_b200_matcher.hpp"

namespace myslam {

Matcher::Matcher(float ratio) : ratio_(ratio) {}

int Matcher::searchByProjection(Frame* frame_curr, Frame* frame_last, const float radius, bool checkRot)
{ return myslam_b200::searchByProjection(frame_curr, frame_last, radius, checkRot); }

int Matcher::searchByProjection(Frame* frame_curr, KeyFrame* keyframe, const float radius, const float distThreshold,
                                const set<MapPoint*>& found, bool checkRot)
{ return myslam_b200::searchByProjection(frame_curr, keyframe, radius, distThreshold, found, checkRot); }

int Matcher::searchByProjection(Frame* frame, const vector<MapPoint*>& mappoints, const float thRadius)
{ return myslam_b200::searchByProjection(frame, mappoints, thRadius, ratio_); }

int Matcher::searchByProjection(KeyFrame* keyframe, Sophus::Sim3& Scw, vector<MapPoint*>& loopMapPoints,
                                vector<MapPoint*>& matchMapPoints, int th)
{ return myslam_b200::searchByProjection(keyframe, Scw, loopMapPoints, matchMapPoints, th); }

int Matcher::searchByBoW(KeyFrame* keyframe, Frame* frame, vector<MapPoint*>& mappointMatches, bool checkRot)
{ return myslam_b200::searchByBoW(keyframe, frame, mappointMatches, checkRot, ratio_); }

int Matcher::searchByBoW(KeyFrame* keyframe1, KeyFrame* keyframe2, vector<MapPoint*>& mappointMatches, bool checkRot)
{ return myslam_b200::searchByBoWKeyFrames(keyframe1, keyframe2, mappointMatches, checkRot, ratio_); }

int Matcher::searchBySim3(KeyFrame* keyframe1, KeyFrame* keyframe2, vector<MapPoint*>& matches12, Sophus::Sim3& S12, const float th)
{ return myslam_b200::searchBySim3(keyframe1, keyframe2, matches12, S12, th); }

int Matcher::computeDistance(const Mat& desp1, const Mat& desp2)
{ return myslam_b200::computeDistance(desp1, desp2); }

int Matcher::searchForTriangulation(KeyFrame* keyframe1, KeyFrame* keyframe2, vector<pair<int, int> >& matchIdxs,
                                    Eigen::Matrix3d& F12, bool checkRot)
{ return myslam_b200::searchForTriangulation(keyframe1, keyframe2, matchIdxs, F12, checkRot); }

int Matcher::fuseMapPoints(KeyFrame* keyframe, vector<MapPoint*>& mappoints, const float& threshold)
{ return myslam_b200::fuseMapPoints(keyframe, mappoints, threshold); }

int Matcher::fuseByPose(KeyFrame* keyframe, Sophus::Sim3& Scw, vector<MapPoint*>& loopMapPoints,
                        vector<MapPoint*>& replaceMapPoints, const float th)
{ return myslam_b200::fuseByPose(keyframe, Scw, loopMapPoints, replaceMapPoints, th); }

}  // namespace myslam
