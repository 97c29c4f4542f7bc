// oracle/compat_myslam/myslam/keyframe.h -- TEST INFRASTRUCTURE.  Shadows the reference's include/myslam/keyframe.h (and,
// through it, frame.h / camera.h / common_include.h) when the reference's OWN src/matcher.cpp is compiled in place for the
// oracle (oracle/Makefile, target _ref/libmatcherref.so): the reference's include/myslam/matcher.h is used as it is, its
// `#include "myslam/keyframe.h"` lands here, and Frame / KeyFrame / MapPoint / Camera / SE3 / Sim3 / Matrix3d /
// FeatureVector become the stand-ins of myslam_stub.hpp (Eigen, Sophus and DBoW3 are not installed in this image).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "myslam_stub.hpp"

namespace Sophus { typedef myslam::Sim3 Sim3; typedef myslam::SE3 SE3; }
namespace Eigen { typedef myslam::Matrix3d Matrix3d; typedef myslam::Vector3d Vector3d; }
namespace DBoW3 { typedef myslam::FeatureVector FeatureVector; }
namespace myslam { using namespace std; using cv::Mat; }      // what common_include.h:34-47 provides
