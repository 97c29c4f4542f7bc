// oracle/compat_myslam/myslam/mappoint.h -- TEST INFRASTRUCTURE: shadows include/myslam/mappoint.h (see keyframe.h here).
#pragma once
#include "myslam/keyframe.h"
