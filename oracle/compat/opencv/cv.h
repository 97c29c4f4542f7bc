// oracle/compat/opencv/cv.h -- TEST INFRASTRUCTURE (CPU oracle shim) for `#include <opencv/cv.h>`
// (/root/reference/include/myslam/ORBextractor.h:26).
#pragma once
#include "../opencv2/core/core.hpp"
#include "../opencv2/imgproc/imgproc.hpp"
#include "../opencv2/features2d/features2d.hpp"
