// oracle/compat/opencv2/features2d/features2d.hpp -- TEST INFRASTRUCTURE (CPU oracle shim).
// cv::FAST (TYPE_9_16) for the calls at /root/reference/src/ORBextractor.cpp:817,822 and
// KeyPointsFilter::retainBest (only referenced by the dead ComputeKeyPointsOld, :1014,:1032).
#pragma once
#include "../core/core.hpp"

namespace cv {

static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
  Mat img = _img.getMat();
  std::vector<cvp::FastKp> tmp;
  cvp::fast9_16(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression, tmp);
  keypoints.clear();
  keypoints.reserve(tmp.size());
  for (size_t i = 0; i < tmp.size(); ++i)
    keypoints.push_back(KeyPoint((float)tmp[i].x, (float)tmp[i].y, 7.f, -1.f, (float)tmp[i].score));
}

struct KeyPointsFilter {
  // keep the n strongest (dead code path in the reference; semantics: OpenCV keeps ties of the n-th response)
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
      if (n == 0) { kps.clear(); return; }
      std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(),
                       [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      float ambiguous = kps[n - 1].response;
      auto end = std::partition(kps.begin() + n, kps.end(), [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
      kps.resize(end - kps.begin());
    }
  }
};

}  // namespace cv
