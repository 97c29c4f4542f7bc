// tests/tools/adapter_check.cpp -- compile/link check of include/orb_b200_adapter.hpp against the OpenCV-compat shim,
// and (with a GPU) a run-time comparison of the adapter with the reference's own class on the same frame.
//   g++ -std=c++11 -Ioracle/compat -Iinclude tests/tools/adapter_check.cpp -Lvo_slam_test_b200/lib -lvoslam_b200 -o /tmp/adapter_check
#include <cstdio>
#include <cstring>
#include "opencv/cv.h"
#include "orb_b200_adapter.hpp"

int main(int argc, char** argv) {
  int ndev = 0;
  orbx_device_count(&ndev);
  if (ndev == 0) { std::printf("adapter links; no CUDA device -> compute skipped\n"); return 0; }
  const int W = 640, H = 480;
  cv::Mat img(H, W, CV_8UC1);
  unsigned s = 12345;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) { s = s * 1664525u + 1013904223u; img.at<uchar>(y, x) = (uchar)((((x / 24) + (y / 24)) & 1) * 120 + 60 + (s >> 29)); }
  ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  ex(img, cv::Mat(), kps, desc);
  std::printf("adapter: %zu keypoints, descriptors %dx%d, levels %d, scale[1]=%.7f\n", kps.size(), desc.rows, desc.cols, ex.GetLevels(),
              ex.GetScaleFactors()[1]);
  if (kps.empty() || desc.rows != (int)kps.size()) return 1;
  int d = myslam_b200::computeDistance(desc.row(0), desc.row(1));
  std::printf("computeDistance(row0,row1) = %d\n", d);
  return 0;
}
