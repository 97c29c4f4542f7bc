// oracle/compat/opencv2/imgproc/imgproc.hpp -- TEST INFRASTRUCTURE (CPU oracle shim).
// resize / copyMakeBorder / GaussianBlur for CV_8UC1, forwarding to oracle/cvprims.h.
// Stands in for the OpenCV calls at /root/reference/src/ORBextractor.cpp:1094,1129,1131,1137.
#pragma once
#include "../core/core.hpp"

namespace cv {

static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double /*fx*/ = 0, double /*fy*/ = 0,
                          int /*interpolation*/ = INTER_LINEAR) {
  Mat src = _src.getMat();
  _dst.create(dsize.height, dsize.width, CV_8UC1);
  Mat dst = _dst.getMat();
  cvp::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// dst is (src.rows+top+bottom) x (src.cols+left+right); src may already be the centre ROI of dst.
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                                  int /*borderType*/) {
  Mat src = _src.getMat();
  _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
  Mat dst = _dst.getMat();
  uchar* centre = dst.data + (size_t)top * dst.step + left;
  if (centre != src.data)
    for (int y = 0; y < src.rows; ++y) std::memcpy(centre + (size_t)y * dst.step, src.data + (size_t)y * src.step, src.cols);
  for (int y = 0; y < dst.rows; ++y) {
    int sy = cvp::reflect101(y - top, src.rows);
    const uchar* srow = centre + (size_t)sy * dst.step;  // interior already in place
    uchar* drow = dst.data + (size_t)y * dst.step;
    if (y < top || y >= top + src.rows)
      for (int x = 0; x < src.cols; ++x) drow[left + x] = srow[x];
    for (int x = 0; x < left; ++x) drow[x] = srow[cvp::reflect101(x - left, src.cols)];
    for (int x = 0; x < right; ++x) drow[left + src.cols + x] = srow[cvp::reflect101(src.cols + x, src.cols)];
  }
}

static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy,
                                int /*borderType*/ = BORDER_REFLECT_101) {
  Mat src = _src.getMat();
  if (ksize.width != 7 || ksize.height != 7 || sx != 2.0 || sy != 2.0) std::abort();  // only the reference's call
  _dst.create(src.rows, src.cols, CV_8UC1);
  Mat dst = _dst.getMat();
  cvp::gaussian_blur7_s2_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}

}  // namespace cv
