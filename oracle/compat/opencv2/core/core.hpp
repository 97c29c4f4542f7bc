// oracle/compat/opencv2/core/core.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Minimal OpenCV-compatible declarations so that the reference's own, unmodified
// /root/reference/src/ORBextractor.cpp compiles *in place* (see oracle/Makefile) without an
// OpenCV installation.  Only the ~35 symbols that file uses are provided.
// The arithmetic primitives forward to oracle/cvprims.h, which is pinned bit-exact against the
// real cv2 4.13.0.  The same shim type-checks the product's adapter header
// (include/orb_b200_adapter.hpp) in tests.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../cvprims.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

static inline int cvRound(double v) { return cvp::cv_round(v); }
static inline int cvRound(float v) { return cvp::cv_round(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvp::cv_floor(v); }
static inline int cvFloor(float v) { return cvp::cv_floor((double)v); }
static inline int cvCeil(double v) { return cvp::cv_ceil(v); }
static inline int cvCeil(float v) { return cvp::cv_ceil((double)v); }

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  template <typename U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator*=(float s) {
    x = (T)(x * s);
    y = (T)(y * s);
    return *this;
  }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size;
  float angle;
  float response;
  int octave;
  int class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
           int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

// 8-bit single-channel matrix header with malloc-backed, ref-counted storage (malloc, not
// operator new, so the parity build's monotonic operator new only sees STL allocations).
class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step;

  Mat() : rows(0), cols(0), data(nullptr), step(0), buf_(nullptr) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), buf_(nullptr) { create(r, c, type); }
  Mat(Size sz, int type) : rows(0), cols(0), data(nullptr), step(0), buf_(nullptr) {
    create(sz.height, sz.width, type);
  }
  // wrap external memory (no ownership)
  Mat(int r, int c, int /*type*/, void* ext, size_t ext_step)
      : rows(r), cols(c), data((uchar*)ext), step(ext_step ? ext_step : (size_t)c), buf_(nullptr) {}
  Mat(const Mat& o) : rows(o.rows), cols(o.cols), data(o.data), step(o.step), buf_(o.buf_) { addref(); }
  Mat& operator=(const Mat& o) {
    if (this != &o) {
      if (o.buf_) ++o.buf_->refs;
      release();
      rows = o.rows; cols = o.cols; data = o.data; step = o.step; buf_ = o.buf_;
    }
    return *this;
  }
  ~Mat() { release(); }

  void create(int r, int c, int /*type*/) {
    if (data && rows == r && cols == c) return;
    release();
    size_t bytes = (size_t)r * c;
    buf_ = (Buf*)std::malloc(sizeof(Buf) + (bytes ? bytes : 1));
    buf_->refs = 1;
    rows = r; cols = c; step = (size_t)c;
    data = (uchar*)(buf_ + 1);
  }
  void release() {
    if (buf_ && --buf_->refs == 0) std::free(buf_);
    buf_ = nullptr; data = nullptr; rows = cols = 0; step = 0;
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  size_t step1() const { return step; }

  Mat operator()(const Rect& r) const {
    Mat m(*this);
    m.data = data + (size_t)r.y * step + r.x;
    m.rows = r.height; m.cols = r.width;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
    return m;
  }
  // Mat::zeros returns an expression; assigning it to a Mat of the same size fills that Mat IN PLACE
  // (OpenCV MatExpr semantics) -- ORBextractor.cpp:1045 relies on this to write into a row range of the
  // caller's descriptor matrix.
  struct ZerosExpr { int rows, cols, type; };
  static ZerosExpr zeros(int r, int c, int type) { ZerosExpr e = {r, c, type}; return e; }
  Mat(const ZerosExpr& e) : rows(0), cols(0), data(nullptr), step(0), buf_(nullptr) { *this = e; }
  Mat& operator=(const ZerosExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, cols);
    return *this;
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const {
    return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T));
  }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

 private:
  struct Buf { long refs; long pad; };
  Buf* buf_;
  void addref() { if (buf_) ++buf_->refs; }
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return *m_; }
  bool empty() const { return m_->empty(); }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  Mat getMat() const { return *m_; }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
  bool empty() const { return m_->empty(); }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

static inline float fastAtan2(float y, float x) { return cvp::fast_atan2(y, x); }

}  // namespace cv
