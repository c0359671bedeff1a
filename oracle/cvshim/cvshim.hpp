// oracle/cvshim/cvshim.hpp — TEST INFRASTRUCTURE ONLY (part of the oracle).
//
// A minimal stand-in for the OpenCV C++ declarations that the reference's src/ORBextractor.cc,
// ORBmatcher.cc, Frame.cc, MapPoint.cc, KeyFrame.cc, Map.cc (and the DBoW2 headers they include)
// use: cv::Mat (8U / 32S / 32F / 64F, views, Mat::zeros / ones / eye expression semantics, small
// float matrix algebra, Mat_ comma initialiser), KeyPoint, Point_, Size, Rect, InputArray /
// OutputArray, FAST, resize, copyMakeBorder, GaussianBlur, fastAtan2, undistortPoints, norm,
// cvRound / cvFloor / cvCeil, KeyPointsFilter, FileStorage / FileNode (declarations only).
// OpenCV's C++ headers are not in this image, so the reference cannot be compiled against the
// real library; with this header those files compile UNMODIFIED, from where they lie under
// /root/reference (oracle/Makefile, target _ref/liborbref.so), which gives the tests the
// reference's own control flow and arithmetic to pin the oracle restatement against.
//
// The arithmetic of the library primitives is NOT OpenCV's code: it is the oracle's restatement
// (orc_resize_linear, orc_fast, orc_gauss7, orc_fast_atan2, orc_undistort_keypoints in
// orb_oracle.cpp; the small-matrix product order below), each pinned bit-for-bit against python
// cv2 4.13.0 (tests/test_oracle_vs_cv2.py, test_oracle_frame.py, test_oracle_search.py).
// Nothing here is copied from OpenCV or from the reference.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

extern "C" {
void orc_resize_linear(const unsigned char* src, int sw, int sh, int sstep, unsigned char* dst, int dw, int dh, int dstep);
int orc_fast(const unsigned char* img, int w, int h, int step, int threshold, int nms, int* xs, int* ys, int* score, int cap);
void orc_gauss7(const unsigned char* src, int w, int h, int sstep, unsigned char* dst, int dstep);
float orc_fast_atan2(float y, float x);
void orc_undistort_keypoints(const void* kin, int n, const float* cam9, void* kout);
// Implemented by oracle/ref_wrap.cpp: allocations made INSIDE a library primitive (scratch buffers, caches)
// are kept out of the per-call bump arena that gives the reference's own allocations a canonical address
// order (see ref_wrap.cpp). They have no influence on the reference's results.
void* cvshim_primitive_enter();
void cvshim_primitive_leave(void* token);
}
struct CvshimPrimitiveScope {
  void* token;
  CvshimPrimitiveScope() : token(cvshim_primitive_enter()) {}
  ~CvshimPrimitiveScope() { cvshim_primitive_leave(token); }
};

// round half to even, as OpenCV's SSE2 / lrint implementations do
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvFloor(float v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvCeil(float v) { return (int)std::ceil(v); }

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
// cv::Point2i(hX*float, 0) in the reference (ORBextractor.cc:717) converts float -> int implicitly at
// the call site (C++ truncation), as with the real cv::Point_<int>(int, int) constructor.
// pt *= scale (:1641): for Point2f and a float factor this is one float product per coordinate
inline Point_<float>& operator*=(Point_<float>& a, float b) { a.x = a.x * b; a.y = a.y * b; return a; }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
};

class KeyPoint {  // field order and sizes of cv::KeyPoint (28 bytes)
 public:
  Point2f pt;
  float size;
  float angle;
  float response;
  int octave;
  int class_id;
  KeyPoint() : pt(0.f, 0.f), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

// Dense 2-D matrix header over a shared buffer (views share the allocation). Depths: 8U (images,
// descriptors), 32S, 32F (poses, calibration, point lists), 64F; 1 or 2 channels (reshape(2) of an N x 2
// point list, Frame.cc:750).
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

inline size_t cvshim_depth_size(int depth) {
  switch (depth) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_64F: return 8; default: return 4; }
}

class _InputArray;
class _OutputArray;

class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step;  // bytes per row
  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), step(0), type_(type) {
    step = step_ ? step_ : (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;  // cv::Mat::create keeps a matching allocation (and ROI)
    type_ = type;
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize() + 64);
    rows = r; cols = c; step = (size_t)c * elemSize(); data = buf_->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf_.reset(); rows = cols = 0; data = nullptr; step = 0; }
  // Mat::zeros / Mat::ones yield a matrix EXPRESSION in OpenCV: assigning it to a matrix of the same size and
  // type fills that matrix in place (it stays a view of its parent) - computeDescriptors (ORBextractor.cc:1514)
  // relies on this to write into the rows of the caller's descriptor matrix.
  struct FillExpr {
    int rows, cols, type;
    double value;
    bool diag;
  };
  static FillExpr zeros(int r, int c, int type) { return FillExpr{r, c, type, 0.0, false}; }
  static FillExpr ones(int r, int c, int type) { return FillExpr{r, c, type, 1.0, false}; }
  static FillExpr eye(int r, int c, int type) { return FillExpr{r, c, type, 1.0, true}; }
  Mat(const FillExpr& e) : Mat() { *this = e; }
  Mat& operator=(const FillExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; y++) {
      memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
      if (e.value != 0.0)
        for (int x = 0; x < cols * channels(); x++)
          if (!e.diag || x == y) set_(y, x, e.value);
    }
    return *this;
  }
  Mat clone() const {
    Mat m;
    if (!data) return m;
    m.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    return m;
  }
  inline void copyTo(const _OutputArray& dst) const;
  inline void convertTo(const _OutputArray& dst, int rtype) const;
  Mat operator()(const Rect& r) const {
    Mat m;
    m.buf_ = buf_; m.type_ = type_; m.rows = r.height; m.cols = r.width; m.step = step;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
  Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat col(int x) const { return colRange(x, x + 1); }
  // reshape(cn): same data, another channel count (continuous rows only; N x 2 C1 <-> N x 1 C2)
  Mat reshape(int cn) const {
    assert(step == (size_t)cols * elemSize());
    Mat m = *this;
    const int total_ch = cols * channels();
    assert(total_ch % cn == 0);
    m.type_ = CV_MAKETYPE(depth(), cn);
    m.cols = total_ch / cn;
    return m;
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  // single index: element i of a row or column vector (cv::Mat::at(int i0))
  template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize1() const { return cvshim_depth_size(depth()); }
  size_t elemSize() const { return elemSize1() * channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t step1() const { return step / elemSize1(); }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  inline Mat t() const;
  inline Mat inv() const;
  inline double dot(const Mat& m) const;
  inline Mat mul(const Mat& m) const;

 private:
  void set_(int y, int x, double v) {
    switch (depth()) {
      case CV_8U: at<uchar>(y, x) = (uchar)v; break;
      case CV_32S: at<int>(y, x) = (int)v; break;
      case CV_32F: at<float>(y, x) = (float)v; break;
      case CV_64F: at<double>(y, x) = v; break;
      default: assert(false);
    }
  }
  int type_;
  std::shared_ptr<std::vector<uchar>> buf_;
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return *m_; }
  bool empty() const { return m_->empty(); }
  Size size() const { return m_->size(); }

 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  // a temporary view (Rwc.copyTo(Twc.rowRange(0,3).colRange(0,3)), KeyFrame.cc): the header is kept by value, and
  // create() with the view's own size and type keeps writing into the parent's buffer
  _OutputArray(const Mat& m) : own_(m), m_(&own_) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void create(Size sz, int type) const { m_->create(sz, type); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
  Mat& getMatRef() const { return *m_; }

 private:
  mutable Mat own_;
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline Mat noArray() { return Mat(); }

inline void Mat::copyTo(const _OutputArray& dst_) const {
  dst_.create(rows, cols, type_);
  Mat dst = dst_.getMat();
  for (int y = 0; y < rows; y++) memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
}
// 8U -> 32F / 32F -> 32F only (Frame.cc:970); may be in place (other type -> new buffer)
inline void Mat::convertTo(const _OutputArray& dst_, int rtype) const {
  assert(channels() == 1 && rtype == CV_32F && (depth() == CV_8U || depth() == CV_32F));
  Mat src = *this;  // keeps the source alive if dst is this matrix
  Mat out(rows, cols, CV_32F);
  for (int y = 0; y < rows; y++)
    for (int x = 0; x < cols; x++) out.at<float>(y, x) = src.depth() == CV_8U ? (float)src.at<uchar>(y, x) : src.at<float>(y, x);
  dst_.getMatRef() = out;
}

// ---- 32F matrix arithmetic. OpenCV evaluates these lazily (MatExpr -> one cv::gemm / addWeighted call); the
// stand-in evaluates eagerly with the same per-element arithmetic for the shapes the reference path uses:
// products of 3x3 / 3x1 / 4x4 float matrices follow cv::gemm's small-matrix case (float products and sums in index
// order); everything else accumulates in double like the general gemm. Both orders are what the oracle restates
// and pins against cv2.gemm (tests/test_oracle_search.py).
inline void cvshim_assert_f32(const Mat& m) { assert(m.type() == CV_32FC1); (void)m; }
inline Mat Mat::t() const {
  cvshim_assert_f32(*this);
  Mat r(cols, rows, CV_32F);
  for (int y = 0; y < rows; y++)
    for (int x = 0; x < cols; x++) r.at<float>(x, y) = at<float>(y, x);
  return r;
}
inline Mat Mat::inv() const { assert(false && "cv::Mat::inv is not on the pinned path"); return Mat(); }
inline double Mat::dot(const Mat& m) const {
  cvshim_assert_f32(*this); cvshim_assert_f32(m);
  assert(rows * cols == m.rows * m.cols);
  double acc = 0;
  const int n = rows * cols;
  for (int i = 0; i < n; i++) {
    const float a = rows == 1 ? at<float>(0, i) : (cols == 1 ? at<float>(i, 0) : at<float>(i / cols, i % cols));
    const float b = m.rows == 1 ? m.at<float>(0, i) : (m.cols == 1 ? m.at<float>(i, 0) : m.at<float>(i / m.cols, i % m.cols));
    acc += (double)a * b;
  }
  return acc;
}
inline Mat Mat::mul(const Mat& m) const {
  cvshim_assert_f32(*this); cvshim_assert_f32(m);
  Mat r(rows, cols, CV_32F);
  for (int y = 0; y < rows; y++)
    for (int x = 0; x < cols; x++) r.at<float>(y, x) = at<float>(y, x) * m.at<float>(y, x);
  return r;
}
inline Mat operator*(const Mat& a, const Mat& b) {
  cvshim_assert_f32(a); cvshim_assert_f32(b);
  assert(a.cols == b.rows);
  Mat r(a.rows, b.cols, CV_32F);
  const int len = a.cols;
  const bool small = len >= 2 && len <= 4 && (len == r.cols || len == r.rows);
  for (int y = 0; y < r.rows; y++)
    for (int x = 0; x < r.cols; x++) {
      if (small) {
        float acc = a.at<float>(y, 0) * b.at<float>(0, x);
        for (int k = 1; k < len; k++) acc = acc + a.at<float>(y, k) * b.at<float>(k, x);
        r.at<float>(y, x) = acc;
      } else {
        double acc = 0;
        for (int k = 0; k < len; k++) acc += (double)a.at<float>(y, k) * b.at<float>(k, x);
        r.at<float>(y, x) = (float)acc;
      }
    }
  return r;
}
#define CVSHIM_ELEMENTWISE(NAME, EXPR)                                               \
  inline Mat NAME(const Mat& a, const Mat& b) {                                      \
    cvshim_assert_f32(a); cvshim_assert_f32(b);                                      \
    assert(a.rows == b.rows && a.cols == b.cols);                                    \
    Mat r(a.rows, a.cols, CV_32F);                                                   \
    for (int y = 0; y < a.rows; y++)                                                 \
      for (int x = 0; x < a.cols; x++) {                                             \
        const float p = a.at<float>(y, x), q = b.at<float>(y, x);                    \
        r.at<float>(y, x) = EXPR;                                                    \
      }                                                                              \
    return r;                                                                        \
  }
CVSHIM_ELEMENTWISE(operator+, p + q)
CVSHIM_ELEMENTWISE(operator-, p - q)
#undef CVSHIM_ELEMENTWISE
inline Mat cvshim_scale(const Mat& a, double s) {  // cv::Mat * scalar = convertTo(alpha): 32F -> 32F scales by (float)alpha
  cvshim_assert_f32(a);
  Mat r(a.rows, a.cols, CV_32F);
  const float fs = (float)s;
  for (int y = 0; y < a.rows; y++)
    for (int x = 0; x < a.cols; x++) r.at<float>(y, x) = a.at<float>(y, x) * fs;
  return r;
}
inline Mat operator*(const Mat& a, double s) { return cvshim_scale(a, s); }
inline Mat operator*(double s, const Mat& a) { return cvshim_scale(a, s); }
inline Mat operator/(const Mat& a, double s) { return cvshim_scale(a, 1.0 / s); }
inline Mat operator-(const Mat& a) { return cvshim_scale(a, -1.0); }
inline Mat operator*(const Mat& a, const Mat::FillExpr& e) { return a * Mat(e); }
inline Mat operator*(double s, const Mat::FillExpr& e) { return cvshim_scale(Mat(e), s); }
inline Mat operator*(const Mat::FillExpr& e, double s) { return cvshim_scale(Mat(e), s); }

// cv::Mat_<float>(r, c) << a, b, c  (Frame.cc:1153)
template <typename T> struct cvshim_depth_of;
template <> struct cvshim_depth_of<float> { enum { value = CV_32F }; };
template <> struct cvshim_depth_of<double> { enum { value = CV_64F }; };
template <> struct cvshim_depth_of<int> { enum { value = CV_32S }; };
template <> struct cvshim_depth_of<uchar> { enum { value = CV_8U }; };
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() : Mat() {}
  Mat_(int r, int c) : Mat(r, c, cvshim_depth_of<T>::value) {}
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};
template <typename T>
struct MatCommaInitializer_ {
  Mat_<T> m;
  int idx;
  template <typename U> MatCommaInitializer_& operator,(U v) {
    m.template at<T>(idx / m.cols, idx % m.cols) = (T)v;
    idx++;
    return *this;
  }
  operator Mat_<T>() const { return m; }
  operator Mat() const { return m; }
};
template <typename T, typename U>
inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
  MatCommaInitializer_<T> ci{m, 0};
  return (ci, v);
}

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
// cv::norm: float inputs, double accumulator
inline double norm(const Mat& a, int normType = NORM_L2) {
  cvshim_assert_f32(a);
  double acc = 0;
  for (int y = 0; y < a.rows; y++)
    for (int x = 0; x < a.cols; x++) {
      const double v = a.at<float>(y, x);
      acc += normType == NORM_L1 ? std::fabs(v) : v * v;
    }
  return normType == NORM_L1 ? acc : std::sqrt(acc);
}
inline double norm(const Mat& a, const Mat& b, int normType = NORM_L2) {
  cvshim_assert_f32(a); cvshim_assert_f32(b);
  assert(a.rows == b.rows && a.cols == b.cols);
  double acc = 0;
  for (int y = 0; y < a.rows; y++)
    for (int x = 0; x < a.cols; x++) {
      const double v = (double)(a.at<float>(y, x) - b.at<float>(y, x));
      acc += normType == NORM_L1 ? std::fabs(v) : v * v;
    }
  return normType == NORM_L1 ? acc : std::sqrt(acc);
}

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::resize for 8-bit single channel, INTER_LINEAR, explicit dsize (the only form the reference uses)
inline void resize(InputArray src_, OutputArray dst_, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR) {
  assert(interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0);
  (void)interpolation;
  Mat src = src_.getMat();
  dst_.create(dsize, CV_8UC1);
  Mat dst = dst_.getMat();
  CvshimPrimitiveScope scope;
  orc_resize_linear(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

inline int cvshim_reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// cv::copyMakeBorder, BORDER_REFLECT_101 (with or without BORDER_ISOLATED: only pixels inside
// src are ever read, which is also what OpenCV does when src is not a sub-matrix, :1716).
// Works in place when src is the interior view of dst (:1695).
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType,
                           const Scalar& = Scalar()) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  (void)borderType;
  Mat src = src_.getMat();
  dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
  Mat dst = dst_.getMat();
  const int w = src.cols, h = src.rows;
  for (int y = 0; y < h; y++) {
    uchar* d = dst.ptr(y + top) + left;
    const uchar* s = src.ptr(y);
    if (d != s) memmove(d, s, (size_t)w);
    for (int k = 1; k <= left; k++) d[-k] = d[cvshim_reflect101(-k, w)];
    for (int k = 1; k <= right; k++) d[w - 1 + k] = d[cvshim_reflect101(w - 1 + k, w)];
  }
  for (int k = 1; k <= top; k++) memcpy(dst.ptr(top - k), dst.ptr(top + cvshim_reflect101(-k, h)), (size_t)dst.cols);
  for (int k = 1; k <= bottom; k++)
    memcpy(dst.ptr(top + h - 1 + k), dst.ptr(top + cvshim_reflect101(h - 1 + k, h)), (size_t)dst.cols);
}

// cv::GaussianBlur, 7x7 sigma 2 BORDER_REFLECT_101 on 8-bit (the only form the reference uses); src may be dst.
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY = 0,
                         int borderType = BORDER_DEFAULT) {
  assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2.0 && sigmaY == 2.0 && borderType == BORDER_REFLECT_101);
  (void)ksize; (void)sigmaX; (void)sigmaY; (void)borderType;
  Mat src = src_.getMat();
  dst_.create(src.rows, src.cols, CV_8UC1);
  CvshimPrimitiveScope scope;
  Mat tmp(src.rows, src.cols, CV_8UC1);
  orc_gauss7(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
  Mat dst = dst_.getMat();
  for (int y = 0; y < src.rows; y++) memcpy(dst.ptr(y), tmp.ptr(y), (size_t)src.cols);
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): FAST-9/16, KeyPoint(x, y, 7.f, -1, score)
inline void FAST(InputArray img_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
  Mat img = img_.getMat();
  keypoints.clear();
  CvshimPrimitiveScope scope;
  const int cap = std::max(0, (img.cols - 6)) * std::max(0, (img.rows - 6));
  if (cap == 0) return;
  std::vector<int> xs(cap), ys(cap), sc(cap);
  int n = orc_fast(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, xs.data(), ys.data(),
                   sc.data(), cap);
  keypoints.reserve(n);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

// cv::undistortPoints(src, dst, K, dist, R = empty, P = K) on an N x 1 two-channel float list (Frame.cc:751, :802);
// arithmetic = the oracle's restatement, pinned bit-for-bit against cv2.undistortPoints (tests/test_oracle_frame.py)
inline void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray dist_, InputArray R_, InputArray P_) {
  Mat src = src_.getMat(), K = K_.getMat(), D = dist_.getMat(), P = P_.getMat();
  assert(R_.empty() && src.type() == CV_32FC2 && src.isContinuous() && K.type() == CV_32FC1 && P.data == K.data);
  (void)P;
  const int n = src.rows * src.cols, nd = D.rows * D.cols;
  float cam[9] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), 0, 0, 0, 0, 0};
  for (int i = 0; i < nd && i < 5; i++) cam[4 + i] = D.at<float>(i);
  if (cam[4] == 0.0f) cam[4] = 1e-30f;  // the oracle entry point treats k1 == 0 as "no distortion" (Frame.cc:728), OpenCV does not
  std::vector<KeyPoint> in(n), out(n);
  const float* s = src.ptr<float>();
  for (int i = 0; i < n; i++) { in[i].pt.x = s[2 * i]; in[i].pt.y = s[2 * i + 1]; }
  orc_undistort_keypoints(in.data(), n, cam, out.data());
  dst_.create(src.rows, src.cols, CV_32FC2);
  float* d = dst_.getMat().ptr<float>();
  for (int i = 0; i < n; i++) { d[2 * i] = out[i].pt.x; d[2 * i + 1] = out[i].pt.y; }
}

// cv::FileStorage / cv::FileNode: only named by DBoW2's TemplatedVocabulary::save/load (YAML vocabulary files),
// which the pinned path never runs. Declarations that let those templates parse; every use aborts.
class FileNode {
 public:
  FileNode operator[](const char*) const { abort(); }
  FileNode operator[](const std::string&) const { abort(); }
  FileNode operator[](int) const { abort(); }
  size_t size() const { abort(); }
  operator int() const { abort(); }
  operator float() const { abort(); }
  operator double() const { abort(); }
  operator std::string() const { abort(); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) { abort(); }
  bool isOpened() const { abort(); }
  void release() {}
  FileNode operator[](const char*) const { abort(); }
  FileNode operator[](const std::string&) const { abort(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { abort(); return fs; }

// Only referenced by the reference's dead ComputeKeyPointsOld (:1188); must link, is never run.
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n < 0 || (int)kps.size() <= n) return;
    std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    float thr = n > 0 ? kps[n - 1].response : 0.f;
    size_t m = n;
    while (m < kps.size() && kps[m].response >= thr && n > 0) m++;
    kps.resize(m);
  }
};

}  // namespace cv
