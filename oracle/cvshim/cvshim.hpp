// oracle/cvshim/cvshim.hpp — TEST INFRASTRUCTURE ONLY (part of the oracle).
//
// A minimal stand-in for the handful of OpenCV C++ declarations that the reference's
// src/ORBextractor.cc uses (cv::Mat, KeyPoint, Point_, Size, Rect, InputArray/OutputArray,
// FAST, resize, copyMakeBorder, GaussianBlur, fastAtan2, cvRound/cvFloor/cvCeil,
// KeyPointsFilter). OpenCV's C++ headers are not in this image, so the reference cannot be
// compiled against the real library; with this header it compiles UNMODIFIED, from where it
// lies under /root/reference (oracle/Makefile, target _ref/liborbref.so), which gives the
// tests the reference's own control flow (cell loop, iniTh/minTh fallback, DistributeOctTree,
// IC_Angle, computeOrbDescriptor, level scaling) to pin the oracle restatement against.
//
// The pixel arithmetic of the five primitives is NOT OpenCV's code: it is the oracle's
// restatement (orc_resize_linear, orc_fast, orc_gauss7, orc_fast_atan2 in orb_oracle.cpp),
// each of which tests/test_oracle_vs_cv2.py pins bit-for-bit against python cv2 4.13.0.
// Nothing here is copied from OpenCV or from the reference.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
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

// 8-bit single-channel matrix header over a shared buffer (views share the allocation).
class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step;
  Mat() : rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c) {
    assert(type == CV_8UC1);
  }
  void create(int r, int c, int type) {
    assert(type == CV_8UC1);
    if (data && r == rows && c == cols) return;  // cv::Mat::create keeps a matching allocation (and ROI)
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c + 64);
    rows = r; cols = c; step = (size_t)c; data = buf_->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf_.reset(); rows = cols = 0; data = nullptr; step = 0; }
  // Mat::zeros yields a matrix EXPRESSION in OpenCV: assigning it to a matrix of the same size and type
  // fills that matrix in place (it stays a view of its parent) - computeDescriptors (:1514) relies on this
  // to write into the rows of the caller's descriptor matrix.
  struct ZerosExpr { int rows, cols; };
  static ZerosExpr zeros(int r, int c, int type) { assert(type == CV_8UC1); (void)type; return ZerosExpr{r, c}; }
  Mat(const ZerosExpr& e) : Mat() { *this = e; }
  Mat& operator=(const ZerosExpr& e) {
    create(e.rows, e.cols, CV_8UC1);
    for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols);
    return *this;
  }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
    return m;
  }
  Mat operator()(const Rect& r) const {
    Mat m;
    m.buf_ = buf_; m.rows = r.height; m.cols = r.width; m.step = step; m.data = data + (size_t)r.y * step + r.x;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
  Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  int type() const { return CV_8UC1; }
  int depth() const { return CV_8U; }
  int channels() const { return 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t step1() const { return step; }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step == (size_t)cols; }

 private:
  std::shared_ptr<std::vector<uchar>> buf_;
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
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void create(Size sz, int type) const { m_->create(sz, type); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
  Mat& getMatRef() const { return *m_; }

 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

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
