// orb_b200_compat.hpp — header-only C++ adapter: the reference's ORBextractor / ORBmatcher
// signatures (include/ORBextractor.h:93-162, include/ORBmatcher.h:57-221 of the reference)
// on top of the C ABI of orb_b200.h. Frame.cc / Tracking.cc keep compiling unchanged when
// this header replaces the two reference headers and liborb_b200.so is linked.
//
// With -DORB_B200_HAVE_OPENCV the real cv::Mat / cv::KeyPoint / cv::InputArray are used. Without
// it (this image has no OpenCV C++ headers) a minimal stand-in with the same binary layout for
// KeyPoint and a {ptr, rows, cols, step} Mat view is used so that the adapter still compiles
// and is exercised by tests/test_compat_build.py.
#pragma once
#include <climits>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"

#ifdef ORB_B200_HAVE_OPENCV
#include <opencv2/core/core.hpp>
namespace orbcv = cv;
#else
namespace orbcv {
struct Point2f { float x, y; };
struct KeyPoint {  // same 28-byte layout as cv::KeyPoint
  Point2f pt; float size, angle, response; int octave, class_id;
};
struct Mat {       // a non-owning or owning u8 matrix, enough for the adapter
  int rows = 0, cols = 0; size_t step = 0; unsigned char* data = nullptr;
  std::vector<unsigned char> storage;
  bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
  void create(int r, int c) { rows = r; cols = c; step = (size_t)c; storage.assign((size_t)r * c, 0); data = storage.data(); }
  void release() { rows = cols = 0; step = 0; data = nullptr; storage.clear(); }
  unsigned char* ptr(int r) { return data + (size_t)r * step; }
  const unsigned char* ptr(int r) const { return data + (size_t)r * step; }
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace orbcv
#endif

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
      : nlevels_(nlevels), scaleFactor_(scaleFactor) {
    orb_params p = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
    if (orb_create(&p, device, 1, &h_) != ORB_OK) throw std::runtime_error(std::string("orb_create: ") + orb_last_error());
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    orb_get_scale_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), nullptr);
    cap_ = orb_max_keypoints(h_);
    mvImagePyramid.resize(nlevels);
  }
  ~ORBextractor() { orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // ORBextractor::operator() (src/ORBextractor.cc:1533): mask is ignored, as in the reference.
  void operator()(orbcv::InputArray image_, orbcv::InputArray /*mask*/, std::vector<orbcv::KeyPoint>& keypoints,
                  orbcv::OutputArray descriptors_) {
#ifdef ORB_B200_HAVE_OPENCV
    if (image_.empty()) return;
    cv::Mat image = image_.getMat();
    CV_Assert(image.type() == CV_8UC1);
#else
    const orbcv::Mat& image = image_;
    if (image.empty()) return;
#endif
    static_assert(sizeof(orbcv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte record");
    kpbuf_.resize(cap_);
    descbuf_.resize((size_t)cap_ * 32);
    std::vector<orb_level_view> views(nlevels_);
    int n = 0;
    const int st = orb_extract(h_, image.data, image.cols, image.rows, (size_t)image.step, kpbuf_.data(), cap_, &n,
                               descbuf_.data(), views.data());
    if (st != ORB_OK) throw std::runtime_error(std::string("orb_extract: ") + orb_last_error());
    keypoints.clear();
    keypoints.resize(n);
    if (n) std::memcpy(static_cast<void*>(keypoints.data()), kpbuf_.data(), (size_t)n * sizeof(orb_keypoint));
#ifdef ORB_B200_HAVE_OPENCV
    if (n == 0) descriptors_.release();
    else {
      descriptors_.create(n, 32, CV_8U);
      cv::Mat d = descriptors_.getMat();
      for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &descbuf_[(size_t)i * 32], 32);
    }
    for (int l = 0; l < nlevels_; l++)  // ROI-like views with >= 19 px of readable border (Frame.cc:855)
      mvImagePyramid[l] = cv::Mat(views[l].height, views[l].width, CV_8UC1, views[l].data, (size_t)views[l].step);
#else
    if (n == 0) descriptors_.release();
    else {
      descriptors_.create(n, 32);
      std::memcpy(descriptors_.data, descbuf_.data(), (size_t)n * 32);
    }
    for (int l = 0; l < nlevels_; l++) {
      mvImagePyramid[l].rows = views[l].height; mvImagePyramid[l].cols = views[l].width;
      mvImagePyramid[l].step = (size_t)views[l].step; mvImagePyramid[l].data = views[l].data;
    }
#endif
  }

  int inline GetLevels() { return nlevels_; }
  float inline GetScaleFactor() { return scaleFactor_; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<orbcv::Mat> mvImagePyramid;  // public data member of the reference (ORBextractor.h:162)

 private:
  orb_extractor* h_ = nullptr;
  int nlevels_, cap_ = 0;
  float scaleFactor_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<orb_keypoint> kpbuf_;
  std::vector<unsigned char> descbuf_;
};

// What SearchForInitialization needs from a Frame (mvKeysUn, mDescriptors, image bounds).
// With OpenCV the real Frame class provides exactly these members.
struct FrameLike {
  std::vector<orbcv::KeyPoint> mvKeysUn;
  orbcv::Mat mDescriptors;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
};

class ORBmatcher {
 public:
  static const int TH_LOW = 50;
  static const int TH_HIGH = 100;
  static const int HISTO_LENGTH = 30;

  ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}
  ~ORBmatcher() { orb_matcher_destroy(m_); }

  // ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083)
  static int DescriptorDistance(const orbcv::Mat& a, const orbcv::Mat& b) { return orb_descriptor_distance(a.ptr(0), b.ptr(0)); }

  // ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:573)
  template <class FrameT>
  int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<orbcv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                              int windowSize = 10) {
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    if (!m_ || cap_ < std::max(n1, n2)) {
      orb_matcher_destroy(m_);
      m_ = nullptr;
      cap_ = std::max(std::max(n1, n2), 2048);
      if (orb_matcher_create(device_, 1, cap_, &m_) != ORB_OK) throw std::runtime_error(std::string("orb_matcher_create: ") + orb_last_error());
    }
    std::vector<float> xy1, xy2, a1, a2;
    std::vector<int32_t> o1, o2;
    std::vector<unsigned char> d1, d2;
    pack(F1, xy1, o1, a1, d1);
    pack(F2, xy2, o2, a2, d2);
    orb_frame_view v1 = {n1, xy1.data(), o1.data(), a1.data(), d1.data()};
    orb_frame_view v2 = {n2, xy2.data(), o2.data(), a2.data(), d2.data()};
    orb_match_params mp = {mfNNratio, mbCheckOrientation ? 1 : 0, windowSize, 0, F2.mnMinX, F2.mnMaxX, F2.mnMinY, F2.mnMaxY};
    vnMatches12.assign(n1, -1);
    int nmatches = 0;
    static_assert(sizeof(orbcv::Point2f) == 8, "Point2f layout");
    const int st = orb_search_for_initialization(m_, &v1, &v2, &mp, reinterpret_cast<float*>(vbPrevMatched.data()),
                                                 vnMatches12.data(), &nmatches, nullptr, nullptr);
    if (st != ORB_OK) throw std::runtime_error(std::string("orb_search_for_initialization: ") + orb_last_error());
    return nmatches;
  }

 protected:
  template <class FrameT>
  static void pack(const FrameT& F, std::vector<float>& xy, std::vector<int32_t>& oct, std::vector<float>& ang,
                   std::vector<unsigned char>& desc) {
    const size_t n = F.mvKeysUn.size();
    xy.resize(2 * n); oct.resize(n); ang.resize(n); desc.resize(32 * n);
    for (size_t i = 0; i < n; i++) {
      xy[2 * i] = F.mvKeysUn[i].pt.x; xy[2 * i + 1] = F.mvKeysUn[i].pt.y;
      oct[i] = F.mvKeysUn[i].octave; ang[i] = F.mvKeysUn[i].angle;
      std::memcpy(&desc[32 * i], F.mDescriptors.ptr((int)i), 32);
    }
  }
  float mfNNratio;
  bool mbCheckOrientation;
  int device_;
  orb_matcher* m_ = nullptr;
  int cap_ = 0;
};

}  // namespace ORB_SLAM2
