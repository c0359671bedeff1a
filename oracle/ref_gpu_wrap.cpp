// oracle/ref_gpu_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REFERENCE's own ORB_SLAM2::Frame / ORBmatcher, compiled unmodified from
// /root/reference/src/{Frame,ORBmatcher,MapPoint,KeyFrame,Map}.cc, but linked against the DROP-IN instead of the
// reference's src/ORBextractor.cc: orb_slam2_detailed_comments_b200/compat/orb_b200_extractor.cpp (constructor +
// operator()), orb_b200_matcher.cpp (DescriptorDistance, SearchForInitialization) and orb_b200_frame.cpp
// (ComputeStereoMatches) on top of liborb_b200.so. Built into oracle/_ref/liborbref_gpu.so by `make -C oracle refgpu`.
// The GPU tests run the reference's real Frame constructors (src/Frame.cc:121-158 stereo, :313-350 monocular) through
// this library and compare every member they fill with what oracle/_ref/liborbref.so (the all-CPU reference) computes.
#include <cstring>
#include <vector>

#include "cvshim.hpp"
#define private public
#define protected public
#include "ORBextractor.h"  // the reference's headers: -I/root/reference/include
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

// (oracle/ref_wrap.cpp is compiled into this library too: its orbref_* entry points - frames from given keypoints, the
// searches on live MapPoint / KeyFrame objects - then run the drop-in for the replaced members and the reference's CPU
// code for everything else, with the same signatures as in the all-CPU library.)

namespace {
struct GpuFrame {
  ORB_SLAM2::Frame f;
  std::string error;
};
thread_local std::string t_error;

void camera(const float* cam9, cv::Mat& K, cv::Mat& dist) {
  K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = cam9[0]; K.at<float>(1, 1) = cam9[1]; K.at<float>(0, 2) = cam9[2]; K.at<float>(1, 2) = cam9[3];
  const int nd = cam9[8] != 0.0f ? 5 : 4;  // Tracking.cc: k3 is appended only when it is non-zero
  dist = cv::Mat(nd, 1, CV_32F);
  for (int i = 0; i < nd; i++) dist.at<float>(i) = cam9[4 + i];
}
}  // namespace

extern "C" {

const char* orbgpu_last_error() { return t_error.c_str(); }

// new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST), src/Tracking.cc:175-188
void* orbgpu_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  try {
    return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  } catch (const std::exception& e) {
    t_error = e.what();
    return nullptr;
  }
}
void orbgpu_extractor_destroy(void* ex) { delete (ORB_SLAM2::ORBextractor*)ex; }

// (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors), src/Frame.cc:437
int orbgpu_extract(void* ex, const unsigned char* img, int w, int h, int step, void* kps, unsigned char* desc, int cap) {
  try {
    cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)step), d;
    std::vector<cv::KeyPoint> k;
    (*(ORB_SLAM2::ORBextractor*)ex)(image, cv::Mat(), k, d);
    const int n = (int)k.size();
    if (n > cap) { t_error = "capacity"; return -1; }
    if (n) {
      memcpy(kps, k.data(), (size_t)n * sizeof(cv::KeyPoint));
      for (int i = 0; i < n; i++) memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    return n;
  } catch (const std::exception& e) {
    t_error = e.what();
    return -1;
  }
}
// mvImagePyramid[l] with its 19-px border, packed (h+38) x (w+38)
int orbgpu_pyramid_level(void* ex, int l, int* w, int* h, unsigned char* dst) {
  ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)ex;
  const cv::Mat& m = e->mvImagePyramid[l];
  *w = m.cols; *h = m.rows;
  if (dst)
    for (int y = -19; y < m.rows + 19; y++) memcpy(dst + (size_t)(y + 19) * (m.cols + 38), m.data + (long long)y * (long long)m.step - 19, m.cols + 38);
  return 0;
}

// Frame(imGray, timeStamp, extractor, voc, K, distCoef, bf, thDepth), src/Frame.cc:313-350 (monocular)
void* orbgpu_frame_mono(void* ex, const unsigned char* img, int w, int h, int step, const float* cam9, float bf, float thDepth) {
  try {
    cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)step), K, dist;
    camera(cam9, K, dist);
    ORB_SLAM2::Frame::mbInitialComputations = true;   // the statics (image bounds, grid cell size) follow this camera
    GpuFrame* g = new GpuFrame();
    g->f = ORB_SLAM2::Frame(image, 0.0, (ORB_SLAM2::ORBextractor*)ex, (ORB_SLAM2::ORBVocabulary*)nullptr, K, dist, bf, thDepth);
    return g;
  } catch (const std::exception& e) {
    t_error = e.what();
    return nullptr;
  }
}
// Frame(imLeft, imRight, timeStamp, extractorLeft, extractorRight, voc, K, distCoef, bf, thDepth), src/Frame.cc:121-158:
// two extractor threads (:146-154), then ComputeStereoMatches (:157)
void* orbgpu_frame_stereo(void* exL, void* exR, const unsigned char* imgL, const unsigned char* imgR, int w, int h, int step,
                          const float* cam9, float bf, float thDepth) {
  try {
    cv::Mat L(h, w, CV_8UC1, (void*)imgL, (size_t)step), R(h, w, CV_8UC1, (void*)imgR, (size_t)step), K, dist;
    camera(cam9, K, dist);
    ORB_SLAM2::Frame::mbInitialComputations = true;
    GpuFrame* g = new GpuFrame();
    g->f = ORB_SLAM2::Frame(L, R, 0.0, (ORB_SLAM2::ORBextractor*)exL, (ORB_SLAM2::ORBextractor*)exR, (ORB_SLAM2::ORBVocabulary*)nullptr, K,
                            dist, bf, thDepth);
    return g;
  } catch (const std::exception& e) {
    t_error = e.what();
    return nullptr;
  }
}
void orbgpu_frame_destroy(void* h) { delete (GpuFrame*)h; }

int orbgpu_frame_n(void* h) { return ((GpuFrame*)h)->f.N; }
int orbgpu_frame_n_right(void* h) { return (int)((GpuFrame*)h)->f.mvKeysRight.size(); }
// which: 0 mvKeys, 1 mvKeysUn, 2 mvKeysRight
void orbgpu_frame_keys(void* h, int which, void* out) {
  const ORB_SLAM2::Frame& f = ((GpuFrame*)h)->f;
  const std::vector<cv::KeyPoint>& v = which == 0 ? f.mvKeys : (which == 1 ? f.mvKeysUn : f.mvKeysRight);
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(cv::KeyPoint));
}
void orbgpu_frame_descriptors(void* h, int right, unsigned char* out) {
  const ORB_SLAM2::Frame& f = ((GpuFrame*)h)->f;
  const cv::Mat& d = right ? f.mDescriptorsRight : f.mDescriptors;
  for (int i = 0; i < d.rows; i++) memcpy(out + (size_t)i * 32, d.ptr(i), 32);
}
void orbgpu_frame_stereo_vectors(void* h, float* uRight, float* depth) {
  const ORB_SLAM2::Frame& f = ((GpuFrame*)h)->f;
  for (int i = 0; i < f.N; i++) { uRight[i] = f.mvuRight[i]; depth[i] = f.mvDepth[i]; }
}
void orbgpu_frame_bounds(float* b4) {
  b4[0] = ORB_SLAM2::Frame::mnMinX; b4[1] = ORB_SLAM2::Frame::mnMaxX; b4[2] = ORB_SLAM2::Frame::mnMinY; b4[3] = ORB_SLAM2::Frame::mnMaxY;
}
// mGrid as CSR: cell = ix*48 + iy, items in the order AssignFeaturesToGrid pushed them
void orbgpu_frame_grid(void* h, int* cellStart /* 64*48+1 */, int* items /* N */) {
  const ORB_SLAM2::Frame& f = ((GpuFrame*)h)->f;
  int pos = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
      cellStart[ix * FRAME_GRID_ROWS + iy] = pos;
      for (size_t idx : f.mGrid[ix][iy]) items[pos++] = (int)idx;
    }
  cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
}
void orbgpu_frame_scale_tables(void* h, int* nlevels, float* scale, float* invScale, float* sigma2, float* invSigma2) {
  const ORB_SLAM2::Frame& f = ((GpuFrame*)h)->f;
  *nlevels = f.mnScaleLevels;
  for (int l = 0; l < f.mnScaleLevels; l++) {
    scale[l] = f.mvScaleFactors[l]; invScale[l] = f.mvInvScaleFactors[l]; sigma2[l] = f.mvLevelSigma2[l]; invSigma2[l] = f.mvInvLevelSigma2[l];
  }
}

// ORBmatcher matcher(nnratio, checkOri); matcher.SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize):
// src/Tracking.cc:915-926 - through the reference's own class declaration, body from orb_b200_matcher.cpp
int orbgpu_search_for_initialization(void* h1, void* h2, float* prevMatchedXY /* in/out, N1 x 2 */, int* matches12 /* N1 */,
                                     int windowSize, float nnratio, int checkOri) {
  try {
    GpuFrame *a = (GpuFrame*)h1, *b = (GpuFrame*)h2;
    std::vector<cv::Point2f> prev(a->f.N);
    for (int i = 0; i < a->f.N; i++) prev[i] = cv::Point2f(prevMatchedXY[2 * i], prevMatchedXY[2 * i + 1]);
    std::vector<int> m12;
    ORB_SLAM2::ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForInitialization(a->f, b->f, prev, m12, windowSize);
    for (int i = 0; i < a->f.N; i++) {
      prevMatchedXY[2 * i] = prev[i].x; prevMatchedXY[2 * i + 1] = prev[i].y;
      matches12[i] = m12[i];
    }
    return n;
  } catch (const std::exception& e) {
    t_error = e.what();
    return -1;
  }
}
int orbgpu_descriptor_distance(const unsigned char* a, const unsigned char* b) {
  cv::Mat ma(1, 32, CV_8UC1, (void*)a), mb(1, 32, CV_8UC1, (void*)b);
  return ORB_SLAM2::ORBmatcher::DescriptorDistance(ma, mb);
}

}  // extern "C"
