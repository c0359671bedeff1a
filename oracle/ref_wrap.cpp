// oracle/ref_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REFERENCE's own ORB_SLAM2::ORBextractor, compiled unmodified from
// /root/reference/src/ORBextractor.cc against the OpenCV stand-in in oracle/cvshim/ (see the
// header of cvshim.hpp for what that does and does not prove). Built into oracle/_ref/liborbref.so
// by `make -C oracle ref`; used by tests/test_oracle_vs_ref.py to pin the oracle restatement and
// by tools/make_golden.py to cross-check the golden vectors. Never loaded by the product.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include <atomic>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>

#include "cvshim.hpp"
// The wrapper (and only the wrapper - the reference's own translation units are compiled untouched) reaches
// Frame's private helpers UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid (Frame.h, `private:`)
// and ORBmatcher's protected ComputeThreeMaxima directly. Access control does not change layout or mangling.
#define private public
#define protected public
#include "ORBextractor.h"  // the reference's headers: -I/root/reference/include
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

// ---------------------------------------------------------------------------------------------------
// Canonical heap order. DistributeOctTree sorts (size, ExtractorNode*) pairs (ORBextractor.cc:926), so
// among nodes of equal size the one at the higher heap address is divided first - and since children are
// pushed to the list front in processing order, even the ORDER of the returned keypoints depends on heap
// addresses. With glibc malloc that order changes from process to process. The oracle and the CUDA path
// use the rule "addresses grow with creation order". To run the unmodified reference under exactly that
// rule, this library (linked with -Bsymbolic, so only code inside it is affected) replaces operator new
// by a per-extractor bump allocator that never reuses an address while `canonical` mode is on; with the
// mode off every allocation goes to malloc, i.e. the reference runs as it would in its own build.
namespace {
const size_t kArenaBytes = size_t(2) << 30;  // virtual; pages are touched lazily
struct Arena {
  char* base;
  size_t off;
};
const int kMaxArenas = 256;
std::atomic<char*> g_arena_base[kMaxArenas];
thread_local Arena* tl_arena = nullptr;

inline bool in_arena(void* p) {
  for (int i = 0; i < kMaxArenas; i++) {
    char* b = g_arena_base[i].load(std::memory_order_relaxed);
    if (b && (char*)p >= b && (char*)p < b + kArenaBytes) return true;
  }
  return false;
}
inline void* ref_alloc(size_t n) {
  if (tl_arena) {
    size_t o = (tl_arena->off + 15) & ~size_t(15);
    if (o + n > kArenaBytes) abort();
    tl_arena->off = o + n;
    return tl_arena->base + o;
  }
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
inline void ref_free(void* p) {
  if (p && !in_arena(p)) free(p);
}
}  // namespace
void* operator new(size_t n) { return ref_alloc(n); }
void* operator new[](size_t n) { return ref_alloc(n); }
void operator delete(void* p) noexcept { ref_free(p); }
void operator delete[](void* p) noexcept { ref_free(p); }
void operator delete(void* p, size_t) noexcept { ref_free(p); }
void operator delete[](void* p, size_t) noexcept { ref_free(p); }

extern "C" void* cvshim_primitive_enter() {
  void* t = tl_arena;
  tl_arena = nullptr;
  return t;
}
extern "C" void cvshim_primitive_leave(void* token) { tl_arena = (Arena*)token; }

namespace {
struct ArenaOwner {  // first member of Ref: constructed first, destroyed last
  Arena a{nullptr, 0};
  int slot = -1;
  void ensure() {
    if (a.base) return;
    void* m = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) abort();
    a.base = (char*)m;
    for (int i = 0; i < kMaxArenas; i++) {
      char* expect = nullptr;
      if (g_arena_base[i].compare_exchange_strong(expect, a.base)) { slot = i; return; }
    }
    abort();
  }
  ~ArenaOwner() {
    if (!a.base) return;
    g_arena_base[slot].store(nullptr);
    munmap(a.base, kArenaBytes);
  }
};

struct Ref {
  ArenaOwner arena;
  ORB_SLAM2::ORBextractor ex;
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  Ref(int n, float sf, int nl, int ini, int mn) : ex(n, sf, nl, ini, mn) {}
};
}  // namespace

extern "C" {

void* orbref_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new Ref(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbref_destroy(void* h) { delete (Ref*)h; }

// ORBextractor::operator()(image, cv::Mat(), keypoints, descriptors); keypoints in cv::KeyPoint's 28-byte layout
// canonical != 0: run under the monotonic allocator (see the top of this file)
int orbref_extract(void* h, const unsigned char* img, int w, int hgt, int step, void* kps, unsigned char* desc, int cap,
                   int canonical) {
  Ref* r = (Ref*)h;
  cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)step);
  cv::Mat mask;
  // drop everything a previous call may have left in the arena, then start it from its base again
  std::vector<cv::KeyPoint>().swap(r->kps);
  r->desc.release();
  for (size_t l = 0; l < r->ex.mvImagePyramid.size(); l++) r->ex.mvImagePyramid[l] = cv::Mat();
  r->arena.a.off = 0;
  if (canonical) {
    r->arena.ensure();
    tl_arena = &r->arena.a;
  }
  r->ex(image, mask, r->kps, r->desc);
  tl_arena = nullptr;
  int n = (int)r->kps.size();
  int m = n < cap ? n : cap;
  static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
  if (kps && m > 0) memcpy(kps, r->kps.data(), (size_t)m * sizeof(cv::KeyPoint));
  if (desc)
    for (int i = 0; i < m; i++) memcpy(desc + (size_t)i * 32, r->desc.ptr(i), 32);
  return n;
}

void orbref_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2) {
  Ref* r = (Ref*)h;
  std::vector<float> a = r->ex.GetScaleFactors(), b = r->ex.GetInverseScaleFactors(), c = r->ex.GetScaleSigmaSquares(),
                     d = r->ex.GetInverseScaleSigmaSquares();
  for (size_t i = 0; i < a.size(); i++) { scale[i] = a[i]; invScale[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = d[i]; }
}

// mvImagePyramid[l] of the last call: size, and a copy INCLUDING the 19-px border around the view
void orbref_level_info(void* h, int l, int* w, int* hgt) {
  Ref* r = (Ref*)h;
  *w = r->ex.mvImagePyramid[l].cols; *hgt = r->ex.mvImagePyramid[l].rows;
}
void orbref_level_copy(void* h, int l, unsigned char* dst /* (w+38) x (h+38) packed */) {
  Ref* r = (Ref*)h;
  const cv::Mat& m = r->ex.mvImagePyramid[l];
  const int E = 19, W = m.cols + 2 * E;
  for (int y = -E; y < m.rows + E; y++) memcpy(dst + (size_t)(y + E) * W, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - E, (size_t)W);
}

// one extractor per thread, frames handed out dynamically; returns the total keypoint count
long orbref_extract_batch_mt(int nfeatures, float sf, int nlevels, int ini, int mn, const unsigned char* imgs, int B, int w,
                             int hgt, int nthreads) {
  std::atomic<int> next(0);
  std::atomic<long> total(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&]() {
      Ref r(nfeatures, sf, nlevels, ini, mn);
      cv::Mat mask;
      for (int i = next++; i < B; i = next++) {
        cv::Mat image(hgt, w, CV_8UC1, (void*)(imgs + (size_t)i * w * hgt), (size_t)w);
        r.ex(image, mask, r.kps, r.desc);
        total += (long)r.kps.size();
      }
    });
  for (auto& x : th) x.join();
  return total.load();
}


// ---------------------------------------------------------------------------------------------------
// ORBmatcher.cc / Frame.cc of the reference, compiled unmodified (oracle/Makefile). Everything they call in
// KeyFrame / MapPoint / Map / DBoW2 is outside the pinned path and is satisfied by generated abort() stubs.

int orbref_descriptor_distance(const unsigned char* a, const unsigned char* b) {
  cv::Mat ma(1, 32, CV_8UC1, (void*)a), mb(1, 32, CV_8UC1, (void*)b);
  return ORB_SLAM2::ORBmatcher::DescriptorDistance(ma, mb);
}

int orbref_matcher_constants(int* th_low, int* th_high, int* histo_length) {
  *th_low = ORB_SLAM2::ORBmatcher::TH_LOW; *th_high = ORB_SLAM2::ORBmatcher::TH_HIGH;
  *histo_length = ORB_SLAM2::ORBmatcher::HISTO_LENGTH;
  return 0;
}

void orbref_three_maxima(const int* counts, int L, int* out3) {
  std::vector<std::vector<int>> histo(L);
  for (int i = 0; i < L; i++) histo[i].assign(counts[i], 0);
  ORB_SLAM2::ORBmatcher m(0.9f, true);
  out3[0] = out3[1] = out3[2] = -1;  // the function only assigns what it finds; every caller starts from -1 (:696)
  m.ComputeThreeMaxima(histo.data(), L, out3[0], out3[1], out3[2]);
}

}  // extern "C"

namespace {
struct RefFrame {
  ORB_SLAM2::Frame f;
  int w, h;
  // Frame's image bounds and grid cell sizes are static members (Frame.cc:45-48), set by the first constructed
  // frame (:176-196); the wrapper re-derives them for this frame's camera before every use with the same two
  // statements as the constructor.
  void activate() {
    f.ComputeImageBounds(cv::Mat(h, w, CV_8UC1));
    ORB_SLAM2::Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (ORB_SLAM2::Frame::mnMaxX - ORB_SLAM2::Frame::mnMinX);
    ORB_SLAM2::Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (ORB_SLAM2::Frame::mnMaxY - ORB_SLAM2::Frame::mnMinY);
  }
};
}  // namespace

extern "C" {

// A Frame from given keypoints / descriptors: default constructor, members filled in, then the reference's own
// UndistortKeyPoints (Frame.cc:724) and AssignFeaturesToGrid (:399). cam9 = fx fy cx cy k1 k2 p1 p2 k3.
void* orbref_frame_create(const void* kps, int n, const unsigned char* desc, const float* cam9, int w, int h) {
  RefFrame* r = new RefFrame();
  r->w = w; r->h = h;
  ORB_SLAM2::Frame& f = r->f;
  f.N = n;
  f.mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.mDescriptors = cv::Mat(n, 32, CV_8UC1);
  if (n) memcpy(f.mDescriptors.data, desc, (size_t)n * 32);
  f.mK = cv::Mat::eye(3, 3, CV_32F);
  f.mK.at<float>(0, 0) = cam9[0]; f.mK.at<float>(1, 1) = cam9[1]; f.mK.at<float>(0, 2) = cam9[2]; f.mK.at<float>(1, 2) = cam9[3];
  const int nd = cam9[8] != 0.0f ? 5 : 4;  // Tracking.cc: k3 is appended only when it is non-zero
  f.mDistCoef = cv::Mat(nd, 1, CV_32F);
  for (int i = 0; i < nd; i++) f.mDistCoef.at<float>(i) = cam9[4 + i];
  r->activate();
  f.UndistortKeyPoints();
  f.mvpMapPoints.assign(n, (ORB_SLAM2::MapPoint*)nullptr);
  f.mvbOutlier.assign(n, false);
  f.AssignFeaturesToGrid();
  return r;
}
void orbref_frame_destroy(void* h) { delete (RefFrame*)h; }

void orbref_frame_keys_un(void* h, void* kps_out) {
  RefFrame* r = (RefFrame*)h;
  if (r->f.N) memcpy(kps_out, r->f.mvKeysUn.data(), (size_t)r->f.N * sizeof(cv::KeyPoint));
}
void orbref_frame_bounds(void* h, float* b4) {
  RefFrame* r = (RefFrame*)h;
  r->activate();
  b4[0] = ORB_SLAM2::Frame::mnMinX; b4[1] = ORB_SLAM2::Frame::mnMaxX; b4[2] = ORB_SLAM2::Frame::mnMinY; b4[3] = ORB_SLAM2::Frame::mnMaxY;
}
// mGrid as CSR: cell = ix*48 + iy, items in the order AssignFeaturesToGrid pushed them
void orbref_frame_grid(void* h, int* cellStart /* 64*48+1 */, int* items /* N */) {
  RefFrame* r = (RefFrame*)h;
  int pos = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
      cellStart[ix * FRAME_GRID_ROWS + iy] = pos;
      for (size_t idx : r->f.mGrid[ix][iy]) items[pos++] = (int)idx;
    }
  cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
}
int orbref_frame_features_in_area(void* h, float x, float y, float rad, int minLevel, int maxLevel, int* out, int cap) {
  RefFrame* r = (RefFrame*)h;
  r->activate();
  std::vector<size_t> v = r->f.GetFeaturesInArea(x, y, rad, minLevel, maxLevel);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}

// ORBmatcher(nnratio, checkOri).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:573
int orbref_search_for_initialization(void* h1, void* h2, float* prevMatchedXY /* in/out, N1 x 2 */, int* matches12 /* N1 */,
                                     int windowSize, float nnratio, int checkOri) {
  RefFrame *a = (RefFrame*)h1, *b = (RefFrame*)h2;
  a->activate();
  std::vector<cv::Point2f> prev(a->f.N);
  for (int i = 0; i < a->f.N; i++) prev[i] = cv::Point2f(prevMatchedXY[2 * i], prevMatchedXY[2 * i + 1]);
  std::vector<int> m12;
  ORB_SLAM2::ORBmatcher matcher(nnratio, checkOri != 0);
  int n = matcher.SearchForInitialization(a->f, b->f, prev, m12, windowSize);
  for (int i = 0; i < a->f.N; i++) {
    prevMatchedXY[2 * i] = prev[i].x; prevMatchedXY[2 * i + 1] = prev[i].y;
    matches12[i] = m12[i];
  }
  return n;
}

// Frame::ComputeStereoMatches (Frame.cc:831) on the keypoints, descriptors and pyramids the two reference
// extractors hold from their last orbref_extract call. Returns the number of keypoints with a depth.
int orbref_stereo_matches(void* hL, void* hR, float mbf, float mb, float* uRight, float* depth) {
  Ref *L = (Ref*)hL, *R = (Ref*)hR;
  ORB_SLAM2::Frame f;
  f.mpORBextractorLeft = &L->ex; f.mpORBextractorRight = &R->ex;
  f.mvKeys = L->kps; f.mvKeysRight = R->kps;
  f.N = (int)f.mvKeys.size();
  f.mDescriptors = L->desc.clone(); f.mDescriptorsRight = R->desc.clone();
  f.mvScaleFactors = L->ex.GetScaleFactors(); f.mvInvScaleFactors = L->ex.GetInverseScaleFactors();
  f.mbf = mbf; f.mb = mb;
  f.ComputeStereoMatches();
  int kept = 0;
  for (int i = 0; i < f.N; i++) { uRight[i] = f.mvuRight[i]; depth[i] = f.mvDepth[i]; kept += f.mvDepth[i] > 0; }
  return kept;
}

}  // extern "C"
