// oracle/ref_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REFERENCE's own ORB_SLAM2::ORBextractor, ORBmatcher, Frame, MapPoint,
// KeyFrame and Map, compiled unmodified from /root/reference/src/*.cc against the OpenCV stand-in in
// oracle/cvshim/ (see the header of cvshim.hpp for what that does and does not prove). Built into
// oracle/_ref/liborbref.so by `make -C oracle ref`; used by tests/test_oracle_vs_ref*.py, the GPU tests
// named *_the_reference_itself, tools/ref_stress*.py and bench.py's CPU arm. Never loaded by the product.
#include <atomic>
#include <chrono>
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
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBmatcher.h"

// Wall time of the last call into the reference's class (the matcher / stereo member only, not the construction of the
// live objects around it): the same wrapper is linked into liborbref.so (CPU bodies) and liborbref_gpu.so (drop-in bodies),
// so tools/gpu_dropin_latency.py reads both sides of the comparison from one clock.
static double g_lastCallUs = 0.0;
struct CallTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~CallTimer() { g_lastCallUs = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
};
extern "C" double orbref_last_call_us() { return g_lastCallUs; }
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
std::atomic<int> g_arena_hi{0};   // slots >= g_arena_hi were never used: operator delete only scans the used ones
thread_local Arena* tl_arena = nullptr;

inline bool in_arena(void* p) {
  const int hi = g_arena_hi.load(std::memory_order_acquire);
  for (int i = 0; i < hi; i++) {
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
      if (g_arena_base[i].compare_exchange_strong(expect, a.base)) {
        slot = i;
        int hi = g_arena_hi.load();
        while (hi < i + 1 && !g_arena_hi.compare_exchange_weak(hi, i + 1)) {}
        return;
      }
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
  int n;
  { CallTimer t; n = matcher.SearchForInitialization(a->f, b->f, prev, m12, windowSize); }
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
  { CallTimer t; f.ComputeStereoMatches(); }
  int kept = 0;
  for (int i = 0; i < f.N; i++) { uRight[i] = f.mvuRight[i]; depth[i] = f.mvDepth[i]; kept += f.mvDepth[i] > 0; }
  return kept;
}


}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// The projection searches of Tracking (ORBmatcher.cc:72 local map, :1710 last frame) on live ORB_SLAM2::MapPoint
// objects (src/MapPoint.cc, src/Map.cc compiled unmodified). The wrapper builds the frames and map points from
// flat arrays in the layout the oracle's entry points use, runs the reference function and flattens
// CurrentFrame.mvpMapPoints back into indices.
namespace {
void set_tracking_state(RefFrame* cur, const float* uright, const unsigned char* occupied0, float mbf, float mb,
                        const float* scaleFactors, int nlevels, const float* cam4, ORB_SLAM2::Map* map, ORB_SLAM2::Frame& helper,
                        std::vector<ORB_SLAM2::MapPoint*>& owned) {
  ORB_SLAM2::Frame& f = cur->f;
  cur->activate();
  f.mvuRight.assign(f.N, -1.f);
  if (uright) f.mvuRight.assign(uright, uright + f.N);
  f.mbf = mbf; f.mb = mb;
  f.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
  f.mnScaleLevels = nlevels;
  ORB_SLAM2::Frame::fx = cam4[0]; ORB_SLAM2::Frame::fy = cam4[1]; ORB_SLAM2::Frame::cx = cam4[2]; ORB_SLAM2::Frame::cy = cam4[3];
  ORB_SLAM2::Frame::invfx = 1.0f / cam4[0]; ORB_SLAM2::Frame::invfy = 1.0f / cam4[1];
  // keypoints that already hold an observed map point: one shared MapPoint with Observations() > 0
  helper.N = 1;
  helper.mvKeysUn.assign(1, cv::KeyPoint());
  helper.mvScaleFactors = f.mvScaleFactors; helper.mnScaleLevels = nlevels;
  helper.mDescriptors = cv::Mat::zeros(1, 32, CV_8UC1);
  helper.SetPose(cv::Mat::eye(4, 4, CV_32F));
  cv::Mat pos = (cv::Mat_<float>(3, 1) << 0.f, 0.f, 1.f);
  ORB_SLAM2::MapPoint* taken = new ORB_SLAM2::MapPoint(pos, map, &helper, 0);
  taken->nObs = 1;
  owned.push_back(taken);
  f.mvpMapPoints.assign(f.N, (ORB_SLAM2::MapPoint*)nullptr);
  for (int i = 0; i < f.N; i++)
    if (occupied0 && occupied0[i]) f.mvpMapPoints[i] = taken;
}
void flatten(const ORB_SLAM2::Frame& f, const std::vector<ORB_SLAM2::MapPoint*>& queries, int* matchOfKp) {
  std::map<ORB_SLAM2::MapPoint*, int> index;
  for (size_t i = 0; i < queries.size(); i++)
    if (queries[i]) index[queries[i]] = (int)i;
  for (int i = 0; i < f.N; i++) {
    auto it = index.find(f.mvpMapPoints[i]);
    matchOfKp[i] = it == index.end() ? -1 : it->second;
  }
}
}  // namespace

extern "C" {

// ORBmatcher(0.9, true).SearchByProjection(CurrentFrame, LastFrame, th, bMono=false), ORBmatcher.cc:1710.
// mpFlags bit0: LastFrame.mvpMapPoints[i] exists and is no outlier; bit1: its Observations() > 0.
// direction 0 / 1 / 2 = neither / bForward / bBackward: realised through LastFrame's pose (tlc.z = 0, +3 mb, -3 mb).
int orbref_search_last_frame(void* hCur, const float* uright, const unsigned char* occupied0, const void* lastKps, int n1,
                             const float* Xw, const unsigned char* mpFlags, const unsigned char* mpDesc, const float* Tcw16,
                             const float* cam4, float mbf, float mb, float th, int direction, const float* scaleFactors, int nlevels,
                             int* matchOfKp) {
  RefFrame* cur = (RefFrame*)hCur;
  ORB_SLAM2::Map map;
  ORB_SLAM2::Frame helper;
  std::vector<ORB_SLAM2::MapPoint*> owned;
  set_tracking_state(cur, uright, occupied0, mbf, mb, scaleFactors, nlevels, cam4, &map, helper, owned);
  cv::Mat Tcw(4, 4, CV_32F);
  memcpy(Tcw.data, Tcw16, 16 * sizeof(float));
  cur->f.SetPose(Tcw);
  ORB_SLAM2::Frame last;
  last.N = n1;
  last.mvKeys.assign((const cv::KeyPoint*)lastKps, (const cv::KeyPoint*)lastKps + n1);
  last.mvKeysUn = last.mvKeys;
  last.mDescriptors = cv::Mat(n1, 32, CV_8UC1);
  if (n1) memcpy(last.mDescriptors.data, mpDesc, (size_t)n1 * 32);
  last.mvScaleFactors = cur->f.mvScaleFactors; last.mnScaleLevels = nlevels;
  // Rlw = I, tlw = d - twc  =>  tlc = Rlw*twc + tlw = d
  cv::Mat twc = cur->f.GetCameraCenter();
  cv::Mat Tlw = cv::Mat::eye(4, 4, CV_32F);
  const float dz = direction == 1 ? 3.f * mb : (direction == 2 ? -3.f * mb : 0.f);
  Tlw.at<float>(0, 3) = -twc.at<float>(0); Tlw.at<float>(1, 3) = -twc.at<float>(1); Tlw.at<float>(2, 3) = dz - twc.at<float>(2);
  last.SetPose(Tlw);
  last.mvpMapPoints.assign(n1, (ORB_SLAM2::MapPoint*)nullptr);
  last.mvbOutlier.assign(n1, false);
  for (int i = 0; i < n1; i++) {
    if (!(mpFlags[i] & 1)) continue;
    cv::Mat pos = (cv::Mat_<float>(3, 1) << Xw[3 * i], Xw[3 * i + 1], Xw[3 * i + 2]);
    ORB_SLAM2::MapPoint* mp = new ORB_SLAM2::MapPoint(pos, &map, &last, i);
    mp->nObs = (mpFlags[i] & 2) ? 1 : 0;
    last.mvpMapPoints[i] = mp;
    owned.push_back(mp);
  }
  ORB_SLAM2::ORBmatcher matcher(0.9f, true);
  int n;
  { CallTimer t; n = matcher.SearchByProjection(cur->f, last, th, false); }
  flatten(cur->f, last.mvpMapPoints, matchOfKp);
  for (ORB_SLAM2::MapPoint* p : owned) delete p;
  cur->f.mvpMapPoints.assign(cur->f.N, (ORB_SLAM2::MapPoint*)nullptr);
  return n;
}

// ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th), ORBmatcher.cc:72, on map points whose tracking fields
// (the output of Frame::isInFrustum) are given.
int orbref_search_local_map(void* hCur, const float* uright, const unsigned char* occupied0, int nmp, const unsigned char* trackInView,
                            const unsigned char* bad, const int* level, const float* viewCos, const float* projX, const float* projY,
                            const float* projXR, const unsigned char* desc, const int* nobs, float th, float nnratio,
                            const float* cam4, float mbf, float mb, const float* scaleFactors, int nlevels, int* matchOfKp) {
  RefFrame* cur = (RefFrame*)hCur;
  ORB_SLAM2::Map map;
  ORB_SLAM2::Frame helper, src;
  std::vector<ORB_SLAM2::MapPoint*> owned;
  set_tracking_state(cur, uright, occupied0, mbf, mb, scaleFactors, nlevels, cam4, &map, helper, owned);
  src.N = nmp;
  src.mvKeysUn.assign(nmp, cv::KeyPoint());
  src.mvScaleFactors = cur->f.mvScaleFactors; src.mnScaleLevels = nlevels;
  src.mDescriptors = cv::Mat(nmp, 32, CV_8UC1);
  if (nmp) memcpy(src.mDescriptors.data, desc, (size_t)nmp * 32);
  src.SetPose(cv::Mat::eye(4, 4, CV_32F));
  std::vector<ORB_SLAM2::MapPoint*> mps(nmp);
  cv::Mat pos = (cv::Mat_<float>(3, 1) << 0.f, 0.f, 1.f);
  for (int i = 0; i < nmp; i++) {
    ORB_SLAM2::MapPoint* mp = new ORB_SLAM2::MapPoint(pos, &map, &src, i);  // mDescriptor = desc row i (MapPoint.cc:107)
    mp->mbTrackInView = trackInView[i] != 0;
    mp->mbBad = bad[i] != 0;
    mp->mnTrackScaleLevel = level[i];
    mp->mTrackViewCos = viewCos[i];
    mp->mTrackProjX = projX[i]; mp->mTrackProjY = projY[i]; mp->mTrackProjXR = projXR[i];
    mp->nObs = nobs[i];
    mps[i] = mp;
    owned.push_back(mp);
  }
  ORB_SLAM2::ORBmatcher matcher(nnratio, true);
  int n;
  { CallTimer t; n = matcher.SearchByProjection(cur->f, mps, th); }
  flatten(cur->f, mps, matchOfKp);
  for (ORB_SLAM2::MapPoint* p : owned) delete p;
  cur->f.mvpMapPoints.assign(cur->f.N, (ORB_SLAM2::MapPoint*)nullptr);
  return n;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// The bag-of-words guided searches (ORBmatcher.cc:247 keyframe -> frame, :729 keyframe -> keyframe, :884
// triangulation) on live ORB_SLAM2::KeyFrame objects (src/KeyFrame.cc compiled unmodified). A DBoW2::FeatureVector
// is given as the vocabulary node id of every feature (-1 = none), the layout the oracle and the device use.
namespace {
struct Side {
  ORB_SLAM2::Frame f;
  ORB_SLAM2::KeyFrame* kf = nullptr;
  std::vector<ORB_SLAM2::MapPoint*> mps;
  ~Side() {
    delete kf;
    for (ORB_SLAM2::MapPoint* p : mps) delete p;
  }
  // hasMp[i]: feature i holds a (good) map point
  void build(const void* kps, int n, const unsigned char* desc, const int* node, const unsigned char* hasMp, const float* uright,
             const float* Tcw16, const float* cam4, const float* scaleFactors, const float* levelSigma2, int nlevels,
             ORB_SLAM2::Map* map, bool asKeyFrame) {
    f.N = n;
    f.mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
    f.mvKeysUn = f.mvKeys;
    f.mDescriptors = cv::Mat(n, 32, CV_8UC1);
    if (n) memcpy(f.mDescriptors.data, desc, (size_t)n * 32);
    for (int i = 0; i < n; i++)
      if (node[i] >= 0) f.mFeatVec.addFeature((DBoW2::NodeId)node[i], (unsigned int)i);
    f.mvuRight.assign(n, -1.f);
    if (uright) f.mvuRight.assign(uright, uright + n);
    f.mvDepth.assign(n, -1.f);
    f.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    f.mvLevelSigma2.assign(levelSigma2, levelSigma2 + nlevels);
    f.mnScaleLevels = nlevels;
    f.mb = 0.5f; f.mbf = 40.f;
    ORB_SLAM2::Frame::fx = cam4[0]; ORB_SLAM2::Frame::fy = cam4[1]; ORB_SLAM2::Frame::cx = cam4[2]; ORB_SLAM2::Frame::cy = cam4[3];
    ORB_SLAM2::Frame::invfx = 1.0f / cam4[0]; ORB_SLAM2::Frame::invfy = 1.0f / cam4[1];
    cv::Mat Tcw = cv::Mat::eye(4, 4, CV_32F);
    if (Tcw16) memcpy(Tcw.data, Tcw16, 16 * sizeof(float));
    f.SetPose(Tcw);
    f.mvpMapPoints.assign(n, (ORB_SLAM2::MapPoint*)nullptr);
    f.mvbOutlier.assign(n, false);
    cv::Mat pos = (cv::Mat_<float>(3, 1) << 0.f, 0.f, 1.f);
    for (int i = 0; i < n; i++)
      if (hasMp && hasMp[i]) {
        ORB_SLAM2::MapPoint* mp = new ORB_SLAM2::MapPoint(pos, map, &f, i);
        mps.push_back(mp);
        f.mvpMapPoints[i] = mp;
      }
    if (asKeyFrame) kf = new ORB_SLAM2::KeyFrame(f, map, (ORB_SLAM2::KeyFrameDatabase*)nullptr);
  }
  int index_of(ORB_SLAM2::MapPoint* p) const {
    if (!p) return -1;
    for (int i = 0; i < f.N; i++)
      if (f.mvpMapPoints[i] == p) return i;
    return -1;
  }
};
const float kIdentityCam[4] = {500.f, 500.f, 320.f, 240.f};
}  // namespace

extern "C" {

// ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches), ORBmatcher.cc:247.
// matchOfKp[j] = keyframe feature whose map point frame feature j received (-1 = none).
int orbref_search_by_bow_frame(const void* kps1, int n1, const unsigned char* desc1, const int* node1, const unsigned char* usable1,
                               const void* kps2, int n2, const unsigned char* desc2, const int* node2, float nnratio, int checkOri,
                               const float* scaleFactors, int nlevels, int* matchOfKp) {
  ORB_SLAM2::Map map;
  Side a, b;
  std::vector<float> s2(scaleFactors, scaleFactors + nlevels);
  a.build(kps1, n1, desc1, node1, usable1, nullptr, nullptr, kIdentityCam, scaleFactors, s2.data(), nlevels, &map, true);
  b.build(kps2, n2, desc2, node2, nullptr, nullptr, nullptr, kIdentityCam, scaleFactors, s2.data(), nlevels, &map, false);
  std::vector<ORB_SLAM2::MapPoint*> matches;
  ORB_SLAM2::ORBmatcher matcher(nnratio, checkOri != 0);
  int n;
  { CallTimer t; n = matcher.SearchByBoW(a.kf, b.f, matches); }
  for (int j = 0; j < n2; j++) matchOfKp[j] = a.index_of(matches[j]);
  return n;
}

// ORBmatcher(nnratio, checkOri).SearchByBoW(pKF1, pKF2, vpMatches12), ORBmatcher.cc:729.
// matches12[i] = keyframe-2 feature whose map point keyframe-1 feature i was matched to (-1 = none).
int orbref_search_by_bow_keyframes(const void* kps1, int n1, const unsigned char* desc1, const int* node1, const unsigned char* usable1,
                                   const void* kps2, int n2, const unsigned char* desc2, const int* node2, const unsigned char* usable2,
                                   float nnratio, int checkOri, const float* scaleFactors, int nlevels, int* matches12) {
  ORB_SLAM2::Map map;
  Side a, b;
  std::vector<float> s2(scaleFactors, scaleFactors + nlevels);
  a.build(kps1, n1, desc1, node1, usable1, nullptr, nullptr, kIdentityCam, scaleFactors, s2.data(), nlevels, &map, true);
  b.build(kps2, n2, desc2, node2, usable2, nullptr, nullptr, kIdentityCam, scaleFactors, s2.data(), nlevels, &map, true);
  std::vector<ORB_SLAM2::MapPoint*> matches;
  ORB_SLAM2::ORBmatcher matcher(nnratio, checkOri != 0);
  int n;
  { CallTimer t; n = matcher.SearchByBoW(a.kf, b.kf, matches); }
  for (int i = 0; i < n1; i++) matches12[i] = b.index_of(matches[i]);
  return n;
}

// ORBmatcher(0.6, checkOri).SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo), ORBmatcher.cc:884.
// Keyframe 1 sits at the identity pose, keyframe 2 at Tcw2 (the epipole :898-908 follows from the two poses).
int orbref_search_for_triangulation(const void* kps1, int n1, const unsigned char* desc1, const int* node1, const unsigned char* hasMp1,
                                    const float* ur1, const void* kps2, int n2, const unsigned char* desc2, const int* node2,
                                    const unsigned char* hasMp2, const float* ur2, const float* Tcw2, const float* cam4, const float* F12,
                                    int onlyStereo, int checkOri, const float* scaleFactors, const float* levelSigma2, int nlevels,
                                    int* matches12) {
  ORB_SLAM2::Map map;
  Side a, b;
  a.build(kps1, n1, desc1, node1, hasMp1, ur1, nullptr, cam4, scaleFactors, levelSigma2, nlevels, &map, true);
  b.build(kps2, n2, desc2, node2, hasMp2, ur2, Tcw2, cam4, scaleFactors, levelSigma2, nlevels, &map, true);
  cv::Mat F(3, 3, CV_32F);
  memcpy(F.data, F12, 9 * sizeof(float));
  std::vector<std::pair<size_t, size_t>> pairs;
  ORB_SLAM2::ORBmatcher matcher(0.6f, checkOri != 0);
  int n;
  { CallTimer t; n = matcher.SearchForTriangulation(a.kf, b.kf, F, pairs, onlyStereo != 0); }
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  for (const auto& p : pairs) matches12[p.first] = (int)p.second;
  return n;
}

}  // extern "C"

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:365) on live objects: map point p is observed by keyframes
// 0 .. n_p-1 at feature index p, keyframe k holding observation k of every map point. mObservations is a
// std::map keyed by KeyFrame* (visited in address order), so keyframe k is the k-th lowest address.
// bestDesc[p] = the 32 bytes mDescriptor holds afterwards.
extern "C" void orbref_distinctive_descriptors(const unsigned char* desc, const int* offsets, int nPoints, unsigned char* bestDesc) {
  ORB_SLAM2::Map map;
  int maxObs = 0;
  for (int p = 0; p < nPoints; p++) maxObs = std::max(maxObs, offsets[p + 1] - offsets[p]);
  std::vector<Side*> sides(maxObs);
  std::vector<cv::KeyPoint> kps(nPoints);
  std::vector<int> node(nPoints, -1);
  const float sf[1] = {1.f};
  for (int k = 0; k < maxObs; k++) {
    std::vector<unsigned char> d((size_t)nPoints * 32, 0);
    for (int p = 0; p < nPoints; p++)
      if (k < offsets[p + 1] - offsets[p]) memcpy(&d[(size_t)p * 32], desc + (size_t)(offsets[p] + k) * 32, 32);
    sides[k] = new Side();
    sides[k]->build(kps.data(), nPoints, d.data(), node.data(), nullptr, nullptr, nullptr, kIdentityCam, sf, sf, 1, &map, true);
  }
  std::vector<ORB_SLAM2::KeyFrame*> kfs(maxObs);
  for (int k = 0; k < maxObs; k++) kfs[k] = sides[k]->kf;
  std::vector<ORB_SLAM2::KeyFrame*> byAddress = kfs;
  std::sort(byAddress.begin(), byAddress.end());
  // observation k must live in the k-th lowest keyframe: permute the descriptor rows accordingly
  for (int k = 0; k < maxObs; k++) {
    ORB_SLAM2::KeyFrame* target = byAddress[k];
    for (int p = 0; p < nPoints; p++)
      if (k < offsets[p + 1] - offsets[p]) memcpy(target->mDescriptors.data + (size_t)p * target->mDescriptors.step, desc + (size_t)(offsets[p] + k) * 32, 32);
  }
  cv::Mat pos = (cv::Mat_<float>(3, 1) << 0.f, 0.f, 1.f);
  for (int p = 0; p < nPoints; p++) {
    ORB_SLAM2::MapPoint mp(pos, byAddress.empty() ? nullptr : byAddress[0], &map);
    const int n = offsets[p + 1] - offsets[p];
    for (int k = 0; k < n; k++) mp.AddObservation(byAddress[k], (size_t)p);
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    if (!d.empty()) memcpy(bestDesc + (size_t)p * 32, d.data, 32);
    else memset(bestDesc + (size_t)p * 32, 0, 32);
  }
  for (Side* s_ : sides) delete s_;
}

