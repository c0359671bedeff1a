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

#include "ORBextractor.h"  // the reference's header: -I/root/reference/include

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

}  // extern "C"
