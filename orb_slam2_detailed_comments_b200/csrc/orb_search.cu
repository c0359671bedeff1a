// orb_search.cu — ordered candidate-set matchers of ORBmatcher on B200 (sm_100a)
// (reference: src/ORBmatcher.cc:72-169 SearchByProjection local map, :1710-1860 SearchByProjection last frame,
//  :247-420 SearchByBoW keyframe -> frame; candidates from src/Frame.cc:590-670 GetFeaturesInArea).
//
// The reference walks the map points in order; each one takes the best candidate keypoint that no earlier
// map point occupies. That chain is split in two:
//   k_search_scan    EIGHT lanes per query (windows hold ~5-30 candidates), all queries of all frames in
//                    parallel: the candidates are enumerated in the reference's visiting order (lanes =
//                    candidates), filtered against the frame's initial occupancy / stereo coordinate, and
//                    the FOUR smallest keys (distance << 22 | visiting rank) are kept (redux.sync min on the
//                    8-lane group; the rank inside the key names the owner lane).
//   k_search_commit  one CTA per frame solves the in-order dependency as a fixed point: the picks of all
//                    queries are re-chosen in parallel sweeps against "tag[idx] = smallest occupying query
//                    that picks idx" until a sweep changes nothing (triangular system: the fixed point is the
//                    sequential result; 1 + longest displacement chain sweeps). A query that lost stored
//                    candidates falls back on its 3rd / 4th, and only if those run out is it re-scored.
// The two smallest keys by (distance, visiting order) are exactly the reference's (best, second best):
// `dist < bestDist` / `else if (dist < bestDist2)` with demotion keeps the lexicographic top two.
// Integer-pipe work (LOP3 / POPC / REDUX); no tensor cores by design.
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

#include "orb_common.cuh"

using namespace orbb200;

namespace {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;   // Frame.h:41-42
constexpr int HISTO_LENGTH = 30;                                                      // ORBmatcher.cc:51
constexpr unsigned kNoKey = 0xffffffffu, kFull = 0xffffffffu;
constexpr int kScanWarps = 8;
constexpr int kMaxCols = 64;
constexpr int kMaxLevels = 16;

struct Top4 {
  unsigned k[4];
  int i[4];
};

__device__ __forceinline__ void top4_clear(Top4& T) {
#pragma unroll
  for (int m = 0; m < 4; m++) { T.k[m] = kNoKey; T.i[m] = -1; }
}

__device__ __forceinline__ void top4_insert(Top4& T, unsigned key, int idx) {
#pragma unroll
  for (int m = 0; m < 4; m++) {
    if (key < T.k[m]) {
      const unsigned tk = T.k[m]; const int ti = T.i[m];
      T.k[m] = key; T.i[m] = idx;
      key = tk; idx = ti;
    }
  }
}

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const u8* d) {
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(d)), b1 = __ldg(reinterpret_cast<const uint4*>(d) + 1);
  const unsigned x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
  const unsigned x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
  // carry-save adders: 4 POPC instead of 8
  const unsigned s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
  const unsigned s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
  const unsigned ones = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
  const unsigned twos = c1 ^ c2 ^ c3, fours = (c1 & c2) | (c3 & (c1 ^ c2));
  return __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
}

// Candidate words: bits 0-15 keypoint index, bits 16-23 its octave, bit 30 (first word only) "the query occupies".
constexpr unsigned kWordIdx = 0xffffu, kWordOccupies = 1u << 30;
__device__ __forceinline__ int word_idx(int w) { return w & (int)kWordIdx; }
__device__ __forceinline__ int word_oct(int w) { return (w >> 16) & 0xff; }

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return G == 32 ? kFull : (((1u << G) - 1u) << (lane & ~(G - 1)));
}

// The four smallest (distance, visiting rank) keys among candidates t = 0..total-1, G lanes per query;
// wordOf(t) returns the candidate word or -1 when the candidate is filtered out before the occupancy test.
// REV: ties go to the LATER candidate (SearchForTriangulation's `dist > bestDist` test), so the rank is stored reversed.
template <int G, bool REV = false, class WordFn, class OccFn>
__device__ __forceinline__ void group_top4(int total, WordFn wordOf, OccFn occ, const uint4 q0, const uint4 q1, const u8* desc,
                                           const float* uright, float ur, float radius, int lane, Top4& T) {
  const unsigned gm = group_mask<G>(lane);
  const int gl = lane & (G - 1);
  top4_clear(T);
  for (int t0 = 0; t0 < total; t0 += G) {
    const int t = t0 + gl;
    unsigned key = kNoKey;
    int w = -1;
    if (t < total) {
      w = wordOf(t);
      if (w >= 0 && !occ(word_idx(w))) {
        bool ok = true;
        if (uright) {   // ORBmatcher.cc:125-130 / :1792-1798
          const float r = uright[word_idx(w)];
          if (r > 0.f && fabsf(__fsub_rn(ur, r)) > radius) ok = false;
        }
        if (ok) key = ((unsigned)hamming256(q0, q1, desc + (size_t)word_idx(w) * 32) << 22) | (unsigned)(REV ? 0x3fffff - t : t);
      }
    }
    for (;;) {
      const unsigned m = __reduce_min_sync(gm, key);
      if (m >= T.k[3]) break;
      const int rk = (int)(m & 0x3fffffu);
      const int mw = __shfl_sync(gm, w, (REV ? 0x3fffff - rk : rk) - t0, G);   // the rank in the key names the owner lane
      top4_insert(T, m, mw);
      if (key == m) key = kNoKey;
    }
  }
}

// Frame::GetFeaturesInArea (Frame.cc:590-670) as a flattened candidate range: the cells (ix, cy0..cy1) of one
// grid column are contiguous in the CSR layout, so the window is <= 64 segments visited in ix order.
struct WindowSegs {
  int seg[kMaxCols];       // first CSR slot of column c
  int cum[kMaxCols + 1];   // candidates before column c
};

template <int G>
__device__ __forceinline__ int window_setup(const orb_proj_query& Q, const float* B4, float invW, float invH, const int* start,
                                            WindowSegs& W, int lane) {
  const unsigned gm = group_mask<G>(lane);
  const int gl = lane & (G - 1);
  const float x = Q.u, y = Q.v, r = Q.radius;
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, B4[0]), r), invW)));
  const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, B4[0]), r), invW)));
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, B4[2]), r), invH)));
  const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, B4[2]), r), invH)));
  if (!(cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0) || cx1 < cx0 || cy1 < cy0) return 0;
  const int ncols = cx1 - cx0 + 1;
  int run = 0;
  for (int c0 = 0; c0 < ncols; c0 += G) {
    const int c = c0 + gl;
    int b = 0, len = 0;
    if (c < ncols) {
      b = start[(cx0 + c) * kGridRows + cy0];
      len = start[(cx0 + c) * kGridRows + cy1 + 1] - b;
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      const int t = __shfl_up_sync(gm, incl, o, G);
      if (gl >= o) incl += t;
    }
    if (c < ncols) { W.seg[c] = b; W.cum[c] = run + incl - len; }
    run += __shfl_sync(gm, incl, G - 1, G);
  }
  if (gl == 0) W.cum[ncols] = run;
  __syncwarp(gm);
  return run;
}

// candidate t of the window -> candidate word, or -1 if the level / circle test rejects it
__device__ __forceinline__ int window_candidate(int t, const WindowSegs& W, const int* items, const orb_keypoint* K,
                                                const orb_proj_query& Q, float r2) {
  int c = 0;
  while (t >= W.cum[c + 1]) c++;
  const int idx = items[W.seg[c] + (t - W.cum[c])];
  const float kx = K[idx].x, ky = K[idx].y;
  const int oct = K[idx].octave;
  if (Q.min_level > 0 || Q.max_level >= 0) {
    if (oct < Q.min_level) return -1;
    if (Q.max_level >= 0 && oct > Q.max_level) return -1;
  }
  const float dx = __fsub_rn(kx, Q.u), dy = __fsub_rn(ky, Q.v);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2 ? (idx | ((oct & 0xff) << 16)) : -1;
}

struct SearchArgs {
  // current frames
  const orb_keypoint* kps; const u8* desc; const float* uright; const u8* occupied; const int* counts;
  const int* cellStart; const int* cellItems;
  float b4[4], invW, invH;
  int cap;
  // queries
  const orb_proj_query* queries; const u8* qdesc; const int* qcounts; int qcap;
  // BoW source (bow != 0): queries are keyframe features in (node, index) order
  int bow;
  const orb_keypoint* kps1; const u8* usable1;
  // SearchForTriangulation (tri != null, on top of the BoW source): usable1 = "has a map point" (inverted), epipolar filters
  const orb_triangulation_pair* tri; const float* ur1;
  float sf[kMaxLevels], sigma2[kMaxLevels];
  // scratch
  uint4* top; int* order1; int* cb; int* ce; int* sorted2; int* node2s; int* mqP; int* meta;
  // params / outputs
  int mode, th, checkOri; float ratio;
  int* matchOfKp; int* matchOfQuery; int* nmatches;
};

// ORBmatcher::CheckDistEpipolarLine (ORBmatcher.cc:205-227) and the epipole test of SearchForTriangulation (:958-965)
__device__ __forceinline__ bool tri_candidate_ok(const SearchArgs& A, const orb_triangulation_pair& P, float x1, float y1, bool stereo1,
                                                 const orb_keypoint& kp2, bool stereo2) {
  if (P.only_stereo && !stereo2) return false;
  const int oct = min(max(kp2.octave, 0), kMaxLevels - 1);
  if (!stereo1 && !stereo2) {
    const float dx = __fsub_rn(P.ex, kp2.x), dy = __fsub_rn(P.ey, kp2.y);
    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, A.sf[oct])) return false;
  }
  const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, P.F12[0]), __fmul_rn(y1, P.F12[3])), P.F12[6]);
  const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, P.F12[1]), __fmul_rn(y1, P.F12[4])), P.F12[7]);
  const float c = __fadd_rn(__fadd_rn(__fmul_rn(x1, P.F12[2]), __fmul_rn(y1, P.F12[5])), P.F12[8]);
  const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp2.x), __fmul_rn(b, kp2.y)), c);
  const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
  if (den == 0.f) return false;
  const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
  return (double)dsqr < __dmul_rn(3.84, (double)A.sigma2[oct]);
}

// One query of the BoW-ordered searches: keyframe feature at position q of the (node, index) order against the
// frame features of the same vocabulary node.
template <int G, class OccFn>
__device__ __forceinline__ void bow_score(const SearchArgs& A, int b, int q, OccFn occ, int lane, Top4& T) {
  top4_clear(T);
  const int i1 = A.order1[(size_t)b * A.qcap + q];
  const int c0 = A.cb[(size_t)b * A.qcap + q], c1 = A.ce[(size_t)b * A.qcap + q];
  const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + i1) * 32);
  const int* s2 = A.sorted2 + (size_t)b * A.cap;
  const u8* D = A.desc + (size_t)b * A.cap * 32;
  const bool flag = A.usable1[(size_t)b * A.qcap + i1] != 0;
  if (!A.tri) {
    if (!flag) return;
    const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
    group_top4<G>(c1 - c0, [&](int t) { return s2[c0 + t]; }, occ, q0, q1, D, nullptr, 0.f, 0.f, lane, T);
  } else {
    const orb_triangulation_pair P = A.tri[b];
    const bool stereo1 = A.ur1 && A.ur1[(size_t)b * A.qcap + i1] >= 0.f;
    if (flag || (P.only_stereo && !stereo1)) return;      // :927-936: has a map point already / monocular feature
    const orb_keypoint kp1 = A.kps1[(size_t)b * A.qcap + i1];
    const orb_keypoint* K2 = A.kps + (size_t)b * A.cap;
    const float* UR2 = A.uright ? A.uright + (size_t)b * A.cap : nullptr;
    const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
    group_top4<G, true>(c1 - c0,
                        [&](int t) {
                          const int i2 = s2[c0 + t];
                          return tri_candidate_ok(A, P, kp1.x, kp1.y, stereo1, K2[i2], UR2 && UR2[i2] >= 0.f) ? i2 : -1;
                        },
                        occ, q0, q1, D, nullptr, 0.f, 0.f, lane, T);
  }
}

// ---- pass 1: score every query against the frame's initial state, kScanGroup lanes per query
constexpr int kScanGroup = 8;
constexpr int kScanQueries = 32 * kScanWarps / kScanGroup;   // queries per CTA

__global__ void __launch_bounds__(32 * kScanWarps) k_search_scan(const SearchArgs A) {
  constexpr int G = kScanGroup;
  __shared__ WindowSegs segs[kScanQueries];
  const int b = blockIdx.y, lane = threadIdx.x & 31, grp = threadIdx.x / G, gl = lane & (G - 1);
  const int q = blockIdx.x * kScanQueries + grp;
  const int nq = A.bow ? A.meta[4 * b] : A.qcounts[b];
  if (q >= nq) return;
  const orb_keypoint* K = A.kps + (size_t)b * A.cap;
  const u8* D = A.desc + (size_t)b * A.cap * 32;
  const u8* occ0 = A.occupied ? A.occupied + (size_t)b * A.cap : nullptr;
  Top4 T;
  top4_clear(T);
  unsigned occupies = kWordOccupies;
  if (!A.bow) {
    const orb_proj_query Q = A.queries[(size_t)b * A.qcap + q];
    if (!(Q.flags & 2)) occupies = 0;
    if (Q.flags & 1) {
      WindowSegs& W = segs[grp];
      const int total = window_setup<G>(Q, A.b4, A.invW, A.invH, A.cellStart + (size_t)b * (kGridCells + 1), W, lane);
      const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + q) * 32);
      const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
      const int* items = A.cellItems + (size_t)b * A.cap;
      const float r2 = __fmul_rn(Q.radius, Q.radius);
      group_top4<G>(total, [&](int t) { return window_candidate(t, W, items, K, Q, r2); },
                    [&](int idx) { return occ0 && occ0[idx]; }, q0, q1, D,
                    A.uright ? A.uright + (size_t)b * A.cap : nullptr, Q.ur, Q.radius, lane, T);
    }
  } else {
    bow_score<G>(A, b, q, [&](int idx) { return occ0 && occ0[idx]; }, lane, T);
  }
  if (gl == 0) {
    uint4* o = A.top + ((size_t)b * A.qcap + q) * 2;
    o[0] = make_uint4(T.k[0], T.k[1], T.k[2], T.k[3]);
    o[1] = make_uint4((unsigned)T.i[0] | (T.k[0] != kNoKey ? occupies : 0u), (unsigned)T.i[1], (unsigned)T.i[2], (unsigned)T.i[3]);
  }
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2035-2077)
__device__ void three_maxima30(const int* histo, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < HISTO_LENGTH; i++) {
    const int s = histo[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
  else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {   // ORBmatcher.cc:1814-1820
  float rot = __fsub_rn(a1, a2);
  if (rot < 0.f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, HISTO_LENGTH / 360.0f));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// acceptance of (best, second) — ORBmatcher.cc:155-158 / :331-336 / :1807; wb / ws = candidate words or -1
__device__ __forceinline__ bool accept_match(int mode, int th, float ratio, unsigned kb, int wb, unsigned ks, int ws) {
  if (wb < 0) return false;
  const int bd = (int)(kb >> 22);
  if (bd > th) return false;
  const int bd2 = ws >= 0 ? (int)(ks >> 22) : 256;
  if (mode == ORB_SEARCH_RATIO_LEVEL) {
    const int l1 = word_oct(wb), l2 = ws >= 0 ? word_oct(ws) : -1;
    if (l1 == l2 && (float)bd > __fmul_rn(ratio, (float)bd2)) return false;
  } else if (mode == ORB_SEARCH_RATIO) {
    if (!((float)bd < __fmul_rn(ratio, (float)bd2))) return false;
  }
  return true;
}

// ---- pass 2: the in-order dependency as a fixed point, one CTA per frame.
// The walk of the reference is the unique solution of  pick(q) = choose(q, {pick(j) : j < q})  where choose takes the
// first (two) stored candidates that no earlier occupying query picked and applies the acceptance rule. The system is
// triangular, so Jacobi iteration converges to exactly that solution: every sweep publishes the current picks
// (tag[idx] = smallest occupying query that picks idx), then all queries re-choose in parallel against the tags
// (a candidate is taken for q iff tag[idx] < q). After sweep t the first t queries are final; the loop ends when a
// sweep changes nothing, which takes 1 + (longest chain of queries that actually displace each other) sweeps — a
// handful on tracking-like input. A query whose four stored candidates ran out is re-scored exactly (warp per
// query) against the same tags.
constexpr int kCommitThreads = 512;

__global__ void __launch_bounds__(kCommitThreads) k_search_commit(const SearchArgs A) {
  extern __shared__ __align__(16) unsigned csm[];
  __shared__ WindowSegs segs[kCommitThreads / 32];
  __shared__ int hist[HISTO_LENGTH];
  __shared__ int sAcc, sRem, sChanged, sNres;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = A.counts[b];
  const int nq = A.bow ? A.meta[4 * b] : A.qcounts[b];
  unsigned* occBits = csm;                                        // cap/32 words: the frame's initial occupancy
  unsigned* tag = occBits + ((A.cap + 31) >> 5);                  // cap words
  int* pick = reinterpret_cast<int*>(tag + A.cap);                // qcap words: keypoint picked by query q, or -1
  int* rlist = pick + A.qcap;                                     // qcap words: queries to re-score in this sweep
  const orb_keypoint* K = A.kps + (size_t)b * A.cap;
  const u8* D = A.desc + (size_t)b * A.cap * 32;
  const u8* occ0 = A.occupied ? A.occupied + (size_t)b * A.cap : nullptr;
  const float* UR = A.uright ? A.uright + (size_t)b * A.cap : nullptr;
  int* mk = A.matchOfKp + (size_t)b * A.cap;
  int* mq = A.bow ? A.mqP + (size_t)b * A.qcap : A.matchOfQuery + (size_t)b * A.qcap;
  const int* order1 = A.order1 + (size_t)b * A.qcap;
  const uint4* top = A.top + (size_t)b * A.qcap * 2;

  for (int w = tid; w < ((A.cap + 31) >> 5); w += kCommitThreads) {
    unsigned bits = 0;
    if (occ0)
      for (int j = 0; j < 32; j++) if (w * 32 + j < n && occ0[w * 32 + j]) bits |= 1u << j;
    occBits[w] = bits;
  }
  for (int i = tid; i < A.cap; i += kCommitThreads) mk[i] = -1;
  for (int q = tid; q < A.qcap; q += kCommitThreads) {
    mq[q] = -1;
    pick[q] = -1;
    if (A.bow) A.matchOfQuery[(size_t)b * A.qcap + q] = -1;
  }
  if (tid < HISTO_LENGTH) hist[tid] = 0;
  if (tid == 0) { sAcc = 0; sRem = 0; }
  __syncthreads();

  const bool needSecond = A.mode != ORB_SEARCH_BEST;
  const int need = needSecond ? 2 : 1;
  for (;;) {
    // ---- publish the picks of the previous sweep
    for (int i = tid; i < n; i += kCommitThreads) tag[i] = kNoKey;
    if (tid == 0) { sChanged = 0; sNres = 0; }
    __syncthreads();
    for (int q = tid; q < nq; q += kCommitThreads) {
      const int p = pick[q];
      if (p >= 0 && (top[2 * q + 1].x & kWordOccupies)) atomicMin(&tag[p], (unsigned)q);
    }
    __syncthreads();
    // ---- every query chooses again
    bool changed = false;
    for (int q = tid; q < nq; q += kCommitThreads) {
      const uint4 k4 = top[2 * q];
      if (k4.x == kNoKey) continue;
      const uint4 i4 = top[2 * q + 1];
      const unsigned k[4] = {k4.x, k4.y, k4.z, k4.w};
      const int w[4] = {(int)(i4.x & ~kWordOccupies), (int)i4.y, (int)i4.z, (int)i4.w};
      int wb = -1, ws = -1, nfree = 0;
      unsigned kb = kNoKey, ks = kNoKey;
#pragma unroll
      for (int m = 0; m < 4; m++)
        if (k[m] != kNoKey && !(tag[word_idx(w[m])] < (unsigned)q)) {   // not taken by an earlier query
          if (wb < 0) { wb = w[m]; kb = k[m]; }
          else if (ws < 0) { ws = w[m]; ks = k[m]; }
          nfree++;
        }
      if (k[3] != kNoKey && nfree < need) {   // more candidates exist than the four stored: exact re-score below
        rlist[atomicAdd(&sNres, 1)] = q;
        continue;
      }
      const int np = accept_match(A.mode, A.th, A.ratio, kb, wb, ks, ws) ? word_idx(wb) : -1;
      if (np != pick[q]) { pick[q] = np; changed = true; }
    }
    __syncthreads();
    const int nres = sNres;
    for (int r = wid; r < nres; r += kCommitThreads / 32) {
      const int qq = rlist[r];
      auto taken = [&](int idx) { return ((occBits[idx >> 5] >> (idx & 31)) & 1u) || tag[idx] < (unsigned)qq; };
      Top4 X;
      if (!A.bow) {
        const orb_proj_query Q = A.queries[(size_t)b * A.qcap + qq];
        WindowSegs& W = segs[wid];
        const int total = window_setup<32>(Q, A.b4, A.invW, A.invH, A.cellStart + (size_t)b * (kGridCells + 1), W, lane);
        const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + qq) * 32);
        const uint4 d0 = __ldg(qd), d1 = __ldg(qd + 1);
        const int* items = A.cellItems + (size_t)b * A.cap;
        const float r2 = __fmul_rn(Q.radius, Q.radius);
        group_top4<32>(total, [&](int t) { return window_candidate(t, W, items, K, Q, r2); }, taken, d0, d1, D, UR, Q.ur, Q.radius,
                       lane, X);
        __syncwarp();
      } else {
        bow_score<32>(A, b, qq, taken, lane, X);
      }
      if (lane == 0) {
        const int np = accept_match(A.mode, A.th, A.ratio, X.k[0], X.i[0], X.k[1], X.i[1]) ? word_idx(X.i[0]) : -1;
        if (np != pick[qq]) { pick[qq] = np; changed = true; }
      }
    }
    if (changed) sChanged = 1;
    __syncthreads();
    if (!sChanged) break;
    __syncthreads();   // everyone has read sChanged before the next sweep resets it
  }
  // ---- the fixed point is the reference's result: write it out
  int nacc = 0;
  for (int q = tid; q < nq; q += kCommitThreads) {
    const int p = pick[q];
    if (p >= 0) {
      atomicMax(&mk[p], q);   // a later query overwrites an earlier one (only possible if that one did not occupy)
      mq[q] = p;
      nacc++;
    }
  }
  if (nacc) atomicAdd(&sAcc, nacc);
  __threadfence_block();
  __syncthreads();
  // rotation consistency (ORBmatcher.cc:1830-1855 / :380-417): histogram of the accepted matches, three maxima,
  // everything else is removed
  if (A.checkOri) {
    auto angle_of = [&](int q) {
      return A.bow ? A.kps1[(size_t)b * A.qcap + order1[q]].angle : A.queries[(size_t)b * A.qcap + q].angle;
    };
    for (int q = tid; q < nq; q += kCommitThreads) {
      const int idx = mq[q];
      if (idx >= 0) atomicAdd(&hist[rot_bin(angle_of(q), K[idx].angle)], 1);
    }
    __syncthreads();
    int i1, i2, i3;
    three_maxima30(hist, i1, i2, i3);
    int nrem = 0;
    for (int q = tid; q < nq; q += kCommitThreads) {
      const int idx = mq[q];
      if (idx < 0) continue;
      const int bin = rot_bin(angle_of(q), K[idx].angle);
      if (bin != i1 && bin != i2 && bin != i3) {
        mq[q] = -1;
        mk[idx] = -1;
        nrem++;
      }
    }
    if (nrem) atomicAdd(&sRem, nrem);
    __syncthreads();
  }
  if (tid == 0) A.nmatches[b] = sAcc - sRem;
  if (A.bow) {   // positions in the (node, index) order -> keyframe feature indices
    for (int i = tid; i < n; i += kCommitThreads) {
      const int p = mk[i];
      if (p >= 0) mk[i] = order1[p];
    }
    for (int p = tid; p < nq; p += kCommitThreads) A.matchOfQuery[(size_t)b * A.qcap + order1[p]] = mq[p];
  }
}

// ---- projection of the last frame's map points (ORBmatcher.cc:1734-1775)
struct ProjectArgs {
  const float* Xw; const u8* flags; const orb_keypoint* last; const int* counts; const float* Tcw; const int* direction;
  float fx, fy, cx, cy, b4[4], mbf, th, sf[kMaxLevels];
  int qcap, nlevels;
  orb_proj_query* out;
};

__global__ void __launch_bounds__(128) k_project_last_frame(const ProjectArgs P) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.qcap) return;
  orb_proj_query q;
  q.u = q.v = q.radius = q.ur = q.angle = 0.f;
  q.min_level = 0; q.max_level = -1; q.flags = 0;
  const size_t at = (size_t)b * P.qcap + i;
  if (i < P.counts[b]) {
    const orb_keypoint kp = P.last[at];
    q.angle = kp.angle;
    const int fl = P.flags[at];
    if (fl & 1) {
      const float* X = P.Xw + at * 3;
      const float* T = P.Tcw + (size_t)b * 16;
      float c[3];
      // cv::Mat Rcw*x3Dw+tcw = cv::gemm 3x3 float case: float products and sums in index order, then the addend
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(T[4 * r], X[0]), __fmul_rn(T[4 * r + 1], X[1])), __fmul_rn(T[4 * r + 2], X[2]));
        c[r] = __fadd_rn(t, T[4 * r + 3]);
      }
      const float invz = __double2float_rn(__ddiv_rn(1.0, (double)c[2]));
      if (!(invz < 0.f)) {
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c[0]), invz), P.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c[1]), invz), P.cy);
        if (!(u < P.b4[0] || u > P.b4[1]) && !(v < P.b4[2] || v > P.b4[3])) {
          const int oct = min(max(kp.octave, 0), P.nlevels - 1);
          const int dir = P.direction ? P.direction[b] : 0;
          q.u = u; q.v = v;
          q.radius = __fmul_rn(P.th, P.sf[oct]);
          q.ur = __fsub_rn(u, __fmul_rn(P.mbf, invz));
          if (dir == 1) { q.min_level = oct; q.max_level = -1; }
          else if (dir == 2) { q.min_level = 0; q.max_level = oct; }
          else { q.min_level = oct - 1; q.max_level = oct + 1; }
          q.flags = 1 | (fl & 2);
        }
      }
    }
  }
  P.out[at] = q;
}

// ---- BoW preparation: FeatureVector order of both sides from per-feature node ids
__device__ void bitonic_sort(unsigned long long* key, int N, int tid, int nthreads) {
  for (int k = 2; k <= N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (N >> 1); t += nthreads) {
        const int l = ((t / j) * 2 * j) + (t % j), r = l + j;
        const bool up = (l & k) == 0;
        const unsigned long long a = key[l], c = key[r];
        if ((a > c) == up) { key[l] = c; key[r] = a; }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_bow_prepare(const SearchArgs A, const int* __restrict__ node1, const int* __restrict__ counts1,
                                                     const int* __restrict__ node2) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ int cnt;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n1 = counts1[b], n2 = A.counts[b];
  int* sorted2 = A.sorted2 + (size_t)b * A.cap;
  int* node2s = A.node2s + (size_t)b * A.cap;
  int* order1 = A.order1 + (size_t)b * A.qcap;
  // frame side
  int N = 1;
  while (N < max(n2, 2)) N <<= 1;
  if (tid == 0) cnt = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < N; i += 256) {
    unsigned long long k = ~0ull;
    if (i < n2) {
      const int nd = node2[(size_t)b * A.cap + i];
      if (nd >= 0) { k = ((unsigned long long)(unsigned)nd << 32) | (unsigned)i; local++; }
    }
    keys[i] = k;
  }
  atomicAdd(&cnt, local);
  __syncthreads();
  bitonic_sort(keys, N, tid, 256);
  const int m2 = cnt;
  for (int i = tid; i < m2; i += 256) { sorted2[i] = (int)(keys[i] & 0xffffffffu); node2s[i] = (int)(keys[i] >> 32); }
  __syncthreads();
  // keyframe side
  N = 1;
  while (N < max(n1, 2)) N <<= 1;
  if (tid == 0) cnt = 0;
  __syncthreads();
  local = 0;
  for (int i = tid; i < N; i += 256) {
    unsigned long long k = ~0ull;
    if (i < n1) {
      const int nd = node1[(size_t)b * A.qcap + i];
      if (nd >= 0) { k = ((unsigned long long)(unsigned)nd << 32) | (unsigned)i; local++; }
    }
    keys[i] = k;
  }
  atomicAdd(&cnt, local);
  __syncthreads();
  bitonic_sort(keys, N, tid, 256);
  const int m1 = cnt;
  if (tid == 0) A.meta[4 * b] = m1;
  for (int p = tid; p < m1; p += 256) {
    order1[p] = (int)(keys[p] & 0xffffffffu);
    const int nd = (int)(keys[p] >> 32);
    int lo = 0, hi = m2;          // lower bound of nd in node2s
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2s[mid] < nd) lo = mid + 1; else hi = mid; }
    const int c0 = lo;
    hi = m2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2s[mid] <= nd) lo = mid + 1; else hi = mid; }
    A.cb[(size_t)b * A.qcap + p] = c0;
    A.ce[(size_t)b * A.qcap + p] = lo;
  }
}

// The same ordering for a FEW pairs per call (the host-memory entry points: one keyframe pair per call). k_bow_prepare is one
// CTA per pair with 66 bitonic steps: 76 us for 2000 features, half of the whole call. Here every feature finds its position
// by counting the smaller keys (keys are distinct: they hold the index), all CTAs of the machine take part, and a second small
// kernel does the binary searches once the frame side is in place.
__global__ void __launch_bounds__(256) k_bow_rank(const SearchArgs A, const int* __restrict__ node1, const int* __restrict__ counts1,
                                                  const int* __restrict__ node2) {
  extern __shared__ __align__(16) unsigned long long keys[];
  const int b = blockIdx.z, side = blockIdx.y, tid = threadIdx.x;
  const int n = side ? counts1[b] : A.counts[b];
  const int stride = side ? A.qcap : A.cap;
  const int* node = (side ? node1 : node2) + (size_t)b * stride;
  if ((int)(blockIdx.x * 256) >= n && blockIdx.x != 0) return;
  const int np = (n + 1) & ~1;
  int valid = 0;
  for (int i = tid; i < np; i += 256) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const int nd = node[i];
      if (nd >= 0) { k = ((unsigned long long)(unsigned)nd << 32) | (unsigned)i; valid++; }
    }
    keys[i] = k;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // number of features with a node: block-wide sum of `valid`
    __shared__ int s_valid;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    if (tid == 0) A.meta[4 * b + (side ? 0 : 1)] = s_valid;
  }
  const int i = blockIdx.x * 256 + tid;
  if (i >= n) return;
  const unsigned long long mine = keys[i];
  if (mine == ~0ull) return;
  int r0 = 0, r1 = 0;
  const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys);
#pragma unroll 4
  for (int j = 0; j < (np >> 1); j++) {
    const ulonglong2 kk = k2[j];   // broadcast
    r0 += kk.x < mine;
    r1 += kk.y < mine;
  }
  const int rank = r0 + r1;
  if (side) {
    A.order1[(size_t)b * A.qcap + rank] = i;
  } else {
    A.sorted2[(size_t)b * A.cap + rank] = i;
    A.node2s[(size_t)b * A.cap + rank] = (int)(mine >> 32);
  }
}

__global__ void __launch_bounds__(256) k_bow_slices(const SearchArgs A, const int* __restrict__ node1) {
  const int b = blockIdx.y, p = blockIdx.x * 256 + threadIdx.x;
  const int m1 = A.meta[4 * b], m2 = A.meta[4 * b + 1];
  if (p >= m1) return;
  const int* node2s = A.node2s + (size_t)b * A.cap;
  const int nd = node1[(size_t)b * A.qcap + A.order1[(size_t)b * A.qcap + p]];
  int lo = 0, hi = m2;          // lower bound of nd in node2s
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2s[mid] < nd) lo = mid + 1; else hi = mid; }
  const int c0 = lo;
  hi = m2;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2s[mid] <= nd) lo = mid + 1; else hi = mid; }
  A.cb[(size_t)b * A.qcap + p] = c0;
  A.ce[(size_t)b * A.qcap + p] = lo;
}

struct ScratchLayout {
  size_t top, order1, cb, ce, sorted2, node2s, mqP, mk, meta, total;
};

ScratchLayout scratch_layout(int batch, int qcap, int cap) {
  ScratchLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += round_up(bytes, (size_t)256); return at; };
  L.top = take((size_t)batch * qcap * 32);
  L.order1 = take((size_t)batch * qcap * 4);
  L.cb = take((size_t)batch * qcap * 4);
  L.ce = take((size_t)batch * qcap * 4);
  L.mqP = take((size_t)batch * qcap * 4);
  L.sorted2 = take((size_t)batch * cap * 4);
  L.node2s = take((size_t)batch * cap * 4);
  L.mk = take((size_t)batch * cap * 4);
  L.meta = take((size_t)batch * 16);
  L.total = o;
  return L;
}

int fill_common(SearchArgs& A, const orb_device_frames* F, int qcap, const orb_search_params* P, void* scratch, int32_t* mk,
                int32_t* mq, int32_t* nm) {
  if (!F || !P || !scratch || !mq || !nm) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (!F->keypoints_un || !F->descriptors || !F->counts || F->batch <= 0 || F->capacity <= 0 || qcap <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad frame description");
  if (P->mode < ORB_SEARCH_BEST || P->mode > ORB_SEARCH_RATIO) ORB_FAIL(ORB_ERR_INVALID, "unknown search mode");
  if (F->capacity > (1 << 16) || qcap > (1 << 20)) ORB_FAIL(ORB_ERR_UNSUPPORTED, "capacity too large for the search kernels");
  memset(&A, 0, sizeof A);
  A.kps = F->keypoints_un; A.desc = F->descriptors; A.uright = F->uright; A.occupied = F->occupied; A.counts = F->counts;
  A.cellStart = F->cell_start; A.cellItems = F->cell_items;
  for (int i = 0; i < 4; i++) A.b4[i] = F->bounds[i];
  A.invW = (float)kGridCols / (F->bounds[1] - F->bounds[0]);   // Frame.cc:184-186
  A.invH = (float)kGridRows / (F->bounds[3] - F->bounds[2]);
  A.cap = F->capacity; A.qcap = qcap;
  const ScratchLayout L = scratch_layout(F->batch, qcap, F->capacity);
  u8* s = (u8*)scratch;
  A.top = (uint4*)(s + L.top); A.order1 = (int*)(s + L.order1); A.cb = (int*)(s + L.cb); A.ce = (int*)(s + L.ce);
  A.sorted2 = (int*)(s + L.sorted2); A.node2s = (int*)(s + L.node2s); A.mqP = (int*)(s + L.mqP); A.meta = (int*)(s + L.meta);
  A.mode = P->mode; A.th = P->th; A.ratio = P->nn_ratio; A.checkOri = P->check_orientation;
  A.matchOfKp = mk ? mk : (int*)(s + L.mk); A.matchOfQuery = mq; A.nmatches = nm;
  return ORB_OK;
}

int launch_search(const SearchArgs& A, int batch, cudaStream_t s) {
  k_search_scan<<<dim3((A.qcap + kScanQueries - 1) / kScanQueries, batch), 32 * kScanWarps, 0, s>>>(A);
  ORB_CUDA(cudaGetLastError());
  const size_t smem = ((size_t)((A.cap + 31) >> 5) + (size_t)A.cap + 2 * (size_t)A.qcap) * 4;
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "capacity too large for the commit kernel's shared memory");
  ORB_CUDA(raise_dynamic_smem(k_search_commit, smem));
  k_search_commit<<<batch, kCommitThreads, smem, s>>>(A);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

}  // namespace

extern "C" {

size_t orb_search_scratch_bytes(int batch, int query_capacity, int capacity) {
  if (batch <= 0 || query_capacity <= 0 || capacity <= 0) return 0;
  return scratch_layout(batch, query_capacity, capacity).total;
}

int orb_project_last_frame_device(int device, const float* d_world_pos, const uint8_t* d_mp_flags,
                                  const orb_keypoint* d_last_keypoints, const int32_t* d_last_counts, int batch,
                                  int query_capacity, const float* d_Tcw, const int32_t* d_direction, const float* cam4,
                                  const float* bounds4, float mbf, float th, const float* scale_factors, int nlevels,
                                  orb_proj_query* d_queries, void* stream) {
  if (!d_world_pos || !d_mp_flags || !d_last_keypoints || !d_last_counts || !d_Tcw || !cam4 || !bounds4 || !scale_factors ||
      !d_queries || batch <= 0 || query_capacity <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  if (nlevels <= 0 || nlevels > kMaxLevels) ORB_FAIL(ORB_ERR_UNSUPPORTED, "1..16 pyramid levels");
  ORB_CUDA(cudaSetDevice(device));
  ProjectArgs P;
  memset(&P, 0, sizeof P);
  P.Xw = d_world_pos; P.flags = d_mp_flags; P.last = d_last_keypoints; P.counts = d_last_counts; P.Tcw = d_Tcw;
  P.direction = d_direction;
  P.fx = cam4[0]; P.fy = cam4[1]; P.cx = cam4[2]; P.cy = cam4[3];
  for (int i = 0; i < 4; i++) P.b4[i] = bounds4[i];
  P.mbf = mbf; P.th = th;
  for (int i = 0; i < nlevels; i++) P.sf[i] = scale_factors[i];
  P.qcap = query_capacity; P.nlevels = nlevels; P.out = d_queries;
  k_project_last_frame<<<dim3((query_capacity + 127) / 128, batch), 128, 0, (cudaStream_t)stream>>>(P);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_search_by_projection_device(int device, const orb_device_frames* frames, const orb_proj_query* d_queries,
                                    const uint8_t* d_query_descriptors, const int32_t* d_query_counts,
                                    int query_capacity, const orb_search_params* params, void* d_scratch,
                                    int32_t* d_match_of_keypoint, int32_t* d_match_of_query, int32_t* d_nmatches,
                                    void* stream) {
  SearchArgs A;
  const int st = fill_common(A, frames, query_capacity, params, d_scratch, d_match_of_keypoint, d_match_of_query, d_nmatches);
  if (st) return st;
  if (!d_queries || !d_query_descriptors || !d_query_counts || !frames->cell_start || !frames->cell_items)
    ORB_FAIL(ORB_ERR_INVALID, "projection search needs queries, their descriptors and the frame grid");
  ORB_CUDA(cudaSetDevice(device));
  A.queries = d_queries; A.qdesc = d_query_descriptors; A.qcounts = d_query_counts;
  return launch_search(A, frames->batch, (cudaStream_t)stream);
}

// the (node id, feature index) ordering of both sides: one CTA per pair for batches, the whole machine for a few pairs
static int launch_bow_prepare(const SearchArgs& A, const int32_t* d_node1, const int32_t* d_counts1, const int32_t* d_node2, int batch,
                              int query_capacity, int capacity, cudaStream_t stream) {
  const int big = std::max(query_capacity, capacity);
  if (batch <= 8 && (size_t)(big + 2) * 8 <= 200 * 1024) {
    const size_t smem = (size_t)(big + 2) * 8;
    ORB_CUDA(raise_dynamic_smem(k_bow_rank, smem));
    k_bow_rank<<<dim3((big + 255) / 256, 2, batch), 256, smem, stream>>>(A, d_node1, d_counts1, d_node2);
    k_bow_slices<<<dim3((query_capacity + 255) / 256, batch), 256, 0, stream>>>(A, d_node1);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
  }
  int N = 2;
  while (N < big) N <<= 1;
  const size_t smem = (size_t)N * 8;
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many features for the BoW ordering kernel");
  ORB_CUDA(raise_dynamic_smem(k_bow_prepare, smem));
  k_bow_prepare<<<batch, 256, smem, stream>>>(A, d_node1, d_counts1, d_node2);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_search_by_bow_device(int device, const orb_keypoint* d_keypoints1, const uint8_t* d_descriptors1,
                             const int32_t* d_node1, const uint8_t* d_usable1, const int32_t* d_counts1,
                             int query_capacity, const orb_device_frames* frames, const int32_t* d_node2,
                             const orb_search_params* params, void* d_scratch, int32_t* d_match_of_keypoint,
                             int32_t* d_match_of_query, int32_t* d_nmatches, void* stream) {
  SearchArgs A;
  const int st = fill_common(A, frames, query_capacity, params, d_scratch, d_match_of_keypoint, d_match_of_query, d_nmatches);
  if (st) return st;
  if (!d_keypoints1 || !d_descriptors1 || !d_node1 || !d_usable1 || !d_counts1 || !d_node2)
    ORB_FAIL(ORB_ERR_INVALID, "null argument");
  ORB_CUDA(cudaSetDevice(device));
  A.bow = 1; A.kps1 = d_keypoints1; A.usable1 = d_usable1; A.qdesc = d_descriptors1; A.qcounts = d_counts1;
  A.uright = nullptr;
  const int stp = launch_bow_prepare(A, d_node1, d_counts1, d_node2, frames->batch, query_capacity, frames->capacity, (cudaStream_t)stream);
  if (stp) return stp;
  return launch_search(A, frames->batch, (cudaStream_t)stream);
}

int orb_search_for_triangulation_device(int device, const orb_keypoint* d_keypoints1, const uint8_t* d_descriptors1,
                                        const int32_t* d_node1, const uint8_t* d_has_mappoint1, const float* d_uright1,
                                        const int32_t* d_counts1, int query_capacity, const orb_device_frames* frames2,
                                        const int32_t* d_node2, const orb_triangulation_pair* d_pairs,
                                        const float* scale_factors, const float* level_sigma2, int nlevels,
                                        int check_orientation, void* d_scratch, int32_t* d_matches12, int32_t* d_nmatches,
                                        void* stream) {
  if (!d_pairs || !scale_factors || !level_sigma2) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (nlevels <= 0 || nlevels > kMaxLevels) ORB_FAIL(ORB_ERR_UNSUPPORTED, "1..16 pyramid levels");
  orb_search_params sp;
  sp.mode = ORB_SEARCH_BEST; sp.th = 50 /* TH_LOW, ORBmatcher.cc:49 */; sp.nn_ratio = 0.f; sp.check_orientation = check_orientation;
  SearchArgs A;
  const int st = fill_common(A, frames2, query_capacity, &sp, d_scratch, nullptr, d_matches12, d_nmatches);
  if (st) return st;
  if (!d_keypoints1 || !d_descriptors1 || !d_node1 || !d_has_mappoint1 || !d_counts1 || !d_node2) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  ORB_CUDA(cudaSetDevice(device));
  A.bow = 1; A.kps1 = d_keypoints1; A.usable1 = d_has_mappoint1; A.qdesc = d_descriptors1; A.qcounts = d_counts1;
  A.tri = d_pairs; A.ur1 = d_uright1;
  for (int i = 0; i < kMaxLevels; i++) {
    A.sf[i] = scale_factors[std::min(i, nlevels - 1)];
    A.sigma2[i] = level_sigma2[std::min(i, nlevels - 1)];
  }
  const int stp = launch_bow_prepare(A, d_node1, d_counts1, d_node2, frames2->batch, query_capacity, frames2->capacity, (cudaStream_t)stream);
  if (stp) return stp;
  return launch_search(A, frames2->batch, (cudaStream_t)stream);
}

// ---- host-memory form of the last-frame search: what a drop-in ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th,
// bMono) calls once per tracked frame. One packed upload, grid + projection + scan + commit, one packed download.
namespace {
struct HostSearchCache {   // per host thread and device: staging buffers and a stream
  int device = -1;
  cudaStream_t stream = nullptr;
  u8* d = nullptr; u8* h = nullptr; size_t bytes = 0;
  ~HostSearchCache() {
    if (device >= 0) { cudaSetDevice(device); cudaFree(d); if (h) cudaFreeHost(h); if (stream) cudaStreamDestroy(stream); }
  }
};
thread_local HostSearchCache t_hs;
}  // namespace

// Shared body: `nq` queries either given on the host (queries != NULL: the local-map search) or produced on the device by
// the projection kernel from the last frame's map points (last != NULL).
static int host_projection_search(int device, int nc, const orb_keypoint* cur_kps, const uint8_t* cur_desc, const float* cur_uright,
                                  const uint8_t* cur_occupied, const float* bounds4, int nq, const orb_proj_query* queries,
                                  const uint8_t* qdesc, const orb_last_frame_search* last, const orb_search_params* sp,
                                  int32_t* match_of_keypoint, int* nmatches) {
  *nmatches = 0;
  for (int i = 0; i < nc; i++) match_of_keypoint[i] = -1;
  if (nc <= 0 || nq <= 0) return ORB_OK;
  ORB_CUDA(cudaSetDevice(device));
  HostSearchCache& C = t_hs;
  if (C.device != device) {
    if (C.device >= 0) { cudaSetDevice(C.device); cudaFree(C.d); if (C.h) cudaFreeHost(C.h); if (C.stream) cudaStreamDestroy(C.stream); cudaSetDevice(device); }
    C.device = device; C.stream = nullptr; C.d = nullptr; C.h = nullptr; C.bytes = 0;
    ORB_CUDA(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
  }
  // packed layout (256-byte aligned pieces): inputs first (one H2D), then device-only scratch, then outputs (one D2H)
  size_t off = 0;
  auto piece = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t oCK = piece((size_t)nc * sizeof(orb_keypoint)), oCD = piece((size_t)nc * 32), oCU = piece((size_t)nc * 4), oCO = piece((size_t)nc);
  const size_t oLK = piece(last ? (size_t)nq * sizeof(orb_keypoint) : 0), oLX = piece(last ? (size_t)nq * 12 : 0), oLF = piece(last ? (size_t)nq : 0);
  const size_t oQD = piece((size_t)nq * 32);
  const size_t oQH = piece(last ? 0 : (size_t)nq * sizeof(orb_proj_query));   // queries given by the host
  const size_t oHdr = piece(256);   // n_cur, n_q, direction, Tcw[16]
  const size_t inEnd = off;
  const size_t oCS = piece((size_t)(kGridCells + 1) * 4), oCI = piece((size_t)nc * 4);
  const size_t oQ = last ? piece((size_t)nq * sizeof(orb_proj_query)) : oQH;
  const size_t oScr = piece(orb_search_scratch_bytes(1, nq, nc));
  const size_t outBegin = off;
  const size_t oMK = piece((size_t)nc * 4), oMQ = piece((size_t)nq * 4), oNM = piece(256);
  const size_t total = off;
  if (total > C.bytes) {
    ORB_CUDA(cudaStreamSynchronize(C.stream));
    cudaFree(C.d); if (C.h) cudaFreeHost(C.h);
    C.d = nullptr; C.h = nullptr; C.bytes = 0;
    const size_t want = total + total / 2;
    ORB_CUDA(cudaMalloc(&C.d, want));
    ORB_CUDA(cudaHostAlloc((void**)&C.h, want, cudaHostAllocDefault));
    C.bytes = want;
  }
  u8 *H = C.h, *D = C.d;
  memcpy(H + oCK, cur_kps, (size_t)nc * sizeof(orb_keypoint));
  memcpy(H + oCD, cur_desc, (size_t)nc * 32);
  if (cur_uright) memcpy(H + oCU, cur_uright, (size_t)nc * 4);
  if (cur_occupied) memcpy(H + oCO, cur_occupied, (size_t)nc); else memset(H + oCO, 0, (size_t)nc);
  memcpy(H + oQD, qdesc, (size_t)nq * 32);
  int32_t* hdr = reinterpret_cast<int32_t*>(H + oHdr);
  hdr[0] = nc; hdr[1] = nq; hdr[2] = 0;
  if (last) {
    memcpy(H + oLK, last->last_keypoints, (size_t)nq * sizeof(orb_keypoint));
    memcpy(H + oLX, last->last_world_pos, (size_t)nq * 12);
    memcpy(H + oLF, last->last_mp_flags, (size_t)nq);
    hdr[2] = last->direction;
    memcpy(hdr + 4, last->Tcw, 16 * sizeof(float));
  } else {
    memcpy(H + oQH, queries, (size_t)nq * sizeof(orb_proj_query));
  }
  cudaStream_t s = C.stream;
  ORB_CUDA(cudaMemcpyAsync(D, H, inEnd, cudaMemcpyHostToDevice, s));
  const int32_t* dHdr = reinterpret_cast<const int32_t*>(D + oHdr);
  int st = orb_assign_features_to_grid_device(device, reinterpret_cast<const orb_keypoint*>(D + oCK), dHdr, 1, nc, bounds4,
                                              reinterpret_cast<int32_t*>(D + oCS), reinterpret_cast<int32_t*>(D + oCI), s);
  if (st) return st;
  if (last) {
    st = orb_project_last_frame_device(device, reinterpret_cast<const float*>(D + oLX), D + oLF, reinterpret_cast<const orb_keypoint*>(D + oLK),
                                       dHdr + 1, 1, nq, reinterpret_cast<const float*>(dHdr + 4), dHdr + 2, last->cam4, bounds4, last->mbf,
                                       last->th, last->scale_factors, last->nlevels, reinterpret_cast<orb_proj_query*>(D + oQ), s);
    if (st) return st;
  }
  orb_device_frames fr;
  memset(&fr, 0, sizeof fr);
  fr.keypoints_un = reinterpret_cast<const orb_keypoint*>(D + oCK); fr.descriptors = D + oCD;
  fr.uright = cur_uright ? reinterpret_cast<const float*>(D + oCU) : nullptr;
  fr.occupied = D + oCO; fr.counts = dHdr;
  fr.cell_start = reinterpret_cast<const int32_t*>(D + oCS); fr.cell_items = reinterpret_cast<const int32_t*>(D + oCI);
  for (int i = 0; i < 4; i++) fr.bounds[i] = bounds4[i];
  fr.batch = 1; fr.capacity = nc;
  st = orb_search_by_projection_device(device, &fr, reinterpret_cast<const orb_proj_query*>(D + oQ), D + oQD, dHdr + 1, nq, sp, D + oScr,
                                       reinterpret_cast<int32_t*>(D + oMK), reinterpret_cast<int32_t*>(D + oMQ),
                                       reinterpret_cast<int32_t*>(D + oNM), s);
  if (st) return st;
  ORB_CUDA(cudaMemcpyAsync(H + outBegin, D + outBegin, total - outBegin, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  memcpy(match_of_keypoint, H + oMK, (size_t)nc * 4);
  *nmatches = *reinterpret_cast<const int32_t*>(H + oNM);
  return ORB_OK;
}

int orb_search_by_projection_last_frame(int device, const orb_last_frame_search* a, int32_t* match_of_keypoint, int* nmatches) {
  if (!a || !match_of_keypoint || !nmatches || !a->cur_keypoints_un || !a->cur_descriptors || !a->last_keypoints || !a->last_world_pos ||
      !a->last_mp_flags || !a->last_mp_descriptors || !a->Tcw || !a->cam4 || !a->bounds4 || !a->scale_factors)
    ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (a->n_cur < 0 || a->n_last < 0) ORB_FAIL(ORB_ERR_INVALID, "negative count");
  const orb_search_params sp = {ORB_SEARCH_BEST, a->th_dist, a->nn_ratio, a->check_orientation};
  return host_projection_search(device, a->n_cur, a->cur_keypoints_un, a->cur_descriptors, a->cur_uright, a->cur_occupied, a->bounds4,
                                a->n_last, nullptr, a->last_mp_descriptors, a, &sp, match_of_keypoint, nmatches);
}

int orb_search_by_projection_host(int device, int n_cur, const orb_keypoint* cur_keypoints_un, const uint8_t* cur_descriptors,
                                  const float* cur_uright, const uint8_t* cur_occupied, const float* bounds4, int n_queries,
                                  const orb_proj_query* queries, const uint8_t* query_descriptors, const orb_search_params* params,
                                  int32_t* match_of_keypoint, int* nmatches) {
  if (!match_of_keypoint || !nmatches || !cur_keypoints_un || !cur_descriptors || !bounds4 || !queries || !query_descriptors || !params)
    ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (n_cur < 0 || n_queries < 0) ORB_FAIL(ORB_ERR_INVALID, "negative count");
  return host_projection_search(device, n_cur, cur_keypoints_un, cur_descriptors, cur_uright, cur_occupied, bounds4, n_queries, queries,
                                query_descriptors, nullptr, params, match_of_keypoint, nmatches);
}

// ---- host-memory forms of the vocabulary-guided searches ------------------------------------------------------------
namespace {
// Packs host arrays into the thread's pinned staging block / device block (256-byte aligned pieces).
struct HostPack {
  HostSearchCache& C;
  size_t off = 0, inEnd = 0, outBegin = 0;
  struct Piece { size_t off, bytes; const void* src; };
  std::vector<Piece> in;
  explicit HostPack(HostSearchCache& c) : C(c) {}
  size_t add(size_t bytes, const void* src = nullptr) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    if (src) in.push_back(Piece{o, bytes, src});
    return o;
  }
  int prepare(int device) {   // call after every piece has been added
    if (C.device != device) {
      if (C.device >= 0) { cudaSetDevice(C.device); cudaFree(C.d); if (C.h) cudaFreeHost(C.h); if (C.stream) cudaStreamDestroy(C.stream); cudaSetDevice(device); }
      C.device = device; C.stream = nullptr; C.d = nullptr; C.h = nullptr; C.bytes = 0;
      ORB_CUDA(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
    }
    if (off > C.bytes) {
      ORB_CUDA(cudaStreamSynchronize(C.stream));
      cudaFree(C.d); if (C.h) cudaFreeHost(C.h);
      C.d = nullptr; C.h = nullptr; C.bytes = 0;
      const size_t want = off + off / 2;
      ORB_CUDA(cudaMalloc(&C.d, want));
      ORB_CUDA(cudaHostAlloc((void**)&C.h, want, cudaHostAllocDefault));
      C.bytes = want;
    }
    for (const Piece& p : in) memcpy(C.h + p.off, p.src, p.bytes);
    return ORB_OK;
  }
};
}  // namespace

int orb_search_by_bow_host(int device, int n1, const orb_keypoint* keypoints1_un, const uint8_t* descriptors1, const int32_t* node1,
                           const uint8_t* usable1, int n2, const orb_keypoint* keypoints2_un, const uint8_t* descriptors2,
                           const int32_t* node2, const uint8_t* occupied2, const orb_search_params* params, int32_t* match_of_keypoint2,
                           int32_t* match_of_query1, int* nmatches) {
  if (!keypoints1_un || !descriptors1 || !node1 || !usable1 || !keypoints2_un || !descriptors2 || !node2 || !params || !nmatches ||
      !match_of_keypoint2 || !match_of_query1)
    ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (n1 < 0 || n2 < 0) ORB_FAIL(ORB_ERR_INVALID, "negative count");
  *nmatches = 0;
  for (int i = 0; i < n2; i++) match_of_keypoint2[i] = -1;
  for (int i = 0; i < n1; i++) match_of_query1[i] = -1;
  if (n1 == 0 || n2 == 0) return ORB_OK;
  ORB_CUDA(cudaSetDevice(device));
  HostPack P(t_hs);
  std::vector<uint8_t> occ0;
  if (!occupied2) { occ0.assign(n2, 0); occupied2 = occ0.data(); }
  const int32_t counts[2] = {n1, n2};
  const size_t oK1 = P.add((size_t)n1 * sizeof(orb_keypoint), keypoints1_un), oD1 = P.add((size_t)n1 * 32, descriptors1);
  const size_t oN1 = P.add((size_t)n1 * 4, node1), oU1 = P.add((size_t)n1, usable1);
  const size_t oK2 = P.add((size_t)n2 * sizeof(orb_keypoint), keypoints2_un), oD2 = P.add((size_t)n2 * 32, descriptors2);
  const size_t oN2 = P.add((size_t)n2 * 4, node2), oO2 = P.add((size_t)n2, occupied2);
  const size_t oCnt = P.add(256, counts);
  const size_t inEnd = P.off;
  const size_t oScr = P.add(orb_search_scratch_bytes(1, n1, n2));
  const size_t outBegin = P.off;
  const size_t oMK = P.add((size_t)n2 * 4), oMQ = P.add((size_t)n1 * 4), oNM = P.add(256);
  int st = P.prepare(device);
  if (st) return st;
  HostSearchCache& C = t_hs;
  u8 *H = C.h, *D = C.d;
  memcpy(H + oCnt, counts, sizeof counts);
  cudaStream_t s = C.stream;
  ORB_CUDA(cudaMemcpyAsync(D, H, inEnd, cudaMemcpyHostToDevice, s));
  const int32_t* dCnt = reinterpret_cast<const int32_t*>(D + oCnt);
  orb_device_frames fr;
  memset(&fr, 0, sizeof fr);
  fr.keypoints_un = reinterpret_cast<const orb_keypoint*>(D + oK2); fr.descriptors = D + oD2;
  fr.occupied = D + oO2; fr.counts = dCnt + 1;
  fr.batch = 1; fr.capacity = n2;
  st = orb_search_by_bow_device(device, reinterpret_cast<const orb_keypoint*>(D + oK1), D + oD1, reinterpret_cast<const int32_t*>(D + oN1),
                                D + oU1, dCnt, n1, &fr, reinterpret_cast<const int32_t*>(D + oN2), params, D + oScr,
                                reinterpret_cast<int32_t*>(D + oMK), reinterpret_cast<int32_t*>(D + oMQ), reinterpret_cast<int32_t*>(D + oNM), s);
  if (st) return st;
  ORB_CUDA(cudaMemcpyAsync(H + outBegin, D + outBegin, P.off - outBegin, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  memcpy(match_of_keypoint2, H + oMK, (size_t)n2 * 4);
  memcpy(match_of_query1, H + oMQ, (size_t)n1 * 4);
  *nmatches = *reinterpret_cast<const int32_t*>(H + oNM);
  return ORB_OK;
}

int orb_search_for_triangulation_host(int device, int n1, const orb_keypoint* keypoints1_un, const uint8_t* descriptors1,
                                      const int32_t* node1, const uint8_t* has_mappoint1, const float* uright1, int n2,
                                      const orb_keypoint* keypoints2_un, const uint8_t* descriptors2, const int32_t* node2,
                                      const uint8_t* has_mappoint2, const float* uright2, const orb_triangulation_pair* pair,
                                      const float* scale_factors, const float* level_sigma2, int nlevels, int check_orientation,
                                      int32_t* matches12, int* nmatches) {
  if (!keypoints1_un || !descriptors1 || !node1 || !has_mappoint1 || !keypoints2_un || !descriptors2 || !node2 || !has_mappoint2 ||
      !pair || !scale_factors || !level_sigma2 || !matches12 || !nmatches)
    ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (n1 < 0 || n2 < 0) ORB_FAIL(ORB_ERR_INVALID, "negative count");
  *nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return ORB_OK;
  ORB_CUDA(cudaSetDevice(device));
  HostPack P(t_hs);
  const int32_t counts[2] = {n1, n2};
  const size_t oK1 = P.add((size_t)n1 * sizeof(orb_keypoint), keypoints1_un), oD1 = P.add((size_t)n1 * 32, descriptors1);
  const size_t oN1 = P.add((size_t)n1 * 4, node1), oH1 = P.add((size_t)n1, has_mappoint1), oR1 = P.add((size_t)n1 * 4, uright1);
  const size_t oK2 = P.add((size_t)n2 * sizeof(orb_keypoint), keypoints2_un), oD2 = P.add((size_t)n2 * 32, descriptors2);
  const size_t oN2 = P.add((size_t)n2 * 4, node2), oH2 = P.add((size_t)n2, has_mappoint2), oR2 = P.add((size_t)n2 * 4, uright2);
  const size_t oPair = P.add(256, pair), oCnt = P.add(256, counts);
  const size_t inEnd = P.off;
  const size_t oScr = P.add(orb_search_scratch_bytes(1, n1, n2));
  const size_t outBegin = P.off;
  const size_t oM12 = P.add((size_t)n1 * 4), oNM = P.add(256);
  int st = P.prepare(device);
  if (st) return st;
  HostSearchCache& C = t_hs;
  u8 *H = C.h, *D = C.d;
  memcpy(H + oPair, pair, sizeof(orb_triangulation_pair));
  memcpy(H + oCnt, counts, sizeof counts);
  cudaStream_t s = C.stream;
  ORB_CUDA(cudaMemcpyAsync(D, H, inEnd, cudaMemcpyHostToDevice, s));
  const int32_t* dCnt = reinterpret_cast<const int32_t*>(D + oCnt);
  orb_device_frames fr;
  memset(&fr, 0, sizeof fr);
  fr.keypoints_un = reinterpret_cast<const orb_keypoint*>(D + oK2); fr.descriptors = D + oD2;
  fr.uright = uright2 ? reinterpret_cast<const float*>(D + oR2) : nullptr;
  fr.occupied = D + oH2; fr.counts = dCnt + 1;
  fr.batch = 1; fr.capacity = n2;
  st = orb_search_for_triangulation_device(device, reinterpret_cast<const orb_keypoint*>(D + oK1), D + oD1, reinterpret_cast<const int32_t*>(D + oN1),
                                           D + oH1, uright1 ? reinterpret_cast<const float*>(D + oR1) : nullptr, dCnt, n1, &fr,
                                           reinterpret_cast<const int32_t*>(D + oN2), reinterpret_cast<const orb_triangulation_pair*>(D + oPair),
                                           scale_factors, level_sigma2, nlevels, check_orientation, D + oScr,
                                           reinterpret_cast<int32_t*>(D + oM12), reinterpret_cast<int32_t*>(D + oNM), s);
  if (st) return st;
  ORB_CUDA(cudaMemcpyAsync(H + outBegin, D + outBegin, P.off - outBegin, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  memcpy(matches12, H + oM12, (size_t)n1 * 4);
  *nmatches = *reinterpret_cast<const int32_t*>(H + oNM);
  return ORB_OK;
}

}  // extern "C"
