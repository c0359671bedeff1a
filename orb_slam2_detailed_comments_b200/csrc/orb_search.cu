// orb_search.cu — ordered candidate-set matchers of ORBmatcher on B200 (sm_100a)
// (reference: src/ORBmatcher.cc:72-169 SearchByProjection local map, :1710-1860 SearchByProjection last frame,
//  :247-420 SearchByBoW keyframe -> frame; candidates from src/Frame.cc:590-670 GetFeaturesInArea).
//
// The reference walks the map points in order; each one takes the best candidate keypoint that no earlier
// map point occupies. That chain is split in two:
//   k_search_scan    one WARP per query, all queries of all frames in parallel: the candidates are
//                    enumerated in the reference's visiting order (lanes = candidates), filtered against the
//                    frame's initial occupancy / stereo coordinate, and the FOUR smallest keys
//                    (distance << 22 | visiting rank) are kept (redux.sync min + ballot).
//   k_search_commit  one warp per frame walks the queries in order, 32 per step. A query whose stored
//                    candidates lost nothing to earlier commits is final; one that lost candidates falls
//                    back on its 3rd / 4th stored candidate, and only if those run out is it re-scored
//                    exactly against the current occupancy. Same-step dependencies are found with a
//                    shared-memory tag per keypoint (atomicMin of the lane) and resolved in lane order.
// The two smallest keys by (distance, visiting order) are exactly the reference's (best, second best):
// `dist < bestDist` / `else if (dist < bestDist2)` with demotion keeps the lexicographic top two.
// Integer-pipe work (LOP3 / POPC / REDUX); no tensor cores by design.
#include <algorithm>
#include <cstring>

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

// The four smallest (distance, visiting rank) keys among candidates t = 0..total-1; idxOf(t) returns the
// keypoint index or -1 when the candidate is filtered out before the occupancy test.
template <class IdxFn, class OccFn>
__device__ __forceinline__ void warp_top4(int total, IdxFn idxOf, OccFn occ, const uint4 q0, const uint4 q1, const u8* desc,
                                          const float* uright, float ur, float radius, int lane, Top4& T) {
  top4_clear(T);
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    unsigned key = kNoKey;
    int idx = -1;
    if (t < total) {
      idx = idxOf(t);
      if (idx >= 0 && !occ(idx)) {
        bool ok = true;
        if (uright) {   // ORBmatcher.cc:125-130 / :1792-1798
          const float r = uright[idx];
          if (r > 0.f && fabsf(__fsub_rn(ur, r)) > radius) ok = false;
        }
        if (ok) key = ((unsigned)hamming256(q0, q1, desc + (size_t)idx * 32) << 22) | (unsigned)t;
      }
    }
    for (;;) {
      const unsigned m = __reduce_min_sync(kFull, key);
      if (m >= T.k[3]) break;
      const int owner = __ffs(__ballot_sync(kFull, key == m)) - 1;
      const int mi = __shfl_sync(kFull, idx, owner);
      top4_insert(T, m, mi);
      if (lane == owner) key = kNoKey;
    }
  }
}

// Frame::GetFeaturesInArea (Frame.cc:590-670) as a flattened candidate range: the cells (ix, cy0..cy1) of one
// grid column are contiguous in the CSR layout, so the window is <= 64 segments visited in ix order.
struct WindowSegs {
  int seg[kMaxCols];       // first CSR slot of column c
  int cum[kMaxCols + 1];   // candidates before column c
};

__device__ __forceinline__ int window_setup(const orb_proj_query& Q, const float* B4, float invW, float invH, const int* start,
                                            WindowSegs& W, int lane) {
  const float x = Q.u, y = Q.v, r = Q.radius;
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, B4[0]), r), invW)));
  const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, B4[0]), r), invW)));
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, B4[2]), r), invH)));
  const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, B4[2]), r), invH)));
  if (!(cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0) || cx1 < cx0 || cy1 < cy0) return 0;
  const int ncols = cx1 - cx0 + 1;
  int run = 0;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    const int c = c0 + lane;
    int b = 0, len = 0;
    if (c < ncols) {
      b = start[(cx0 + c) * kGridRows + cy0];
      len = start[(cx0 + c) * kGridRows + cy1 + 1] - b;
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (c < ncols) { W.seg[c] = b; W.cum[c] = run + incl - len; }
    run += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) W.cum[ncols] = run;
  __syncwarp();
  return run;
}

// candidate t of the window -> keypoint index, or -1 if the level / circle test rejects it
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
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2 ? idx : -1;
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
  // scratch
  uint4* top; int* order1; int* cb; int* ce; int* sorted2; int* node2s; int* mqP; int* meta;
  // params / outputs
  int mode, th, checkOri; float ratio;
  int* matchOfKp; int* matchOfQuery; int* nmatches;
};

// ---- pass 1: score every query against the frame's initial state
__global__ void __launch_bounds__(32 * kScanWarps) k_search_scan(const SearchArgs A) {
  __shared__ WindowSegs segs[kScanWarps];
  const int b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kScanWarps + wid;
  const int nq = A.bow ? A.meta[4 * b] : A.qcounts[b];
  if (q >= nq) return;
  const orb_keypoint* K = A.kps + (size_t)b * A.cap;
  const u8* D = A.desc + (size_t)b * A.cap * 32;
  const u8* occ0 = A.occupied ? A.occupied + (size_t)b * A.cap : nullptr;
  Top4 T;
  top4_clear(T);
  if (!A.bow) {
    const orb_proj_query Q = A.queries[(size_t)b * A.qcap + q];
    if (Q.flags & 1) {
      const int total = window_setup(Q, A.b4, A.invW, A.invH, A.cellStart + (size_t)b * (kGridCells + 1), segs[wid], lane);
      const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + q) * 32);
      const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
      const int* items = A.cellItems + (size_t)b * A.cap;
      const float r2 = __fmul_rn(Q.radius, Q.radius);
      const WindowSegs& W = segs[wid];
      warp_top4(total, [&](int t) { return window_candidate(t, W, items, K, Q, r2); },
                [&](int idx) { return occ0 && occ0[idx]; }, q0, q1, D,
                A.uright ? A.uright + (size_t)b * A.cap : nullptr, Q.ur, Q.radius, lane, T);
    }
  } else {
    const int i1 = A.order1[(size_t)b * A.qcap + q];
    if (A.usable1[(size_t)b * A.qcap + i1]) {
      const int c0 = A.cb[(size_t)b * A.qcap + q], c1 = A.ce[(size_t)b * A.qcap + q];
      const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + i1) * 32);
      const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
      const int* s2 = A.sorted2 + (size_t)b * A.cap;
      warp_top4(c1 - c0, [&](int t) { return s2[c0 + t]; }, [&](int) { return false; }, q0, q1, D, nullptr, 0.f, 0.f, lane, T);
    }
  }
  if (lane == 0) {
    uint4* o = A.top + ((size_t)b * A.qcap + q) * 2;
    o[0] = make_uint4(T.k[0], T.k[1], T.k[2], T.k[3]);
    o[1] = make_uint4((unsigned)T.i[0], (unsigned)T.i[1], (unsigned)T.i[2], (unsigned)T.i[3]);
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

// acceptance of (best, second) — ORBmatcher.cc:155-158 / :331-336 / :1807
__device__ __forceinline__ bool accept_match(int mode, int th, float ratio, unsigned kb, int ib, unsigned ks, int is,
                                             const orb_keypoint* K) {
  if (ib < 0) return false;
  const int bd = (int)(kb >> 22);
  if (bd > th) return false;
  const int bd2 = is >= 0 ? (int)(ks >> 22) : 256;
  if (mode == ORB_SEARCH_RATIO_LEVEL) {
    const int l1 = K[ib].octave, l2 = is >= 0 ? K[is].octave : -1;
    if (l1 == l2 && (float)bd > __fmul_rn(ratio, (float)bd2)) return false;
  } else if (mode == ORB_SEARCH_RATIO) {
    if (!((float)bd < __fmul_rn(ratio, (float)bd2))) return false;
  }
  return true;
}

// ---- pass 2: in-order commit, one warp per frame
__global__ void __launch_bounds__(32) k_search_commit(const SearchArgs A) {
  extern __shared__ __align__(16) unsigned csm[];
  __shared__ WindowSegs segs;
  __shared__ int hist[HISTO_LENGTH];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = A.counts[b];
  const int nq = A.bow ? A.meta[4 * b] : A.qcounts[b];
  unsigned* occBits = csm;                               // cap/32 words
  unsigned* tag = csm + ((A.cap + 31) >> 5);             // cap words: lowest lane that claims the keypoint this step
  const orb_keypoint* K = A.kps + (size_t)b * A.cap;
  const u8* D = A.desc + (size_t)b * A.cap * 32;
  const u8* occ0 = A.occupied ? A.occupied + (size_t)b * A.cap : nullptr;
  const float* UR = A.uright ? A.uright + (size_t)b * A.cap : nullptr;
  int* mk = A.matchOfKp + (size_t)b * A.cap;
  int* mq = A.bow ? A.mqP + (size_t)b * A.qcap : A.matchOfQuery + (size_t)b * A.qcap;
  const int* order1 = A.order1 + (size_t)b * A.qcap;

  for (int w = lane; w < ((A.cap + 31) >> 5); w += 32) {
    unsigned bits = 0;
    if (occ0)
      for (int j = 0; j < 32; j++) {
        const int i = w * 32 + j;
        if (i < n && occ0[i]) bits |= 1u << j;
      }
    occBits[w] = bits;
  }
  for (int i = lane; i < A.cap; i += 32) { tag[i] = kNoKey; mk[i] = -1; }
  for (int q = lane; q < A.qcap; q += 32) {
    mq[q] = -1;
    if (A.bow) A.matchOfQuery[(size_t)b * A.qcap + q] = -1;
  }
  if (lane < HISTO_LENGTH) hist[lane] = 0;
  __syncwarp();

  const bool needSecond = A.mode != ORB_SEARCH_BEST;
  int nacc = 0;
  auto occ = [&](int idx) { return (occBits[idx >> 5] >> (idx & 31)) & 1u; };
  auto commit = [&](int q, int idx, bool occupies, float angle) {
    atomicMax(&mk[idx], q);      // a later query overwrites an earlier one (only possible if that one did not occupy)
    mq[q] = idx;
    if (occupies) atomicOr(&occBits[idx >> 5], 1u << (idx & 31));
    if (A.checkOri) atomicAdd(&hist[rot_bin(angle, K[idx].angle)], 1);
    nacc++;
  };

  for (int q0 = 0; q0 < nq; q0 += 32) {
    const int q = q0 + lane;
    Top4 T;
    top4_clear(T);
    bool occupies = true;
    float angle = 0.f;
    if (q < nq) {
      const uint4* t = A.top + ((size_t)b * A.qcap + q) * 2;
      const uint4 k4 = t[0], i4 = t[1];
      T.k[0] = k4.x; T.k[1] = k4.y; T.k[2] = k4.z; T.k[3] = k4.w;
      T.i[0] = (int)i4.x; T.i[1] = (int)i4.y; T.i[2] = (int)i4.z; T.i[3] = (int)i4.w;
      if (A.bow) {
        angle = A.kps1[(size_t)b * A.qcap + order1[q]].angle;
      } else {
        const orb_proj_query& Q = A.queries[(size_t)b * A.qcap + q];
        occupies = (Q.flags & 2) != 0;
        angle = Q.angle;
      }
    }
    bool pending = T.k[0] != kNoKey;
    while (__any_sync(kFull, pending)) {
      int ib = -1, is = -1;
      unsigned kb = kNoKey, ks = kNoKey;
#pragma unroll
      for (int m = 0; m < 4; m++)
        if (T.k[m] != kNoKey && !occ(T.i[m])) {
          if (ib < 0) { ib = T.i[m]; kb = T.k[m]; }
          else if (is < 0) { is = T.i[m]; ks = T.k[m]; }
        }
      const bool complete = T.k[3] == kNoKey;
      const bool needFull = pending && !complete && (ib < 0 || (needSecond && is < 0));
      const bool acc = pending && !needFull && accept_match(A.mode, A.th, A.ratio, kb, ib, ks, is, K);
      const bool claims = acc && occupies;
      if (claims) atomicMin(&tag[ib], (unsigned)lane);
      __syncwarp();
      bool conflicted = false;
      if (pending && !needFull) {
        if (ib >= 0 && tag[ib] < (unsigned)lane) conflicted = true;
        if (needSecond && is >= 0 && tag[is] < (unsigned)lane) conflicted = true;
      }
      const unsigned badMask = __ballot_sync(kFull, pending && (needFull || conflicted));
      const int first = badMask ? __ffs(badMask) - 1 : 32;
      __syncwarp();
      if (claims) tag[ib] = kNoKey;
      if (pending && lane < first) {
        if (acc) commit(q, ib, occupies, angle);
        pending = false;
      }
      __syncwarp();
      if (first < 32 && __shfl_sync(kFull, (int)needFull, first)) {
        // exact re-score of query q0 + first against the current occupancy (its stored candidates ran out)
        const int qq = q0 + first;
        Top4 X;
        if (!A.bow) {
          const orb_proj_query Q = A.queries[(size_t)b * A.qcap + qq];
          const int total = window_setup(Q, A.b4, A.invW, A.invH, A.cellStart + (size_t)b * (kGridCells + 1), segs, lane);
          const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + qq) * 32);
          const uint4 d0 = __ldg(qd), d1 = __ldg(qd + 1);
          const int* items = A.cellItems + (size_t)b * A.cap;
          const float r2 = __fmul_rn(Q.radius, Q.radius);
          warp_top4(total, [&](int t) { return window_candidate(t, segs, items, K, Q, r2); }, occ, d0, d1, D, UR, Q.ur, Q.radius,
                    lane, X);
        } else {
          const int i1 = order1[qq];
          const int c0 = A.cb[(size_t)b * A.qcap + qq], c1 = A.ce[(size_t)b * A.qcap + qq];
          const uint4* qd = reinterpret_cast<const uint4*>(A.qdesc + ((size_t)b * A.qcap + i1) * 32);
          const uint4 d0 = __ldg(qd), d1 = __ldg(qd + 1);
          const int* s2 = A.sorted2 + (size_t)b * A.cap;
          warp_top4(c1 - c0, [&](int t) { return s2[c0 + t]; }, occ, d0, d1, D, nullptr, 0.f, 0.f, lane, X);
        }
        if (lane == first) {
          if (accept_match(A.mode, A.th, A.ratio, X.k[0], X.i[0], X.k[1], X.i[1], K)) commit(q, X.i[0], occupies, angle);
          pending = false;
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  __threadfence_block();
  // rotation consistency (ORBmatcher.cc:1830-1855 / :380-417)
  int nrem = 0;
  if (A.checkOri) {
    int i1, i2, i3;
    three_maxima30(hist, i1, i2, i3);
    for (int q = lane; q < nq; q += 32) {
      const int idx = mq[q];
      if (idx < 0) continue;
      const float a1 = A.bow ? A.kps1[(size_t)b * A.qcap + order1[q]].angle : A.queries[(size_t)b * A.qcap + q].angle;
      const int bin = rot_bin(a1, K[idx].angle);
      if (bin != i1 && bin != i2 && bin != i3) {
        mq[q] = -1;
        mk[idx] = -1;
        nrem++;
      }
    }
  }
  int tot = nacc - nrem;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
  if (lane == 0) A.nmatches[b] = tot;
  if (A.bow) {   // positions in the (node, index) order -> keyframe feature indices
    __syncwarp();
    __threadfence_block();
    for (int i = lane; i < n; i += 32) {
      const int p = mk[i];
      if (p >= 0) mk[i] = order1[p];
    }
    for (int p = lane; p < nq; p += 32) A.matchOfQuery[(size_t)b * A.qcap + order1[p]] = mq[p];
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

struct ScratchLayout {
  size_t top, order1, cb, ce, sorted2, node2s, mqP, meta, total;
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
  L.meta = take((size_t)batch * 16);
  L.total = o;
  return L;
}

int fill_common(SearchArgs& A, const orb_device_frames* F, int qcap, const orb_search_params* P, void* scratch, int32_t* mk,
                int32_t* mq, int32_t* nm) {
  if (!F || !P || !scratch || !mk || !mq || !nm) ORB_FAIL(ORB_ERR_INVALID, "null argument");
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
  A.matchOfKp = mk; A.matchOfQuery = mq; A.nmatches = nm;
  return ORB_OK;
}

int launch_search(const SearchArgs& A, int batch, cudaStream_t s) {
  k_search_scan<<<dim3((A.qcap + kScanWarps - 1) / kScanWarps, batch), 32 * kScanWarps, 0, s>>>(A);
  ORB_CUDA(cudaGetLastError());
  const size_t smem = ((size_t)((A.cap + 31) >> 5) + (size_t)A.cap) * 4;
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "capacity too large for the commit kernel's shared memory");
  if (smem > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_search_commit<<<batch, 32, smem, s>>>(A);
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
  A.uright = nullptr; A.occupied = nullptr;
  int N = 2;
  while (N < std::max(query_capacity, frames->capacity)) N <<= 1;
  const size_t smem = (size_t)N * 8;
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many features for the BoW ordering kernel");
  if (smem > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_bow_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bow_prepare<<<frames->batch, 256, smem, (cudaStream_t)stream>>>(A, d_node1, d_counts1, d_node2);
  ORB_CUDA(cudaGetLastError());
  return launch_search(A, frames->batch, (cudaStream_t)stream);
}

}  // extern "C"
