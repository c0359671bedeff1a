// orb_match.cu — B200 (sm_100a) implementation of the ORBmatcher hot path
// (reference: src/ORBmatcher.cc:573-717 SearchForInitialization, :2035 ComputeThreeMaxima,
//  :2083 DescriptorDistance; candidate window from src/Frame.cc:590-670).
//
// Integer-pipe work (LOP3 / POPC / IADD3 / VIMNMX); no tensor cores by design.
//   k_match_pairs    one CTA per frame pair. The descriptors of frame 2 live in REGISTERS
//                    (thread t owns candidates t, t+256, ...) together with their
//                    vMatchedDistance state, so the sequential row walk of the reference
//                    costs one broadcast load of the row + XOR/POPC + one block reduction.
//   k_allpairs       one CTA per (row keyframe, column-keyframe span): row descriptors and
//                    their running (best, second) in registers, column descriptors broadcast
//                    from shared memory; no cross-thread reduction in the inner loop.
//   k_hamming_matrix plain distance matrix (parity aid).
#include <algorithm>
#include <climits>
#include <cstring>
#include <cmath>
#include <vector>

#include <dlfcn.h>

#include <mutex>
#include <string>
#include <vector>

#include "orb_common.cuh"

using namespace orbb200;

namespace {

constexpr int kMatchThreads = 256;
constexpr int TH_LOW = 50;        // ORBmatcher.cc:49
constexpr int HISTO_LENGTH = 30;  // ORBmatcher.cc:51
constexpr unsigned kNoKey = 0xffffffffu;

// 256-bit Hamming distance. Carry-save form: 7 words are compressed to (ones, twos, fours)
// with 4 full adders (2 LOP3 each), so only 4 POPC are issued instead of 8.
__device__ __forceinline__ int hamming8(const unsigned a[8], const unsigned b[8]) {
#if defined(ORB_NAIVE_POPC)
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
  return d;
#else
  const unsigned x0 = a[0] ^ b[0], x1 = a[1] ^ b[1], x2 = a[2] ^ b[2], x3 = a[3] ^ b[3];
  const unsigned x4 = a[4] ^ b[4], x5 = a[5] ^ b[5], x6 = a[6] ^ b[6], x7 = a[7] ^ b[7];
  const unsigned s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
  const unsigned s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
  const unsigned ones = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
  const unsigned twos = c1 ^ c2 ^ c3, fours = (c1 & c2) | (c3 & (c1 ^ c2));
  return __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
#endif
}

// two smallest keys of the union of two (k1<=k2) pairs
__device__ __forceinline__ void merge2(unsigned& k1, unsigned& k2, unsigned o1, unsigned o2) {
  const unsigned hi = max(k1, o1);
  k1 = min(k1, o1);
  k2 = min(hi, min(k2, o2));
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2035-2077)
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < L; i++) {
    const int s = histo[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
  else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

struct PairArgs {
  // per-pair strides are implied: every frame holds n1 / n2 keypoints
  const u8* desc1; const u8* desc2;       // pair p: desc1 + p*stride1*32 ...
  const float* ang1; const float* ang2;
  const float* xy1; const float* xy2;     // windowed mode only
  const int* oct1; const int* oct2;       // windowed mode only
  float* prev;                            // windowed mode: vbPrevMatched (in/out), n1 x 2
  int n1, n2;
  long long stride1, stride2;             // keypoints between consecutive pairs
  float nnratio; int checkOri; int window;
  float minX, minY, invW, invH;
  int* matches12; int* nmatches; int* best; int* second;  // best/second optional
};

// dynamic smem: m12[n1] int, m21[n2] int, binOf[n1] u8 (4-byte padded)
template <int CPT, bool WINDOWED>
__global__ void __launch_bounds__(kMatchThreads) k_match_pairs(const PairArgs A) {
  extern __shared__ __align__(16) int msm[];
  __shared__ unsigned s_red[2][kMatchThreads / 32][2];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_nmatch;
  __shared__ int s_keep[3];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p = blockIdx.x;
  const int n1 = A.n1, n2 = A.n2;
  int* m12 = msm;
  int* m21 = msm + n1;
  u8* binOf = (u8*)(m21 + n2);
  const u8* D1 = A.desc1 + (size_t)p * A.stride1 * 32;
  const u8* D2 = A.desc2 + (size_t)p * A.stride2 * 32;
  const float* ang1 = A.ang1 + (size_t)p * A.stride1;
  const float* ang2 = A.ang2 + (size_t)p * A.stride2;

  // frame-2 descriptors + vMatchedDistance state into registers
  unsigned b[CPT][8];
  int md[CPT];
  float cx[WINDOWED ? CPT : 1], cy[WINDOWED ? CPT : 1];
  int cg[WINDOWED ? CPT : 1];
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    const int i2 = tid + c * kMatchThreads;
    md[c] = INT_MAX;
    if (i2 < n2) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(D2 + (size_t)i2 * 32));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(D2 + (size_t)i2 * 32) + 1);
      b[c][0] = lo.x; b[c][1] = lo.y; b[c][2] = lo.z; b[c][3] = lo.w;
      b[c][4] = hi.x; b[c][5] = hi.y; b[c][6] = hi.z; b[c][7] = hi.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) b[c][k] = 0;
      md[c] = -1;  // never a candidate: vMatchedDistance <= dist always
    }
    if (WINDOWED) {
      cg[c] = -1;
      cx[c] = cy[c] = 0.f;
      if (i2 < n2) {
        const float* xy2 = A.xy2 + (size_t)p * A.stride2 * 2;
        const float x = xy2[2 * i2], y = xy2[2 * i2 + 1];
        cx[c] = x; cy[c] = y;
        // Frame::PosInGrid (Frame.cc:682-698) + level filter of GetFeaturesInArea(...,0,0)
        const int gx = (int)roundf(__fmul_rn(__fsub_rn(x, A.minX), A.invW));
        const int gy = (int)roundf(__fmul_rn(__fsub_rn(y, A.minY), A.invH));
        const int o = (A.oct2 + (size_t)p * A.stride2)[i2];
        if (gx >= 0 && gx < 64 && gy >= 0 && gy < 48 && o == 0) cg[c] = gx | (gy << 8);
      }
    }
  }
  for (int i = tid; i < n1; i += kMatchThreads) { m12[i] = -1; binOf[i] = 255; }
  for (int i = tid; i < n2; i += kMatchThreads) m21[i] = -1;
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) s_nmatch = 0;
  __syncthreads();

  const float factor = HISTO_LENGTH / 360.0f;  // this fork: ORBmatcher.cc:585-586
  int nmatches = 0;                            // tracked by thread 0
#pragma unroll 1
  for (int i1 = 0; i1 < n1; i1++) {
    int gx0 = 0, gx1 = 63, gy0 = 0, gy1 = 47;
    float qx = 0.f, qy = 0.f, r2 = 0.f;
    bool rowActive = true;
    if (WINDOWED) {
      // level1 > 0 -> continue (:599); Frame::GetFeaturesInArea cell range (Frame.cc:603-634)
      const int o1 = (A.oct1 + (size_t)p * A.stride1)[i1];
      const float* pv = A.prev + (size_t)p * A.stride1 * 2;
      qx = pv[2 * i1]; qy = pv[2 * i1 + 1];
      const float r = (float)A.window;
      r2 = __fmul_rn(r, r);
      gx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(qx, A.minX), r), A.invW)));
      gx1 = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(qx, A.minX), r), A.invW)));
      gy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(qy, A.minY), r), A.invH)));
      gy1 = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(qy, A.minY), r), A.invH)));
      rowActive = o1 <= 0 && gx0 < 64 && gx1 >= 0 && gy0 < 48 && gy1 >= 0;
    }
    unsigned k1 = kNoKey, k2 = kNoKey;
    if (rowActive) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32) + 1);
      const unsigned a[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int c = 0; c < CPT; c++) {
        const int dist = hamming8(a, b[c]);
        bool ok = md[c] > dist;  // vMatchedDistance[i2] <= dist -> skip (:627)
        if (WINDOWED) {
          const int gx = cg[c] & 0xff, gy = cg[c] >> 8;
          const float dx = __fsub_rn(cx[c], qx), dy = __fsub_rn(cy[c], qy);
          ok = ok && cg[c] >= 0 && gx >= gx0 && gx <= gx1 && gy >= gy0 && gy <= gy1 &&
               __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2;  // circular window (Frame.cc:664)
        }
        const unsigned key = ok ? ((unsigned)dist << 16) | (unsigned)(tid + c * kMatchThreads) : kNoKey;
        merge2(k1, k2, key, kNoKey);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
      merge2(k1, k2, o1, o2);
    }
    const int buf = i1 & 1;
    if (lane == 0) { s_red[buf][wid][0] = k1; s_red[buf][wid][1] = k2; }
    __syncthreads();
    k1 = kNoKey; k2 = kNoKey;
#pragma unroll
    for (int w = 0; w < kMatchThreads / 32; w++) merge2(k1, k2, s_red[buf][w][0], s_red[buf][w][1]);
    const int best = k1 == kNoKey ? INT_MAX : (int)(k1 >> 16);
    const int second = k2 == kNoKey ? INT_MAX : (int)(k2 >> 16);
    const int bestIdx = (int)(k1 & 0xffffu);
    if (tid == 0 && A.best) {
      A.best[(size_t)p * n1 + i1] = best;
      A.second[(size_t)p * n1 + i1] = second;
    }
    if (best <= TH_LOW && (float)best < __fmul_rn((float)second, A.nnratio)) {  // :644-647
#pragma unroll
      for (int c = 0; c < CPT; c++)
        if (bestIdx == tid + c * kMatchThreads) md[c] = best;  // vMatchedDistance[bestIdx2] = bestDist
      if (tid == 0) {
        const int prevOwner = m21[bestIdx];
        if (prevOwner >= 0) { m12[prevOwner] = -1; nmatches--; }  // :650-654
        m12[i1] = bestIdx;
        m21[bestIdx] = i1;
        nmatches++;
        if (A.checkOri) {
          float rot = __fsub_rn(ang1[i1], ang2[bestIdx]);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          int bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == HISTO_LENGTH) bin = 0;
          if (bin >= 0 && bin < HISTO_LENGTH) { binOf[i1] = (u8)bin; s_hist[bin]++; }
        }
      }
    }
  }
  if (tid == 0) {
    s_nmatch = nmatches;
    int a = -1, b2 = -1, c = -1;
    if (A.checkOri) three_maxima(s_hist, HISTO_LENGTH, a, b2, c);
    s_keep[0] = a; s_keep[1] = b2; s_keep[2] = c;
  }
  __syncthreads();
  if (A.checkOri) {
    int dropped = 0;
    for (int i = tid; i < n1; i += kMatchThreads) {
      const int bin = binOf[i];
      if (bin != 255 && bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2] && m12[i] >= 0) {
        m12[i] = -1;  // :692-706
        dropped++;
      }
    }
    if (dropped) atomicSub(&s_nmatch, dropped);
  }
  __syncthreads();
  int* out12 = A.matches12 + (size_t)p * n1;
  for (int i = tid; i < n1; i += kMatchThreads) {
    const int m = m12[i];
    out12[i] = m;
    if (WINDOWED && m >= 0) {  // :712-714
      float* pv = A.prev + (size_t)p * A.stride1 * 2;
      const float* xy2 = A.xy2 + (size_t)p * A.stride2 * 2;
      pv[2 * i] = xy2[2 * m];
      pv[2 * i + 1] = xy2[2 * m + 1];
    }
  }
  if (tid == 0) A.nmatches[p] = s_nmatch;
}

// ------------------------------------------------------------------------------------------
// Windowed SearchForInitialization, latency form (one pair per call: Tracking::MonocularInitialization, Tracking.cc:915-926).
// k_match_pairs walks the rows of frame 1 inside ONE CTA with a block reduction per row: 0.65 us per row, 1.3 ms for 2000
// keypoints - no faster than the reference's own loop on a host core. The distances are not sequential, only
// vMatchedDistance (:627) and the steal (:650-654) are. So:
//   k_sfi_scan  every active row (octave 0, :599) gets a warp somewhere on the machine: all frame-2 keypoints are tested
//               against the window exactly as GetFeaturesInArea does (cell range, octave 0, circle), and the candidates
//               with distance <= dmax are stored sorted, up to kSfiK per row. dmax is the largest distance that can still
//               change a decision: a best above TH_LOW is rejected anyway, and a second best with
//               (dmax + 1) * nnratio > TH_LOW accepts every best <= TH_LOW (:644-647), exactly as "no second" does.
//   k_sfi_walk  one warp per pair walks the rows in order over the stored lists (shared memory): a candidate is skipped
//               iff vMatchedDistance[i2] <= its distance, the first two that are not give best and second. A row whose
//               list overflowed and ran out is re-scored exactly by the warp. Then the rotation histogram and the
//               outputs, as in k_match_pairs.
// The per-row best / second distances (a test aid of this library, not part of the reference's interface) need the exact
// second best: callers that ask for them get k_match_pairs.
// ------------------------------------------------------------------------------------------
constexpr int kSfiK = 8;               // stored candidates per row
constexpr int kSfiRowsPerWarp = 2;
constexpr int kSfiChunk = 1024;        // rows staged per step of the walk

struct SfiRow {   // GetFeaturesInArea window of row i1 (Frame.cc:603-634)
  float qx, qy, r2;
  int gx0, gx1, gy0, gy1;
  bool active;
};

__device__ __forceinline__ SfiRow sfi_row(const PairArgs& A, int p, int i1) {
  SfiRow R;
  const int o1 = (A.oct1 + (size_t)p * A.stride1)[i1];
  const float* pv = A.prev + (size_t)p * A.stride1 * 2;
  R.qx = pv[2 * i1]; R.qy = pv[2 * i1 + 1];
  const float r = (float)A.window;
  R.r2 = __fmul_rn(r, r);
  R.gx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(R.qx, A.minX), r), A.invW)));
  R.gx1 = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(R.qx, A.minX), r), A.invW)));
  R.gy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(R.qy, A.minY), r), A.invH)));
  R.gy1 = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(R.qy, A.minY), r), A.invH)));
  R.active = o1 <= 0 && R.gx0 < 64 && R.gx1 >= 0 && R.gy0 < 48 && R.gy1 >= 0;   // level1 > 0 -> continue (:599)
  return R;
}

// candidate i2 of the row's window? (octave 0, Frame::PosInGrid Frame.cc:682-698, cell range, circle Frame.cc:664)
__device__ __forceinline__ bool sfi_in_window(const PairArgs& A, int p, const SfiRow& R, int i2) {
  if ((A.oct2 + (size_t)p * A.stride2)[i2] != 0) return false;
  const float* xy2 = A.xy2 + (size_t)p * A.stride2 * 2;
  const float x = xy2[2 * i2], y = xy2[2 * i2 + 1];
  const int gx = (int)roundf(__fmul_rn(__fsub_rn(x, A.minX), A.invW));
  const int gy = (int)roundf(__fmul_rn(__fsub_rn(y, A.minY), A.invH));
  if (gx < 0 || gx >= 64 || gy < 0 || gy >= 48) return false;
  if (gx < R.gx0 || gx > R.gx1 || gy < R.gy0 || gy > R.gy1) return false;
  const float dx = __fsub_rn(x, R.qx), dy = __fsub_rn(y, R.qy);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < R.r2;
}

__device__ __forceinline__ int sfi_distance(const unsigned a[8], const u8* D2, int i2) {
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(D2 + (size_t)i2 * 32));
  const uint4 hi = __ldg(reinterpret_cast<const uint4*>(D2 + (size_t)i2 * 32) + 1);
  const unsigned b[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  return hamming8(a, b);
}

__global__ void __launch_bounds__(256) k_sfi_scan(const PairArgs A, unsigned* __restrict__ keys, int* __restrict__ cnt, int dmax) {
  __shared__ unsigned s_list[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, p = blockIdx.y;
  const unsigned lt = (1u << lane) - 1u;
  const u8* D1 = A.desc1 + (size_t)p * A.stride1 * 32;
  const u8* D2 = A.desc2 + (size_t)p * A.stride2 * 32;
  for (int r = 0; r < kSfiRowsPerWarp; r++) {
    const int i1 = (blockIdx.x * 8 + wid) * kSfiRowsPerWarp + r;
    if (i1 >= A.n1) break;
    const SfiRow R = sfi_row(A, p, i1);
    int total = 0;
    if (R.active) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32) + 1);
      const unsigned a[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      for (int base = 0; base < A.n2; base += 32) {
        const int i2 = base + lane;
        bool ok = i2 < A.n2 && sfi_in_window(A, p, R, i2);
        unsigned key = kNoKey;
        if (ok) {
          const int dist = sfi_distance(a, D2, i2);
          ok = dist <= dmax;
          key = ((unsigned)dist << 16) | (unsigned)i2;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = total + __popc(m & lt);
          if (pos < 32) s_list[wid][pos] = key;
        }
        total += __popc(m);
      }
      __syncwarp();
      // sort the (<= 32) stored keys by counting; keys are distinct (they hold the index)
      const int have = min(total, 32);
      const unsigned mine = lane < have ? s_list[wid][lane] : kNoKey;
      int rank = 0;
      for (int j = 0; j < have; j++) rank += s_list[wid][j] < mine;
      if (lane < have && rank < kSfiK) keys[((size_t)p * A.n1 + i1) * kSfiK + rank] = mine;
      __syncwarp();
    }
    // more than 32 candidates <= dmax: the stored ones are not the smallest - the walk re-scores the row (flag 0x200)
    if (lane == 0) cnt[(size_t)p * A.n1 + i1] = total > 32 ? 0x200 : (min(total, kSfiK) | (total > kSfiK ? 0x100 : 0));
  }
}

// dynamic smem: m12[n1] int, m21[n2] int, vmd[n2] u16 (4-byte padded), binOf[n1] u8 (4-byte padded), keys[kSfiChunk * kSfiK], cnt[kSfiChunk],
// act[kSfiChunk] u16
__global__ void __launch_bounds__(256) k_sfi_walk(const PairArgs A, const unsigned* __restrict__ keys, const int* __restrict__ cnt,
                                                   int dmax) {
  extern __shared__ __align__(16) int wsm[];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_nmatch;
  __shared__ int s_keep[3];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, p = blockIdx.x;
  const int n1 = A.n1, n2 = A.n2;
  int* m12 = wsm;
  int* m21 = m12 + n1;
  unsigned short* vmd = reinterpret_cast<unsigned short*>(m21 + n2);
  u8* binOf = reinterpret_cast<u8*>(vmd + ((n2 + 1) & ~1));
  unsigned* skeys = reinterpret_cast<unsigned*>(binOf + ((n1 + 3) & ~3));
  int* scnt = reinterpret_cast<int*>(skeys + kSfiChunk * kSfiK);
  unsigned short* act = reinterpret_cast<unsigned short*>(scnt + kSfiChunk);
  const u8* D1 = A.desc1 + (size_t)p * A.stride1 * 32;
  const u8* D2 = A.desc2 + (size_t)p * A.stride2 * 32;
  const float* ang1 = A.ang1 + (size_t)p * A.stride1;
  const float* ang2 = A.ang2 + (size_t)p * A.stride2;
  for (int i = tid; i < n1; i += 256) { m12[i] = -1; binOf[i] = 255; }
  for (int i = tid; i < n2; i += 256) { m21[i] = -1; vmd[i] = 0xffffu; }
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  const float factor = HISTO_LENGTH / 360.0f;
  int nmatches = 0;   // warp 0
  for (int c0 = 0; c0 < n1; c0 += kSfiChunk) {
    const int rows = min(kSfiChunk, n1 - c0);
    __syncthreads();
    for (int i = tid; i < rows; i += 256) scnt[i] = cnt[(size_t)p * n1 + c0 + i];
    for (int i = tid; i < rows * kSfiK; i += 256) skeys[i] = keys[((size_t)p * n1 + c0) * kSfiK + i];
    __syncthreads();
    if (wid != 0) continue;
    // the chunk's active rows, in order
    int nAct = 0;
    for (int base = 0; base < rows; base += 32) {
      const bool on = base + lane < rows && scnt[base + lane] != 0;
      const unsigned mm = __ballot_sync(0xffffffffu, on);
      if (on) act[nAct + __popc(mm & ((1u << lane) - 1u))] = (unsigned short)(base + lane);
      nAct += __popc(mm);
    }
    __syncwarp();
    // Four rows per step, eight lanes each (a stored list has <= 8 entries). Rows of a step are independent unless an
    // earlier row of the step takes a keypoint that a later one has among its candidates (vMatchedDistance :627, the steal
    // :650-654): the step then ends before that row. A row that needs the exact re-scoring is done alone by the warp.
    const int g = lane >> 3, gl = lane & 7;
    int pos = 0;
    while (pos < nAct) {
      const int row = pos + g < nAct ? (int)act[pos + g] : -1;
      const int ci = row >= 0 ? scnt[row] : 0;
      const int nk = ci & 0xff;
      const unsigned key = gl < nk ? skeys[row * kSfiK + gl] : kNoKey;
      const bool free_ = key != kNoKey && (int)vmd[key & 0xffffu] > (int)(key >> 16);   // vMatchedDistance[i2] <= dist -> skip (:627)
      const unsigned mAll = __ballot_sync(0xffffffffu, free_);
      unsigned m = (mAll >> (8 * g)) & 0xffu;
      const bool exactNeeded = (ci & 0x300) && (__popc(m) < 2 || (ci & 0x200));
      unsigned kb = kNoKey, ks = kNoKey;
      int L = 1;
      if (__shfl_sync(0xffffffffu, (int)exactNeeded, 0)) {
        // the stored list of the first row ran out (or never held the smallest): exact top two against the current state
        const int i1 = c0 + (int)act[pos];
        const SfiRow R = sfi_row(A, p, i1);
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(D1 + (size_t)i1 * 32) + 1);
        const unsigned a[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        unsigned k1 = kNoKey, k2 = kNoKey;
        for (int i2 = lane; i2 < n2; i2 += 32) {
          if (!sfi_in_window(A, p, R, i2)) continue;
          const int dist = sfi_distance(a, D2, i2);
          if (dist > dmax || !((int)vmd[i2] > dist)) continue;
          merge2(k1, k2, ((unsigned)dist << 16) | (unsigned)i2, kNoKey);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
          merge2(k1, k2, o1, o2);
        }
        kb = g == 0 ? k1 : kNoKey; ks = g == 0 ? k2 : kNoKey;   // only the first row is decided in this step
      } else {
        // best / second of every group: its first two free entries (the lists are sorted)
        const int f1 = m ? __ffs(m) - 1 : 0;
        const unsigned k1 = __shfl_sync(0xffffffffu, key, 8 * g + f1);
        if (m) { kb = k1; m &= m - 1; }
        const int f2 = m ? __ffs(m) - 1 : 0;
        const unsigned k2 = __shfl_sync(0xffffffffu, key, 8 * g + f2);
        if (m) ks = k2;
      }
      const int best = kb == kNoKey ? INT_MAX : (int)(kb >> 16);
      const int second = ks == kNoKey ? INT_MAX : (int)(ks >> 16);
      const bool accept = row >= 0 && best <= TH_LOW && (float)best < __fmul_rn((float)second, A.nnratio);   // :644-647
      const int bestIdx = (int)(kb & 0xffffu);
      if (!__shfl_sync(0xffffffffu, (int)exactNeeded, 0)) {
        // how many leading rows of the step stand: row k + 1 falls if it needs the exact path or a standing row takes one of
        // its candidates
        unsigned hit = 0;
#pragma unroll
        for (int gp = 0; gp < 3; gp++) {
          const int accP = __shfl_sync(0xffffffffu, (int)accept, 8 * gp), bP = __shfl_sync(0xffffffffu, bestIdx, 8 * gp);
          hit |= __ballot_sync(0xffffffffu, g > gp && accP && key != kNoKey && (int)(key & 0xffffu) == bP);
          const int okNext = __shfl_sync(0xffffffffu, (int)(row >= 0 && !exactNeeded), 8 * (gp + 1));
          if (L == gp + 1 && okNext && !((hit >> (8 * (gp + 1))) & 0xffu)) L = gp + 2;
        }
      }
      int delta = 0;
      if (gl == 0 && g < L && accept) {
        const int i1 = c0 + row;
        vmd[bestIdx] = (unsigned short)best;   // vMatchedDistance[bestIdx2] = bestDist
        const int prevOwner = m21[bestIdx];
        if (prevOwner >= 0) { m12[prevOwner] = -1; delta--; }   // :650-654
        m12[i1] = bestIdx;
        m21[bestIdx] = i1;
        delta++;
        if (A.checkOri) {
          float rot = __fsub_rn(ang1[i1], ang2[bestIdx]);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          int bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == HISTO_LENGTH) bin = 0;
          if (bin >= 0 && bin < HISTO_LENGTH) { binOf[i1] = (u8)bin; atomicAdd(&s_hist[bin], 1); }
        }
      }
      nmatches += __reduce_add_sync(0xffffffffu, delta);
      __syncwarp();   // the next step reads vmd
      pos += L;
    }
  }
  if (tid == 0) {
    s_nmatch = nmatches;
    int a = -1, b2 = -1, c = -1;
    if (A.checkOri) three_maxima(s_hist, HISTO_LENGTH, a, b2, c);
    s_keep[0] = a; s_keep[1] = b2; s_keep[2] = c;
  }
  __syncthreads();
  if (A.checkOri) {
    int dropped = 0;
    for (int i = tid; i < n1; i += 256) {
      const int bin = binOf[i];
      if (bin != 255 && bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2] && m12[i] >= 0) {
        m12[i] = -1;  // :692-706
        dropped++;
      }
    }
    if (dropped) atomicSub(&s_nmatch, dropped);
  }
  __syncthreads();
  int* out12 = A.matches12 + (size_t)p * n1;
  for (int i = tid; i < n1; i += 256) {
    const int mm = m12[i];
    out12[i] = mm;
    if (mm >= 0) {  // :712-714
      float* pv = A.prev + (size_t)p * A.stride1 * 2;
      const float* xy2 = A.xy2 + (size_t)p * A.stride2 * 2;
      pv[2 * i] = xy2[2 * mm];
      pv[2 * i + 1] = xy2[2 * mm + 1];
    }
  }
  if (tid == 0) A.nmatches[p] = s_nmatch;
}

// ------------------------------------------------------------------------------------------
// Brute-force SearchForInitialization, throughput form. The reference's row walk is sequential only
// through vMatchedDistance (:627) and the steal (:650-654); the distances themselves are not. So:
//   phase 1  every thread owns RPT rows of frame 1 (descriptors + the two smallest keys in
//            registers) and streams frame 2's descriptors from shared memory (broadcast loads): no
//            cross-thread reduction, no barrier in the inner loop. Keys = dist << 16 | index.
//   phase 2  one thread walks the rows in order. A row's top-2 is still exact if neither candidate
//            has meanwhile been matched with a distance <= the row's (the reference would skip it).
//            Otherwise the decision is often still forced (best > TH_LOW: reject; best < ratio *
//            lower bound of second: accept); only the remaining rows are recomputed exactly against
//            the current vMatchedDistance by the whole CTA.
// Results are identical to the sequential algorithm; when the caller asks for the per-row best /
// second distances every non-exact row is recomputed.
// ------------------------------------------------------------------------------------------
constexpr int kBfThreads = 512;
constexpr int kBfMaxN = 2048;

template <int RPT>
__global__ void __launch_bounds__(kBfThreads, 2) k_match_pairs_bf(const PairArgs A) {
  extern __shared__ __align__(16) unsigned char bfsm[];
  __shared__ unsigned s_red[kBfThreads / 32][2];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_nmatch, s_row, s_need, s_exact;
  __shared__ int s_keep[3];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p = blockIdx.x;
  const int n1 = A.n1, n2 = A.n2;
  // region 0: frame-2 descriptors (phase 1), then the rows' top-2 keys (phase 2)
  const size_t region0 = (size_t)(n2 * 32 > n1 * 8 ? n2 * 32 : n1 * 8);
  uint4* bdesc = reinterpret_cast<uint4*>(bfsm);
  uint2* keys = reinterpret_cast<uint2*>(bfsm);
  unsigned short* md = reinterpret_cast<unsigned short*>(bfsm + region0);   // vMatchedDistance (0xffff = INT_MAX)
  short* m12 = reinterpret_cast<short*>(md + n2);
  short* m21 = m12 + n1;
  short* accIdx = m21 + n2;   // frame-2 index a row was matched to when it was accepted (-1: never)
  int* tag = reinterpret_cast<int*>(accIdx + n1);   // byte offset region0 + 4*(n1+n2): 4-byte aligned;   // lowest lane of the current step matched to a candidate
  u8* binOf = reinterpret_cast<u8*>(tag + n2);
  const u8* D1 = A.desc1 + (size_t)p * A.stride1 * 32;
  const u8* D2 = A.desc2 + (size_t)p * A.stride2 * 32;
  const float* ang1 = A.ang1 + (size_t)p * A.stride1;
  const float* ang2 = A.ang2 + (size_t)p * A.stride2;

  {
    const uint4* src = reinterpret_cast<const uint4*>(D2);
    for (int t = tid; t < n2 * 2; t += kBfThreads) bdesc[t] = __ldg(src + t);
  }
  for (int i = tid; i < n2; i += kBfThreads) { md[i] = 0xffffu; m21[i] = -1; tag[i] = INT_MAX; }
  for (int i = tid; i < n1; i += kBfThreads) { m12[i] = -1; accIdx[i] = -1; binOf[i] = 255; }
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) { s_nmatch = 0; s_row = 0; s_need = -1; s_exact = -1; }
  unsigned a[RPT][8];
  unsigned k1[RPT], k2[RPT];
#pragma unroll
  for (int r = 0; r < RPT; r++) {
    const int row = tid + r * kBfThreads;
    k1[r] = kNoKey; k2[r] = kNoKey;
    if (row < n1) {
      const uint4* src = reinterpret_cast<const uint4*>(D1 + (size_t)row * 32);
      const uint4 lo = __ldg(src), hi = __ldg(src + 1);
      a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = lo.z; a[r][3] = lo.w;
      a[r][4] = hi.x; a[r][5] = hi.y; a[r][6] = hi.z; a[r][7] = hi.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) a[r][k] = 0;
    }
  }
  __syncthreads();
  // ---- phase 1: two smallest keys of every row over all of frame 2
#pragma unroll 2
  for (int c = 0; c < n2; c++) {
    const uint4 lo = bdesc[2 * c], hi = bdesc[2 * c + 1];
    const unsigned bb[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const unsigned x = ((unsigned)hamming8(a[r], bb) << 16) + (unsigned)c;
      k2[r] = min(k2[r], max(k1[r], x));
      k1[r] = min(k1[r], x);
    }
  }
  __syncthreads();   // everybody is done with the descriptors: region 0 becomes the key table
#pragma unroll
  for (int r = 0; r < RPT; r++) {
    const int row = tid + r * kBfThreads;
    if (row < n1) keys[row] = make_uint2(k1[r], k2[r]);
  }
  __syncthreads();

  // ---- phase 2: in-order commit by warp 0, 32 rows per step, with cooperative exact recomputation
  // on demand. The 32 rows are decided in parallel against the state before the step; the longest
  // prefix in which no row touches a candidate that an EARLIER row of the step has just been matched
  // to is committed (those commits are independent of each other), the rest is re-evaluated.
  const float factor = HISTO_LENGTH / 360.0f;  // this fork: ORBmatcher.cc:585-586
  const bool wantDist = A.best != nullptr;
  for (;;) {
    if (wid == 0) {
      int row0 = s_row, nmatches = s_nmatch, need = -1;
      const int exactRow = s_exact;
      while (row0 < n1) {
        const int row = row0 + lane;
        const bool active = row < n1;
        const uint2 kk = active ? keys[row] : make_uint2(kNoKey, kNoKey);
        const int d1 = kk.x == kNoKey ? INT_MAX : (int)(kk.x >> 16), c1 = (int)(kk.x & 0xffffu);
        const int d2 = kk.y == kNoKey ? INT_MAX : (int)(kk.y >> 16), c2 = (int)(kk.y & 0xffffu);
        int best = d1, second = d2, bestIdx = c1;
        bool decided = !active, accept = false;
        if (active) {
          if (row == exactRow) {
            decided = true;   // computed against the current vMatchedDistance
          } else {
            const bool s1 = kk.x != kNoKey && (int)md[c1] <= d1;   // would the reference skip it? (:627)
            const bool s2 = kk.y != kNoKey && (int)md[c2] <= d2;
            if (!s1 && !s2) decided = true;
            else if (!wantDist) {
              if (!s1) {                       // best stands, second >= d2
                if (d1 > TH_LOW) { decided = true; best = INT_MAX; }
                else if ((float)d1 < __fmul_rn((float)d2, A.nnratio)) { decided = true; accept = true; }
              } else if (d2 > TH_LOW) {        // best >= d2 > TH_LOW whichever candidate it is
                decided = true; best = INT_MAX;
              }
            }
          }
          if (decided && !accept) accept = best <= TH_LOW && (float)best < __fmul_rn((float)second, A.nnratio);  // :644-647
        }
        // which rows depend on a match made by an earlier row of this step?
        if (accept) atomicMin(&tag[bestIdx], lane);
        __syncwarp();
        const bool conflict = active && ((kk.x != kNoKey && tag[c1] < lane) || (kk.y != kNoKey && tag[c2] < lane));
        __syncwarp();
        if (accept) tag[bestIdx] = INT_MAX;
        const unsigned undecided = __ballot_sync(0xffffffffu, active && !decided);
        const unsigned bad = undecided | __ballot_sync(0xffffffffu, conflict);
        const int P = min(bad ? __ffs(bad) - 1 : 32, n1 - row0);
        const bool commit = lane < P;
        int stolen = 0;
        if (commit) {
          if (wantDist) {
            A.best[(size_t)p * n1 + row] = best;
            A.second[(size_t)p * n1 + row] = second;
          }
          if (accept) {
            const int prevOwner = m21[bestIdx];
            if (prevOwner >= 0) { m12[prevOwner] = -1; stolen = 1; }  // :650-654
            m12[row] = (short)bestIdx;
            m21[bestIdx] = (short)row;
            md[bestIdx] = (unsigned short)best;
            accIdx[row] = (short)bestIdx;   // the rotation bin is filled in afterwards, in parallel
          }
        }
        nmatches += __popc(__ballot_sync(0xffffffffu, commit && accept)) - __popc(__ballot_sync(0xffffffffu, stolen != 0));
        __syncwarp();
        row0 += P;
        if (P < 32 && row0 < n1 && ((undecided >> P) & 1u)) { need = row0; break; }
      }
      if (lane == 0) { s_row = row0; s_nmatch = nmatches; s_need = need; }
    }
    __syncthreads();
    const int need = s_need;
    if (need < 0) break;
    // exact 2-NN of row `need` against the current vMatchedDistance, by the whole CTA
    {
      const uint4* src = reinterpret_cast<const uint4*>(D1 + (size_t)need * 32);
      const uint4 lo = __ldg(src), hi = __ldg(src + 1);
      const unsigned ar[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      unsigned q1 = kNoKey, q2 = kNoKey;
      for (int c = tid; c < n2; c += kBfThreads) {
        const uint4* bs = reinterpret_cast<const uint4*>(D2 + (size_t)c * 32);
        const uint4 blo = __ldg(bs), bhi = __ldg(bs + 1);
        const unsigned bb[8] = {blo.x, blo.y, blo.z, blo.w, bhi.x, bhi.y, bhi.z, bhi.w};
        const int d = hamming8(ar, bb);
        if ((int)md[c] > d || md[c] == 0xffffu) merge2(q1, q2, ((unsigned)d << 16) | (unsigned)c, kNoKey);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned o1 = __shfl_xor_sync(0xffffffffu, q1, o), o2 = __shfl_xor_sync(0xffffffffu, q2, o);
        merge2(q1, q2, o1, o2);
      }
      if (lane == 0) { s_red[wid][0] = q1; s_red[wid][1] = q2; }
      __syncthreads();
      if (tid == 0) {
        q1 = kNoKey; q2 = kNoKey;
        for (int w = 0; w < kBfThreads / 32; w++) merge2(q1, q2, s_red[w][0], s_red[w][1]);
        keys[need] = make_uint2(q1, q2);
        s_exact = need;
      }
      __syncthreads();
    }
  }
  // rotation histogram (:663-677): every row that was accepted counts, also if it was robbed later
  if (A.checkOri) {
    for (int i = tid; i < n1; i += kBfThreads) {
      const int j = accIdx[i];
      if (j >= 0) {
        float rot = __fsub_rn(ang1[i], ang2[j]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == HISTO_LENGTH) bin = 0;
        if (bin >= 0 && bin < HISTO_LENGTH) { binOf[i] = (u8)bin; atomicAdd(&s_hist[bin], 1); }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int x = -1, y = -1, z = -1;
    if (A.checkOri) three_maxima(s_hist, HISTO_LENGTH, x, y, z);
    s_keep[0] = x; s_keep[1] = y; s_keep[2] = z;
  }
  __syncthreads();
  if (A.checkOri) {
    int dropped = 0;
    for (int i = tid; i < n1; i += kBfThreads) {
      const int bin = binOf[i];
      if (bin != 255 && bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2] && m12[i] >= 0) {
        m12[i] = -1;  // :692-706
        dropped++;
      }
    }
    if (dropped) atomicSub(&s_nmatch, dropped);
  }
  __syncthreads();
  int* out12 = A.matches12 + (size_t)p * n1;
  for (int i = tid; i < n1; i += kBfThreads) out12[i] = m12[i];
  if (tid == 0) A.nmatches[p] = s_nmatch;
}

// ------------------------------------------------------------------------------------------
constexpr int kApThreads = 256;

// grid: (column spans, row keyframes). Rows of keyframe i in registers (RPT per thread), each
// column keyframe's descriptors are staged in shared memory and broadcast.
template <int RPT>
__global__ void __launch_bounds__(kApThreads) k_allpairs(const u8* __restrict__ all, int nKF, int nDesc, int rowBegin,
                                                         int colBegin, int colEnd, int colsPerBlock, float nnratio,
                                                         int* __restrict__ counts) {
  extern __shared__ __align__(16) uint4 bsm[];  // nDesc x 2 uint4
  __shared__ int s_cnt[kApThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i = rowBegin + blockIdx.y;
  const int j0 = colBegin + blockIdx.x * colsPerBlock, j1 = min(j0 + colsPerBlock, colEnd);
  unsigned a[RPT][8];
#pragma unroll
  for (int r = 0; r < RPT; r++) {
    const int row = tid + r * kApThreads;
    if (row < nDesc) {
      const uint4* src = reinterpret_cast<const uint4*>(all + ((size_t)i * nDesc + row) * 32);
      const uint4 lo = __ldg(src), hi = __ldg(src + 1);
      a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = lo.z; a[r][3] = lo.w;
      a[r][4] = hi.x; a[r][5] = hi.y; a[r][6] = hi.z; a[r][7] = hi.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) a[r][k] = 0;
    }
  }
  for (int j = j0; j < j1; j++) {
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(all + (size_t)j * nDesc * 32);
    for (int t = tid; t < nDesc * 2; t += kApThreads) bsm[t] = __ldg(src + t);
    __syncthreads();
    int best[RPT], second[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) { best[r] = INT_MAX; second[r] = INT_MAX; }
#pragma unroll 2
    for (int c = 0; c < nDesc; c++) {
      const uint4 lo = bsm[2 * c], hi = bsm[2 * c + 1];
      const unsigned bb[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int r = 0; r < RPT; r++) {
        const int d = hamming8(a[r], bb);
        second[r] = min(second[r], max(best[r], d));
        best[r] = min(best[r], d);
      }
    }
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < RPT; r++)
      if (tid + r * kApThreads < nDesc && best[r] <= TH_LOW && (float)best[r] < __fmul_rn((float)second[r], nnratio)) cnt++;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kApThreads / 32; w++) t += s_cnt[w];
      counts[(size_t)blockIdx.y * nKF + j] = t;
    }
  }
}

__global__ void k_hamming_matrix(const u8* __restrict__ A, int na, const u8* __restrict__ B, int nb, int* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= nb) return;
  const uint4* pa = reinterpret_cast<const uint4*>(A + (size_t)i * 32);
  const uint4* pb = reinterpret_cast<const uint4*>(B + (size_t)j * 32);
  const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
  const unsigned a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const unsigned b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  out[(size_t)i * nb + j] = hamming8(a, b);
}

// ---- integer pipe microbenchmarks (roofline denominators for matching) -------------------
template <int WHAT>
__global__ void __launch_bounds__(256) k_int_pipe(unsigned* out, int iters, unsigned seed) {
  unsigned x0 = threadIdx.x ^ seed, x1 = x0 * 3u + 1u, x2 = x0 * 5u + 2u, x3 = x0 * 7u + 3u;
  unsigned x4 = x0 * 11u + 4u, x5 = x0 * 13u + 5u, x6 = x0 * 17u + 6u, x7 = x0 * 19u + 7u;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (WHAT == 0) {  // POPC: 8 independent chains
        x0 = __popc(x0) + seed; x1 = __popc(x1) + seed; x2 = __popc(x2) + seed; x3 = __popc(x3) + seed;
        x4 = __popc(x4) + seed; x5 = __popc(x5) + seed; x6 = __popc(x6) + seed; x7 = __popc(x7) + seed;
      } else if (WHAT == 1) {  // LOP3: 8 independent chains of a non-trivial 3-input function
        x0 = (x0 & x1) ^ seed; x1 = (x1 & x2) ^ seed; x2 = (x2 & x3) ^ seed; x3 = (x3 & x4) ^ seed;
        x4 = (x4 & x5) ^ seed; x5 = (x5 & x6) ^ seed; x6 = (x6 & x7) ^ seed; x7 = (x7 & x0) ^ seed;
      } else {                 // the matcher's mix: 1 POPC per 4 LOP3 (2 POPC + 8 LOP3 per step)
        x0 = __popc(x0) + seed; x1 = (x1 & x2) ^ seed; x2 = (x2 & x3) ^ seed; x3 = (x3 & x5) ^ seed;
        x5 = (x5 & x6) ^ seed; x4 = __popc(x4) + seed; x6 = (x6 & x7) ^ seed; x7 = (x7 & x1) ^ seed;
        x1 = (x1 ^ x3) & seed; x2 = (x2 ^ x5) | seed;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

template <int CPT, bool W>
int launch_match(const PairArgs& A, int pairs, size_t smem, cudaStream_t s) {
  ORB_CUDA(raise_dynamic_smem(k_match_pairs<CPT, W>, smem));
  k_match_pairs<CPT, W><<<pairs, kMatchThreads, smem, s>>>(A);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

template <int RPT>
int launch_match_bf(const PairArgs& A, int pairs, size_t smem, cudaStream_t s) {
  ORB_CUDA(raise_dynamic_smem(k_match_pairs_bf<RPT>, smem));
  k_match_pairs_bf<RPT><<<pairs, kBfThreads, smem, s>>>(A);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int dispatch_match_bf(const PairArgs& A, int pairs, cudaStream_t s) {
  const size_t region0 = (size_t)std::max(A.n2 * 32, A.n1 * 8);
  const size_t smem = round_up(region0 + (size_t)A.n2 * 2 + (size_t)A.n1 * 2 + (size_t)A.n2 * 2 + (size_t)A.n1 * 2 + 4 + (size_t)A.n2 * 4 + (size_t)A.n1 + 16, (size_t)16);
  const int rpt = (A.n1 + kBfThreads - 1) / kBfThreads;
  if (rpt <= 1) return launch_match_bf<1>(A, pairs, smem, s);
  if (rpt <= 2) return launch_match_bf<2>(A, pairs, smem, s);
  return launch_match_bf<4>(A, pairs, smem, s);
}

template <bool W>
int dispatch_match(const PairArgs& A, int pairs, cudaStream_t s) {
  if (!W && A.n1 <= kBfMaxN && A.n2 <= kBfMaxN && A.n1 > 0 && A.n2 > 0) return dispatch_match_bf(A, pairs, s);
  const size_t smem = (size_t)(A.n1 + A.n2) * 4 + round_up((size_t)A.n1, (size_t)16);
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many keypoints per frame for the matcher's shared memory");
  const int cpt = (A.n2 + kMatchThreads - 1) / kMatchThreads;
  if (cpt <= 1) return launch_match<1, W>(A, pairs, smem, s);
  if (cpt <= 2) return launch_match<2, W>(A, pairs, smem, s);
  if (cpt <= 4) return launch_match<4, W>(A, pairs, smem, s);
  if (cpt <= 8) return launch_match<8, W>(A, pairs, smem, s);
  if (cpt <= 12) return launch_match<12, W>(A, pairs, smem, s);
  // the monocular initialisation extractor asks for 2*nFeatures (src/Tracking.cc:188): 4000 (+3 per level) on KITTI
  if (cpt <= 16) return launch_match<16, W>(A, pairs, smem, s);
  if (cpt <= 20) return launch_match<20, W>(A, pairs, smem, s);
  ORB_FAIL(ORB_ERR_UNSUPPORTED, "more than 5120 keypoints in frame 2");
}

}  // namespace

struct orb_matcher {
  int device = 0;
  int maxPairs = 0, maxKp = 0;
  cudaStream_t stream = nullptr;
  // staging for the host single-pair entry point: one device block and one pinned host block with the same packed
  // layout [desc1|desc2|ang1|ang2|xy2|oct1|oct2|prev|m12|best|second|nm] - one upload, one download, one synchronisation
  u8* d_stage = nullptr; u8* h_stage = nullptr; size_t stageBytes = 0;
  // orb_match_allpairs_nccl: the exchange runs on its own stream, one event per peer block
  cudaStream_t commStream = nullptr;
  cudaEvent_t evReady = nullptr;
  std::vector<cudaEvent_t> evBlock;
};

extern "C" {

int orb_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  // ORBmatcher.cc:2083-2103 computes the same number with a SWAR popcount on 8 x 32 bits.
  uint64_t x[4], y[4];
  memcpy(x, a, 32);
  memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
         __builtin_popcountll(x[3] ^ y[3]);
}

int orb_matcher_create(int device, int max_pairs, int max_keypoints, orb_matcher** out) {
  if (!out || max_keypoints <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  *out = nullptr;
  ORB_CUDA(cudaSetDevice(device));
  orb_matcher* m = new orb_matcher();
  m->device = device;
  m->maxPairs = std::max(1, max_pairs);
  m->maxKp = max_keypoints;
  const size_t K = (size_t)max_keypoints;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  m->stageBytes = K * (2 * 32 + 2 * 4 + 8 + 2 * 4 + 8 + 3 * 4 + kSfiK * 4 + 4) + 18 * 16;
  if (e == cudaSuccess) e = cudaMalloc(&m->d_stage, m->stageBytes);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&m->h_stage, m->stageBytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    orb_matcher_destroy(m);
    return cuda_fail(e, "orb_matcher_create", __FILE__, __LINE__);
  }
  *out = m;
  return ORB_OK;
}

int orb_matcher_destroy(orb_matcher* m) {
  if (!m) return ORB_OK;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->d_stage);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->commStream) cudaStreamDestroy(m->commStream);
  if (m->evReady) cudaEventDestroy(m->evReady);
  for (cudaEvent_t ev : m->evBlock) cudaEventDestroy(ev);
  delete m;
  return ORB_OK;
}

int orb_search_for_initialization(orb_matcher* m, const orb_frame_view* f1, const orb_frame_view* f2,
                                  const orb_match_params* mp, float* prev_matched, int32_t* matches12, int* nmatches,
                                  int32_t* best, int32_t* second) {
  if (!m || !f1 || !f2 || !mp || !matches12 || !nmatches) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  const int n1 = f1->n, n2 = f2->n;
  if (n1 < 0 || n2 < 0 || n1 > m->maxKp || n2 > m->maxKp) ORB_FAIL(ORB_ERR_INVALID, "keypoint count exceeds matcher capacity");
  if (mp->mode == 0 && !prev_matched) ORB_FAIL(ORB_ERR_INVALID, "windowed mode needs prev_matched");
  *nmatches = 0;
  for (int i = 0; i < n1; i++) {
    matches12[i] = -1;
    if (best) best[i] = INT_MAX;
    if (second) second[i] = INT_MAX;
  }
  if (n1 == 0 || n2 == 0) return ORB_OK;
  ORB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  const bool windowed = mp->mode == 0;
  if (windowed && (!f1->octave || !f2->octave || !f2->xy)) ORB_FAIL(ORB_ERR_INVALID, "windowed mode needs octaves and frame-2 positions");
  // packed layout of this call (16-byte aligned pieces)
  size_t off = 0;
  auto piece = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  const size_t oD1 = piece((size_t)n1 * 32), oD2 = piece((size_t)n2 * 32), oA1 = piece((size_t)n1 * 4), oA2 = piece((size_t)n2 * 4);
  const size_t oXY2 = piece(windowed ? (size_t)n2 * 8 : 0), oO1 = piece(windowed ? (size_t)n1 * 4 : 0);
  const size_t oO2 = piece(windowed ? (size_t)n2 * 4 : 0), oPrev = piece(windowed ? (size_t)n1 * 8 : 0);
  const size_t inEnd = off;
  const size_t oM12 = piece((size_t)n1 * 4), oBest = piece((size_t)n1 * 4), oSecond = piece((size_t)n1 * 4), oNm = piece(16);
  const size_t outBegin = windowed ? oPrev : oM12, outEnd = off;
  // device-only scratch of the latency form (k_sfi_scan / k_sfi_walk): candidate lists of the rows
  const bool latencyForm = windowed && !best && !second && n2 <= 65535;
  const size_t oKeys = piece(latencyForm ? (size_t)n1 * kSfiK * 4 : 0), oCnt = piece(latencyForm ? (size_t)n1 * 4 : 0);
  if (off > m->stageBytes) ORB_FAIL(ORB_ERR_INVALID, "keypoint count exceeds matcher capacity");
  u8* H = m->h_stage; u8* D = m->d_stage;
  memcpy(H + oD1, f1->descriptors, (size_t)n1 * 32); memcpy(H + oD2, f2->descriptors, (size_t)n2 * 32);
  memcpy(H + oA1, f1->angle, (size_t)n1 * 4); memcpy(H + oA2, f2->angle, (size_t)n2 * 4);
  if (windowed) {
    memcpy(H + oXY2, f2->xy, (size_t)n2 * 8);
    memcpy(H + oO1, f1->octave, (size_t)n1 * 4); memcpy(H + oO2, f2->octave, (size_t)n2 * 4);
    memcpy(H + oPrev, prev_matched, (size_t)n1 * 8);
  }
  ORB_CUDA(cudaMemcpyAsync(D, H, inEnd, cudaMemcpyHostToDevice, s));
  PairArgs A;
  memset(&A, 0, sizeof A);
  A.desc1 = D + oD1; A.desc2 = D + oD2;
  A.ang1 = reinterpret_cast<float*>(D + oA1); A.ang2 = reinterpret_cast<float*>(D + oA2);
  A.n1 = n1; A.n2 = n2; A.stride1 = 0; A.stride2 = 0;
  A.nnratio = mp->nnratio; A.checkOri = mp->check_orientation; A.window = mp->window;
  A.matches12 = reinterpret_cast<int*>(D + oM12); A.nmatches = reinterpret_cast<int*>(D + oNm);
  A.best = reinterpret_cast<int*>(D + oBest); A.second = reinterpret_cast<int*>(D + oSecond);
  int st;
  if (windowed) {
    A.xy2 = reinterpret_cast<float*>(D + oXY2); A.oct1 = reinterpret_cast<int*>(D + oO1); A.oct2 = reinterpret_cast<int*>(D + oO2);
    A.prev = reinterpret_cast<float*>(D + oPrev);
    A.minX = mp->min_x; A.minY = mp->min_y;
    A.invW = 64.f / (mp->max_x - mp->min_x);  // Frame.cc:184-186
    A.invH = 48.f / (mp->max_y - mp->min_y);
    if (latencyForm) {
      // the largest distance that can change a decision (see k_sfi_scan)
      int dmax = 255;
      for (int d = TH_LOW; d < 256; d++)
        if ((float)(d + 1) * mp->nnratio > (float)TH_LOW) { dmax = d; break; }
      unsigned* keys = reinterpret_cast<unsigned*>(D + oKeys);
      int* cnt = reinterpret_cast<int*>(D + oCnt);
      const size_t smem = (size_t)(n1 + n2) * 4 + (size_t)((n2 + 1) & ~1) * 2 + (size_t)((n1 + 3) & ~3) + (size_t)kSfiChunk * (kSfiK + 1) * 4 + (size_t)kSfiChunk * 2;
      if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many keypoints per frame for the matcher's shared memory");
      ORB_CUDA(raise_dynamic_smem(k_sfi_walk, smem));
      const int rowsPerCta = 8 * kSfiRowsPerWarp;
      k_sfi_scan<<<dim3((n1 + rowsPerCta - 1) / rowsPerCta, 1), 256, 0, s>>>(A, keys, cnt, dmax);
      k_sfi_walk<<<1, 256, smem, s>>>(A, keys, cnt, dmax);
      ORB_CUDA(cudaGetLastError());
      st = ORB_OK;
    } else {
      st = dispatch_match<true>(A, 1, s);
    }
  } else {
    st = dispatch_match<false>(A, 1, s);
  }
  if (st) return st;
  ORB_CUDA(cudaMemcpyAsync(H + outBegin, D + outBegin, outEnd - outBegin, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  memcpy(matches12, H + oM12, (size_t)n1 * 4);
  *nmatches = *reinterpret_cast<const int*>(H + oNm);
  if (best) memcpy(best, H + oBest, (size_t)n1 * 4);
  if (second) memcpy(second, H + oSecond, (size_t)n1 * 4);
  if (windowed) memcpy(prev_matched, H + oPrev, (size_t)n1 * 8);
  return ORB_OK;
}

int orb_match_pairs_device(orb_matcher* m, const uint8_t* d_descriptors, const float* d_angles, int pairs, int n,
                           float nnratio, int check_orientation, int32_t* d_matches12, int32_t* d_nmatches,
                           void* stream) {
  if (!m || !d_descriptors || !d_angles || !d_matches12 || !d_nmatches || pairs <= 0 || n <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(m->device));
  PairArgs A;
  memset(&A, 0, sizeof A);
  A.desc1 = d_descriptors; A.desc2 = d_descriptors + (size_t)n * 32;
  A.ang1 = d_angles; A.ang2 = d_angles + n;
  A.n1 = n; A.n2 = n; A.stride1 = 2LL * n; A.stride2 = 2LL * n;
  A.nnratio = nnratio; A.checkOri = check_orientation;
  A.matches12 = d_matches12; A.nmatches = d_nmatches;
  return dispatch_match<false>(A, pairs, stream ? (cudaStream_t)stream : m->stream);
}

int orb_match_allpairs_device(orb_matcher* m, const uint8_t* d_all, int n_kf, int n_desc, int row_begin, int row_end,
                              int col_begin, int col_end, float nnratio, int32_t* d_counts, void* stream) {
  if (!m || !d_all || !d_counts || n_kf <= 0 || n_desc <= 0 || row_begin < 0 || row_end > n_kf || row_end <= row_begin ||
      col_begin < 0 || col_end > n_kf || col_end <= col_begin)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
  const int rows = row_end - row_begin;
  const size_t smem = (size_t)n_desc * 32;
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "more than 6400 descriptors per keyframe");
  // enough column spans to fill the machine several times over, but long enough to amortise
  // the register load of the row keyframe
  const int ncols = col_end - col_begin;
  int spans = std::max(1, std::min(ncols, (148 * 8 + rows - 1) / rows));
  const int colsPerBlock = (ncols + spans - 1) / spans;
  spans = (ncols + colsPerBlock - 1) / colsPerBlock;
  const int rpt = (n_desc + kApThreads - 1) / kApThreads;
  dim3 grid(spans, rows);
#define ORB_AP(R)                                                                                                  \
  do {                                                                                                             \
    ORB_CUDA(raise_dynamic_smem(k_allpairs<R>, smem));         \
    k_allpairs<R><<<grid, kApThreads, smem, s>>>(d_all, n_kf, n_desc, row_begin, col_begin, col_end, colsPerBlock, nnratio, d_counts); \
  } while (0)
  if (rpt <= 1) ORB_AP(1);
  else if (rpt <= 2) ORB_AP(2);
  else if (rpt <= 4) ORB_AP(4);
  else if (rpt <= 8) ORB_AP(8);
  else ORB_FAIL(ORB_ERR_UNSUPPORTED, "more than 2048 descriptors per keyframe");
#undef ORB_AP
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

// ---- all-pairs with the descriptor exchange over NCCL -------------------------------------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already uses - e.g. the one bundled with
// PyTorch - or the system's), so liborb_b200.so has no link-time dependency on it and single-GPU users never load it.
namespace {
typedef struct { char internal[128]; } orb_nccl_id;   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(orb_nccl_id*) = nullptr;
  int (*CommInitRank)(void**, int, orb_nccl_id, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.error = std::string("libnccl.so.2 cannot be loaded: ") + dlerror(); return; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) { ok = false; api.error = std::string("libnccl misses ") + n; } return p; };
    api.GetUniqueId = (int (*)(orb_nccl_id*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, orb_nccl_id, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclBroadcast");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) { dlclose(api.lib); api.lib = nullptr; }
  });
  return &api;
}
int nccl_fail(NcclApi* a, int rc, const char* what) {
  set_last_error(std::string(what) + " failed: " + (a->GetErrorString ? a->GetErrorString(rc) : "?"));
  return ORB_ERR_CUDA;
}
#define ORB_NCCL(call, what)                          \
  do {                                                \
    const int rc_ = (call);                           \
    if (rc_ != 0) return nccl_fail(api, rc_, what);   \
  } while (0)
constexpr int kNcclUint8 = 1;   // ncclUint8 / ncclChar family: ncclInt8 = 0, ncclUint8 = 1
}  // namespace

void orb_shard_range(int total, int rank, int world, int* begin, int* end) {
  const int base = total / world, rem = total % world;
  *begin = rank * base + std::min(rank, rem);
  *end = *begin + base + (rank < rem ? 1 : 0);
}

int orb_nccl_unique_id(void* id128) {
  if (!id128) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  NcclApi* api = nccl_api();
  if (!api->lib) ORB_FAIL(ORB_ERR_UNSUPPORTED, api->error);
  ORB_NCCL(api->GetUniqueId(reinterpret_cast<orb_nccl_id*>(id128)), "ncclGetUniqueId");
  return ORB_OK;
}

int orb_nccl_comm_create(int device, int rank, int world, const void* id128, void** comm) {
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  NcclApi* api = nccl_api();
  if (!api->lib) ORB_FAIL(ORB_ERR_UNSUPPORTED, api->error);
  ORB_CUDA(cudaSetDevice(device));
  orb_nccl_id id;
  memcpy(&id, id128, sizeof id);
  ORB_NCCL(api->CommInitRank(comm, world, id, rank), "ncclCommInitRank");
  return ORB_OK;
}

int orb_nccl_comm_destroy(void* comm) {
  if (!comm) return ORB_OK;
  NcclApi* api = nccl_api();
  if (!api->lib) ORB_FAIL(ORB_ERR_UNSUPPORTED, api->error);
  ORB_NCCL(api->CommDestroy(comm), "ncclCommDestroy");
  return ORB_OK;
}

int orb_match_allpairs_nccl(orb_matcher* m, void* nccl_comm, int rank, int world, const uint8_t* d_local_desc, int n_kf, int n_desc,
                            float nnratio, uint8_t* d_all, int32_t* d_counts, void* stream) {
  if (!m || !d_local_desc || !d_all || !d_counts || n_kf <= 0 || n_desc <= 0 || world < 1 || rank < 0 || rank >= world || n_kf < world)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  if (world > 1 && !nccl_comm) ORB_FAIL(ORB_ERR_INVALID, "a communicator is needed for more than one rank");
  ORB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
  int rb, re;
  orb_shard_range(n_kf, rank, world, &rb, &re);
  const size_t kfBytes = (size_t)n_desc * 32;
  ORB_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)(re - rb) * n_kf * sizeof(int32_t), s));
  if (d_local_desc != d_all + (size_t)rb * kfBytes)
    ORB_CUDA(cudaMemcpyAsync(d_all + (size_t)rb * kfBytes, d_local_desc, (size_t)(re - rb) * kfBytes, cudaMemcpyDeviceToDevice, s));
  if (world == 1) return orb_match_allpairs_device(m, d_all, n_kf, n_desc, rb, re, 0, n_kf, nnratio, d_counts, s);
  NcclApi* api = nccl_api();
  if (!api->lib) ORB_FAIL(ORB_ERR_UNSUPPORTED, api->error);
  if (!m->commStream) {
    ORB_CUDA(cudaStreamCreateWithFlags(&m->commStream, cudaStreamNonBlocking));
    ORB_CUDA(cudaEventCreateWithFlags(&m->evReady, cudaEventDisableTiming));
  }
  while ((int)m->evBlock.size() < world) {
    cudaEvent_t ev;
    ORB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    m->evBlock.push_back(ev);
  }
  // the exchange starts once this rank's block sits in d_all; every rank issues the broadcasts in the same order
  // (root 0, 1, ...), each followed by an event, on the communication stream
  ORB_CUDA(cudaEventRecord(m->evReady, s));
  ORB_CUDA(cudaStreamWaitEvent(m->commStream, m->evReady, 0));
  for (int r = 0; r < world; r++) {
    int a, b;
    orb_shard_range(n_kf, r, world, &a, &b);
    uint8_t* blk = d_all + (size_t)a * kfBytes;
    ORB_NCCL(api->Broadcast(blk, blk, (size_t)(b - a) * kfBytes, kNcclUint8, r, nccl_comm, m->commStream), "ncclBroadcast");
    ORB_CUDA(cudaEventRecord(m->evBlock[r], m->commStream));
  }
  // compute: the own block at once (its broadcast only reads it), then the peers' blocks as they land, k_allpairs on the
  // compute stream overlapping the remaining transfers
  int st = orb_match_allpairs_device(m, d_all, n_kf, n_desc, rb, re, rb, re, nnratio, d_counts, s);
  if (st) return st;
  for (int k = 1; k < world; k++) {
    const int r = (rank + k) % world;
    int a, b;
    orb_shard_range(n_kf, r, world, &a, &b);
    ORB_CUDA(cudaStreamWaitEvent(s, m->evBlock[r], 0));
    st = orb_match_allpairs_device(m, d_all, n_kf, n_desc, rb, re, a, b, nnratio, d_counts, s);
    if (st) return st;
  }
  ORB_CUDA(cudaStreamWaitEvent(s, m->evBlock[rank], 0));   // d_all may be reused once `s` has passed this point
  return ORB_OK;
}

int orb_hamming_matrix_device(orb_matcher* m, const uint8_t* d_a, int na, const uint8_t* d_b, int nb, int32_t* d_out,
                              void* stream) {
  if (!m || !d_a || !d_b || !d_out || na <= 0 || nb <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(m->device));
  k_hamming_matrix<<<dim3((nb + 127) / 128, na), 128, 0, stream ? (cudaStream_t)stream : m->stream>>>(d_a, na, d_b, nb, d_out);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_matcher_synchronize(orb_matcher* m, void* stream) {
  if (!m) ORB_FAIL(ORB_ERR_INVALID, "null handle");
  ORB_CUDA(cudaSetDevice(m->device));
  ORB_CUDA(cudaStreamSynchronize(stream ? (cudaStream_t)stream : m->stream));
  return ORB_OK;
}

int orb_int_pipe_peak(int device, int what, double* ops_per_s) {
  if (!ops_per_s) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  ORB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ORB_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  unsigned* d_out = nullptr;
  ORB_CUDA(cudaMalloc(&d_out, (size_t)blocks * 256 * 4));
  cudaEvent_t e0, e1;
  ORB_CUDA(cudaEventCreate(&e0));
  ORB_CUDA(cudaEventCreate(&e1));
  float bestMs = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    ORB_CUDA(cudaEventRecord(e0));
    if (what == 0) k_int_pipe<0><<<blocks, 256>>>(d_out, iters, 0x9e3779b9u + rep);
    else if (what == 1) k_int_pipe<1><<<blocks, 256>>>(d_out, iters, 0x9e3779b9u + rep);
    else k_int_pipe<2><<<blocks, 256>>>(d_out, iters, 0x9e3779b9u + rep);
    ORB_CUDA(cudaEventRecord(e1));
    ORB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ORB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) bestMs = std::min(bestMs, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  // per loop iteration: 8 unrolled x 8 chains of the measured op (plus one dependent add/xor each,
  // which issues on the other pipe for POPC and is the same op class for LOP3: counted once)
  // what = 2 reports "mix units" per second: one unit = 1 POPC + 4 LOP3 (16 units per loop iteration)
  const double ops = (double)blocks * 256.0 * iters * (what == 2 ? 16.0 : 64.0);
  *ops_per_s = ops / (bestMs * 1e-3);
  return ORB_OK;
}

}  // extern "C"
