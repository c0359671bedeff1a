// orb_extract.cu — B200 (sm_100a) implementation of ORBextractor::operator()
// (reference: src/ORBextractor.cc:1533-1649 and the functions it calls).
//
// Data layout in HBM, per frame of a chunk (all offsets in LevelGeom):
//   pyramid slab : nlevels bordered u8 planes, row pitch multiple of 32 B, interior column 0
//                  at byte 32 of a row (16 B aligned), 19 border rows/cols of REFLECT_101 content
//   blur slab    : nlevels compact u8 planes (pitch multiple of 32 B)
//   candidates   : per level a uint2 list {x | y<<16, score} (unordered; order is recovered
//                  from (x,y) because the reference's candidate order is a function of position)
//   kept         : per level a uint2 list in the reference's final list order
// Kernels (one launch each over a whole chunk of frames):
//   k_level0_border, k_resize_border (x nlevels-1)  ComputePyramid      :1655-1724
//   k_fast_cells                                    ComputeKeyPointsOctTree + cv::FAST :1037-1142
//   k_quadtree                                      DistributeOctTree   :688-1033
//   k_blur7                                         cv::GaussianBlur    :1607-1615
//   k_describe                                      IC_Angle :94-141, computeOrbDescriptor :153-204,
//                                                   level->image scaling :1633-1642
//   k_stereo (stereo entry points only)             Frame::ComputeStereoMatches, src/Frame.cc:831-1082
// A batch call cuts the batch into chunks of max_batch frames (optionally alternating between two
// workspace lanes / streams, see orb_set_lanes).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <cuda.h>
#include <cuda_pipeline.h>

#include "../../include/orb_pattern_data.h"
#include "orb_common.cuh"

namespace orbb200 {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

cudaError_t raise_dynamic_smem_impl(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;   // (device, kernel) -> largest size set so far
  if (bytes <= 48 * 1024) return cudaSuccess;                     // always launchable
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = granted[std::make_pair(dev, func)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}
}  // namespace orbb200

using namespace orbb200;

namespace {

constexpr int kEdge = 19;       // EDGE_THRESHOLD, ORBextractor.cc:81
constexpr int kLeftPad = 32;    // bytes left of interior column 0 in a bordered row (>= kEdge)
constexpr int kMaxLevels = 16;
constexpr int kHalfPatch = 15;  // HALF_PATCH_SIZE, ORBextractor.cc:80
constexpr int kQtThreads = 256;      // k_quadtree: threads per CTA for batches ...
constexpr int kQtMaxThreads = 1024;  // ... and for the few-frames case (one CTA per level: more threads shorten every phase)
constexpr int kQtSmemKeys = 2048;   // candidates of a level cached in shared memory by k_quadtree (4096 cost a resident CTA per SM on KITTI: 3 x 69 KB instead of 4 x 56 KB, 0.99 vs 0.85 ms; levels with more candidates read them from global memory)

struct LevelGeom {
  int w, h;               // interior size
  int pitch;              // bordered plane row pitch
  long long off;          // byte offset of interior (0,0) inside the frame's pyramid slab
  int bpitch;             // blurred plane pitch
  long long boff;         // byte offset of the blurred plane inside the frame's blur slab
  int wCell, hCell;       // FAST cell size (ORBextractor.cc:1070-1071)
  int nColsAll;           // nCols of the reference (row stride of the candidate order)
  int nCols, nRows;       // cells that survive the skip rules (:1083, :1101)
  int cellBase;           // first block of this level in the k_fast_cells grid
  int nfeat;              // mnFeaturesPerLevel[level]
  int nIni;               // quadtree roots (:695)
  float hX;               // root width (:697)
  int candOff, candCap;   // candidate list slot inside the per-frame candidate array
  int keptOff, keptCap;
  int tapX, tapY;         // offsets of this level's resize taps in the tap table
  int blurTileBase, blurTilesX, blurTilesY;
  unsigned blurTilesXMagic;   // ceil(2^32 / blurTilesX): tile row = umulhi(tile, magic)
  int borderTileBase, borderTilesX, borderTilesY;
  float scale;            // mvScaleFactor[level]
  float patch;            // keypoint size = int(31*scale) (:1164)
};

struct Geom {
  int nlevels, W, H;
  int iniTh, minTh;
  int totalCells, totalBlurTiles;
  LevelGeom lv[kMaxLevels];
};

__constant__ int c_umax[kHalfPatch + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// ------------------------------------------------------------------------------------------
// Pyramid
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
  p = p < 0 ? -p : p;
  return p >= n ? 2 * (n - 1) - p : p;
}

// Level 0: copyMakeBorder(image, 19, REFLECT_101)  (ORBextractor.cc:1716)
// One thread writes 16 bytes (one aligned STG.128) of a bordered row. Byte 16*g of a row is
// interior column 16*(g-2); columns left of -19 / right of w+18 are row padding.
__global__ void __launch_bounds__(256) k_level0_border(const Geom g, const u8* __restrict__ img, size_t step,
                                                       size_t frameStride, u8* __restrict__ pyr,
                                                       size_t pyrStride) {
  const LevelGeom& L = g.lv[0];
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  const int by = blockIdx.y * blockDim.y + threadIdx.y;
  const int f = blockIdx.z;
  if (gi * 16 >= L.pitch || by >= L.h + 2 * kEdge) return;
  const int c0 = 16 * (gi - 2);
  const int y = reflect101(by - kEdge, L.h);
  const u8* row = img + (size_t)f * frameStride + (size_t)y * step;
  uint4 out;
  if (c0 >= 16 && c0 + 20 <= L.w) {
    // interior: 5 aligned words + funnel shifts (source rows have arbitrary alignment)
    const size_t addr = reinterpret_cast<size_t>(row + c0);
    const unsigned* wp = reinterpret_cast<const unsigned*>(addr & ~(size_t)3);
    const unsigned sh = (unsigned)(addr & 3) * 8;
    const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
    out.x = __funnelshift_r(w0, w1, sh);
    out.y = __funnelshift_r(w1, w2, sh);
    out.z = __funnelshift_r(w2, w3, sh);
    out.w = __funnelshift_r(w3, w4, sh);
  } else {
    unsigned v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      unsigned acc = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int x = min(max(reflect101(c0 + 4 * q + j, L.w), 0), L.w - 1);
        acc |= (unsigned)__ldg(row + x) << (8 * j);
      }
      v[q] = acc;
    }
    out = make_uint4(v[0], v[1], v[2], v[3]);
  }
  u8* dst = pyr + (size_t)f * pyrStride + L.off + (long long)(by - kEdge) * L.pitch - kLeftPad + 16 * gi;
  *reinterpret_cast<uint4*>(dst) = out;
}

// Level l>0: resize(level l-1 -> l, INTER_LINEAR) fused with copyMakeBorder(REFLECT_101 |
// ISOLATED) (ORBextractor.cc:1677, :1695). Every thread of the BORDERED domain recomputes the
// interior pixels it mirrors, so the plane is written exactly once, 4 pixels (one aligned
// STG.32) per thread. taps: {source index, c0 | c1<<16} per destination column / row
// (11-bit fixed point, cv::resize INTER_LINEAR 8-bit).
__device__ __forceinline__ unsigned resize_px_slow(const u8* r0, const u8* r1, int2 tx, int sw, int cy0, int cy1) {
  const int sx0 = tx.x, sx1 = min(sx0 + 1, sw - 1);
  const int cx0 = tx.y & 0xffff, cx1 = tx.y >> 16;
  const int h0 = r0[sx0] * cx0 + r0[sx1] * cx1;
  const int h1 = r1[sx0] * cx0 + r1[sx1] * cx1;
  const int v = (((cy0 * (h0 >> 4)) >> 16) + ((cy1 * (h1 >> 4)) >> 16) + 2) >> 2;
  return (unsigned)min(max(v, 0), 255);
}

// one 4-pixel word of bordered row (interior row index y already reflected)
// tightEnd: the source is a caller's image and this is its last frame - the 12-byte windows of the fast paths must not
// run past the end of the last row (the pyramid planes have 19 px of border and slack after every row instead).
__device__ __forceinline__ unsigned resize_row_word(const LevelGeom& D, const LevelGeom& S, const u8* src, long long spitch,
                                                    const int2* __restrict__ taps, int c0, int y, bool tightEnd) {
  const int2 ty = __ldg(taps + D.tapY + y);
  const int sy0 = ty.x, sy1 = min(sy0 + 1, S.h - 1);
  const int cy0 = ty.y & 0xffff, cy1 = ty.y >> 16;
  const unsigned cy0s = (unsigned)cy0 << 16, cy1s = (unsigned)cy1 << 16;
  const u8* r0 = src + (long long)sy0 * spitch;
  const u8* r1 = src + (long long)sy1 * spitch;
  unsigned out = 0;
  bool fast = c0 >= 0 && c0 + 3 < D.w;
  int4 ta, tb;
  if (fast) {
    const int4* tp4 = reinterpret_cast<const int4*>(taps + D.tapX + c0);  // tapX and c0 are even: 16 B aligned
    ta = __ldg(tp4);
    tb = __ldg(tp4 + 1);
    fast = tb.z + 1 - ta.x <= 7;  // the 4 pixels' source span fits one 8-byte window
    if (tightEnd && sy1 == S.h - 1 && ta.x + 12 > S.w) fast = false;
  }
  if (fast) {
    const int a = ta.x;
    const size_t a0 = reinterpret_cast<size_t>(r0 + a), a1 = reinterpret_cast<size_t>(r1 + a);
    const unsigned* p0 = reinterpret_cast<const unsigned*>(a0 & ~(size_t)3);
    const unsigned* p1 = reinterpret_cast<const unsigned*>(a1 & ~(size_t)3);
    const unsigned s0 = (unsigned)(a0 & 3) * 8, s1 = (unsigned)(a1 & 3) * 8;
    const unsigned u0 = p0[0], u1 = p0[1], u2 = p0[2], v0 = p1[0], v1 = p1[1], v2 = p1[2];
    const unsigned lo0 = __funnelshift_r(u0, u1, s0), hi0 = __funnelshift_r(u1, u2, s0);
    const unsigned lo1 = __funnelshift_r(v0, v1, s1), hi1 = __funnelshift_r(v1, v2, s1);
    const int sx[4] = {ta.x, ta.z, tb.x, tb.z};
    const unsigned cf[4] = {(unsigned)ta.y, (unsigned)ta.w, (unsigned)tb.y, (unsigned)tb.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned sel = (unsigned)(sx[j] - a) * 0x11u + 0x10u;   // bytes (d, d+1) of the window
      const unsigned q0 = __byte_perm(lo0, hi0, sel), q1 = __byte_perm(lo1, hi1, sel);
      const unsigned h0 = __dp2a_lo(cf[j], q0, 0u);         // c0*p[sx] + c1*p[sx+1]  (<= 255*2048)
      const unsigned h1 = __dp2a_lo(cf[j], q1, 0u);
      // ((cy*(h>>4))>>16) == umulhi(cy<<16, h>>4); the sum is <= 1020, so the result needs no clamp
      const unsigned v = (__umulhi(cy0s, h0 >> 4) + __umulhi(cy1s, h1 >> 4) + 2u) >> 2;
      out |= v << (8 * j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int x = min(max(reflect101(c0 + j, D.w), 0), D.w - 1);
      out |= resize_px_slow(r0, r1, __ldg(taps + D.tapX + x), S.w, cy0, cy1) << (8 * j);
    }
  }
  return out;
}

// 4 consecutive bordered positions p0..p0+3 (interior coordinates, may be negative or >= n) map,
// under REFLECT_101, into a window of 4 consecutive interior positions [lo, lo+3]: ascending
// (inside), descending (in the border) or, where the group straddles the far turning point, a
// mix. perm holds, per position j, the index 0..3 inside the window (nibble j) - directly a PRMT
// selector for the 4 output bytes. Returns false where no such window exists.
__device__ __forceinline__ bool reflect_group(int p0, int n, int& lo, unsigned& perm) {
  if (p0 >= 0 && p0 + 3 < n) { lo = p0; perm = 0x3210u; return true; }
  if (p0 + 3 <= 0) { lo = -p0 - 3; perm = 0x0123u; return -p0 < n; }                  // p = 0 mirrors onto itself
  if (p0 >= n - 1) { lo = 2 * (n - 1) - p0 - 3; perm = 0x0123u; return lo >= 0; }     // so does p = n-1
  if (p0 < 0 || n < 4) return false;
  lo = n - 4;                                                                         // p0 < n-1 < p0+3
  perm = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) perm |= (unsigned)(reflect101(p0 + j, n) - lo) << (4 * j);
  return true;
}

// Thread = 4 columns x 4 rows of the bordered plane. Blocks that do not straddle a reflection
// turning point share the horizontal pass: the 4 output rows read at most 6 consecutive source
// rows (scale <= 4/3), each filtered once; border blocks are the mirrored copy of such a block.
// ext != nullptr: the source level is read from there (frame stride extStride, row pitch extPitch) instead of the
// pyramid slab - level 1 can be made from the input image itself, so that it does not wait for k_level0_border.
__global__ void __launch_bounds__(256) k_resize_border(const Geom g, int l, u8* __restrict__ pyr, size_t pyrStride,
                                                       const int2* __restrict__ taps, const u8* __restrict__ ext,
                                                       size_t extStride, int extPitch) {
  const LevelGeom& D = g.lv[l];
  const LevelGeom& S = g.lv[l - 1];
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  const int by0 = 4 * (blockIdx.y * blockDim.y + threadIdx.y);
  const int f = blockIdx.z;
  if (gi * 4 >= D.pitch || by0 >= D.h + 2 * kEdge) return;
  const int c0 = 4 * gi - kLeftPad;
  const u8* src = ext ? ext + (size_t)f * extStride : pyr + (size_t)f * pyrStride + S.off;
  const long long spitch = ext ? extPitch : S.pitch;
  const bool tightEnd = ext != nullptr && f == (int)gridDim.z - 1;
  u8* dst = pyr + (size_t)f * pyrStride + D.off + (long long)(by0 - kEdge) * D.pitch + c0;
  const int nrows = min(4, D.h + 2 * kEdge - by0);
  int xlo = 0, ylo = 0;
  unsigned permX = 0x3210u, permY = 0x3210u;
  bool blockFast = nrows == 4 && reflect_group(c0, D.w, xlo, permX) && reflect_group(by0 - kEdge, D.h, ylo, permY);
  int2 tx[4], ty[4];
  if (blockFast) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      tx[k] = __ldg(taps + D.tapX + xlo + k);
      ty[k] = __ldg(taps + D.tapY + ylo + k);
    }
    // the 4 pixels' source span must fit one 8-byte window; source rows advance by 1 or 2 per
    // output row and the block must fit rows base .. base+5
    blockFast = tx[3].x + 1 - tx[0].x <= 7 && ty[1].x - ty[0].x >= 1 && ty[2].x - ty[1].x >= 1 && ty[3].x - ty[2].x >= 1 &&
                ty[3].x - ty[0].x <= 4;
    if (tightEnd && ty[0].x + 5 >= S.h - 1 && tx[0].x + 12 > S.w) blockFast = false;
  }
  if (blockFast) {
    // The integer ALU pipe (shifts, selects, logic, byte permutes) bounds this kernel: the choice of the two source rows
    // of an output row - uniform over the warp, which shares by0 - is a branch instead of 8 selects, and the 4 results
    // are packed with 3 byte permutes. (Doing the right shifts as high multiplies on the FMA pipe measured 8 % slower.)
    const int a = tx[0].x, base = ty[0].x;
    unsigned sel[4];
#pragma unroll
    for (int j = 0; j < 4; j++) sel[j] = (unsigned)(tx[j].x - a) * 0x11u + 0x10u;   // bytes (d, d+1) of the window
    unsigned hq[6][4];   // horizontal pass of source rows base..base+5, already >> 4
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const u8* q = src + (long long)min(base + i, S.h - 1) * spitch + a;
      const unsigned mis = (unsigned)(reinterpret_cast<size_t>(q) & 3);
      const unsigned* p = reinterpret_cast<const unsigned*>(q - mis);
      const unsigned u0 = p[0], u1 = p[1], u2 = p[2];
      const unsigned lo = __funnelshift_r(u0, u1, mis * 8), hi = __funnelshift_r(u1, u2, mis * 8);
#pragma unroll
      for (int j = 0; j < 4; j++) hq[i][j] = __dp2a_lo((unsigned)tx[j].y, __byte_perm(lo, hi, sel[j]), 0u) >> 4;
    }
    unsigned orow[4];   // window rows ylo .. ylo+3, bytes already in bordered column order
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned cy0s = (unsigned)(ty[k].y & 0xffff) << 16, cy1s = (unsigned)(ty[k].y >> 16) << 16;
      unsigned v[4];
      // ((cy*(h>>4))>>16) == umulhi(cy<<16, h>>4); the sum is <= 1020, so the result needs no clamp
      if (ty[k].x - base > k) {                    // source rows base+k+1, base+k+2 (warp-uniform)
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (__umulhi(cy0s, hq[k + 1][j]) + __umulhi(cy1s, hq[k + 2][j]) + 2u) >> 2;
      } else {                                     // source rows base+k, base+k+1
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (__umulhi(cy0s, hq[k][j]) + __umulhi(cy1s, hq[k + 1][j]) + 2u) >> 2;
      }
      const unsigned out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
      orow[k] = __byte_perm(out, 0u, permX);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned sel = (permY >> (4 * k)) & 3u;   // which window row lands on bordered row by0+k
      const unsigned v = sel == 0 ? orow[0] : (sel == 1 ? orow[1] : (sel == 2 ? orow[2] : orow[3]));
      *reinterpret_cast<unsigned*>(dst + (long long)k * D.pitch) = v;
    }
  } else {
    for (int k = 0; k < nrows; k++) {
      const int y = reflect101(by0 + k - kEdge, D.h);
      *reinterpret_cast<unsigned*>(dst + (long long)k * D.pitch) = resize_row_word(D, S, src, spitch, taps, c0, y, tightEnd);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Pyramid, second form (the default): no divergent edge paths inside a warp, and the resize as a separable
// shared-memory tile kernel.
// ------------------------------------------------------------------------------------------
// Level 0 (copyMakeBorder :1716). The 16-byte groups of a bordered row are of two kinds: interior groups (an aligned
// window of the source row, 5 word loads + funnel shifts) and the few groups at either end that hold mirrored or padding
// bytes (byte loads). Mixed in one warp the long edge path used to run beside 30 idle lanes in two warps of three
// (11.8 of 32 lanes active); here a warp does ONE kind: blocks [0, nbInt) copy interior groups, one warp per bordered
// row, the remaining blocks do the edge groups, 4 rows x 8 groups per warp.
// one warp: the interior groups of bordered row `by`
__device__ __forceinline__ void level0_interior_row(const LevelGeom& L, const u8* src, size_t step, u8* dstBase, int by, int lane) {
  const int gA = 3, gB = max(gA, (L.w - 20) / 16 + 3);          // interior groups: 16 <= c0 and c0 + 20 <= w
  const u8* row = src + (size_t)reflect101(by - kEdge, L.h) * step;
  u8* dst = dstBase + (long long)by * L.pitch;
  // the row's groups in rounds of 4 x 32: all loads of a round are issued before its first store (a warp lives for one row:
  // with a load - store chain per group it spent its life waiting, long-scoreboard stalls 22 warps per issue slot)
  const unsigned sh = (unsigned)(reinterpret_cast<size_t>(row) & 3) * 8;   // 16 * (gi - 2) keeps the alignment of `row`
  const unsigned* wrow = reinterpret_cast<const unsigned*>(reinterpret_cast<size_t>(row) & ~(size_t)3);
  for (int g0 = gA + lane; g0 < gB; g0 += 128) {
    unsigned w[4][5];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int gi = g0 + 32 * k;
      if (gi < gB) {
        const unsigned* wp = wrow + 4 * (gi - 2);
#pragma unroll
        for (int j = 0; j < 5; j++) w[k][j] = __ldg(wp + j);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int gi = g0 + 32 * k;
      if (gi < gB) {
        uint4 out;
        out.x = __funnelshift_r(w[k][0], w[k][1], sh);
        out.y = __funnelshift_r(w[k][1], w[k][2], sh);
        out.z = __funnelshift_r(w[k][2], w[k][3], sh);
        out.w = __funnelshift_r(w[k][3], w[k][4], sh);
        *reinterpret_cast<uint4*>(dst + 16 * gi) = out;
      }
    }
  }
}
// one warp: the edge groups (<= 8) of the 4 bordered rows by4 .. by4+3
__device__ __forceinline__ void level0_edge_rows(const LevelGeom& L, const u8* src, size_t step, u8* dstBase, int by4, int lane) {
  const int rows = L.h + 2 * kEdge, groups = L.pitch >> 4;
  const int gA = 3, gB = max(gA, (L.w - 20) / 16 + 3);
  const int nEdge = gA + (groups - gB);                          // <= 8 for every supported width (checked by the host)
  const int by = by4 + (lane >> 3);
  const int k = lane & 7;
  if (by >= rows || k >= nEdge) return;
  const int gi = k < gA ? k : gB + (k - gA);
  const int c0 = 16 * (gi - 2);
  const u8* row = src + (size_t)reflect101(by - kEdge, L.h) * step;
  unsigned v[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int x = min(max(reflect101(c0 + 4 * q + j, L.w), 0), L.w - 1);
      acc |= (unsigned)__ldg(row + x) << (8 * j);
    }
    v[q] = acc;
  }
  *reinterpret_cast<uint4*>(dstBase + (long long)by * L.pitch + 16 * gi) = make_uint4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256) k_level0_border2(const Geom g, const u8* __restrict__ img, size_t step,
                                                        size_t frameStride, u8* __restrict__ pyr, size_t pyrStride,
                                                        int nbInt) {
  const LevelGeom& L = g.lv[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.y;
  const u8* src = img + (size_t)f * frameStride;
  u8* dstBase = pyr + (size_t)f * pyrStride + L.off - (long long)kEdge * L.pitch - kLeftPad;
  if ((int)blockIdx.x < nbInt) {
    const int by = blockIdx.x * 8 + wid;
    if (by < L.h + 2 * kEdge) level0_interior_row(L, src, step, dstBase, by, lane);
  } else {
    level0_edge_rows(L, src, step, dstBase, (((int)blockIdx.x - nbInt) * 8 + wid) * 4, lane);
  }
}

// Level l > 0, interior pixels (resize :1677): a thread owns 4 output columns - their taps, the byte selectors and the
// shift of its 8-byte source window stay in registers - and walks DOWN the source rows of a band of output rows. Every
// source row is filtered horizontally exactly once per band ((c0*p[sx] + c1*p[sx+1]) >> 4: 3 aligned words, 2 funnel
// shifts, PRMT + IDP.2A + shift per pixel), the previous row's result stays in registers, and an output row is emitted
// when its second source row has arrived (2 IMAD.HI + add + shift per pixel, 3 PRMT pack the word, one aligned STG.32).
// The next row's words are requested before the current one is used. No shared memory, no barrier, no divergence: all
// control flow depends on the row only. (The first-round kernel re-filtered 6 source rows per 4x4 block at 20-24 of 32
// lanes: ~46 thread instructions per pixel; a shared-memory tile version measured slower still - its per-tile set-up
// and store pass are amortised over only 8 pixels per thread.) The border is written afterwards by k_fill_borders.
// Per band the warp keeps the rows' vertical taps in registers (lane i: output row j0+i) and a 64-bit mask of the source
// rows after which an output row is due, so the row loop has a fixed trip count, one predictable branch and no global
// loads besides the pixels. Output rows use strictly increasing source rows (scale factor > 1; checked by the host).
constexpr int kRzThreads = 64, kRzBand = 32;
// One warp: 32 column groups starting at column cFirst, output rows [j0, j1) (at most 32) of level l of one frame
// (`frame` = the frame's pyramid slab). CG: source pixels are read with ld.global.cg - the single-launch pyramid reads
// rows another CTA of the same launch has just written, which must not be served from a stale L1 line.
template <bool CG, bool WAIT = false>
__device__ __forceinline__ void resize_strip_warp(const Geom& g, int l, u8* __restrict__ frame, const int2* __restrict__ taps,
                                                  int cFirst, int j0, int j1, int lane) {
  const LevelGeom& D = g.lv[l];
  const LevelGeom& S = g.lv[l - 1];
  // column groups cover the BORDERED width: columns -20 .. w+18 in aligned words; a border column takes the taps of the
  // interior column it mirrors (REFLECT_101), so the left / right border of an interior row costs ~10 extra groups per row
  const int cLast = ((D.w + kEdge - 1 + 20) & ~3) - 20;                     // first column of the last group
  if (cFirst > cLast || j0 >= j1) return;                                    // warp-uniform
  const int cMine = cFirst + 4 * lane;
  const bool active = cMine <= cLast;
  const int c0 = active ? cMine : cLast;                                    // idle lanes shadow the last column group
  int2 tx[4];
#pragma unroll
  for (int j = 0; j < 4; j++) tx[j] = __ldg(taps + D.tapX + reflect101(min(max(c0 + j, -kEdge), D.w + kEdge - 1), D.w));
  const int sxMin = min(min(tx[0].x, tx[1].x), min(tx[2].x, tx[3].x));
  const int a0 = kLeftPad + sxMin;                                 // byte offset of the window inside a bordered source row
  const unsigned sh = (unsigned)(a0 & 3) * 8;
  unsigned sel[4];
#pragma unroll
  for (int j = 0; j < 4; j++) sel[j] = (unsigned)(tx[j].x - sxMin) * 0x11u + 0x10u;   // bytes (d, d+1) of the window
  // vertical taps of the band: lane i holds row j0+i
  const bool rowValid = j0 + lane < j1;
  const int2 tyMine = __ldg(taps + D.tapY + min(j0 + lane, j1 - 1));
  const int r0 = __shfl_sync(0xffffffffu, tyMine.x, 0);            // first source row of the band
  const int ds = tyMine.x - r0;                                    // < 64 for bands of <= 32 rows (scale factor < 2)
  const unsigned mlo = __reduce_or_sync(0xffffffffu, rowValid && ds < 32 ? 1u << ds : 0u);
  const unsigned mhi = __reduce_or_sync(0xffffffffu, rowValid && ds >= 32 ? 1u << (ds - 32) : 0u);
  unsigned long long due = (unsigned long long)mlo | ((unsigned long long)mhi << 32);   // bit t: a row starts at r0+t
  const int nIter = __shfl_sync(0xffffffffu, ds, j1 - 1 - j0) + 1; // source rows r0+1 .. r0+nIter are filtered in the loop
  const int spw = S.pitch >> 2, dpw = D.pitch >> 2;
  const int lastRow = S.h - 1;
  const unsigned* q = reinterpret_cast<const unsigned*>(frame + S.off - kLeftPad) + (a0 >> 2) + (long long)r0 * spw;
  unsigned* dp = reinterpret_cast<unsigned*>(frame + D.off + (long long)j0 * D.pitch + c0);
  auto ld = [](const unsigned* p) { return CG ? __ldcg(p) : *p; };

  // programmatic dependent launch (k_resize_strip_pdl): everything above read only the tap tables; the pixels of the level
  // below are complete once the launch before this one has finished
  if (WAIT) asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned hc[4], hn[4];
  {
    const unsigned q0 = ld(q), q1 = ld(q + 1), q2 = ld(q + 2);
    const unsigned lo = __funnelshift_r(q0, q1, sh), hi = __funnelshift_r(q1, q2, sh);
#pragma unroll
    for (int j = 0; j < 4; j++) hc[j] = __dp2a_lo((unsigned)tx[j].y, __byte_perm(lo, hi, sel[j]), 0u) >> 4;
  }
  int r = r0;                                                      // source row held in hc
  if (r < lastRow) q += spw;                                       // the bottom row is its own successor (:A.2 clamp)
  unsigned u0 = ld(q), u1 = ld(q + 1), u2 = ld(q + 2);
  int jj = 0;
#pragma unroll 2
  for (int t = 0; t < nIter; t++) {
    {
      const unsigned lo = __funnelshift_r(u0, u1, sh), hi = __funnelshift_r(u1, u2, sh);
#pragma unroll
      for (int j = 0; j < 4; j++) hn[j] = __dp2a_lo((unsigned)tx[j].y, __byte_perm(lo, hi, sel[j]), 0u) >> 4;
    }
    r++;
    if (r < lastRow) q += spw;                                     // request the row after it now
    u0 = ld(q); u1 = ld(q + 1); u2 = ld(q + 2);
    if ((unsigned)due & 1u) {                                      // an output row uses (hc, hn); warp-uniform
      const unsigned cy = (unsigned)__shfl_sync(0xffffffffu, tyMine.y, jj);
      // (((cy0*h0)>>16) + ((cy1*h1)>>16) + 2) >> 2 with plain 32-bit multiplies (cy*h < 2^27; IMAD.HI is slow): the +2 rides
      // on the first product as 2<<16, a 16x2 SIMD add sums the two upper halves without the carry of the lower halves,
      // and the result (<= 255, no clamp needed) sits at bit 18
      const unsigned cy0 = cy & 0xffffu, cy1 = cy >> 16;
      unsigned v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = __vadd2(cy0 * hc[k] + 0x20000u, cy1 * hn[k]) >> 18;
      if (active) *dp = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
      dp += dpw;
      jj++;
    }
    due >>= 1;
#pragma unroll
    for (int k = 0; k < 4; k++) hc[k] = hn[k];
  }
}

__global__ void __launch_bounds__(kRzThreads) k_resize_strip(const Geom g, int l, u8* __restrict__ pyr, size_t pyrStride,
                                                             const int2* __restrict__ taps, int bandRows) {
  const int j0 = blockIdx.y * bandRows;
  resize_strip_warp<false>(g, l, pyr + (size_t)blockIdx.z * pyrStride, taps, 4 * (blockIdx.x * kRzThreads + (threadIdx.x & ~31)) - 20, j0,
                           min(j0 + bandRows, g.lv[l].h), threadIdx.x & 31);
}

// The same kernel for the per-frame call, where a level is a few microseconds of work and the chain of eight dependent
// launches is the long pole: launched with programmatic stream serialization, it lets the next level's CTAs start and set
// up their taps while this one is still running; they wait for this grid to finish before they touch its pixels.
__global__ void __launch_bounds__(kRzThreads) k_resize_strip_pdl(const Geom g, int l, u8* __restrict__ pyr, size_t pyrStride,
                                                                 const int2* __restrict__ taps, int bandRows) {
  asm volatile("griddepcontrol.launch_dependents;");
  const int j0 = blockIdx.y * bandRows;
  resize_strip_warp<false, true>(g, l, pyr + (size_t)blockIdx.z * pyrStride, taps, 4 * (blockIdx.x * kRzThreads + (threadIdx.x & ~31)) - 20,
                                 j0, min(j0 + bandRows, g.lv[l].h), threadIdx.x & 31);
}

// Top / bottom border rows of levels 1.. (copyMakeBorder(REFLECT_101) :1695): the strips have written every interior row
// over the whole bordered width, so a border row is an aligned 16-byte copy of the interior row it mirrors. One launch for
// all levels, one warp per row.
struct BorderJobs { int base[kMaxLevels + 1]; };   // first block of every level (5 blocks of 8 rows each)

// one warp: border row k (0..18 top, 19..37 bottom) of level l of one frame
template <bool CG>
__device__ __forceinline__ void border_row_copy(const LevelGeom& L, u8* frame, int k, int lane) {
  if (k >= 2 * kEdge) return;
  const int by = k < kEdge ? k : L.h + k;                         // bordered row
  u8* plane = frame + L.off - kLeftPad;                           // byte 0 of interior row 0
  const uint4* srow = reinterpret_cast<const uint4*>(plane + (long long)reflect101(by - kEdge, L.h) * L.pitch);
  uint4* drow = reinterpret_cast<uint4*>(plane + (long long)(by - kEdge) * L.pitch);
  for (int gi = lane; gi < (L.pitch >> 4); gi += 32) drow[gi] = CG ? __ldcg(srow + gi) : srow[gi];
}

__global__ void __launch_bounds__(256) k_fill_borders(const Geom g, u8* __restrict__ pyr, size_t pyrStride, const BorderJobs jobs) {
  int l = 1;
#pragma unroll 1
  while (l + 1 < g.nlevels && (int)blockIdx.x >= jobs.base[l + 1]) l++;
  border_row_copy<false>(g.lv[l], pyr + (size_t)blockIdx.y * pyrStride, (blockIdx.x - jobs.base[l]) * 8 + (threadIdx.x >> 5),
                         threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------
// The whole pyramid of a few frames in ONE launch (the per-frame drop-in call: 9 dependent launches cost ~45 us of
// launch latency for ~10 us of work). Persistent CTAs draw work items from a counter, in an order in which every item's
// producers come earlier: level-0 bands (8 bordered rows), then per level bands of 4 output rows, then the top / bottom
// border rows. An item waits (one thread polls, acquire) until the bands of the level below that hold its source rows
// are flagged done (release after a __threadfence of every writer); the waited-for items were drawn earlier, by CTAs that
// are running and never wait on later items, so the scheme cannot deadlock whatever the number of resident CTAs.
// ------------------------------------------------------------------------------------------
struct PyrItem { short kind, level; int band; int depLevel, depFirst, depLast; };   // kind 0 level-0 band, 1 strip band, 2 border rows
struct PyrPlan { int nItems; int bandBase[kMaxLevels + 1]; };                       // flags of level l start at bandBase[l]
constexpr int kPyrFusedBand = 4;    // output rows per strip item
constexpr int kPyrFusedMaxFrames = 4;

__global__ void __launch_bounds__(256) k_pyramid_fused(const Geom g, const u8* __restrict__ img, size_t step, size_t frameStride,
                                                       u8* __restrict__ pyr, size_t pyrStride, const int2* __restrict__ taps,
                                                       const PyrItem* __restrict__ items, const PyrPlan plan, int B,
                                                       int* __restrict__ flags, int* __restrict__ counter) {
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int totalBands = plan.bandBase[g.nlevels];
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    const int gi = s_item;
    __syncthreads();
    if (gi >= plan.nItems * B) return;
    const int f = gi % B;
    const PyrItem it = items[gi / B];
    int* fl = flags + (size_t)f * totalBands;
    if (it.depFirst <= it.depLast) {
      if (threadIdx.x == 0) {
        for (int d = it.depFirst; d <= it.depLast; d++) {
          const int* p = fl + plan.bandBase[it.depLevel] + d;
          while (atomicAdd(const_cast<int*>(p), 0) == 0) __nanosleep(64);
        }
        __threadfence();
      }
      __syncthreads();
    }
    u8* frame = pyr + (size_t)f * pyrStride;
    if (it.kind == 0) {
      const LevelGeom& L = g.lv[0];
      const u8* src = img + (size_t)f * frameStride;
      u8* dstBase = frame + L.off - (long long)kEdge * L.pitch - kLeftPad;
      const int by = it.band * 8 + wid;
      if (by < L.h + 2 * kEdge) level0_interior_row(L, src, step, dstBase, by, lane);
      if (wid < 2) level0_edge_rows(L, src, step, dstBase, it.band * 8 + 4 * wid, lane);
    } else if (it.kind == 1) {
      const LevelGeom& D = g.lv[it.level];
      const int j0 = it.band * kPyrFusedBand, j1 = min(j0 + kPyrFusedBand, D.h);
      for (int c = -20 + 128 * wid; c <= D.w + kEdge - 1; c += 128 * 8) resize_strip_warp<true>(g, it.level, frame, taps, c, j0, j1, lane);
    } else {
      for (int k = it.band * 8 + wid; k < min(it.band * 8 + 8, 2 * kEdge); k += 8) border_row_copy<true>(g.lv[it.level], frame, k, lane);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && it.kind != 2) atomicExch(fl + plan.bandBase[it.level] + it.band, 1);
  }
}

// ------------------------------------------------------------------------------------------
// FAST-9/16 per 30-px cell with iniTh/minTh fallback
// ------------------------------------------------------------------------------------------
// Score S(p) = largest threshold for which p is still a FAST-9 corner
//            = max(max_arcs min d, max_arcs min(-d)) - 1,  d_k = I(p) - I(circle_k).
//
// Two pixels per thread, packed as s16x2 and processed with the DPX min/max instructions
// (VIMNMX / VIMNMX3 .S16x2). The shared-memory tile is stored "pair interleaved": word j of a
// tile row holds (T[j], T[j+S]), S = ceil(cw/2), so that the 16 circle samples of the pixel pair
// (x, x+S) are 16 aligned 32-bit loads. Differences are kept biased, d' = d + 255 in [0,510], so
// a plain 32-bit subtraction never borrows across the halves.
struct FastDiffs { unsigned d[16]; };

__device__ __forceinline__ void fast_load_diffs(const unsigned* p, int tp, FastDiffs& D) {
  const unsigned c = p[0] + 0x00FF00FFu;
  const unsigned* rm3 = p - 3 * tp; const unsigned* rm2 = p - 2 * tp; const unsigned* rm1 = p - tp;
  const unsigned* rp1 = p + tp;     const unsigned* rp2 = p + 2 * tp; const unsigned* rp3 = p + 3 * tp;
  D.d[0] = c - rp3[0];   D.d[1] = c - rp3[1];   D.d[2] = c - rp2[2];    D.d[3] = c - rp1[3];
  D.d[4] = c - p[3];     D.d[5] = c - rm1[3];   D.d[6] = c - rm2[2];    D.d[7] = c - rm3[1];
  D.d[8] = c - rm3[0];   D.d[9] = c - rm3[-1];  D.d[10] = c - rm2[-2];  D.d[11] = c - rm1[-3];
  D.d[12] = c - p[-3];   D.d[13] = c - rp1[-3]; D.d[14] = c - rp2[-2];  D.d[15] = c - rp3[-1];
}

// Prefilter tests work on the raw pixel pairs, without forming differences. With d_k = c - p_k:
//   min over pairs of max(d_k, d_k+8) = c - A,  A = max over pairs of min(p_k, p_k+8)
//   max over pairs of min(d_k, d_k+8) = c - B,  B = min over pairs of max(p_k, p_k+8)
// and a corner at threshold th needs c - A > th (all of some 9-arc darker ... every arc holds one pixel
// of each antipodal pair) or B - c > th. Both comparisons are made per 16-bit half with a 0x200 bias so
// that no borrow crosses the halves: bit 9 of (c + K - A) is set iff A < c - th, K = 0x200 - (th + 1).
// Stage 0 applies this test to the two antipodal pairs on the axes, inline in k_fast_cells (the vertical pair comes from
// a register window that slides down the column); stage 1:
// The same test on 8 of the 16 circle pixels: the 4 pairs of the even positions.
__device__ __forceinline__ unsigned fast_bound4(const unsigned* p, int tp, unsigned K) {
  const unsigned c = p[0];
  const unsigned p0 = p[3 * tp], p8 = p[-3 * tp], p4 = p[3], p12 = p[-3];
  const unsigned p2 = p[2 * tp + 2], p10 = p[-2 * tp - 2], p6 = p[-2 * tp + 2], p14 = p[2 * tp - 2];
  const unsigned A = __vmaxs2(__vimax3_s16x2(__vmins2(p0, p8), __vmins2(p4, p12), __vmins2(p2, p10)), __vmins2(p6, p14));
  const unsigned B = __vmins2(__vimin3_s16x2(__vmaxs2(p0, p8), __vmaxs2(p4, p12), __vmaxs2(p2, p10)), __vmaxs2(p6, p14));
  return ((c + K - A) | (B + K - c)) & 0x02000200u;
}

// Exact S + 256 in each half: sliding 9-window min / max over the circular sequence, two
// 3-input stages each.
__device__ __forceinline__ unsigned fast_exact(const FastDiffs& D) {
  unsigned mn3[16], mx3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    mn3[k] = __vimin3_s16x2(D.d[k], D.d[(k + 1) & 15], D.d[(k + 2) & 15]);
    mx3[k] = __vimax3_s16x2(D.d[k], D.d[(k + 1) & 15], D.d[(k + 2) & 15]);
  }
  unsigned a[16], b[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    a[k] = __vimin3_s16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
    b[k] = __vimax3_s16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
  }
  unsigned A = __vimax3_s16x2(__vimax3_s16x2(a[0], a[1], a[2]), __vimax3_s16x2(a[3], a[4], a[5]), __vimax3_s16x2(a[6], a[7], a[8]));
  A = __vimax3_s16x2(A, __vimax3_s16x2(a[9], a[10], a[11]), __vimax3_s16x2(a[12], a[13], a[14]));
  A = __vmaxs2(A, a[15]);
  unsigned B = __vimin3_s16x2(__vimin3_s16x2(b[0], b[1], b[2]), __vimin3_s16x2(b[3], b[4], b[5]), __vimin3_s16x2(b[6], b[7], b[8]));
  B = __vimin3_s16x2(B, __vimin3_s16x2(b[9], b[10], b[11]), __vimin3_s16x2(b[12], b[13], b[14]));
  B = __vmins2(B, b[15]);
  return __vmaxs2(A, 0x01FE01FEu - B);
}

// One WARP per cell, persistent warps striding over all (frame, level, cell) items of the chunk.
// The cell's detection window is x in [19+j*wCell, min(18+(j+1)*wCell, w-20)] (sub-image
// [iniX,maxX) minus FAST's own 3-px frame), so windows tile the level and the 3x3 NMS never sees
// across a cell boundary (scores outside the window count as 0).
// Per cell: raw bytes arrive by 16-byte cp.async (prefetched while the previous cell is being scored) ->
// pair tile -> 4-pixel prefilter, one lane per column, survivors kept as a bit mask per lane -> queue ->
// 8-pixel prefilter on the queue -> exact score on the queue (dense, no divergence) -> 3x3 NMS on the hit list ->
// iniTh / minTh decision -> one atomicAdd per cell reserves the output range. Only __syncwarp() is needed: warps
// run out of phase and hide each other's latencies.
//
// Pair tile: the two pixels of a word are VERTICAL neighbours at distance R = ceil(ch / 2): word (r, T) =
// (I[r][T], I[r + R][T]), one word per column, R + 6 rows. Both pixels see the 16 circle samples at the same word
// offsets, and the tile is built from the raw rows word-wise: two aligned 32-bit loads (rows r and r + R), six byte
// permutes and one 16-byte store per 4 columns (the former (x, x + S) horizontal pairing needed two byte loads per word).
constexpr int kFastWarps = 11;  // upper bound; the host picks the warps per CTA that pack an SM's shared memory best
constexpr int kFastThreads = 32 * kFastWarps;
constexpr int kFastRun = 16;   // consecutive cells a warp grabs per atomic ...
constexpr int kFastTailRun = 2; // ... except for the last ~one run per warp of the chunk, handed out in short runs so that
                                // all warps finish together (the kernel is persistent: the tail is pure imbalance)

struct FastSmemLayout {   // per-warp shared memory carve-up (in bytes), sized for the largest cell
  int rawPitchWords, rawBytes, tileBytes, hitsBytes, scBytes, total;
};

struct CellDesc {
  const u8* src;          // 16-byte aligned address at or left of tile byte (0,0) = image (x0-3, y0-3)
  int ox;                 // image column x0-3 sits `ox` (0..15) bytes into a raw row
  int pitch, cw, ch, x0, y0, l, f, rw;
};

// The levels can be dealt to two GROUPS by the shared memory their cells need, one launch per group (ORB_B200_FAST_SPLIT=1):
// a warp's carve-up is sized for the largest cell of its launch, and the few levels with tall cells (e.g. 32 x 40 at KITTI's
// level 5, where 73 detection rows make two cell rows) cost every warp 1.4 KB, i.e. two resident warps per SM. Off by
// default (one group holding all levels): the second launch costs more than the occupancy returns, see build_geom.
struct FastGroup {
  int nPos;                       // levels of this group
  int posLevel[kMaxLevels];       // their level indices, in item order
  int posBase[kMaxLevels + 1];    // first cell of every position in the group's cell numbering; [nPos] = cells per frame
};

// Walks the (frame, level, cell-row, cell-column) items of a contiguous item range without
// divisions after the first item.
struct CellCursor {
  int f, l, ci, cj, p;
};

__device__ __forceinline__ void cursor_init(const Geom& g, const FastGroup& G, int item, CellCursor& k) {
  const int cells = G.posBase[G.nPos];
  k.f = item / cells;
  const int cc = item - k.f * cells;
  int p = 0;
#pragma unroll 1
  while (p + 1 < G.nPos && cc >= G.posBase[p + 1]) p++;
  const int cell = cc - G.posBase[p];
  k.p = p;
  k.l = G.posLevel[p];
  k.ci = cell / g.lv[k.l].nCols;
  k.cj = cell - k.ci * g.lv[k.l].nCols;
}

__device__ __forceinline__ void cursor_next(const Geom& g, const FastGroup& G, CellCursor& k) {
  if (++k.cj == g.lv[k.l].nCols) {
    k.cj = 0;
    if (++k.ci == g.lv[k.l].nRows) {
      k.ci = 0;
      if (++k.p == G.nPos) { k.p = 0; k.f++; }
      k.l = G.posLevel[k.p];
    }
  }
}

__device__ __forceinline__ bool fast_cell_desc(const Geom& g, const u8* pyr, size_t pyrStride, const CellCursor& k, CellDesc& c) {
  const LevelGeom& L = g.lv[k.l];
  c.x0 = kEdge + k.cj * L.wCell;
  c.y0 = kEdge + k.ci * L.hCell;
  c.cw = min(c.x0 + L.wCell - 1, L.w - kEdge - 1) - c.x0 + 1;
  c.ch = min(c.y0 + L.hCell - 1, L.h - kEdge - 1) - c.y0 + 1;
  c.l = k.l;
  c.f = k.f;
  c.pitch = L.pitch;
  // interior column 0 sits at byte kLeftPad (a multiple of 16) of a 32-byte aligned row: 16-byte chunks
  const int xs = (c.x0 - 3) & ~15;
  c.ox = (c.x0 - 3) - xs;
  c.rw = (c.ox + c.cw + 6 + 15) >> 4;             // chunks per row (<= 6)
  c.src = pyr + (size_t)k.f * pyrStride + L.off + (long long)(c.y0 - 3) * L.pitch + xs;
  return c.cw > 0 && c.ch > 0;
}

__device__ __forceinline__ void fast_prefetch(const CellDesc& c, unsigned* raw, int rawPitchWords, int lane) {
  // (ch+6) rows x rw (<= 6) 16-byte chunks: 8 rows x 4 chunks per step (4 rows x 8 chunks for wide cells)
  const int rows = c.ch + 6;
  const int sh = c.rw <= 4 ? 2 : 3;               // warp-uniform
  const int cx = lane & ((1 << sh) - 1), r0 = lane >> sh, dr = 32 >> sh;
  if (cx < c.rw) {
    const u8* s = c.src + (r0 * c.pitch + 16 * cx);
    unsigned* d = raw + r0 * rawPitchWords + 4 * cx;
    const int sstep = dr * c.pitch, dstep = dr * rawPitchWords;
    for (int r = r0; r < rows; r += dr) {
      __pipeline_memcpy_async(d, s, 16);
      s += sstep;
      d += dstep;
    }
  }
  __pipeline_commit();
}

// flags: bit 0 = skip the 8-pixel prefilter (stage 1)
__global__ void __launch_bounds__(kFastThreads) k_fast_cells(const Geom g, const u8* __restrict__ pyr, size_t pyrStride,
                                                             uint2* __restrict__ cand, int* __restrict__ candCount,
                                                             int candTotal, int nItems, const FastSmemLayout lay,
                                                             const FastGroup G, int* __restrict__ workCounter, int run, int tailRun,
                                                             int tailMul, int flags) {
  extern __shared__ __align__(16) u8 smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u8* base = smem + wid * lay.total;
  unsigned* raw = reinterpret_cast<unsigned*>(base);
  unsigned* tile = reinterpret_cast<unsigned*>(base + lay.rawBytes);
  // hits (produced, from the front) and the queue (consumed, stored at the back) share one buffer
  // of 2*R*cw entries: a batch of 32 queue items yields at most 64 hits, so writes never reach
  // the unread part of the queue (2*(e0+32) <= R*cw + e0 + 32 whenever e0 + 32 <= R*cw)
  unsigned short* hits = reinterpret_cast<unsigned short*>(base + lay.rawBytes + lay.tileBytes);
  u8* sc = base + lay.rawBytes + lay.tileBytes + lay.hitsBytes;
  const unsigned ltmask = (1u << lane) - 1u;
  // the score plane is all zero between cells: cleared once here, and after every cell at the positions it wrote
  for (int i = lane; i < lay.scBytes / 4; i += 32) reinterpret_cast<unsigned*>(sc)[i] = 0u;

  // work distribution: a warp grabs runs of consecutive cells from a global counter of work units
  const int tailItems = min(nItems, (int)(gridDim.x * (blockDim.x >> 5)) * run * tailMul);
  const int longUnits = (nItems - tailItems) / run;
  CellDesc c;
  CellCursor cur_k;
  int left = 0;          // cells left in the current run
  bool have = false;
  auto advance = [&]() {
    // moves to the next cell with a non-empty detection window; false when the work is exhausted
    for (;;) {
      if (left == 0) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(workCounter, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        const int start = unit < longUnits ? unit * run : longUnits * run + (unit - longUnits) * tailRun;
        if (start >= nItems) return false;
        left = min(unit < longUnits ? run : tailRun, nItems - start);
        cursor_init(g, G, start, cur_k);
      } else {
        cursor_next(g, G, cur_k);
      }
      left--;
      if (fast_cell_desc(g, pyr, pyrStride, cur_k, c)) return true;
    }
  };
  have = advance();
  if (have) fast_prefetch(c, raw, lay.rawPitchWords, lane);

  while (have) {
    const int cw = c.cw, ch = c.ch;
    const int R = (ch + 1) >> 1;            // pair distance: word (r, T) = rows (r, r + R) of the tile
    const int xoff = (c.ox & 3) + 3;        // tile column of cell column 0
    const int tp = (xoff + cw + 3 + 3) & ~3; // tile pitch in words (whole raw words: multiple of 4)
    const int sp = cw + 2;                  // score pitch (1-px zero apron)
    __pipeline_wait_prior(0);
    __syncwarp();
    // raw rows -> pair tile, word-wise: lane = (row group q, raw word k); 4 tile words per lane and step
    {
      const int wpr = tp >> 2;                       // raw words per tile row (<= 18)
      const int rps = wpr <= 8 ? 4 : (wpr <= 10 ? 3 : (wpr <= 16 ? 2 : 1));   // rows per step (32 / wpr)
      const int q = wpr <= 8 ? lane >> 3 : (wpr <= 10 ? lane / 10 : (wpr <= 16 ? lane >> 4 : 0));
      const int k = wpr <= 8 ? lane & 7 : (wpr <= 10 ? lane - 10 * q : (wpr <= 16 ? lane & 15 : lane));
      if (q < rps && k < wpr) {
        const int rpw = lay.rawPitchWords;
        const unsigned* ra = raw + (c.ox >> 2) + k;
        uint4* t = reinterpret_cast<uint4*>(tile + q * tp + 4 * k);
        // rows r and r + R of the raw buffer; for odd ch the partner of tile row R+5 would be raw row ch+6, which is not
        // fetched (one more raw row would cost a resident warp per SM): the row above stands in, its half of the word
        // only feeds a pixel that is masked later
        const int rows = R + 6, last = ch + 5;
        const unsigned* pa = ra + q * rpw;
        const int hiOff = R * rpw, stepA = rps * rpw;
        for (int r = q; r < rows; r += rps, pa += stepA) {
          const unsigned a = pa[0], b = pa[r + R > last ? hiOff - rpw : hiOff];
          const unsigned x01 = __byte_perm(a, b, 0x5410), x23 = __byte_perm(a, b, 0x7632);
          uint4 o;
          o.x = __byte_perm(x01, 0u, 0x4240);
          o.y = __byte_perm(x01, 0u, 0x4341);
          o.z = __byte_perm(x23, 0u, 0x4240);
          o.w = __byte_perm(x23, 0u, 0x4341);
          *t = o;
          t += (rps * tp) >> 2;
        }
      }
    }
    __syncwarp();
    // the raw buffer is free again: prefetch the next cell of this warp
    const CellDesc cur = c;
    have = advance();
    if (have) fast_prefetch(c, raw, lay.rawPitchWords, lane);

    // Two threshold passes like the reference (:1111-1124): cv::FAST(iniTh) first; only if its
    // result (after NMS) is empty, cv::FAST(minTh). Scoring pixels at iniTh first keeps the exact
    // scoring away from the many weak corners of the ~97 % of cells that have a strong one.
    unsigned short* queue = hits + R * cw;   // the queue fills [R*cw, R*cw + nq), hits grow from 0
    int th = g.iniTh, nh = 0, total = 0;
#pragma unroll 1
    for (;;) {
      // bit 9 of a half of the prefilter word is set iff a corner at th is possible there (see fast_bound4)
      const unsigned K = 0x02000200u - (unsigned)(th + 1) * 0x00010001u;
      // ---- prefilter, stage 0 (the two antipodal pairs on the axes, every pixel pair). A lane walks DOWN its column with
      // the 7 rows of the vertical arm in registers (sliding window): per step one new row word and the two horizontal
      // neighbours are loaded. The outcome of step i is shifted into a per-lane bit mask; the queue is written after the
      // walk from the masks (one warp scan instead of a ballot, two population counts and a store per step).
      int nq = 0;
      for (int xb = 0; xb < cw; xb += 32) {        // one sweep (cw <= 32) or two
        const int x = xb + lane;
        const unsigned* col = tile + (xoff + min(x, cw - 1));
        unsigned w0 = col[0], w1 = col[tp], w2 = col[2 * tp], w3 = col[3 * tp], w4 = col[4 * tp], w5 = col[5 * tp];
        const unsigned* pc = col + 3 * tp;           // centre of step 0
        const int tp3 = 3 * tp;
        unsigned mask = 0;
#pragma unroll 8
        for (int i = 0; i < R; i++) {
          const unsigned w6 = pc[tp3];
          const unsigned pl = pc[-3], pr = pc[3];
          const unsigned A = __vmaxs2(__vmins2(w6, w0), __vmins2(pr, pl));
          const unsigned B = __vmins2(__vmaxs2(w6, w0), __vmaxs2(pr, pl));
          const unsigned b = ((w3 + K - A) | (B + K - w3)) & 0x02000200u;
          mask = (mask >> 1) | (b ? 0x80000000u : 0u);
          w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
          pc += tp;
        }
        mask = x < cw ? mask >> (32 - R) : 0u;       // bit r = the pair (r, r + R) of this column passed (1 <= R <= 30)
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        unsigned short* qp = queue + nq + (incl - cnt);
        while (mask) {
          const int r = __ffs(mask) - 1;
          mask &= mask - 1;
          *qp++ = (unsigned short)((r << 6) | x);
        }
        nq += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();
      if (!(flags & 1)) {
        int nq1 = 0;
        for (int e0 = 0; e0 < nq; e0 += 32) {
          const int e = e0 + lane;
          bool pass = false;
          int i = 0;
          if (e < nq) {
            i = queue[e];
            pass = fast_bound4(tile + ((i >> 6) + 3) * tp + ((i & 63) + xoff), tp, K) != 0u;
          }
          const unsigned m = __ballot_sync(0xffffffffu, pass);   // every lane has read its entry by now
          __syncwarp();                                          // (memory ordering of the in-place compaction, for racecheck)
          if (pass) queue[nq1 + __popc(m & ltmask)] = (unsigned short)i;
          nq1 += __popc(m);
        }
        nq = nq1;
        __syncwarp();
      }
      // ---- exact score of the queued pairs
      nh = 0;
      for (int e0 = 0; e0 < nq; e0 += 32) {
        const int e = e0 + lane;
        int sLo = 0, sHi = 0, r = 0, x = 0;
        if (e < nq) {
          const int i = queue[e];
          r = i >> 6; x = i & 63;
          FastDiffs D;
          fast_load_diffs(tile + (r + 3) * tp + (x + xoff), tp, D);
          const unsigned s2 = fast_exact(D);
          sLo = (int)(s2 & 0xffffu) - 256;
          sHi = r + R < ch ? (int)(s2 >> 16) - 256 : 0;
        }
        const bool hLo = sLo >= th, hHi = sHi >= th;
        const unsigned mLo = __ballot_sync(0xffffffffu, hLo), mHi = __ballot_sync(0xffffffffu, hHi);
        if (hLo) {
          sc[(r + 1) * sp + x + 1] = (u8)sLo;
          hits[nh + __popc(mLo & ltmask)] = (unsigned short)((r << 6) | x);
        }
        nh += __popc(mLo);
        if (hHi) {
          sc[(r + R + 1) * sp + x + 1] = (u8)sHi;
          hits[nh + __popc(mHi & ltmask)] = (unsigned short)(((r + R) << 6) | x);
        }
        nh += __popc(mHi);
      }
      __syncwarp();
      // ---- strict 3x3 maximum inside the cell (scores below th count as 0, as in cv::FAST); every hit scores >= th
      total = 0;
      for (int e0 = 0; e0 < nh; e0 += 32) {
        const int e = e0 + lane;
        bool keep = false;
        if (e < nh) {
          const int p = hits[e];
          const int r = p >> 6, cx = p & 63;
          const u8* q = sc + (r + 1) * sp + cx + 1;
          const int s = q[0];
          const int m = max(max(max(q[-1], q[1]), max(q[-sp - 1], q[-sp])), max(max(q[-sp + 1], q[sp - 1]), max(q[sp], q[sp + 1])));
          keep = s > m;
          if (keep) hits[e] = (unsigned short)(p | 0x8000);
        }
        total += __popc(__ballot_sync(0xffffffffu, keep));
      }
      if (total > 0 || th == g.minTh) break;
      th = g.minTh;   // nothing at iniTh: redo the cell at minTh (the scores already written stay valid)
      __syncwarp();
    }
    // ---- emit: reserve with one atomic, write; give the score plane back all zero
    const LevelGeom& L = g.lv[cur.l];
    int basei = 0;
    if (total > 0) {
      if (lane == 0) basei = atomicAdd(candCount + cur.f * g.nlevels + cur.l, total);
      basei = __shfl_sync(0xffffffffu, basei, 0);
    }
    uint2* out = cand + (size_t)cur.f * candTotal + L.candOff;
    for (int e0 = 0; e0 < nh; e0 += 32) {
      const int e = e0 + lane;
      bool keep = false;
      int r = 0, cx = 0, s = 0;
      if (e < nh) {
        const int p = hits[e];
        r = (p >> 6) & 0x1ff; cx = p & 63;
        u8* q = sc + (r + 1) * sp + cx + 1;
        s = *q;
        keep = (p & 0x8000) != 0;
      }
      __syncwarp();   // all scores of the batch are read before any is cleared (a position appears once)
      if (e < nh) sc[(r + 1) * sp + cx + 1] = 0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int idx = basei + __popc(m & ltmask);
        if (idx < L.candCap) out[idx] = make_uint2((unsigned)(cur.x0 + cx) | ((unsigned)(cur.y0 + r) << 16), (unsigned)s);
      }
      basei += __popc(m);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Quadtree distribution (DistributeOctTree), one CTA per (level, frame)
// ------------------------------------------------------------------------------------------
struct QtNode {
  short x0, x1, y0, y1;  // [x0,x1) x [y0,y1) relative to the 16-px detection border
  int count;             // keys inside
  int seq;               // creation order; list order of the reference == descending seq
};

__device__ __forceinline__ int qt_quadrant(int xr, int yr, const QtNode& nd) {
  const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2)  (:608)
  const int my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
  return (xr < mx ? 0 : 1) + (yr < my ? 0 : 2);       // n1, n2, n3, n4 (:644-662)
}

// In-place exclusive scan of a[0..n) in shared memory by the whole CTA; returns the total.
__device__ int block_scan_excl(int* a, int n, int* tmp /* >= 34 ints */) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + T - 1) / T;
  const int b = min(tid * per, n), e = min(b + per, n);
  int s = 0;
  for (int i = b; i < e; i++) s += a[i];
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) tmp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < (T >> 5) ? tmp[lane] : 0;
    int inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += t;
    }
    tmp[lane] = inc2 - v;
    if (lane == 31) tmp[33] = inc2;
  }
  __syncthreads();
  int base = tmp[wid] + incl - s;
  for (int i = b; i < e; i++) {
    const int t = a[i];
    a[i] = base;
    base += t;
  }
  const int total = tmp[33];
  __syncthreads();
  return total;
}

// The reference algorithm, restated level-synchronously:
//  * a "full pass" splits every multi-key node; its result set is order independent;
//  * the "sorted partial pass" walks multi-key nodes by (count, creation order) descending and
//    stops the moment the list holds >= N nodes: each split's gain (non-empty children - 1) is
//    independent of the others, so the stop position is a prefix-sum search;
//  * every new node is pushed to the list front, so final list order == descending creation seq.
// Canonical tie rule (the reference sorts by heap address): later-created node first.
__global__ void __launch_bounds__(kQtMaxThreads) k_quadtree(const Geom g, const uint2* __restrict__ cand,
                                                         const int* __restrict__ candCount,
                                                         unsigned short* __restrict__ keyNode,
                                                         uint2* __restrict__ kept, int* __restrict__ keptCount,
                                                         int candTotal, int keptTotal, int nodeCap, int seqWords,
                                                         int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char qsm[];
  __shared__ int s_tmp[34];
  __shared__ int s_nc, s_nexp, s_cut, s_roots;
  // grid = (frames, levels): CTAs are dispatched x-fastest, so all frames of level 0 (the most candidates, the longest
  // CTAs) start first and the short top levels fill the tail
  const int l = blockIdx.y, f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const LevelGeom& L = g.lv[l];
  int n = candCount[f * g.nlevels + l];
  if (n > L.candCap) {
    n = L.candCap;
    if (tid == 0) atomicOr(overflow, 1);
  }
  if (n == 0) {
    if (tid == 0) keptCount[f * g.nlevels + l] = 0;
    return;
  }
  const uint2* C = cand + (size_t)f * candTotal + L.candOff;
  unsigned short* KN = keyNode + (size_t)f * candTotal + L.candOff;
  uint2* K = kept + (size_t)f * keptTotal + L.keptOff;

  QtNode* nodes = (QtNode*)qsm;
  QtNode* nodes2 = nodes + nodeCap;
  int* childCnt = (int*)(nodes2 + nodeCap);
  int* childIdx = childCnt + 4 * nodeCap;
  int* slot = childIdx + 4 * nodeCap;
  int* nodeRank = slot + nodeCap;
  int* candList = nodeRank + nodeCap;
  int* byRank = candList + nodeCap;
  unsigned long long* key64 = (unsigned long long*)(byRank + nodeCap);
  // positions and node ids of up to kQtSmemKeys candidates live in shared memory (they are re-read
  // in every refinement round); larger levels fall back to the global scratch
  unsigned* s_xy = (unsigned*)(key64 + nodeCap);
  unsigned short* s_kn = (unsigned short*)(s_xy + kQtSmemKeys);
  // ranking by creation sequence without comparing all pairs: a bitmap over the sequence numbers (one bit per number
  // handed out so far) and the running population count of its words; rank = set bits above
  unsigned* bm = (unsigned*)(s_kn + kQtSmemKeys);
  int* bmPre = (int*)(bm + seqWords);
  const bool inSmem = n <= kQtSmemKeys;
  const unsigned* XY;          // generic pointers: shared or global
  unsigned short* KNp;
  if (inSmem) {
    for (int k = tid; k < n; k += T) s_xy[k] = C[k].x;
    XY = s_xy;
    KNp = s_kn;
  } else {
    XY = nullptr;
    KNp = KN;
  }
  __syncthreads();

  const int N = L.nfeat;
  const int nIni = L.nIni;
  const float hX = L.hX;
  const int Hp = L.h - 32;

  // roots (:700-741); empty roots are erased (:745-763)
  if (tid < nIni) {
    QtNode r;
    r.x0 = (short)(int)__fmul_rn(hX, (float)tid);
    r.x1 = (short)(int)__fmul_rn(hX, (float)(tid + 1));
    r.y0 = 0;
    r.y1 = (short)Hp;
    r.count = 0;
    r.seq = nIni - 1 - tid;
    nodes[tid] = r;
  }
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    const int xr = (int)((inSmem ? XY[k] : C[k].x) & 0xffffu) - 16;
    int r = (int)__fdiv_rn((float)xr, hX);
    r = min(max(r, 0), nIni - 1);
    KNp[k] = (unsigned short)r;
    atomicAdd(&nodes[r].count, 1);
  }
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int i = 0; i < nIni; i++) {
      if (nodes[i].count > 0) { slot[i] = m; nodes2[m++] = nodes[i]; }
      else slot[i] = -1;
    }
    s_roots = m;
  }
  __syncthreads();
  for (int k = tid; k < n; k += T) KNp[k] = (unsigned short)slot[KNp[k]];
  int numNodes = s_roots;
  { QtNode* t = nodes; nodes = nodes2; nodes2 = t; }
  int seqCounter = nIni;
  bool partial = false;
  __syncthreads();

  for (;;) {
    const int prev = numNodes;
    for (int i = tid; i < 4 * numNodes; i += T) childCnt[i] = 0;
    if (tid == 0) { s_nc = 0; s_nexp = 0; s_cut = 0x7fffffff; }
    __syncthreads();
    for (int i = tid; i < numNodes; i += T) {
      slot[i] = 1;
      nodeRank[i] = -1;
      if (nodes[i].count > 1) candList[atomicAdd(&s_nc, 1)] = i;
    }
    for (int k = tid; k < n; k += T) {
      const int nd = KNp[k];
      const QtNode Nd = nodes[nd];
      if (Nd.count > 1) {
        const unsigned xy = inSmem ? XY[k] : C[k].x;
        atomicAdd(&childCnt[4 * nd + qt_quadrant((int)(xy & 0xffffu) - 16, (int)(xy >> 16) - 16, Nd)], 1);
      }
    }
    __syncthreads();
    const int nc = s_nc;
    if (nc == 0) break;  // nothing left to split: list size unchanged (:890)
    for (int i = tid; i < nc; i += T) {
      const QtNode& Nd = nodes[candList[i]];
      key64[i] = (partial ? ((unsigned long long)(unsigned)Nd.count << 32) : 0ull) | (unsigned)Nd.seq;
    }
    __syncthreads();
    // rank of every node to split: by (key count, creation sequence) descending in the sorted partial pass, by
    // creation sequence descending in a full pass (there the order only fixes the children's sequence numbers)
    const int words = (seqCounter + 31) >> 5;
    const bool byBitmap = !partial && words <= seqWords;
    if (byBitmap) {
      for (int w = tid; w < words; w += T) bm[w] = 0u;
      __syncthreads();
      for (int i = tid; i < nc; i += T) {
        const unsigned sq = (unsigned)key64[i];
        atomicOr(&bm[sq >> 5], 1u << (sq & 31));
      }
      __syncthreads();
      for (int w = tid; w < words; w += T) bmPre[w] = __popc(bm[w]);
      __syncthreads();
      block_scan_excl(bmPre, words, s_tmp);
    }
    for (int i = tid; i < nc; i += T) {
      const unsigned long long ki = key64[i];
      int r = 0;
      if (byBitmap) {
        const unsigned sq = (unsigned)ki;
        r = nc - 1 - (bmPre[sq >> 5] + __popc(bm[sq >> 5] & ((1u << (sq & 31)) - 1u)));
      } else {
        for (int j = 0; j < nc; j++) r += key64[j] > ki;
      }
      const int nd = candList[i];
      const int ne = (childCnt[4 * nd] > 0) + (childCnt[4 * nd + 1] > 0) + (childCnt[4 * nd + 2] > 0) + (childCnt[4 * nd + 3] > 0);
      nodeRank[nd] = r;
      byRank[r] = ne - 1;
    }
    __syncthreads();
    int cut = nc - 1;
    if (partial) {
      // first rank at which the running list size reaches N (:982)
      if (tid < 32) {
        int running = numNodes;
        for (int base = 0; base < nc; base += 32) {
          const int v = base + tid < nc ? byRank[base + tid] : 0;
          int incl = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += t;
          }
          const unsigned hit = __ballot_sync(0xffffffffu, base + tid < nc && running + incl >= N);
          if (hit) {
            if (tid == 0) s_cut = base + __ffs(hit) - 1;
            break;
          }
          running += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      __syncthreads();
      cut = min(s_cut, nc - 1);
    }
    for (int i = tid; i < nc; i += T) {
      const int nd = candList[i];
      if (nodeRank[nd] <= cut) slot[nd] = byRank[nodeRank[nd]] + 1;
      else nodeRank[nd] = -1;
    }
    __syncthreads();
    const int total = block_scan_excl(slot, numNodes, s_tmp);
    for (int i = tid; i < numNodes; i += T) {
      const QtNode Nd = nodes[i];
      const int base = slot[i], r = nodeRank[i];
      if (r < 0) {
        nodes2[base] = Nd;
      } else {
        const int mx = Nd.x0 + ((Nd.x1 - Nd.x0 + 1) >> 1), my = Nd.y0 + ((Nd.y1 - Nd.y0 + 1) >> 1);
        int j = 0, nexp = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int c = childCnt[4 * i + q];
          if (c > 0) {
            QtNode ch;
            ch.x0 = (q & 1) ? (short)mx : Nd.x0;
            ch.x1 = (q & 1) ? Nd.x1 : (short)mx;
            ch.y0 = (q & 2) ? (short)my : Nd.y0;
            ch.y1 = (q & 2) ? Nd.y1 : (short)my;
            ch.count = c;
            ch.seq = seqCounter + 4 * r + q;
            nodes2[base + j] = ch;
            childIdx[4 * i + q] = base + j;
            j++;
            nexp += c > 1;
          }
        }
        if (nexp) atomicAdd(&s_nexp, nexp);
      }
    }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
      const int nd = KNp[k];
      if (nodeRank[nd] < 0) {
        KNp[k] = (unsigned short)slot[nd];
      } else {
        const unsigned xy = inSmem ? XY[k] : C[k].x;
        KNp[k] = (unsigned short)childIdx[4 * nd + qt_quadrant((int)(xy & 0xffffu) - 16, (int)(xy >> 16) - 16, nodes[nd])];
      }
    }
    __syncthreads();
    seqCounter += 4 * (cut + 1);
    numNodes = total;
    { QtNode* t = nodes; nodes = nodes2; nodes2 = t; }
    const int nToExpand = s_nexp;
    __syncthreads();
    if (numNodes >= N || numNodes == prev) break;                 // :890, :987
    if (!partial && numNodes + 3 * nToExpand > N) partial = true;  // :908
  }

  // best response per node, earliest candidate wins ties (:1004-1029); the reference's
  // candidate order (cell-row-major, row-major inside a cell) is a function of (x,y).
  for (int i = tid; i < numNodes; i += T) key64[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    const uint2 c = C[k];
    const int x = (int)(c.x & 0xffffu) - kEdge, y = (int)(c.x >> 16) - kEdge;
    const int ci = y / L.hCell, cj = x / L.wCell;
    const unsigned ord = (unsigned)(((ci * L.nColsAll + cj) * L.hCell + (y - ci * L.hCell)) * L.wCell + (x - cj * L.wCell));
    const unsigned long long pk = ((unsigned long long)c.y << 56) | ((unsigned long long)(0x7fffffffu - ord) << 24) | (unsigned)k;
    atomicMax(&key64[KNp[k]], pk);
  }
  __syncthreads();
  // final list order == descending creation sequence
  const int wordsF = (seqCounter + 31) >> 5;
  const bool bitmapF = wordsF <= seqWords;
  if (bitmapF) {
    for (int w = tid; w < wordsF; w += T) bm[w] = 0u;
    __syncthreads();
    for (int i = tid; i < numNodes; i += T) {
      const unsigned sq = (unsigned)nodes[i].seq;
      atomicOr(&bm[sq >> 5], 1u << (sq & 31));
    }
    __syncthreads();
    for (int w = tid; w < wordsF; w += T) bmPre[w] = __popc(bm[w]);
    __syncthreads();
    block_scan_excl(bmPre, wordsF, s_tmp);
  }
  for (int i = tid; i < numNodes; i += T) {
    const int s = nodes[i].seq;
    int r = 0;
    if (bitmapF) {
      r = numNodes - 1 - (bmPre[s >> 5] + __popc(bm[s >> 5] & ((1u << (s & 31)) - 1u)));
    } else {
      for (int j = 0; j < numNodes; j++) r += nodes[j].seq > s;
    }
    if (r < L.keptCap) K[r] = C[(int)(key64[i] & 0xffffffu)];
  }
  if (tid == 0) {
    keptCount[f * g.nlevels + l] = min(numNodes, L.keptCap);
    if (numNodes > L.keptCap) atomicOr(overflow, 2);
  }
}

// ------------------------------------------------------------------------------------------
// 7x7 sigma=2 Gaussian blur, OpenCV 4.x 8-bit fixed-point path: k = [18,34,48,56,48,34,18]/256
// horizontally into u16, vertically (+32768)>>16. Reads the bordered plane, whose border IS the
// REFLECT_101 content the reference's blur of the clone() re-synthesises.
// ------------------------------------------------------------------------------------------
// Integer dot-product instructions do the taps: the horizontal pass is two IDP.4A per pixel on
// byte windows built with funnel shifts, the vertical pass four IDP.2A per pixel on u16 pairs
// (two tile rows interleaved per 32-bit word). One CTA blurs a 128 x 32 tile.
// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers, used by k_blur7 and k_describe_tma.
// Measured on B200: the innermost start coordinate of a byte tensor must sit on a 16-byte boundary (an unaligned one
// raises "illegal instruction"), so boxes are fetched from the aligned column at or left of the wanted one.
struct PyrMaps {
  CUtensorMap m[kMaxLevels];
};
constexpr int kBlurTW = 128, kBlurTH = 64;
constexpr int kBlurChunks = (kBlurTW + 32) / 16;  // 16-byte chunks per input tile row: image x0-16 .. x0+143
constexpr int kBlurInP = kBlurChunks * 4 + 4;     // tile pitch in words (16-B aligned rows, 4 words pad)
constexpr int kBlurRows = kBlurTH + 6;        // input rows y0-3 .. y0+34
constexpr int kBlurPairs = kBlurRows / 2;

template <bool TMA>
__global__ void __launch_bounds__(256) k_blur7(const Geom g, const u8* __restrict__ pyr, size_t pyrStride,
                                               u8* __restrict__ blur, size_t blurStride,
                                               const __grid_constant__ PyrMaps bmaps) {
  __shared__ __align__(128) unsigned in[kBlurRows * kBlurInP];
  __shared__ __align__(8) unsigned long long s_bbar;
  __shared__ int s_level;
  constexpr int inP = TMA ? kBlurChunks * 4 : kBlurInP;   // the TMA box is dense: 160-byte rows
  __shared__ __align__(16) unsigned hp[kBlurPairs * kBlurTW];  // (H[2p][x], H[2p+1][x]) as u16 pairs
  const int f = blockIdx.y, tid = threadIdx.x;
  // one thread finds the tile's level (a dependent chain of parameter loads) and starts the fetch; the rest read it
  // after the barrier they wait at anyway
  if (tid == 0) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blurTileBase) l++;
    s_level = l;
    if (TMA) {
      // one bulk tensor copy per CTA: rows y0-3 .. y0+66, bytes x0-16 .. x0+143 of the bordered plane (rows / bytes
      // outside the plane arrive as zeros; they only feed outputs that are not stored)
      const LevelGeom& L0 = g.lv[l];
      const int t0 = blockIdx.x - L0.blurTileBase;
      const int ty0 = L0.blurTilesX == 1 ? t0 : (int)__umulhi((unsigned)t0, L0.blurTilesXMagic), tx0 = t0 - ty0 * L0.blurTilesX;
      mbar_init(&s_bbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&s_bbar, kBlurRows * kBlurChunks * 16);
      tma_load_3d(in, &bmaps.m[l], kLeftPad + tx0 * kBlurTW - 16, kEdge + ty0 * kBlurTH - 3, f, &s_bbar);
    }
  }
  __syncthreads();
  const int l = s_level;
  const LevelGeom& L = g.lv[l];
  const int t = blockIdx.x - L.blurTileBase;
  const int ty = L.blurTilesX == 1 ? t : (int)__umulhi((unsigned)t, L.blurTilesXMagic), tx = t - ty * L.blurTilesX;   // t < 2^16
  const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
  if (TMA) {
    mbar_wait(&s_bbar, 0);
  } else {
    const u8* src = pyr + (size_t)f * pyrStride + L.off;
    const int xlast = L.pitch - kLeftPad - 16;  // last 16-byte chunk of a bordered row (covers col w+18)
    for (int i = tid; i < kBlurRows * kBlurChunks; i += 256) {
      const int r = i / kBlurChunks, c = i - r * kBlurChunks;
      const int y = min(y0 + r - 3, L.h + kEdge - 1), x = min(x0 - 16 + 16 * c, xlast);
      __pipeline_memcpy_async(in + r * kBlurInP + 4 * c, src + (long long)y * L.pitch + x, 16);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
  }
  // thread = a fixed group of 4 columns (xg) and every 8th row pair: all addresses advance by compile-time strides
  const int xg = tid & 31, r8 = tid >> 5;
  // horizontal: k = [18,34,48,56,48,34,18]; output x needs tile bytes (x+1 .. x+7)
  {
    const unsigned* p = in + (2 * r8) * inP + xg + 3;   // image x0+4xg-4 = tile byte 4xg+12
    unsigned* ho = hp + r8 * kBlurTW + 4 * xg;
#pragma unroll
    for (int it = 0; it < (kBlurPairs + 7) / 8; it++) {
      if (8 * it + r8 < kBlurPairs) {
        unsigned hrow[2][4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          // output x+j needs tile bytes (1+j .. 7+j) of (w0,w1,w2): the taps are applied with
          // pre-shifted coefficient words instead of shifting the data
          const unsigned w0 = p[(16 * it + q) * inP], w1 = p[(16 * it + q) * inP + 1], w2 = p[(16 * it + q) * inP + 2];
          hrow[q][0] = __dp4a(w0, 0x30221200u, __dp4a(w1, 0x12223038u, 0u));
          hrow[q][1] = __dp4a(w0, 0x22120000u, __dp4a(w1, 0x22303830u, __dp4a(w2, 0x00000012u, 0u)));
          hrow[q][2] = __dp4a(w0, 0x12000000u, __dp4a(w1, 0x30383022u, __dp4a(w2, 0x00001222u, 0u)));
          hrow[q][3] = __dp4a(w1, 0x38302212u, __dp4a(w2, 0x00122230u, 0u));
        }
        uint4 o;
        o.x = hrow[0][0] | (hrow[1][0] << 16);
        o.y = hrow[0][1] | (hrow[1][1] << 16);
        o.z = hrow[0][2] | (hrow[1][2] << 16);
        o.w = hrow[0][3] | (hrow[1][3] << 16);
        *reinterpret_cast<uint4*>(ho + 8 * it * kBlurTW) = o;
      }
    }
  }
  __syncthreads();
  // vertical: out rows 2yp, 2yp+1 need tile rows 2yp .. 2yp+7 = row pairs yp .. yp+3
  const int x = x0 + 4 * xg;
  if (x >= L.w) return;
  u8* d0 = blur + (size_t)f * blurStride + L.boff + (long long)(y0 + 2 * r8) * L.bpitch + x;
  const long long dstep = 16LL * L.bpitch;
  const unsigned* hq = hp + r8 * kBlurTW + 4 * xg;
#pragma unroll
  for (int it = 0; it < kBlurTH / 16; it++) {
    const int y = y0 + 2 * (8 * it + r8);
    if (y >= L.h) break;
    const uint4 q0 = *reinterpret_cast<const uint4*>(hq + (8 * it + 0) * kBlurTW);
    const uint4 q1 = *reinterpret_cast<const uint4*>(hq + (8 * it + 1) * kBlurTW);
    const uint4 q2 = *reinterpret_cast<const uint4*>(hq + (8 * it + 2) * kBlurTW);
    const uint4 q3 = *reinterpret_cast<const uint4*>(hq + (8 * it + 3) * kBlurTW);
    const unsigned c0[4] = {q0.x, q0.y, q0.z, q0.w}, c1[4] = {q1.x, q1.y, q1.z, q1.w};
    const unsigned c2[4] = {q2.x, q2.y, q2.z, q2.w}, c3[4] = {q3.x, q3.y, q3.z, q3.w};
    unsigned ev[4], od[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      unsigned e = __dp2a_lo(c0[j], 0x2212u, 32768u);   // rows 0,1 x (k0,k1)
      e = __dp2a_lo(c1[j], 0x3830u, e);                 // rows 2,3 x (k2,k3)
      e = __dp2a_lo(c2[j], 0x2230u, e);                 // rows 4,5 x (k4,k5)
      e = __dp2a_lo(c3[j], 0x0012u, e);                 // row 6 x k6
      unsigned o = __dp2a_lo(c0[j], 0x1200u, 32768u);   // row 1 x k0
      o = __dp2a_lo(c1[j], 0x3022u, o);                 // rows 2,3 x (k1,k2)
      o = __dp2a_lo(c2[j], 0x3038u, o);                 // rows 4,5 x (k3,k4)
      o = __dp2a_lo(c3[j], 0x1222u, o);                 // rows 6,7 x (k5,k6)
      ev[j] = e;
      od[j] = o;
    }
    // the result of pixel j is byte 2 of its accumulator (sum / 2^16, < 256): three PRMT pack four of them
    const unsigned o0 = __byte_perm(__byte_perm(ev[0], ev[1], 0x0062), __byte_perm(ev[2], ev[3], 0x0062), 0x5410);
    const unsigned o1 = __byte_perm(__byte_perm(od[0], od[1], 0x0062), __byte_perm(od[2], od[3], 0x0062), 0x5410);
    *reinterpret_cast<unsigned*>(d0) = o0;
    if (y + 1 < L.h) *reinterpret_cast<unsigned*>(d0 + L.bpitch) = o1;
    d0 += dstep;
  }
}

// ------------------------------------------------------------------------------------------
// Orientation + descriptor + output record, one warp per kept keypoint
// ------------------------------------------------------------------------------------------
// cv::fastAtan2 (float polynomial, no FMA contraction).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
  const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

constexpr int kPatchWords = 12;   // blurred patch: 37 rows x 48 bytes (37 + 8-byte alignment slack), 6 chunks of 8 B
constexpr int kMomWords = 10;     // unblurred patch: 31 rows x 40 bytes (31 + 8-byte alignment slack), 5 chunks of 8 B
constexpr int kDescSlots = 16;    // keypoint slots per CTA (8 warps x 2: 28 KB of patches per CTA keeps 8 CTAs per SM)
constexpr int kDescPerWarp = kDescSlots / 8;
constexpr int kDescRounds = 1;    // rounds of kDescSlots keypoints per CTA of k_describe_tma (2 amortises the per-lane set-up but needs 64 registers: 4 instead of 5 CTAs per SM, measured 3.27 vs 3.17 ms)
constexpr int kPatchBufWords = 37 * kPatchWords;  // one buffer holds either patch
constexpr int kDescSmem = 8 * kDescPerWarp * kPatchBufWords * 4;

template <int ROWS, int WORDS>
__device__ __forceinline__ void stage_patch(unsigned* dst, const u8* src, int pitch, int lane) {
  // ROWS x WORDS/2 8-byte cp.async copies (the LDGSTS issue rate, not bytes, bounds this kernel),
  // 32 per step; (row, chunk) of every step is a compile-time function of the lane
  constexpr int CH = WORDS / 2, N = ROWS * CH;
#pragma unroll
  for (int k = 0; k < (N + 31) / 32; k++) {
    const int i = lane + 32 * k;
    if (32 * (k + 1) <= N || i < N) {
      const int r = i / CH, c = i - r * CH;
      __pipeline_memcpy_async(dst + 2 * i, src + (r * pitch + 8 * c), 8);
    }
  }
  __pipeline_commit();
}

// One CTA handles 32 consecutive output slots of a frame, 4 per warp.
//  A  each warp streams the 31x31 unblurred patches of its keypoints into shared memory
//     (cp.async) and sums the intensity-centroid moments with IDP.4A: lane <-> patch row, the
//     disc mask and the u weights live in registers; then fastAtan2 (IC_Angle :94-141).
//  B  ONE thread per keypoint evaluates cos/sin in double precision and rounds to float (matches
//     glibc's cosf/sinf; per warp this would issue the same FP64 stream 32 times more often).
//  C  each warp streams the 37-row blurred patches (issued before B, so the copies overlap it)
//     and evaluates the 256 steered-BRIEF comparisons from shared memory (:153-204).
__global__ void __launch_bounds__(256) k_describe(const Geom g, const u8* __restrict__ pyr, size_t pyrStride,
                                                  const u8* __restrict__ blur, size_t blurStride,
                                                  const uint2* __restrict__ kept, const int* __restrict__ keptCount,
                                                  int keptTotal, const signed char* __restrict__ pattern,
                                                  orb_keypoint* __restrict__ outK, u8* __restrict__ outD,
                                                  int* __restrict__ outN, int cap, int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned s_buf[];   // [8 warps][kDescPerWarp][kPatchBufWords]
  __shared__ int s_lvl[kDescSlots];
  __shared__ float s_angle[kDescSlots], s_cos[kDescSlots], s_sin[kDescSlots];
  __shared__ int s_prefix[kMaxLevels + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.y;
  const int slot0 = blockIdx.x * kDescSlots;
  unsigned* wbuf = s_buf + wid * (kDescPerWarp * kPatchBufWords);
  // concatenate levels in ascending octave (:1585-1645)
  if (threadIdx.x == 0) {
    int total = 0;
    for (int q = 0; q < g.nlevels; q++) {
      s_prefix[q] = total;
      total += keptCount[f * g.nlevels + q];
    }
    s_prefix[g.nlevels] = total;
    if (blockIdx.x == 0) {
      outN[f] = min(total, cap);
      if (total > cap) atomicOr(overflow, 4);
    }
  }
  // disc mask / (u+15) weights of this lane's patch row v = lane-15: 8 words of 4 bytes
  unsigned w1[8], wu[8];
  {
    const int v = lane - kHalfPatch;
    const int d = lane < 31 ? c_umax[v < 0 ? -v : v] : -1;
    // columns 15-d .. 15+d of the row as a 31-bit mask; a nibble of it becomes 4 bytes of 0 / 1 with one multiply
    // (bit j -> bit 8j: the partial products of 0x00204081 do not collide), the weights (u + 15) with a second one
    const unsigned rowMask = d < 0 ? 0u : ((2u << (2 * d)) - 1u) << (kHalfPatch - d);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned ones = (((rowMask >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u;
      w1[k] = ones;
      wu[k] = (ones * 0xffu) & (0x03020100u + 0x04040404u * (unsigned)k);
    }
  }
  __syncthreads();
  const int total = min(s_prefix[g.nlevels], cap);
  if (slot0 >= total) return;

  // ---- phase A: stage unblurred patches, moments, fastAtan2
  int lvl[kDescPerWarp], px[kDescPerWarp], py[kDescPerWarp], resp[kDescPerWarp];
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    const int slotIdx = slot0 + wid * kDescPerWarp + k;
    lvl[k] = -1; px[k] = py[k] = resp[k] = 0;
    if (slotIdx < total) {
      // octave of output slot slotIdx: the number of level prefixes it has passed (lane q looks at level q)
      const int l = __popc(__ballot_sync(0xffffffffu, lane >= 1 && lane < g.nlevels && slotIdx >= s_prefix[lane]));
      const LevelGeom& L = g.lv[l];
      const uint2 rec = kept[(size_t)f * keptTotal + L.keptOff + (slotIdx - s_prefix[l])];
      lvl[k] = l; px[k] = (int)(rec.x & 0xffffu); py[k] = (int)(rec.x >> 16); resp[k] = (int)rec.y;
      const int xa = (px[k] - kHalfPatch) & ~7;
      stage_patch<31, kMomWords>(wbuf + k * kPatchBufWords,
                                 pyr + (size_t)f * pyrStride + L.off + (long long)(py[k] - kHalfPatch) * L.pitch + xa, L.pitch, lane);
    } else {
      __pipeline_commit();
    }
  }
  float angle[kDescPerWarp];
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    angle[k] = 0.f;
    __pipeline_wait_prior(kDescPerWarp - 1 - k);
    __syncwarp();
    if (lvl[k] >= 0) {
      const int mis = (px[k] - kHalfPatch) & 7;   // patch byte 0 sits `mis` bytes into the staged row
      const unsigned* row = wbuf + k * kPatchBufWords + (lane < 31 ? lane : 30) * kMomWords + (mis >> 2);
      const unsigned sh = (unsigned)(mis & 3) * 8;
      unsigned s1 = 0, s2 = 0;
      unsigned prev = row[0];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const unsigned nxt = row[q + 1];
        const unsigned wv = __funnelshift_r(prev, nxt, sh);   // patch bytes 4q .. 4q+3 of this row
        s1 = __dp4a(wv, w1[q], s1);
        s2 = __dp4a(wv, wu[q], s2);
        prev = nxt;
      }
      const int m10 = __reduce_add_sync(0xffffffffu, (int)s2 - kHalfPatch * (int)s1);   // sum u*I
      const int m01 = __reduce_add_sync(0xffffffffu, (lane - kHalfPatch) * (int)s1);    // sum v*I
      angle[k] = fast_atan2_deg((float)m01, (float)m10);
    }
  }
  __syncwarp();
  // the buffers are free: start streaming the blurred patches, then publish the angles
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    const int sl = wid * kDescPerWarp + k;
    if (lvl[k] >= 0) {
      const LevelGeom& L = g.lv[lvl[k]];
      const int xa = (px[k] - 18) & ~7;
      stage_patch<37, kPatchWords>(wbuf + k * kPatchBufWords,
                                   blur + (size_t)f * blurStride + L.boff + (long long)(py[k] - 18) * L.bpitch + xa, L.bpitch, lane);
    } else {
      __pipeline_commit();
    }
    if (lane == 0) { s_lvl[sl] = lvl[k]; s_angle[sl] = angle[k]; }
  }
  __syncthreads();
  // ---- phase B: cos / sin, one thread per keypoint
  if (threadIdx.x < kDescSlots && s_lvl[threadIdx.x] >= 0) {
    const float ang = __fmul_rn(s_angle[threadIdx.x], (float)(3.14159265358979323846 / 180.f));
    double sn, cs;
    sincos((double)ang, &sn, &cs);
    s_cos[threadIdx.x] = (float)cs;
    s_sin[threadIdx.x] = (float)sn;
  }
  __syncthreads();
  // ---- phase C: steered BRIEF on the blurred patch; lane i produces descriptor byte i
  const int4* pp = reinterpret_cast<const int4*>(pattern) + lane * 2;
  const int4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
  const int words[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    __pipeline_wait_prior(kDescPerWarp - 1 - k);
    __syncwarp();
    if (lvl[k] < 0) continue;
    const int sl = wid * kDescPerWarp + k, slotIdx = slot0 + sl;
    const LevelGeom& L = g.lv[lvl[k]];
    const int x = px[k], y = py[k];
    const float a = s_cos[sl], b = s_sin[sl];
    const u8* cb = reinterpret_cast<const u8*>(wbuf + k * kPatchBufWords) + 18 * (4 * kPatchWords) + (x - ((x - 18) & ~7));
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      const int wd = words[bit];
      const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff);
      const float x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
      const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
      const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
      const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
      const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
      const int t0 = cb[r0 * (4 * kPatchWords) + c0], t1 = cb[r1 * (4 * kPatchWords) + c1];
      val |= (t0 < t1) << bit;
    }
    const size_t o = (size_t)f * cap + slotIdx;
    outD[o * 32 + lane] = (u8)val;
    if (lane == 0) {
      orb_keypoint kp;
      kp.x = lvl[k] ? __fmul_rn((float)x, L.scale) : (float)x;
      kp.y = lvl[k] ? __fmul_rn((float)y, L.scale) : (float)y;
      kp.size = L.patch;
      kp.angle = angle[k];
      kp.response = (float)resp[k];
      kp.octave = lvl[k];
      kp.class_id = -1;
      outK[o] = kp;
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_describe_tma: the same computation with the patches fetched by the TMA unit. Every pyramid level (bordered
// plane) and every blurred level is described by a 3-D tensor map (x, y, frame); ONE lane per keypoint issues
// cp.async.bulk.tensor for the 32x31 moment patch, later for the 48x37 BRIEF patch, and the warp waits on the
// keypoint's mbarrier. That replaces ~380 LDGSTS per keypoint (the issue-rate limiter of k_describe) by two
// instructions. The TMA unit wants the innermost start coordinate on a 16-byte boundary (measured on B200: an
// unaligned byte coordinate raises "illegal instruction"), so the boxes are 48 / 64 bytes wide and start at the
// aligned column at or left of x-15 / x-18.
// ------------------------------------------------------------------------------------------
struct DescMaps {
  CUtensorMap pyr[kMaxLevels];
  CUtensorMap blur[kMaxLevels];
  CUtensorMap blur48[kMaxLevels];   // 48-byte wide boxes: enough when the patch starts <= 11 bytes into its 16-byte chunk
};
constexpr int kTmaMomW = 48, kTmaPatchW = 64;   // 31 / 37 px + up to 15 bytes of alignment slack
constexpr int kTmaMomBytes = kTmaMomW * 31, kTmaPatchBytes = kTmaPatchW * 37;
constexpr int kTmaBufBytes = 2432;   // 64 x 37 rounded up to a multiple of 128 (TMA destination alignment)
constexpr int kDescTmaSmem = kDescSlots * kTmaBufBytes + 128;

__global__ void __launch_bounds__(256) k_describe_tma(const Geom g, const __grid_constant__ DescMaps maps,
                                                      const uint2* __restrict__ kept, const int* __restrict__ keptCount,
                                                      int keptTotal, const signed char* __restrict__ pattern,
                                                      orb_keypoint* __restrict__ outK, u8* __restrict__ outD,
                                                      int* __restrict__ outN, int cap, int* __restrict__ overflow) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[kDescSlots];
  __shared__ int s_prefix[kMaxLevels + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.y;
  // aligned by an OFFSET, not by integer arithmetic on the pointer: the compiler keeps the shared state space (LDS with 32-bit
  // addresses; the former round trip through size_t made every patch load a generic 64-bit LD)
  unsigned char* bufs = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
  unsigned char* wbuf = bufs + wid * (kDescPerWarp * kTmaBufBytes);
  if (threadIdx.x == 0) {
    int total = 0;
    for (int q = 0; q < g.nlevels; q++) {
      s_prefix[q] = total;
      total += keptCount[f * g.nlevels + q];
    }
    s_prefix[g.nlevels] = total;
    if (blockIdx.x == 0) {
      outN[f] = min(total, cap);
      if (total > cap) atomicOr(overflow, 4);
    }
  }
  if (threadIdx.x < kDescSlots) mbar_init(&s_bar[threadIdx.x], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // disc mask / (u+15) weights of this lane's patch row v = lane-15: 8 + 8 words of 4 bytes (computed while the first patch is
  // in flight; a constant table read from global memory measured slower: 3.16 vs 2.88 ms)
  unsigned w1[8], wu[8];
  {
    const int v = lane - kHalfPatch;
    const int d = lane < 31 ? c_umax[v < 0 ? -v : v] : -1;
    // columns 15-d .. 15+d of the row as a 31-bit mask; a nibble of it becomes 4 bytes of 0 / 1 with one multiply
    // (bit j -> bit 8j: the partial products of 0x00204081 do not collide), the weights (u + 15) with a second one
    const unsigned rowMask = d < 0 ? 0u : ((2u << (2 * d)) - 1u) << (kHalfPatch - d);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned ones = (((rowMask >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u;
      w1[k] = ones;
      wu[k] = (ones * 0xffu) & (0x03020100u + 0x04040404u * (unsigned)k);
    }
  }
  __syncthreads();
  const int total = min(s_prefix[g.nlevels], cap);
  // a CTA does kDescRounds rounds of 16 keypoints: the per-lane set-up above (disc masks, level prefix) is paid once
#pragma unroll 1
  for (int round = 0; round < kDescRounds; round++) {
  const int slot0 = (blockIdx.x * kDescRounds + round) * kDescSlots;
  if (slot0 >= total) break;
  const unsigned par = 0u;   // every barrier completes twice per round: parities 0, 1 in every round

  // ---- phase A: fetch the unblurred patches, moments, fastAtan2
  int lvl[kDescPerWarp], px[kDescPerWarp], py[kDescPerWarp], resp[kDescPerWarp];
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    const int slotIdx = slot0 + wid * kDescPerWarp + k;
    lvl[k] = -1; px[k] = py[k] = resp[k] = 0;
    if (slotIdx < total) {
      // octave of output slot slotIdx: the number of level prefixes it has passed (lane q looks at level q)
      const int l = __popc(__ballot_sync(0xffffffffu, lane >= 1 && lane < g.nlevels && slotIdx >= s_prefix[lane]));
      const LevelGeom& L = g.lv[l];
      const uint2 rec = kept[(size_t)f * keptTotal + L.keptOff + (slotIdx - s_prefix[l])];
      lvl[k] = l; px[k] = (int)(rec.x & 0xffffu); py[k] = (int)(rec.x >> 16); resp[k] = (int)rec.y;
      if (lane == 0) {
        unsigned long long* bar = &s_bar[wid * kDescPerWarp + k];
        mbar_expect_tx(bar, kTmaMomBytes);
        tma_load_3d(wbuf + k * kTmaBufBytes, &maps.pyr[l], kLeftPad + ((px[k] - kHalfPatch) & ~15), kEdge + py[k] - kHalfPatch, f, bar);
      }
    }
  }
  float angle[kDescPerWarp];
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    angle[k] = 0.f;
    if (lvl[k] >= 0) {
      mbar_wait(&s_bar[wid * kDescPerWarp + k], par);
      const int mis = (px[k] - kHalfPatch) & 15;   // patch byte 0 sits `mis` bytes into the fetched row
      const unsigned char* rowb = wbuf + k * kTmaBufBytes + (lane < 31 ? lane : 30) * kTmaMomW;
      const unsigned sh = (unsigned)(mis & 3) * 8;
      unsigned s1 = 0, s2 = 0;
      {
        // the whole 48-byte row with three 16-byte loads (rows are 12 words apart: a quarter warp covers all 32 banks, no
        // conflicts; one word per load was a 4-way conflict), then a warp-uniform shift by mis / 4 words
        const uint4* r4 = reinterpret_cast<const uint4*>(rowb);
        const uint4 A = r4[0], B = r4[1], C = r4[2];
        unsigned W[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
        if (mis & 8) {
#pragma unroll
          for (int j = 0; j < 10; j++) W[j] = W[j + 2];
        }
        if (mis & 4) {
#pragma unroll
          for (int j = 0; j < 9; j++) W[j] = W[j + 1];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const unsigned wv = __funnelshift_r(W[q], W[q + 1], sh);   // patch bytes 4q .. 4q+3 of this row
          s1 = __dp4a(wv, w1[q], s1);
          s2 = __dp4a(wv, wu[q], s2);
        }
      }
      const int m10 = __reduce_add_sync(0xffffffffu, (int)s2 - kHalfPatch * (int)s1);   // sum u*I
      const int m01 = __reduce_add_sync(0xffffffffu, (lane - kHalfPatch) * (int)s1);    // sum v*I
      angle[k] = fast_atan2_deg((float)m01, (float)m10);
    }
  }
  __syncwarp();
  // the buffers are free: fetch the blurred patches
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    const int sl = wid * kDescPerWarp + k;
    if (lane == 0 && lvl[k] >= 0) {
      mbar_expect_tx(&s_bar[sl], kTmaPatchBytes);
      tma_load_3d(wbuf + k * kTmaBufBytes, &maps.blur[lvl[k]], (px[k] - 18) & ~15, py[k] - 18, f, &s_bar[sl]);
    }
  }
  // ---- phase B: cos / sin in double precision, rounded to float (glibc's cosf / sinf): lane k of the warp does the
  // warp's keypoint k while the patch is in flight. No CTA-wide step: every warp runs on its own from here (a single
  // warp doing all 16 keypoints of the CTA kept the other seven waiting at two barriers: 13 % of the stall samples).
  float cosMine = 0.f, sinMine = 0.f;
  {
    float myAngle = 0.f;
    bool mine = false;
#pragma unroll
    for (int k = 0; k < kDescPerWarp; k++)
      if (lane == k) { myAngle = angle[k]; mine = lvl[k] >= 0; }
    if (mine) {
      const float ang = __fmul_rn(myAngle, (float)(3.14159265358979323846 / 180.f));
      double sn, cs;
      sincos((double)ang, &sn, &cs);
      cosMine = (float)cs;
      sinMine = (float)sn;
    }
  }
  __syncwarp();
  // ---- phase C: steered BRIEF on the blurred patch; lane i produces descriptor byte i (its 8 test pairs are re-read
  // per round, so that they do not stay in registers beside the moment masks)
  const int4* pp = reinterpret_cast<const int4*>(pattern) + lane * 2;
  const int4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
  const int words[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
  for (int k = 0; k < kDescPerWarp; k++) {
    if (lvl[k] < 0) continue;
    const int sl = wid * kDescPerWarp + k, slotIdx = slot0 + sl;
    mbar_wait(&s_bar[sl], par ^ 1u);
    const LevelGeom& L = g.lv[lvl[k]];
    const int x = px[k], y = py[k];
    const float a = __shfl_sync(0xffffffffu, cosMine, k), b = __shfl_sync(0xffffffffu, sinMine, k);
    const u8* cb = wbuf + k * kTmaBufBytes + 18 * kTmaPatchW + 18 + ((x - 18) & 15);
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      const int wd = words[bit];
      const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff);
      const float x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
      const float fr0 = __fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)), fc0 = __fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b));
      const float fr1 = __fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)), fc1 = __fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b));
      int t0, t1;
      {
        // cvRound = round to nearest even = what adding 1.5 * 2^23 does to the mantissa (|v| < 2^22): the integer is the low
        // bits of the sum. Keeps the conversions off the 16-lane XU pipe; the bias 65 * 0x4B400000 leaves with the base address.
        constexpr float kMagic = 12582912.f;
        constexpr unsigned kBias = 65u * 0x4B400000u;
        const unsigned i0 = (unsigned)__float_as_int(__fadd_rn(fr0, kMagic)) * (unsigned)kTmaPatchW + (unsigned)__float_as_int(__fadd_rn(fc0, kMagic));
        const unsigned i1 = (unsigned)__float_as_int(__fadd_rn(fr1, kMagic)) * (unsigned)kTmaPatchW + (unsigned)__float_as_int(__fadd_rn(fc1, kMagic));
        t0 = cb[(int)(i0 - kBias)]; t1 = cb[(int)(i1 - kBias)];
      }
      val |= (t0 < t1) << bit;
    }
    const size_t o = (size_t)f * cap + slotIdx;
    outD[o * 32 + lane] = (u8)val;
    if (lane == 0) {
      orb_keypoint kp;
      kp.x = lvl[k] ? __fmul_rn((float)x, L.scale) : (float)x;
      kp.y = lvl[k] ? __fmul_rn((float)y, L.scale) : (float)y;
      kp.size = L.patch;
      kp.angle = angle[k];
      kp.response = (float)resp[k];
      kp.octave = lvl[k];
      kp.class_id = -1;
      outK[o] = kp;
    }
  }
  __syncwarp();
  }   // round
}

// ------------------------------------------------------------------------------------------
// k_describe_ring: the arithmetic of k_describe_tma as a per-warp software pipeline. A warp owns KPW consecutive output
// slots and no CTA-wide step exists. ncu on k_describe_tma (16 keypoints per CTA, two per warp) showed a chain of four
// dependent memory latencies per CTA (level counts -> kept record -> moment patch -> blurred patch) that only other CTAs
// could hide (58 % of the warp slots active), and lane-serial work done per keypoint: cos / sin in double precision on 2
// of 32 lanes, fastAtan2 and the keypoint record on one. Here
//   * lane j loads the record of the warp's keypoint j (one coalesced load) and finds its level from a warp scan of the
//     level counts;
//   * phase A streams the 48 x 31 moment patches through a ring of three buffers (TMA two keypoints ahead), lane j keeps
//     (m01, m10) of keypoint j;
//   * fastAtan2, cos / sin run once for all KPW keypoints, one per lane, while the first blurred patches are in flight;
//   * phase C streams the 64 x 37 blurred patches through a ring of two buffers in the same memory (TMA one keypoint ahead);
//   * lane j writes the keypoint record of keypoint j.
// The moment rows are read with three 16-byte loads per lane (conflict free; one word per load was a 4-way bank conflict)
// and the BRIEF coordinates are rounded by the magic-number addition instead of F2I (the 16-lane XU pipe was 67 % busy).
// ------------------------------------------------------------------------------------------
constexpr int kMomBufBytes = 1536;                 // 48 x 31 rounded up to a multiple of 128
// per warp: DEPTH blurred patches or DEPTH + 1 moment patches in the same memory
constexpr int desc_ring_smem(int depth) { return 8 * depth * kTmaBufBytes + 128; }

template <int KPW, int MINB, int DEPTH>
__global__ void __launch_bounds__(256, MINB) k_describe_ring(const Geom g, const __grid_constant__ DescMaps maps,
                                                       const uint2* __restrict__ kept, const int* __restrict__ keptCount,
                                                       int keptTotal, const signed char* __restrict__ pattern,
                                                       orb_keypoint* __restrict__ outK, u8* __restrict__ outD,
                                                       int* __restrict__ outN, int cap, int* __restrict__ overflow) {
  static_assert(KPW >= 2 && KPW <= 32, "one lane per keypoint");
  constexpr int kRingBytes = DEPTH * kTmaBufBytes, NA = DEPTH + 1;   // NA moment buffers / barriers
  static_assert(NA * kMomBufBytes <= kRingBytes, "moment ring must fit in the blurred-patch ring");
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[8 * NA];
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.y;
  unsigned char* ring = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u) + wid * kRingBytes;
  unsigned long long* bar = s_bar + wid * NA;
  const unsigned ringA = smem_u32(ring), barA = smem_u32(bar);   // shared-window addresses for the TMA / mbarrier instructions
  // level prefix by a warp scan: lane q holds the number of kept keypoints on levels < q
  const int cnt = lane < g.nlevels ? keptCount[f * g.nlevels + lane] : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < kMaxLevels; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  const int excl = incl - cnt;
  const int totalAll = __shfl_sync(kFull, incl, kMaxLevels - 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    outN[f] = min(totalAll, cap);
    if (totalAll > cap) atomicOr(overflow, 4);
  }
  const int base = (blockIdx.x * 8 + wid) * KPW;
  const int n = min(KPW, min(totalAll, cap) - base);
  if (n <= 0) return;   // the whole warp leaves; nothing below is CTA-wide
  if (lane < NA) mbar_init(&bar[lane], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  // lane j: record of keypoint base + j (lanes >= n repeat the last one; they never write)
  int myL = 0, myResp;
  unsigned myRec;   // x | y << 14 | level << 28 (coordinates < 16384: the host launches k_describe_tma for larger images)
  {
    const int slot = base + min(lane, n - 1);
#pragma unroll 1
    for (int q = 1; q < g.nlevels; q++) myL += slot >= __shfl_sync(kFull, excl, q);
    const int first = __shfl_sync(kFull, excl, myL);
    const uint2 rec = kept[(size_t)f * keptTotal + g.lv[myL].keptOff + (slot - first)];
    myRec = (rec.x & 0x3fffu) | ((rec.x >> 16) << 14) | ((unsigned)myL << 28);
    myResp = (int)rec.y;
  }
  auto fetch_moment = [&](int j, int slot) {
    const unsigned rec = __shfl_sync(kFull, myRec, j);
    if (lane == 0) {
      const int x = (int)(rec & 0x3fffu), y = (int)((rec >> 14) & 0x3fffu);
      const unsigned b = barA + 8u * slot;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kTmaMomBytes) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(ringA + slot * kMomBufBytes), "l"(&maps.pyr[rec >> 28]), "r"(kLeftPad + ((x - kHalfPatch) & ~15)),
                     "r"(kEdge + y - kHalfPatch), "r"(f), "r"(b) : "memory");
    }
  };
  // the blurred patch: columns x-18 .. x+18 from the 16-byte chunk at or left of x-18; a 48-byte box holds them when the
  // patch starts <= 11 bytes into the chunk (3 of 4 keypoints: fewer L2 sectors, and rows 12 words apart spread the BRIEF
  // samples over all banks where 16 words fold them onto two sets of 16)
  auto fetch_blurred = [&](int j, int slot) {
    const unsigned rec = __shfl_sync(kFull, myRec, j);
    if (lane == 0) {
      const int x = (int)(rec & 0x3fffu), y = (int)((rec >> 14) & 0x3fffu);
      const bool narrow = ((x - 18) & 15) <= 11;
      const unsigned b = barA + 8u * slot;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((narrow ? 48 : kTmaPatchW) * 37) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(ringA + slot * kTmaBufBytes), "l"(narrow ? &maps.blur48[rec >> 28] : &maps.blur[rec >> 28]), "r"((x - 18) & ~15),
                     "r"(y - 18), "r"(f), "r"(b) : "memory");
    }
  };
  auto wait_bar = [&](int slot, unsigned par) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(barA + 8u * slot), "r"(par) : "memory");
  };
#pragma unroll
  for (int j = 0; j < DEPTH; j++)
    if (j < n) fetch_moment(j, j);
  unsigned parity = 0;   // bit b: the phase bar[b] completes next
  // ---- phase A: intensity-centroid moments of every keypoint
  int myM10 = 0, myM01 = 0;
  {
    // disc mask / (u + 15) weights of this lane's patch row v = lane - 15: 8 + 8 words of 4 bytes (computed, not loaded: the
    // first patches are in flight meanwhile, and a table read from global memory measured slower)
    unsigned w1[8], wu[8];
    {
      const int v = lane - kHalfPatch;
      const int d = lane < 31 ? c_umax[v < 0 ? -v : v] : -1;
      const unsigned rowMask = d < 0 ? 0u : ((2u << (2 * d)) - 1u) << (kHalfPatch - d);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const unsigned ones = (((rowMask >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u;
        w1[k] = ones;
        wu[k] = (ones * 0xffu) & (0x03020100u + 0x04040404u * (unsigned)k);
      }
    }
    int rs = 0;   // ring slot of keypoint j = j % NA
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      if (j + DEPTH < n) fetch_moment(j + DEPTH, rs == 0 ? NA - 1 : rs - 1);   // the slot keypoint j - 1 was read from, one __syncwarp ago
      wait_bar(rs, (parity >> rs) & 1u);
      parity ^= 1u << rs;
      const int mis = ((int)(__shfl_sync(kFull, myRec, j) & 0x3fffu) - kHalfPatch) & 15;   // patch byte 0 sits `mis` bytes into the fetched row
      const uint4* r4 = reinterpret_cast<const uint4*>(ring + rs * kMomBufBytes + (lane < 31 ? lane : 30) * kTmaMomW);
      const uint4 A = r4[0], B = r4[1], C = r4[2];
      unsigned W[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
      if (mis & 8) {
#pragma unroll
        for (int q = 0; q < 10; q++) W[q] = W[q + 2];
      }
      if (mis & 4) {
#pragma unroll
        for (int q = 0; q < 9; q++) W[q] = W[q + 1];
      }
      const unsigned sh = (unsigned)(mis & 3) * 8;
      unsigned s1 = 0, s2 = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const unsigned wv = __funnelshift_r(W[q], W[q + 1], sh);   // patch bytes 4q .. 4q+3 of this row
        s1 = __dp4a(wv, w1[q], s1);
        s2 = __dp4a(wv, wu[q], s2);
      }
      // sum u*I, sum v*I over the rows: one REDUX each
      const int m10 = __reduce_add_sync(kFull, (int)s2 - kHalfPatch * (int)s1);
      const int m01 = __reduce_add_sync(kFull, (lane - kHalfPatch) * (int)s1);
      if (lane == j) { myM10 = m10; myM01 = m01; }
      __syncwarp();
      rs = rs == NA - 1 ? 0 : rs + 1;
    }
  }
  // ---- the ring is free (every moment row was read before the last __syncwarp): start on the blurred patches
#pragma unroll
  for (int j = 0; j < DEPTH; j++)
    if (j < n) fetch_blurred(j, j);
  // ---- phase B, one keypoint per lane: fastAtan2; cos / sin in double precision, rounded to float (glibc's cosf / sinf)
  const float myAngle = fast_atan2_deg((float)myM01, (float)myM10);
  float myCos, mySin;
  {
    const float ang = __fmul_rn(myAngle, (float)(3.14159265358979323846 / 180.f));
    double sn, cs;
    sincos((double)ang, &sn, &cs);
    myCos = (float)cs;
    mySin = (float)sn;
  }
  // ---- phase C: steered BRIEF on the blurred patch; lane i produces descriptor byte i
  {
    const int4* pp = reinterpret_cast<const int4*>(pattern) + lane * 2;
    const int4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    const int words[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    int bs = 0;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      wait_bar(bs, (parity >> bs) & 1u);
      parity ^= 1u << bs;
      const float a = __shfl_sync(kFull, myCos, j), b = __shfl_sync(kFull, mySin, j);
      const int m = ((int)(__shfl_sync(kFull, myRec, j) & 0x3fffu) - 18) & 15;
      const unsigned pitch = m <= 11 ? 48u : (unsigned)kTmaPatchW;
      // cvRound = round to nearest even = what adding 1.5 * 2^23 does to the mantissa (|v| < 2^22): the integer is in the low
      // bits of the sum; the bias (pitch + 1) * 0x4B400000 of (row * pitch + column) leaves with the base address
      constexpr float kMagic = 12582912.f;
      // (a warp-wide reduction hands the compiler a value it knows to be uniform: the base goes into a uniform register and
      // the loads address [index + base] without an add each)
      const unsigned ubase = __reduce_max_sync(kFull, (unsigned)(bs * kTmaBufBytes + 18 + m) + 18u * pitch - (pitch + 1u) * 0x4B400000u);
      const u8* cb = ring + (int)ubase;
      int val = 0;
#pragma unroll
      for (int bit = 0; bit < 8; bit++) {
        const int wd = words[bit];
        const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff);
        const float x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
        const float fr0 = __fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)), fc0 = __fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b));
        const float fr1 = __fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)), fc1 = __fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b));
        const unsigned i0 = (unsigned)__float_as_int(__fadd_rn(fr0, kMagic)) * pitch + (unsigned)__float_as_int(__fadd_rn(fc0, kMagic));
        const unsigned i1 = (unsigned)__float_as_int(__fadd_rn(fr1, kMagic)) * pitch + (unsigned)__float_as_int(__fadd_rn(fc1, kMagic));
        const int t0 = cb[(int)i0], t1 = cb[(int)i1];
        val |= (t0 < t1) << bit;
      }
      outD[((size_t)f * cap + base + j) * 32 + lane] = (u8)val;
      __syncwarp();   // every lane has read its samples: the buffer may be refilled
      if (j + DEPTH < n) fetch_blurred(j + DEPTH, bs);
      bs = bs == DEPTH - 1 ? 0 : bs + 1;
    }
  }
  if (lane < n) {
    const LevelGeom& L = g.lv[myL];
    orb_keypoint kp;
    const int myX = (int)(myRec & 0x3fffu), myY = (int)((myRec >> 14) & 0x3fffu);
    kp.x = myL ? __fmul_rn((float)myX, L.scale) : (float)myX;
    kp.y = myL ? __fmul_rn((float)myY, L.scale) : (float)myY;
    kp.size = L.patch;
    kp.angle = myAngle;
    kp.response = (float)myResp;
    kp.octave = myL;
    kp.class_id = -1;
    outK[(size_t)f * cap + base + lane] = kp;
  }
}

typedef void (*DescribeTmaFn)(const Geom, const DescMaps, const uint2*, const int*, int, const signed char*, orb_keypoint*, u8*, int*, int, int*);
// Measured (KITTI, 2048 frames): 8 keypoints per warp 2.33 ms, 4: 2.76, 16: 2.47; a ring one buffer deeper: 2.76; a register
// budget for 4 CTAs per SM (64 registers, 4 bytes spilled): 4.0 ms.
constexpr int kDescRingDepth = 2;
static DescribeTmaFn describe_ring_variant(int kpw) {
  return kpw == 4 ? k_describe_ring<4, 3, kDescRingDepth> : (kpw == 16 ? k_describe_ring<16, 3, kDescRingDepth> : k_describe_ring<8, 3, kDescRingDepth>);
}

// ------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (Frame.cc:831-1082): one CTA per stereo pair (frames 2p, 2p+1 of the
// chunk), one warp per left keypoint.
//   coarse : best Hamming among right keypoints whose row band (+-2*scale) holds the left row, within
//            one octave and the disparity range (first wins = lowest right index) (:867-948)
//   refine : 11x11 SAD slide of +-5 px on the keypoint's pyramid level, parabola fit (:951-1064)
//   prune  : drop matches whose SAD is >= 1.5*1.4*median (:1069-1081)
// ------------------------------------------------------------------------------------------
constexpr int kStereoThreads = 512;

__device__ __forceinline__ int hamming_words(const unsigned a[8], const uint4 lo, const uint4 hi) {
  return __popc(a[0] ^ lo.x) + __popc(a[1] ^ lo.y) + __popc(a[2] ^ lo.z) + __popc(a[3] ^ lo.w) + __popc(a[4] ^ hi.x) +
         __popc(a[5] ^ hi.y) + __popc(a[6] ^ hi.z) + __popc(a[7] ^ hi.w);
}

__global__ void __launch_bounds__(kStereoThreads) k_stereo(const Geom g, const u8* __restrict__ pyr, size_t pyrStride,
                                                           const orb_keypoint* __restrict__ kps, const u8* __restrict__ desc,
                                                           const int* __restrict__ counts, int cap, float mbf, float mb,
                                                           const float* __restrict__ invScale, float* __restrict__ uRight,
                                                           float* __restrict__ depth, int entCap, int2* __restrict__ gsad,
                                                           int* __restrict__ gcnt) {
  // gridDim.y = G > 1 (few pairs per call: the latency case): the left keypoints of a pair are dealt to G CTAs, matches go
  // to a global list, and the CTA that finishes last (ticket) does the median cut over all of them.
  extern __shared__ __align__(16) unsigned char ssm[];
  __shared__ int s_n;
  __shared__ int s_scan[34];
  __shared__ float s_median;
  __shared__ int s_last;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.y, part = blockIdx.y;
  constexpr int kW = kStereoThreads / 32;
  const int fL = 2 * pair, fR = 2 * pair + 1;
  const int N = counts[fL], Nr = counts[fR];
  float* rx = reinterpret_cast<float*>(ssm);                       // right keypoint x
  int* rband = reinterpret_cast<int*>(rx + cap);                   // minr | maxr << 16
  int2* sad = reinterpret_cast<int2*>(rband + cap);                // (SAD best, left index)
  int* rowStart = reinterpret_cast<int*>(sad + cap);               // [nRows + 1] row index of the right keypoints
  unsigned short* ent = reinterpret_cast<unsigned short*>(rowStart + (g.H + 2));   // vRowIndices, flattened (:846-865)
  u8* roct = reinterpret_cast<u8*>(ent + entCap);
  const int nRows = g.H;
  const orb_keypoint* KL = kps + (size_t)fL * cap;
  const orb_keypoint* KR = kps + (size_t)fR * cap;
  const u8* DL = desc + (size_t)fL * cap * 32;
  const u8* DR = desc + (size_t)fR * cap * 32;
  float* uR = uRight + (size_t)pair * cap;
  float* dp = depth + (size_t)pair * cap;
  for (int i = tid; i < cap; i += kStereoThreads)
    if ((i % (kW * G)) / kW == part) { uR[i] = -1.0f; dp[i] = -1.0f; }   // every entry has one owner CTA
  for (int i = tid; i < Nr; i += kStereoThreads) {
    const orb_keypoint k = KR[i];
    const float r = __fmul_rn(2.0f, g.lv[k.octave].scale);         // :858
    const int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));
    rx[i] = k.x;
    rband[i] = (max(minr, 0) & 0xffff) | (min(maxr, 0xffff) << 16);
    roct[i] = (u8)k.octave;
  }
  for (int i = tid; i <= nRows; i += kStereoThreads) rowStart[i] = 0;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // vRowIndices: count, scan, fill (order inside a row is irrelevant: ties are broken by the packed key)
  for (int i = tid; i < Nr; i += kStereoThreads) {
    const int band = rband[i];
    for (int y = band & 0xffff; y <= min(band >> 16, nRows - 1); y++) atomicAdd(&rowStart[y], 1);
  }
  __syncthreads();
  const int total = block_scan_excl(rowStart, nRows + 1, s_scan);
  const bool indexed = total <= entCap;     // otherwise fall back to scanning every right keypoint
  if (indexed) {
    // rowStart[y] now holds the start of row y; fill using a second cursor array aliased on `sad`
    int* cursor = reinterpret_cast<int*>(sad);
    for (int i = tid; i < nRows; i += kStereoThreads) cursor[i] = rowStart[i];
    __syncthreads();
    for (int i = tid; i < Nr; i += kStereoThreads) {
      const int band = rband[i];
      for (int y = band & 0xffff; y <= min(band >> 16, nRows - 1); y++) ent[atomicAdd(&cursor[y], 1)] = (unsigned short)i;
    }
  }
  __syncthreads();
  const float maxD = __fdiv_rn(mbf, mb);   // minZ = mb, minD = 0 (:885-887)
  const int TH_HIGH = 100, thOrbDist = 75;
  for (int iL = wid + kW * part; iL < N; iL += kW * G) {
    const orb_keypoint kl = KL[iL];
    const int levelL = kl.octave, row = (int)kl.y;
    const float uL = kl.x, minU = __fsub_rn(uL, maxD), maxU = uL;
    if (maxU < 0.f) continue;
    const uint4* dl = reinterpret_cast<const uint4*>(DL + (size_t)iL * 32);
    const uint4 dlo = __ldg(dl), dhi = __ldg(dl + 1);
    const unsigned a[8] = {dlo.x, dlo.y, dlo.z, dlo.w, dhi.x, dhi.y, dhi.z, dhi.w};
    unsigned key = 0xffffffffu;
    if (indexed) {
      if (row < 0 || row >= nRows) continue;
      const int e1 = rowStart[row + 1];
      for (int e = rowStart[row] + lane; e < e1; e += 32) {
        const int iR = ent[e], o = roct[iR];
        const float x = rx[iR];
        if (o >= levelL - 1 && o <= levelL + 1 && x >= minU && x <= maxU) {
          const uint4* dr = reinterpret_cast<const uint4*>(DR + (size_t)iR * 32);
          const int d = hamming_words(a, __ldg(dr), __ldg(dr + 1));
          if (d < TH_HIGH) key = min(key, ((unsigned)d << 16) | (unsigned)iR);
        }
      }
    } else {
      for (int iR = lane; iR < Nr; iR += 32) {
        const int band = rband[iR], o = roct[iR];
        const float x = rx[iR];
        if (row >= (band & 0xffff) && row <= (band >> 16) && o >= levelL - 1 && o <= levelL + 1 && x >= minU && x <= maxU) {
          const uint4* dr = reinterpret_cast<const uint4*>(DR + (size_t)iR * 32);
          const int d = hamming_words(a, __ldg(dr), __ldg(dr + 1));
          if (d < TH_HIGH) key = min(key, ((unsigned)d << 16) | (unsigned)iR);
        }
      }
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (key == 0xffffffffu || (int)(key >> 16) >= thOrbDist) continue;
    const int bestIdxR = (int)(key & 0xffffu);
    // ---- SAD refinement on pyramid level levelL
    const LevelGeom& L = g.lv[levelL];
    const float sf = invScale[levelL];
    const float scaleduL = roundf(__fmul_rn(kl.x, sf)), scaledvL = roundf(__fmul_rn(kl.y, sf));
    const float scaleduR0 = roundf(__fmul_rn(rx[bestIdxR], sf));
    const float iniu = scaleduR0 - 10.f, endu = scaleduR0 + 11.f;   // L + w = 10 (this fork: :990-992)
    if (iniu < 0.f || endu >= (float)L.w) continue;
    const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
    const u8* IL = pyr + (size_t)fL * pyrStride + L.off + (long long)cv * L.pitch + cu;
    const u8* IR = pyr + (size_t)fR * pyrStride + L.off + (long long)cv * L.pitch + cr;
    const int cL = IL[0];
    // |(IL - cL) - (IR_q - cR_q)| = |(IL + (cR_q - cL)) - IR_q|: one add and one SAD instruction per pixel and shift
    int acc[11], kq[11];
#pragma unroll
    for (int q = 0; q < 11; q++) { acc[q] = 0; kq[q] = (int)IR[q - 5] - cL; }
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int pI = lane + 32 * t;
      if (pI < 121) {
        const int dy = pI / 11 - 5, dx = pI % 11 - 5;
        const int il = (int)IL[dy * L.pitch + dx];
        const u8* rrow = IR + dy * L.pitch + dx;
#pragma unroll
        for (int q = 0; q < 11; q++) acc[q] = (int)__sad(il + kq[q], (int)rrow[q - 5], (unsigned)acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 11; q++) acc[q] = __reduce_add_sync(0xffffffffu, acc[q]);   // one REDUX per shift (ten instructions as a shuffle tree)
    if (lane == 0) {
      int best = 0x7fffffff, bestinc = 0;
#pragma unroll
      for (int q = 0; q < 11; q++)
        if (acc[q] < best) { best = acc[q]; bestinc = q - 5; }
      if (bestinc != -5 && bestinc != 5) {
        float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int q = 1; q < 10; q++)
          if (q - 5 == bestinc) { d1 = (float)acc[q - 1]; d2 = (float)acc[q]; d3 = (float)acc[q + 1]; }
        const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
        if (!(deltaR < -1.f || deltaR > 1.f)) {
          float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          if (disparity >= 0.f && disparity < maxD) {
            if (disparity <= 0.f) {
              disparity = (float)0.01;
              bestuR = (float)((double)uL - 0.01);
            }
            dp[iL] = __fdiv_rn(mbf, disparity);
            uR[iL] = bestuR;
            if (G == 1) sad[atomicAdd(&s_n, 1)] = make_int2(best, iL);
            else gsad[(size_t)pair * cap + atomicAdd(&gcnt[2 * pair], 1)] = make_int2(best, iL);
          }
        }
      }
    }
  }
  if (G > 1) __threadfence();   // this CTA's matches are visible device-wide before its ticket is drawn
  __syncthreads();
  if (G > 1) {
    if (tid == 0) s_last = atomicAdd(&gcnt[2 * pair + 1], 1) == G - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int ng = *reinterpret_cast<volatile int*>(&gcnt[2 * pair]);
    for (int i = tid; i < ng; i += kStereoThreads) sad[i] = __ldcg(&gsad[(size_t)pair * cap + i]);
    if (tid == 0) s_n = ng;
    __syncthreads();
  }
  const int n = s_n;
  if (n == 0) return;
  // median of the SAD values = element n/2 of the sorted list (only the value matters): two-pass
  // radix select on shared-memory histograms (SAD <= 121 * 1020 < 2^17: 9 high bits, then 8 low bits)
  {
    int* hist = rowStart;   // the row index is no longer needed; >= 512 ints (nRows + 1 >= 512 is NOT
                            // guaranteed, so the histogram lives in the (larger) entry array when needed)
    if (nRows + 1 < 512) hist = reinterpret_cast<int*>(ent);
    const int kth = n / 2;
    for (int i = tid; i < 512; i += kStereoThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kStereoThreads) atomicAdd(&hist[min(sad[i].x >> 8, 511)], 1);
    __syncthreads();
    if (tid == 0) {
      int acc = 0, b = 0;
      for (; b < 512; b++) { if (acc + hist[b] > kth) break; acc += hist[b]; }
      s_scan[0] = b;          // bin of the median
      s_scan[1] = kth - acc;  // rank inside the bin
    }
    __syncthreads();
    const int bin = s_scan[0], rank = s_scan[1];
    for (int i = tid; i < 256; i += kStereoThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kStereoThreads) {
      const int v = sad[i].x;
      if (min(v >> 8, 511) == bin) atomicAdd(&hist[v & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int acc = 0, b = 0;
      for (; b < 256; b++) { if (acc + hist[b] > rank) break; acc += hist[b]; }
      s_median = (float)((bin << 8) | b);
    }
    __syncthreads();
  }
  const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), s_median);
  for (int i = tid; i < n; i += kStereoThreads) {
    if (!((float)sad[i].x < thDist)) {
      uR[sad[i].y] = -1.0f;
      dp[sad[i].y] = -1.0f;
    }
  }
}

inline int cv_round_f(float v) { return (int)lrintf(v); }

}  // namespace

// ==========================================================================================
// Host side
// ==========================================================================================
struct orb_extractor {
  orb_params p;
  int device = 0;
  int maxBatch = 1;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> perLevel;
  cudaStream_t stream = nullptr;

  // geometry of the current image size
  Geom g;
  bool haveGeom = false;
  size_t pyrStride = 0, blurStride = 0;
  int candTotal = 0, keptTotal = 0, nodeCap = 0, maxKp = 0;
  size_t qtSmem = 0;
  int qtSeqWords = 0;
  // k_fast_cells: one launch per group of levels (see FastGroup)
  int fastGroups = 1;
  FastGroup fastGroup[2] = {};
  FastSmemLayout fastLayG[2] = {};
  size_t fastSmemG[2] = {0, 0};
  int fastWarpsG[2] = {4, 4}, fastBlocksG[2] = {0, 0};
  BorderJobs borderJobs = {};    // block ranges of k_fill_borders
  // k_pyramid_fused (few frames per call): item table + dependency flags / work counter
  std::vector<PyrItem> pyrItems;
  PyrPlan pyrPlan = {};
  PyrItem* d_pyrItems = nullptr;
  int* d_pyrFlags = nullptr;     // kPyrFusedMaxFrames x bands flags, then the work counter
  bool pyrFused = false;         // ORB_B200_PYR_FUSED=1: the single-launch pyramid for <= 4 frames (measured slower: 56 vs 47 us)
  int numSMs = 148;
  int borderBlocks = 0;
  bool pyrTiled = true;          // k_level0_border2 + k_resize_strip + k_fill_borders (ORB_B200_PYR=0: the first-round kernels)
  std::vector<int2> taps;

  // device workspace (sized for maxBatch frames of the current geometry)
  u8* d_pyr = nullptr; u8* d_blur = nullptr;
  uint2* d_cand = nullptr; int* d_candCount = nullptr; unsigned short* d_keyNode = nullptr;
  uint2* d_kept = nullptr; int* d_keptCount = nullptr; int2* d_taps = nullptr;
  signed char* d_pattern = nullptr; int* d_overflow = nullptr; int* d_work = nullptr;
  int wsFrames = 0;
  // two workspace lanes (orb_set_lanes / ORB_B200_LANES=1 for one): consecutive chunks of a batch call run on two
  // streams, so the tail of every kernel of one chunk and the latency-bound phases (quadtree) overlap the other chunk's
  // kernels: +6.7 % frames/s, +14 % stereo pairs/s on 256-frame chunks (measured, round 2). Single-call entry points
  // use lane 0 only.
  int lanes = 2, lastLane = 0;
  long long hostChunks = 0;      // chunks issued by the host batch entry points so far (staging buffer parity)
  bool asyncPending = false;     // orb_extract_batch_host_async work may still be in flight
  DescMaps descMaps[2];          // TMA tensor maps of the two workspace lanes (k_describe_tma)
  int descRing = 8;   // keypoints per warp of k_describe_ring (4, 8, 16); 0 = k_describe_tma
  bool pdl = true;   // per-frame calls: the resize chain as programmatic dependent launches (ORB_B200_PDL=0: plain launches)
  int rzSmallFrom = 3, rzSmallBand = 16;   // k_resize_strip: levels >= rzSmallFrom use bands of rzSmallBand rows (measured: pyramid 2.77 -> 2.71 ms)
  PyrMaps blurMaps[2];           // ... and of the blur input tiles (k_blur7)
  bool blurTma = false;
  bool descTma = false;          // maps are valid for the current workspace
  cudaStream_t laneStream[2] = {nullptr, nullptr};
  cudaEvent_t evFork = nullptr, evJoin[2] = {nullptr, nullptr};
  // k_blur7 needs only the pyramid: with blurFork != 0 it runs on a side stream of the lane, beside k_quadtree
  // (blurFork = 2, the latency-bound kernel of the chain) or beside k_fast_cells + k_quadtree (blurFork = 1), and
  // k_describe waits for both (ORB_B200_BLUR_FORK; off while the per-stage timing of the bench is on)
  int blurFork = 0;
  cudaStream_t blurStream[2] = {nullptr, nullptr};
  cudaEvent_t evBlurGo[2] = {nullptr, nullptr}, evBlurDone[2] = {nullptr, nullptr};
  // k_level0_border (a copy, HBM-bound) runs on the same side stream beside the resize chain (ALU-bound): level 1 is
  // made from the input image itself, the pyramid's level 0 is only needed from k_fast_cells on. Measured +0.5 %
  // device-resident and nothing end to end, so it is a switch (ORB_B200_L0_FORK=1), off by default and off while the
  // per-stage timing of the bench is on
  int l0Fork = 0;
  cudaEvent_t evL0Go[2] = {nullptr, nullptr}, evL0Done[2] = {nullptr, nullptr};
  // staging for the host entry points: two sets, so that the H2D copy of chunk i+1 and the D2H
  // copy of chunk i-1 overlap the kernels of chunk i (copy streams + events)
  int fastTailRun = kFastTailRun, fastTailMul = 1, fastFlags = 1;   // bit 0: skip the 8-pixel prefilter (measured: -5 % FAST)
  u8* d_in[2] = {nullptr, nullptr}; size_t d_inBytes = 0;
  orb_keypoint* d_kps[2] = {nullptr, nullptr}; u8* d_desc[2] = {nullptr, nullptr}; int* d_n[2] = {nullptr, nullptr};
  int stageFrames = 0, stageCap = 0;
  cudaStream_t sIn = nullptr, sOut = nullptr;
  cudaEvent_t evIn[2] = {nullptr, nullptr}, evDone[2] = {nullptr, nullptr}, evOut[2] = {nullptr, nullptr};
  u8* hostPyr = nullptr; size_t hostPyrBytes = 0;   // pinned copy of the last single-call pyramid (mvImagePyramid views)
  // pinned staging of the single-call entry points (orb_extract / orb_extract_stereo): the image goes up from here and
  // counts, overflow flag, keypoints, descriptors (and the stereo vectors) come back in one batch of async copies
  // followed by ONE stream synchronisation, instead of pageable copies with a round trip each
  u8* h_in = nullptr; size_t h_inBytes = 0;
  u8* h_out = nullptr; size_t h_outBytes = 0;
  // ... and their kernel sequence (2 memsets + 14 launches, all on fixed buffers) is replayed as a CUDA graph, captured
  // again whenever a pointer, the image size or the capacity changes (ORB_B200_GRAPH=0 launches kernel by kernel)
  struct GraphKey { const void* p[8]; int w, h, cap, frames, flags; float mbf, mb; };   // mbf / mb are baked into k_stereo's launch
  u8* d_stereoScratch[2] = {nullptr, nullptr}; size_t stereoScratchBytes[2] = {0, 0};   // per workspace lane: match lists + counters of k_stereo's split mode
  bool useGraph = true;
  cudaGraphExec_t callGraph[2] = {nullptr, nullptr};   // [0] one frame, [1] stereo pair
  GraphKey callKey[2] = {};
  int callLaunches[2] = {0, 0};
  int lastLaunches = 0;
  int lastChunkFrames = 0;
  float* d_invScale = nullptr;          // mvInvScaleFactor on the device (stereo refinement)
  float* d_uRight[2] = {nullptr, nullptr}; float* d_depth[2] = {nullptr, nullptr};  // host-path staging
  int stereoOutCap = 0;                 // floats allocated in d_uRight[0] / d_depth[0]
  // the last orb_extract call's results are still on the device (pyramid in workspace frame 0, keypoints / descriptors /
  // count in the first staging set): what orb_stereo_match pairs up (0 = nothing valid)
  int singleCap = 0, singleW = 0, singleH = 0;
  // host wall-clock breakdown of the last orb_extract call (microseconds): staging copy of the image into pinned memory,
  // enqueue (H2D + graph launch + D2H requests), wait for the device, copy-out of the results
  double callUs[4] = {0, 0, 0, 0};
  // last work enqueued on a caller's stream (device entry points): recorded here so that the handle can wait for it
  // without a device-wide synchronisation (which would also wait for - and fail on - other handles' capturing streams)
  cudaEvent_t evUser = nullptr;
  bool userPending = false;
  // optional per-stage CUDA-event timing (bench roofline): 6 boundary events per chunk
  bool profile = false;
  std::vector<cudaEvent_t> evPool;
  size_t evUsed = 0;
  double stageMs[5] = {0, 0, 0, 0, 0};
  long long stageLaunches[5] = {0, 0, 0, 0, 0};
};

namespace {

// Waits for everything THIS handle has in flight: its own streams and the last work it enqueued on a caller's stream.
// Never cudaDeviceSynchronize: another handle may be capturing a CUDA graph on another host thread (src/Frame.cc:146-154
// runs two extractors from two threads), and a device-wide wait on a capturing stream is an error.
int sync_handle(orb_extractor* e) {
  cudaStream_t ss[] = {e->stream, e->laneStream[0], e->laneStream[1], e->blurStream[0], e->blurStream[1], e->sIn, e->sOut};
  for (cudaStream_t s : ss)
    if (s) ORB_CUDA(cudaStreamSynchronize(s));
  if (e->userPending && e->evUser) {
    ORB_CUDA(cudaEventSynchronize(e->evUser));
    e->userPending = false;
  }
  return ORB_OK;
}

// Remembers work enqueued on a caller's stream (see sync_handle).
int mark_user_stream(orb_extractor* e, cudaStream_t s) {
  if (s == e->stream) return ORB_OK;
  if (!e->evUser) ORB_CUDA(cudaEventCreateWithFlags(&e->evUser, cudaEventDisableTiming));
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return ORB_OK; }
  ORB_CUDA(cudaEventRecord(e->evUser, s));
  e->userPending = true;
  return ORB_OK;
}

void build_tables(orb_extractor* e) {
  // ORBextractor.cc:469-526. scaleFactor is a double member initialised from the float argument.
  const int nl = e->p.nlevels;
  const double sf = (double)e->p.scale_factor;
  e->scale.assign(nl, 1.0f);
  e->sigma2.assign(nl, 1.0f);
  for (int i = 1; i < nl; i++) {
    e->scale[i] = (float)((double)e->scale[i - 1] * sf);
    e->sigma2[i] = e->scale[i] * e->scale[i];
  }
  e->invScale.resize(nl);
  e->invSigma2.resize(nl);
  for (int i = 0; i < nl; i++) {
    e->invScale[i] = 1.0f / e->scale[i];
    e->invSigma2[i] = 1.0f / e->sigma2[i];
  }
  e->perLevel.assign(nl, 0);
  const float factor = (float)(1.0 / sf);
  float want = (float)e->p.nfeatures * (1.0f - factor) / (1.0f - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; l++) {
    e->perLevel[l] = cv_round_f(want);
    sum += e->perLevel[l];
    want *= factor;
  }
  e->perLevel[nl - 1] = std::max(e->p.nfeatures - sum, 0);
}

// resize taps of cv::resize INTER_LINEAR 8-bit (OpenCV 4.x): see DESIGN.md / SURVEY A.2
void make_taps(int dsize, int ssize, int2* out) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  for (int d = 0; d < dsize; d++) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)std::floor(fx);
    fx -= (float)s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; fx = 0.f; }
    const int c0 = cv_round_f((1.f - fx) * 2048.f), c1 = cv_round_f(fx * 2048.f);
    out[d].x = s;
    out[d].y = (c0 & 0xffff) | (c1 << 16);
  }
}

int build_geom(orb_extractor* e, int W, int H) {
  Geom& g = e->g;
  memset(&g, 0, sizeof g);
  const int nl = e->p.nlevels;
  g.nlevels = nl; g.W = W; g.H = H; g.iniTh = e->p.ini_th_fast; g.minTh = e->p.min_th_fast;
  long long pyrOff = 0, blurOff = 0;
  int candOff = 0, keptOff = 0, cellBase = 0, blurBase = 0, tapOff = 0;
  int nodeCap = 0, maxCw = 0, maxCh = 0, maxKp = 0;
  e->taps.clear();
  for (int l = 0; l < nl; l++) {
    LevelGeom& L = g.lv[l];
    L.w = cv_round_f((float)W * e->invScale[l]);   // :1663
    L.h = cv_round_f((float)H * e->invScale[l]);
    if (L.w > 32000 || L.h > 32000) ORB_FAIL(ORB_ERR_UNSUPPORTED, "image larger than 32000 px per side");
    const int Wp = L.w - 32, Hp = L.h - 32;        // detection window size (:1052-1060)
    if (Wp < 30 || Hp < 30) ORB_FAIL(ORB_ERR_UNSUPPORTED, "a pyramid level is smaller than one 30-px FAST cell (the reference divides by zero)");
    L.pitch = round_up(kLeftPad + L.w + kEdge, 32);
    L.off = pyrOff + (long long)kEdge * L.pitch + kLeftPad;
    pyrOff += round_up((long long)L.pitch * (L.h + 2 * kEdge), 256LL);
    L.bpitch = round_up(L.w, 32);
    L.boff = blurOff;
    blurOff += round_up((long long)L.bpitch * L.h, 256LL);
    const float width = (float)Wp, height = (float)Hp;
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    L.wCell = (int)std::ceil(width / (float)nCols);
    L.hCell = (int)std::ceil(height / (float)nRows);
    L.nColsAll = nCols;
    L.nCols = 0; L.nRows = 0;
    for (int j = 0; j < nCols; j++) if (16 + j * L.wCell < (L.w - 16) - 6) L.nCols = j + 1;  // :1083
    for (int i = 0; i < nRows; i++) if (16 + i * L.hCell < (L.h - 16) - 3) L.nRows = i + 1;  // :1101
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    maxCw = std::max(maxCw, L.wCell); maxCh = std::max(maxCh, L.hCell);
    L.nfeat = e->perLevel[l];
    L.nIni = (int)std::round((float)Wp / (float)Hp);  // :695
    if (L.nIni < 1) ORB_FAIL(ORB_ERR_UNSUPPORTED, "portrait aspect ratio > 2:1 (the reference divides by zero)");
    L.hX = (float)Wp / (float)L.nIni;
    L.candOff = candOff;
    L.candCap = std::min(1 << 24, std::max(1024, (L.w * L.h) / 6));
    candOff += round_up(L.candCap, 4);
    L.keptOff = keptOff;
    L.keptCap = std::max(L.nfeat + 3, 4 * L.nIni) + 1;   // == level_kept_cap(), orb_max_keypoints_for_size
    keptOff += round_up(L.keptCap, 4);
    maxKp += L.keptCap;
    nodeCap = std::max(nodeCap, L.keptCap + 3);
    L.blurTileBase = blurBase;
    L.blurTilesX = (L.w + kBlurTW - 1) / kBlurTW;
    L.blurTilesXMagic = L.blurTilesX == 1 ? 0u : (unsigned)((0x100000000ull + L.blurTilesX - 1) / L.blurTilesX);
    L.blurTilesY = (L.h + kBlurTH - 1) / kBlurTH;
    blurBase += L.blurTilesX * L.blurTilesY;
    L.scale = e->scale[l];
    L.patch = (float)(int)(31.f * e->scale[l]);  // :1164
    if (l > 0) {
      L.tapX = tapOff; tapOff += round_up(L.w, 2) + 8;  // even offsets: int4 loads of two taps; slack for 8-tap reads
      L.tapY = tapOff; tapOff += round_up(L.h, 2);
      e->taps.resize(tapOff);
      make_taps(L.w, g.lv[l - 1].w, e->taps.data() + L.tapX);
      make_taps(L.h, g.lv[l - 1].h, e->taps.data() + L.tapY);
    }
  }
  g.totalCells = cellBase;
  g.totalBlurTiles = blurBase;
  e->pyrStride = (size_t)pyrOff;
  e->blurStride = (size_t)blurOff;
  e->candTotal = candOff;
  e->keptTotal = keptOff;
  e->nodeCap = round_up(nodeCap, 2);
  e->maxKp = maxKp;
  {
    if (maxCw > 60 || maxCh > 60) ORB_FAIL(ORB_ERR_UNSUPPORTED, "FAST cell larger than 60 px");
    int maxW = kFastWarps;
    if (const char* ev = getenv("ORB_B200_FAST_WARPS")) maxW = std::max(1, std::min(kFastWarps, atoi(ev)));
    if (const char* ev = getenv("ORB_B200_FAST_TAIL_RUN")) e->fastTailRun = std::max(1, std::min(kFastRun, atoi(ev)));
    if (const char* ev = getenv("ORB_B200_FAST_TAIL_MUL")) e->fastTailMul = std::max(0, std::min(64, atoi(ev)));
    if (const char* ev = getenv("ORB_B200_FAST_FLAGS")) e->fastFlags = atoi(ev);
    // one group by default: measured on KITTI, giving the 11 % of cells that are 38-40 rows tall their own launch lets the
    // rest run at 22 instead of 20 resident warps per SM, but the second launch's ragged end costs more (4.56 vs 4.45 ms)
    bool allowSplit = false;
    if (const char* ev = getenv("ORB_B200_FAST_SPLIT")) allowSplit = atoi(ev) != 0;
    // per-warp carve-up for cells up to cw x ch, and the warps per CTA that keep most warps resident under 227 KB
    auto layout = [&](int cw, int ch, FastSmemLayout& y, int& warps, int& resident) {
      const int Rm = (ch + 1) / 2;                               // pair distance of the tallest cell
      y.rawPitchWords = 4 * ((15 + cw + 6 + 15) / 16);           // 16-byte chunks
      y.rawBytes = round_up(y.rawPitchWords * 4 * (ch + 6), 16);
      y.tileBytes = round_up(4 * ((3 + cw + 6 + 3) & ~3) * (Rm + 6), 16);
      y.hitsBytes = round_up(2 * (2 * Rm * cw), 16);             // hits + queue share it (see k_fast_cells)
      y.scBytes = round_up((cw + 2) * (ch + 2) + 4, 16);
      y.total = y.rawBytes + y.tileBytes + y.hitsBytes + y.scBytes;
      warps = 1; resident = 0;
      for (int w = 1; w <= maxW; w++) {
        const long long perCta = (long long)y.total * w + 1024;
        const int r = (int)std::min<long long>(32, (227 * 1024) / perCta) * w;
        if (perCta <= 200 * 1024 && r > resident) { resident = r; warps = w; }
      }
    };
    // levels by the size of their cells (what one warp needs), smallest first; every prefix is a candidate first group
    std::vector<int> order(nl);
    for (int l = 0; l < nl; l++) order[l] = l;
    auto need = [&](int l) { FastSmemLayout y; int w, r; layout(g.lv[l].wCell, g.lv[l].hCell, y, w, r); return y.total; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return need(a) < need(b); });
    auto cellsOf = [&](int l) { return g.lv[l].nCols * g.lv[l].nRows; };
    double bestCost = 1e300;
    int bestSplit = nl;                                           // levels order[0 .. split) form the first group
    for (int split = 1; split <= nl; split++) {
      if (!allowSplit && split != nl) continue;
      double cost = 0;
      bool ok = true;
      for (int gi = 0; gi < 2; gi++) {
        const int a = gi == 0 ? 0 : split, b = gi == 0 ? split : nl;
        if (a >= b) continue;
        int cw = 0, ch = 0, cells = 0;
        for (int k = a; k < b; k++) { cw = std::max(cw, g.lv[order[k]].wCell); ch = std::max(ch, g.lv[order[k]].hCell); cells += cellsOf(order[k]); }
        FastSmemLayout y; int w, r;
        layout(cw, ch, y, w, r);
        if (r == 0) { ok = false; break; }
        // a launch's time ~ cells / resident warps (measured: 18 -> 20 resident warps = -6 %), plus a tail per launch
        cost += (double)cells / r + 0.02 * g.totalCells / 20.0;
      }
      if (ok && cost < bestCost) { bestCost = cost; bestSplit = split; }
    }
    if (bestCost >= 1e300) ORB_FAIL(ORB_ERR_UNSUPPORTED, "FAST cell too large");
    e->fastGroups = bestSplit == nl ? 1 : 2;
    for (int gi = 0; gi < e->fastGroups; gi++) {
      const int a = gi == 0 ? 0 : bestSplit, b = gi == 0 ? bestSplit : nl;
      FastGroup& G = e->fastGroup[gi];
      memset(&G, 0, sizeof G);
      std::vector<int> lv(order.begin() + a, order.begin() + b);
      std::sort(lv.begin(), lv.end());                            // inside a group: ascending level (large levels first)
      int cw = 0, ch = 0, base = 0;
      for (int l : lv) {
        G.posLevel[G.nPos] = l;
        G.posBase[G.nPos++] = base;
        base += cellsOf(l);
        cw = std::max(cw, g.lv[l].wCell); ch = std::max(ch, g.lv[l].hCell);
      }
      G.posBase[G.nPos] = base;
      int resident = 0;
      layout(cw, ch, e->fastLayG[gi], e->fastWarpsG[gi], resident);
      e->fastSmemG[gi] = (size_t)e->fastLayG[gi].total * e->fastWarpsG[gi];
      if (e->fastSmemG[gi] > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "FAST cell too large");
    }
  }
  {
    // k_resize_strip: the 4 columns of a thread must read one 8-byte source window (scale factor < ~1.7)
    bool ok = true;
    if (const char* ev = getenv("ORB_B200_PYR")) ok = atoi(ev) != 0;
    for (int l = 1; l < nl && ok; l++) {
      const LevelGeom& D = g.lv[l];
      const int2* tX = e->taps.data() + D.tapX;
      auto refl = [](int p, int n) { p = p < 0 ? -p : p; return p >= n ? 2 * (n - 1) - p : p; };
      for (int c = -20; c <= D.w + kEdge - 1; c += 4) {          // every aligned word of the bordered width
        int lo = 1 << 30, hi = -1;
        for (int j = 0; j < 4; j++) {
          const int sx = tX[refl(std::min(std::max(c + j, -kEdge), D.w + kEdge - 1), D.w)].x;
          lo = std::min(lo, sx); hi = std::max(hi, sx);
        }
        if (hi + 1 - lo > 7) ok = false;
      }
      // ... and output rows must start at strictly increasing source rows, at most 63 rows apart within a band
      const int2* tY = e->taps.data() + D.tapY;
      for (int j = 1; j < D.h; j++)
        if (tY[j].x <= tY[j - 1].x) ok = false;
      for (int j = 0; j < D.h; j += kRzBand)
        if (tY[std::min(j + kRzBand, D.h) - 1].x - tY[j].x > 63) ok = false;
    }
    // k_level0_border2 / k_fill_borders: at most 8 edge groups per bordered row
    const int groups0 = g.lv[0].pitch >> 4, gB0 = std::max(3, (g.lv[0].w - 20) / 16 + 3);
    if (3 + (groups0 - gB0) > 8) ok = false;
    int blocks = 0;
    for (int l = 1; l < nl; l++) {
      e->borderJobs.base[l] = blocks;
      blocks += (2 * kEdge + 7) / 8;
    }
    e->borderJobs.base[0] = 0;
    e->borderJobs.base[nl] = blocks;
    e->borderBlocks = blocks;
    // item table of k_pyramid_fused: producers before consumers
    {
      auto refl = [](int p, int n) { p = p < 0 ? -p : p; return p >= n ? 2 * (n - 1) - p : p; };
      e->pyrItems.clear();
      PyrPlan& P = e->pyrPlan;
      int base = 0;
      P.bandBase[0] = 0;
      const int nb0 = (g.lv[0].h + 2 * kEdge + 7) / 8;
      for (int b = 0; b < nb0; b++) e->pyrItems.push_back(PyrItem{0, 0, b, 0, 1, 0});
      base += nb0;
      for (int l = 1; l < nl; l++) {
        P.bandBase[l] = base;
        const LevelGeom& D = g.lv[l];
        const LevelGeom& S = g.lv[l - 1];
        const int2* tY = e->taps.data() + D.tapY;
        const int nb = (D.h + kPyrFusedBand - 1) / kPyrFusedBand;
        for (int b = 0; b < nb; b++) {
          const int j0 = b * kPyrFusedBand, j1 = std::min(j0 + kPyrFusedBand, D.h);
          const int sy0 = tY[j0].x, sy1 = std::min(tY[j1 - 1].x + 1, S.h - 1);
          PyrItem it{1, (short)l, b, l - 1, 0, 0};
          if (l == 1) { it.depFirst = (sy0 + kEdge) / 8; it.depLast = (sy1 + kEdge) / 8; }
          else { it.depFirst = sy0 / kPyrFusedBand; it.depLast = sy1 / kPyrFusedBand; }
          e->pyrItems.push_back(it);
        }
        base += nb;
      }
      P.bandBase[nl] = base;
      for (int l = 1; l < nl; l++) {
        const LevelGeom& D = g.lv[l];
        for (int b = 0; b * 8 < 2 * kEdge; b++) {
          int lo = 1 << 30, hi = -1;
          for (int k = b * 8; k < std::min(b * 8 + 8, 2 * kEdge); k++) {
            const int by = k < kEdge ? k : D.h + k;
            const int r = refl(by - kEdge, D.h);
            lo = std::min(lo, r); hi = std::max(hi, r);
          }
          e->pyrItems.push_back(PyrItem{2, (short)l, b, l, lo / kPyrFusedBand, hi / kPyrFusedBand});
        }
      }
      P.nItems = (int)e->pyrItems.size();
      if (const char* ev = getenv("ORB_B200_PYR_FUSED")) e->pyrFused = atoi(ev) != 0;
    }
    e->pyrTiled = ok;
  }
  // sequence-number bitmap of k_quadtree: every refinement round hands out 4 numbers per split node; rounds are bounded
  // by the halvings of the longer side (larger counts fall back to the all-pairs ranking inside the kernel)
  {
    int rounds = 2;
    for (int side = std::max(g.lv[0].w, g.lv[0].h); side > 1; side >>= 1) rounds++;
    e->qtSeqWords = std::min(2048, (g.lv[0].nIni + 4 * rounds * e->nodeCap + 31) / 32 + 1);
  }
  e->qtSmem = (size_t)e->nodeCap * (2 * sizeof(QtNode) + 4 * 4 * 2 + 4 * 4 + 8) + (size_t)kQtSmemKeys * 6 + (size_t)e->qtSeqWords * 8;
  if (e->qtSmem > 220 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "features per level too large for the quadtree kernel's shared memory");
  return ORB_OK;
}

struct Lane {   // the workspace slice of one lane
  u8* pyr; u8* blur; uint2* cand; int* candCount; unsigned short* keyNode; uint2* kept; int* keptCount; int* work;
};

Lane lane_of(const orb_extractor* e, int lane) {
  const size_t F = (size_t)e->wsFrames * lane;
  Lane L;
  L.pyr = e->d_pyr + F * e->pyrStride;
  L.blur = e->d_blur + F * e->blurStride;
  L.cand = e->d_cand + F * e->candTotal;
  L.candCount = e->d_candCount + F * kMaxLevels;
  L.keyNode = e->d_keyNode + F * e->candTotal;
  L.kept = e->d_kept + F * e->keptTotal;
  L.keptCount = e->d_keptCount + F * kMaxLevels;
  L.work = e->d_work + 2 * lane;
  return L;
}

void free_workspace(orb_extractor* e) {
  cudaFree(e->d_pyr); cudaFree(e->d_blur); cudaFree(e->d_cand); cudaFree(e->d_candCount);
  cudaFree(e->d_keyNode); cudaFree(e->d_kept); cudaFree(e->d_keptCount); cudaFree(e->d_taps);
  e->d_pyr = e->d_blur = nullptr; e->d_cand = nullptr; e->d_candCount = nullptr; e->d_keyNode = nullptr;
  e->d_kept = nullptr; e->d_keptCount = nullptr; e->d_taps = nullptr;
  e->wsFrames = 0;
}

// TMA tensor maps of the workspace (k_describe_tma): per level, the bordered pyramid plane and the blurred plane as
// (x, y, frame) byte tensors. cuTensorMapEncodeTiled is taken from the driver through the runtime, so the library
// still links against cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int build_desc_maps(orb_extractor* e) {
  e->descTma = false;
  e->blurTma = false;
  if (const char* ev = getenv("ORB_B200_DESC_TMA"))
    if (atoi(ev) == 0) return ORB_OK;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return ORB_OK;   // k_describe (cp.async staging) is used instead
  }
  const EncodeTiledFn encode = (EncodeTiledFn)sym;
  const Geom& g = e->g;
  for (int lane = 0; lane < e->lanes; lane++) {
    const Lane W = lane_of(e, lane);
    for (int l = 0; l < g.nlevels; l++) {
      const LevelGeom& L = g.lv[l];
      const cuuint32_t ones[3] = {1, 1, 1};
      {
        u8* base = W.pyr + (L.off - (long long)kEdge * L.pitch - kLeftPad);
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kEdge), (cuuint64_t)e->wsFrames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)e->pyrStride};
        const cuuint32_t box[3] = {kTmaMomW, 31, 1};
        if (encode(&e->descMaps[lane].pyr[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ORB_OK;
      }
      {
        u8* base = W.blur + L.boff;
        const cuuint64_t dims[3] = {(cuuint64_t)L.bpitch, (cuuint64_t)L.h, (cuuint64_t)e->wsFrames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.bpitch, (cuuint64_t)e->blurStride};
        const cuuint32_t box[3] = {kTmaPatchW, 37, 1};
        if (encode(&e->descMaps[lane].blur[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ORB_OK;
        const cuuint32_t box48[3] = {48, 37, 1};
        if (encode(&e->descMaps[lane].blur48[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box48, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ORB_OK;
      }
    }
  }
  e->descRing = 8;   // keypoints per warp of k_describe_ring; 0 = k_describe_tma
  if (const char* ev = getenv("ORB_B200_DESC_RING")) e->descRing = atoi(ev);
  if (e->descRing != 0 && e->descRing != 4 && e->descRing != 8 && e->descRing != 16) e->descRing = 8;
  ORB_CUDA(raise_dynamic_smem(describe_ring_variant(e->descRing ? e->descRing : 8), desc_ring_smem(kDescRingDepth)));
  ORB_CUDA(raise_dynamic_smem(k_describe_tma, kDescTmaSmem));
  e->descTma = true;
  // blur input tiles: (kBlurTW + 32) x (kBlurTH + 6) bytes of the bordered plane
  bool blurOk = true;
  if (const char* ev = getenv("ORB_B200_BLUR_TMA"))
    if (atoi(ev) == 0) blurOk = false;
  for (int lane = 0; blurOk && lane < e->lanes; lane++) {
    const Lane W = lane_of(e, lane);
    for (int l = 0; blurOk && l < g.nlevels; l++) {
      const LevelGeom& L = g.lv[l];
      const cuuint32_t ones[3] = {1, 1, 1};
      u8* base = W.pyr + (L.off - (long long)kEdge * L.pitch - kLeftPad);
      const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kEdge), (cuuint64_t)e->wsFrames};
      const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)e->pyrStride};
      const cuuint32_t box[3] = {kBlurChunks * 16, kBlurRows, 1};
      if (encode(&e->blurMaps[lane].m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, ones,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        blurOk = false;
    }
  }
  e->blurTma = blurOk;
  return ORB_OK;
}

int ensure_geom(orb_extractor* e, int W, int H, int frames) {
  ORB_CUDA(cudaSetDevice(e->device));
  const bool newGeom = !e->haveGeom || e->g.W != W || e->g.H != H;
  if (newGeom) {
    e->haveGeom = false;
    int st = build_geom(e, W, H);
    if (st) return st;
    { const int ss_ = sync_handle(e); if (ss_) return ss_; }
    free_workspace(e);
    ORB_CUDA(cudaMalloc(&e->d_taps, std::max<size_t>(1, e->taps.size()) * sizeof(int2)));
    ORB_CUDA(cudaMemcpy(e->d_taps, e->taps.data(), e->taps.size() * sizeof(int2), cudaMemcpyHostToDevice));
    cudaFree(e->d_pyrItems); cudaFree(e->d_pyrFlags);
    e->d_pyrItems = nullptr; e->d_pyrFlags = nullptr;
    ORB_CUDA(cudaMalloc(&e->d_pyrItems, std::max<size_t>(1, e->pyrItems.size()) * sizeof(PyrItem)));
    ORB_CUDA(cudaMemcpy(e->d_pyrItems, e->pyrItems.data(), e->pyrItems.size() * sizeof(PyrItem), cudaMemcpyHostToDevice));
    ORB_CUDA(cudaMalloc(&e->d_pyrFlags, ((size_t)kPyrFusedMaxFrames * e->pyrPlan.bandBase[e->g.nlevels] + 1) * sizeof(int)));
    ORB_CUDA(raise_dynamic_smem(k_quadtree, e->qtSmem));
    ORB_CUDA(raise_dynamic_smem(k_fast_cells, std::max(e->fastSmemG[0], e->fastSmemG[1])));
    ORB_CUDA(raise_dynamic_smem(k_describe, kDescSmem));
    {
      int perSM = 0, dev = 0, sms = 0;
      ORB_CUDA(cudaGetDevice(&dev));
      ORB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      for (int gi = 0; gi < e->fastGroups; gi++) {
        ORB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_fast_cells, 32 * e->fastWarpsG[gi], e->fastSmemG[gi]));
        e->fastBlocksG[gi] = std::max(1, perSM) * std::max(1, sms);
      }
      e->numSMs = std::max(1, sms);
    }
    e->haveGeom = true;
  }
  frames = std::min(frames, e->maxBatch);
  if (frames > e->wsFrames) {
    { const int ss_ = sync_handle(e); if (ss_) return ss_; }
    int2* keepTaps = e->d_taps; e->d_taps = nullptr;
    free_workspace(e);
    e->d_taps = keepTaps;
    const size_t F = (size_t)frames * e->lanes;
    ORB_CUDA(cudaMalloc(&e->d_pyr, F * e->pyrStride));
    ORB_CUDA(cudaMalloc(&e->d_blur, F * e->blurStride));
    ORB_CUDA(cudaMalloc(&e->d_cand, F * e->candTotal * sizeof(uint2)));
    ORB_CUDA(cudaMalloc(&e->d_keyNode, F * e->candTotal * sizeof(unsigned short)));
    ORB_CUDA(cudaMalloc(&e->d_candCount, F * kMaxLevels * sizeof(int)));
    ORB_CUDA(cudaMalloc(&e->d_kept, F * e->keptTotal * sizeof(uint2)));
    ORB_CUDA(cudaMalloc(&e->d_keptCount, F * kMaxLevels * sizeof(int)));
    e->wsFrames = frames;
    const int st = build_desc_maps(e);
    if (st) return st;
  }
  return ORB_OK;
}

int stage_mark(orb_extractor* e, cudaStream_t s) {
  if (!e->profile) return ORB_OK;
  if (e->evUsed == e->evPool.size()) {
    cudaEvent_t ev;
    ORB_CUDA(cudaEventCreate(&ev));
    e->evPool.push_back(ev);
  }
  ORB_CUDA(cudaEventRecord(e->evPool[e->evUsed++], s));
  return ORB_OK;
}

// One chunk (<= wsFrames frames) through the whole pipeline, asynchronous on `s`.
int run_chunk(orb_extractor* e, const u8* d_img, int B, size_t step, size_t frameStride, orb_keypoint* d_kps,
              int cap, int* d_counts, u8* d_desc, cudaStream_t s, int lane = 0) {
  const Geom& g = e->g;
  const Lane W = lane_of(e, lane);
  e->lastLane = lane;
  const int nl = g.nlevels;
  int launches = 0, st;
  if ((st = stage_mark(e, s))) return st;
  const bool l0fork = !e->pyrTiled && !e->profile && e->l0Fork && nl > 1;
  int pyrLaunches = nl;
  if (e->pyrTiled && e->pyrFused && B <= kPyrFusedMaxFrames && lane == 0) {
    // few frames: the whole pyramid in one dependency-driven launch
    const size_t flagInts = (size_t)B * e->pyrPlan.bandBase[nl];
    ORB_CUDA(cudaMemsetAsync(e->d_pyrFlags, 0, ((size_t)kPyrFusedMaxFrames * e->pyrPlan.bandBase[nl] + 1) * sizeof(int), s));
    (void)flagInts;
    const int blocks = std::min(e->pyrPlan.nItems * B, 2 * e->numSMs);
    k_pyramid_fused<<<blocks, 256, 0, s>>>(g, d_img, step, frameStride, W.pyr, e->pyrStride, e->d_taps, e->d_pyrItems, e->pyrPlan, B,
                                           e->d_pyrFlags, e->d_pyrFlags + (size_t)kPyrFusedMaxFrames * e->pyrPlan.bandBase[nl]);
    launches++;
    pyrLaunches = 1;
  } else if (e->pyrTiled) {
    pyrLaunches = nl > 1 ? nl + 1 : nl;
    const LevelGeom& L0 = g.lv[0];
    const int rows0 = L0.h + 2 * kEdge;
    const int nbInt = (rows0 + 7) / 8, nbEdge = (rows0 + 31) / 32;
    k_level0_border2<<<dim3(nbInt + nbEdge, B), 256, 0, s>>>(g, d_img, step, frameStride, W.pyr, e->pyrStride, nbInt);
    launches++;
    // few frames (the per-frame drop-in call): short bands so that a level still fills the machine
    for (int l = 1; l < nl; l++) {
      const LevelGeom& L = g.lv[l];
      // the upper levels have few strips: shorter bands there keep more than one wave of CTAs in flight
      const int bandRows = B >= 8 ? (l >= e->rzSmallFrom ? e->rzSmallBand : kRzBand) : 4;
      dim3 grid(((L.w + kEdge - 1 + 20) / 4 + 1 + kRzThreads - 1) / kRzThreads, (L.h + bandRows - 1) / bandRows, B);
      if (B < 8 && e->pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(kRzThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ORB_CUDA(cudaLaunchKernelEx(&cfg, k_resize_strip_pdl, g, l, W.pyr, e->pyrStride, (const int2*)e->d_taps, bandRows));
      } else {
        k_resize_strip<<<grid, kRzThreads, 0, s>>>(g, l, W.pyr, e->pyrStride, e->d_taps, bandRows);
      }
      launches++;
    }
    if (nl > 1) {
      k_fill_borders<<<dim3(e->borderBlocks, B), 256, 0, s>>>(g, W.pyr, e->pyrStride, e->borderJobs);
      launches++;
    }
  } else {
  {
    const LevelGeom& L = g.lv[0];
    dim3 grid((L.pitch / 16 + 31) / 32, (L.h + 2 * kEdge + 7) / 8, B);
    cudaStream_t ls = l0fork ? e->blurStream[lane] : s;
    if (l0fork) {
      ORB_CUDA(cudaEventRecord(e->evL0Go[lane], s));
      ORB_CUDA(cudaStreamWaitEvent(ls, e->evL0Go[lane], 0));
    }
    k_level0_border<<<grid, dim3(32, 8), 0, ls>>>(g, d_img, step, frameStride, W.pyr, e->pyrStride);
    if (l0fork) ORB_CUDA(cudaEventRecord(e->evL0Done[lane], ls));
    launches++;
  }
  for (int l = 1; l < nl; l++) {
    const LevelGeom& L = g.lv[l];
    dim3 grid((L.pitch / 4 + 31) / 32, (L.h + 2 * kEdge + 31) / 32, B);
    const bool fromImage = l0fork && l == 1;
    k_resize_border<<<grid, dim3(32, 8), 0, s>>>(g, l, W.pyr, e->pyrStride, e->d_taps, fromImage ? d_img : nullptr,
                                                 fromImage ? frameStride : 0, fromImage ? (int)step : 0);
    launches++;
  }
  if (l0fork) ORB_CUDA(cudaStreamWaitEvent(s, e->evL0Done[lane], 0));
  }
  const int fork = e->profile ? 0 : e->blurFork;
  cudaStream_t bs = fork ? e->blurStream[lane] : s;
  if (fork == 1) {
    ORB_CUDA(cudaEventRecord(e->evBlurGo[lane], s));
    ORB_CUDA(cudaStreamWaitEvent(bs, e->evBlurGo[lane], 0));
  }
  ORB_CUDA(cudaMemsetAsync(W.candCount, 0, (size_t)B * nl * sizeof(int), s));
  ORB_CUDA(cudaMemsetAsync(W.work, 0, 2 * sizeof(int), s));
  if ((st = stage_mark(e, s))) return st;
  for (int gi = 0; gi < e->fastGroups; gi++) {
    const FastGroup& G = e->fastGroup[gi];
    const int nItems = G.posBase[G.nPos] * B;
    if (nItems == 0) continue;
    const int blocks = std::min(e->fastBlocksG[gi], (nItems + e->fastWarpsG[gi] - 1) / e->fastWarpsG[gi]);
    // cells per grab: long runs amortise the atomic and keep a warp on neighbouring cells, but every warp should get
    // at least ~6 grabs or the launch ends ragged (the small group of tall cells has ~12 cells per warp)
    const int run = std::max(1, std::min(kFastRun, nItems / (blocks * e->fastWarpsG[gi] * 6)));
    k_fast_cells<<<blocks, 32 * e->fastWarpsG[gi], e->fastSmemG[gi], s>>>(g, W.pyr, e->pyrStride, W.cand, W.candCount, e->candTotal, nItems,
                                                                    e->fastLayG[gi], G, W.work + gi, run, std::min(run, e->fastTailRun),
                                                                    e->fastTailMul, e->fastFlags);
    launches++;
  }
  if ((st = stage_mark(e, s))) return st;
  if (fork == 2) {
    ORB_CUDA(cudaEventRecord(e->evBlurGo[lane], s));
    ORB_CUDA(cudaStreamWaitEvent(bs, e->evBlurGo[lane], 0));
  }
  k_quadtree<<<dim3(B, nl), B <= 4 ? kQtMaxThreads : kQtThreads, e->qtSmem, s>>>(g, W.cand, W.candCount, W.keyNode, W.kept,
                                                       W.keptCount, e->candTotal, e->keptTotal, e->nodeCap, e->qtSeqWords,
                                                       e->d_overflow);
  launches++;
  if ((st = stage_mark(e, s))) return st;
  if (e->blurTma)
    k_blur7<true><<<dim3(g.totalBlurTiles, B), 256, 0, bs>>>(g, W.pyr, e->pyrStride, W.blur, e->blurStride, e->blurMaps[lane]);
  else
    k_blur7<false><<<dim3(g.totalBlurTiles, B), 256, 0, bs>>>(g, W.pyr, e->pyrStride, W.blur, e->blurStride, e->blurMaps[lane]);
  if (fork) {
    ORB_CUDA(cudaEventRecord(e->evBlurDone[lane], bs));
    ORB_CUDA(cudaStreamWaitEvent(s, e->evBlurDone[lane], 0));
  }
  launches++;
  if ((st = stage_mark(e, s))) return st;
  const int slots = std::min(cap, e->maxKp);
  // a few frames per launch: the ring's serial chain of 8 keypoints per warp is the long pole (15 vs 11 us for one frame)
  if (e->descTma && e->descRing && B >= 8 && g.lv[0].w < 16384 && g.lv[0].h < 16384) {
    const int per = 8 * e->descRing;
    auto kfn = describe_ring_variant(e->descRing);
    kfn<<<dim3((slots + per - 1) / per, B), 256, desc_ring_smem(kDescRingDepth), s>>>(g, e->descMaps[lane], W.kept, W.keptCount, e->keptTotal, e->d_pattern, d_kps,
                                                                     d_desc, d_counts, cap, e->d_overflow);
  } else if (e->descTma)
    k_describe_tma<<<dim3((slots + kDescSlots * kDescRounds - 1) / (kDescSlots * kDescRounds), B), 256, kDescTmaSmem, s>>>(g, e->descMaps[lane], W.kept, W.keptCount,
                                                                                            e->keptTotal, e->d_pattern, d_kps, d_desc,
                                                                                            d_counts, cap, e->d_overflow);
  else
    k_describe<<<dim3((slots + kDescSlots - 1) / kDescSlots, B), 256, kDescSmem, s>>>(g, W.pyr, e->pyrStride, W.blur, e->blurStride,
                                                                W.kept, W.keptCount, e->keptTotal, e->d_pattern,
                                                                d_kps, d_desc, d_counts, cap, e->d_overflow);
  launches++;
  if ((st = stage_mark(e, s))) return st;
  ORB_CUDA(cudaGetLastError());
  if (e->profile) {
    e->stageLaunches[0] += pyrLaunches; e->stageLaunches[1] += e->fastGroups; e->stageLaunches[2]++; e->stageLaunches[3]++; e->stageLaunches[4]++;
  }
  e->lastLaunches += launches;
  e->lastChunkFrames = B;
  return ORB_OK;
}

// Frame::ComputeStereoMatches for the B/2 stereo pairs of the chunk that was just extracted
// (frames 2p = left, 2p+1 = right; the chunk's pyramids are still in the workspace).
// match lists (pairs x cap) + counters (2 per pair) of k_stereo's split mode, one block per workspace lane; grown outside of any
// graph capture and before the lanes of a multi-chunk call are forked
static size_t stereo_counter_bytes(int pairs) { return round_up((size_t)2 * pairs * sizeof(int), (size_t)256); }
// CTAs per pair: a pair alone in a call is a latency problem (16 CTAs), a chunk of 128 pairs at one CTA per pair is less than
// one wave of the machine (2 CTAs per pair), many more pairs fill it by themselves
// (128-pair chunks, ms per 1024 pairs incl. extraction: 1 CTA per pair 13.41, 2: 13.27, 3: 13.36, 4: 13.43, 8: 13.61 - every CTA of
// a pair repeats the pair's set-up)
static int stereo_split(int pairs) { return pairs <= 8 ? 16 : (pairs <= 32 ? 4 : (pairs <= 192 ? 2 : 1)); }
int ensure_stereo_scratch(orb_extractor* e, int cap, int pairs, int lane, cudaStream_t s) {
  const size_t need = (size_t)pairs * cap * sizeof(int2) + stereo_counter_bytes(pairs);
  if (need <= e->stereoScratchBytes[lane]) return ORB_OK;
  ORB_CUDA(cudaStreamSynchronize(s));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  cudaFree(e->d_stereoScratch[lane]);
  e->d_stereoScratch[lane] = nullptr; e->stereoScratchBytes[lane] = 0;
  if (e->callGraph[1]) { cudaGraphExecDestroy(e->callGraph[1]); e->callGraph[1] = nullptr; }
  ORB_CUDA(cudaMalloc(&e->d_stereoScratch[lane], need));
  e->stereoScratchBytes[lane] = need;
  return ORB_OK;
}

int run_stereo(orb_extractor* e, int B, const orb_keypoint* d_kps, int cap, const int* d_counts, const u8* d_desc, float mbf,
               float mb, float* d_uRight, float* d_depth, cudaStream_t s, int lane = 0) {
  // right keypoints (x, row band, octave), SAD list, row index (H+2 ints) and up to 12 index entries
  // per keypoint (more -> the kernel scans all right keypoints instead of using the index)
  size_t fixed = (size_t)cap * (4 + 4 + 8 + 1) + (size_t)(e->g.H + 2) * 4 + 64;
  int entCap = std::max(cap * 12, 1024);   // >= 1024 entries: the median histogram may live there
  if (fixed + (size_t)entCap * 2 > 100 * 1024) entCap = (int)std::max<long long>(0, (100 * 1024 - (long long)fixed) / 2);
  const size_t smem = round_up(fixed + (size_t)entCap * 2, (size_t)16);
  if (smem > 200 * 1024 || (size_t)e->g.H * 4 > (size_t)cap * 8)
    ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many keypoints / rows for the stereo kernel's shared memory");
  if (cap > 65535) ORB_FAIL(ORB_ERR_UNSUPPORTED, "stereo matching supports at most 65535 keypoints per frame");
  ORB_CUDA(raise_dynamic_smem(k_stereo, smem));
  // deal every pair to G CTAs (see stereo_split); without a scratch block of this lane that is large enough: one CTA per pair
  const int pairs = B / 2;
  int G = stereo_split(pairs);
  if (G > 1 && e->stereoScratchBytes[lane] < (size_t)pairs * cap * sizeof(int2) + stereo_counter_bytes(pairs)) G = 1;
  if (G > 1) ORB_CUDA(cudaMemsetAsync(e->d_stereoScratch[lane], 0, (size_t)2 * pairs * sizeof(int), s));
  int* gcnt = reinterpret_cast<int*>(e->d_stereoScratch[lane]);
  int2* gsad = G > 1 ? reinterpret_cast<int2*>(e->d_stereoScratch[lane] + stereo_counter_bytes(pairs)) : nullptr;
  k_stereo<<<dim3(pairs, G), kStereoThreads, smem, s>>>(e->g, lane_of(e, lane).pyr, e->pyrStride, d_kps, d_desc, d_counts, cap, mbf, mb,
                                                        e->d_invScale, d_uRight, d_depth, entCap, gsad, gcnt);
  ORB_CUDA(cudaGetLastError());
  e->lastLaunches++;
  return ORB_OK;
}

// Fork the caller's stream into the lane streams / join them back (used when a call spans
// several chunks; a single chunk runs directly on the caller's stream).
int lanes_fork(orb_extractor* e, cudaStream_t s, int nChunks) {
  if (e->lanes < 2 || nChunks < 2) return ORB_OK;
  ORB_CUDA(cudaEventRecord(e->evFork, s));
  for (int l = 0; l < 2; l++) ORB_CUDA(cudaStreamWaitEvent(e->laneStream[l], e->evFork, 0));
  return ORB_OK;
}
int lanes_join(orb_extractor* e, cudaStream_t s, int nChunks) {
  if (e->lanes < 2 || nChunks < 2) return ORB_OK;
  for (int l = 0; l < 2; l++) {
    ORB_CUDA(cudaEventRecord(e->evJoin[l], e->laneStream[l]));
    ORB_CUDA(cudaStreamWaitEvent(s, e->evJoin[l], 0));
  }
  return ORB_OK;
}

int ensure_stage(orb_extractor* e, size_t inBytes, int frames, int cap) {
  if (!e->sIn) {
    ORB_CUDA(cudaStreamCreateWithFlags(&e->sIn, cudaStreamNonBlocking));
    ORB_CUDA(cudaStreamCreateWithFlags(&e->sOut, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      ORB_CUDA(cudaEventCreateWithFlags(&e->evIn[b], cudaEventDisableTiming));
      ORB_CUDA(cudaEventCreateWithFlags(&e->evDone[b], cudaEventDisableTiming));
      ORB_CUDA(cudaEventCreateWithFlags(&e->evOut[b], cudaEventDisableTiming));
    }
  }
  if (inBytes > e->d_inBytes) {
    { const int ss_ = sync_handle(e); if (ss_) return ss_; }
    for (int b = 0; b < 2; b++) {
      cudaFree(e->d_in[b]);
      e->d_in[b] = nullptr;
      ORB_CUDA(cudaMalloc(&e->d_in[b], inBytes));
    }
    e->d_inBytes = inBytes;
  }
  if ((size_t)frames * cap > (size_t)e->stageFrames * e->stageCap || frames > e->stageFrames) {
    { const int ss_ = sync_handle(e); if (ss_) return ss_; }
    for (int b = 0; b < 2; b++) {
      cudaFree(e->d_kps[b]); cudaFree(e->d_desc[b]); cudaFree(e->d_n[b]);
      e->d_kps[b] = nullptr; e->d_desc[b] = nullptr; e->d_n[b] = nullptr;
      ORB_CUDA(cudaMalloc(&e->d_kps[b], (size_t)frames * cap * sizeof(orb_keypoint)));
      ORB_CUDA(cudaMalloc(&e->d_desc[b], (size_t)frames * cap * 32));
      ORB_CUDA(cudaMalloc(&e->d_n[b], (size_t)frames * sizeof(int)));
    }
    e->stageFrames = frames;
    e->stageCap = cap;
  }
  return ORB_OK;
}

// run_chunk (+ run_stereo for a pair) of a single-call entry point through a captured graph
int run_call(orb_extractor* e, int frames, int width, int height, size_t dFrame, int capacity, bool stereo, float mbf, float mb,
             cudaStream_t s) {
  auto direct = [&]() -> int {
    int st = run_chunk(e, e->d_in[0], frames, width, dFrame, e->d_kps[0], capacity, e->d_n[0], e->d_desc[0], s);
    if (st) return st;
    if (stereo) st = run_stereo(e, frames, e->d_kps[0], capacity, e->d_n[0], e->d_desc[0], mbf, mb, e->d_uRight[0], e->d_depth[0], s, 0);
    return st;
  };
  if (!e->useGraph || e->profile || e->blurFork || e->l0Fork) return direct();
  const int slot = stereo ? 1 : 0;
  const Lane W = lane_of(e, 0);
  orb_extractor::GraphKey key = {};
  key.p[0] = W.pyr; key.p[1] = W.blur; key.p[2] = W.cand; key.p[3] = W.kept; key.p[4] = e->d_in[0]; key.p[5] = e->d_kps[0];
  key.p[6] = e->d_desc[0]; key.p[7] = stereo ? (const void*)e->d_uRight[0] : (const void*)e->d_n[0];
  key.w = width; key.h = height; key.cap = capacity; key.frames = frames;
  key.flags = (e->descTma ? 1 : 0) | (e->blurTma ? 2 : 0) | (stereo ? 4 : 0);
  key.mbf = stereo ? mbf : 0.f; key.mb = stereo ? mb : 0.f;   // compared exactly (bitwise, by the memcmp below)
  if (!e->callGraph[slot] || memcmp(&key, &e->callKey[slot], sizeof key) != 0) {
    if (e->callGraph[slot]) { cudaGraphExecDestroy(e->callGraph[slot]); e->callGraph[slot] = nullptr; }
    const int before = e->lastLaunches;
    cudaGraph_t graph = nullptr;
    ORB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int st = direct();
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (st || ce != cudaSuccess || !graph) {   // could not be captured: run it the plain way from now on
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      e->useGraph = false;
      e->lastLaunches = before;
      return direct();
    }
    const cudaError_t ie = cudaGraphInstantiate(&e->callGraph[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      e->callGraph[slot] = nullptr;
      e->useGraph = false;
      e->lastLaunches = before;
      return direct();
    }
    e->callKey[slot] = key;
    e->callLaunches[slot] = e->lastLaunches - before;
    e->lastLaunches = before;
  }
  ORB_CUDA(cudaGraphLaunch(e->callGraph[slot], s));
  e->lastLaunches += e->callLaunches[slot];
  e->lastChunkFrames = frames;
  e->lastLane = 0;
  return ORB_OK;
}

int ensure_pinned(orb_extractor* e, size_t inBytes, size_t outBytes) {
  if (inBytes > e->h_inBytes) {
    if (e->h_in) cudaFreeHost(e->h_in);
    e->h_in = nullptr; e->h_inBytes = 0;
    ORB_CUDA(cudaHostAlloc((void**)&e->h_in, inBytes, cudaHostAllocDefault));
    e->h_inBytes = inBytes;
  }
  if (outBytes > e->h_outBytes) {
    if (e->h_out) cudaFreeHost(e->h_out);
    e->h_out = nullptr; e->h_outBytes = 0;
    ORB_CUDA(cudaHostAlloc((void**)&e->h_out, outBytes, cudaHostAllocDefault));
    e->h_outBytes = outBytes;
  }
  return ORB_OK;
}

void stage_rows(u8* dst, const u8* src, int width, int height, size_t step) {
  if (step == (size_t)width) { memcpy(dst, src, (size_t)width * height); return; }
  for (int y = 0; y < height; y++) memcpy(dst + (size_t)y * width, src + (size_t)y * step, (size_t)width);
}

int overflow_status(orb_extractor* e, int flag, cudaStream_t s) {
  if (flag) {
    cudaMemsetAsync(e->d_overflow, 0, sizeof(int), s);
    ORB_FAIL(ORB_ERR_CAPACITY, flag & 4 ? "keypoint output capacity too small" : "internal candidate/keypoint list overflow");
  }
  return ORB_OK;
}

int check_overflow(orb_extractor* e, cudaStream_t s) {
  int flag = 0;
  ORB_CUDA(cudaMemcpyAsync(&flag, e->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  if (flag) {
    cudaMemsetAsync(e->d_overflow, 0, sizeof(int), s);
    ORB_FAIL(ORB_ERR_CAPACITY, flag & 4 ? "keypoint output capacity too small" : "internal candidate/keypoint list overflow");
  }
  return ORB_OK;
}

}  // namespace

extern "C" {

const char* orb_last_error(void) { return g_last_error.c_str(); }

int orb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int orb_create(const orb_params* params, int device, int max_batch, orb_extractor** out) {
  if (!params || !out) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (params->nlevels < 1 || params->nlevels > kMaxLevels || params->nfeatures < 1 || !(params->scale_factor > 1.0f) ||
      params->min_th_fast < 1 || params->ini_th_fast < params->min_th_fast || params->ini_th_fast > 254)
    ORB_FAIL(ORB_ERR_INVALID, "bad extractor parameters");
  *out = nullptr;
  ORB_CUDA(cudaSetDevice(device));
  orb_extractor* e = new orb_extractor();
  e->p = *params;
  e->device = device;
  e->maxBatch = std::max(1, max_batch);
  build_tables(e);
  if (const char* ev = getenv("ORB_B200_LANES")) e->lanes = atoi(ev) >= 2 ? 2 : 1;
  if (const char* ev = getenv("ORB_B200_PDL")) e->pdl = atoi(ev) != 0;
  if (const char* ev = getenv("ORB_B200_RZ_SMALL_FROM")) e->rzSmallFrom = std::max(1, atoi(ev));
  if (const char* ev = getenv("ORB_B200_RZ_SMALL_BAND")) e->rzSmallBand = std::max(4, std::min(kRzBand, atoi(ev)));
  if (const char* ev = getenv("ORB_B200_BLUR_FORK")) e->blurFork = std::max(0, std::min(2, atoi(ev)));
  if (const char* ev = getenv("ORB_B200_L0_FORK")) e->l0Fork = atoi(ev) != 0;
  if (const char* ev = getenv("ORB_B200_GRAPH")) e->useGraph = atoi(ev) != 0;
  cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  for (int l = 0; l < 2 && err == cudaSuccess; l++) {
    err = cudaStreamCreateWithFlags(&e->laneStream[l], cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evJoin[l], cudaEventDisableTiming);
  }
  if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming);
  for (int l = 0; l < 2 && err == cudaSuccess; l++) {
    err = cudaStreamCreateWithFlags(&e->blurStream[l], cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evBlurGo[l], cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evBlurDone[l], cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evL0Go[l], cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->evL0Done[l], cudaEventDisableTiming);
  }
  if (err == cudaSuccess) err = cudaMalloc(&e->d_pattern, sizeof ORB_BIT_PATTERN_31);
  if (err == cudaSuccess) err = cudaMemcpy(e->d_pattern, ORB_BIT_PATTERN_31, sizeof ORB_BIT_PATTERN_31, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&e->d_overflow, sizeof(int));
  if (err == cudaSuccess) err = cudaMalloc(&e->d_work, 4 * sizeof(int));
  if (err == cudaSuccess) err = cudaMalloc(&e->d_invScale, kMaxLevels * sizeof(float));
  if (err == cudaSuccess) err = cudaMemcpy(e->d_invScale, e->invScale.data(), e->invScale.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemset(e->d_overflow, 0, sizeof(int));
  if (err != cudaSuccess) {
    delete e;
    return cuda_fail(err, "orb_create", __FILE__, __LINE__);
  }
  *out = e;
  return ORB_OK;
}

int orb_destroy(orb_extractor* e) {
  if (!e) return ORB_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  free_workspace(e);
  cudaFree(e->d_taps); cudaFree(e->d_pattern); cudaFree(e->d_overflow); cudaFree(e->d_work); cudaFree(e->d_invScale);
  cudaFree(e->d_pyrItems); cudaFree(e->d_pyrFlags);
  for (int b = 0; b < 2; b++) { cudaFree(e->d_uRight[b]); cudaFree(e->d_depth[b]); }
  for (int b = 0; b < 2; b++) {
    cudaFree(e->d_in[b]); cudaFree(e->d_kps[b]); cudaFree(e->d_desc[b]); cudaFree(e->d_n[b]);
    if (e->evIn[b]) cudaEventDestroy(e->evIn[b]);
    if (e->evDone[b]) cudaEventDestroy(e->evDone[b]);
    if (e->evOut[b]) cudaEventDestroy(e->evOut[b]);
  }
  for (int k = 0; k < 2; k++)
    if (e->callGraph[k]) cudaGraphExecDestroy(e->callGraph[k]);
  cudaFree(e->d_stereoScratch[0]); cudaFree(e->d_stereoScratch[1]);
  if (e->hostPyr) cudaFreeHost(e->hostPyr);
  if (e->h_in) cudaFreeHost(e->h_in);
  if (e->h_out) cudaFreeHost(e->h_out);
  if (e->sIn) cudaStreamDestroy(e->sIn);
  if (e->sOut) cudaStreamDestroy(e->sOut);
  for (cudaEvent_t ev : e->evPool) cudaEventDestroy(ev);
  for (int l = 0; l < 2; l++) {
    if (e->laneStream[l]) cudaStreamDestroy(e->laneStream[l]);
    if (e->evJoin[l]) cudaEventDestroy(e->evJoin[l]);
  }
  if (e->evFork) cudaEventDestroy(e->evFork);
  for (int l = 0; l < 2; l++) {
    if (e->blurStream[l]) cudaStreamDestroy(e->blurStream[l]);
    if (e->evBlurGo[l]) cudaEventDestroy(e->evBlurGo[l]);
    if (e->evBlurDone[l]) cudaEventDestroy(e->evBlurDone[l]);
    if (e->evL0Go[l]) cudaEventDestroy(e->evL0Go[l]);
    if (e->evL0Done[l]) cudaEventDestroy(e->evL0Done[l]);
  }
  if (e->evUser) cudaEventDestroy(e->evUser);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return ORB_OK;
}

int orb_get_scale_tables(const orb_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                         int32_t* per_level) {
  if (!e) ORB_FAIL(ORB_ERR_INVALID, "null handle");
  for (int l = 0; l < e->p.nlevels; l++) {
    if (scale) scale[l] = e->scale[l];
    if (inv_scale) inv_scale[l] = e->invScale[l];
    if (sigma2) sigma2[l] = e->sigma2[l];
    if (inv_sigma2) inv_sigma2[l] = e->invSigma2[l];
    if (per_level) per_level[l] = e->perLevel[l];
  }
  return ORB_OK;
}

// Per level the quadtree returns at most max(nfeatures_l + 3, 4 * nIni) nodes (a split adds <= 3; the first pass turns
// nIni roots into <= 4 * nIni children), nIni = round(width' / height') roots (ORBextractor.cc:695).
static int level_kept_cap(int nfeat, int nIni) { return std::max(nfeat + 3, 4 * nIni) + 1; }

int orb_max_keypoints(const orb_extractor* e) {
  if (!e) return 0;
  if (e->haveGeom) return e->maxKp;   // exact for the image size in use
  // no image seen yet: valid for aspect ratios up to 8:1 (orb_max_keypoints_for_size is exact)
  int m = 0;
  for (int l = 0; l < e->p.nlevels; l++) m += level_kept_cap(e->perLevel[l], 8);
  return m;
}

int orb_max_keypoints_for_size(const orb_extractor* e, int width, int height) {
  if (!e || width <= 0 || height <= 0) return 0;
  int m = 0;
  for (int l = 0; l < e->p.nlevels; l++) {
    const int w = cv_round_f((float)width * e->invScale[l]), h = cv_round_f((float)height * e->invScale[l]);
    const int Wp = w - 32, Hp = h - 32;
    if (Wp < 30 || Hp < 30) return 0;   // orb_extract would return ORB_ERR_UNSUPPORTED
    m += level_kept_cap(e->perLevel[l], std::max(1, (int)std::round((float)Wp / (float)Hp)));
  }
  return m;
}

// The other entry points reuse the staging buffers without events: let asynchronous batches finish first.
static int drain_async(orb_extractor* e) {
  e->singleCap = 0;   // whatever runs next overwrites the single-call results
  if (!e->asyncPending) return ORB_OK;
  e->asyncPending = false;
  if (e->sIn) ORB_CUDA(cudaStreamSynchronize(e->sIn));
  for (int l = 0; l < 2; l++)
    if (e->laneStream[l]) ORB_CUDA(cudaStreamSynchronize(e->laneStream[l]));
  ORB_CUDA(cudaStreamSynchronize(e->stream));
  if (e->sOut) ORB_CUDA(cudaStreamSynchronize(e->sOut));
  return ORB_OK;
}

int orb_extract_batch_device(orb_extractor* e, const uint8_t* d_images, int batch, int width, int height, size_t step,
                             size_t frame_stride, orb_keypoint* d_keypoints, int capacity, int32_t* d_counts,
                             uint8_t* d_descriptors, void* stream) {
  if (!e || !d_images || !d_keypoints || !d_counts || !d_descriptors) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (batch <= 0 || width <= 0 || height <= 0 || capacity <= 0 || step < (size_t)width) ORB_FAIL(ORB_ERR_INVALID, "bad size");
  int st = drain_async(e);
  if (st) return st;
  st = ensure_geom(e, width, height, batch);
  if (st) return st;
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  e->lastLaunches = 0;
  const int nChunks = (batch + e->wsFrames - 1) / e->wsFrames;
  const bool multi = e->lanes >= 2 && nChunks >= 2;
  if ((st = lanes_fork(e, s, nChunks))) return st;
  int ci = 0;
  for (int b0 = 0; b0 < batch; b0 += e->wsFrames, ci++) {
    const int B = std::min(e->wsFrames, batch - b0);
    const int lane = multi ? (ci & 1) : 0;
    st = run_chunk(e, d_images + (size_t)b0 * frame_stride, B, step, frame_stride, d_keypoints + (size_t)b0 * capacity,
                   capacity, d_counts + b0, d_descriptors + (size_t)b0 * capacity * 32, multi ? e->laneStream[lane] : s, lane);
    if (st) return st;
  }
  st = lanes_join(e, s, nChunks);
  if (st) return st;
  return mark_user_stream(e, s);
}

int orb_synchronize(orb_extractor* e, void* stream) {
  if (!e) ORB_FAIL(ORB_ERR_INVALID, "null handle");
  ORB_CUDA(cudaSetDevice(e->device));
  if (!stream) {   // everything the handle has in flight, including the copy streams of the host entry points
    e->asyncPending = false;
    if (e->sIn) ORB_CUDA(cudaStreamSynchronize(e->sIn));
    for (int l = 0; l < 2; l++)
      if (e->laneStream[l]) ORB_CUDA(cudaStreamSynchronize(e->laneStream[l]));
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    if (e->sOut) ORB_CUDA(cudaStreamSynchronize(e->sOut));
  }
  return check_overflow(e, stream ? (cudaStream_t)stream : e->stream);
}

int orb_last_launch_count(const orb_extractor* e) { return e ? e->lastLaunches : 0; }

int orb_last_call_breakdown(const orb_extractor* e, double* us4) {
  if (!e || !us4) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  for (int k = 0; k < 4; k++) us4[k] = e->callUs[k];
  return ORB_OK;
}

int orb_set_lanes(orb_extractor* e, int lanes) {
  if (!e || lanes < 1 || lanes > 2) ORB_FAIL(ORB_ERR_INVALID, "lanes must be 1 or 2");
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  if (lanes != e->lanes) {
    // the workspace is sized per lane count: drop it, it is re-allocated by the next call
    int2* keepTaps = e->d_taps; e->d_taps = nullptr;
    free_workspace(e);
    e->d_taps = keepTaps;
    e->lanes = lanes;
  }
  return ORB_OK;
}

int orb_set_profiling(orb_extractor* e, int enable) {
  if (!e) ORB_FAIL(ORB_ERR_INVALID, "null handle");
  e->profile = enable != 0;
  return ORB_OK;
}

int orb_get_stage_times(orb_extractor* e, double* ms5, long long* launches5) {
  if (!e || !ms5 || !launches5) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  for (size_t i = 0; i + 5 < e->evUsed; i += 6)
    for (int k = 0; k < 5; k++) {
      float ms = 0.f;
      ORB_CUDA(cudaEventElapsedTime(&ms, e->evPool[i + k], e->evPool[i + k + 1]));
      e->stageMs[k] += ms;
    }
  e->evUsed = 0;
  for (int k = 0; k < 5; k++) {
    ms5[k] = e->stageMs[k];
    launches5[k] = e->stageLaunches[k];
    e->stageMs[k] = 0;
    e->stageLaunches[k] = 0;
  }
  return ORB_OK;
}

static int batch_host_impl(orb_extractor* e, const uint8_t* images, int batch, int width, int height, size_t step,
                           size_t frame_stride, orb_keypoint* keypoints, int capacity, int32_t* counts,
                           uint8_t* descriptors, bool wait) {
  if (!e || !images || !keypoints || !counts || !descriptors) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (batch <= 0 || width <= 0 || height <= 0 || capacity <= 0 || step < (size_t)width) ORB_FAIL(ORB_ERR_INVALID, "bad size");
  e->singleCap = 0;
  int st = ensure_geom(e, width, height, batch);
  if (st) return st;
  const int chunk = e->wsFrames;
  const size_t dFrame = (size_t)width * height;
  st = ensure_stage(e, dFrame * chunk, chunk, capacity);
  if (st) return st;
  cudaStream_t s = e->stream;
  e->lastLaunches = 0;
  // the chunk counter lives in the handle: consecutive asynchronous calls continue one pipeline
  for (int b0 = 0; b0 < batch; b0 += chunk, e->hostChunks++) {
    const long long ci = e->hostChunks;
    const int B = std::min(chunk, batch - b0);
    const int b = (int)(ci & 1);
    // H2D of this chunk: its staging buffer must no longer be read by the kernels of chunk ci-2
    if (ci >= 2) ORB_CUDA(cudaStreamWaitEvent(e->sIn, e->evDone[b], 0));
    if (step == (size_t)width && frame_stride == dFrame) {
      ORB_CUDA(cudaMemcpyAsync(e->d_in[b], images + (size_t)b0 * frame_stride, dFrame * B, cudaMemcpyHostToDevice, e->sIn));
    } else {
      for (int k = 0; k < B; k++)
        ORB_CUDA(cudaMemcpy2DAsync(e->d_in[b] + (size_t)k * dFrame, width, images + (size_t)(b0 + k) * frame_stride, step,
                                   width, height, cudaMemcpyHostToDevice, e->sIn));
    }
    ORB_CUDA(cudaEventRecord(e->evIn[b], e->sIn));
    // kernels (lane b of the workspace, on its own stream): wait for the input, and for the D2H of
    // chunk ci-2 that still reads the output buffers
    cudaStream_t cs = e->lanes >= 2 ? e->laneStream[b] : s;
    const int lane = e->lanes >= 2 ? b : 0;
    ORB_CUDA(cudaStreamWaitEvent(cs, e->evIn[b], 0));
    if (ci >= 2) ORB_CUDA(cudaStreamWaitEvent(cs, e->evOut[b], 0));
    st = run_chunk(e, e->d_in[b], B, width, dFrame, e->d_kps[b], capacity, e->d_n[b], e->d_desc[b], cs, lane);
    if (st) return st;
    ORB_CUDA(cudaEventRecord(e->evDone[b], cs));
    // D2H on its own stream
    ORB_CUDA(cudaStreamWaitEvent(e->sOut, e->evDone[b], 0));
    ORB_CUDA(cudaMemcpyAsync(counts + b0, e->d_n[b], (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, e->sOut));
    ORB_CUDA(cudaMemcpyAsync(keypoints + (size_t)b0 * capacity, e->d_kps[b], (size_t)B * capacity * sizeof(orb_keypoint),
                             cudaMemcpyDeviceToHost, e->sOut));
    ORB_CUDA(cudaMemcpyAsync(descriptors + (size_t)b0 * capacity * 32, e->d_desc[b], (size_t)B * capacity * 32,
                             cudaMemcpyDeviceToHost, e->sOut));
    ORB_CUDA(cudaEventRecord(e->evOut[b], e->sOut));
  }
  if (!wait) {
    e->asyncPending = true;
    return ORB_OK;
  }
  e->asyncPending = false;
  ORB_CUDA(cudaStreamSynchronize(e->sOut));
  for (int l = 0; l < 2; l++) ORB_CUDA(cudaStreamSynchronize(e->laneStream[l]));
  return check_overflow(e, s);
}

int orb_extract_batch_host(orb_extractor* e, const uint8_t* images, int batch, int width, int height, size_t step,
                           size_t frame_stride, orb_keypoint* keypoints, int capacity, int32_t* counts,
                           uint8_t* descriptors) {
  return batch_host_impl(e, images, batch, width, height, step, frame_stride, keypoints, capacity, counts, descriptors, true);
}

int orb_extract_batch_host_async(orb_extractor* e, const uint8_t* images, int batch, int width, int height, size_t step,
                                 size_t frame_stride, orb_keypoint* keypoints, int capacity, int32_t* counts,
                                 uint8_t* descriptors) {
  return batch_host_impl(e, images, batch, width, height, step, frame_stride, keypoints, capacity, counts, descriptors, false);
}

int orb_extract(orb_extractor* e, const uint8_t* image, int width, int height, size_t step, orb_keypoint* keypoints,
                int capacity, int* n, uint8_t* descriptors, orb_level_view* pyramid) {
  if (!e || !n) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (!image || width == 0 || height == 0) return ORB_OK;  // empty image: outputs untouched (:1537)
  if (!keypoints || !descriptors || capacity <= 0) ORB_FAIL(ORB_ERR_INVALID, "null output");
  int st = drain_async(e);
  if (st) return st;
  // a handle that may serve as the left eye of orb_stereo_match (max_batch >= 2) keeps a second frame slot
  const int slots = e->maxBatch >= 2 ? 2 : 1;
  st = ensure_geom(e, width, height, slots);
  if (st) return st;
  const size_t dFrame = (size_t)width * height;
  st = ensure_stage(e, dFrame * slots, slots, capacity);
  if (st) return st;
  cudaStream_t s = e->stream;
  e->lastLaunches = 0;
  const int m = std::min(capacity, e->maxKp);      // entries that can come back
  const size_t oK = 64, oD = oK + round_up((size_t)m * sizeof(orb_keypoint), (size_t)64);
  st = ensure_pinned(e, dFrame, oD + (size_t)m * 32);
  if (st) return st;
  const auto t0 = std::chrono::steady_clock::now();
  {
    // the image goes up in two halves: the upload of the first overlaps the staging copy of the second
    const int h0 = height / 2;
    const size_t b0 = (size_t)h0 * width;
    stage_rows(e->h_in, image, width, h0, step);
    if (b0) ORB_CUDA(cudaMemcpyAsync(e->d_in[0], e->h_in, b0, cudaMemcpyHostToDevice, s));
    stage_rows(e->h_in + b0, image + (size_t)h0 * step, width, height - h0, step);
    ORB_CUDA(cudaMemcpyAsync(e->d_in[0] + b0, e->h_in + b0, dFrame - b0, cudaMemcpyHostToDevice, s));
  }
  const auto t1 = std::chrono::steady_clock::now();
  st = run_call(e, 1, width, height, dFrame, capacity, false, 0.f, 0.f, s);
  if (st) return st;
  int* hc = reinterpret_cast<int*>(e->h_out);
  ORB_CUDA(cudaMemcpyAsync(hc, e->d_n[0], sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(hc + 1, e->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oK, e->d_kps[0], (size_t)m * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oD, e->d_desc[0], (size_t)m * 32, cudaMemcpyDeviceToHost, s));
  if (pyramid) {
    if (e->pyrStride > e->hostPyrBytes) {
      if (e->hostPyr) cudaFreeHost(e->hostPyr);
      e->hostPyr = nullptr; e->hostPyrBytes = 0;
      ORB_CUDA(cudaHostAlloc((void**)&e->hostPyr, e->pyrStride, cudaHostAllocDefault));
      e->hostPyrBytes = e->pyrStride;
    }
    ORB_CUDA(cudaMemcpyAsync(e->hostPyr, e->d_pyr, e->pyrStride, cudaMemcpyDeviceToHost, s));
  }
  const auto t2 = std::chrono::steady_clock::now();
  ORB_CUDA(cudaStreamSynchronize(s));
  const auto t3 = std::chrono::steady_clock::now();
  st = overflow_status(e, hc[1], s);
  if (st) return st;
  const int cnt = std::min(hc[0], m);
  if (cnt > 0) {
    memcpy(keypoints, e->h_out + oK, (size_t)cnt * sizeof(orb_keypoint));
    memcpy(descriptors, e->h_out + oD, (size_t)cnt * 32);
  }
  *n = cnt;
  {
    const auto t4 = std::chrono::steady_clock::now();
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::micro>(b - a).count();
    };
    e->callUs[0] = us(t0, t1); e->callUs[1] = us(t1, t2); e->callUs[2] = us(t2, t3); e->callUs[3] = us(t3, t4);
  }
  e->singleCap = capacity; e->singleW = width; e->singleH = height;
  if (pyramid)
    for (int l = 0; l < e->g.nlevels; l++) {
      pyramid[l].data = e->hostPyr + e->g.lv[l].off;
      pyramid[l].width = e->g.lv[l].w;
      pyramid[l].height = e->g.lv[l].h;
      pyramid[l].step = e->g.lv[l].pitch;
    }
  return ORB_OK;
}

int orb_extract_stereo_batch_device(orb_extractor* e, const uint8_t* d_images, int pairs, int width, int height, size_t step,
                                    size_t frame_stride, orb_keypoint* d_keypoints, int capacity, int32_t* d_counts,
                                    uint8_t* d_descriptors, float mbf, float mb, float* d_uright, float* d_depth,
                                    void* stream) {
  if (!e || !d_images || !d_keypoints || !d_counts || !d_descriptors || !d_uright || !d_depth) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (pairs <= 0 || width <= 0 || height <= 0 || capacity <= 0 || step < (size_t)width || !(mb > 0.f) || !(mbf > 0.f))
    ORB_FAIL(ORB_ERR_INVALID, "bad size or stereo baseline");
  int st = drain_async(e);
  if (st) return st;
  st = ensure_geom(e, width, height, std::max(2, 2 * pairs));
  if (st) return st;
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  e->lastLaunches = 0;
  const int chunk = std::max(2, e->wsFrames & ~1);
  if (chunk > e->wsFrames) ORB_FAIL(ORB_ERR_INVALID, "max_batch must be >= 2 for stereo");
  const int batch = 2 * pairs;
  const int nChunks = (batch + chunk - 1) / chunk;
  const bool multi = e->lanes >= 2 && nChunks >= 2;
  {
    const int chunkPairs = std::min(chunk, batch) / 2;
    if (stereo_split(chunkPairs) > 1)
      for (int l = 0; l < (multi ? 2 : 1); l++)
        if ((st = ensure_stereo_scratch(e, capacity, chunkPairs, l, s))) return st;
  }
  if ((st = lanes_fork(e, s, nChunks))) return st;
  int ci = 0;
  for (int b0 = 0; b0 < batch; b0 += chunk, ci++) {
    const int B = std::min(chunk, batch - b0);
    const int lane = multi ? (ci & 1) : 0;
    cudaStream_t ls = multi ? e->laneStream[lane] : s;
    st = run_chunk(e, d_images + (size_t)b0 * frame_stride, B, step, frame_stride, d_keypoints + (size_t)b0 * capacity,
                   capacity, d_counts + b0, d_descriptors + (size_t)b0 * capacity * 32, ls, lane);
    if (st) return st;
    st = run_stereo(e, B, d_keypoints + (size_t)b0 * capacity, capacity, d_counts + b0,
                    d_descriptors + (size_t)b0 * capacity * 32, mbf, mb, d_uright + (size_t)(b0 / 2) * capacity,
                    d_depth + (size_t)(b0 / 2) * capacity, ls, lane);
    if (st) return st;
  }
  st = lanes_join(e, s, nChunks);
  if (st) return st;
  return mark_user_stream(e, s);
}

int orb_extract_stereo(orb_extractor* e, const uint8_t* left, const uint8_t* right, int width, int height, size_t step,
                       float mbf, float mb, orb_keypoint* kps_left, int capacity, int* n_left, uint8_t* desc_left,
                       orb_keypoint* kps_right, int* n_right, uint8_t* desc_right, float* uright, float* depth) {
  if (!e || !n_left || !n_right) ORB_FAIL(ORB_ERR_INVALID, "null argument");
  if (!left || !right || width == 0 || height == 0) return ORB_OK;  // empty image: outputs untouched
  if (!kps_left || !kps_right || !desc_left || !desc_right || !uright || !depth || capacity <= 0 || !(mb > 0.f) || !(mbf > 0.f))
    ORB_FAIL(ORB_ERR_INVALID, "null output or bad stereo baseline");
  if (e->maxBatch < 2) ORB_FAIL(ORB_ERR_INVALID, "max_batch must be >= 2 for stereo");
  int st = drain_async(e);
  if (st) return st;
  st = ensure_geom(e, width, height, 2);
  if (st) return st;
  const size_t dFrame = (size_t)width * height;
  st = ensure_stage(e, dFrame * 2, 2, capacity);
  if (st) return st;
  if (capacity > e->stereoOutCap) {   // grows only (a new pointer also means a new graph capture)
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_uRight[0]); cudaFree(e->d_depth[0]);
    e->d_uRight[0] = e->d_depth[0] = nullptr;
    e->stereoOutCap = 0;
    ORB_CUDA(cudaMalloc(&e->d_uRight[0], (size_t)capacity * sizeof(float)));
    ORB_CUDA(cudaMalloc(&e->d_depth[0], (size_t)capacity * sizeof(float)));
    e->stereoOutCap = capacity;
  }
  cudaStream_t s = e->stream;
  e->lastLaunches = 0;
  const int m = std::min(capacity, e->maxKp);
  const size_t szK = round_up((size_t)m * sizeof(orb_keypoint), (size_t)64), szD = (size_t)m * 32, szF = round_up((size_t)m * sizeof(float), (size_t)64);
  const size_t oKL = 64, oKR = oKL + szK, oDL = oKR + szK, oDR = oDL + szD, oU = oDR + szD, oZ = oU + szF;
  st = ensure_pinned(e, 2 * dFrame, oZ + szF);
  if (st) return st;
  st = ensure_stereo_scratch(e, capacity, 1, 0, s);
  if (st) return st;
  stage_rows(e->h_in, left, width, height, step);
  stage_rows(e->h_in + dFrame, right, width, height, step);
  ORB_CUDA(cudaMemcpyAsync(e->d_in[0], e->h_in, 2 * dFrame, cudaMemcpyHostToDevice, s));
  st = run_call(e, 2, width, height, dFrame, capacity, true, mbf, mb, s);
  if (st) return st;
  int* hc = reinterpret_cast<int*>(e->h_out);
  ORB_CUDA(cudaMemcpyAsync(hc, e->d_n[0], 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(hc + 2, e->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oKL, e->d_kps[0], (size_t)m * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oKR, e->d_kps[0] + capacity, (size_t)m * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oDL, e->d_desc[0], szD, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oDR, e->d_desc[0] + (size_t)capacity * 32, szD, cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oU, e->d_uRight[0], (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + oZ, e->d_depth[0], (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  st = overflow_status(e, hc[2], s);
  if (st) return st;
  const int cnt[2] = {std::min(hc[0], m), std::min(hc[1], m)};
  if (cnt[0] > 0) {
    memcpy(kps_left, e->h_out + oKL, (size_t)cnt[0] * sizeof(orb_keypoint));
    memcpy(desc_left, e->h_out + oDL, (size_t)cnt[0] * 32);
    memcpy(uright, e->h_out + oU, (size_t)cnt[0] * sizeof(float));
    memcpy(depth, e->h_out + oZ, (size_t)cnt[0] * sizeof(float));
  }
  if (cnt[1] > 0) {
    memcpy(kps_right, e->h_out + oKR, (size_t)cnt[1] * sizeof(orb_keypoint));
    memcpy(desc_right, e->h_out + oDR, (size_t)cnt[1] * 32);
  }
  *n_left = cnt[0];
  *n_right = cnt[1];
  return ORB_OK;
}

int orb_stereo_match(orb_extractor* left, orb_extractor* right, float mbf, float mb, float* uright, float* depth, int* n_left) {
  if (!left || !right || !uright || !depth || left == right) ORB_FAIL(ORB_ERR_INVALID, "null or identical handles");
  if (!(mb > 0.f) || !(mbf > 0.f)) ORB_FAIL(ORB_ERR_INVALID, "bad stereo baseline");
  if (left->device != right->device) ORB_FAIL(ORB_ERR_INVALID, "both extractors must live on the same device");
  if (!left->singleCap || !right->singleCap)
    ORB_FAIL(ORB_ERR_INVALID, "orb_stereo_match needs the results of an orb_extract call on both handles");
  if (left->singleCap != right->singleCap || left->singleW != right->singleW || left->singleH != right->singleH ||
      memcmp(&left->p, &right->p, sizeof(orb_params)) != 0)
    ORB_FAIL(ORB_ERR_INVALID, "left and right extractors differ in parameters, image size or capacity");
  if (left->maxBatch < 2 || left->wsFrames < 2) ORB_FAIL(ORB_ERR_INVALID, "the left extractor must be created with max_batch >= 2");
  ORB_CUDA(cudaSetDevice(left->device));
  orb_extractor* e = left;
  cudaStream_t s = e->stream;
  const int cap = e->singleCap;
  const Lane WL = lane_of(left, 0), WR = lane_of(right, 0);
  // the right eye becomes frame 1 of the left handle's chunk: pyramid, keypoints, descriptors, count (device to device)
  ORB_CUDA(cudaMemcpyAsync(WL.pyr + e->pyrStride, WR.pyr, e->pyrStride, cudaMemcpyDeviceToDevice, s));
  ORB_CUDA(cudaMemcpyAsync(e->d_kps[0] + cap, right->d_kps[0], (size_t)cap * sizeof(orb_keypoint), cudaMemcpyDeviceToDevice, s));
  ORB_CUDA(cudaMemcpyAsync(e->d_desc[0] + (size_t)cap * 32, right->d_desc[0], (size_t)cap * 32, cudaMemcpyDeviceToDevice, s));
  ORB_CUDA(cudaMemcpyAsync(e->d_n[0] + 1, right->d_n[0], sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (cap > e->stereoOutCap) {
    ORB_CUDA(cudaStreamSynchronize(s));
    cudaFree(e->d_uRight[0]); cudaFree(e->d_depth[0]);
    e->d_uRight[0] = e->d_depth[0] = nullptr;
    e->stereoOutCap = 0;
    ORB_CUDA(cudaMalloc(&e->d_uRight[0], (size_t)cap * sizeof(float)));
    ORB_CUDA(cudaMalloc(&e->d_depth[0], (size_t)cap * sizeof(float)));
    e->stereoOutCap = cap;
  }
  const int m = std::min(cap, e->maxKp);
  const size_t szF = round_up((size_t)m * sizeof(float), (size_t)64);
  int st = ensure_pinned(e, 0, 64 + 2 * szF);
  if (st) return st;
  st = ensure_stereo_scratch(e, cap, 1, 0, s);
  if (st) return st;
  e->lastLaunches = 0;
  st = run_stereo(e, 2, e->d_kps[0], cap, e->d_n[0], e->d_desc[0], mbf, mb, e->d_uRight[0], e->d_depth[0], s, 0);
  if (st) return st;
  int* hc = reinterpret_cast<int*>(e->h_out);
  ORB_CUDA(cudaMemcpyAsync(hc, e->d_n[0], sizeof(int), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + 64, e->d_uRight[0], (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaMemcpyAsync(e->h_out + 64 + szF, e->d_depth[0], (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
  ORB_CUDA(cudaStreamSynchronize(s));
  const int cnt = std::min(hc[0], m);
  if (cnt > 0) {
    memcpy(uright, e->h_out + 64, (size_t)cnt * sizeof(float));
    memcpy(depth, e->h_out + 64 + szF, (size_t)cnt * sizeof(float));
  }
  if (n_left) *n_left = cnt;
  return ORB_OK;
}

// ---- stage access (parity tests) ---------------------------------------------------------
int orb_stage_level_size(const orb_extractor* e, int level, int* width, int* height) {
  if (!e || !e->haveGeom || level < 0 || level >= e->g.nlevels) ORB_FAIL(ORB_ERR_INVALID, "no geometry / bad level");
  *width = e->g.lv[level].w;
  *height = e->g.lv[level].h;
  return ORB_OK;
}

static int stage_check(const orb_extractor* e, int frame, int level) {
  if (!e || !e->haveGeom || level < 0 || level >= e->g.nlevels || frame < 0 || frame >= e->lastChunkFrames)
    ORB_FAIL(ORB_ERR_INVALID, "no geometry / bad frame or level");
  return ORB_OK;
}

int orb_stage_copy_level(orb_extractor* e, int frame, int level, uint8_t* dst) {
  int st = stage_check(e, frame, level);
  if (st) return st;
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  const LevelGeom& L = e->g.lv[level];
  const u8* src = lane_of(e, e->lastLane).pyr + (size_t)frame * e->pyrStride + L.off - (long long)kEdge * L.pitch - kEdge;
  ORB_CUDA(cudaMemcpy2D(dst, L.w + 2 * kEdge, src, L.pitch, L.w + 2 * kEdge, L.h + 2 * kEdge, cudaMemcpyDeviceToHost));
  return ORB_OK;
}

int orb_stage_copy_blur(orb_extractor* e, int frame, int level, uint8_t* dst) {
  int st = stage_check(e, frame, level);
  if (st) return st;
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  const LevelGeom& L = e->g.lv[level];
  ORB_CUDA(cudaMemcpy2D(dst, L.w, lane_of(e, e->lastLane).blur + (size_t)frame * e->blurStride + L.boff, L.bpitch, L.w, L.h,
                        cudaMemcpyDeviceToHost));
  return ORB_OK;
}

static int copy_list(orb_extractor* e, const uint2* d_list, int cnt, int32_t* xs, int32_t* ys, int32_t* score,
                     int capacity) {
  const int m = std::min(cnt, capacity);
  if (m <= 0) return ORB_OK;
  std::vector<uint2> tmp(m);
  ORB_CUDA(cudaMemcpy(tmp.data(), d_list, (size_t)m * sizeof(uint2), cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; i++) {
    xs[i] = (int)(tmp[i].x & 0xffffu);
    ys[i] = (int)(tmp[i].x >> 16);
    score[i] = (int)tmp[i].y;
  }
  return ORB_OK;
}

int orb_stage_copy_candidates(orb_extractor* e, int frame, int level, int32_t* xs, int32_t* ys, int32_t* score,
                              int capacity, int* n) {
  int st = stage_check(e, frame, level);
  if (st) return st;
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  int cnt = 0;
  ORB_CUDA(cudaMemcpy(&cnt, lane_of(e, e->lastLane).candCount + frame * e->g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  cnt = std::min(cnt, e->g.lv[level].candCap);
  *n = cnt;
  return copy_list(e, lane_of(e, e->lastLane).cand + (size_t)frame * e->candTotal + e->g.lv[level].candOff, cnt, xs, ys, score, capacity);
}

int orb_stage_copy_kept(orb_extractor* e, int frame, int level, int32_t* xs, int32_t* ys, int32_t* score, int capacity,
                        int* n) {
  int st = stage_check(e, frame, level);
  if (st) return st;
  ORB_CUDA(cudaSetDevice(e->device));
  { const int ss_ = sync_handle(e); if (ss_) return ss_; }
  int cnt = 0;
  ORB_CUDA(cudaMemcpy(&cnt, lane_of(e, e->lastLane).keptCount + frame * e->g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  *n = cnt;
  return copy_list(e, lane_of(e, e->lastLane).kept + (size_t)frame * e->keptTotal + e->g.lv[level].keptOff, cnt, xs, ys, score, capacity);
}

}  // extern "C"
