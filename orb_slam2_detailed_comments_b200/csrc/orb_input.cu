// orb_input.cu — the step before the extraction path, and the other all-pairs Hamming consumer:
//   k_cvt_gray        cv::cvtColor(..., CV_{RGB,BGR,RGBA,BGRA}2GRAY)   (src/Tracking.cc:250-276, 310-324, 369-383)
//   k_remap_linear    cv::remap(..., INTER_LINEAR), CV_32FC1 maps       (Examples/Stereo/stereo_euroc.cc:181-188)
//   k_distinctive     MapPoint::ComputeDistinctiveDescriptors           (src/MapPoint.cc:365-448)
// so that raw camera frames can stay on the device from the copy engine to the keypoints. HBM-bound byte work:
// coalesced word loads / stores, no tensor cores by design.
#include <algorithm>
#include <climits>
#include <cstring>

#include "orb_common.cuh"

using namespace orbb200;

namespace {

// OpenCV 4.x RGB2Gray<uchar>: 15-bit coefficients, rounded
__device__ __forceinline__ unsigned gray15(unsigned r, unsigned g, unsigned b) {
  return (r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15;
}

// thread = 16 consecutive pixels in the LINEAR pixel order of the whole batch: when source and destination are
// densely packed (the usual case for a batch of raw frames) every thread does three (RGB) or four (RGBA) 16-byte
// loads and one 16-byte store whatever the image width; padded rows / odd strides take the per-pixel path.
template <int CH>
__global__ void __launch_bounds__(256) k_cvt_gray(const u8* __restrict__ src, int w, int h, size_t sstep, size_t sframe, int blueFirst,
                                                  u8* __restrict__ dst, size_t dstep, size_t dframe, long long total, int dense) {
  const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (p0 >= total) return;
  const int ri = blueFirst ? 2 : 0, bi = blueFirst ? 0 : 2;
  if (dense && p0 + 16 <= total) {
    const u8* s = src + p0 * CH;
    unsigned in[4 * CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(s) + i);
      in[4 * i] = v.x; in[4 * i + 1] = v.y; in[4 * i + 2] = v.z; in[4 * i + 3] = v.w;
    }
    unsigned out[4];
#pragma unroll
    for (int p = 0; p < 16; p++) {
      unsigned c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int byte = p * CH + k;
        c[k] = (in[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
      }
      const unsigned g = gray15(c[ri], c[1], c[bi]);
      if ((p & 3) == 0) out[p >> 2] = g;
      else out[p >> 2] |= g << (8 * (p & 3));
    }
    *reinterpret_cast<uint4*>(dst + p0) = make_uint4(out[0], out[1], out[2], out[3]);
  } else {
    const long long plane = (long long)w * h;
    for (long long p = p0; p < min(p0 + 16, total); p++) {
      const long long f = p / plane, r = p - f * plane;
      const int y = (int)(r / w), x = (int)(r - (long long)y * w);
      const u8* s = src + (size_t)f * sframe + (size_t)y * sstep + (size_t)x * CH;
      dst[(size_t)f * dframe + (size_t)y * dstep + x] = (u8)gray15(s[ri], s[1], s[bi]);
    }
  }
}

// Densely packed batches: one CTA converts 4096 consecutive pixels. The 12 / 16 KB of interleaved source bytes are
// brought in by ONE bulk async copy (cp.async.bulk, the TMA unit: no per-thread load instructions and no L1 tag
// traffic; the per-thread 16-byte loads at a 48-byte stride of the kernel above keep the LSU queue full at 40 % of
// the HBM rate). RGB: thread t reads its 48 bytes with three conflict-free LDS.128 and stores one STG.128; RGBA:
// thread t converts the 16-byte groups t, t+256, t+512, t+768 and stores four coalesced words.
constexpr int kCvtTilePx = 4096;

template <int CH>
__global__ void __launch_bounds__(256) k_cvt_gray_bulk(const u8* __restrict__ src, int blueFirst, u8* __restrict__ dst) {
  __shared__ __align__(128) u8 buf[kCvtTilePx * CH];
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  const size_t p0 = (size_t)blockIdx.x * kCvtTilePx;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, kCvtTilePx * CH);
    bulk_load_1d(buf, src + p0 * CH, kCvtTilePx * CH, &bar);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  const int ri = blueFirst ? 2 : 0, bi = blueFirst ? 0 : 2;
  if (CH == 3) {
    const uint4* q = reinterpret_cast<const uint4*>(buf + tid * 48);
    const uint4 v0 = q[0], v1 = q[1], v2 = q[2];
    const unsigned in[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    unsigned out[4];
#pragma unroll
    for (int p = 0; p < 16; p++) {
      unsigned c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int byte = p * 3 + k;
        c[k] = (in[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
      }
      const unsigned g = gray15(c[ri], c[1], c[bi]);
      if ((p & 3) == 0) out[p >> 2] = g;
      else out[p >> 2] |= g << (8 * (p & 3));
    }
    *reinterpret_cast<uint4*>(dst + p0 + tid * 16) = make_uint4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint4 v = *reinterpret_cast<const uint4*>(buf + (size_t)(i * 256 + tid) * 16);
      const unsigned px[4] = {v.x, v.y, v.z, v.w};
      unsigned o = 0;
#pragma unroll
      for (int k = 0; k < 4; k++)
        o |= gray15((px[k] >> (8 * ri)) & 0xffu, (px[k] >> 8) & 0xffu, (px[k] >> (8 * bi)) & 0xffu) << (8 * k);
      *reinterpret_cast<unsigned*>(dst + p0 + (size_t)(i * 256 + tid) * 4) = o;
    }
  }
}

// cv::remap, bilinear, fixed point: coordinates rounded to 1/32 px (cvRound(x*32)), weights (32-fy)(32-fx)*32 ...
// (OpenCV's table rint((1-fy)(1-fx)*2^15) is exact in float; its only saturated entry, 32768 at fx = fy = 0, gives the
// same pixel), result (sum + 2^14) >> 15, taps outside the source = 0 (BORDER_CONSTANT). thread = 4 consecutive
// output pixels in the linear order of the plane (the maps are dense, so no row arithmetic is needed) of up to
// kRemapFrames frames: the maps are shared by all frames, so coordinates and weights are computed once.
constexpr int kRemapFrames = 4;

__global__ void __launch_bounds__(256) k_remap_linear(const u8* __restrict__ src, int sw, int sh, size_t sstep, size_t sframe,
                                                      const float* __restrict__ mapx, const float* __restrict__ mapy, int dw, int dh,
                                                      u8* __restrict__ dst, size_t dstep, size_t dframe, int batch, int denseDst) {
  const int plane = dw * dh;
  const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, f0 = blockIdx.y * kRemapFrames;
  if (l4 >= plane) return;
  const int nf = min(kRemapFrames, batch - f0);
  unsigned g[kRemapFrames][4];
#pragma unroll
  for (int k = 0; k < kRemapFrames; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) g[k][i] = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (l4 + i >= plane) continue;
    const int sx = __float2int_rn(__fmul_rn(__ldg(mapx + l4 + i), 32.f)), sy = __float2int_rn(__fmul_rn(__ldg(mapy + l4 + i), 32.f));
    const int X = min(32767, max(-32768, sx >> 5)), Y = min(32767, max(-32768, sy >> 5));
    const int fx = sx & 31, fy = sy & 31;
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    const bool inside = (unsigned)X < (unsigned)(sw - 1) && (unsigned)Y < (unsigned)(sh - 1);
    const bool y0 = Y >= 0 && Y < sh, y1 = Y + 1 >= 0 && Y + 1 < sh, c0 = X >= 0 && X < sw, c1 = X + 1 >= 0 && X + 1 < sw;
    const long long o00 = (long long)Y * (long long)sstep + X;
#pragma unroll
    for (int k = 0; k < kRemapFrames; k++) {
      if (k >= nf) break;
      const u8* p = src + (size_t)(f0 + k) * sframe + o00;
      int p00 = 0, p01 = 0, p10 = 0, p11 = 0;
      if (inside) {
        p00 = p[0]; p01 = p[1]; p10 = p[sstep]; p11 = p[sstep + 1];
      } else {
        if (y0 && c0) p00 = p[0];
        if (y0 && c1) p01 = p[1];
        if (y1 && c0) p10 = p[sstep];
        if (y1 && c1) p11 = p[sstep + 1];
      }
      g[k][i] = (unsigned)((p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15);
    }
  }
#pragma unroll
  for (int k = 0; k < kRemapFrames; k++) {
    if (k >= nf) break;
    u8* D = dst + (size_t)(f0 + k) * dframe;
    if (denseDst && l4 + 4 <= plane) {
      *reinterpret_cast<unsigned*>(D + l4) = g[k][0] | (g[k][1] << 8) | (g[k][2] << 16) | (g[k][3] << 24);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (l4 + i < plane) {
          const int y = (l4 + i) / dw, x = (l4 + i) - y * dw;
          D[(size_t)y * dstep + x] = (u8)g[k][i];
        }
    }
  }
}

// MapPoint::ComputeDistinctiveDescriptors: one warp per map point, lanes own rows of the N x N distance matrix
// (descriptor j is a broadcast load); the row's distances go to the warp's shared-memory slice, its median
// (the element of rank (int)(0.5*(N-1))) is found by bisection on the value; smallest median, first row wins.
constexpr int kDistWarps = 4;

__device__ __forceinline__ int hamming_words(const unsigned a[8], const uint4 b0, const uint4 b1) {
  return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) + __popc(a[4] ^ b1.x) +
         __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

__global__ void __launch_bounds__(32 * kDistWarps) k_distinctive(const u8* __restrict__ desc, const int* __restrict__ offsets, int nPoints,
                                                                 int maxObs, int* __restrict__ bestIdx, u8* __restrict__ bestDesc) {
  extern __shared__ __align__(16) unsigned short dsm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * kDistWarps + wid;
  if (p >= nPoints) return;
  unsigned short* rows = dsm + (size_t)wid * 32 * maxObs;   // maxObs x 32, lane-minor: conflict-free
  const int o = offsets[p], N = offsets[p + 1] - o;
  const u8* D = desc + (size_t)o * 32;
  if (N > maxObs) {   // caller's bound is wrong: report instead of overrunning shared memory
    if (lane == 0) bestIdx[p] = -2;
    return;
  }
  const int rank = (int)(0.5 * (N - 1));
  unsigned best = 0xffffffffu;
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    unsigned key = 0xffffffffu;
    if (i < N) {
      unsigned a[8];
      const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(D + (size_t)i * 32)), a1 = __ldg(reinterpret_cast<const uint4*>(D + (size_t)i * 32) + 1);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      unsigned short* r = rows + lane;
      for (int j = 0; j < N; j++) {
        const uint4* q = reinterpret_cast<const uint4*>(D + (size_t)j * 32);
        r[j * 32] = (unsigned short)hamming_words(a, __ldg(q), __ldg(q + 1));
      }
      // smallest v with #{d <= v} > rank
      int lo = 0, hi = 256;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
        for (int j = 0; j < N; j++) c += r[j * 32] <= mid;
        if (c > rank) hi = mid; else lo = mid + 1;
      }
      key = ((unsigned)lo << 20) | (unsigned)i;
    }
    best = min(best, __reduce_min_sync(0xffffffffu, key));
  }
  if (N <= 0) {
    if (lane == 0) bestIdx[p] = -1;
    return;
  }
  const int bi = (int)(best & 0xfffffu);
  if (lane == 0) bestIdx[p] = bi;
  if (bestDesc && lane < 8) reinterpret_cast<unsigned*>(bestDesc + (size_t)p * 32)[lane] = reinterpret_cast<const unsigned*>(D + (size_t)bi * 32)[lane];
}

}  // namespace

extern "C" {

int orb_cvt_color_gray_device(int device, const uint8_t* d_src, int width, int height, size_t src_step, size_t src_frame_stride,
                              int batch, int code, uint8_t* d_gray, size_t gray_step, size_t gray_frame_stride, void* stream) {
  if (!d_src || !d_gray || width <= 0 || height <= 0 || batch <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  if (code < ORB_RGB2GRAY || code > ORB_BGRA2GRAY) ORB_FAIL(ORB_ERR_INVALID, "unknown colour conversion code");
  if (batch > 65535) ORB_FAIL(ORB_ERR_UNSUPPORTED, "more than 65535 frames per call");
  const int ch = (code == ORB_RGBA2GRAY || code == ORB_BGRA2GRAY) ? 4 : 3;
  const int blueFirst = (code == ORB_BGR2GRAY || code == ORB_BGRA2GRAY) ? 1 : 0;
  if (src_step < (size_t)width * ch || gray_step < (size_t)width) ORB_FAIL(ORB_ERR_INVALID, "row step smaller than a row");
  ORB_CUDA(cudaSetDevice(device));
  const long long total = (long long)batch * width * height;
  const int dense = src_step == (size_t)width * ch && src_frame_stride == src_step * height && gray_step == (size_t)width &&
                    gray_frame_stride == gray_step * height && ((size_t)d_src & 15) == 0 && ((size_t)d_gray & 15) == 0;
  long long done = 0;
  if (dense && total >= kCvtTilePx) {   // whole 4096-pixel tiles through the bulk-copy kernel
    const long long tiles = total / kCvtTilePx;
    if (tiles > 0x7fffffffLL) ORB_FAIL(ORB_ERR_UNSUPPORTED, "batch too large");
    if (ch == 4) k_cvt_gray_bulk<4><<<(unsigned)tiles, 256, 0, (cudaStream_t)stream>>>(d_src, blueFirst, d_gray);
    else k_cvt_gray_bulk<3><<<(unsigned)tiles, 256, 0, (cudaStream_t)stream>>>(d_src, blueFirst, d_gray);
    ORB_CUDA(cudaGetLastError());
    done = tiles * kCvtTilePx;
  }
  if (done < total) {   // the tail of a dense batch, or everything when rows are padded
    const long long rest = total - done;
    const unsigned blocks = (unsigned)(((rest + 15) / 16 + 255) / 256);
    // in dense mode the tail is itself a dense run of pixels that starts 16-byte aligned (4096 * ch bytes per tile)
    const u8* s0 = dense ? d_src + done * ch : d_src;
    u8* g0 = dense ? d_gray + done : d_gray;
    if (ch == 4)
      k_cvt_gray<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(s0, width, height, src_step, src_frame_stride, blueFirst, g0, gray_step,
                                                              gray_frame_stride, rest, dense);
    else
      k_cvt_gray<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(s0, width, height, src_step, src_frame_stride, blueFirst, g0, gray_step,
                                                              gray_frame_stride, rest, dense);
    ORB_CUDA(cudaGetLastError());
  }
  return ORB_OK;
}

int orb_remap_linear_device(int device, const uint8_t* d_src, int src_width, int src_height, size_t src_step, size_t src_frame_stride,
                            int batch, const float* d_map_x, const float* d_map_y, int dst_width, int dst_height, uint8_t* d_dst,
                            size_t dst_step, size_t dst_frame_stride, void* stream) {
  if (!d_src || !d_map_x || !d_map_y || !d_dst || src_width <= 0 || src_height <= 0 || dst_width <= 0 || dst_height <= 0 || batch <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  if (batch > 65535 * kRemapFrames) ORB_FAIL(ORB_ERR_UNSUPPORTED, "too many frames per call");
  if (src_step < (size_t)src_width || dst_step < (size_t)dst_width) ORB_FAIL(ORB_ERR_INVALID, "row step smaller than a row");
  ORB_CUDA(cudaSetDevice(device));
  if ((long long)dst_width * dst_height > (1ll << 30)) ORB_FAIL(ORB_ERR_UNSUPPORTED, "destination plane too large");
  const int plane = dst_width * dst_height;
  const int denseDst = dst_step == (size_t)dst_width && (dst_frame_stride & 3) == 0 && ((size_t)d_dst & 3) == 0;
  const dim3 grid(((plane + 3) / 4 + 255) / 256, (batch + kRemapFrames - 1) / kRemapFrames);
  k_remap_linear<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_width, src_height, src_step, src_frame_stride, d_map_x, d_map_y,
                                                         dst_width, dst_height, d_dst, dst_step, dst_frame_stride, batch, denseDst);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_distinctive_descriptors_device(int device, const uint8_t* d_descriptors, const int32_t* d_offsets, int n_points,
                                       int max_observations, int32_t* d_best_index, uint8_t* d_best_descriptor, void* stream) {
  if (!d_descriptors || !d_offsets || !d_best_index || n_points <= 0 || max_observations <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  const size_t smem = (size_t)kDistWarps * 32 * max_observations * sizeof(unsigned short);
  if (smem > 200 * 1024) ORB_FAIL(ORB_ERR_UNSUPPORTED, "more than 800 observations per map point");
  ORB_CUDA(cudaSetDevice(device));
  ORB_CUDA(raise_dynamic_smem(k_distinctive, smem));
  k_distinctive<<<(n_points + kDistWarps - 1) / kDistWarps, 32 * kDistWarps, smem, (cudaStream_t)stream>>>(
      d_descriptors, d_offsets, n_points, max_observations, d_best_index, d_best_descriptor);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

}  // extern "C"
