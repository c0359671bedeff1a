// orb_frame.cu — the step right after extraction in Frame's constructors, device resident
// (reference: src/Frame.cc): UndistortKeyPoints :724-776 and ComputeImageBounds :779-829
// (cv::undistortPoints with P = K), AssignFeaturesToGrid :399-423 / PosInGrid :682-698,
// GetFeaturesInArea :590-670. Keeps the keypoints of a batch on the device between extraction and
// the grid-guided matchers. SURVEY.md 8(f) "next #2".
#include <cmath>
#include <cstring>

#include "orb_common.cuh"

using namespace orbb200;

namespace {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;  // FRAME_GRID_COLS / ROWS (Frame.h:55-60)

// cv::undistortPoints(src, dst, K, dist, Mat(), K): iterative inverse of the (k1,k2,p1,p2,k3)
// model, 5 fixed iterations in double (default TermCriteria(COUNT, 5, 0.01)), re-projected with
// P = K. Evaluated in OpenCV's operation order without FMA contraction.
__host__ __device__ inline void undistort_point(const orb_camera& c, float u, float v, float& ou, float& ov) {
  const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy, k1 = c.k1, k2 = c.k2, p1 = c.p1, p2 = c.p2, k3 = c.k3;
  const double ifx = 1. / fx, ify = 1. / fy;
#ifdef __CUDA_ARCH__
#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DD(a, b) __ddiv_rn((a), (b))
#else
#define DM(a, b) ((a) * (b))
#define DA(a, b) ((a) + (b))
#define DD(a, b) ((a) / (b))
#endif
  double x = DM(DA((double)u, -cx), ifx), y = DM(DA((double)v, -cy), ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = DA(DM(x, x), DM(y, y));
    const double den = DA(1., DM(DA(DM(DA(DM(k3, r2), k2), r2), k1), r2));   // 1 + ((k3*r2 + k2)*r2 + k1)*r2
    const double icdist = DD(1., den);                                       // numerator: k4..k6 are zero
    if (icdist < 0) { x = DM(DA((double)u, -cx), ifx); y = DM(DA((double)v, -cy), ify); break; }
    const double dX = DA(DM(DM(DM(2., p1), x), y), DM(p2, DA(r2, DM(DM(2., x), x))));
    const double dY = DA(DM(p1, DA(r2, DM(DM(2., y), y))), DM(DM(DM(2., p2), x), y));
    x = DM(DA(x0, -dX), icdist);
    y = DM(DA(y0, -dY), icdist);
  }
  ou = (float)DA(DM(fx, x), cx);
  ov = (float)DA(DM(fy, y), cy);
#undef DM
#undef DA
#undef DD
}

__global__ void k_undistort(const orb_keypoint* __restrict__ in, const int* __restrict__ counts, int cap, const orb_camera cam,
                            orb_keypoint* __restrict__ out) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[f]) return;
  orb_keypoint k = in[(size_t)f * cap + i];
  if (cam.k1 != 0.0f) undistort_point(cam, k.x, k.y, k.x, k.y);   // mDistCoef.at<float>(0)==0 -> mvKeysUn = mvKeys
  out[(size_t)f * cap + i] = k;
}

// Frame::PosInGrid
__device__ __forceinline__ bool pos_in_grid(float x, float y, float minX, float minY, float invW, float invH, int& gx, int& gy) {
  gx = (int)roundf(__fmul_rn(__fsub_rn(x, minX), invW));
  gy = (int)roundf(__fmul_rn(__fsub_rn(y, minY), invH));
  return gx >= 0 && gx < kGridCols && gy >= 0 && gy < kGridRows;
}

// One CTA per frame: mGrid[ix][iy] as CSR (cell = ix*48 + iy), items in insertion (= index) order.
__global__ void __launch_bounds__(256) k_grid_assign(const orb_keypoint* __restrict__ kps, const int* __restrict__ counts, int cap,
                                                     float minX, float minY, float invW, float invH,
                                                     int* __restrict__ cellStart, int* __restrict__ cellItems) {
  __shared__ int cnt[kGridCells + 1];
  __shared__ int wsum[9];
  const int f = blockIdx.x, tid = threadIdx.x, n = counts[f];
  const orb_keypoint* K = kps + (size_t)f * cap;
  int* start = cellStart + (size_t)f * (kGridCells + 1);
  int* items = cellItems + (size_t)f * cap;
  for (int i = tid; i <= kGridCells; i += 256) cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    int gx, gy;
    if (pos_in_grid(K[i].x, K[i].y, minX, minY, invW, invH, gx, gy)) atomicAdd(&cnt[gx * kGridRows + gy], 1);
  }
  __syncthreads();
  // exclusive scan of 3073 entries: 12 (+1) per thread, then across threads
  {
    const int per = (kGridCells + 1 + 255) / 256;
    const int b = min(tid * per, kGridCells + 1), e = min(b + per, kGridCells + 1);
    int s = 0;
    for (int i = b; i < e; i++) s += cnt[i];
    const int lane = tid & 31, wid = tid >> 5;
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (tid == 0) {
      int acc = 0;
      for (int w = 0; w < 8; w++) { const int t = wsum[w]; wsum[w] = acc; acc += t; }
    }
    __syncthreads();
    int base = wsum[wid] + incl - s;
    for (int i = b; i < e; i++) { const int t = cnt[i]; cnt[i] = base; base += t; }
  }
  __syncthreads();
  for (int i = tid; i <= kGridCells; i += 256) start[i] = cnt[i];
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    int gx, gy;
    if (pos_in_grid(K[i].x, K[i].y, minX, minY, invW, invH, gx, gy)) items[atomicAdd(&cnt[gx * kGridRows + gy], 1)] = i;
  }
  __syncthreads();
  // restore insertion order inside every cell (cells are short: insertion sort)
  for (int c = tid; c < kGridCells; c += 256) {
    const int b = start[c], e = start[c + 1];
    for (int i = b + 1; i < e; i++) {
      const int v = items[i];
      int j = i - 1;
      while (j >= b && items[j] > v) { items[j + 1] = items[j]; j--; }
      items[j + 1] = v;
    }
  }
}

// One warp per query; cells are visited ix-outer / iy-inner and items in cell order, exactly like
// the reference, so the output order (and with it first-wins tie breaking of the best-only
// matchers) is reproduced.
__global__ void __launch_bounds__(128) k_features_in_area(const orb_keypoint* __restrict__ kps, int cap,
                                                          const int* __restrict__ cellStart, const int* __restrict__ cellItems,
                                                          float minX, float minY, float invW, float invH,
                                                          const orb_area_query* __restrict__ queries, int nq,
                                                          int* __restrict__ out, int outCap, int* __restrict__ outCount) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= nq) return;
  const orb_area_query Q = queries[q];
  const orb_keypoint* K = kps + (size_t)Q.frame * cap;
  const int* start = cellStart + (size_t)Q.frame * (kGridCells + 1);
  const int* items = cellItems + (size_t)Q.frame * cap;
  int* o = out + (size_t)q * outCap;
  int n = 0;
  const float x = Q.x, y = Q.y, r = Q.r;
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
  const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
  const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
  if (cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0) {
    const bool check = Q.min_level > 0 || Q.max_level >= 0;
    const float r2 = __fmul_rn(r, r);
    for (int ix = cx0; ix <= cx1; ix++) {
      // cells (ix, cy0..cy1) are contiguous in the CSR layout
      const int b = start[ix * kGridRows + cy0], e = start[ix * kGridRows + cy1 + 1];
      for (int i0 = b; i0 < e; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        int idx = 0;
        if (i < e) {
          idx = items[i];
          const orb_keypoint k = K[idx];
          keep = true;
          if (check) {
            if (k.octave < Q.min_level) keep = false;
            if (Q.max_level >= 0 && k.octave > Q.max_level) keep = false;
          }
          const float dx = __fsub_rn(k.x, x), dy = __fsub_rn(k.y, y);
          keep = keep && __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2;   // this fork: circular window (:664)
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int pos = n + __popc(m & ((1u << lane) - 1u));
          if (pos < outCap) o[pos] = idx;
        }
        n += __popc(m);
      }
    }
  }
  if (lane == 0) outCount[q] = n;
}

}  // namespace

extern "C" {

int orb_compute_image_bounds(const orb_camera* cam, int width, int height, float* bounds4) {
  if (!cam || !bounds4 || width <= 0 || height <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  if (cam->k1 != 0.0f) {
    float x[4], y[4];
    const float px[4] = {0.f, (float)width, 0.f, (float)width}, py[4] = {0.f, 0.f, (float)height, (float)height};
    for (int i = 0; i < 4; i++) undistort_point(*cam, px[i], py[i], x[i], y[i]);
    bounds4[0] = fminf(x[0], x[2]); bounds4[1] = fmaxf(x[1], x[3]);
    bounds4[2] = fminf(y[0], y[1]); bounds4[3] = fmaxf(y[2], y[3]);
  } else {
    bounds4[0] = 0.f; bounds4[1] = (float)width; bounds4[2] = 0.f; bounds4[3] = (float)height;
  }
  return ORB_OK;
}

int orb_undistort_keypoints_device(int device, const orb_keypoint* d_keypoints, const int32_t* d_counts, int batch, int capacity,
                                   const orb_camera* cam, orb_keypoint* d_keypoints_un, void* stream) {
  if (!d_keypoints || !d_counts || !cam || !d_keypoints_un || batch <= 0 || capacity <= 0) ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(device));
  k_undistort<<<dim3((capacity + 127) / 128, batch), 128, 0, (cudaStream_t)stream>>>(d_keypoints, d_counts, capacity, *cam, d_keypoints_un);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_assign_features_to_grid_device(int device, const orb_keypoint* d_keypoints_un, const int32_t* d_counts, int batch,
                                       int capacity, const float* bounds4, int32_t* d_cell_start, int32_t* d_cell_items,
                                       void* stream) {
  if (!d_keypoints_un || !d_counts || !bounds4 || !d_cell_start || !d_cell_items || batch <= 0 || capacity <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(device));
  const float invW = (float)kGridCols / (bounds4[1] - bounds4[0]), invH = (float)kGridRows / (bounds4[3] - bounds4[2]);  // Frame.cc:184-186
  k_grid_assign<<<batch, 256, 0, (cudaStream_t)stream>>>(d_keypoints_un, d_counts, capacity, bounds4[0], bounds4[2], invW, invH,
                                                         d_cell_start, d_cell_items);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

int orb_get_features_in_area_device(int device, const orb_keypoint* d_keypoints_un, int capacity, const float* bounds4,
                                    const int32_t* d_cell_start, const int32_t* d_cell_items, const orb_area_query* d_queries,
                                    int n_queries, int32_t* d_out, int out_capacity, int32_t* d_out_counts, void* stream) {
  if (!d_keypoints_un || !bounds4 || !d_cell_start || !d_cell_items || !d_queries || !d_out || !d_out_counts || n_queries <= 0 ||
      out_capacity <= 0)
    ORB_FAIL(ORB_ERR_INVALID, "bad argument");
  ORB_CUDA(cudaSetDevice(device));
  const float invW = (float)kGridCols / (bounds4[1] - bounds4[0]), invH = (float)kGridRows / (bounds4[3] - bounds4[2]);
  k_features_in_area<<<(n_queries * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      d_keypoints_un, capacity, d_cell_start, d_cell_items, bounds4[0], bounds4[2], invW, invH, d_queries, n_queries, d_out,
      out_capacity, d_out_counts);
  ORB_CUDA(cudaGetLastError());
  return ORB_OK;
}

}  // extern "C"
