// Probe: which TMA tile loads are legal for byte tensors (unaligned inner coordinate, box sizes)?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned char* out) {
  __shared__ __align__(128) unsigned char buf[BW * BH];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  asm volatile("fence.mbarrier_init.release.cluster;");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(buf)), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(&bar)), "r"(0));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}
template <int BW, int BH>
int run(EncodeTiledFn enc, unsigned char* d, int pitch, int rows, int frames, size_t fstride, int x, int y, int z, const unsigned char* h) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
  cuuint32_t box[3] = {BW, BH, 1}, ones[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("  encode failed %d\n", (int)r); return 1; }
  unsigned char* out; cudaMalloc(&out, BW * BH);
  k<BW, BH><<<1, 128>>>(m, x, y, z, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  box %dx%d at (%d,%d,%d): %s\n", BW, BH, x, y, z, cudaGetErrorString(e)); return 2; }
  unsigned char* ho = (unsigned char*)malloc(BW * BH);
  cudaMemcpy(ho, out, BW * BH, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; r2++)
    for (int c = 0; c < BW; c++) {
      const int yy = y + r2, xx = x + c;
      const unsigned char want = (yy >= 0 && yy < rows && xx >= 0 && xx < pitch) ? h[(size_t)z * fstride + (size_t)yy * pitch + xx] : 0;
      bad += ho[r2 * BW + c] != want;
    }
  printf("  box %dx%d at (%d,%d,%d): ok, %d mismatching bytes\n", BW, BH, x, y, z, bad);
  cudaFree(out); free(ho);
  return 0;
}
int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) { printf("no entry point\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  const int pitch = 1312, rows = 414, frames = 3; const size_t fstride = (size_t)pitch * rows;
  unsigned char* h = (unsigned char*)malloc(fstride * frames);
  for (size_t i = 0; i < fstride * frames; i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char* d; cudaMalloc(&d, fstride * frames); cudaMemcpy(d, h, fstride * frames, cudaMemcpyHostToDevice);
  const int xs[] = {64, 48, 36, 33, 7, -3};
  for (int xi = 0; xi < 6; xi++) {
    if (run<32, 31>(enc, d, pitch, rows, frames, fstride, xs[xi], 100, 1, h) == 2) return 0;
    if (run<48, 37>(enc, d, pitch, rows, frames, fstride, xs[xi], 21, 2, h) == 2) return 0;
    if (run<64, 8>(enc, d, pitch, rows, frames, fstride, xs[xi], 5, 0, h) == 2) return 0;
  }
  return 0;
}
