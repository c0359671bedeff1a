// Shared host/device helpers for liborb_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/orb_b200.h"

typedef unsigned char u8;

namespace orbb200 {

void set_last_error(const std::string& msg);

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  set_last_error(buf);
  return ORB_ERR_CUDA;
}

#define ORB_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) return orbb200::cuda_fail(e_, #call, __FILE__, __LINE__);    \
  } while (0)

#define ORB_FAIL(code, msg)            \
  do {                                 \
    orbb200::set_last_error(msg);      \
    return (code);                     \
  } while (0)

template <typename T>
inline T round_up(T v, T m) { return (v + m - 1) / m * m; }

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION (per device), not to a handle: several live
// handles with different geometries (the reference keeps ORBextractor(nFeatures) and ORBextractor(2*nFeatures) alive
// side by side, src/Tracking.cc:175-188) and several host threads (src/Frame.cc:146-154) share it. It is therefore
// only ever RAISED, under a lock, to the largest size any launch has asked for so far (defined in orb_extract.cu).
cudaError_t raise_dynamic_smem_impl(const void* func, size_t bytes);
template <typename K>
inline cudaError_t raise_dynamic_smem(K* kernel, size_t bytes) {
  return raise_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}


#ifdef __CUDACC__
// ---- TMA (cp.async.bulk[.tensor]) + mbarrier helpers (sm_90+; this library is sm_100a only)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned addresses and size), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

}  // namespace orbb200
