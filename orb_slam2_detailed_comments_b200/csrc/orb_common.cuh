// Shared host/device helpers for liborb_b200.so (sm_100a only).
#pragma once
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

}  // namespace orbb200
