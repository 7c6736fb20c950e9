// Shared helpers for libocrb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/ocrb200.h"

namespace ocrb {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return OCRB_ECUDA;
  }
  return OCRB_OK;
}

#define OCRB_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::ocrb::set_error(__VA_ARGS__);      \
      return OCRB_EINVAL;                  \
    }                                      \
  } while (0)

#define OCRB_CUDA(call)                                                   \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) {                                             \
      ::ocrb::set_error("%s: %s", #call, cudaGetErrorString(e__));        \
      return OCRB_ECUDA;                                                  \
    }                                                                     \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace ocrb
