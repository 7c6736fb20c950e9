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

// ───────────── programmatic dependent launch (PDL) ─────────────
// A kernel launched through launch_pdl() may begin while its predecessor in the stream is still running; it
// must execute pdl_wait() before it touches anything the predecessor (or an earlier kernel) writes, and should
// call pdl_launch_dependents() as early as possible so that ITS successor can start its independent prologue
// (barrier init, TMEM allocation, weight prefetch).  OCRB_PDL=0 in the environment turns the attribute off.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled(int bit = 1);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_bit(int bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(bit) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_bit(1, kernel, grid, block, smem, st, args...);
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace ocrb
