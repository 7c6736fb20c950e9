// Control for scripts/emu_sanitize.sh: a kernel with a missing __syncthreads must be reported by ThreadSanitizer when it runs
// on the emulator (and must be clean with the barrier, -DWITH_BARRIER) -- otherwise a clean report of the real kernels says nothing.
#include "cuda_emu.h"
#include <cstdio>

__global__ void neighbour_exchange(int *out) {
  __shared__ int s[256];
  s[threadIdx.x] = (int)threadIdx.x;
#ifdef WITH_BARRIER
  __syncthreads();
#endif
  out[threadIdx.x] = s[(threadIdx.x + 1) & 255];
}

int main() {
  int out[256];
  emu::launch(dim3(2), dim3(256), 0, [&] { neighbour_exchange(out); });
  std::printf("selftest done %d\n", out[3]);
  return 0;
}
