// Experiment: cost of a kernel boundary inside a CUDA graph on B200: normal edges vs programmatic (PDL) edges,
// and a software grid barrier inside one persistent kernel.  nvcc -arch=sm_100a -O3 boundary.cu -o boundary
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(int *buf, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  if (threadIdx.x == 0) buf[blockIdx.x] += 1;
}
__global__ void k_persist(int *buf, unsigned *bar, int phases) {
  // software grid barrier: one atomic per CTA per phase, spin on the counter
  for (int p = 0; p < phases; ++p) {
    if (threadIdx.x == 0) {
      buf[blockIdx.x] += 1;
      __threadfence();
      atomicAdd(bar, 1u);
      const unsigned target = (unsigned)(p + 1) * gridDim.x;
      while (*((volatile unsigned *)bar) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
}
int main() {
  int *buf; unsigned *bar;
  cudaMalloc(&buf, 4096); cudaMemset(buf, 0, 4096);
  cudaMalloc(&bar, 4); cudaMemset(bar, 0, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  const int N = 200, grid = 148, threads = 320;
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < N; ++i) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = threads; cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = pdl;
      cudaLaunchKernelEx(&cfg, k_chain, buf, pdl);
    }
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st); for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st); cudaEventRecord(e1, st);
    cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("graph chain of %d kernels (grid %d x %d thr), pdl=%d: %.2f us per kernel boundary\n", N, grid, threads, pdl, ms * 1e3 / (5 * N));
  }
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_persist<<<grid, threads, 0, st>>>(buf, bar, 10); cudaStreamSynchronize(st); cudaMemset(bar, 0, 4);
    cudaEventRecord(e0, st); k_persist<<<grid, threads, 0, st>>>(buf, bar, 1000); cudaEventRecord(e1, st);
    cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("persistent kernel, software grid barrier: %.2f us per barrier (%s)\n", ms * 1e3 / 1000, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
