// One-shot all-reduce over NVLink peer memory for the tensor-parallel decode step (BASELINE configs[4]).
//
// After a row-parallel GEMM (o_proj / down_proj: HF base_model_tp_plan "rowwise") every rank holds a [B, hidden] bf16
// partial product; the residual stream needs x += sum over ranks.  The messages are tiny (B x 8192 bf16 = 48 KB at
// B = 3) and there are 160 of them per decode step, so latency is everything: NCCL through torch.distributed costs
// ~40 us per call inside the CUDA graph, more than the weight streaming of the step.  Here every rank's partial lives
// in a buffer that all peers have mapped (CUDA IPC); one kernel per rank
//   1. tells every peer "my partial k is complete" (one remote flag store per peer and CTA),
//   2. waits until every peer has said the same,
//   3. reads all partials straight from peer memory, adds them in rank order (fp32; identical bits on every rank),
//      rounds to bf16 (the all-reduce result) and adds the residual (bf16), as HF does in two steps.
// No second barrier: partials alternate between two slots, and a rank can only start call k+2 after every peer has
// announced call k+1, i.e. finished reading call k (see DESIGN.md "Tensor parallel").
#include "common.cuh"
#include <cuda.h>
#include <string.h>

namespace ocrb {

typedef __nv_bfloat16 bf16;
constexpr int AR_MAX_WORLD = 8;
constexpr int AR_CTAS = 16;
constexpr int AR_THREADS = 256;

struct ArPeers {
  const bf16 *data[AR_MAX_WORLD];     // peer r's partial slot for this call (device pointers valid on this GPU)
  int *flags[AR_MAX_WORLD];           // peer r's flag array [AR_CTAS][AR_MAX_WORLD]
};

__global__ void __launch_bounds__(AR_THREADS)
allreduce_residual_kernel(ArPeers peers, int world, int rank, int *__restrict__ seq, bf16 *__restrict__ x, long long ldx,
                          int rows, int dim, long long ld_part) {
  const int cta = blockIdx.x, tid = threadIdx.x;
  // Launched as a normal kernel: a PDL launch (next GEMM prefetching during the exchange) measured no faster at TP-2
  // and made the 2-rank tiny-config test hang (profiles/r01_notes.md).
  __shared__ int s_k;
  if (tid == 0) s_k = seq[cta] + 1;
  __syncthreads();
  const int k = s_k;
  if (tid < world) {
    // the partial was written by the GEMM kernel that precedes this launch in the stream: complete and visible
    int *dst = peers.flags[tid] + cta * AR_MAX_WORLD + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(k) : "memory");
  }
  if (tid < world) {
    const int *src = peers.flags[rank] + cta * AR_MAX_WORLD + tid;
    int v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (v < k && clock64() - t0 > 60000000000LL) {
        printf("ocrb all-reduce: rank %d never saw rank %d announce call %d (flag %d)\n", rank, tid, k, v);
        __trap();
      }
    } while (v < k);
  }
  __syncthreads();
  const int vec_per_row = dim >> 3;
  const long long total = (long long)rows * vec_per_row;
  for (long long i = (long long)cta * AR_THREADS + tid; i < total; i += (long long)AR_CTAS * AR_THREADS) {
    const int r = (int)(i / vec_per_row), v = (int)(i % vec_per_row);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    uint4 raw[AR_MAX_WORLD];
#pragma unroll
    for (int p = 0; p < AR_MAX_WORLD; ++p)
      if (p < world) {
        const bf16 *src = peers.data[p] + (size_t)r * ld_part + v * 8;
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw[p].x), "=r"(raw[p].y), "=r"(raw[p].z), "=r"(raw[p].w)
                     : "l"(src));
      }
#pragma unroll
    for (int p = 0; p < AR_MAX_WORLD; ++p)
      if (p < world) {
        const bf16 *pe = reinterpret_cast<const bf16 *>(&raw[p]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(pe[e]);
      }
    uint4 xa = *reinterpret_cast<const uint4 *>(x + (size_t)r * ldx + v * 8);
    bf16 *xe = reinterpret_cast<bf16 *>(&xa);
#pragma unroll
    for (int e = 0; e < 8; ++e) xe[e] = __float2bfloat16_rn(__bfloat162float(xe[e]) + bf16_round(acc[e]));
    *reinterpret_cast<uint4 *>(x + (size_t)r * ldx + v * 8) = xa;
  }
  __syncthreads();
  if (tid == 0) seq[cta] = k;
}

}  // namespace ocrb

using namespace ocrb;

/* IPC plumbing: handle + offset of a device pointer inside its cudaMalloc allocation (torch tensors are
 * sub-allocations of large blocks). */
extern "C" int ocrb_comm_ipc_handle(const void *ptr, void *handle64, int64_t *offset) {
  OCRB_REQUIRE(ptr && handle64 && offset, "comm_ipc_handle: null pointer");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
  static RangeFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
      set_error("comm_ipc_handle: cuMemGetAddressRange unavailable");
      return OCRB_ECUDA;
    }
    fn = (RangeFn)p;
  }
  if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) {
    set_error("comm_ipc_handle: cuMemGetAddressRange failed");
    return OCRB_ECUDA;
  }
  cudaIpcMemHandle_t h;
  OCRB_CUDA(cudaIpcGetMemHandle(&h, (void *)base));
  memcpy(handle64, &h, sizeof(h));
  *offset = (int64_t)((CUdeviceptr)ptr - base);
  return OCRB_OK;
}

extern "C" int ocrb_comm_ipc_open(const void *handle64, int64_t offset, void **ptr) {
  OCRB_REQUIRE(handle64 && ptr, "comm_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void *base = nullptr;
  OCRB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = (char *)base + offset;
  return OCRB_OK;
}

/* x[rows, dim] += bf16(sum over ranks of partial_r[rows, dim]).  data_ptrs / flag_ptrs: host arrays of `world` device
 * pointers valid on THIS GPU (own buffer at index `rank`): each rank's partial for this call and each rank's flag array
 * (int32 [16][8], zero-initialised once).  seq: int32[16] private to this rank, zero-initialised once.  Every rank must
 * issue the same sequence of calls.  rows * dim must be a multiple of 8; dim % 8 == 0. */
extern "C" int ocrb_allreduce_residual_bf16(void *x, int64_t ldx, const void *const *data_ptrs, void *const *flag_ptrs,
                                            int32_t world, int32_t rank, int32_t *seq, int32_t rows, int32_t dim,
                                            int64_t ld_part, void *stream) {
  OCRB_REQUIRE(x && data_ptrs && flag_ptrs && seq, "allreduce_residual_bf16: null pointer");
  OCRB_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce_residual_bf16: bad world/rank");
  OCRB_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && ldx % 8 == 0 && ld_part % 8 == 0, "allreduce_residual_bf16: bad sizes");
  ArPeers peers;
  for (int r = 0; r < AR_MAX_WORLD; ++r) {
    peers.data[r] = (const bf16 *)(r < world ? data_ptrs[r] : nullptr);
    peers.flags[r] = (int *)(r < world ? flag_ptrs[r] : nullptr);
  }
  allreduce_residual_kernel<<<AR_CTAS, AR_THREADS, 0, (cudaStream_t)stream>>>(peers, world, rank, seq, (bf16 *)x, ldx, rows,
                                                                              dim, ld_part);
  return check_launch("allreduce_residual_kernel");
}
