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

// ───────────── vocab-split lm_head: greedy token by a (max, lowest index) exchange instead of gathering the logits ─────────────
// HF's plan leaves lm_head "colwise_rep" (every rank ends up with all logits).  A greedy step only needs the arg max: each
// rank reduces its own [B, V / world] slice to one (value, global index) pair per sequence, publishes it in a peer-mapped
// slot, and reads the other ranks' pairs -- 8 bytes per rank and sequence instead of an all-gather of B x V bf16 followed
// by an arg max over the full vocabulary.  Equal values resolve to the LOWEST global index, which is what the first-index
// arg max over the concatenated logits returns, so the tokens equal the gather route's bit for bit and are identical on
// every rank.  Same flag protocol and two-slot argument as the all-reduce above (one CTA per sequence, counters in seq).
struct AmPeers {
  const unsigned long long *pairs[AR_MAX_WORLD];   // peer r's pair slots [2][max_rows]
  int *flags[AR_MAX_WORLD];                        // peer r's flag array [max_rows][AR_MAX_WORLD]
};

__global__ void __launch_bounds__(512)
tp_argmax_step_kernel(AmPeers peers, int world, int rank, int *__restrict__ seq, int max_rows,
                      const bf16 *__restrict__ logits, long long ldl, int vl, int eos, int pad, int max_new,
                      int32_t *__restrict__ out_tokens, int32_t *__restrict__ next_ids, int32_t *__restrict__ finished,
                      int32_t *__restrict__ ctx_len, const int32_t *__restrict__ step, int advance_ctx) {
  __shared__ float s_val[16];
  __shared__ int s_idx[16];
  __shared__ unsigned long long s_pair[AR_MAX_WORLD];
  __shared__ int s_k;
  const int b = blockIdx.x, tid = threadIdx.x;
  const bf16 *row = logits + (size_t)b * ldl;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  const int nvec = vl >> 3;
  for (int v = tid; v < nvec; v += 512) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(row + v * 8);
    const bf16 *e = reinterpret_cast<const bf16 *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = __bfloat162float(e[k]);
      if (f > best) { best = f; best_i = v * 8 + k; }  // ascending index within a thread: first max kept
    }
  }
  for (int i = (nvec << 3) + tid; i < vl; i += 512) {
    const float f = __bfloat162float(row[i]);
    if (f > best || (f == best && i < best_i)) { best = f; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = best_i; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < 16; ++k)
      if (s_val[k] > best || (s_val[k] == best && s_idx[k] < best_i)) { best = s_val[k]; best_i = s_idx[k]; }
    const int k = seq[b] + 1;
    s_k = k;
    // a slice without any finite-comparable value keeps best_i = 0x7fffffff, which loses every tie
    const unsigned gi = best_i == 0x7fffffff ? 0x7fffffffu : (unsigned)(rank * vl + best_i);
    const unsigned long long pr = ((unsigned long long)__float_as_uint(best) << 32) | gi;
    unsigned long long *mine = const_cast<unsigned long long *>(peers.pairs[rank]) + (size_t)(k & 1) * max_rows + b;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(mine), "l"(pr) : "memory");
    __threadfence_system();
  }
  __syncthreads();
  const int k = s_k;
  if (tid < world) {
    int *dst = peers.flags[tid] + b * AR_MAX_WORLD + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(k) : "memory");
    const int *src = peers.flags[rank] + b * AR_MAX_WORLD + tid;
    int v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (v < k && clock64() - t0 > 60000000000LL) {
        printf("ocrb tp arg max: rank %d never saw rank %d announce step %d (flag %d)\n", rank, tid, k, v);
        __trap();
      }
    } while (v < k);
    const unsigned long long *pp = peers.pairs[tid] + (size_t)(k & 1) * max_rows + b;
    unsigned long long pr;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(pr) : "l"(pp) : "memory");
    s_pair[tid] = pr;
  }
  __syncthreads();
  if (tid == 0) {
    float bv = __uint_as_float((unsigned)(s_pair[0] >> 32));
    unsigned bi = (unsigned)s_pair[0];
    for (int p = 1; p < world; ++p) {
      const float v = __uint_as_float((unsigned)(s_pair[p] >> 32));
      const unsigned i = (unsigned)s_pair[p];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    int tok = (int)bi;
    const int st = step[0];
    if (finished[b]) tok = pad;
    if (st < max_new) out_tokens[(size_t)b * max_new + st] = tok;
    if (tok == eos) finished[b] = 1;
    next_ids[b] = tok;
    if (advance_ctx) ctx_len[b] += 1;
    seq[b] = k;
  }
}

__global__ void tp_step_increment_kernel(int32_t *step) { step[0] += 1; }

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

/* Greedy step of a vocab-split lm_head without gathering the logits: every rank passes its [B, vl] slice (vl = V / world,
 * rank r holds vocabulary rows [r*vl, (r+1)*vl)), the kernel exchanges one (max, global index) pair per sequence through
 * peer memory and then does the bookkeeping of ocrb_argmax_step with the winning token (ties -> lowest index, identical
 * on all ranks).  pair_ptrs / flag_ptrs: host arrays of `world` device pointers valid on THIS GPU: each rank's pair slots
 * (uint64 [2][max_rows]) and flag array (int32 [max_rows][8]), zero-initialised once; seq: int32[max_rows] private to
 * this rank, zero-initialised once.  B <= max_rows.  Every rank must issue the same sequence of calls. */
extern "C" int ocrb_tp_argmax_step(const void *logits_local, int64_t ldl, int32_t B, int32_t vl, const void *const *pair_ptrs,
                                   void *const *flag_ptrs, int32_t world, int32_t rank, int32_t *seq, int32_t max_rows,
                                   int32_t eos, int32_t pad, int32_t max_new, int32_t *out_tokens, int32_t *next_ids,
                                   int32_t *finished, int32_t *ctx_len, int32_t *step, int32_t advance_ctx, void *stream) {
  OCRB_REQUIRE(logits_local && pair_ptrs && flag_ptrs && seq && out_tokens && next_ids && finished && ctx_len && step,
               "tp_argmax_step: null pointer");
  OCRB_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "tp_argmax_step: bad world/rank");
  OCRB_REQUIRE(B > 0 && B <= max_rows && vl > 0 && ldl % 8 == 0 && max_new > 0 && (long long)vl * world < 0x7fffffffLL,
               "tp_argmax_step: bad sizes");
  AmPeers peers;
  for (int r = 0; r < AR_MAX_WORLD; ++r) {
    peers.pairs[r] = (const unsigned long long *)(r < world ? pair_ptrs[r] : nullptr);
    peers.flags[r] = (int *)(r < world ? flag_ptrs[r] : nullptr);
  }
  tp_argmax_step_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(peers, world, rank, seq, max_rows, (const bf16 *)logits_local, ldl,
                                                             vl, eos, pad, max_new, out_tokens, next_ids, finished, ctx_len,
                                                             step, advance_ctx);
  int rc = check_launch("tp_argmax_step_kernel");
  if (rc) return rc;
  tp_step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  return check_launch("tp_step_increment_kernel");
}
