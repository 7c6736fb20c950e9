// Variable-length flash attention for the vision tower (windowed / full, non-causal, hd 80;
// HF modeling_qwen2_5_vl.py:231-283) and for decoder prefill (causal GQA, hd 128; :718-760).
// bf16 operands, fp32 scores / softmax / accumulation, online softmax, one pass over K/V.
// v1 datapath: mma.sync.m16n8k16 with ldmatrix-fed fragments (attention is ~3 % of the read's
// FLOPs; the dense GEMMs are on tcgen05 -- see gemm_tcgen05.cu).  A tcgen05 version is listed
// in DESIGN.md as the follow-up.
#include "common.cuh"
#include <math.h>

namespace ocrb {

typedef __nv_bfloat16 bf16;

constexpr int FA_BM = 64, FA_BN = 64, FA_THREADS = 128;

__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// Copy a [rows<=64][HD] tile (global row stride ld) into padded shared memory, zero-filling rows >= valid.
template <int HD>
__device__ __forceinline__ void load_tile(bf16 *dst, const bf16 *src, long long ld, int valid_rows) {
  constexpr int LDS = HD + 8;
  constexpr int CH = HD / 8;
  for (int i = threadIdx.x; i < 64 * CH; i += FA_THREADS) {
    const int r = i / CH, c = i - r * CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid_rows) v = *reinterpret_cast<const uint4 *>(src + (size_t)r * ld + c * 8);
    *reinterpret_cast<uint4 *>(dst + r * LDS + c * 8) = v;
  }
}

template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
flash_varlen_kernel(const bf16 *__restrict__ Q, long long ldq, const bf16 *__restrict__ K, long long ldk,
                    const bf16 *__restrict__ V, long long ldv, bf16 *__restrict__ O, long long ldo,
                    const int32_t *__restrict__ cu_seqlens, int n_q, int n_kv, float scale_log2, int causal) {
  constexpr int LDS = HD + 8;
  constexpr int KSTEPS = HD / 16;  // k-steps of QK^T
  constexpr int DT = HD / 8;       // n-tiles of the output
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16 *sQ = reinterpret_cast<bf16 *>(fa_smem);
  bf16 *sK = sQ + 64 * LDS;
  bf16 *sV = sK + 64 * LDS;

  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int s0 = cu_seqlens[seq], s1 = cu_seqlens[seq + 1];
  const int len = s1 - s0;
  const int q0 = qt * FA_BM;
  if (q0 >= len) return;
  const int kvh = head / (n_q / n_kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile<HD>(sQ, Q + (size_t)(s0 + q0) * ldq + (size_t)head * HD, ldq, min(64, len - q0));
  __syncthreads();
  uint32_t qf[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3],
                sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);

  float o[DT][4];
#pragma unroll
  for (int t = 0; t < DT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int row_in_tile = warp * 16 + (lane >> 2);  // and +8
  const int qrow0 = q0 + row_in_tile, qrow1 = qrow0 + 8;

  const int kv_end = causal ? min(len, q0 + FA_BM) : len;
  for (int k0 = 0; k0 < kv_end; k0 += FA_BN) {
    __syncthreads();  // previous tile fully consumed
    const int valid = min(64, len - k0);
    load_tile<HD>(sK, K + (size_t)(s0 + k0) * ldk + (size_t)kvh * HD, ldk, valid);
    load_tile<HD>(sV, V + (size_t)(s0 + k0) * ldv + (size_t)kvh * HD, ldv, valid);
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {  // pairs of key n-tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, sK + (tp * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[2 * tp], qf[ks], b0, b1);
        mma_bf16(s[2 * tp + 1], qf[ks], b2, b3);
      }
    }
    // scale (log2 domain) + mask
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + t * 8 + (lane & 3) * 2 + (e & 1);
        const int qrow = (e < 2) ? qrow0 : qrow1;
        float v = s[t][e] * scale_log2;
        if (key >= len || (causal && key > qrow)) v = -INFINITY;
        s[t][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
    float alpha[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      m_new[r] = fmaxf(m_run[r], tmax[r]);
      // rows past the end of the sequence see only -inf: keep them finite
      const float m_safe = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
      alpha[r] = exp2f(m_run[r] - m_safe);
      m_run[r] = m_new[r];
      m_new[r] = m_safe;
      l_run[r] *= alpha[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float p0 = exp2f(s[t][0] - m_new[0]), p1 = exp2f(s[t][1] - m_new[0]);
      const float p2 = exp2f(s[t][2] - m_new[1]), p3 = exp2f(s[t][3] - m_new[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int j = t >> 1;
      if ((t & 1) == 0) { pf[j][0] = pack_bf16(p0, p1); pf[j][1] = pack_bf16(p2, p3); }
      else              { pf[j][2] = pack_bf16(p0, p1); pf[j][3] = pack_bf16(p2, p3); }
    }
#pragma unroll
    for (int t = 0; t < DT; ++t) {
      o[t][0] *= alpha[0]; o[t][1] *= alpha[0];
      o[t][2] *= alpha[1]; o[t][3] *= alpha[1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {       // k16 steps over the 64 keys
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {  // pairs of d n-tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, sV + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8);
        mma_bf16(o[2 * dp], pf[j], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[j], b2, b3);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  bf16 *orow0 = O + (size_t)(s0 + qrow0) * ldo + (size_t)head * HD;
  bf16 *orow1 = O + (size_t)(s0 + qrow1) * ldo + (size_t)head * HD;
#pragma unroll
  for (int t = 0; t < DT; ++t) {
    const int d = t * 8 + (lane & 3) * 2;
    if (qrow0 < len) *reinterpret_cast<__nv_bfloat162 *>(orow0 + d) = __floats2bfloat162_rn(o[t][0] * inv0, o[t][1] * inv0);
    if (qrow1 < len) *reinterpret_cast<__nv_bfloat162 *>(orow1 + d) = __floats2bfloat162_rn(o[t][2] * inv1, o[t][3] * inv1);
  }
}

// ───────────── windowed vision attention: every sequence is ONE tile (<= 64 tokens, non-causal) ─────────────
// The 28 windowed blocks of the vision tower (HF modeling_qwen2_5_vl.py:231-283 with cu_window_seqlens) are 3 360
// independent (window, head) problems of 64 x 64 x 80 per 3 pages: 123 MB of q / k / v / o per block, ~20 us of HBM time.
// One CTA per problem (flash_varlen_kernel) spends its life waiting for its own loads: 98 us per block.  Here the CTAs
// are persistent and double-buffered: the cp.async loads of problem i+1 are in flight while problem i is computed.
// Same arithmetic, in the same order, as flash_varlen_kernel on a single tile (bit-equal results).
__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;                       // 0: the 16 bytes are zero-filled, src is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(FA_THREADS, 3)
window_attn_kernel(const bf16 *__restrict__ Q, long long ldq, const bf16 *__restrict__ K, long long ldk,
                   const bf16 *__restrict__ V, long long ldv, bf16 *__restrict__ O, long long ldo,
                   const int32_t *__restrict__ cu_seqlens, int n_items, int n_heads, float scale_log2) {
  constexpr int LDS = HD + 8;
  constexpr int KSTEPS = HD / 16;
  constexpr int DT = HD / 8;
  constexpr int CH = HD / 8;
  constexpr int TILE = 64 * LDS;                        // elements of one staged operand
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16 *sbase = reinterpret_cast<bf16 *>(fa_smem);      // [2 buffers][Q | K | V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto issue = [&](int item, int buf) {
    if (item < n_items) {
      const int seq = item / n_heads, head = item - seq * n_heads;
      const int s0 = cu_seqlens[seq], len = cu_seqlens[seq + 1] - s0;
      bf16 *dst = sbase + (size_t)buf * 3 * TILE;
      for (int i = threadIdx.x; i < 3 * 64 * CH; i += FA_THREADS) {
        const int which = i / (64 * CH), j = i - which * (64 * CH);
        const int r = j / CH, c = j - r * CH;
        const bf16 *src = which == 0 ? Q : (which == 1 ? K : V);
        const long long ld = which == 0 ? ldq : (which == 1 ? ldk : ldv);
        const bool ok = r < len;
        cp_async16(dst + which * TILE + r * LDS + c * 8, src + (size_t)(s0 + (ok ? r : 0)) * ld + (size_t)head * HD + c * 8, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  issue(blockIdx.x, 0);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
    issue(item + gridDim.x, buf ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const int seq = item / n_heads, head = item - seq * n_heads;
    const int s0 = cu_seqlens[seq], len = cu_seqlens[seq + 1] - s0;
    const bf16 *sQ = sbase + (size_t)buf * 3 * TILE, *sK = sQ + TILE, *sV = sK + TILE;
    uint32_t qf[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    float o[DT][4];
#pragma unroll
    for (int t = 0; t < DT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
    const int qrow0 = warp * 16 + (lane >> 2), qrow1 = qrow0 + 8;
    float s[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, sK + (tp * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[2 * tp], qf[ks], b0, b1);
        mma_bf16(s[2 * tp + 1], qf[ks], b2, b3);
      }
    }
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = t * 8 + (lane & 3) * 2 + (e & 1);
        float v = s[t][e] * scale_log2;
        if (key >= len) v = -INFINITY;
        s[t][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
    float m_new[2], l_run[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      m_new[r] = (tmax[r] == -INFINITY) ? 0.f : tmax[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float p0 = exp2f(s[t][0] - m_new[0]), p1 = exp2f(s[t][1] - m_new[0]);
      const float p2 = exp2f(s[t][2] - m_new[1]), p3 = exp2f(s[t][3] - m_new[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int j = t >> 1;
      if ((t & 1) == 0) { pf[j][0] = pack_bf16(p0, p1); pf[j][1] = pack_bf16(p2, p3); }
      else              { pf[j][2] = pack_bf16(p0, p1); pf[j][3] = pack_bf16(p2, p3); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, sV + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8);
        mma_bf16(o[2 * dp], pf[j], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[j], b2, b3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    bf16 *orow0 = O + (size_t)(s0 + qrow0) * ldo + (size_t)head * HD;
    bf16 *orow1 = O + (size_t)(s0 + qrow1) * ldo + (size_t)head * HD;
#pragma unroll
    for (int t = 0; t < DT; ++t) {
      const int d = t * 8 + (lane & 3) * 2;
      if (qrow0 < len) *reinterpret_cast<__nv_bfloat162 *>(orow0 + d) = __floats2bfloat162_rn(o[t][0] * inv0, o[t][1] * inv0);
      if (qrow1 < len) *reinterpret_cast<__nv_bfloat162 *>(orow1 + d) = __floats2bfloat162_rn(o[t][2] * inv1, o[t][3] * inv1);
    }
    __syncthreads();                                    // this buffer is refilled by the next iteration's loads
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int HD>
static int launch_window(const bf16 *q, long long ldq, const bf16 *k, long long ldk, const bf16 *v, long long ldv, bf16 *out,
                         long long ldo, const int32_t *cu, int n_seq, int n_heads, float scale, cudaStream_t st) {
  const size_t smem = (size_t)2 * 3 * 64 * (HD + 8) * sizeof(bf16);
  static bool attr_set = false;
  static int n_sm = 0;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(window_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
    attr_set = true;
  }
  const long long items = (long long)n_seq * n_heads;
  const int grid = (int)(items < 3LL * n_sm ? items : 3LL * n_sm);
  window_attn_kernel<HD><<<grid, FA_THREADS, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, cu, (int)items, n_heads,
                                                          scale * 1.4426950408889634f);
  return check_launch("window_attn_kernel");
}

template <int HD>
static int launch_flash(const bf16 *q, long long ldq, const bf16 *k, long long ldk, const bf16 *v, long long ldv, bf16 *out,
                        long long ldo, const int32_t *cu, int n_seq, int max_seqlen, int n_q, int n_kv, float scale,
                        int causal, cudaStream_t st) {
  const size_t smem = (size_t)3 * 64 * (HD + 8) * sizeof(bf16);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    OCRB_CUDA(cudaFuncSetAttribute(flash_varlen_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  flash_varlen_kernel<HD><<<dim3(cdiv(max_seqlen, FA_BM), n_q, n_seq), FA_THREADS, smem, st>>>(
      q, ldq, k, ldk, v, ldv, out, ldo, cu, n_q, n_kv, scale_log2, causal);
  return check_launch("flash_varlen_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_attention_varlen(const void *q, int64_t ldq, const void *k, int64_t ldk, const void *v, int64_t ldv,
                                     void *out, int64_t ldo, const int32_t *cu_seqlens, int32_t n_seq, int32_t max_seqlen,
                                     int32_t n_q, int32_t n_kv, int32_t hd, float scale, int32_t causal, void *stream) {
  OCRB_REQUIRE(q && k && v && out && cu_seqlens, "attention_varlen: null pointer");
  OCRB_REQUIRE(n_seq > 0 && max_seqlen > 0 && n_kv > 0 && n_q % n_kv == 0, "attention_varlen: bad sizes");
  OCRB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention_varlen: strides must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (!causal && max_seqlen <= 64 && n_q == n_kv && (long long)n_seq * n_q < (1ll << 30)) {
    // windowed vision blocks: one tile per sequence -> persistent double-buffered workers
    if (hd == 80)
      return launch_window<80>((const bf16 *)q, ldq, (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)out, ldo, cu_seqlens,
                               n_seq, n_q, scale, st);
    if (hd == 64)
      return launch_window<64>((const bf16 *)q, ldq, (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)out, ldo, cu_seqlens,
                               n_seq, n_q, scale, st);
  }
  if (hd == 80)
    return launch_flash<80>((const bf16 *)q, ldq, (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)out, ldo, cu_seqlens,
                            n_seq, max_seqlen, n_q, n_kv, scale, causal, st);
  if (hd == 128)
    return launch_flash<128>((const bf16 *)q, ldq, (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)out, ldo, cu_seqlens,
                             n_seq, max_seqlen, n_q, n_kv, scale, causal, st);
  if (hd == 64)
    return launch_flash<64>((const bf16 *)q, ldq, (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)out, ldo, cu_seqlens,
                            n_seq, max_seqlen, n_q, n_kv, scale, causal, st);
  set_error("attention_varlen: head_dim %d not supported (64, 80, 128)", hd);
  return OCRB_EUNSUPPORTED;
}
