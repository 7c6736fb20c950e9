// Batched greedy decode hot path (HF generation loop, utils.py:2743-2806, one token per sequence per step): paged
// split-KV attention over the cache.  (The weight-streaming linears of the step are in skinny.cu.)
//
// HBM-bound: every cached K and V byte of every sequence is read exactly once per step (B * ctx * 57 344 B for the 7B
// config over 28 layers).  Flash-decoding layout:
//   * grid = (key ranges, KV heads, sequences); a CTA owns `chunk` consecutive keys of one (sequence, KV head);
//   * the G query heads that share the KV head are the 16 rows of mma.sync.m16n8k16 tiles (rows >= G are zero);
//   * every WARP is an independent online-softmax worker: it takes the 16-key tiles t = warp, warp + 4, ... of the CTA's
//     range, streams their K/V rows with cp.async into its private 3-stage shared-memory ring (issued BEFORE
//     griddepcontrol.wait: the cache was written by earlier steps, not by the preceding qkv GEMM), keeps its running
//     (max, sum, O[16 x hd]) in registers and never meets another warp inside the loop (no __syncthreads);
//   * the four warps merge once through shared memory and the CTA stores one fp32 partial (max, sum, o[hd]) per head;
//     decode_attn_combine_kernel folds the partials of a sequence's key ranges.
// mRoPE of the new token's q / k (bf16 arithmetic, as HF) and the append of its k, v to the cache are fused: the warp
// whose tile holds position ctx builds that row in shared memory and writes it to the cache.
// The arithmetic of a sequence depends only on (its context length, chunk) -- never on B or on the other sequences.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ocrb {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float rope_elem_bf16(const bf16 *vec, int i, int hd, const bf16 *c, const bf16 *s) {
  const int half = hd >> 1;
  const float x = __bfloat162float(vec[i]);
  const float other = (i < half) ? -__bfloat162float(vec[i + half]) : __bfloat162float(vec[i - half]);
  const float t = bf16_round(x * __bfloat162float(c[i]));
  const float u = bf16_round(other * __bfloat162float(s[i]));
  return bf16_round(t + u);
}

constexpr int DA_TILE = 16;            // keys per warp tile (one k-step of the P.V product)
constexpr int DA_WARPS = 4;
constexpr int DA_STAGES = 3;           // per-warp cp.async ring depth
constexpr int DA_MAXG = 16;            // query heads per KV head (rows of the m16 tile)

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

template <int HD>
struct DaSmem {
  static constexpr int PITCH = HD + 8;                       // bf16 per shared row (conflict-free ldmatrix)
  static constexpr int TILE_ELEMS = DA_TILE * PITCH;         // one K (or V) tile
  static constexpr int STAGE_ELEMS = 2 * TILE_ELEMS;         // K then V
  static constexpr int WARP_ELEMS = DA_STAGES * STAGE_ELEMS;
  static constexpr size_t BYTES = (size_t)(16 * PITCH + DA_WARPS * WARP_ELEMS) * sizeof(bf16) + DA_WARPS * 32 * sizeof(float);
};

template <int HD>
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attn_kernel(const bf16 *__restrict__ qkv, long long ldqkv, bf16 *__restrict__ k_cache, bf16 *__restrict__ v_cache,
                   const int32_t *__restrict__ block_table, int max_pages, const int32_t *__restrict__ ctx_len,
                   int page_size, int n_q, int n_kv, const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT,
                   float scale, float *__restrict__ split_ws, int n_splits, int chunk) {
  using SM = DaSmem<HD>;
  constexpr int PITCH = SM::PITCH;
  constexpr int NJ = HD / 8;                                  // 8-wide output column tiles of O
  extern __shared__ __align__(16) uint8_t da_smem[];
  bf16 *sQ = reinterpret_cast<bf16 *>(da_smem);
  bf16 *ring = sQ + 16 * PITCH;
  float *s_ml = reinterpret_cast<float *>(ring + DA_WARPS * SM::WARP_ELEMS);   // [warp][16 rows][m, l]
  const int G = n_q / n_kv;
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) pdl_launch_dependents();
  const int ctx = ctx_len[b];            // written by the previous step's argmax kernel: complete long before this launch
  const int total = ctx + 1;
  const int k0 = split * chunk;
  const int k1 = min(total, k0 + chunk);
  float *ws = split_ws + (((size_t)b * n_q + (size_t)kvh * G) * n_splits + split) * (HD + 2);
  const size_t ws_head = (size_t)n_splits * (HD + 2);
  if (k0 >= k1) {
    pdl_wait();                          // split_ws may still be read by the previous layer's combine kernel
    for (int g = tid; g < G; g += DA_WARPS * 32) { ws[g * ws_head] = -INFINITY; ws[g * ws_head + 1] = 0.f; }
    return;
  }
  const int n_tiles = (k1 - k0 + DA_TILE - 1) / DA_TILE;     // tiles of this CTA; warp w takes w, w + 4, ...
  const int my_tiles = (n_tiles - warp + DA_WARPS - 1) / DA_WARPS;
  const size_t tok_stride = (size_t)n_kv * HD;
  const int32_t *bt = block_table + (size_t)b * max_pages;
  bf16 *wring = ring + warp * SM::WARP_ELEMS;

  // K / V rows of one tile: 16 rows x (HD / 8) 16-byte pieces each, spread over the 32 lanes.  Rows of tokens that are
  // not cached yet (the new token, written below) or beyond the context are zero-filled.
  auto issue_tile = [&](int i) {
    if (i < my_tiles) {
      const int key0 = k0 + (warp + i * DA_WARPS) * DA_TILE;
      bf16 *sk = wring + (i % DA_STAGES) * SM::STAGE_ELEMS, *sv = sk + SM::TILE_ELEMS;
      const size_t row0 = (size_t)bt[key0 / page_size] * page_size + key0 % page_size;   // page_size % 16 == 0: one page per tile
      constexpr int PIECES = HD / 8;                         // 16-byte pieces per row
#pragma unroll
      for (int j = 0; j < DA_TILE * PIECES / 32; ++j) {
        const int idx = lane + j * 32;
        const int r = idx / PIECES, c = idx % PIECES;
        bf16 *dk = sk + r * PITCH + c * 8, *dv = sv + r * PITCH + c * 8;
        if (key0 + r < ctx) {
          const size_t off = (row0 + r) * tok_stride + (size_t)kvh * HD + c * 8;
          cp_async16(dk, k_cache + off);
          cp_async16(dv, v_cache + off);
        } else {
          *reinterpret_cast<uint4 *>(dk) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4 *>(dv) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    cp_async_commit();                   // one group per ring slot use, empty or not, so the wait counts stay uniform
  };
#pragma unroll
  for (int i = 0; i < DA_STAGES; ++i) issue_tile(i);

  pdl_wait();                            // qkv comes from the preceding skinny GEMM
  const bf16 *row = qkv + (size_t)b * ldqkv;
  const bf16 *c = cosT + (size_t)b * HD, *s = sinT + (size_t)b * HD;
  for (int i = tid; i < 16 * HD; i += DA_WARPS * 32) {
    const int g = i / HD, d = i % HD;
    float v = 0.f;
    if (g < G) v = rope_elem_bf16(row + (size_t)(kvh * G + g) * HD, d, HD, c, s);
    sQ[g * PITCH + d] = __float2bfloat16_rn(v);
  }
  __syncthreads();

  const int m = lane >> 3, l8 = lane & 7;
  const int r0 = lane >> 2, cq = (lane & 3) * 2;
  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) ldmatrix_x4(qf[kk], sQ + ((m & 1) * 8 + l8) * PITCH + kk * 16 + (m >> 1) * 8);
  float o[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float run_m[2] = {-INFINITY, -INFINITY}, run_l[2] = {0.f, 0.f};

  for (int i = 0; i < my_tiles; ++i) {
    cp_async_wait<DA_STAGES - 1>();
    __syncwarp();
    const int key0 = k0 + (warp + i * DA_WARPS) * DA_TILE;
    bf16 *sk = wring + (i % DA_STAGES) * SM::STAGE_ELEMS, *sv = sk + SM::TILE_ELEMS;
    if (ctx >= key0 && ctx < key0 + DA_TILE) {
      // this tile holds the new token: rope its key, place k / v in the tile and append them to the cache
      const bf16 *knew = row + (size_t)n_q * HD + (size_t)kvh * HD;
      const bf16 *vnew = row + (size_t)(n_q + n_kv) * HD + (size_t)kvh * HD;
      const size_t dst = ((size_t)bt[ctx / page_size] * page_size + ctx % page_size) * tok_stride + (size_t)kvh * HD;
      const int r = ctx - key0;
      for (int d = lane; d < HD; d += 32) {
        const bf16 kr = __float2bfloat16_rn(rope_elem_bf16(knew, d, HD, c, s));
        const bf16 vr = vnew[d];
        sk[r * PITCH + d] = kr;
        sv[r * PITCH + d] = vr;
        k_cache[dst + d] = kr;
        v_cache[dst + d] = vr;
      }
      __syncwarp();
    }
    // ---- S = Q.K^T for the 16 keys of the tile (two 8-key column tiles) ----
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t bb[4];
      ldmatrix_x4(bb, sk + ((m >> 1) * 8 + l8) * PITCH + kk * 16 + (m & 1) * 8);
      mma_bf16_16816(acc[0], qf[kk], bb[0], bb[1]);
      mma_bf16_16816(acc[1], qf[kk], bb[2], bb[3]);
    }
    // ---- online softmax: this thread holds rows r0 (e = 0, 1) and r0 + 8 (e = 2, 3), keys j * 8 + cq + (e & 1) ----
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = key0 + j * 8 + cq + (e & 1) < total;
        acc[j][e] = ok ? acc[j][e] * scale : -INFINITY;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], acc[j][e]);
      }
    float corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tmax[h] = fmaxf(tmax[h], __shfl_xor_sync(0xffffffffu, tmax[h], 1));
      tmax[h] = fmaxf(tmax[h], __shfl_xor_sync(0xffffffffu, tmax[h], 2));
      const float mn = fmaxf(run_m[h], tmax[h]);              // finite: every processed tile has a live key
      corr[h] = __expf(run_m[h] - mn);                        // exp(-inf) = 0 on the first tile
      run_m[h] = mn;
      run_l[h] *= corr[h];
    }
    uint32_t pa[4];
    {
      float pv[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[j][e] = __expf(acc[j][e] - run_m[e >> 1]);
          run_l[e >> 1] += pv[j][e];
        }
      pa[0] = pack_bf16(pv[0][0], pv[0][1]);
      pa[1] = pack_bf16(pv[0][2], pv[0][3]);
      pa[2] = pack_bf16(pv[1][0], pv[1][1]);
      pa[3] = pack_bf16(pv[1][2], pv[1][3]);
    }
    // ---- O = O * corr + P.V ----
#pragma unroll
    for (int jj = 0; jj < HD / 16; ++jj) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, sv + ((m & 1) * 8 + l8) * PITCH + jj * 16 + (m >> 1) * 8);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float (&oo)[4] = o[jj * 2 + t];
        oo[0] *= corr[0]; oo[1] *= corr[0]; oo[2] *= corr[1]; oo[3] *= corr[1];
      }
      mma_bf16_16816(o[jj * 2], pa, bb[0], bb[1]);
      mma_bf16_16816(o[jj * 2 + 1], pa, bb[2], bb[3]);
    }
    __syncwarp();                        // every lane is done with this ring slot
    issue_tile(i + DA_STAGES);
  }
  cp_async_wait<0>();
  // ---- merge the four warps (fixed order: warp 0..3), one fp32 partial per head ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    run_l[h] += __shfl_xor_sync(0xffffffffu, run_l[h], 1);
    run_l[h] += __shfl_xor_sync(0xffffffffu, run_l[h], 2);
  }
  __syncthreads();                       // all warps left their rings: the space is reused for the fp32 O tiles
  float *sO = reinterpret_cast<float *>(ring);               // [warp][16][HD + 4]
  constexpr int OP = HD + 4;
  static_assert((size_t)DA_WARPS * 16 * OP * sizeof(float) <= (size_t)DA_WARPS * SM::WARP_ELEMS * sizeof(bf16), "merge buffer");
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    *reinterpret_cast<float2 *>(sO + ((size_t)warp * 16 + r0) * OP + j * 8 + cq) = make_float2(o[j][0], o[j][1]);
    *reinterpret_cast<float2 *>(sO + ((size_t)warp * 16 + r0 + 8) * OP + j * 8 + cq) = make_float2(o[j][2], o[j][3]);
  }
  if ((lane & 3) == 0) {
    s_ml[(warp * 16 + r0) * 2] = run_m[0];
    s_ml[(warp * 16 + r0) * 2 + 1] = run_l[0];
    s_ml[(warp * 16 + r0 + 8) * 2] = run_m[1];
    s_ml[(warp * 16 + r0 + 8) * 2 + 1] = run_l[1];
  }
  __syncthreads();
  for (int idx = tid; idx < G * HD; idx += DA_WARPS * 32) {
    const int g = idx / HD, d = idx % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) M = fmaxf(M, s_ml[(w * 16 + g) * 2]);
    float L = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float mw = s_ml[(w * 16 + g) * 2];
      const float f = (mw == -INFINITY) ? 0.f : __expf(mw - M);
      L += s_ml[(w * 16 + g) * 2 + 1] * f;
      a += sO[((size_t)w * 16 + g) * OP + d] * f;
    }
    ws[g * ws_head + 2 + d] = a;
    if (d == 0) { ws[g * ws_head] = M; ws[g * ws_head + 1] = L; }
  }
}

// grid = (n_q, B), hd threads
__global__ void decode_attn_combine_kernel(const float *__restrict__ split_ws, int n_q, int n_splits, int hd,
                                           bf16 *__restrict__ out, long long ldo) {
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  pdl_wait();
  // Trigger only now: the o_proj GEMM behind this kernel then overlaps this (short) kernel, not the attention
  // kernel before it (DESIGN.md "PDL").
  if (d == 0) pdl_launch_dependents();
  const float *ws = split_ws + ((size_t)b * n_q + h) * n_splits * (hd + 2);
  float M = -INFINITY;
  for (int s = 0; s < n_splits; ++s) M = fmaxf(M, ws[(size_t)s * (hd + 2)]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float m = ws[(size_t)s * (hd + 2)];
    if (m == -INFINITY) continue;
    const float f = __expf(m - M);
    L += ws[(size_t)s * (hd + 2) + 1] * f;
    acc += ws[(size_t)s * (hd + 2) + 2 + d] * f;
  }
  out[(size_t)b * ldo + (size_t)h * hd + d] = __float2bfloat16_rn(acc / L);
}

template <int HD>
static int launch_decode_attn(const bf16 *qkv, long long ldqkv, bf16 *kc, bf16 *vc, const int32_t *bt, int max_pages,
                              const int32_t *ctx_len, int B, int page_size, int n_q, int n_kv, const bf16 *cosT,
                              const bf16 *sinT, float scale, float *split_ws, int n_splits, int chunk, cudaStream_t st) {
  constexpr size_t smem = DaSmem<HD>::BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(decode_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  OCRB_CUDA(launch_pdl_bit(2, decode_attn_kernel<HD>, dim3(n_splits, n_kv, B), dim3(DA_WARPS * 32), smem, st, qkv, ldqkv, kc, vc,
                           bt, max_pages, ctx_len, page_size, n_q, n_kv, cosT, sinT, scale, split_ws, n_splits, chunk));
  return check_launch("decode_attn_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_decode_attention(const void *qkv, int64_t ldqkv, void *k_cache, void *v_cache,
                                     const int32_t *block_table, int32_t max_pages, const int32_t *ctx_len, int32_t B,
                                     int32_t page_size, int32_t n_q, int32_t n_kv, int32_t hd, const void *cosT,
                                     const void *sinT, float scale, void *out, int64_t ldo, float *split_ws,
                                     int32_t n_splits, void *stream) {
  OCRB_REQUIRE(qkv && k_cache && v_cache && block_table && ctx_len && cosT && sinT && out && split_ws,
               "decode_attention: null pointer");
  OCRB_REQUIRE(B > 0 && n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= DA_MAXG && (hd == 64 || hd == 128) && n_splits > 0,
               "decode_attention: needs n_q / n_kv <= 16 query heads per KV head and hd of 64 or 128");
  OCRB_REQUIRE(page_size > 0 && page_size % DA_TILE == 0, "decode_attention: page_size must be a multiple of 16");
  OCRB_REQUIRE(B <= 65535 && n_kv <= 65535, "decode_attention: grid too large");
  const int max_ctx = max_pages * page_size;
  const int chunk = cdiv(cdiv(max_ctx, n_splits), DA_TILE) * DA_TILE;     // keys per CTA, whole tiles
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (hd == 128)
    rc = launch_decode_attn<128>((const bf16 *)qkv, (long long)ldqkv, (bf16 *)k_cache, (bf16 *)v_cache, block_table,
                                 (int)max_pages, ctx_len, (int)B, (int)page_size, (int)n_q, (int)n_kv, (const bf16 *)cosT,
                                 (const bf16 *)sinT, scale, split_ws, (int)n_splits, chunk, st);
  else
    rc = launch_decode_attn<64>((const bf16 *)qkv, (long long)ldqkv, (bf16 *)k_cache, (bf16 *)v_cache, block_table,
                                (int)max_pages, ctx_len, (int)B, (int)page_size, (int)n_q, (int)n_kv, (const bf16 *)cosT,
                                (const bf16 *)sinT, scale, split_ws, (int)n_splits, chunk, st);
  if (rc) return rc;
  OCRB_CUDA(launch_pdl_bit(4, decode_attn_combine_kernel, dim3(n_q, B), dim3(hd), 0, st, (const float *)split_ws, (int)n_q,
                           (int)n_splits, (int)hd, (bf16 *)out, (long long)ldo));
  return check_launch("decode_attn_combine_kernel");
}
