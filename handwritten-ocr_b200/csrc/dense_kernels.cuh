// Kernels of the small bandwidth-bound ops of the VLM (HF modeling_qwen2_5_vl.py): RMSNorm, vision RoPE, text mRoPE, row
// gathers, paged-KV prefill write, greedy argmax + step bookkeeping, residual add.  All mirror HF's rounding points (fp32
// math, bf16 stores where HF materialises bf16 tensors).  Definitions only -- no launches -- so that tests/emu can compile this
// file for the host (OCRB_EMU) and run the kernels thread by thread under the host sanitizers and against their twins.
#pragma once
#ifndef OCRB_EMU
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#endif
#include <math.h>
#include <stdint.h>

namespace ocrb {

#ifdef OCRB_EMU
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
#endif

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ───────────── RMSNorm: one CTA per row ─────────────
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const bf16 *__restrict__ x, long long ldx, const bf16 *__restrict__ w, bf16 *__restrict__ y,
               long long ldy, int dim, float eps) {
  __shared__ float s_part[8];
  const bf16 *xr = x + (size_t)blockIdx.x * ldx;
  bf16 *yr = y + (size_t)blockIdx.x * ldy;
  float ss = 0.f;
  const int nvec = dim >> 3;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(xr + v * 8);
    const bf16 *e = reinterpret_cast<const bf16 *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = __bfloat162float(e[k]);
      ss = fmaf(f, f, ss);
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += s_part[k];
  const float rstd = rsqrtf(tot / (float)dim + eps);
  for (int v = threadIdx.x; v < nvec; v += 256) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(xr + v * 8);
    const uint4 wraw = *reinterpret_cast<const uint4 *>(w + v * 8);
    const bf16 *e = reinterpret_cast<const bf16 *>(&raw);
    const bf16 *we = reinterpret_cast<const bf16 *>(&wraw);
    uint4 o;
    bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float n = bf16_round(__bfloat162float(e[k]) * rstd);
      oe[k] = __float2bfloat16_rn(__bfloat162float(we[k]) * n);
    }
    *reinterpret_cast<uint4 *>(yr + v * 8) = o;
  }
}

// One warp per row, the row stays in registers between the statistics and the scaling (one pass over memory, no block
// barrier): dim <= 4096.  16-byte accesses; lanes stride the row so every request is a full 512-byte line group.
__global__ void __launch_bounds__(256)
rmsnorm_warp_kernel(const bf16 *__restrict__ x, long long ldx, const bf16 *__restrict__ w, bf16 *__restrict__ y,
                    long long ldy, int rows, int dim, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16 *xr = x + (size_t)row * ldx;
  bf16 *yr = y + (size_t)row * ldy;
  const int nvec = dim >> 3;
  uint4 raw[16];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int v = lane + i * 32;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (v < nvec) raw[i] = *reinterpret_cast<const uint4 *>(xr + v * 8);
    const bf16 *e = reinterpret_cast<const bf16 *>(&raw[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = __bfloat162float(e[k]);
      ss = fmaf(f, f, ss);
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / (float)dim + eps);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 wraw = __ldg(reinterpret_cast<const uint4 *>(w + v * 8));
      const bf16 *e = reinterpret_cast<const bf16 *>(&raw[i]);
      const bf16 *we = reinterpret_cast<const bf16 *>(&wraw);
      uint4 o;
      bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float n = bf16_round(__bfloat162float(e[k]) * rstd);
        oe[k] = __float2bfloat16_rn(__bfloat162float(we[k]) * n);
      }
      *reinterpret_cast<uint4 *>(yr + v * 8) = o;
    }
  }
}

// ───────────── vision RoPE (fp32 math, unfused like eager torch) ─────────────
// 16-byte version: a thread rotates 8 (x1, x2) pairs -- x1 from the first half of a head, x2 from the second half.
// Same per-element arithmetic as the scalar kernel below (which stays for head dims whose half is not a multiple of 8).
__global__ void __launch_bounds__(256)
rope_vision_vec_kernel(bf16 *__restrict__ qkv, int S, int heads, int hd, const float *__restrict__ cosT,
                       const float *__restrict__ sinT) {
  const int half = hd >> 1, nv = half >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = S * 2 * heads * nv;
  if (idx >= total) return;
  const int v = idx % nv;
  int r = idx / nv;
  const int h = r % heads;
  r /= heads;
  const int which = r & 1;
  const int s = r >> 1;
  bf16 *p = qkv + ((size_t)s * 3 + which) * heads * hd + (size_t)h * hd + v * 8;
  const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + half);
  const float *c = cosT + (size_t)s * hd + v * 8, *sn = sinT + (size_t)s * hd + v * 8;
  float c1[8], c2[8], s1[8], s2[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    *reinterpret_cast<float4 *>(c1 + 4 * q) = __ldg(reinterpret_cast<const float4 *>(c + 4 * q));
    *reinterpret_cast<float4 *>(c2 + 4 * q) = __ldg(reinterpret_cast<const float4 *>(c + half + 4 * q));
    *reinterpret_cast<float4 *>(s1 + 4 * q) = __ldg(reinterpret_cast<const float4 *>(sn + 4 * q));
    *reinterpret_cast<float4 *>(s2 + 4 * q) = __ldg(reinterpret_cast<const float4 *>(sn + half + 4 * q));
  }
  const bf16 *ae = reinterpret_cast<const bf16 *>(&a), *be = reinterpret_cast<const bf16 *>(&b);
  uint4 oa, ob;
  bf16 *oae = reinterpret_cast<bf16 *>(&oa), *obe = reinterpret_cast<bf16 *>(&ob);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float x1 = __bfloat162float(ae[k]), x2 = __bfloat162float(be[k]);
    oae[k] = __float2bfloat16_rn(__fadd_rn(__fmul_rn(x1, c1[k]), __fmul_rn(-x2, s1[k])));
    obe[k] = __float2bfloat16_rn(__fadd_rn(__fmul_rn(x2, c2[k]), __fmul_rn(x1, s2[k])));
  }
  *reinterpret_cast<uint4 *>(p) = oa;
  *reinterpret_cast<uint4 *>(p + half) = ob;
}


// qkv: [S, 3, heads, hd]; rotates q (slot 0) and k (slot 1) in place.
__global__ void __launch_bounds__(256)
rope_vision_kernel(bf16 *__restrict__ qkv, int S, int heads, int hd, const float *__restrict__ cosT,
                   const float *__restrict__ sinT) {
  const int half = hd >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * 2 * heads * half;
  if (idx >= total) return;
  const int i = (int)(idx % half);
  long long r = idx / half;
  const int h = (int)(r % heads);
  r /= heads;
  const int which = (int)(r % 2);
  const int s = (int)(r / 2);
  bf16 *p = qkv + ((size_t)s * 3 + which) * heads * hd + (size_t)h * hd;
  const float x1 = __bfloat162float(p[i]), x2 = __bfloat162float(p[i + half]);
  const float c1 = cosT[(size_t)s * hd + i], c2 = cosT[(size_t)s * hd + i + half];
  const float s1 = sinT[(size_t)s * hd + i], s2 = sinT[(size_t)s * hd + i + half];
  const float o1 = __fadd_rn(__fmul_rn(x1, c1), __fmul_rn(-x2, s1));
  const float o2 = __fadd_rn(__fmul_rn(x2, c2), __fmul_rn(x1, s2));
  p[i] = __float2bfloat16_rn(o1);
  p[i + half] = __float2bfloat16_rn(o2);
}

// ───────────── text mRoPE in bf16 arithmetic (every op rounds to bf16 as eager torch does) ─────────────
__device__ __forceinline__ void rope_bf16_pair(bf16 &a, bf16 &b, bf16 c1, bf16 s1, bf16 c2, bf16 s2) {
  const float x1 = __bfloat162float(a), x2 = __bfloat162float(b);
  const float t1 = bf16_round(x1 * __bfloat162float(c1));
  const float u1 = bf16_round(-x2 * __bfloat162float(s1));
  const float t2 = bf16_round(x2 * __bfloat162float(c2));
  const float u2 = bf16_round(x1 * __bfloat162float(s2));
  a = __float2bfloat16_rn(t1 + u1);
  b = __float2bfloat16_rn(t2 + u2);
}

// 16-byte version of the text mRoPE below (head-dim half a multiple of 8, 16-byte aligned rows): same arithmetic.
__global__ void __launch_bounds__(256)
rope_text_vec_kernel(bf16 *__restrict__ q, long long ldq, bf16 *__restrict__ k, long long ldk, int T, int n_q, int n_kv,
                     int hd, const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT) {
  const int half = hd >> 1, nv = half >> 3;
  const int heads = n_q + n_kv;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * heads * nv) return;
  const int v = idx % nv;
  const int r = idx / nv;
  const int h = r % heads, t = r / heads;
  bf16 *p = ((h < n_q) ? q + (size_t)t * ldq + (size_t)h * hd : k + (size_t)t * ldk + (size_t)(h - n_q) * hd) + v * 8;
  const bf16 *c = cosT + (size_t)t * hd + v * 8, *sn = sinT + (size_t)t * hd + v * 8;
  uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + half);
  const uint4 c1 = __ldg(reinterpret_cast<const uint4 *>(c)), c2 = __ldg(reinterpret_cast<const uint4 *>(c + half));
  const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(sn)), s2 = __ldg(reinterpret_cast<const uint4 *>(sn + half));
  bf16 *ae = reinterpret_cast<bf16 *>(&a), *be = reinterpret_cast<bf16 *>(&b);
  const bf16 *c1e = reinterpret_cast<const bf16 *>(&c1), *c2e = reinterpret_cast<const bf16 *>(&c2);
  const bf16 *s1e = reinterpret_cast<const bf16 *>(&s1), *s2e = reinterpret_cast<const bf16 *>(&s2);
#pragma unroll
  for (int e = 0; e < 8; ++e) rope_bf16_pair(ae[e], be[e], c1e[e], s1e[e], c2e[e], s2e[e]);
  *reinterpret_cast<uint4 *>(p) = a;
  *reinterpret_cast<uint4 *>(p + half) = b;
}

__global__ void __launch_bounds__(256)
rope_text_kernel(bf16 *__restrict__ q, long long ldq, bf16 *__restrict__ k, long long ldk, int T, int n_q, int n_kv,
                 int hd, const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT) {
  const int half = hd >> 1;
  const int heads = n_q + n_kv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)T * heads * half;
  if (idx >= total) return;
  const int i = (int)(idx % half);
  long long r = idx / half;
  const int h = (int)(r % heads);
  const int t = (int)(r / heads);
  bf16 *p = (h < n_q) ? q + (size_t)t * ldq + (size_t)h * hd : k + (size_t)t * ldk + (size_t)(h - n_q) * hd;
  const bf16 *c = cosT + (size_t)t * hd, *s = sinT + (size_t)t * hd;
  rope_bf16_pair(p[i], p[i + half], c[i], s[i], c[i + half], s[i + half]);
}

// cos/sin for one decode step: pos[b] = ctx_len[b] + rope_delta[b] (text tokens: t = h = w = pos).
__global__ void decode_rope_table_kernel(const int32_t *__restrict__ ctx_len, const int32_t *__restrict__ rope_delta,
                                         const float *__restrict__ inv_freq, int B, int hd, bf16 *__restrict__ cosT,
                                         bf16 *__restrict__ sinT) {
  const int half = hd >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const float pos = (float)(ctx_len[b] + rope_delta[b]);
  const float f = __fmul_rn(inv_freq[i], pos);
  const bf16 c = __float2bfloat16_rn(cosf(f)), s = __float2bfloat16_rn(sinf(f));
  cosT[(size_t)b * hd + i] = c;
  cosT[(size_t)b * hd + i + half] = c;
  sinT[(size_t)b * hd + i] = s;
  sinT[(size_t)b * hd + i + half] = s;
}

// ───────────── row gathers ─────────────
__global__ void __launch_bounds__(256)
rows_copy_kernel(const bf16 *__restrict__ src, long long lds, const int32_t *__restrict__ src_idx, bf16 *__restrict__ dst,
                 long long ldd, const int32_t *__restrict__ dst_idx, int n_rows, int nvec) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rows * nvec) return;
  const int row = (int)(idx / nvec), v = (int)(idx - (long long)row * nvec);
  const long long sr = src_idx ? src_idx[row] : row;
  const long long dr = dst_idx ? dst_idx[row] : row;
  const uint4 val = *reinterpret_cast<const uint4 *>(src + sr * lds + v * 8);
  *reinterpret_cast<uint4 *>(dst + dr * ldd + v * 8) = val;
}

// ───────────── paged KV: prefill write ─────────────
__global__ void __launch_bounds__(256)
kv_write_prefill_kernel(const bf16 *__restrict__ k, long long ldk, const bf16 *__restrict__ v, long long ldv,
                        bf16 *__restrict__ k_cache, bf16 *__restrict__ v_cache, const int32_t *__restrict__ block_table,
                        int max_pages, const int32_t *__restrict__ cu_seqlens, int n_seq, int T, int page_size,
                        int row_vec /* n_kv*hd/8 */, int hd_vec /* hd/8 */) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * row_vec) return;
  const int t = (int)(idx / row_vec), vv = (int)(idx - (long long)t * row_vec);
  int s = 0;
  while (s + 1 < n_seq && t >= cu_seqlens[s + 1]) ++s;
  const int pos = t - cu_seqlens[s];
  const int page = block_table[(size_t)s * max_pages + pos / page_size];
  // cache layout [page][kv head][token in page][hd]: the 16 tokens of a (page, kv head) are one contiguous block
  const int kvh = vv / hd_vec, dv = vv - kvh * hd_vec;
  const int n_kv = row_vec / hd_vec;
  const size_t dst = ((((size_t)page * n_kv + kvh) * page_size + pos % page_size) * hd_vec + dv) * 8;
  *reinterpret_cast<uint4 *>(k_cache + dst) = *reinterpret_cast<const uint4 *>(k + (size_t)t * ldk + vv * 8);
  *reinterpret_cast<uint4 *>(v_cache + dst) = *reinterpret_cast<const uint4 *>(v + (size_t)t * ldv + vv * 8);
}

// ───────────── greedy argmax + per-step bookkeeping: one CTA per sequence ─────────────
__global__ void __launch_bounds__(512)
argmax_step_kernel(const bf16 *__restrict__ logits, long long ldl, int V, int eos, int pad, int max_new,
                   int32_t *__restrict__ out_tokens, int32_t *__restrict__ next_ids, int32_t *__restrict__ finished,
                   int32_t *__restrict__ ctx_len, int32_t *__restrict__ step, int advance_ctx) {
  __shared__ float s_val[16];
  __shared__ int s_idx[16];
  const int b = blockIdx.x;
  const bf16 *row = logits + (size_t)b * ldl;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  const int nvec = V >> 3;
  for (int v = threadIdx.x; v < nvec; v += 512) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(row + v * 8);
    const bf16 *e = reinterpret_cast<const bf16 *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = __bfloat162float(e[k]);
      if (f > best) { best = f; best_i = v * 8 + k; }  // ascending index within a thread: first max kept
    }
  }
  for (int i = (nvec << 3) + threadIdx.x; i < V; i += 512) {
    const float f = __bfloat162float(row[i]);
    if (f > best || (f == best && i < best_i)) { best = f; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = best_i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 16; ++k)
      if (s_val[k] > best || (s_val[k] == best && s_idx[k] < best_i)) { best = s_val[k]; best_i = s_idx[k]; }
    int tok = best_i;
    const int st = step[0];
    if (finished[b]) tok = pad;
    if (st < max_new) out_tokens[(size_t)b * max_new + st] = tok;
    if (tok == eos) finished[b] = 1;
    next_ids[b] = tok;
    if (advance_ctx) ctx_len[b] += 1;
  }
}

__global__ void step_increment_kernel(int32_t *step) { step[0] += 1; }


// x[r, :] = bf16(x[r, :] + y[r, :]) -- the residual add that follows a tensor-parallel all-reduce (the fused residual
// epilogue of the GEMMs cannot be used there: the sum over ranks has to happen first).
__global__ void __launch_bounds__(256)
residual_add_kernel(bf16 *__restrict__ x, long long ldx, const bf16 *__restrict__ y, long long ldy, int rows, int vec_per_row) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * vec_per_row) return;
  const int r = (int)(idx / vec_per_row), v = (int)(idx % vec_per_row);
  uint4 a = *reinterpret_cast<const uint4 *>(x + (size_t)r * ldx + v * 8);
  const uint4 b = *reinterpret_cast<const uint4 *>(y + (size_t)r * ldy + v * 8);
  bf16 *ae = reinterpret_cast<bf16 *>(&a);
  const bf16 *be = reinterpret_cast<const bf16 *>(&b);
#pragma unroll
  for (int k = 0; k < 8; ++k) ae[k] = __float2bfloat16_rn(__bfloat162float(ae[k]) + __bfloat162float(be[k]));
  *reinterpret_cast<uint4 *>(x + (size_t)r * ldx + v * 8) = a;
}

}  // namespace ocrb
