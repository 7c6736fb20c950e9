// Batched greedy decode hot path (HF generation loop, utils.py:2743-2806, one token per sequence
// per step): weight-streaming skinny GEMM ("GEMV", HBM-bound: every weight byte is read exactly once
// per step for all B sequences) with HF's rounding points fused in, and split-KV paged attention.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ocrb {

typedef __nv_bfloat16 bf16;

constexpr int GV_WARPS = 8;       // warps per CTA
constexpr int GV_ROWS = 2;        // weight rows per warp (a gate/up pair under SwiGLU)
constexpr int GV_MAXB = 8;        // sequences per pass

__device__ __forceinline__ uint4 ld_stream(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void unpack8(const uint4 &raw, float *f) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}

__device__ __forceinline__ float silu_bf16(float g) {
  // torch silu on bf16: computed in fp32 as x / (1 + exp(-x)), rounded to bf16
  return bf16_round(g / (1.0f + expf(-g)));
}

__device__ __forceinline__ float gelu_erf_bf16(float x) {
  return bf16_round(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)));
}

// D[b, n] = epilogue( sum_k X[b,k] * W[n,k] ).  One warp owns GV_ROWS consecutive weight rows (under
// SwiGLU: rows n and n+64 of a 128-row [gate64|up64] tile) and streams them once with 16-byte loads;
// X (optionally RMS-normalised on the fly, HF rounding) sits in shared memory as bf16.
template <int NB, bool NORM>
__global__ void __launch_bounds__(GV_WARPS * 32)
gemv_kernel(const bf16 *__restrict__ X, long long ldx, const bf16 *__restrict__ W, long long ldw, bf16 *__restrict__ D,
            long long ldd, int N, int K, const bf16 *__restrict__ bias, const bf16 *__restrict__ residual, long long ldr,
            int epilogue, const bf16 *__restrict__ norm_w, float eps, bool x_in_smem) {
  extern __shared__ __align__(16) uint8_t gv_smem[];
  bf16 *xs = reinterpret_cast<bf16 *>(gv_smem);  // [NB][K] when x_in_smem
  __shared__ float s_red[GV_WARPS][NB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvec = K >> 3;

  if (NORM) {
    // HF RMSNorm prologue: rstd per sequence, then xs = bf16( w * bf16(x * rstd) )
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) ss[b] = 0.f;
    for (int v = threadIdx.x; v < kvec; v += GV_WARPS * 32) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(X + (size_t)b * ldx + v * 8);
        float f[8];
        unpack8(raw, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) ss[b] = fmaf(f[k], f[k], ss[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float v = ss[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][b] = v;
    }
    __syncthreads();
    float rstd[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) t += s_red[w][b];
      rstd[b] = rsqrtf(t / (float)K + eps);
    }
    for (int v = threadIdx.x; v < kvec; v += GV_WARPS * 32) {
      const uint4 wraw = *reinterpret_cast<const uint4 *>(norm_w + v * 8);
      float wf[8];
      unpack8(wraw, wf);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(X + (size_t)b * ldx + v * 8);
        float f[8];
        unpack8(raw, f);
        uint4 o;
        bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) oe[k] = __float2bfloat16_rn(wf[k] * bf16_round(f[k] * rstd[b]));
        *reinterpret_cast<uint4 *>(xs + (size_t)b * K + v * 8) = o;
      }
    }
    __syncthreads();
  } else if (x_in_smem) {
    for (int v = threadIdx.x; v < kvec; v += GV_WARPS * 32)
#pragma unroll
      for (int b = 0; b < NB; ++b)
        *reinterpret_cast<uint4 *>(xs + (size_t)b * K + v * 8) =
            *reinterpret_cast<const uint4 *>(X + (size_t)b * ldx + v * 8);
    __syncthreads();
  }
  const bf16 *xsrc = (NORM || x_in_smem) ? xs : X;
  const long long xstride = (NORM || x_in_smem) ? (long long)K : ldx;

  // rows of this warp
  const int pair = blockIdx.x * GV_WARPS + warp;
  int n0, n1, out_col;
  if (epilogue == OCRB_EPI_SWIGLU) {
    const int tile = pair >> 6, j = pair & 63;
    n0 = tile * 128 + j;       // gate row
    n1 = n0 + 64;              // up row
    out_col = tile * 64 + j;
    if (n1 >= N) return;
  } else {
    n0 = pair * 2;
    n1 = n0 + 1;
    out_col = n0;
    if (n0 >= N) return;
  }
  const bool has1 = n1 < N;
  const bf16 *w0 = W + (size_t)n0 * ldw;
  const bf16 *w1 = W + (size_t)(has1 ? n1 : n0) * ldw;

  float acc0[NB], acc1[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = 0.f;

  int v = lane;
  // main loop, 2x unrolled: 4 independent 16-byte weight loads in flight per lane
  for (; v + 32 < kvec; v += 64) {
    const uint4 a0 = ld_stream(w0 + v * 8), a1 = ld_stream(w1 + v * 8);
    const uint4 c0 = ld_stream(w0 + (v + 32) * 8), c1 = ld_stream(w1 + (v + 32) * 8);
    float fa0[8], fa1[8], fc0[8], fc1[8];
    unpack8(a0, fa0); unpack8(a1, fa1); unpack8(c0, fc0); unpack8(c1, fc1);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float x[8];
      unpack8(*reinterpret_cast<const uint4 *>(xsrc + (size_t)b * xstride + v * 8), x);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc0[b] = fmaf(fa0[k], x[k], acc0[b]); acc1[b] = fmaf(fa1[k], x[k], acc1[b]); }
      unpack8(*reinterpret_cast<const uint4 *>(xsrc + (size_t)b * xstride + (v + 32) * 8), x);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc0[b] = fmaf(fc0[k], x[k], acc0[b]); acc1[b] = fmaf(fc1[k], x[k], acc1[b]); }
    }
  }
  for (; v < kvec; v += 32) {
    const uint4 a0 = ld_stream(w0 + v * 8), a1 = ld_stream(w1 + v * 8);
    float fa0[8], fa1[8];
    unpack8(a0, fa0); unpack8(a1, fa1);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float x[8];
      unpack8(*reinterpret_cast<const uint4 *>(xsrc + (size_t)b * xstride + v * 8), x);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc0[b] = fmaf(fa0[k], x[k], acc0[b]); acc1[b] = fmaf(fa1[k], x[k], acc1[b]); }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc0[b] += __shfl_xor_sync(0xffffffffu, acc0[b], o);
      acc1[b] += __shfl_xor_sync(0xffffffffu, acc1[b], o);
    }
  }
  // epilogue: lane b finishes sequence b
  if (lane < NB) {
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (lane == b) { r0 = acc0[b]; r1 = acc1[b]; }
    if (bias) {
      r0 += __bfloat162float(bias[n0]);
      if (has1) r1 += __bfloat162float(bias[n1]);
    }
    r0 = bf16_round(r0);
    r1 = bf16_round(r1);
    bf16 *drow = D + (size_t)lane * ldd;
    if (epilogue == OCRB_EPI_SWIGLU) {
      drow[out_col] = __float2bfloat16_rn(silu_bf16(r0) * r1);
    } else {
      if (epilogue == OCRB_EPI_RESIDUAL) {
        const bf16 *rr = residual + (size_t)lane * ldr;
        r0 += __bfloat162float(rr[n0]);
        if (has1) r1 += __bfloat162float(rr[n1]);
      } else if (epilogue == OCRB_EPI_GELU) {
        r0 = gelu_erf_bf16(r0);
        r1 = gelu_erf_bf16(r1);
      }
      if (has1) {
        __nv_bfloat162 o2 = __floats2bfloat162_rn(r0, r1);
        *reinterpret_cast<__nv_bfloat162 *>(drow + n0) = o2;
      } else {
        drow[n0] = __float2bfloat16_rn(r0);
      }
    }
  }
}

// ───────────── paged decode attention, split over the KV length ─────────────
// grid = (n_splits, n_kv, B), 128 threads.  Each CTA serves all G = n_q/n_kv query heads of one KV
// head over one chunk of the context.  Phase 1: lane-per-key dot products (q in shared memory, fp32).
// Phase 2: chunk max / exp / sum.  Phase 3: thread-per-dim P.V.  Partial (m, l, acc) go to split_ws;
// a second kernel merges the splits.  The CTA with split 0 also applies mRoPE to the new token's
// q,k (bf16 arithmetic, as HF) and appends k,v to the cache; every CTA re-derives the roped q,k it
// needs from qkv, so there is no inter-CTA dependency.
constexpr int DA_MAXG = 8;
constexpr int DA_THREADS = 256;
constexpr int DA_TPK = 4;          // threads per key in the score phase

__device__ __forceinline__ float rope_elem_bf16(const bf16 *vec, int i, int hd, const bf16 *c, const bf16 *s) {
  const int half = hd >> 1;
  const float x = __bfloat162float(vec[i]);
  const float other = (i < half) ? -__bfloat162float(vec[i + half]) : __bfloat162float(vec[i - half]);
  const float t = bf16_round(x * __bfloat162float(c[i]));
  const float u = bf16_round(other * __bfloat162float(s[i]));
  return bf16_round(t + u);
}

__global__ void __launch_bounds__(DA_THREADS)
decode_attn_partial_kernel(const bf16 *__restrict__ qkv, long long ldqkv, bf16 *__restrict__ k_cache,
                           bf16 *__restrict__ v_cache, const int32_t *__restrict__ block_table, int max_pages,
                           const int32_t *__restrict__ ctx_len, int page_size, int n_q, int n_kv, int hd,
                           const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT, float scale,
                           float *__restrict__ split_ws, int n_splits, int chunk) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int G = n_q / n_kv;
  float *s_q = reinterpret_cast<float *>(da_smem);   // [G][hd]
  float *s_knew = s_q + G * hd;                      // [hd] roped new key
  float *s_p = s_knew + hd;                          // [chunk][8]  scores, then probabilities (key-major)
  float *s_red = s_p + (size_t)chunk * DA_MAXG;      // [4 warps][G][hd] partial P.V
  int *s_row = reinterpret_cast<int *>(s_red + (size_t)(DA_THREADS / 32) * G * hd);   // [chunk] cache row of each key
  __shared__ float s_m[DA_MAXG], s_l[DA_MAXG];
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) pdl_launch_dependents();
  pdl_wait();                          // qkv comes from the preceding skinny GEMM
  const int ctx = ctx_len[b];          // tokens already cached; the new token sits at position ctx
  const int total = ctx + 1;
  const int k0 = split * chunk;
  const int k1 = min(total, k0 + chunk);
  float *ws = split_ws + (((size_t)b * n_q + (size_t)kvh * G) * n_splits + split) * (hd + 2);
  const size_t ws_head = (size_t)n_splits * (hd + 2);
  if (k0 >= k1 && split != 0) {        // chunk beyond the context: publish an empty partial and leave
    for (int g = tid; g < G; g += DA_THREADS) { ws[g * ws_head] = -INFINITY; ws[g * ws_head + 1] = 0.f; }
    return;
  }
  const bf16 *row = qkv + (size_t)b * ldqkv;
  const bf16 *c = cosT + (size_t)b * hd, *s = sinT + (size_t)b * hd;
  const bf16 *knew = row + (size_t)n_q * hd + (size_t)kvh * hd;
  const bf16 *vnew = row + (size_t)(n_q + n_kv) * hd + (size_t)kvh * hd;
  for (int i = tid; i < G * hd; i += DA_THREADS) {
    const int g = i / hd, d = i - g * hd;
    s_q[i] = rope_elem_bf16(row + (size_t)(kvh * G + g) * hd, d, hd, c, s);
  }
  for (int d = tid; d < hd; d += DA_THREADS) s_knew[d] = rope_elem_bf16(knew, d, hd, c, s);
  const int32_t *bt = block_table + (size_t)b * max_pages;
  for (int i = tid; i < k1 - k0; i += DA_THREADS) {
    const int key = k0 + i;
    s_row[i] = (key < ctx) ? bt[key / page_size] * page_size + key % page_size : -1;   // -1: the new token (not cached yet)
  }
  __syncthreads();
  const size_t tok_stride = (size_t)n_kv * hd;
  if (split == 0) {
    // append the new token's k (roped) and v
    const int page = bt[ctx / page_size];
    const size_t dst = ((size_t)page * page_size + ctx % page_size) * tok_stride + (size_t)kvh * hd;
    for (int d = tid; d < hd; d += DA_THREADS) {
      k_cache[dst + d] = __float2bfloat16_rn(s_knew[d]);
      v_cache[dst + d] = vnew[d];
    }
  }
  const int nk = k1 - k0;
  // phase 1: scores.  DA_TPK threads share a key (each owns hd/DA_TPK contiguous dims, fetched before the first
  // FMA), partial dot products meet through two xor-shuffles: short dependent chains, 8 warps per CTA.
  {
    const int part = tid & (DA_TPK - 1);
    const int dpp = hd / DA_TPK;                 // dims per part (multiple of 8)
    const int dbase = part * dpp;
    for (int base = k0; base < k1; base += DA_THREADS / DA_TPK) {   // block-uniform trip count (shuffles below)
      const int key = base + tid / DA_TPK;
      const bool live = key < k1;
      float sc[DA_MAXG];
#pragma unroll
      for (int g = 0; g < DA_MAXG; ++g) sc[g] = 0.f;
      if (live) {
        const int r = s_row[key - k0];
        if (r < 0) {
          for (int d = dbase; d < dbase + dpp; ++d) {
            const float kv = s_knew[d];
#pragma unroll
            for (int g = 0; g < DA_MAXG; ++g)
              if (g < G) sc[g] = fmaf(s_q[g * hd + d], kv, sc[g]);
          }
        } else {
          const bf16 *kr = k_cache + (size_t)r * tok_stride + (size_t)kvh * hd + dbase;
          if (dpp == 32) {
            uint4 raw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) raw[j] = *reinterpret_cast<const uint4 *>(kr + j * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float kf[8];
              unpack8(raw[j], kf);
#pragma unroll
              for (int g = 0; g < DA_MAXG; ++g)
                if (g < G) {
                  const float4 qa = *reinterpret_cast<const float4 *>(s_q + g * hd + dbase + j * 8);
                  const float4 qb = *reinterpret_cast<const float4 *>(s_q + g * hd + dbase + j * 8 + 4);
                  float a = sc[g];
                  a = fmaf(qa.x, kf[0], a); a = fmaf(qa.y, kf[1], a); a = fmaf(qa.z, kf[2], a); a = fmaf(qa.w, kf[3], a);
                  a = fmaf(qb.x, kf[4], a); a = fmaf(qb.y, kf[5], a); a = fmaf(qb.z, kf[6], a); a = fmaf(qb.w, kf[7], a);
                  sc[g] = a;
                }
            }
          } else {
            for (int d8 = 0; d8 < dpp; d8 += 8) {
              float kf[8];
              unpack8(*reinterpret_cast<const uint4 *>(kr + d8), kf);
#pragma unroll
              for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int g = 0; g < DA_MAXG; ++g)
                  if (g < G) sc[g] = fmaf(s_q[g * hd + dbase + d8 + e], kf[e], sc[g]);
            }
          }
        }
      }
#pragma unroll
      for (int g = 0; g < DA_MAXG; ++g) {
        sc[g] += __shfl_xor_sync(0xffffffffu, sc[g], 1);
        sc[g] += __shfl_xor_sync(0xffffffffu, sc[g], 2);
      }
      if (live && part == 0) {
        float4 lo = make_float4(sc[0] * scale, sc[1] * scale, sc[2] * scale, sc[3] * scale);
        float4 hi = make_float4(sc[4] * scale, sc[5] * scale, sc[6] * scale, sc[7] * scale);
        *reinterpret_cast<float4 *>(s_p + (size_t)(key - k0) * DA_MAXG) = lo;
        *reinterpret_cast<float4 *>(s_p + (size_t)(key - k0) * DA_MAXG + 4) = hi;
      }
    }
  }
  __syncthreads();
  // phase 2: per-head max / exp / sum (warp w handles heads w, w+4, ...)
  for (int g = warp; g < G; g += DA_THREADS / 32) {
    float m = -INFINITY;
    for (int i = lane; i < nk; i += 32) m = fmaxf(m, s_p[i * DA_MAXG + g]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int i = lane; i < nk; i += 32) {
      const float p = __expf(s_p[i * DA_MAXG + g] - m);
      s_p[i * DA_MAXG + g] = p;
      l += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) { s_m[g] = m; s_l[g] = l; }
  }
  __syncthreads();
  // phase 3: P.V.  Warp w takes keys w, w+4, ...; lane l owns dims [4l, 4l+4) (+128 per pass for hd > 128),
  // so every V row is read once, coalesced, 8 bytes per lane.  Partial sums meet in shared memory.
  for (int d0 = 0; d0 < hd; d0 += 128) {
    const int d = d0 + lane * 4;
    const bool d_ok = d < hd;
    float acc[DA_MAXG][4];
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
    constexpr int NW = DA_THREADS / 32;
    constexpr int VB = 8;                      // V rows in flight per lane
    for (int i0 = warp; i0 < nk; i0 += NW * VB) {
      uint2 raw[VB];
#pragma unroll
      for (int j = 0; j < VB; ++j) {
        const int i = i0 + j * NW;
        raw[j] = make_uint2(0, 0);
        if (i < nk && d_ok) {
          const int r = s_row[i];
          const bf16 *vp = (r < 0) ? (vnew + d) : (v_cache + (size_t)r * tok_stride + (size_t)kvh * hd + d);
          raw[j] = *reinterpret_cast<const uint2 *>(vp);
        }
      }
#pragma unroll
      for (int j = 0; j < VB; ++j) {
        const int i = i0 + j * NW;
        if (i < nk) {
          const float v0 = __uint_as_float(raw[j].x << 16), v1 = __uint_as_float(raw[j].x & 0xffff0000u);
          const float v2 = __uint_as_float(raw[j].y << 16), v3 = __uint_as_float(raw[j].y & 0xffff0000u);
          const float4 pa = *reinterpret_cast<const float4 *>(s_p + (size_t)i * DA_MAXG);
          const float4 pb = *reinterpret_cast<const float4 *>(s_p + (size_t)i * DA_MAXG + 4);
          const float pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
          for (int g = 0; g < DA_MAXG; ++g)
            if (g < G) {
              acc[g][0] = fmaf(pp[g], v0, acc[g][0]);
              acc[g][1] = fmaf(pp[g], v1, acc[g][1]);
              acc[g][2] = fmaf(pp[g], v2, acc[g][2]);
              acc[g][3] = fmaf(pp[g], v3, acc[g][3]);
            }
        }
      }
    }
    if (d_ok) {
#pragma unroll
      for (int g = 0; g < DA_MAXG; ++g)
        if (g < G)
          *reinterpret_cast<float4 *>(s_red + ((size_t)warp * G + g) * hd + d) =
              make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
    }
  }
  __syncthreads();
  for (int i = tid; i < G * hd; i += DA_THREADS) {
    const int g = i / hd, d = i - g * hd;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < DA_THREADS / 32; ++w) a += s_red[((size_t)w * G + g) * hd + d];
    ws[g * ws_head + 2 + d] = a;
  }
  for (int g = tid; g < G; g += DA_THREADS) { ws[g * ws_head] = s_m[g]; ws[g * ws_head + 1] = s_l[g]; }
}

// ───────────── tensor-core variant (hd = 128, chunk <= 64 keys, G <= 16 query heads per KV head) ─────────────
// The SIMT kernel above is bound by per-warp instruction latency (~5k dependent instructions per warp).  Here the
// G query heads of one KV head are the M = 16 rows of mma.sync.m16n8k16 tiles (rows >= G are zero), so a 64-key chunk
// costs 32 MMAs per warp: S = Q.K^T with K rows as the "col" operand, softmax in fp32 in shared memory, O = P.V with
// V through ldmatrix.trans.  K/V rows are fetched with cp.async BEFORE griddepcontrol.wait (they were written by
// earlier steps, not by the preceding qkv GEMM), so the loads overlap the predecessor's tail under PDL.
constexpr int DM_CH = 64;              // keys per CTA
constexpr int DM_HD = 128;
constexpr int DM_PITCH = DM_HD + 8;    // bf16 elements per shared row (272 B: conflict-free ldmatrix)
constexpr int DM_SP = DM_CH + 8;       // P row pitch (bf16)
constexpr int DM_SS = DM_CH + 4;       // S row pitch (fp32)

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
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

__global__ void __launch_bounds__(128)
decode_attn_mma_kernel(const bf16 *__restrict__ qkv, long long ldqkv, bf16 *__restrict__ k_cache, bf16 *__restrict__ v_cache,
                       const int32_t *__restrict__ block_table, int max_pages, const int32_t *__restrict__ ctx_len,
                       int page_size, int n_q, int n_kv, const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT,
                       float scale, float *__restrict__ split_ws, int n_splits, int chunk, int early) {
  constexpr int hd = DM_HD;
  __shared__ __align__(16) bf16 sQ[16 * DM_PITCH];
  __shared__ __align__(16) bf16 sK[DM_CH * DM_PITCH];
  __shared__ __align__(16) bf16 sV[DM_CH * DM_PITCH];
  __shared__ __align__(16) float sS[16 * DM_SS];
  __shared__ __align__(16) bf16 sP[16 * DM_SP];
  __shared__ float s_m[16], s_l[16];
  const int G = n_q / n_kv;
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) pdl_launch_dependents();
  if (early == 0) pdl_wait();
  const int ctx = ctx_len[b];          // written by the previous step's argmax kernel: complete long before this launch
  const int total = ctx + 1;
  const int k0 = split * chunk;
  const int k1 = min(total, k0 + chunk);
  float *ws = split_ws + (((size_t)b * n_q + (size_t)kvh * G) * n_splits + split) * (hd + 2);
  const size_t ws_head = (size_t)n_splits * (hd + 2);
  if (k0 >= k1) {
    if (early != 0) pdl_wait();        // split_ws may still be read by the previous layer's combine kernel
    for (int g = tid; g < G; g += 128) { ws[g * ws_head] = -INFINITY; ws[g * ws_head + 1] = 0.f; }
    return;
  }
  const int nk = k1 - k0;
  const size_t tok_stride = (size_t)n_kv * hd;
  const int32_t *bt = block_table + (size_t)b * max_pages;
  if (early == 1) pdl_wait();
  // ---- K / V rows of already cached tokens: 16-byte cp.async, 16 per thread; other rows are zero-filled ----
  {
    const int c16 = tid & 15;                      // 16-byte column of the row
#pragma unroll
    for (int i = 0; i < DM_CH / 8; ++i) {
      const int r = (tid >> 4) + i * 8;
      const int key = k0 + r;
      bf16 *dk = sK + r * DM_PITCH + c16 * 8, *dv = sV + r * DM_PITCH + c16 * 8;
      if (key < ctx) {
        const size_t row = (size_t)bt[key / page_size] * page_size + key % page_size;
        cp_async16(dk, k_cache + row * tok_stride + (size_t)kvh * hd + c16 * 8);
        cp_async16(dv, v_cache + row * tok_stride + (size_t)kvh * hd + c16 * 8);
      } else if (key != ctx) {
        *reinterpret_cast<uint4 *>(dk) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(dv) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < 16 * DM_SP / 2; i += 128) reinterpret_cast<uint32_t *>(sP)[i] = 0u;
  if (early == 2) pdl_wait();          // qkv comes from the preceding skinny GEMM
  const bf16 *row = qkv + (size_t)b * ldqkv;
  const bf16 *c = cosT + (size_t)b * hd, *s = sinT + (size_t)b * hd;
  for (int i = tid; i < 16 * hd; i += 128) {
    const int g = i >> 7, d = i & 127;
    float v = 0.f;
    if (g < G) v = rope_elem_bf16(row + (size_t)(kvh * G + g) * hd, d, hd, c, s);
    sQ[g * DM_PITCH + d] = __float2bfloat16_rn(v);
  }
  if (ctx >= k0 && ctx < k0 + chunk) {
    // this chunk holds the new token: rope its key, place k/v in the tile and append them to the cache
    const bf16 *knew = row + (size_t)n_q * hd + (size_t)kvh * hd;
    const bf16 *vnew = row + (size_t)(n_q + n_kv) * hd + (size_t)kvh * hd;
    const size_t dst = ((size_t)bt[ctx / page_size] * page_size + ctx % page_size) * tok_stride + (size_t)kvh * hd;
    const int r = ctx - k0;
    for (int d = tid; d < hd; d += 128) {
      const bf16 kr = __float2bfloat16_rn(rope_elem_bf16(knew, d, hd, c, s));
      const bf16 vr = vnew[d];
      sK[r * DM_PITCH + d] = kr;
      sV[r * DM_PITCH + d] = vr;
      k_cache[dst + d] = kr;
      v_cache[dst + d] = vr;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // ---- S = Q.K^T: warp w owns keys [16w, 16w+16) ----
  {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int m = lane >> 3, l8 = lane & 7;
#pragma unroll
    for (int kk = 0; kk < hd / 16; ++kk) {
      uint32_t a[4], bb[4];
      ldmatrix_x4(a, sQ + ((m & 1) * 8 + l8) * DM_PITCH + kk * 16 + (m >> 1) * 8);
      ldmatrix_x4(bb, sK + (warp * 16 + (m >> 1) * 8 + l8) * DM_PITCH + kk * 16 + (m & 1) * 8);
      mma_bf16_16816(acc[0], a, bb[0], bb[1]);
      mma_bf16_16816(acc[1], a, bb[2], bb[3]);
    }
    const int r0 = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = warp * 16 + j * 8 + cq;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = (col + e) < nk;
        sS[r0 * DM_SS + col + e] = ok ? acc[j][e] * scale : -INFINITY;
        sS[(r0 + 8) * DM_SS + col + e] = ok ? acc[j][2 + e] * scale : -INFINITY;
      }
    }
  }
  __syncthreads();
  // ---- softmax statistics of the chunk (fp32) and P in bf16: warp w handles heads w, w+4, ... ----
  for (int g = warp; g < G; g += 4) {
    const float x0 = sS[g * DM_SS + lane], x1 = sS[g * DM_SS + 32 + lane];
    float mx = fmaxf(x0, x1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float p0 = __expf(x0 - mx), p1 = __expf(x1 - mx);
    float l = p0 + p1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    sP[g * DM_SP + lane] = __float2bfloat16_rn(p0);
    sP[g * DM_SP + 32 + lane] = __float2bfloat16_rn(p1);
    if (lane == 0) { s_m[g] = mx; s_l[g] = l; }
  }
  __syncthreads();
  // ---- O = P.V: warp w owns dims [32w, 32w+32) ----
  {
    float o[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    const int m = lane >> 3, l8 = lane & 7;
#pragma unroll
    for (int ks = 0; ks < DM_CH / 16; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(a, sP + ((m & 1) * 8 + l8) * DM_SP + ks * 16 + (m >> 1) * 8);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, sV + (ks * 16 + (m & 1) * 8 + l8) * DM_PITCH + warp * 32 + jj * 16 + (m >> 1) * 8);
        mma_bf16_16816(o[jj * 2], a, bb[0], bb[1]);
        mma_bf16_16816(o[jj * 2 + 1], a, bb[2], bb[3]);
      }
    }
    const int r0 = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = warp * 32 + j * 8 + cq;
      if (r0 < G) *reinterpret_cast<float2 *>(ws + r0 * ws_head + 2 + d) = make_float2(o[j][0], o[j][1]);
      if (r0 + 8 < G) *reinterpret_cast<float2 *>(ws + (r0 + 8) * ws_head + 2 + d) = make_float2(o[j][2], o[j][3]);
    }
  }
  for (int g = tid; g < G; g += 128) { ws[g * ws_head] = s_m[g]; ws[g * ws_head + 1] = s_l[g]; }
}

// grid = (n_q, B), hd threads
__global__ void decode_attn_combine_kernel(const float *__restrict__ split_ws, int n_q, int n_splits, int hd,
                                           bf16 *__restrict__ out, long long ldo) {
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  pdl_wait();
  // Trigger only now: the o_proj GEMM behind this kernel then overlaps this (short) kernel, not the attention
  // kernel before it.  Letting the GEMM prologue start during decode_attn_mma_kernel faulted on small grids
  // (illegal address, cause not established -- DESIGN.md "PDL"), and measured no faster.
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

template <int NB>
static int launch_gemv(const bf16 *X, long long ldx, const bf16 *W, long long ldw, bf16 *D, long long ldd, int N, int K,
                       const bf16 *bias, const bf16 *residual, long long ldr, int epilogue, const bf16 *norm_w, float eps,
                       cudaStream_t st) {
  const int pairs = (epilogue == OCRB_EPI_SWIGLU) ? N / 2 : (N + 1) / 2;
  const int blocks = cdiv(pairs, GV_WARPS);
  const size_t xbytes = (size_t)NB * K * sizeof(bf16);
  const bool norm = norm_w != nullptr;
  const bool x_in_smem = norm || xbytes <= 48 * 1024;
  OCRB_REQUIRE(!norm || xbytes <= 200 * 1024, "gemv_bf16: fused RMSNorm needs B*K*2 <= 200 KiB of shared memory");
  const size_t smem = x_in_smem ? xbytes : 0;
  if (norm) {
    static size_t attr_smem = 48 * 1024;  // raise the dynamic shared-memory limit once per size class
    if (smem > attr_smem) {
      OCRB_CUDA(cudaFuncSetAttribute(gemv_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    gemv_kernel<NB, true><<<blocks, GV_WARPS * 32, smem, st>>>(X, ldx, W, ldw, D, ldd, N, K, bias, residual, ldr, epilogue,
                                                                norm_w, eps, true);
  } else {
    gemv_kernel<NB, false><<<blocks, GV_WARPS * 32, smem, st>>>(X, ldx, W, ldw, D, ldd, N, K, bias, residual, ldr, epilogue,
                                                                 nullptr, eps, x_in_smem);
  }
  return check_launch("gemv_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_gemv_bf16(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int32_t B,
                              int32_t N, int32_t K, const void *bias, const void *residual, int64_t ldr, int32_t epilogue,
                              const void *norm_w, float eps, void *stream) {
  OCRB_REQUIRE(A && W && D, "gemv_bf16: null pointer");
  OCRB_REQUIRE(B >= 1 && B <= GV_MAXB, "gemv_bf16: B must be in 1..8 (use ocrb_gemm_bf16 for larger batches)");
  OCRB_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemv_bf16: K and strides must be multiples of 8");
  OCRB_REQUIRE(epilogue >= 0 && epilogue <= 3, "gemv_bf16: bad epilogue");
  OCRB_REQUIRE(epilogue != OCRB_EPI_RESIDUAL || residual, "gemv_bf16: residual epilogue without residual");
  OCRB_REQUIRE(epilogue != OCRB_EPI_SWIGLU || N % 128 == 0, "gemv_bf16: SwiGLU needs N (packed gate|up rows) % 128 == 0");
  OCRB_REQUIRE(epilogue == OCRB_EPI_SWIGLU || ldd % 2 == 0, "gemv_bf16: ldd must be even");
  cudaStream_t st = (cudaStream_t)stream;
  const bf16 *X = (const bf16 *)A, *Wp = (const bf16 *)W, *bp = (const bf16 *)bias, *rp = (const bf16 *)residual,
             *nw = (const bf16 *)norm_w;
  bf16 *Dp = (bf16 *)D;
#define GV_CASE(nb) \
  case nb: return launch_gemv<nb>(X, lda, Wp, ldw, Dp, ldd, N, K, bp, rp, ldr, epilogue, nw, eps, st);
  switch (B) {
    GV_CASE(1) GV_CASE(2) GV_CASE(3) GV_CASE(4) GV_CASE(5) GV_CASE(6) GV_CASE(7) GV_CASE(8)
  }
#undef GV_CASE
  return OCRB_EINVAL;
}

extern "C" int ocrb_decode_attention(const void *qkv, int64_t ldqkv, void *k_cache, void *v_cache,
                                     const int32_t *block_table, int32_t max_pages, const int32_t *ctx_len, int32_t B,
                                     int32_t page_size, int32_t n_q, int32_t n_kv, int32_t hd, const void *cosT,
                                     const void *sinT, float scale, void *out, int64_t ldo, float *split_ws,
                                     int32_t n_splits, void *stream) {
  OCRB_REQUIRE(qkv && k_cache && v_cache && block_table && ctx_len && cosT && sinT && out && split_ws,
               "decode_attention: null pointer");
  OCRB_REQUIRE(B > 0 && n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= DA_MAXG && hd % 32 == 0 && hd <= 256 && n_splits > 0,
               "decode_attention: bad sizes");
  const int max_ctx = max_pages * page_size;
  const int chunk = cdiv(max_ctx, n_splits);
  const int G = n_q / n_kv;
  const size_t smem = ((size_t)G * hd + hd + (size_t)DA_MAXG * chunk + (size_t)(DA_THREADS / 32) * G * hd) * sizeof(float) +
                      (size_t)chunk * sizeof(int);
  OCRB_REQUIRE(smem <= 200 * 1024, "decode_attention: chunk too large, raise n_splits");
  cudaStream_t st = (cudaStream_t)stream;
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    OCRB_CUDA(cudaFuncSetAttribute(decode_attn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  static int use_mma = -1, early = 2;
  if (use_mma < 0) {
    const char *e = getenv("OCRB_ATTN_SIMT");
    use_mma = (e && e[0] == '1') ? 0 : 1;
    const char *e2 = getenv("OCRB_ATTN_EARLY");
    if (e2) early = atoi(e2);
  }
  if (use_mma && hd == DM_HD && G <= 16 && chunk <= DM_CH) {
    OCRB_CUDA(launch_pdl_bit(2, decode_attn_mma_kernel, dim3(n_splits, n_kv, B), dim3(128), 0, st, (const bf16 *)qkv, (long long)ldqkv,
                         (bf16 *)k_cache, (bf16 *)v_cache, block_table, (int)max_pages, ctx_len, (int)page_size, (int)n_q,
                         (int)n_kv, (const bf16 *)cosT, (const bf16 *)sinT, scale, split_ws, (int)n_splits, chunk, early));
    int rc = check_launch("decode_attn_mma_kernel");
    if (rc) return rc;
    OCRB_CUDA(launch_pdl_bit(4, decode_attn_combine_kernel, dim3(n_q, B), dim3(hd), 0, st, (const float *)split_ws, (int)n_q,
                         (int)n_splits, (int)hd, (bf16 *)out, (long long)ldo));
    return check_launch("decode_attn_combine_kernel");
  }
  OCRB_CUDA(launch_pdl_bit(2, decode_attn_partial_kernel, dim3(n_splits, n_kv, B), dim3(DA_THREADS), smem, st, (const bf16 *)qkv,
                       (long long)ldqkv, (bf16 *)k_cache, (bf16 *)v_cache, block_table, (int)max_pages, ctx_len,
                       (int)page_size, (int)n_q, (int)n_kv, (int)hd, (const bf16 *)cosT, (const bf16 *)sinT, scale, split_ws,
                       (int)n_splits, chunk));
  int rc = check_launch("decode_attn_partial_kernel");
  if (rc) return rc;
  OCRB_CUDA(launch_pdl_bit(4, decode_attn_combine_kernel, dim3(n_q, B), dim3(hd), 0, st, (const float *)split_ws, (int)n_q,
                       (int)n_splits, (int)hd, (bf16 *)out, (long long)ldo));
  return check_launch("decode_attn_combine_kernel");
}
