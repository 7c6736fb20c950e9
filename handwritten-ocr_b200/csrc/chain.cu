// A CHAIN of dependent weight-streaming linears of one decode step in ONE persistent launch (HF generation loop,
// utils.py:2743-2806; decoder layer modeling_qwen2_5_vl.py:839-879): o_proj + residual -> RMSNorm -> gate/up + SwiGLU ->
// down_proj + residual -> RMSNorm -> qkv of the next layer (or final norm + lm_head).  Same arithmetic, rounding points
// and stream-K split as skinny_gemm_kernel (skinny.cu) -- a linear produces the same bits through either path -- but:
//   * HBM never drains between the linears: the weight producer walks the whole chain and runs AHEAD of the
//     dependencies (weights never depend on activations); one CTA per SM with a ring as deep as shared memory
//     allows (12 x 18 KiB at B <= 16: 29 MB in flight over 148 SMs ~ 4.5 us of HBM time) rides over the bubbles
//     between dependent linears instead of paying a launch + pipeline fill + drain for each of them;
//   * the dependencies are device-side: every CTA bumps a counter when its part of linear g is stored, the
//     activation producer of linear g+1 polls it (ld.acquire.gpu) before it issues the activation TMA loads;
//   * an RMSNorm in front of a linear is done in the chain: row r is normalised by one warp of CTA r (same routine and
//     summation order as skinny_norm_rows_kernel), announced through a second counter;
//   * counters and stream-K flags are returned to zero by the last CTA to leave, so a CUDA graph can replay the launch.
//
// Warp roles (224 threads): 0 = weight TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue (TMEM lane
// quadrant = warp % 4), 6 = activation TMA producer + RMSNorm rows.
#include "skinny_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace ocrb {

constexpr int CH_MAXD = OCRB_CHAIN_MAX;
constexpr int CH_THREADS = 224;
constexpr int CH_TRACE_SLOTS = 64;

struct alignas(64) ChainDesc {
  CUtensorMap map_w;                 // [N, K] weights, box [128 x 64]
  CUtensorMap map_x;                 // [B, K] activations (the normalised copy when norm_w != nullptr), box [BP x 64]
  const bf16 *Xraw; long long ldx;   // rows to normalise (norm_w != nullptr)
  bf16 *xn;                          // normalised rows [B][K]
  const bf16 *norm_w; float eps;
  bf16 *D; long long ldd;
  const bf16 *bias;
  const bf16 *residual; long long ldr;
  int N, K, num_tiles, num_kb, epilogue;
  int res_early;                     // the residual does not come from the linear right before: fetch it ahead
};

struct ChainParams {
  ChainDesc d[CH_MAXD];
  int n_desc, B;
  float *partials;                   // [grid][BC][128] fp32 stream-K partials (one slot per CTA, reused along the chain)
  int *flags;                        // [CH_MAXD][SK_MAX_GRID] partial-ready flags
  int *done;                         // [CH_MAXD] CTAs whose part of linear g is stored
  int *norm_done;                    // [CH_MAXD] rows normalised for linear g
  int *exit_count;
  unsigned long long *trace;         // optional [grid][64] globaltimer stamps
};

__device__ __forceinline__ void ch_stamp(const ChainParams &p, int slot) {
  if (p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[(size_t)blockIdx.x * CH_TRACE_SLOTS + slot] = t;
  }
}

// Bounded poll of a device counter (a protocol bug must trap, not hang the GPU).
__device__ __forceinline__ void ch_wait_count(const int *ctr, int target, const char *what, int g) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  if (v >= target) return;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v < target && clock64() - t0 > 4000000000LL) {
      printf("ocrb chain: %s of linear %d stuck at %d of %d (CTA %d)\n", what, g, v, target, blockIdx.x);
      __trap();
    }
  } while (v < target);
}
__device__ __forceinline__ void ch_add_release(int *ctr, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(v) : "memory");
}

// One row through HF's RMSNorm (modeling_qwen2_5_vl.py:66-71), one warp: identical arithmetic and summation order to
// skinny_norm_rows_kernel, so the bits do not depend on which path normalised the row.
__device__ __forceinline__ void ch_norm_row(const bf16 *xr, bf16 *yr, const bf16 *w, int K, float eps, int lane) {
  const int kvec = K >> 3;
  if (kvec <= 512) {
    uint4 raw[16], wraw[16];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int v = lane + i * 32;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (v < kvec) raw[i] = __ldcg(reinterpret_cast<const uint4 *>(xr + v * 8));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {               // the weight row travels with the activation row: one round trip, not two
      const int v = lane + i * 32;
      wraw[i] = make_uint4(0, 0, 0, 0);
      if (v < kvec) wraw[i] = __ldg(reinterpret_cast<const uint4 *>(w + v * 8));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float f[8];
      unpack8f(raw[i], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) ss = fmaf(f[k], f[k], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rs = rsqrtf(ss / (float)K + eps);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int v = lane + i * 32;
      if (v < kvec) {
        float f[8], wf[8];
        unpack8f(raw[i], f);
        unpack8f(wraw[i], wf);
        uint4 o;
        bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
        for (int e = 0; e < 8; ++e) oe[e] = __float2bfloat16_rn(wf[e] * bf16_round(f[e] * rs));
        *reinterpret_cast<uint4 *>(yr + v * 8) = o;
      }
    }
  } else {
    const float rs = sk_one_row_rstd(xr, K, eps, lane);
    for (int v = lane; v < kvec; v += 32) {
      float f[8], wf[8];
      unpack8f(__ldcg(reinterpret_cast<const uint4 *>(xr + v * 8)), f);
      unpack8f(*reinterpret_cast<const uint4 *>(w + v * 8), wf);
      uint4 o;
      bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) oe[e] = __float2bfloat16_rn(wf[e] * bf16_round(f[e] * rs));
      *reinterpret_cast<uint4 *>(yr + v * 8) = o;
    }
  }
}

template <int BP>
struct ChainCfg {
  // ring depth: as deep as 227 KiB allow next to the barriers and the SwiGLU exchange buffer
  static constexpr int ST = (BP <= 16) ? 12 : ((BP <= 32) ? 10 : ((BP <= 64) ? 9 : ((BP <= 96) ? 7 : 6)));
  static constexpr uint32_t X_BYTES = BP * SK_BK * 2;
  static constexpr uint32_t STAGE_BYTES = SK_W_BYTES + X_BYTES;
  static constexpr int ACC_STRIDE = (BP <= 16) ? 16 : (BP <= 32 ? 32 : (BP <= 64 ? 64 : 128));
  static constexpr int TMEM_COLS = (2 * ACC_STRIDE < 32) ? 32 : 2 * ACC_STRIDE;
  static constexpr size_t SMEM = (size_t)ST * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 64 * 16 * sizeof(float) + 64;
};

template <int BP, int BC>
__global__ void __launch_bounds__(CH_THREADS, 1)
skinny_chain_kernel(const __grid_constant__ ChainParams p) {
  using C = ChainCfg<BP>;
  constexpr int ST = C::ST;
  constexpr uint32_t X_BYTES = C::X_BYTES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int ACC_STRIDE = C::ACC_STRIDE, TMEM_COLS = C::TMEM_COLS;
  extern __shared__ uint8_t ch_smem_raw[];
  uint8_t *smem = ch_smem_raw + ((1024u - (smem_u32(ch_smem_raw) & 1023u)) & 1023u);
  uint64_t *full_w = reinterpret_cast<uint64_t *>(smem + ST * STAGE_BYTES);
  uint64_t *full_x = full_w + ST;
  uint64_t *empty = full_x + ST;
  uint64_t *tmem_full = empty + ST;            // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  int *s_ticket = reinterpret_cast<int *>(tmem_slot + 1);
  float *s_up = reinterpret_cast<float *>(smem + ST * STAGE_BYTES + 512);     // [64][CC] SwiGLU exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  if (threadIdx.x == 0) ch_stamp(p, 0);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.n_desc; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.d[g].map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.d[g].map_x) : "memory");
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) { ch_stamp(p, 1); pdl_launch_dependents(); }

  if (warp == 0) {
    // ───────────── weight producer: the whole chain, never waits for a dependency ─────────────
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int s = 0;
      uint32_t round = 0;
      for (int g = 0; g < p.n_desc; ++g) {
        const ChainDesc &d = p.d[g];
        SkSpan sp;
        sp.init(cta, G, d.num_tiles, d.num_kb);
        const int n_units = sp.num_units();
        SkCursor cur;
        cur.init(sp);
        for (int it = 0; it < n_units; ++it) {
          mbar_wait(&empty[s], (round & 1u) ^ 1u);
          mbar_expect_tx(&full_w[s], SK_W_BYTES);
          tma_load_2d_hint(smem + s * STAGE_BYTES, &d.map_w, &full_w[s], cur.kb * SK_BK, cur.tile * SK_BM, policy);
          cur.advance(sp);
          if (++s == ST) { s = 0; ++round; }
        }
      }
    }
  } else if (warp == 1) {
    // ───────────── MMA issuer ─────────────
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BP >> 3) << 17) | ((uint32_t)(SK_BM >> 4) << 24);
      int s = 0;
      uint32_t round = 0;
      int segc = 0;                              // segments so far along the chain (TMEM buffer = segc & 1)
      for (int g = 0; g < p.n_desc; ++g) {
        const ChainDesc &d = p.d[g];
        SkSpan sp;
        sp.init(cta, G, d.num_tiles, d.num_kb);
        const int n_segs = sp.num_units() > 0 ? sp.num_segs() : 0;
        for (int seg = 0; seg < n_segs; ++seg, ++segc) {
          int tile, kb0, nkb;
          sp.seg(seg, tile, kb0, nkb);
          const int acc = segc & 1;
          mbar_wait(&tmem_empty[acc], ((segc >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + acc * ACC_STRIDE;
          for (int i = 0; i < nkb; ++i) {
            mbar_wait(&full_w[s], round & 1u);
            mbar_wait(&full_x[s], round & 1u);
            if (seg == 0 && i == 0) ch_stamp(p, 8 + g * 8 + 2);      // first k-block of linear g ready
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + SK_W_BYTES);
#pragma unroll
            for (int k = 0; k < SK_BK / UMMA_K; ++k)
              umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            if (++s == ST) { s = 0; ++round; }
          }
          umma_commit(&tmem_full[acc]);
        }
      }
    }
  } else if (warp == 6) {
    // ───────────── activation producer: dependencies, RMSNorm rows, activation TMA ─────────────
    pdl_wait();                                     // the first linear's input comes from the preceding kernel
    int s = 0;
    uint32_t round = 0;
    for (int g = 0; g < p.n_desc; ++g) {
      const ChainDesc &d = p.d[g];
      if (d.norm_w && cta < p.B) {
        // HBM is saturated by the weight stream when the rows become ready: a cold read of the norm weights then waits
        // microseconds in the DRAM queues.  Pull them into L2 before the dependency wait.
        for (int off = lane * 128; off < d.K * 2; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(d.norm_w) + off));
      }
      if (g > 0) {
        if (lane == 0) ch_wait_count(p.done + (g - 1), G, "completion", g - 1);
        __syncwarp();
      }
      if (lane == 0) ch_stamp(p, 8 + g * 8 + 0);       // inputs of linear g complete
      if (d.norm_w) {
        int rows = 0;
        for (int r = cta; r < p.B; r += G, ++rows)
          ch_norm_row(d.Xraw + (size_t)r * d.ldx, d.xn + (size_t)r * d.K, d.norm_w, d.K, d.eps, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          if (rows) ch_add_release(p.norm_done + g, rows);
          ch_wait_count(p.norm_done + g, p.B, "RMSNorm", g);
          ch_stamp(p, 8 + g * 8 + 1);                  // normalised rows complete
        }
        __syncwarp();
      }
      if (lane == 0) {
        // the activations were written through the generic proxy (other CTAs' epilogues / norm warps), TMA reads
        // them through the async proxy
        asm volatile("fence.proxy.async;" ::: "memory");
        SkSpan sp;
        sp.init(cta, G, d.num_tiles, d.num_kb);
        const int n_units = sp.num_units();
        SkCursor cur;
        cur.init(sp);
        for (int it = 0; it < n_units; ++it) {
          mbar_wait(&empty[s], (round & 1u) ^ 1u);
          mbar_expect_tx(&full_x[s], X_BYTES);
          tma_load_2d(smem + s * STAGE_BYTES + SK_W_BYTES, &d.map_x, &full_x[s], cur.kb * SK_BK, 0);
          cur.advance(sp);
          if (++s == ST) { s = 0; ++round; }
        }
      }
      __syncwarp();
    }
  } else {
    // ───────────── epilogue warps 2..5 ─────────────
    pdl_wait();                                     // residual / bias consumers; partial slots of earlier launches
    const int quad = warp & 3;
    const int et = quad * 32 + lane;             // TMEM lane = weight row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int CC = (BC < 16) ? BC : 16;
    int segc = 0;
    for (int g = 0; g < p.n_desc; ++g) {
      const ChainDesc &d = p.d[g];
      const int KB = d.num_kb;
      SkSpan sp;
      sp.init(cta, G, d.num_tiles, KB);
      const int n_segs = sp.num_units() > 0 ? sp.num_segs() : 0;
      int *flags = p.flags + g * SK_MAX_GRID;
      for (int seg = 0; seg < n_segs; ++seg, ++segc) {
        int tile, kb0, nkb;
        sp.seg(seg, tile, kb0, nkb);
        const int acc = segc & 1;
        const bool finishes = (kb0 + nkb == KB);
        const int n = tile * SK_BM + et;
        const bool n_ok = n < d.N;
        const bool use_res = finishes && d.epilogue == OCRB_EPI_RESIDUAL && n_ok;
        uint32_t rr_next[CC];
        auto load_res = [&](int c0, uint32_t (&rr)[CC]) {
#pragma unroll
          for (int i = 0; i < CC; ++i)
            rr[i] = (c0 + i < p.B) ? (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(d.residual + (size_t)(c0 + i) * d.ldr + n)) : 0u;
        };
        if (use_res && d.res_early) load_res(0, rr_next);
        const float bv = (finishes && d.bias && n_ok) ? __bfloat162float(d.bias[n]) : 0.f;   // cold HBM read: before the wait
        mbar_wait(&tmem_full[acc], (segc >> 1) & 1);
        tcgen05_fence_after();
        if (use_res && !d.res_early) load_res(0, rr_next);
        if (et == 0 && seg == n_segs - 1) ch_stamp(p, 8 + g * 8 + 3);     // last segment of linear g accumulated
        auto load_chunk = [&](int c0, float (&v)[CC]) {
          uint32_t r[CC];
          tmem_ld_cols<CC>(lane_addr + acc * ACC_STRIDE + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CC; ++i) v[i] = __uint_as_float(r[i]);
          if (c0 + CC >= BC) {                         // last TMEM read of this segment: the MMA warp may reuse the buffer
            tcgen05_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
        };
        if (!finishes) {
          // partial span: publish fp32 partials, then the flag
          float *slot = p.partials + (size_t)cta * BC * 128;
#pragma unroll 1
          for (int c0 = 0; c0 < BC; c0 += CC) {
            float v[CC];
            load_chunk(c0, v);
#pragma unroll
            for (int i = 0; i < CC; ++i) slot[(c0 + i) * 128 + et] = v[i];
          }
          __threadfence();
          named_bar_sync(1, 128);
          if (et == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + cta), "r"(1) : "memory");
        } else {
          int c_first = cta;                           // first contributing CTA (== cta: none)
          if (kb0 > 0) {
            // this CTA finishes a tile that earlier CTAs started: their partials are added in k order, then ours
            const int tile_first = tile * KB;
            const int total = d.num_tiles * KB;
            const int per = total / G, rem = total % G;
            const int big = rem * (per + 1);
            c_first = (tile_first < big) ? tile_first / (per + 1) : rem + (tile_first - big) / per;
            for (int cb = c_first; cb < cta; cb += 128) {
              if (cb + et < cta) {
                const int c = cb + et;
                int f;
                const long long t0 = clock64();
                do {
                  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(flags + c) : "memory");
                  if (!f && clock64() - t0 > 4000000000LL) {
                    printf("ocrb chain: partial of CTA %d for linear %d never arrived (CTA %d)\n", c, g, cta);
                    __trap();
                  }
                } while (!f);
              }
            }
            named_bar_sync(1, 128);
          }
#pragma unroll 1
          for (int c0 = 0; c0 < BC; c0 += CC) {
            float v[CC];
            uint32_t rr[CC];
#pragma unroll
            for (int i = 0; i < CC; ++i) rr[i] = rr_next[i];
            if (use_res && c0 + CC < BC) load_res(c0 + CC, rr_next);
            load_chunk(c0, v);
            if (c_first < cta) {
              // same summation order as skinny_gemm_kernel: contributors in k order, own accumulator last
              float sum[CC];
#pragma unroll
              for (int i = 0; i < CC; ++i) sum[i] = 0.f;
              constexpr int FX = (CC <= 4) ? 8 : ((CC <= 8) ? 4 : 2);   // contributors fetched together
              for (int cb = c_first; cb < cta; cb += FX) {
                const int nc = min(FX, cta - cb);
                float pv[FX][CC];
#pragma unroll
                for (int j = 0; j < FX; ++j) {
                  const float *slot = p.partials + (size_t)(cb + (j < nc ? j : 0)) * BC * 128;
#pragma unroll
                  for (int i = 0; i < CC; ++i) pv[j][i] = __ldcg(slot + (c0 + i) * 128 + et);
                }
#pragma unroll
                for (int j = 0; j < FX; ++j)
                  if (j < nc) {
#pragma unroll
                    for (int i = 0; i < CC; ++i) sum[i] += pv[j][i];
                  }
              }
#pragma unroll
              for (int i = 0; i < CC; ++i) v[i] = sum[i] + v[i];
            }
            // ───── epilogue math on the complete accumulator (HF rounding points), as in skinny_gemm_kernel ─────
            float o[CC];
#pragma unroll
            for (int i = 0; i < CC; ++i) o[i] = bf16_round(v[i] + bv);
            if (d.epilogue == OCRB_EPI_SWIGLU) {
              // tile rows 0..63 = gate, 64..127 = up of output columns tile*64 + j
              if (et >= 64) {
#pragma unroll
                for (int i = 0; i < CC; ++i) s_up[(et - 64) * CC + i] = o[i];
              }
              named_bar_sync(1, 128);
              if (et < 64) {
#pragma unroll
                for (int i = 0; i < CC; ++i) o[i] = sk_silu(o[i]) * s_up[et * CC + i];
                if (n_ok) {
                  bf16 *dcol = d.D + (size_t)c0 * d.ldd + (tile * 64 + et);
#pragma unroll
                  for (int i = 0; i < CC; ++i)
                    if (c0 + i < p.B) dcol[(size_t)i * d.ldd] = __float2bfloat16_rn(o[i]);
                }
              }
              named_bar_sync(1, 128);
            } else {
              if (d.epilogue == OCRB_EPI_RESIDUAL) {
#pragma unroll
                for (int i = 0; i < CC; ++i) o[i] += __uint_as_float(rr[i] << 16);
              } else if (d.epilogue == OCRB_EPI_GELU) {
#pragma unroll
                for (int i = 0; i < CC; ++i) o[i] = sk_gelu(o[i]);
              }
              if (n_ok) {
                bf16 *dcol = d.D + (size_t)c0 * d.ldd + n;
#pragma unroll
                for (int i = 0; i < CC; ++i)
                  if (c0 + i < p.B) dcol[(size_t)i * d.ldd] = __float2bfloat16_rn(o[i]);
              }
            }
          }
        }
      }
      // this CTA's part of linear g is stored (or published): announce it
      __threadfence();
      named_bar_sync(1, 128);
      if (et == 0) {
        ch_add_release(p.done + g, 1);
        ch_stamp(p, 8 + g * 8 + 4);
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  // The last CTA to leave returns counters and flags to zero (every other CTA has finished polling by then).
  if (threadIdx.x == 0) {
    __threadfence();
    *s_ticket = atomicAdd(p.exit_count, 1);
    ch_stamp(p, 63);
  }
  __syncthreads();
  if (*s_ticket == G - 1) {
    for (int i = threadIdx.x; i < p.n_desc * SK_MAX_GRID; i += CH_THREADS) p.flags[i] = 0;
    if (threadIdx.x < CH_MAXD) { p.done[threadIdx.x] = 0; p.norm_done[threadIdx.x] = 0; }
    if (threadIdx.x == 0) *p.exit_count = 0;
  }
}

static int ch_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BP, int BC>
static int launch_chain(const ChainParams &p, int grid, cudaStream_t st) {
  constexpr size_t smem = ChainCfg<BP>::SMEM;
  static_assert(smem <= 227 * 1024, "chain: shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(skinny_chain_kernel<BP, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  OCRB_CUDA(launch_pdl(skinny_chain_kernel<BP, BC>, dim3(grid), dim3(CH_THREADS), smem, st, p));
  return check_launch("skinny_chain_kernel");
}

}  // namespace ocrb

using namespace ocrb;

static unsigned long long *g_ch_trace = nullptr;
/* debug hook (not in the public header): device buffer [grid][64] of globaltimer stamps for the next launches */
extern "C" void ocrb_chain_set_trace(void *buf) { g_ch_trace = (unsigned long long *)buf; }

extern "C" int64_t ocrb_chain_workspace_bytes(void) {
  return (int64_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float)             /* partials */
         + (int64_t)(CH_MAXD * SK_MAX_GRID + 3 * 16) * sizeof(int)         /* flags + counters */
         + 256 + (int64_t)CH_MAXD * SK_MAXBP * 8192 * sizeof(bf16);        /* normalised rows */
}

extern "C" int ocrb_skinny_chain_bf16(const ocrb_chain_linear *lin, int32_t n, int32_t B, void *workspace, void *stream) {
  OCRB_REQUIRE(lin && workspace, "skinny_chain_bf16: null pointer");
  OCRB_REQUIRE(n >= 1 && n <= CH_MAXD, "skinny_chain_bf16: 1..%d linears per chain", CH_MAXD);
  OCRB_REQUIRE(B >= 1 && B <= SK_MAXBP, "skinny_chain_bf16: B must be in 1..128");
  const int grid = ch_sm_count();
  OCRB_REQUIRE(grid <= SK_MAX_GRID, "skinny_chain_bf16: more SMs than workspace slots");
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.n_desc = n;
  p.B = B;
  char *ws = (char *)workspace;
  p.partials = (float *)ws;
  ws += (size_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float);
  p.flags = (int *)ws;
  p.done = p.flags + CH_MAXD * SK_MAX_GRID;
  p.norm_done = p.done + 16;
  p.exit_count = p.norm_done + 16;
  ws += (size_t)(CH_MAXD * SK_MAX_GRID + 3 * 16) * sizeof(int) + 256;
  ws = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  bf16 *xn_base = (bf16 *)ws;
  p.trace = g_ch_trace;
  const int BPsel = B <= 16 ? 16 : (B <= 32 ? 32 : (B <= 64 ? 64 : (B <= 96 ? 96 : 128)));
  for (int g = 0; g < n; ++g) {
    const ocrb_chain_linear &l = lin[g];
    ChainDesc &d = p.d[g];
    OCRB_REQUIRE(l.X && l.W && l.D, "skinny_chain_bf16: null pointer in linear %d", g);
    OCRB_REQUIRE(l.N > 0 && l.K > 0 && l.K % 8 == 0 && l.ldx % 8 == 0 && l.ldw % 8 == 0,
                 "skinny_chain_bf16: K and row strides must be multiples of 8 (linear %d)", g);
    OCRB_REQUIRE(((uintptr_t)l.X & 15) == 0 && ((uintptr_t)l.W & 15) == 0 && (!l.norm_w || ((uintptr_t)l.norm_w & 15) == 0),
                 "skinny_chain_bf16: X, W, norm_w must be 16-byte aligned (linear %d)", g);
    OCRB_REQUIRE(l.epilogue >= 0 && l.epilogue <= 3, "skinny_chain_bf16: bad epilogue (linear %d)", g);
    OCRB_REQUIRE(l.epilogue != OCRB_EPI_RESIDUAL || l.residual, "skinny_chain_bf16: residual epilogue without residual");
    OCRB_REQUIRE(l.epilogue != OCRB_EPI_SWIGLU || l.N % 128 == 0, "skinny_chain_bf16: SwiGLU needs packed N %% 128 == 0");
    OCRB_REQUIRE(!l.norm_w || l.K <= 8192, "skinny_chain_bf16: the RMSNorm prologue supports K <= 8192");
    d.N = l.N;
    d.K = l.K;
    d.num_tiles = cdiv(l.N, SK_BM);
    d.num_kb = cdiv(l.K, SK_BK);
    OCRB_REQUIRE((long long)d.num_tiles * d.num_kb < (1ll << 30), "skinny_chain_bf16: problem too large");
    d.epilogue = l.epilogue;
    d.bias = (const bf16 *)l.bias;
    d.residual = (l.epilogue == OCRB_EPI_RESIDUAL) ? (const bf16 *)l.residual : nullptr;
    d.ldr = l.ldr;
    d.D = (bf16 *)l.D;
    d.ldd = l.ldd;
    d.eps = l.eps;
    // the residual may be fetched ahead unless the linear right before writes it
    d.res_early = (g == 0 || lin[g - 1].D != l.residual) ? 1 : 0;
    int rc = make_tensor_map_bf16(&d.map_w, l.W, l.N, l.K, l.ldw, SK_BM);
    if (rc) return rc;
    const void *xsrc = l.X;
    long long ldx = l.ldx;
    if (l.norm_w) {
      d.norm_w = (const bf16 *)l.norm_w;
      d.Xraw = (const bf16 *)l.X;
      d.ldx = l.ldx;
      d.xn = xn_base + (size_t)g * SK_MAXBP * 8192;
      xsrc = d.xn;
      ldx = l.K;
    }
    rc = make_tensor_map_bf16(&d.map_x, xsrc, B, l.K, ldx, BPsel);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 4) return launch_chain<16, 4>(p, grid, st);
  if (B <= 8) return launch_chain<16, 8>(p, grid, st);
  if (B <= 16) return launch_chain<16, 16>(p, grid, st);
  if (B <= 32) return launch_chain<32, 32>(p, grid, st);
  if (B <= 64) return launch_chain<64, 64>(p, grid, st);
  if (B <= 96) return launch_chain<96, 96>(p, grid, st);
  return launch_chain<128, 128>(p, grid, st);
}
