// One decode step's dependent work in ONE persistent launch (HF generation loop, utils.py:2743-2806; decoder layer
// modeling_qwen2_5_vl.py:839-879):  RMSNorm + qkv -> paged attention (+ mRoPE, KV append, split combine) -> o_proj +
// residual -> RMSNorm -> gate/up + SwiGLU -> down_proj + residual -> ... -> final norm + lm_head, as a list of ops
// ("plan") walked by every CTA.  Same arithmetic, rounding points and splits as the one-launch-per-op kernels
// (skinny_gemm_kernel in skinny.cu, decode_attn_kernel / decode_attn_combine_kernel in decode.cu), so an op produces the
// bits it produces there; what changes is the schedule:
//   * HBM never drains between ops.  Everything the step STREAMS -- weight tiles and cached K / V tiles -- is
//     independent of the step's activations, so one producer thread walks the whole plan and runs AHEAD of the
//     dependencies through a shared-memory ring as deep as 227 KiB allow (11 x 18 KiB at B <= 16), and a prefetch
//     thread runs further ahead still, pulling the next tiles into L2 (cp.async.bulk.prefetch.tensor) while the
//     ring is full: the bubbles between dependent ops (stream-K fix-up, RMSNorm, attention) are spent streaming;
//   * dependencies are device-side: every CTA bumps a counter when its part of op g is stored, the consumer of op
//     g+1 polls it (ld.acquire.gpu) before it touches the activations;
//   * an RMSNorm in front of a linear is done in the plan: row r is normalised by one warp of CTA r (same routine and
//     summation order as skinny_norm_rows_kernel), announced through a second counter;
//   * attention: the K / V tiles of a work item (sequence, KV head, 128-key range -- the items of decode_attn_kernel)
//     travel through the same ring as units of 32 keys; the four epilogue warps are the flash-decoding workers
//     (one item each at a time, mma.sync m16n8k16, Q fragments built in registers with the mRoPE applied), then
//     fold the key ranges of their share of (sequence, head) rows;
//   * counters are returned to zero by the last CTA to leave and stream-K flags carry a launch epoch, so a CUDA graph
//     can replay the launch.
//
// Warp roles (256 threads): 0 = ring producer (TMA: weights, K / V), 1 = TMEM alloc + MMA issuer, 2..5 = epilogue /
// attention workers (TMEM lane quadrant = warp % 4), 6 = activation TMA producer + RMSNorm rows + attention schedule,
// 7 = L2 prefetcher.
#include "skinny_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>

namespace ocrb {

constexpr int CH_MAXD = OCRB_CHAIN_MAX;       // linears per parameter-space chain (ocrb_skinny_chain_bf16)
constexpr int CH_MAX_OPS = OCRB_CHAIN_PLAN_MAX_OPS;
constexpr int CH_THREADS = 256;
constexpr int CH_KIND_LINEAR = 0, CH_KIND_ATTN = 1;
constexpr int CH_XN_SLOTS = 4;                // rotating buffers of normalised rows
constexpr int CH_XN_MAX_K = 8192;
constexpr int CH_MAX_ITEMS = 96;              // attention work items per CTA
constexpr int CH_MAX_UNITS = 384;             // 32-key ring units per CTA and attention op
constexpr int CH_HD = 128;                    // head dim of the fused attention
constexpr uint32_t CH_TILE_BYTES = 2 * 2 * 16 * 128;   // one 16-key tile: K atoms (2 x 2 KiB) then V atoms

struct alignas(64) ChainDesc {
  CUtensorMap map_w;                 // linear: [N, K] weights, box [128 x 64];  attention: the layer's K cache, box [16 x 64]
  CUtensorMap map_x;                 // linear: [B, K] activations (normalised copy if norm_w), box [BP x 64];  attention: V cache
  int kind, pad0;
  // ---- linear ----
  const bf16 *Xraw; long long ldx;   // rows to normalise (norm_w != nullptr)
  bf16 *xn;                          // normalised rows [B][K]
  const bf16 *norm_w; float eps;
  bf16 *D; long long ldd;
  const bf16 *bias;
  const bf16 *residual; long long ldr;
  int N, K, num_tiles, num_kb, epilogue;
  int res_early;                     // the residual does not come from the op right before: fetch it ahead
  const bf16 *W; long long ldw;      // raw weights (L2 prefetcher)
  // ---- attention (decode.cu semantics) ----
  const bf16 *qkv; long long ldqkv;
  bf16 *k_cache, *v_cache;
  const int32_t *block_table; int max_pages;
  const int32_t *ctx_len;
  int page_size, n_q, n_kv, n_splits, chunk;
  const bf16 *cosT, *sinT;
  float scale;
  float *split_ws;
  bf16 *att; long long ldo;
};

struct ChainCtl {
  int n_desc, B;
  float *partials;                   // [grid][BC][128] fp32 stream-K partials (one slot per CTA, reused along the plan)
  int *flags;                        // [CH_MAX_OPS][SK_MAX_GRID] partial-ready flags (value = launch epoch + 1)
  int *done;                         // [CH_MAX_OPS] CTAs whose part of op g is stored
  int *stage2;                       // [CH_MAX_OPS] linear: rows normalised; attention: CTAs whose combined rows are stored
  int *exit_count;
  int *epoch;
  int pf_ahead;                      // ring units the L2 prefetcher stays ahead of the ring producer (0: off)
  unsigned long long *trace;         // optional [grid][trace_slots] globaltimer stamps
  int trace_slots;
};

struct ChainParams {                 // parameter-space variant: a short chain of linears
  ChainDesc d[CH_MAXD];
  ChainCtl ctl;
};

struct AttnItem { int b, kvh, k0, nkeys, split; };
struct AttnUnit { int row0, row1; unsigned short item; unsigned char t, ntiles; };

__device__ __forceinline__ void ch_stamp(const ChainCtl &c, int slot) {
  if (c.trace && slot < c.trace_slots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    c.trace[(size_t)blockIdx.x * c.trace_slots + slot] = t;
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
      printf("ocrb chain: %s of op %d stuck at %d of %d (CTA %d)\n", what, g, v, target, blockIdx.x);
      __trap();
    }
  } while (v < target);
}
__device__ __forceinline__ void ch_add_release(int *ctr, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
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

// ───────────── attention helpers (arithmetic of decode.cu, inputs read through L2: they were written in this launch) ─────────────
__device__ __forceinline__ float ch_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float ch_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t ch_pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
// mRoPE of two adjacent elements of a head vector, HF's bf16 rounding points (decode.cu rope_elem_bf16): x = the
// elements, o = their rotate-half partners (sgn = -1 for the first half of the vector), cw / sw = cos / sin
__device__ __forceinline__ uint32_t ch_rope_math(uint32_t x, uint32_t o, uint32_t cw, uint32_t sw, float sgn) {
  const float t0 = bf16_round(ch_bf16_lo(x) * ch_bf16_lo(cw)), t1 = bf16_round(ch_bf16_hi(x) * ch_bf16_hi(cw));
  const float u0 = bf16_round(sgn * ch_bf16_lo(o) * ch_bf16_lo(sw)), u1 = bf16_round(sgn * ch_bf16_hi(o) * ch_bf16_hi(sw));
  return ch_pack_bf16(bf16_round(t0 + u0), bf16_round(t1 + u1));
}
__device__ __forceinline__ uint32_t ch_rope_pair(const bf16 *vec, int i, const bf16 *c, const bf16 *s) {
  constexpr int half = CH_HD / 2;
  const uint32_t x = __ldcg(reinterpret_cast<const unsigned int *>(vec + i));
  const uint32_t o = __ldcg(reinterpret_cast<const unsigned int *>(vec + (i < half ? i + half : i - half)));
  const uint32_t cw = *reinterpret_cast<const unsigned int *>(c + i);
  const uint32_t sw = *reinterpret_cast<const unsigned int *>(s + i);
  return ch_rope_math(x, o, cw, sw, (i < half) ? -1.f : 1.f);
}
__device__ __forceinline__ void ch_ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ch_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ uint32_t ch_movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void ch_mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int BP>
struct ChainCfg {
  // ring depth: as deep as 227 KiB allow next to the barriers, the SwiGLU exchange buffer and the attention schedule
  static constexpr int ST = (BP <= 16) ? 11 : ((BP <= 32) ? 10 : ((BP <= 64) ? 8 : ((BP <= 96) ? 7 : 6)));
  static constexpr uint32_t X_BYTES = BP * SK_BK * 2;
  static constexpr uint32_t STAGE_BYTES = SK_W_BYTES + X_BYTES;
  static constexpr int ACC_STRIDE = (BP <= 16) ? 16 : (BP <= 32 ? 32 : (BP <= 64 ? 64 : 128));
  static constexpr int TMEM_COLS = (2 * ACC_STRIDE < 32) ? 32 : 2 * ACC_STRIDE;
  static constexpr uint32_t OFF_BARS = ST * STAGE_BYTES;
  static constexpr uint32_t OFF_SUP = OFF_BARS + 512;                                  // [16][128] fp32 SwiGLU exchange
  static constexpr uint32_t OFF_ITEMS = OFF_SUP + 128 * 16 * sizeof(float);
  static constexpr uint32_t OFF_UNITS = OFF_ITEMS + CH_MAX_ITEMS * sizeof(AttnItem);
  static constexpr uint32_t OFF_GSTART = OFF_UNITS + CH_MAX_UNITS * sizeof(AttnUnit);
  static constexpr uint32_t OFF_END = OFF_GSTART + (CH_MAX_ITEMS / 4 + 2) * sizeof(int);
  static constexpr size_t SMEM = (size_t)OFF_END + 1024 /*align*/ + 64;
};

template <int BP, int BC>
__device__ __forceinline__ void chain_body(const ChainDesc *__restrict__ descs, const ChainCtl &ctl) {
  using C = ChainCfg<BP>;
  constexpr int ST = C::ST;
  constexpr uint32_t X_BYTES = C::X_BYTES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int ACC_STRIDE = C::ACC_STRIDE, TMEM_COLS = C::TMEM_COLS;
  extern __shared__ uint8_t ch_smem_raw[];
  uint8_t *smem = ch_smem_raw + ((1024u - (smem_u32(ch_smem_raw) & 1023u)) & 1023u);
  uint64_t *full_w = reinterpret_cast<uint64_t *>(smem + C::OFF_BARS);
  uint64_t *full_x = full_w + ST;
  uint64_t *empty = full_x + ST;
  uint64_t *tmem_full = empty + ST;            // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint64_t *tab_bar = tmem_empty + 2;          // attention schedule built
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tab_bar + 1);
  int *s_ticket = reinterpret_cast<int *>(tmem_slot + 1);
  volatile int *s_ld_pos = reinterpret_cast<volatile int *>(tmem_slot + 2);   // ring units issued so far (for the prefetcher)
  int *s_tab_n = reinterpret_cast<int *>(tmem_slot + 3);                      // [0] live items, [1] units, [2] groups
  float *s_up = reinterpret_cast<float *>(smem + C::OFF_SUP);
  AttnItem *s_items = reinterpret_cast<AttnItem *>(smem + C::OFF_ITEMS);
  AttnUnit *s_units = reinterpret_cast<AttnUnit *>(smem + C::OFF_UNITS);
  int *s_gstart = reinterpret_cast<int *>(smem + C::OFF_GSTART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_desc = ctl.n_desc;
  if (threadIdx.x == 0) ch_stamp(ctl, 0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    mbar_init(tab_bar, 1);
    *s_ld_pos = 0;
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
  if (threadIdx.x == 0) { ch_stamp(ctl, 1); pdl_launch_dependents(); }

  if (warp == 0) {
    // ───────────── ring producer: weights and K / V tiles of the whole plan, never waits for a dependency ─────────────
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int pos = 0;                               // ring units issued so far: slot pos % ST, use pos / ST
      bool tab_ready = false;
      for (int g = 0; g < n_desc; ++g) {
        const ChainDesc &d = descs[g];
        if (d.kind == CH_KIND_LINEAR) {
          SkSpan sp;
          sp.init(cta, G, d.num_tiles, d.num_kb);
          const int n_units = sp.num_units();
          SkCursor cur;
          cur.init(sp);
          for (int it = 0; it < n_units; ++it, ++pos) {
            const int s = pos % ST;
            mbar_wait(&empty[s], ((uint32_t)(pos / ST) & 1u) ^ 1u);
            mbar_expect_tx(&full_w[s], SK_W_BYTES);
            tma_load_2d_hint(smem + s * STAGE_BYTES, &d.map_w, &full_w[s], cur.kb * SK_BK, cur.tile * SK_BM, policy);
            cur.advance(sp);
            *s_ld_pos = pos + 1;
          }
        } else {
          if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
          const int n_units = s_tab_n[1];
          for (int u = 0; u < n_units; ++u, ++pos) {
            const int s = pos % ST;
            const AttnUnit e = s_units[u];
            mbar_wait(&empty[s], ((uint32_t)(pos / ST) & 1u) ^ 1u);
            mbar_expect_tx(&full_w[s], (uint32_t)e.ntiles * CH_TILE_BYTES);
            uint8_t *dst = smem + s * STAGE_BYTES;
            for (int i = 0; i < e.ntiles; ++i) {
              const int row = i ? e.row1 : e.row0;
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                tma_load_2d(dst + i * CH_TILE_BYTES + a * 2048, &d.map_w, &full_w[s], a * 64, row);
                tma_load_2d(dst + i * CH_TILE_BYTES + 4096 + a * 2048, &d.map_x, &full_w[s], a * 64, row);
              }
            }
            *s_ld_pos = pos + 1;
          }
        }
      }
    }
  } else if (warp == 7) {
    // ───────────── L2 prefetcher: the same unit stream, up to pf_ahead units beyond the ring producer ─────────────
    // Whole warp, LSU path (prefetch.global.L2 of the 128 row segments of a weight tile): TMA prefetches would queue in
    // front of the dependency-critical activation loads of the same SM.
    if (ctl.pf_ahead > 0) {
      int pos = 0;
      bool tab_ready = false;
      const int ahead = ctl.pf_ahead;
      // Prefetch only while the ring producer is STALLED (ring full, its consumer waiting on a dependency): while it
      // streams, HBM is busy anyway.  "Stalled" = no unit issued for ~1 us (a unit takes ~0.4 us when streaming).
      int seen = 0;
      long long t_seen = clock64();
      auto throttle = [&](int pos_) -> bool {      // false: the producer has passed this unit, nothing to prefetch
        int go = 0;
        if (lane == 0) {
          const long long t0 = clock64();
          for (;;) {
            const int cur = *s_ld_pos;
            const long long now = clock64();
            if (cur != seen) { seen = cur; t_seen = now; }
            if (pos_ < cur) { go = 0; break; }
            if (pos_ < cur + ahead && now - t_seen > 2000) { go = 1; break; }
            if (now - t0 > 8000000000LL) { go = 0; break; }        // the ring is stuck: its own timeouts will report it
            __nanosleep(100);
          }
        }
        return __shfl_sync(0xffffffffu, go, 0) != 0;
      };
      for (int g = 0; g < n_desc; ++g) {
        const ChainDesc &d = descs[g];
        if (d.kind == CH_KIND_LINEAR) {
          SkSpan sp;
          sp.init(cta, G, d.num_tiles, d.num_kb);
          const int n_units = sp.num_units();
          SkCursor cur;
          cur.init(sp);
          for (int it = 0; it < n_units; ++it, ++pos) {
            if (throttle(pos)) {
              const char *base = reinterpret_cast<const char *>(d.W + (size_t)cur.tile * SK_BM * d.ldw + (size_t)cur.kb * SK_BK);
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int rowi = lane + r * 32;
                if (cur.tile * SK_BM + rowi < d.N)
                  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (size_t)rowi * d.ldw * 2));
              }
            }
            cur.advance(sp);
          }
        } else {
          if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
          pos += s_tab_n[1];                   // K / V tiles are not prefetched (small next to the weights at the batch sizes this helps)
        }
      }
    }
  } else if (warp == 1) {
    // ───────────── MMA issuer ─────────────
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BP >> 3) << 17) | ((uint32_t)(SK_BM >> 4) << 24);
      int pos = 0;
      uint32_t xphase = 0;                       // bit s = parity of the next completion of full_x[s] (attention units skip it)
      int segc = 0;                              // segments so far along the plan (TMEM buffer = segc & 1)
      bool tab_ready = false;
      for (int g = 0; g < n_desc; ++g) {
        const ChainDesc &d = descs[g];
        if (d.kind != CH_KIND_LINEAR) {
          if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
          pos += s_tab_n[1];
          continue;
        }
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
          for (int i = 0; i < nkb; ++i, ++pos) {
            const int s = pos % ST;
            mbar_wait(&full_w[s], (uint32_t)(pos / ST) & 1u);
            mbar_wait(&full_x[s], (xphase >> s) & 1u);
            xphase ^= 1u << s;
            if (seg == 0 && i == 0) ch_stamp(ctl, 8 + g * 8 + 2);      // first k-block of op g ready
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + SK_W_BYTES);
#pragma unroll
            for (int k = 0; k < SK_BK / UMMA_K; ++k)
              umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          umma_commit(&tmem_full[acc]);
        }
      }
    }
  } else if (warp == 6) {
    // ───────────── activation producer: attention schedule, dependencies, RMSNorm rows, activation TMA ─────────────
    pdl_wait();                                     // the first op's input (and ctx_len) come from preceding kernels
    // ---- attention schedule of this CTA (the same for every attention op of the plan: only the caches differ) ----
    {
      int ga = -1;
      for (int g = 0; g < n_desc; ++g)
        if (descs[g].kind == CH_KIND_ATTN) { ga = g; break; }
      if (ga >= 0) {
        const ChainDesc &d = descs[ga];
        const int pairs = ctl.B * d.n_kv;
        const int n_items = pairs * d.n_splits;
        // live items in item order (item = cta + G * j): (key range, sequence, kv head), as decode_attn_kernel
        int n_live = 0;
        for (int j0 = 0; cta + G * j0 < n_items; j0 += 32) {
          const int it = cta + G * (j0 + lane);
          bool live = false;
          AttnItem a = {0, 0, 0, 0, 0};
          if (it < n_items) {
            a.split = it / pairs;
            const int pair = it - a.split * pairs;
            a.b = pair / d.n_kv;
            a.kvh = pair - a.b * d.n_kv;
            const int total = d.ctx_len[a.b] + 1;
            a.k0 = a.split * d.chunk;
            a.nkeys = min(total, a.k0 + d.chunk) - a.k0;
            live = a.nkeys > 0;
          }
          const unsigned m = __ballot_sync(0xffffffffu, live);
          const int slot = n_live + __popc(m & ((1u << lane) - 1u));
          if (live && slot < CH_MAX_ITEMS) s_items[slot] = a;
          n_live += __popc(m);
        }
        if (n_live > CH_MAX_ITEMS) {
          if (lane == 0) printf("ocrb chain: %d attention items on CTA %d exceed the schedule (%d)\n", n_live, cta, CH_MAX_ITEMS);
          __trap();
        }
        __syncwarp();
        // ring order: groups of four items (one per worker warp), their 32-key units interleaved
        int n_units = 0, n_groups = 0;
        if (lane == 0) {
          for (int q = 0; q * 4 < n_live; ++q) {
            s_gstart[q] = n_units;
            const int chunk_units = (d.chunk + 31) / 32;
            for (int t = 0; t < chunk_units; ++t)
              for (int w = 0; w < 4; ++w) {
                const int jl = q * 4 + w;
                if (jl < n_live && t * 32 < s_items[jl].nkeys) {
                  if (n_units < CH_MAX_UNITS) {
                    AttnUnit e;
                    e.item = (unsigned short)jl;
                    e.t = (unsigned char)t;
                    e.ntiles = (unsigned char)((s_items[jl].nkeys - t * 32 > 16) ? 2 : 1);
                    e.row0 = e.row1 = 0;
                    s_units[n_units] = e;
                  }
                  ++n_units;
                }
              }
            n_groups = q + 1;
          }
          s_gstart[n_groups] = n_units;
          s_tab_n[0] = n_live;
          s_tab_n[1] = n_units;
          s_tab_n[2] = n_groups;
        }
        n_units = __shfl_sync(0xffffffffu, n_units, 0);
        if (n_units > CH_MAX_UNITS) {
          if (lane == 0) printf("ocrb chain: %d attention units on CTA %d exceed the schedule (%d)\n", n_units, cta, CH_MAX_UNITS);
          __trap();
        }
        __syncwarp();
        for (int u = lane; u < n_units; u += 32) {   // cache rows of the unit's tiles (block-table lookups in parallel)
          AttnUnit e = s_units[u];
          const AttnItem a = s_items[e.item];
          const int32_t *bt = d.block_table + (size_t)a.b * d.max_pages;
          const int key0 = a.k0 + e.t * 32;
          e.row0 = (bt[key0 / d.page_size] * d.n_kv + a.kvh) * d.page_size + key0 % d.page_size;
          if (e.ntiles > 1) {
            const int key1 = key0 + 16;
            e.row1 = (bt[key1 / d.page_size] * d.n_kv + a.kvh) * d.page_size + key1 % d.page_size;
          }
          s_units[u] = e;
        }
        __syncwarp();
      } else if (lane == 0) {
        s_tab_n[0] = s_tab_n[1] = s_tab_n[2] = 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tab_bar);
    }
    int pos = 0;
    for (int g = 0; g < n_desc; ++g) {
      const ChainDesc &d = descs[g];
      if (d.kind != CH_KIND_LINEAR) {
        pos += s_tab_n[1];
        continue;
      }
      if (d.norm_w && cta < ctl.B) {
        // HBM is saturated by the weight stream when the rows become ready: a cold read of the norm weights then waits
        // microseconds in the DRAM queues.  Pull them into L2 before the dependency wait.
        for (int off = lane * 128; off < d.K * 2; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(d.norm_w) + off));
      }
      if (g > 0) {
        if (lane == 0)
          ch_wait_count((descs[g - 1].kind == CH_KIND_LINEAR ? ctl.done : ctl.stage2) + (g - 1), G, "completion", g - 1);
        __syncwarp();
      }
      if (lane == 0) ch_stamp(ctl, 8 + g * 8 + 0);       // inputs of op g complete
      if (d.norm_w) {
        int rows = 0;
        for (int r = cta; r < ctl.B; r += G, ++rows)
          ch_norm_row(d.Xraw + (size_t)r * d.ldx, d.xn + (size_t)r * d.K, d.norm_w, d.K, d.eps, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          if (rows) ch_add_release(ctl.stage2 + g, rows);
          ch_wait_count(ctl.stage2 + g, ctl.B, "RMSNorm", g);
          ch_stamp(ctl, 8 + g * 8 + 1);                  // normalised rows complete
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
        for (int it = 0; it < n_units; ++it, ++pos) {
          const int s = pos % ST;
          mbar_wait(&empty[s], ((uint32_t)(pos / ST) & 1u) ^ 1u);
          mbar_expect_tx(&full_x[s], X_BYTES);
          tma_load_2d(smem + s * STAGE_BYTES + SK_W_BYTES, &d.map_x, &full_x[s], cur.kb * SK_BK, 0);
          cur.advance(sp);
        }
      } else {
        SkSpan sp;
        sp.init(cta, G, d.num_tiles, d.num_kb);
        pos += sp.num_units();
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp <= 5) {
    // ───────────── epilogue warps 2..5 (attention workers in attention ops) ─────────────
    pdl_wait();                                     // residual / bias consumers; partial slots of earlier launches
    const int quad = warp & 3;
    const int et = quad * 32 + lane;             // TMEM lane = weight row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int CC = (BC < 16) ? BC : 16;
    const int E1 = *reinterpret_cast<volatile int *>(ctl.epoch) + 1;      // this launch's flag value
    int segc = 0;
    int pos = 0;
    bool tab_ready = false;
    for (int g = 0; g < n_desc; ++g) {
      const ChainDesc &d = descs[g];
      if (d.kind == CH_KIND_ATTN) {
        // ═════════════ paged decode attention: this warp is a flash-decoding worker (decode.cu decode_attn_kernel) ═════════════
        if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
        const int n_live = s_tab_n[0], n_units = s_tab_n[1], n_groups = s_tab_n[2];
        const int Gq = d.n_q / d.n_kv;
        const int m = lane >> 3, l8 = lane & 7;
        const int r0 = lane >> 2, cq = (lane & 3) * 2;
        bool waited = false;
        for (int q = 0; q < n_groups; ++q) {
          const int jl = q * 4 + quad;
          if (jl >= n_live) continue;
          const AttnItem a = s_items[jl];
          if (!waited) {                            // qkv of this step comes from the linear right before
            if (g > 0) {
              if (lane == 0) ch_wait_count((descs[g - 1].kind == CH_KIND_LINEAR ? ctl.done : ctl.stage2) + (g - 1), G, "completion", g - 1);
              __syncwarp();
            }
            waited = true;
            if (et == 0) ch_stamp(ctl, 8 + g * 8 + 0);       // qkv of the step complete
          }
          const int ctx = d.ctx_len[a.b];
          const int total = ctx + 1;
          const bf16 *row = d.qkv + (size_t)a.b * d.ldqkv;
          const bf16 *c = d.cosT + (size_t)a.b * CH_HD, *s_ = d.sinT + (size_t)a.b * CH_HD;
          // Q^T operand fragments (lane r0 = query head, heads >= Gq zero), mRoPE applied.
          // Element j of a lane = columns (j >> 1) * 16 + (j & 1) * 8 + cq, +1; its rotate-half partner (column +- 64) is
          // element j ^ 8 of the same lane, so one batch of independent loads (q rows, cos, sin) feeds the whole tile.
          uint32_t qf[CH_HD / 16][4];
          {
            const bf16 *q0 = row + (size_t)(a.kvh * Gq + r0) * CH_HD;
            const bool v0 = r0 < Gq;
            uint32_t x0[16], cw[16], sw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = (j >> 1) * 16 + (j & 1) * 8 + cq;
              x0[j] = v0 ? __ldcg(reinterpret_cast<const unsigned int *>(q0 + col)) : 0u;
              cw[j] = *reinterpret_cast<const unsigned int *>(c + col);
              sw[j] = *reinterpret_cast<const unsigned int *>(s_ + col);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float sgn = (j < 8) ? -1.f : 1.f;
              qf[j >> 1][(j & 1) * 2] = ch_rope_math(x0[j], x0[j ^ 8], cw[j], sw[j], sgn);
              qf[j >> 1][(j & 1) * 2 + 1] = 0u;              // rows 8-15 of the old head-major tile: unused (Gq <= 8)
            }
          }
          if (et == 0 && q == 0) ch_stamp(ctl, 8 + g * 8 + 1);   // Q fragments built
          // transposed products (decode.cu): S^T[key][head] = K . Q^T, O^T[dim][head] += V^T . P^T; heads on the N = 8 side
          constexpr int KK = CH_HD / 16;
          float o[KK][4];
#pragma unroll
          for (int mt = 0; mt < KK; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
          float run_m[2] = {-INFINITY, -INFINITY}, run_l[2] = {0.f, 0.f};     // heads cq, cq + 1

          for (int u = s_gstart[q]; u < s_gstart[q + 1]; ++u) {
            const AttnUnit e = s_units[u];
            if (e.item != jl) continue;
            const int upos = pos + u;
            const int s = upos % ST;
            mbar_wait(&full_w[s], (uint32_t)(upos / ST) & 1u);
            if (et == 0 && q == 0 && e.t == 0) ch_stamp(ctl, 8 + g * 8 + 2);   // first K / V unit in shared memory
            uint8_t *slot = smem + s * STAGE_BYTES;
            bool patched = false;
            for (int ti = 0; ti < e.ntiles; ++ti) {
              const int key0 = a.k0 + e.t * 32 + ti * 16;
              uint8_t *tk = slot + ti * CH_TILE_BYTES, *tv = tk + 4096;
              const uint32_t sk = smem_u32(tk), sv = sk + 4096;
              if (ctx >= key0 && ctx < key0 + 16) {
                // this tile holds the new token: rope its key, place k / v in the (swizzled) tile and append them to the cache
                const bf16 *knew = row + (size_t)d.n_q * CH_HD + (size_t)a.kvh * CH_HD;
                const bf16 *vnew = row + (size_t)(d.n_q + d.n_kv) * CH_HD + (size_t)a.kvh * CH_HD;
                const int r = ctx - key0;
                const int32_t *bt = d.block_table + (size_t)a.b * d.max_pages;
                const size_t dst = ((size_t)(bt[ctx / d.page_size] * d.n_kv + a.kvh) * d.page_size + ctx % d.page_size) * CH_HD;
                for (int dd = lane * 2; dd < CH_HD; dd += 64) {
                  const uint32_t kr = ch_rope_pair(knew, dd, c, s_);
                  const uint32_t vr = __ldcg(reinterpret_cast<const unsigned int *>(vnew + dd));
                  const uint32_t off = (uint32_t)(dd >> 6) * 2048 + r * 128 + ((((dd & 63) >> 3) ^ (r & 7)) << 4) + (dd & 7) * 2;
                  *reinterpret_cast<uint32_t *>(tk + off) = kr;
                  *reinterpret_cast<uint32_t *>(tv + off) = vr;
                  *reinterpret_cast<uint32_t *>(d.k_cache + dst + dd) = kr;
                  *reinterpret_cast<uint32_t *>(d.v_cache + dst + dd) = vr;
                }
                patched = true;
                __syncwarp();
              }
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
              {
                const int kr = (m >> 1) * 8 + l8;                     // key row this lane addresses
#pragma unroll
                for (int kk = 0; kk < KK; ++kk) {
                  uint32_t bb[4];
                  const int ch = (kk & 3) * 2 + (m & 1);              // 16-byte chunk inside the 64-dim atom
                  ch_ldmatrix_x4(bb, sk + (kk >> 2) * 2048 + kr * 128 + ((ch ^ (kr & 7)) << 4));
                  const uint32_t ka[4] = {bb[0], bb[2], bb[1], bb[3]};
                  ch_mma_16816(acc, ka, qf[kk][0], qf[kk][2]);
                }
              }
              const bool ok0 = key0 + r0 < total, ok1 = key0 + r0 + 8 < total;
              float corr[2];
              bool moved = false;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                acc[j] = ok0 ? acc[j] * d.scale : -INFINITY;
                acc[2 + j] = ok1 ? acc[2 + j] * d.scale : -INFINITY;
                float tmax = fmaxf(acc[j], acc[2 + j]);
                tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 4));
                tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 8));
                tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 16));
                const float mn = fmaxf(run_m[j], tmax);               // finite: every processed tile has a live key
                corr[j] = __expf(run_m[j] - mn);                      // exp(-inf) = 0 on the first tile
                run_m[j] = mn;
                run_l[j] *= corr[j];
                moved |= corr[j] != 1.0f;
              }
              float pv[4];
#pragma unroll
              for (int ee = 0; ee < 4; ++ee) {
                pv[ee] = __expf(acc[ee] - run_m[ee & 1]);
                run_l[ee & 1] += pv[ee];
              }
              const uint32_t pb0 = ch_movmatrix_trans(ch_pack_bf16(pv[0], pv[1]));
              const uint32_t pb1 = ch_movmatrix_trans(ch_pack_bf16(pv[2], pv[3]));
              {
                const bool rescale = __any_sync(0xffffffffu, moved);
                const int vr = (m & 1) * 8 + l8;                      // key row this lane addresses
#pragma unroll
                for (int jj = 0; jj < KK; ++jj) {
                  uint32_t bb[4];
                  const int ch = (jj & 3) * 2 + (m >> 1);
                  ch_ldmatrix_x4_trans(bb, sv + (jj >> 2) * 2048 + vr * 128 + ((ch ^ (vr & 7)) << 4));
                  const uint32_t va[4] = {bb[0], bb[2], bb[1], bb[3]};
                  float (&oo)[4] = o[jj];
                  if (rescale) { oo[0] *= corr[0]; oo[1] *= corr[1]; oo[2] *= corr[0]; oo[3] *= corr[1]; }
                  ch_mma_16816(oo, va, pb0, pb1);
                }
              }
            }
            if (patched) fence_proxy_async_smem();      // generic writes to a slot the TMA unit will overwrite
            __syncwarp();                               // every lane is done with this ring slot
            if (lane == 0) mbar_arrive(&empty[s]);
          }
          // ---- one fp32 partial (max, sum, o[hd]) per head of this item, straight from the registers ----
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            run_l[j] += __shfl_xor_sync(0xffffffffu, run_l[j], 4);
            run_l[j] += __shfl_xor_sync(0xffffffffu, run_l[j], 8);
            run_l[j] += __shfl_xor_sync(0xffffffffu, run_l[j], 16);
          }
          float *ws = d.split_ws + (((size_t)a.b * d.n_q + (size_t)a.kvh * Gq) * d.n_splits + a.split) * (CH_HD + 2);
          const size_t ws_head = (size_t)d.n_splits * (CH_HD + 2);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int h = cq + j;
            if (h < Gq) {
              float *wh = ws + (size_t)h * ws_head;
#pragma unroll
              for (int mt = 0; mt < KK; ++mt) {
                wh[2 + mt * 16 + r0] = o[mt][j];
                wh[2 + mt * 16 + r0 + 8] = o[mt][2 + j];
              }
              if (r0 == 0) { wh[0] = run_m[j]; wh[1] = run_l[j]; }
            }
          }
        }
        pos += n_units;
        // partials of this CTA's items are stored: announce, then fold the key ranges of this CTA's (sequence, head) rows
        __threadfence();
        named_bar_sync(1, 128);
        if (et == 0) {
          ch_add_release(ctl.done + g, 1);
          ch_stamp(ctl, 8 + g * 8 + 3);
          ch_wait_count(ctl.done + g, G, "attention partials", g);
        }
        named_bar_sync(1, 128);
        {
          // decode_attn_combine_kernel: per (sequence, head) row M = max m_s, L = sum l_s e^(m_s - M), o = sum o_s e^(m_s - M) / L
          const int rows = ctl.B * d.n_q;
          for (int r = cta * 4 + quad; r < rows; r += 4 * G) {
            const int b = r / d.n_q, h = r - b * d.n_q;
            const int live = (d.ctx_len[b] + 1 + d.chunk - 1) / d.chunk;
            const float *ws = d.split_ws + ((size_t)b * d.n_q + h) * d.n_splits * (CH_HD + 2);
            // same arithmetic and order as decode_attn_combine_kernel; the loads do not wait for one another: lane s holds
            // (m_s, l_s) of key range s, the o_s rows are fetched eight ranges at a time
            float M = -INFINITY;
            for (int s0 = 0; s0 < live; s0 += 32) {
              const float mm = (s0 + lane < live) ? __ldcg(ws + (size_t)(s0 + lane) * (CH_HD + 2)) : -INFINITY;
              float mx = mm;
#pragma unroll
              for (int o_ = 16; o_ > 0; o_ >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o_));
              M = fmaxf(M, mx);
            }
            float L = 0.f, a4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s0 = 0; s0 < live; s0 += 32) {
              const bool ok = s0 + lane < live;
              const float mm = ok ? __ldcg(ws + (size_t)(s0 + lane) * (CH_HD + 2)) : -INFINITY;
              const float ll = ok ? __ldcg(ws + (size_t)(s0 + lane) * (CH_HD + 2) + 1) : 0.f;
              const float fl = (mm == -INFINITY) ? 0.f : __expf(mm - M);
              const int ns = min(32, live - s0);
              for (int sb = 0; sb < ns; sb += 8) {
                float2 v0[8], v1[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int s = s0 + sb + (sb + i < ns ? i : 0);
                  v0[i] = __ldcg(reinterpret_cast<const float2 *>(ws + (size_t)s * (CH_HD + 2) + 2 + lane * 4));
                  v1[i] = __ldcg(reinterpret_cast<const float2 *>(ws + (size_t)s * (CH_HD + 2) + 4 + lane * 4));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float f = __shfl_sync(0xffffffffu, fl, (sb + i) & 31);
                  const float lf = __shfl_sync(0xffffffffu, ll, (sb + i) & 31);
                  if (sb + i < ns && f != 0.f) {
                    L += lf * f;
                    a4[0] += v0[i].x * f; a4[1] += v0[i].y * f; a4[2] += v1[i].x * f; a4[3] += v1[i].y * f;
                  }
                }
              }
            }
            uint2 pk;
            pk.x = ch_pack_bf16(a4[0] / L, a4[1] / L);
            pk.y = ch_pack_bf16(a4[2] / L, a4[3] / L);
            *reinterpret_cast<uint2 *>(d.att + (size_t)b * d.ldo + (size_t)h * CH_HD + lane * 4) = pk;
          }
        }
        __threadfence();
        named_bar_sync(1, 128);
        if (et == 0) {
          ch_add_release(ctl.stage2 + g, 1);
          ch_stamp(ctl, 8 + g * 8 + 4);
        }
        continue;
      }
      // ═════════════ linear: stream-K segments of this CTA (skinny_gemm_kernel's epilogue) ═════════════
      const int KB = d.num_kb;
      SkSpan sp;
      sp.init(cta, G, d.num_tiles, KB);
      const int n_segs = sp.num_units() > 0 ? sp.num_segs() : 0;
      pos += sp.num_units();
      int *flags = ctl.flags + (size_t)g * SK_MAX_GRID;
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
            rr[i] = (c0 + i < ctl.B) ? (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(d.residual + (size_t)(c0 + i) * d.ldr + n)) : 0u;
        };
        if (use_res && d.res_early) load_res(0, rr_next);
        const float bv = (finishes && d.bias && n_ok) ? __bfloat162float(d.bias[n]) : 0.f;   // cold HBM read: before the wait
        mbar_wait(&tmem_full[acc], (segc >> 1) & 1);
        tcgen05_fence_after();
        if (use_res && !d.res_early) load_res(0, rr_next);
        if (et == 0 && seg == n_segs - 1) ch_stamp(ctl, 8 + g * 8 + 3);     // last segment of op g accumulated
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
          float *slot = ctl.partials + (size_t)cta * BC * 128;
#pragma unroll 1
          for (int c0 = 0; c0 < BC; c0 += CC) {
            float v[CC];
            load_chunk(c0, v);
#pragma unroll
            for (int i = 0; i < CC; ++i) slot[(c0 + i) * 128 + et] = v[i];
          }
          __threadfence();
          named_bar_sync(1, 128);
          if (et == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + cta), "r"(E1) : "memory");
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
                  if (f != E1 && clock64() - t0 > 4000000000LL) {
                    printf("ocrb chain: partial of CTA %d for op %d never arrived (CTA %d)\n", c, g, cta);
                    __trap();
                  }
                } while (f != E1);
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
                  const float *slot = ctl.partials + (size_t)(cb + (j < nc ? j : 0)) * BC * 128;
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
              // tile rows 0..63 = gate, 64..127 = up of output columns tile*64 + j.  Both halves publish their CC values
              // (column-major: conflict-free), then the gate threads finish the first half of the chunk's sequences and the up
              // threads the second half -- with the up threads only handing over, the 64 gate threads did all the SiLU / product
              // / store work of a tile (13 us per tile at B = 96)
#pragma unroll
              for (int i = 0; i < CC; ++i) s_up[i * 128 + et] = o[i];
              named_bar_sync(1, 128);
              {
                constexpr int HC = CC / 2;
                const int r = et & 63, cb = (et < 64) ? 0 : HC;
                if (tile * SK_BM + r < d.N) {
                  bf16 *dcol = d.D + (size_t)(c0 + cb) * d.ldd + (tile * 64 + r);
#pragma unroll
                  for (int i = 0; i < HC; ++i) {
                    const float val = sk_silu(s_up[(cb + i) * 128 + r]) * s_up[(cb + i) * 128 + 64 + r];
                    if (c0 + cb + i < ctl.B) dcol[(size_t)i * d.ldd] = __float2bfloat16_rn(val);
                  }
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
                  if (c0 + i < ctl.B) dcol[(size_t)i * d.ldd] = __float2bfloat16_rn(o[i]);
              }
            }
          }
        }
      }
      // this CTA's part of op g is stored (or published): announce it
      __threadfence();
      named_bar_sync(1, 128);
      if (et == 0) {
        ch_add_release(ctl.done + g, 1);
        ch_stamp(ctl, 8 + g * 8 + 4);
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  // The last CTA to leave returns the counters to zero and moves the flag epoch on (every other CTA has finished polling).
  if (threadIdx.x == 0) {
    __threadfence();
    *s_ticket = atomicAdd(ctl.exit_count, 1);
    ch_stamp(ctl, 7);
  }
  __syncthreads();
  if (*s_ticket == G - 1) {
    for (int i = threadIdx.x; i < n_desc; i += CH_THREADS) { ctl.done[i] = 0; ctl.stage2[i] = 0; }
    if (threadIdx.x == 0) {
      *ctl.epoch = *reinterpret_cast<volatile int *>(ctl.epoch) + 1;
      *ctl.exit_count = 0;
    }
  }
}

template <int BP, int BC>
__global__ void __launch_bounds__(CH_THREADS, 1)
skinny_chain_kernel(const __grid_constant__ ChainParams p) {
  chain_body<BP, BC>(p.d, p.ctl);
}

template <int BP, int BC>
__global__ void __launch_bounds__(CH_THREADS, 1)
decode_plan_kernel(const ChainDesc *__restrict__ descs, const __grid_constant__ ChainCtl ctl) {
  chain_body<BP, BC>(descs, ctl);
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

template <int BP, int BC>
static int launch_plan(const ChainDesc *descs, const ChainCtl &ctl, int grid, cudaStream_t st) {
  constexpr size_t smem = ChainCfg<BP>::SMEM;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(decode_plan_kernel<BP, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  OCRB_CUDA(launch_pdl(decode_plan_kernel<BP, BC>, dim3(grid), dim3(CH_THREADS), smem, st, descs, ctl));
  return check_launch("decode_plan_kernel");
}

// workspace layout (ocrb_chain_workspace_bytes): partials | flags | done | stage2 | exit_count, epoch | normalised rows
struct ChainWs {
  float *partials;
  int *flags, *done, *stage2, *exit_count, *epoch;
  bf16 *xn;
};
static size_t ch_ws_ints() { return (size_t)CH_MAX_OPS * SK_MAX_GRID + 2 * (size_t)CH_MAX_OPS + 64; }
static ChainWs ch_carve(void *workspace) {
  ChainWs w;
  char *ws = (char *)workspace;
  w.partials = (float *)ws;
  ws += (size_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float);
  w.flags = (int *)ws;
  w.done = w.flags + (size_t)CH_MAX_OPS * SK_MAX_GRID;
  w.stage2 = w.done + CH_MAX_OPS;
  w.exit_count = w.stage2 + CH_MAX_OPS;
  w.epoch = w.exit_count + 16;
  ws += ch_ws_ints() * sizeof(int);
  ws = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  w.xn = (bf16 *)ws;
  return w;
}

static int ch_pf_ahead() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("OCRB_CHAIN_PREFETCH");
    v = e ? atoi(e) : 0;       // measured: no gain, see profiles/r02_notes.md
    if (v < 0) v = 0;
  }
  return v;
}

static unsigned long long *g_ch_trace = nullptr;
static int g_ch_trace_slots = 0;

static void ch_fill_ctl(ChainCtl &c, const ChainWs &w, int n, int B) {
  c.n_desc = n;
  c.B = B;
  c.partials = w.partials;
  c.flags = w.flags;
  c.done = w.done;
  c.stage2 = w.stage2;
  c.exit_count = w.exit_count;
  c.epoch = w.epoch;
  c.pf_ahead = ch_pf_ahead();
  c.trace = g_ch_trace;
  c.trace_slots = g_ch_trace_slots;
}

static int ch_bp(int B) { return B <= 16 ? 16 : (B <= 32 ? 32 : (B <= 64 ? 64 : (B <= 96 ? 96 : 128))); }

// Fills one linear descriptor; `prev_D` = output of the op right before (nullptr: none).
static int ch_fill_linear(ChainDesc &d, const ocrb_chain_linear &l, int g, int B, const void *prev_D, bf16 *xn_slot) {
  OCRB_REQUIRE(l.X && l.W && l.D, "chain: null pointer in linear %d", g);
  OCRB_REQUIRE(l.N > 0 && l.K > 0 && l.K % 8 == 0 && l.ldx % 8 == 0 && l.ldw % 8 == 0,
               "chain: K and row strides must be multiples of 8 (linear %d)", g);
  OCRB_REQUIRE(((uintptr_t)l.X & 15) == 0 && ((uintptr_t)l.W & 15) == 0 && (!l.norm_w || ((uintptr_t)l.norm_w & 15) == 0),
               "chain: X, W, norm_w must be 16-byte aligned (linear %d)", g);
  OCRB_REQUIRE(l.epilogue >= 0 && l.epilogue <= 3, "chain: bad epilogue (linear %d)", g);
  OCRB_REQUIRE(l.epilogue != OCRB_EPI_RESIDUAL || l.residual, "chain: residual epilogue without residual (linear %d)", g);
  OCRB_REQUIRE(l.epilogue != OCRB_EPI_SWIGLU || l.N % 128 == 0, "chain: SwiGLU needs packed N %% 128 == 0 (linear %d)", g);
  OCRB_REQUIRE(!l.norm_w || l.K <= CH_XN_MAX_K, "chain: the RMSNorm prologue supports K <= 8192 (linear %d)", g);
  d.kind = CH_KIND_LINEAR;
  d.N = l.N;
  d.K = l.K;
  d.num_tiles = cdiv(l.N, SK_BM);
  d.num_kb = cdiv(l.K, SK_BK);
  OCRB_REQUIRE((long long)d.num_tiles * d.num_kb < (1ll << 30), "chain: problem too large (linear %d)", g);
  d.epilogue = l.epilogue;
  d.bias = (const bf16 *)l.bias;
  d.residual = (l.epilogue == OCRB_EPI_RESIDUAL) ? (const bf16 *)l.residual : nullptr;
  d.ldr = l.ldr;
  d.D = (bf16 *)l.D;
  d.ldd = l.ldd;
  d.eps = l.eps;
  d.res_early = (prev_D != l.residual) ? 1 : 0;
  d.W = (const bf16 *)l.W;
  d.ldw = l.ldw;      // the residual may be fetched ahead unless the op right before writes it
  int rc = make_tensor_map_bf16(&d.map_w, l.W, l.N, l.K, l.ldw, SK_BM);
  if (rc) return rc;
  const void *xsrc = l.X;
  long long ldx = l.ldx;
  if (l.norm_w) {
    d.norm_w = (const bf16 *)l.norm_w;
    d.Xraw = (const bf16 *)l.X;
    d.ldx = l.ldx;
    d.xn = xn_slot;
    xsrc = d.xn;
    ldx = l.K;
  }
  return make_tensor_map_bf16(&d.map_x, xsrc, B, l.K, ldx, ch_bp(B));
}

struct PlanHeader { uint32_t magic; int32_t n, B, has_attn; int32_t pad[12]; };     // 64 bytes in front of the descriptors
constexpr uint32_t CH_PLAN_MAGIC = 0x0c4a1b20u;

// plans built in this process: device pointer -> (ops, rows), so that a run with the wrong count or batch is refused
// instead of walking garbage descriptors
static std::mutex g_plan_mu;
static std::map<const void *, std::pair<int, int>> g_plans;

}  // namespace ocrb

using namespace ocrb;

/* debug hook (not in the public header): device buffer [grid][slots] of globaltimer stamps for the next launches */
extern "C" void ocrb_chain_set_trace(void *buf, int32_t slots) {
  g_ch_trace = (unsigned long long *)buf;
  g_ch_trace_slots = buf ? slots : 0;
}

extern "C" int64_t ocrb_chain_workspace_bytes(void) {
  return (int64_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float) + (int64_t)ch_ws_ints() * sizeof(int) + 512 +
         (int64_t)CH_XN_SLOTS * SK_MAXBP * CH_XN_MAX_K * sizeof(bf16);
}

#define CH_DISPATCH(FN, ...)                                   \
  do {                                                         \
    if (B <= 4) return FN<16, 4>(__VA_ARGS__);                 \
    if (B <= 8) return FN<16, 8>(__VA_ARGS__);                 \
    if (B <= 16) return FN<16, 16>(__VA_ARGS__);               \
    if (B <= 32) return FN<32, 32>(__VA_ARGS__);               \
    if (B <= 64) return FN<64, 64>(__VA_ARGS__);               \
    if (B <= 96) return FN<96, 96>(__VA_ARGS__);               \
    return FN<128, 128>(__VA_ARGS__);                          \
  } while (0)

extern "C" int ocrb_skinny_chain_bf16(const ocrb_chain_linear *lin, int32_t n, int32_t B, void *workspace, void *stream) {
  OCRB_REQUIRE(lin && workspace, "skinny_chain_bf16: null pointer");
  OCRB_REQUIRE(n >= 1 && n <= CH_MAXD, "skinny_chain_bf16: 1..%d linears per chain", CH_MAXD);
  OCRB_REQUIRE(B >= 1 && B <= SK_MAXBP, "skinny_chain_bf16: B must be in 1..128");
  const int grid = ch_sm_count();
  OCRB_REQUIRE(grid <= SK_MAX_GRID, "skinny_chain_bf16: more SMs than workspace slots");
  const ChainWs w = ch_carve(workspace);
  ChainParams p;
  memset(&p, 0, sizeof(p));
  ch_fill_ctl(p.ctl, w, n, B);
  for (int g = 0; g < n; ++g) {
    int rc = ch_fill_linear(p.d[g], lin[g], g, B, g ? lin[g - 1].D : nullptr, w.xn + (size_t)(g % CH_XN_SLOTS) * SK_MAXBP * CH_XN_MAX_K);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  CH_DISPATCH(launch_chain, p, grid, st);
}

extern "C" int64_t ocrb_chain_plan_bytes(int32_t n_ops) {
  return (int64_t)sizeof(PlanHeader) + (int64_t)(n_ops > 0 ? n_ops : 0) * sizeof(ChainDesc);
}

extern "C" int ocrb_chain_plan_build(const ocrb_chain_op *ops, int32_t n, int32_t B, void *workspace, void *plan, void *stream) {
  OCRB_REQUIRE(ops && workspace && plan, "chain_plan_build: null pointer");
  OCRB_REQUIRE(n >= 1 && n <= CH_MAX_OPS, "chain_plan_build: 1..%d ops per plan", CH_MAX_OPS);
  OCRB_REQUIRE(B >= 1 && B <= SK_MAXBP, "chain_plan_build: B must be in 1..128");
  OCRB_REQUIRE(((uintptr_t)plan & 63) == 0, "chain_plan_build: the plan buffer must be 64-byte aligned");
  const int grid = ch_sm_count();
  OCRB_REQUIRE(grid <= SK_MAX_GRID, "chain_plan_build: more SMs than workspace slots");
  const ChainWs w = ch_carve(workspace);
  std::vector<char> host(sizeof(PlanHeader) + (size_t)n * sizeof(ChainDesc), 0);
  PlanHeader *h = (PlanHeader *)host.data();
  ChainDesc *descs = (ChainDesc *)(host.data() + sizeof(PlanHeader));
  h->magic = CH_PLAN_MAGIC;
  h->n = n;
  h->B = B;
  int n_norm = 0;
  const ocrb_chain_attention *a0 = nullptr;
  for (int g = 0; g < n; ++g) {
    const ocrb_chain_op &op = ops[g];
    const void *prev_D = nullptr;
    if (g) prev_D = ops[g - 1].kind == CH_KIND_LINEAR ? ops[g - 1].lin.D : ops[g - 1].att.out;
    if (op.kind == CH_KIND_LINEAR) {
      bf16 *slot = w.xn + (size_t)(n_norm % CH_XN_SLOTS) * SK_MAXBP * CH_XN_MAX_K;
      if (op.lin.norm_w) ++n_norm;
      int rc = ch_fill_linear(descs[g], op.lin, g, B, prev_D, slot);
      if (rc) return rc;
    } else {
      OCRB_REQUIRE(op.kind == CH_KIND_ATTN, "chain_plan_build: bad kind %d (op %d)", op.kind, g);
      const ocrb_chain_attention &a = op.att;
      OCRB_REQUIRE(a.qkv && a.k_cache && a.v_cache && a.block_table && a.ctx_len && a.cosT && a.sinT && a.out && a.split_ws,
                   "chain_plan_build: null pointer in attention op %d", g);
      OCRB_REQUIRE(a.hd == CH_HD && a.n_kv > 0 && a.n_q % a.n_kv == 0 && a.n_q / a.n_kv <= 8,
                   "chain_plan_build: fused attention needs hd 128 and <= 8 query heads per KV head (op %d)", g);
      OCRB_REQUIRE(a.page_size > 0 && a.page_size % 16 == 0 && a.n_splits > 0 && a.n_cache_pages > 0,
                   "chain_plan_build: page_size must be a multiple of 16 (op %d)", g);
      OCRB_REQUIRE(a.ldqkv % 2 == 0 && a.ldo % 4 == 0 && ((uintptr_t)a.qkv & 3) == 0 && ((uintptr_t)a.out & 7) == 0 &&
                   ((uintptr_t)a.k_cache & 15) == 0 && ((uintptr_t)a.v_cache & 15) == 0,
                   "chain_plan_build: attention operand alignment (op %d)", g);
      ChainDesc &d = descs[g];
      d.kind = CH_KIND_ATTN;
      d.qkv = (const bf16 *)a.qkv;
      d.ldqkv = a.ldqkv;
      d.k_cache = (bf16 *)a.k_cache;
      d.v_cache = (bf16 *)a.v_cache;
      d.block_table = a.block_table;
      d.max_pages = a.max_pages;
      d.ctx_len = a.ctx_len;
      d.page_size = a.page_size;
      d.n_q = a.n_q;
      d.n_kv = a.n_kv;
      d.n_splits = a.n_splits;
      const int max_ctx = a.max_pages * a.page_size;
      d.chunk = cdiv(cdiv(max_ctx, a.n_splits), 16) * 16;        // keys per work item, whole tiles (decode.cu)
      d.cosT = (const bf16 *)a.cosT;
      d.sinT = (const bf16 *)a.sinT;
      d.scale = a.scale;
      d.split_ws = a.split_ws;
      d.att = (bf16 *)a.out;
      d.ldo = a.ldo;
      const long long rows = (long long)a.n_cache_pages * a.n_kv * a.page_size;
      int rc = make_tensor_map_bf16(&d.map_w, a.k_cache, rows, CH_HD, CH_HD, 16);
      if (rc) return rc;
      rc = make_tensor_map_bf16(&d.map_x, a.v_cache, rows, CH_HD, CH_HD, 16);
      if (rc) return rc;
      if (!a0) {
        a0 = &a;
        const long long items = (long long)B * a.n_kv * a.n_splits;
        OCRB_REQUIRE(cdiv(items, grid) <= CH_MAX_ITEMS && cdiv(items, grid) * cdiv(d.chunk, 32) <= CH_MAX_UNITS,
                     "chain_plan_build: %lld attention items exceed the per-CTA schedule (%d items / %d units per CTA)", items,
                     CH_MAX_ITEMS, CH_MAX_UNITS);
        h->has_attn = 1;
      } else {
        // one schedule serves every attention op of the plan: same geometry, only the caches differ
        OCRB_REQUIRE(a.block_table == a0->block_table && a.ctx_len == a0->ctx_len && a.max_pages == a0->max_pages &&
                     a.page_size == a0->page_size && a.n_kv == a0->n_kv && a.n_q == a0->n_q && a.n_splits == a0->n_splits,
                     "chain_plan_build: the attention ops of one plan must share block table, context lengths and geometry (op %d)", g);
      }
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  OCRB_CUDA(cudaMemcpyAsync(plan, host.data(), host.size(), cudaMemcpyHostToDevice, st));
  OCRB_CUDA(cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    g_plans[plan] = std::make_pair((int)n, (int)B);
  }
  return OCRB_OK;
}

extern "C" int ocrb_chain_plan_run(const void *plan, int32_t n, int32_t B, void *workspace, void *stream) {
  OCRB_REQUIRE(plan && workspace, "chain_plan_run: null pointer");
  OCRB_REQUIRE(n >= 1 && n <= CH_MAX_OPS && B >= 1 && B <= SK_MAXBP, "chain_plan_run: bad n / B");
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(plan);
    OCRB_REQUIRE(it != g_plans.end() && it->second.first == n && it->second.second == B,
                 "chain_plan_run: not a plan built by ocrb_chain_plan_build for %d ops and %d rows", n, B);
  }
  const int grid = ch_sm_count();
  const ChainWs w = ch_carve(workspace);
  ChainCtl ctl;
  memset(&ctl, 0, sizeof(ctl));
  ch_fill_ctl(ctl, w, n, B);
  const ChainDesc *descs = (const ChainDesc *)((const char *)plan + sizeof(PlanHeader));
  cudaStream_t st = (cudaStream_t)stream;
  CH_DISPATCH(launch_plan, descs, ctl, grid, st);
}
