// Weight-streaming "skinny" GEMM for batched greedy decode (HF generation loop utils.py:2743-2806: one new
// token per sequence per step, so every nn.Linear of the decoder sees only B <= 64 activation rows):
//
//     D[b, n] = epilogue( sum_k X[b, k] * W[n, k] )          X: [B, K] bf16 (B <= 128),  W: [N, K] bf16
//
// HBM-bound: each weight byte must cross HBM exactly once per step whatever B is.  Design (sm_100a):
//   * swap-AB on the 5th-gen tensor cores: the 128 x 64 WEIGHT tile is the UMMA "A" operand (M = 128 weight
//     rows), the activations are the "B" operand (N = BP = 16/32/64/96/128 batch columns), fp32 accumulators
//     [128 lanes x BP columns] in TMEM, double buffered so the epilogue of one tile overlaps the MMAs of the next;
//   * weights arrive through a TMA ring (cp.async.bulk.tensor, 128B swizzle, L2 evict-first) that never
//     depends on the activations, so it starts before the previous kernel's results are needed;
//   * stream-K: the (tile, k-block) iteration space is cut into gridDim.x equal contiguous spans, one per SM,
//     so a 3584x3584 o_proj (28 tiles) fills the machine as evenly as the 152064-row lm_head.  A span that
//     does not finish its tile publishes an fp32 partial; the CTA that finishes the tile adds the partials in
//     k order (deterministic, independent of B => batch-invariant results);
//   * activations: TMA boxes of X itself ([BP rows x 64], rows >= B zero-filled by the hardware), issued after
//     griddepcontrol.wait.  With an RMSNorm in front (qkv, gate/up, lm_head) the B rows are normalised ONCE by
//     skinny_norm_rows_kernel (HF rounding, modeling_qwen2_5_vl.py:66-71) into a scratch buffer and read from there:
//     fusing the norm into the GEMM (every CTA re-normalising the k-slices of every tile it streams, through producer
//     warps writing the UMMA swizzled layout) was built and measured slower -- 3.58 vs 3.41 ms per decode step at B = 3,
//     and 2.5x slower at B = 64 (profiles/r01_notes.md);
//   * epilogue warps: tcgen05.ld (only the live accumulator columns, 16 at a time) -> bias -> bf16 round -> residual /
//     SwiGLU / GELU with HF's rounding points.
//
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue (TMEM lane quadrant =
// warp % 4).
#include "skinny_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace ocrb {

// 65..128 activation rows: ring depth and CTAs per SM (two CTAs of consecutive launches on one SM = PDL overlap)
#ifndef SK_ST_BIG
#define SK_ST_BIG 6
#endif
#ifndef SK_OCC_BIG
#define SK_OCC_BIG 1
#endif


// Tensor-parallel row-parallel linear (o_proj / down_proj of the decode step, HF base_model_tp_plan "rowwise"): the
// all-reduce of the bf16 partial products and the residual add run INSIDE the cluster kernel's epilogue, over NVLink peer
// memory (see the TP section of skinny_cluster_kernel).  world == 0: not a tensor-parallel call.
constexpr int SK_TP_UNITS = 512;   // (tile, cluster rank) exchange units: 128 tiles x cluster of 4
constexpr int SK_EPI_TP = 4;       // internal epilogue code (not part of the public OCRB_EPI_* set)
struct SkinnyTp {
  const bf16 *data[8];             // every rank's partial slot of this call (device pointers valid on this GPU; own at [rank])
  int *flags[8];                   // every rank's flag array [SK_TP_UNITS][8]
  int *seq;                        // this rank's call counters [SK_TP_UNITS]
  bf16 *x; long long ldx;          // residual stream [B, N], updated in place
  long long ldp;                   // row stride of the partial slots
  int world, rank;
  int mode;                        // bit 0 (experiment): every thread fences before the announcement; bit 1: LL protocol
  unsigned long long *ll[8];       // LL protocol: every rank's receive buffer, cells [2 slots][8 sources][64 rows][ldp / 2]
  int *ll_seq;                     // LL protocol: [0] global call index, [1] ticket of the call in flight
};
constexpr int SK_LL_ROWS = 64;

struct SkinnyParams {
  SkinnyTp tp;
  const bf16 *X; long long ldx;
  bf16 *D; long long ldd;
  const bf16 *bias;
  const bf16 *residual; long long ldr;
  const bf16 *norm_w; float eps;
  int B, N, K;
  int epilogue;
  int num_tiles, num_kb;
  unsigned long long *trace;   // optional [grid][64] globaltimer stamps (debug / profiling), or nullptr
  float *partials;       // [SK_MAX_GRID][BP][128] fp32
  int *flags;            // [SK_MAX_GRID]
};


__device__ __forceinline__ void sk_stamp(const SkinnyParams &p, int slot) {
  if (p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[(size_t)blockIdx.x * 64 + slot] = t;
  }
}



// RMSNorm of the B activation rows into a scratch buffer, for B > 16: above that the fused producers (every CTA
// re-normalising the k-slices of every tile it streams) cost more than the weights' HBM time, so the rows are
// normalised once and the GEMM reads them through TMA like any un-normed input.  Same statistics routine as the fused
// path (one warp per row, fixed summation order), so a row gets the same bits whatever B is.
__global__ void __launch_bounds__(256)
skinny_norm_rows_kernel(SkinnyParams p, bf16 *__restrict__ xn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int b = blockIdx.x * 8 + warp;             // one warp per row, no block-level synchronisation
  const int kvec = p.K >> 3;
  // the norm weights do not depend on the predecessor: fetch the first batch before the wait
  uint4 wraw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int v = lane + i * 32;
    wraw[i] = make_uint4(0, 0, 0, 0);
    if (v < kvec) wraw[i] = *reinterpret_cast<const uint4 *>(p.norm_w + v * 8);
  }
  pdl_wait();
  if (b >= p.B) return;
  const bf16 *xr = p.X + (size_t)b * p.ldx;
  bf16 *yr = xn + (size_t)b * p.K;
  if (kvec <= 512) {
    // K <= 4096: the row stays in registers between the statistics and the scaling (one pass over memory).
    // Same summation order as sk_one_row_rstd (lane-strided, 16 loads per batch, xor-shuffle tree).
    uint4 raw[16];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int v = lane + i * 32;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (v < kvec) raw[i] = __ldcg(reinterpret_cast<const uint4 *>(xr + v * 8));
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
    const float rs = rsqrtf(ss / (float)p.K + p.eps);
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
    const float rs = sk_one_row_rstd(xr, p.K, p.eps, lane);
    for (int v = lane; v < kvec; v += 32) {
      float f[8], wf[8];
      unpack8f(__ldcg(reinterpret_cast<const uint4 *>(xr + v * 8)), f);
      unpack8f(*reinterpret_cast<const uint4 *>(p.norm_w + v * 8), wf);
      uint4 o;
      bf16 *oe = reinterpret_cast<bf16 *>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) oe[e] = __float2bfloat16_rn(wf[e] * bf16_round(f[e] * rs));
      *reinterpret_cast<uint4 *>(yr + v * 8) = o;
    }
  }
}

// BP = MMA N (activation rows staged per k-block: 16/32/64); BC = accumulator columns the epilogue actually reads,
// publishes and stores (4/8/16/32/64 >= B): with B = 3 sequences the fix-up moves 4 columns, not 16.
template <int BP, int BC>
__global__ void __launch_bounds__(SK_THREADS, (BP > 64) ? SK_OCC_BIG : 2)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, SkinnyParams p) {
  // ring depth: 4 x 24 KiB at BP = 64 keeps two CTAs per SM; above that one CTA per SM with 6 stages (96 KiB of weights in flight)
  constexpr int ST = (BP > 64) ? SK_ST_BIG : ((BP >= 64) ? 4 : SK_STAGES);
  constexpr uint32_t X_BYTES = BP * SK_BK * 2;
  constexpr uint32_t STAGE_BYTES = SK_W_BYTES + X_BYTES;
  constexpr int ACC_STRIDE = (BP <= 16) ? 16 : (BP <= 32 ? 32 : (BP <= 64 ? 64 : 128));   // TMEM columns between the two accumulators
  constexpr int TMEM_COLS = (2 * ACC_STRIDE < 32) ? 32 : 2 * ACC_STRIDE;              // power of two >= 32
  extern __shared__ uint8_t sk_smem_raw[];
  // 1024-byte alignment for the swizzle atoms, computed as an OFFSET into the __shared__ array so that the compiler keeps
  // the shared address space (LDS / STS instead of generic LD / ST, and no false aliasing with the global stores)
  uint8_t *smem = sk_smem_raw + ((1024u - (smem_u32(sk_smem_raw) & 1023u)) & 1023u);
  uint64_t *full_w = reinterpret_cast<uint64_t *>(smem + ST * STAGE_BYTES);
  uint64_t *full_x = full_w + ST;
  uint64_t *empty = full_x + ST;
  uint64_t *tmem_full = empty + ST;     // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint64_t *fix_bar = tmem_empty + 2;          // partials of the other contributors landed in the (idle) ring
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fix_bar + 1);
  float *s_up = reinterpret_cast<float *>(tmem_slot + 2);          // [CC][128] SwiGLU exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sk_stamp(p, 0);            // CTA start
  const int KB = p.num_kb;
  SkSpan sp;
  sp.init(blockIdx.x, gridDim.x, p.num_tiles, KB);
  const int n_units = sp.num_units();
  const int n_segs = sp.num_segs();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 1);                  // activation k-slice: one TMA transaction
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    mbar_init(fix_bar, 1);
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
  if (threadIdx.x == 0) { sk_stamp(p, 1); pdl_launch_dependents(); }   // the next kernel may start its own weight prefetch now
  if (warp >= 2) pdl_wait();                        // activations / residual / workspace belong to predecessors
  if (threadIdx.x == 64) sk_stamp(p, 2);            // wait returned

  if (warp == 0) {
    // ───────────── TMA producer ─────────────
    // Weights never depend on the previous kernel: their loads start at once (under PDL: while the predecessor is still
    // running).  Without a fused RMSNorm the activation k-slices are TMA boxes of X itself ([BP rows x 64], rows >= B
    // zero-filled by the hardware); they do depend on the predecessor, so they are issued after griddepcontrol.wait.
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      SkCursor cur, xc;
      cur.init(sp);
      xc = cur;
      const int head = n_units < ST ? n_units : ST;
      for (int it = 0; it < head; ++it) {           // first pass over the ring: slots are free, weights only
        mbar_expect_tx(&full_w[it], SK_W_BYTES);
        tma_load_2d_hint(smem + it * STAGE_BYTES, &map_w, &full_w[it], cur.kb * SK_BK, cur.tile * SK_BM, policy);
        cur.advance(sp);
      }
      {
        pdl_wait();
        for (int it = 0; it < head; ++it) {
          mbar_expect_tx(&full_x[it], X_BYTES);
          tma_load_2d(smem + it * STAGE_BYTES + SK_W_BYTES, &map_x, &full_x[it], xc.kb * SK_BK, 0);
          xc.advance(sp);
        }
      }
      int s = head == ST ? 0 : head;
      uint32_t ph = head == ST ? 1 : 0;
      for (int it = head; it < n_units; ++it) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full_w[s], SK_W_BYTES);
        tma_load_2d_hint(smem + s * STAGE_BYTES, &map_w, &full_w[s], cur.kb * SK_BK, cur.tile * SK_BM, policy);
        {
          mbar_expect_tx(&full_x[s], X_BYTES);
          tma_load_2d(smem + s * STAGE_BYTES + SK_W_BYTES, &map_x, &full_x[s], cur.kb * SK_BK, 0);
        }
        cur.advance(sp);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ───────────── MMA issuer ─────────────
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BP >> 3) << 17) | ((uint32_t)(SK_BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      for (int seg = 0; seg < n_segs; ++seg) {
        int tile, kb0, nkb;
        sp.seg(seg, tile, kb0, nkb);
        const int acc = seg & 1;
        mbar_wait(&tmem_empty[acc], ((seg >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + acc * ACC_STRIDE;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_w[s], ph);
          if (seg == 0 && i == 0) sk_stamp(p, 3);   // first weight tile landed
          if (p.trace && seg == 0 && i < 24) sk_stamp(p, 16 + 2 * i);
          mbar_wait(&full_x[s], ph);
          if (seg == 0 && i == 0) sk_stamp(p, 4);   // first activation slice ready
          if (p.trace && seg == 0 && i < 24) sk_stamp(p, 17 + 2 * i);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + SK_W_BYTES);
#pragma unroll
          for (int k = 0; k < SK_BK / UMMA_K; ++k)
            umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == ST) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ───────────── epilogue warps 2..5 ─────────────
    const int quad = warp & 3;
    const int et = quad * 32 + lane;             // TMEM lane = weight row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int seg = 0; seg < n_segs; ++seg) {
      int tile, kb0, nkb;
      sp.seg(seg, tile, kb0, nkb);
      const int acc = seg & 1;
      // The BC live accumulator columns are processed in chunks of CC <= 16, so the register footprint (and with it
      // the two-CTAs-per-SM co-residency PDL needs) is the same for 4 and for 64 sequences.
      constexpr int CC = (BC < 16) ? BC : 16;
      const bool finishes = (kb0 + nkb == KB);
      const int n = tile * SK_BM + et;
      const bool n_ok = n < p.N;
      // Residual values are fetched one chunk AHEAD of their use and the first chunk before the accumulator is even
      // complete: the residual is usually updated in place (D == residual), so the compiler must keep every load behind
      // the stores that precede it in program order -- loading inside the store loop serialised one L2 round trip per
      // sequence (60 us of a 64 us o_proj at B = 96).
      const bool use_res = finishes && p.epilogue == OCRB_EPI_RESIDUAL && n_ok;
      uint32_t rr_next[CC];
      auto load_res = [&](int c0, uint32_t (&rr)[CC]) {
#pragma unroll
        for (int i = 0; i < CC; ++i)
          rr[i] = (c0 + i < p.B) ? (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(p.residual + (size_t)(c0 + i) * p.ldr + n)) : 0u;
      };
      if (use_res) load_res(0, rr_next);
      mbar_wait(&tmem_full[acc], (seg >> 1) & 1);
      if (et == 0 && seg == 0) sk_stamp(p, 5);      // first segment accumulated
      if (et == 0 && seg == n_segs - 1) sk_stamp(p, 6);   // last segment accumulated
      tcgen05_fence_after();
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
        float *slot = p.partials + (size_t)blockIdx.x * BC * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < BC; c0 += CC) {
          float v[CC];
          load_chunk(c0, v);
#pragma unroll
          for (int i = 0; i < CC; ++i) slot[(c0 + i) * 128 + et] = v[i];   // plain stores: __threadfence + release flag publish them
        }
        __threadfence();
        named_bar_sync(1, 128);
        if (et == 0) {
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags + blockIdx.x), "r"(1) : "memory");
          sk_stamp(p, 7);                           // partial published
        }
      } else {
        int c_first = (int)blockIdx.x;               // first contributing CTA (== blockIdx.x: none)
        constexpr uint32_t PART_BYTES = BC * 128 * sizeof(float);
        constexpr int FIX_SLOTS = (ST * STAGE_BYTES / PART_BYTES) < 8 ? (int)(ST * STAGE_BYTES / PART_BYTES) : 8;
        bool in_smem = false;
        if (kb0 > 0) {
          // this CTA finishes a tile that earlier CTAs started: their partials are added in k order, then ours.
          // First contributing CTA = the one whose span contains the tile's first unit.
          const int tile_first = tile * KB;
          const int total = p.num_tiles * KB;
          const int per = total / (int)gridDim.x, rem = total % (int)gridDim.x;
          const int big = rem * (per + 1);
          c_first = (tile_first < big) ? tile_first / (per + 1) : rem + (tile_first - big) / per;
          // all contributors' flags first (one polling thread per contributor), then stream their partials
          for (int cb = c_first; cb < (int)blockIdx.x; cb += 128) {
            if (cb + et < (int)blockIdx.x) {
              const int c = cb + et;
              int f;
              const long long t0 = clock64();
              do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.flags + c) : "memory");
                if (!f && clock64() - t0 > 4000000000LL) {
                  printf("ocrb skinny gemm: partial of CTA %d never arrived (CTA %d)\n", c, blockIdx.x);
                  __trap();
                }
              } while (!f);
            }
          }
          named_bar_sync(1, 128);
          // This is the CTA's LAST segment (only the head tile of a span can have been started by others), all its
          // MMAs have completed and the producer has nothing left to load: the weight ring is idle.  Fetch every
          // contributor's [BC][128] fp32 partial into it with ONE batch of bulk copies -- a single L2 round trip
          // instead of one per 16-column chunk per pair of contributors (at B = 96 the chunked loads cost ~30 us).
          if ((int)blockIdx.x - c_first <= FIX_SLOTS) {
            in_smem = true;
            if (et == 0) {
              asm volatile("fence.proxy.async;" ::: "memory");     // partials were written through the generic proxy
              const int nc = (int)blockIdx.x - c_first;
              mbar_expect_tx(fix_bar, (uint32_t)nc * PART_BYTES);
              for (int j = 0; j < nc; ++j)
                bulk_load(smem + (size_t)j * PART_BYTES, p.partials + (size_t)(c_first + j) * BC * 128, PART_BYTES, fix_bar);
            }
            mbar_wait(fix_bar, 0);
            if (et == 0) sk_stamp(p, 10);              // contributors' partials are in shared memory
          }
        }
        const float bv = (p.bias && n_ok) ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BC; c0 += CC) {
          float v[CC];
          uint32_t rr[CC];
#pragma unroll
          for (int i = 0; i < CC; ++i) rr[i] = rr_next[i];
          if (use_res && c0 + CC < BC) load_res(c0 + CC, rr_next);
          load_chunk(c0, v);
          if (p.trace && et == 0 && seg == n_segs - 1) sk_stamp(p, 40 + c0 / CC);
          if (in_smem) {
            // same summation order as the register path below: contributors in k order, own accumulator last
            float sum[CC];
#pragma unroll
            for (int i = 0; i < CC; ++i) sum[i] = 0.f;
            const int nc = (int)blockIdx.x - c_first;
            for (int j = 0; j < nc; ++j) {
              const float *slot = reinterpret_cast<const float *>(smem + (size_t)j * PART_BYTES);
#pragma unroll
              for (int i = 0; i < CC; ++i) sum[i] += slot[(c0 + i) * 128 + et];
            }
#pragma unroll
            for (int i = 0; i < CC; ++i) v[i] = sum[i] + v[i];
          } else if (c_first < (int)blockIdx.x) {
            float sum[CC];
#pragma unroll
            for (int i = 0; i < CC; ++i) sum[i] = 0.f;
            constexpr int FX = (CC <= 4) ? 8 : ((CC <= 8) ? 4 : 2);   // contributors fetched together
            for (int cb = c_first; cb < (int)blockIdx.x; cb += FX) {
              const int nc = min(FX, (int)blockIdx.x - cb);
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
          // ───── epilogue math on the complete accumulator (HF rounding points) ─────
          // All CC columns are computed unconditionally (columns >= B hold zeros) and only the stores are predicated:
          // with the arithmetic inside `if (column < B)` every column became its own basic block and the 16 columns of
          // a chunk ran as 16 serial latency chains (1.6 us per chunk at B = 96).
          float o[CC];
#pragma unroll
          for (int i = 0; i < CC; ++i) o[i] = bf16_round(v[i] + bv);
          if (p.epilogue == OCRB_EPI_SWIGLU) {
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
              if (tile * SK_BM + r < p.N) {
                bf16 *dcol = p.D + (size_t)(c0 + cb) * p.ldd + (tile * 64 + r);
#pragma unroll
                for (int i = 0; i < HC; ++i) {
                  const float val = sk_silu(s_up[(cb + i) * 128 + r]) * s_up[(cb + i) * 128 + 64 + r];
                  if (c0 + cb + i < p.B) dcol[(size_t)i * p.ldd] = __float2bfloat16_rn(val);
                }
              }
            }
            named_bar_sync(1, 128);
          } else {
            if (p.epilogue == OCRB_EPI_RESIDUAL) {
#pragma unroll
              for (int i = 0; i < CC; ++i) o[i] += __uint_as_float(rr[i] << 16);
            } else if (p.epilogue == OCRB_EPI_GELU) {
#pragma unroll
              for (int i = 0; i < CC; ++i) o[i] = sk_gelu(o[i]);
            }
            if (n_ok) {
              bf16 *dcol = p.D + (size_t)c0 * p.ldd + n;
#pragma unroll
              for (int i = 0; i < CC; ++i)
                if (c0 + i < p.B) dcol[(size_t)i * p.ldd] = __float2bfloat16_rn(o[i]);
            }
          }
        }
        if (et == 0 && seg == n_segs - 1) sk_stamp(p, 11);   // epilogue of the last segment stored
        if (c_first < (int)blockIdx.x) {
          named_bar_sync(1, 128);                    // every thread is done reading the partials
          for (int c = c_first + et; c < (int)blockIdx.x; c += 128) p.flags[c] = 0;   // consumed: ready for the next launch
          if (et == 0) sk_stamp(p, 8);               // fix-up done
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) sk_stamp(p, 9);             // CTA end
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ───────────── cluster split-K variant for the narrow linears (qkv, o_proj, down_proj: fewer than 148 / 4 tiles) ─────────────
// A thread-block cluster of SKC_CS (4, 3 or 2) CTAs owns ONE 128-row weight tile; CTA r streams the r-th part of K through the
// same TMA ring / tcgen05 pipeline as above and ends with a [128 x BP] fp32 partial in TMEM.  The partials never go
// through global memory: the accumulator columns are dealt out to the CTAs of the cluster in groups of 8
// (group g belongs to CTA g mod SKC_CS), every CTA writes the groups it does not own straight into the owner's shared
// memory (st.shared::cluster, the weight ring is idle by then), and after one cluster barrier each CTA adds the SKC_CS
// partials of its own columns in k order (deterministic; the split depends only on K, never on B) and runs the
// epilogue for them.  Compared with the stream-K fix-up through L2 this removes the publish / poll / fetch round trips
// (~2.5 us at B = 3) and divides the epilogue of a tile by the cluster size (at B = 96: 24 columns per CTA, not 96).
static int sm_count();

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t local_saddr, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// TP: the tensor-parallel instantiation (exchange in the epilogue, below); the single-GPU instantiation carries none of it.
template <int BP, int BC, int SKC_CS, bool TP = false>
__global__ void __launch_bounds__(SK_THREADS, (BP > 64) ? SK_OCC_BIG : 2)
skinny_cluster_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, SkinnyParams p) {
  constexpr int ST = (BP > 64) ? SK_ST_BIG : ((BP >= 64) ? 4 : SK_STAGES);
  constexpr uint32_t X_BYTES = BP * SK_BK * 2;
  constexpr uint32_t STAGE_BYTES = SK_W_BYTES + X_BYTES;
  constexpr int TMEM_COLS = (BP < 32) ? 32 : ((BP <= 32) ? 32 : (BP <= 64 ? 64 : 128));
  constexpr int NG = (BC + 7) / 8;                      // 8-column groups of the accumulator
  constexpr int LG = (NG + SKC_CS - 1) / SKC_CS;        // groups a CTA owns at most
  constexpr uint32_t SLOT_BYTES = LG * 8 * 128 * sizeof(float);      // one source CTA's contribution to my groups
  static_assert(SKC_CS * SLOT_BYTES <= ST * STAGE_BYTES, "receive buffer must fit the idle weight ring");
  extern __shared__ uint8_t sk_smem_raw[];
  uint8_t *smem = sk_smem_raw + ((1024u - (smem_u32(sk_smem_raw) & 1023u)) & 1023u);
  uint64_t *full_w = reinterpret_cast<uint64_t *>(smem + ST * STAGE_BYTES);
  uint64_t *full_x = full_w + ST;
  uint64_t *empty = full_x + ST;
  uint64_t *tmem_full = empty + ST;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = (int)blockIdx.x / SKC_CS;
  const int KB = p.num_kb;
  const int per = KB / SKC_CS, rem = KB % SKC_CS;
  const int kb_begin = rank * per + (rank < rem ? rank : rem);
  const int nkb = per + (rank < rem ? 1 : 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
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
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (warp >= 2) pdl_wait();

  if (warp == 0) {
    // ───────────── TMA producer: weights at once, activation slices after griddepcontrol.wait ─────────────
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      const int head = nkb < ST ? nkb : ST;
      for (int it = 0; it < head; ++it) {
        mbar_expect_tx(&full_w[it], SK_W_BYTES);
        tma_load_2d_hint(smem + it * STAGE_BYTES, &map_w, &full_w[it], (kb_begin + it) * SK_BK, tile * SK_BM, policy);
      }
      pdl_wait();
      for (int it = 0; it < head; ++it) {
        mbar_expect_tx(&full_x[it], X_BYTES);
        tma_load_2d(smem + it * STAGE_BYTES + SK_W_BYTES, &map_x, &full_x[it], (kb_begin + it) * SK_BK, 0);
      }
      int s = head == ST ? 0 : head;
      uint32_t ph = head == ST ? 1 : 0;
      for (int it = head; it < nkb; ++it) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full_w[s], SK_W_BYTES);
        tma_load_2d_hint(smem + s * STAGE_BYTES, &map_w, &full_w[s], (kb_begin + it) * SK_BK, tile * SK_BM, policy);
        mbar_expect_tx(&full_x[s], X_BYTES);
        tma_load_2d(smem + s * STAGE_BYTES + SK_W_BYTES, &map_x, &full_x[s], (kb_begin + it) * SK_BK, 0);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ───────────── MMA issuer ─────────────
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BP >> 3) << 17) | ((uint32_t)(SK_BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_w[s], ph);
        mbar_wait(&full_x[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bdesc = make_smem_desc(sa + SK_W_BYTES);
#pragma unroll
        for (int k = 0; k < SK_BK / UMMA_K; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  }

  // ───────────── exchange + epilogue ─────────────
  const int quad = warp & 3;
  const int et = quad * 32 + lane;                       // TMEM lane = weight row inside the tile (epilogue warps only)
  const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
  float *recv = reinterpret_cast<float *>(smem);         // [source CTA][local group][8 columns][128 rows], in the idle ring
  if (warp >= 2) {
    mbar_wait(tmem_full, 0);                             // this CTA's MMAs are complete: its ring is idle
    tcgen05_fence_after();
  }
  cluster_arrive();                                      // barrier A: every CTA of the cluster may now be written to
  cluster_wait();
  if (warp >= 2) {
    const uint32_t recv_s = smem_u32(recv);
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      uint32_t r[8];
      tmem_ld_cols<8>(lane_addr + g * 8, r);
      tmem_ld_wait();
      const uint32_t owner = (uint32_t)(g % SKC_CS);
      const uint32_t off = ((uint32_t)rank * LG + (uint32_t)(g / SKC_CS)) * 8 * 128 * sizeof(float) + (uint32_t)et * sizeof(float);
      if (owner == (uint32_t)rank) {
#pragma unroll
        for (int i = 0; i < 8; ++i) recv[(off >> 2) + i * 128] = (nkb > 0) ? __uint_as_float(r[i]) : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) st_cluster_f32(recv_s + off + i * 128 * sizeof(float), owner, (nkb > 0) ? __uint_as_float(r[i]) : 0.f);
      }
    }
    tcgen05_fence_before();
  }
  cluster_arrive();                                      // barrier B: all partials delivered (release / acquire)
  cluster_wait();
  if (warp >= 2) {
    const int n = tile * SK_BM + et;
    const bool n_ok = n < p.N;
    const float bv = (p.bias && n_ok) ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll 1
    for (int lg = 0; lg < LG; ++lg) {
      const int g = lg * SKC_CS + rank;                  // global 8-column group this CTA owns
      if (g >= NG) break;
      const int c0 = g * 8;
      uint32_t rr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rr[i] = 0u;
      if (p.epilogue == OCRB_EPI_RESIDUAL && n_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < p.B) rr[i] = (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(p.residual + (size_t)(c0 + i) * p.ldr + n));
      }
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float sum = 0.f;
#pragma unroll
        for (int src = 0; src < SKC_CS; ++src) sum += recv[((src * LG + lg) * 8 + i) * 128 + et];     // k order
        o[i] = bf16_round(sum + bv);
      }
      if (p.epilogue == OCRB_EPI_RESIDUAL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += __uint_as_float(rr[i] << 16);
      } else if (p.epilogue == OCRB_EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sk_gelu(o[i]);
      }
      if (TP && (p.tp.mode & 2)) {
        // LL protocol (as NCCL's low-latency one): the partial travels WITH its flag -- two bf16 values of adjacent rows
        // and the call index in one 8-byte cell, pushed by one 8-byte store (atomic over NVLink) into every peer's
        // receive buffer.  No fence, no flag round trip: the receiver polls its own memory for cells of this call.
        const uint32_t kk = (uint32_t)(p.tp.ll_seq[0] + 1);
        const size_t half = (size_t)(p.tp.ldp >> 1);
        const size_t base = ((size_t)(kk & 1u) * 8 + p.tp.rank) * SK_LL_ROWS;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t mine = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(o[i]));
          const uint32_t up = __shfl_down_sync(0xffffffffu, mine, 1);
          if ((lane & 1) == 0 && n_ok && c0 + i < p.B) {
            const unsigned long long cell = ((unsigned long long)kk << 32) | (up << 16) | mine;
            const size_t idx = (base + (size_t)(c0 + i)) * half + (size_t)(n >> 1);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < p.tp.world && q != p.tp.rank)
                asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p.tp.ll[q] + idx), "l"(cell) : "memory");
          }
        }
      } else if (n_ok) {
        bf16 *dcol = p.D + (size_t)c0 * p.ldd + n;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < p.B) dcol[(size_t)i * p.ldd] = __float2bfloat16_rn(o[i]);
      }
    }
    if (TP && (p.tp.mode & 2) && rank * 8 < p.B) {
      // LL receive: the peers' cells of my rows arrive in MY memory; sum in rank order with my own partial in its place
      const uint32_t kk = (uint32_t)(p.tp.ll_seq[0] + 1);
      const size_t half = (size_t)(p.tp.ldp >> 1);
      const unsigned long long *mybuf = p.tp.ll[p.tp.rank];
#pragma unroll 1
      for (int lg = 0; lg < LG; ++lg) {
        const int c0 = (lg * SKC_CS + rank) * 8;
        if (c0 >= p.B) break;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          if (c0 + i >= p.B) break;
          float sum = 0.f;
#pragma unroll
          for (int src = 0; src < SKC_CS; ++src) sum += recv[((src * LG + lg) * 8 + i) * 128 + et];     // k order, as above
          const float own = bf16_round(sum + bv);
          // the cells of ALL sources are requested together and re-read until each carries this call's index: one L2
          // round trip when the peers were not later than this rank, whatever the world size
          uint32_t pb[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) pb[q] = 0u;
          if ((lane & 1) == 0 && n_ok) {
            const unsigned long long *cp = mybuf + (((size_t)(kk & 1u) * 8) * SK_LL_ROWS + (size_t)(c0 + i)) * half + (size_t)(n >> 1);
            const size_t qstride = (size_t)SK_LL_ROWS * half;
            unsigned long long cell[8];
            bool all;
            const long long t0 = clock64();
            do {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < p.tp.world && q != p.tp.rank)
                  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(cell[q]) : "l"(cp + (size_t)q * qstride) : "memory");
              all = true;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < p.tp.world && q != p.tp.rank) all = all && ((uint32_t)(cell[q] >> 32) == kk);
              if (!all && clock64() - t0 > 60000000000LL) {
                printf("ocrb fused all-reduce (LL): rank %d never received call %u from every peer (tile %d row %d)\n", p.tp.rank, kk, tile, c0 + i);
                __trap();
              }
            } while (!all);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < p.tp.world && q != p.tp.rank) pb[q] = (uint32_t)cell[q];
          }
          __syncwarp();
          float acc = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < p.tp.world) {
              const uint32_t got = __shfl_sync(0xffffffffu, pb[q], lane & ~1);
              const float v = (q == p.tp.rank) ? own : __uint_as_float(((lane & 1) ? (got >> 16) : (got & 0xffffu)) << 16);
              acc += v;
            }
          if (n_ok) {
            bf16 *xp = p.tp.x + (size_t)(c0 + i) * p.tp.ldx + n;
            *xp = __float2bfloat16_rn(__bfloat162float(*xp) + bf16_round(acc));
          }
        }
      }
      named_bar_sync(1, 128);
      if (et == 0) {
        const int units_per_tile = (p.B + 7) / 8 < SKC_CS ? (p.B + 7) / 8 : SKC_CS;
        const int n_units = (int)(gridDim.x / SKC_CS) * units_per_tile;
        const int t = atomicAdd(&p.tp.ll_seq[1], 1);
        if (t == n_units - 1) {                          // last unit of this call on this rank: publish the call index
          p.tp.ll_seq[1] = 0;
          __threadfence();
          p.tp.ll_seq[0] = (int)kk;
        }
      }
    }
    // ───────────── tensor parallel: all-reduce + residual add of this CTA's columns, in place of a second kernel ─────────────
    // p.D is this rank's peer-mapped partial slot: the bf16 partials above are what HF's rowwise plan sums.  Every CTA
    // that owns live columns is one exchange unit (tile, cluster rank) -- the same units on every rank, since the tiling
    // depends on (N, K, B) only: it announces "my partial of call k is written" to every peer (one remote flag store
    // each), waits for the peers' announcements of the same unit, reads their partials of its 128 rows x <= 8 LG columns
    // straight from peer memory, adds them in rank order in fp32 (identical bits on every rank), rounds to bf16 and adds
    // the residual -- the arithmetic of allreduce_residual_kernel (csrc/comm.cu), so both routes give the same bits.
    // Units exchange independently: a tile's reduction overlaps the weight streaming of the clusters still at work, and
    // the 160 all-reduce launches of a 72B-class decode step disappear.  Slot reuse is safe for the reason given in
    // comm.cu: a peer announces call k + 1 only after griddepcontrol.wait, i.e. after its call k has completely finished.
    if (TP && !(p.tp.mode & 2) && rank * 8 < p.B) {
      const int unit = tile * SKC_CS + rank;
      const int k = p.tp.seq[unit] + 1;
      // the partial stores of all 128 threads happen-before the barrier, the announcing threads' st.release.sys after it:
      // release is cumulative, so one fence per announcing thread (inside the release) orders them all
      if (p.tp.mode & 1) __threadfence_system();
      named_bar_sync(1, 128);
      if (et < p.tp.world) {
        int *dst = p.tp.flags[et] + unit * 8 + p.tp.rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(k) : "memory");
        const int *src = p.tp.flags[p.tp.rank] + unit * 8 + et;
        int v;
        const long long t0 = clock64();
        do {
          asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
          if (v < k && clock64() - t0 > 60000000000LL) {
            printf("ocrb fused all-reduce: rank %d unit %d never saw rank %d announce call %d (flag %d)\n", p.tp.rank, unit, et, k, v);
            __trap();
          }
        } while (v < k);
      }
      named_bar_sync(1, 128);
      if (n_ok) {
#pragma unroll 1
        for (int lg = 0; lg < LG; ++lg) {
          const int c0 = (lg * SKC_CS + rank) * 8;
          if (c0 >= p.B) break;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (c0 + i >= p.B) break;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < p.tp.world) {
                unsigned short raw;
                asm volatile("ld.volatile.global.u16 %0, [%1];" : "=h"(raw) : "l"(p.tp.data[q] + (size_t)(c0 + i) * p.tp.ldp + n));
                acc += __uint_as_float((uint32_t)raw << 16);
              }
            bf16 *xp = p.tp.x + (size_t)(c0 + i) * p.tp.ldx + n;
            *xp = __float2bfloat16_rn(__bfloat162float(*xp) + bf16_round(acc));
          }
        }
      }
      named_bar_sync(1, 128);
      if (et == 0) p.tp.seq[unit] = k;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int BP, int BC, int CS, bool TP = false>
static int launch_skinny_cluster(const CUtensorMap &mw, const CUtensorMap &mx, const SkinnyParams &p, cudaStream_t st) {
  constexpr int ST = (BP > 64) ? SK_ST_BIG : ((BP >= 64) ? 4 : SK_STAGES);
  constexpr size_t smem = (size_t)ST * (SK_W_BYTES + BP * SK_BK * 2) + 1024 /*align*/ + 320 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(skinny_cluster_kernel<BP, BC, CS, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_tiles * CS);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(1) ? 2 : 1;
  OCRB_CUDA(cudaLaunchKernelEx(&cfg, skinny_cluster_kernel<BP, BC, CS, TP>, mw, mx, p));
  return check_launch("skinny_cluster_kernel");
}

// Clusters of size CS that can be resident at once, ONE CTA per SM (GPC packing leaves some SMs unusable for clusters):
// queried for the widest instantiation (BP = 128: one CTA per SM by its shared memory), so the answer -- and with it
// the k-split of a linear -- is the same for every batch size.  0 when the query fails.
template <int CS>
static int cluster_cap() {
  static int n = -1;
  if (n < 0) {
    constexpr size_t smem = (size_t)SK_ST_BIG * (SK_W_BYTES + 128 * SK_BK * 2) + 1024 + 320;
    n = 0;
    if (cudaFuncSetAttribute(skinny_cluster_kernel<128, 128, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CS * 64);
      cfg.blockDim = dim3(SK_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CS;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int m = 0;
      if (cudaOccupancyMaxActiveClusters(&m, skinny_cluster_kernel<128, 128, CS>, &cfg) == cudaSuccess) n = m;
      cudaGetLastError();
    }
    const char *e = getenv("OCRB_SK_CLUSTER_DEBUG");
    if (e && e[0] == '1') fprintf(stderr, "ocrb: resident clusters of %d CTAs (one per SM): %d\n", CS, n);
  }
  return n;
}

// Cluster size for a linear with `tiles` 128-row tiles: the largest of 4 / 3 / 2 whose clusters are all resident in one
// wave (0: keep stream-K).  Depends on the tile count and the device only.
static int pick_cluster_size(int tiles) {
  if (tiles <= cluster_cap<4>()) return 4;
  if (tiles <= cluster_cap<3>()) return 3;
  if (tiles <= cluster_cap<2>()) return 2;
  return 0;
}

template <int BP, int BC, bool TP = false>
static int launch_skinny_cluster_cs(int cs, const CUtensorMap &mw, const CUtensorMap &mx, const SkinnyParams &p, cudaStream_t st) {
  if (cs == 4) return launch_skinny_cluster<BP, BC, 4, TP>(mw, mx, p, st);
  if (cs == 3) return launch_skinny_cluster<BP, BC, 3, TP>(mw, mx, p, st);
  return launch_skinny_cluster<BP, BC, 2, TP>(mw, mx, p, st);
}

template <int BP, int BC>
static int launch_skinny(const CUtensorMap &mw, const CUtensorMap &mx, const SkinnyParams &p, int grid, cudaStream_t st) {
  constexpr int ST = (BP > 64) ? SK_ST_BIG : ((BP >= 64) ? 4 : SK_STAGES);
  constexpr int CC = (BC < 16) ? BC : 16;
  constexpr size_t smem = (size_t)ST * (SK_W_BYTES + BP * SK_BK * 2) + 1024 /*align*/ + 320 /*barriers*/ +
                          128 * CC * sizeof(float) + 64;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<BP, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  OCRB_CUDA(launch_pdl(skinny_gemm_kernel<BP, BC>, dim3(grid), dim3(SK_THREADS), smem, st, mw, mx, p));
  return check_launch("skinny_gemm_kernel");
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ocrb

using namespace ocrb;

static unsigned long long *g_sk_trace = nullptr;
/* debug hook (not in the public header): device buffer [grid][16] of globaltimer stamps for the next launches */
extern "C" void ocrb_skinny_set_trace(void *buf) { g_sk_trace = (unsigned long long *)buf; }

extern "C" int64_t ocrb_skinny_workspace_bytes(void) {
  return (int64_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float) + (int64_t)SK_MAX_GRID * sizeof(int) + 256 +
         (int64_t)SK_MAXBP * SK_MAX_NORM_K * sizeof(bf16);
}

// Set by ocrb_skinny_rowparallel_tp_bf16 around its call of ocrb_skinny_gemm_bf16 (the library is driven by one host
// thread per process): when the linear qualifies for the cluster kernel, the exchange is fused into its epilogue and
// g_tp_fused says so; otherwise the GEMM only writes the partial and the caller launches the separate all-reduce kernel.
static const SkinnyTp *g_tp_ctx = nullptr;
static bool g_tp_fused = false;

extern "C" int ocrb_skinny_gemm_bf16(const void *X, int64_t ldx, const void *W, int64_t ldw, void *D, int64_t ldd, int32_t B,
                                     int32_t N, int32_t K, const void *bias, const void *residual, int64_t ldr,
                                     int32_t epilogue, const void *norm_w, float eps, void *workspace, void *stream) {
  OCRB_REQUIRE(X && W && D && workspace, "skinny_gemm_bf16: null pointer");
  OCRB_REQUIRE(B >= 1 && B <= SK_MAXBP, "skinny_gemm_bf16: B must be in 1..128 (use ocrb_gemm_bf16 for larger batches)");
  OCRB_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0,
               "skinny_gemm_bf16: K and row strides must be multiples of 8");
  OCRB_REQUIRE(((uintptr_t)X & 15) == 0 && ((uintptr_t)W & 15) == 0 && (!norm_w || ((uintptr_t)norm_w & 15) == 0),
               "skinny_gemm_bf16: X, W, norm_w must be 16-byte aligned");
  OCRB_REQUIRE(epilogue >= 0 && epilogue <= 3, "skinny_gemm_bf16: bad epilogue");
  OCRB_REQUIRE(epilogue != OCRB_EPI_RESIDUAL || residual, "skinny_gemm_bf16: residual epilogue without residual");
  OCRB_REQUIRE(epilogue != OCRB_EPI_SWIGLU || N % 128 == 0, "skinny_gemm_bf16: SwiGLU needs packed N % 128 == 0");
  SkinnyParams p;
  memset(&p.tp, 0, sizeof(p.tp));
  p.X = (const bf16 *)X; p.ldx = ldx;
  p.D = (bf16 *)D; p.ldd = ldd;
  p.bias = (const bf16 *)bias;
  p.residual = (epilogue == OCRB_EPI_RESIDUAL) ? (const bf16 *)residual : nullptr;
  p.ldr = ldr;
  p.norm_w = (const bf16 *)norm_w; p.eps = eps;
  p.B = B; p.N = N; p.K = K;
  p.epilogue = epilogue;
  p.num_tiles = cdiv(N, SK_BM);
  p.num_kb = cdiv(K, SK_BK);
  p.trace = g_sk_trace;
  p.partials = (float *)workspace;
  p.flags = (int *)((char *)workspace + (size_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float));
  const long long total = (long long)p.num_tiles * p.num_kb;
  OCRB_REQUIRE(total < (1ll << 30), "skinny_gemm_bf16: problem too large (tiles x k-blocks must be < 2^30)");
  // one CTA per SM; spans of at least 4 k-blocks so tiny problems do not pay 148 fix-ups.
  // The split depends only on (N, K) and the SM count -- never on B -- so results are batch-invariant.
  int grid = sm_count();
  if (grid > SK_MAX_GRID) grid = SK_MAX_GRID;
  if (total / 4 < grid) grid = (int)(total / 4 > 1 ? total / 4 : 1);
  {
    static int grid_override = -1;          // experiments only (scripts/exp): OCRB_SK_GRID
    if (grid_override < 0) {
      const char *e = getenv("OCRB_SK_GRID");
      grid_override = e ? atoi(e) : 0;
    }
    if (grid_override > 0 && grid_override < grid) grid = grid_override;
  }
  CUtensorMap mw;
  int rc = make_tensor_map_bf16(&mw, W, N, K, ldw, SK_BM);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int BPsel = B <= 16 ? 16 : (B <= 32 ? 32 : (B <= 64 ? 64 : (B <= 96 ? 96 : 128)));
  if (norm_w) {
    // normalise the rows once (same rstd routine as the fused path), then treat them as a plain input
    OCRB_REQUIRE(K <= SK_MAX_NORM_K, "skinny_gemm_bf16: the RMSNorm prologue supports K <= 32768");
    bf16 *xn = (bf16 *)((char *)workspace + (size_t)SK_MAX_GRID * SK_MAXBP * 128 * sizeof(float) +
                        (size_t)SK_MAX_GRID * sizeof(int) + 256);
    OCRB_CUDA(launch_pdl(skinny_norm_rows_kernel, dim3(cdiv(B, 8)), dim3(256), 0, st, p, xn));
    rc = check_launch("skinny_norm_rows_kernel");
    if (rc) return rc;
    p.X = xn;
    p.ldx = K;
    p.norm_w = nullptr;
    X = xn;
    ldx = K;
    norm_w = nullptr;
  }
  CUtensorMap mx;
  rc = make_tensor_map_bf16(&mx, X, B, K, ldx, BPsel);
  if (rc) return rc;
  // Narrow linears (few tiles, K of at least 16 k-blocks, no SwiGLU pairing): a cluster of 4 / 3 / 2 CTAs per tile with the
  // split-K reduction through distributed shared memory.  The choice depends on (N, K) and the device only.
  {
    static int use_cluster = -1;
    if (use_cluster < 0) {
      const char *e = getenv("OCRB_SK_CLUSTER");
      use_cluster = e ? atoi(e) : 1;
    }
    const int cs = (use_cluster && epilogue != OCRB_EPI_SWIGLU && p.num_tiles * 2 <= sm_count() && p.num_kb >= 16)
                       ? pick_cluster_size(p.num_tiles) : 0;
    if (cs > 0) {
      if (g_tp_ctx && epilogue == OCRB_EPI_NONE && !bias && p.num_tiles * cs <= SK_TP_UNITS && B <= 64) {
        // tensor-parallel instantiation: the exchange is the epilogue (decode batches only)
        p.tp = *g_tp_ctx;
        p.epilogue = SK_EPI_TP;
        g_tp_fused = true;
        if (B <= 8) return launch_skinny_cluster_cs<16, 8, true>(cs, mw, mx, p, st);
        if (B <= 16) return launch_skinny_cluster_cs<16, 16, true>(cs, mw, mx, p, st);
        if (B <= 32) return launch_skinny_cluster_cs<32, 32, true>(cs, mw, mx, p, st);
        return launch_skinny_cluster_cs<64, 64, true>(cs, mw, mx, p, st);
      }
      if (B <= 8) return launch_skinny_cluster_cs<16, 8>(cs, mw, mx, p, st);
      if (B <= 16) return launch_skinny_cluster_cs<16, 16>(cs, mw, mx, p, st);
      if (B <= 32) return launch_skinny_cluster_cs<32, 32>(cs, mw, mx, p, st);
      if (B <= 64) return launch_skinny_cluster_cs<64, 64>(cs, mw, mx, p, st);
      if (B <= 96) return launch_skinny_cluster_cs<96, 96>(cs, mw, mx, p, st);
      return launch_skinny_cluster_cs<128, 128>(cs, mw, mx, p, st);
    }
  }
  if (B <= 4) return launch_skinny<16, 4>(mw, mx, p, grid, st);
  if (B <= 8) return launch_skinny<16, 8>(mw, mx, p, grid, st);
  if (B <= 16) return launch_skinny<16, 16>(mw, mx, p, grid, st);
  if (B <= 32) return launch_skinny<32, 32>(mw, mx, p, grid, st);
  if (B <= 64) return launch_skinny<64, 64>(mw, mx, p, grid, st);
  // 65..128 rows (a folder batch: 32 pages x 3 candidates): still ONE pass over the weights; 28 / 32 KiB stages, one CTA per SM
  if (B <= 96) return launch_skinny<96, 96>(mw, mx, p, grid, st);
  return launch_skinny<128, 128>(mw, mx, p, grid, st);
}

extern "C" int ocrb_allreduce_residual_bf16(void *x, int64_t ldx, const void *const *data_ptrs, void *const *flag_ptrs,
                                            int32_t world, int32_t rank, int32_t *seq, int32_t rows, int32_t dim,
                                            int64_t ld_part, void *stream);

extern "C" int ocrb_skinny_rowparallel_tp_bf16(const void *X, int64_t ldx, const void *W, int64_t ldw, int32_t B, int32_t N,
                                               int32_t K, void *x, int64_t ldx_res, const void *const *data_ptrs,
                                               int64_t ld_part, void *const *fused_flag_ptrs, int32_t *fused_seq,
                                               void *const *flag_ptrs, int32_t *seq, void *const *ll_ptrs, int32_t *ll_seq,
                                               int32_t world, int32_t rank, void *workspace, int32_t allow_fused,
                                               void *stream) {
  OCRB_REQUIRE(x && data_ptrs && fused_flag_ptrs && fused_seq && flag_ptrs && seq, "skinny_rowparallel_tp_bf16: null pointer");
  OCRB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "skinny_rowparallel_tp_bf16: bad world/rank");
  OCRB_REQUIRE(N % 8 == 0 && ldx_res % 8 == 0 && ld_part % 8 == 0, "skinny_rowparallel_tp_bf16: N and strides must be multiples of 8");
  SkinnyTp tp;
  memset(&tp, 0, sizeof(tp));
  for (int r = 0; r < world; ++r) {
    tp.data[r] = (const bf16 *)data_ptrs[r];
    tp.flags[r] = (int *)fused_flag_ptrs[r];
  }
  tp.seq = fused_seq;
  tp.x = (bf16 *)x;
  tp.ldx = ldx_res;
  tp.ldp = ld_part;
  tp.world = world;
  tp.rank = rank;
  {
    static int mode = -1;
    if (mode < 0) {
      const char *e = getenv("OCRB_TP_FUSED_MODE");
      mode = e ? atoi(e) : 0;
    }
    tp.mode = mode & 1;
  }
  if (allow_fused == 2) {
    OCRB_REQUIRE(ll_ptrs && ll_seq && B <= SK_LL_ROWS && ld_part % 2 == 0, "skinny_rowparallel_tp_bf16: LL route needs its buffers and B <= 64");
    for (int r = 0; r < world; ++r) tp.ll[r] = (unsigned long long *)ll_ptrs[r];
    tp.ll_seq = ll_seq;
    tp.mode |= 2;
  }
  g_tp_ctx = allow_fused ? &tp : nullptr;
  g_tp_fused = false;
  const int rc = ocrb_skinny_gemm_bf16(X, ldx, W, ldw, (void *)data_ptrs[rank], ld_part, B, N, K, nullptr, nullptr, 0,
                                       OCRB_EPI_NONE, nullptr, 0.f, workspace, stream);
  g_tp_ctx = nullptr;
  if (rc) return rc;
  if (g_tp_fused) return OCRB_OK;
  return ocrb_allreduce_residual_bf16(x, ldx_res, data_ptrs, flag_ptrs, world, rank, seq, B, N, ld_part, stream);
}

/* 1 when the last ocrb_skinny_rowparallel_tp_bf16 call ran the exchange inside the GEMM (debug / tests). */
extern "C" int ocrb_skinny_rowparallel_tp_was_fused(void) { return g_tp_fused ? 1 : 0; }
