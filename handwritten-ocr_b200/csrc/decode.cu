// Batched greedy decode hot path (HF generation loop, utils.py:2743-2806, one token per sequence per step): paged
// split-KV attention over the cache.  (The weight-streaming linears of the step are in skinny.cu.)
//
// HBM-bound: every cached K and V byte of every sequence is read exactly once per step (B * ctx * 57 344 B for the 7B
// config over 28 layers).  Layout and schedule:
//   * cache layout [page][kv head][16 tokens][hd]: the K (or V) rows a 16-key tile needs are ONE contiguous 4 KiB block,
//     fetched by the TMA unit (cp.async.bulk.tensor, two [16 x 64]-column boxes, 128B swizzle) straight into shared memory;
//   * work item = (sequence, KV head, range of `chunk` keys).  The products are computed TRANSPOSED so that the wide side
//     of mma.sync.m16n8k16 (M = 16) is the 16 keys of a tile and the narrow side (N = 8) the G <= 8 query heads that
//     share the KV head: S^T[key][head] = K . Q^T and O^T[dim][head] += V^T . P^T -- 16 mma per 16-key tile for G <= 8
//     (32 with the heads on the M side, 9 of 16 rows dead at G = 7); P^T reaches its operand layout through two
//     movmatrix transposes;
//   * every WARP is an independent flash-decoding worker: it owns whole work items (item = warp * grid + cta, then
//     strided), streams their 16-key tiles through its private shared-memory ring (lane 0 issues the TMA loads, an
//     mbarrier per slot counts the bytes), keeps the running (max, sum, O[16 x hd]) in registers and writes one fp32
//     partial (max, sum, o[hd]) per head and item -- there is no block-level barrier anywhere in the kernel, so the
//     memory pipe of an SM never drains while one warp finishes an item and another is in the middle of its own;
//   * decode_attn_combine_kernel folds the partials of a sequence's key ranges.
// mRoPE of the new token's q / k (bf16 arithmetic, as HF) and the append of its k, v to the cache are fused: the warp
// whose tile holds position ctx builds that row in shared memory and writes it to the cache.
// The arithmetic of a sequence depends only on (its context length, chunk) -- never on B or on the other sequences.
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ocrb {

__device__ __forceinline__ float rope_elem_bf16(const bf16 *vec, int i, int hd, const bf16 *c, const bf16 *s) {
  const int half = hd >> 1;
  const float x = __bfloat162float(vec[i]);
  const float other = (i < half) ? -__bfloat162float(vec[i + half]) : __bfloat162float(vec[i - half]);
  const float t = bf16_round(x * __bfloat162float(c[i]));
  const float u = bf16_round(other * __bfloat162float(s[i]));
  return bf16_round(t + u);
}

constexpr int DA_TILE = 16;            // keys per tile (one k-step of the P.V product) = one (page, kv head) block at page size 16
// workers (warps) per CTA x per-warp TMA ring depth: 6 x 4 keeps four 8 KiB tiles in flight per worker (a lone worker is
// latency-bound by its ring depth: tile time ~0.4 us against ~1.5 us of memory latency); 8 x 2 is the experiment (OCRB_ATTN_CFG=82)
constexpr int DA_MAXG = 16;            // query heads per KV head (rows of the m16 tile)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// Shared memory per warp: Q staging [16][HD + 8] bf16, then (1024-byte aligned) the ring of tiles.  One tile slot:
// K atoms then V atoms; an atom = [16 keys][64 dims] bf16 = 2 KiB in the 128B-swizzled layout the TMA unit writes
// (16-byte chunk c of row r sits at chunk c ^ (r & 7)).
template <int HD, int DA_WARPS, int DA_STAGES>
struct DaSmem {
  static constexpr int ATOMS = HD / 64;                      // 64-dim column groups per K (or V) tile
  static constexpr uint32_t ATOM_BYTES = DA_TILE * 128;
  static constexpr uint32_t TILE_BYTES = 2 * ATOMS * ATOM_BYTES;      // K + V
  static constexpr int QPITCH = HD + 8;
  static constexpr uint32_t Q_BYTES = 16 * QPITCH * 2;
  static constexpr uint32_t RING_BYTES = DA_WARPS * DA_STAGES * TILE_BYTES;
  static constexpr size_t BYTES = 1024 + RING_BYTES + DA_WARPS * Q_BYTES + DA_WARPS * DA_STAGES * 8;
};

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// NT = 8-head column tiles of the transposed products (1: up to 8 query heads per KV head, 2: up to 16)
template <int HD, int DA_WARPS, int DA_STAGES, int NT>
__global__ void __launch_bounds__(DA_WARPS * 32, 1)
decode_attn_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                   const bf16 *__restrict__ qkv, long long ldqkv, bf16 *__restrict__ k_cache, bf16 *__restrict__ v_cache,
                   const int32_t *__restrict__ block_table, int max_pages, const int32_t *__restrict__ ctx_len, int B,
                   int page_size, int n_q, int n_kv, const bf16 *__restrict__ cosT, const bf16 *__restrict__ sinT,
                   float scale, float *__restrict__ split_ws, int n_splits, int chunk) {
  using SM = DaSmem<HD, DA_WARPS, DA_STAGES>;
  constexpr int QP = SM::QPITCH;
  constexpr int KK = HD / 16;                                 // k-steps of K . Q^T = 16-dim row tiles of O^T
  extern __shared__ uint8_t da_smem_raw[];
  uint8_t *smem = da_smem_raw + ((1024u - (smem_u32(da_smem_raw) & 1023u)) & 1023u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *ring = smem + (size_t)warp * DA_STAGES * SM::TILE_BYTES;
  bf16 *sQ = reinterpret_cast<bf16 *>(smem + SM::RING_BYTES + (size_t)warp * SM::Q_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::RING_BYTES + DA_WARPS * SM::Q_BYTES) + warp * DA_STAGES;
  const int G = n_q / n_kv;
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
#pragma unroll
    for (int s = 0; s < DA_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows >= G of the Q tile stay zero for the whole kernel
  for (int i = lane; i < 16 * QP / 2; i += 32) reinterpret_cast<uint32_t *>(sQ)[i] = 0u;
  __syncwarp();

  const int n_items = B * n_kv * n_splits;
  const int m = lane >> 3, l8 = lane & 7;
  const int r0 = lane >> 2, cq = (lane & 3) * 2;
  const uint32_t ring_s = smem_u32(ring);
  uint32_t phase_bits = 0;                                    // bit s = parity of the next completion of ring slot s
  bool waited = false;                                        // griddepcontrol.wait executed (first item only)

  // items in the order (key range, sequence, kv head): neighbouring workers get the same range of different sequences,
  // so live and dead ranges are spread evenly; worker id = warp * grid + cta spreads a small batch over all SMs
  for (int item = warp * (int)gridDim.x + (int)blockIdx.x; item < n_items; item += DA_WARPS * (int)gridDim.x) {
    const int split = item / (B * n_kv);
    const int pair = item - split * (B * n_kv);
    const int b = pair / n_kv, kvh = pair - b * n_kv;
    const int ctx = ctx_len[b];          // written by the previous step's argmax kernel: complete long before this launch
    const int total = ctx + 1;
    const int k0 = split * chunk;
    const int k1 = min(total, k0 + chunk);
    float *ws = split_ws + (((size_t)b * n_q + (size_t)kvh * G) * n_splits + split) * (HD + 2);
    const size_t ws_head = (size_t)n_splits * (HD + 2);
    if (k0 >= k1) {
      if (!waited) { pdl_wait(); waited = true; }             // split_ws may still be read by the previous layer's combine kernel
      for (int g = lane; g < G; g += 32) { ws[g * ws_head] = -INFINITY; ws[g * ws_head + 1] = 0.f; }
      continue;
    }
    const int n_tiles = (k1 - k0 + DA_TILE - 1) / DA_TILE;
    const int32_t *bt = block_table + (size_t)b * max_pages;

    // one tile = the K and the V block of 16 consecutive keys of this kv head: 2 * ATOMS boxes of [16 rows x 64 dims]
    auto issue_tile = [&](int i) {
      if (i < n_tiles && lane == 0) {
        const int s = i % DA_STAGES;
        const int key0 = k0 + i * DA_TILE;
        const int row = (bt[key0 / page_size] * n_kv + kvh) * page_size + key0 % page_size;
        uint8_t *dst = ring + (size_t)s * SM::TILE_BYTES;
        mbar_expect_tx(&bars[s], SM::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < SM::ATOMS; ++a) {
          tma_load_2d(dst + a * SM::ATOM_BYTES, &map_k, &bars[s], a * 64, row);
          tma_load_2d(dst + (SM::ATOMS + a) * SM::ATOM_BYTES, &map_v, &bars[s], a * 64, row);
        }
      }
    };
#pragma unroll
    for (int i = 0; i < DA_STAGES; ++i) issue_tile(i);

    if (!waited) { pdl_wait(); waited = true; }               // qkv comes from the preceding skinny GEMM
    const bf16 *row = qkv + (size_t)b * ldqkv;
    const bf16 *c = cosT + (size_t)b * HD, *s_ = sinT + (size_t)b * HD;
    // mRoPE of the G query heads into the staging tile, 8 elements (16 bytes) per lane and step: the element chunk, its
    // rotate-half partner chunk (+- HD / 2) and the cos / sin chunks of up to 4 steps are fetched together -- one L2 round trip
    // per item instead of one per element (the scalar loop cost ~3 us of a ~7 us item).  Same arithmetic as rope_elem_bf16.
    {
      constexpr int CH = HD / 8;                               // 16-byte chunks per head
      for (int i0 = 0; i0 < G * CH; i0 += 4 * 32) {
        uint4 xv[4], ov[4], cv[4], sv4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = i0 + u * 32 + lane;
          if (idx < G * CH) {
            const int g = idx / CH, ch = idx - g * CH;
            const bf16 *qh = row + (size_t)(kvh * G + g) * HD;
            xv[u] = __ldg(reinterpret_cast<const uint4 *>(qh + ch * 8));
            ov[u] = __ldg(reinterpret_cast<const uint4 *>(qh + (ch ^ (CH / 2)) * 8));
            cv[u] = __ldg(reinterpret_cast<const uint4 *>(c + ch * 8));
            sv4[u] = __ldg(reinterpret_cast<const uint4 *>(s_ + ch * 8));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = i0 + u * 32 + lane;
          if (idx < G * CH) {
            const int g = idx / CH, ch = idx - g * CH;
            const float sgn = (ch < CH / 2) ? -1.f : 1.f;
            const bf16 *xe = reinterpret_cast<const bf16 *>(&xv[u]), *oe = reinterpret_cast<const bf16 *>(&ov[u]);
            const bf16 *ce = reinterpret_cast<const bf16 *>(&cv[u]), *se = reinterpret_cast<const bf16 *>(&sv4[u]);
            uint4 outv;
            bf16 *re = reinterpret_cast<bf16 *>(&outv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = bf16_round(__bfloat162float(xe[e]) * __bfloat162float(ce[e]));
              const float uu = bf16_round(sgn * __bfloat162float(oe[e]) * __bfloat162float(se[e]));
              re[e] = __float2bfloat16_rn(bf16_round(t + uu));
            }
            *reinterpret_cast<uint4 *>(sQ + g * QP + ch * 8) = outv;
          }
        }
      }
    }
    __syncwarp();
    uint32_t qf[HD / 16][4];
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) ldmatrix_x4(qf[kk], smem_u32(sQ + ((m & 1) * 8 + l8) * QP + kk * 16 + (m >> 1) * 8));
    // transposed accumulators: o[mt][nt] = O^T tile of dims mt*16 .. +15 x heads nt*8 .. +7
    float o[KK][NT][4];
#pragma unroll
    for (int mt = 0; mt < KK; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f;
    // running max / sum of the heads this lane sees: head = nt * 8 + cq + j (replicated over the 8 lanes with the same cq)
    float run_m[NT][2], run_l[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { run_m[nt][0] = run_m[nt][1] = -INFINITY; run_l[nt][0] = run_l[nt][1] = 0.f; }

    for (int i = 0; i < n_tiles; ++i) {
      const int s = i % DA_STAGES;
      mbar_wait(&bars[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      const int key0 = k0 + i * DA_TILE;
      const uint32_t sk = ring_s + (uint32_t)s * SM::TILE_BYTES, sv = sk + SM::ATOMS * SM::ATOM_BYTES;
      if (ctx >= key0 && ctx < key0 + DA_TILE) {
        // this tile holds the new token: rope its key, place k / v in the (swizzled) tile and append them to the cache
        const bf16 *knew = row + (size_t)n_q * HD + (size_t)kvh * HD;
        const bf16 *vnew = row + (size_t)(n_q + n_kv) * HD + (size_t)kvh * HD;
        const int r = ctx - key0;
        const size_t dst = ((size_t)(bt[ctx / page_size] * n_kv + kvh) * page_size + ctx % page_size) * HD;
        uint8_t *tk = ring + (size_t)s * SM::TILE_BYTES, *tv = tk + SM::ATOMS * SM::ATOM_BYTES;
        for (int d = lane; d < HD; d += 32) {
          const bf16 kr = __float2bfloat16_rn(rope_elem_bf16(knew, d, HD, c, s_));
          const bf16 vr = vnew[d];
          const uint32_t off = (uint32_t)(d >> 6) * SM::ATOM_BYTES + r * 128 + ((((d & 63) >> 3) ^ (r & 7)) << 4) + (d & 7) * 2;
          *reinterpret_cast<bf16 *>(tk + off) = kr;
          *reinterpret_cast<bf16 *>(tv + off) = vr;
          k_cache[dst + d] = kr;
          v_cache[dst + d] = vr;
        }
        __syncwarp();
      }
      // ---- S^T[16 keys][heads] = K . Q^T: the K tile is the M side, the heads the N side ----
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      {
        const int kr = (m >> 1) * 8 + l8;                     // key row this lane addresses
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          uint32_t bb[4];
          const int ch = (kk & 3) * 2 + (m & 1);              // 16-byte chunk inside the 64-dim atom
          ldmatrix_x4(bb, sk + (kk >> 2) * SM::ATOM_BYTES + kr * 128 + ((ch ^ (kr & 7)) << 4));
          // bb = (keys 0-7, dims 0-7), (keys 0-7, dims 8-15), (keys 8-15, dims 0-7), (keys 8-15, dims 8-15) of this k-step:
          // as an A operand the middle two swap
          const uint32_t ka[4] = {bb[0], bb[2], bb[1], bb[3]};
          mma_bf16_16816(acc[0], ka, qf[kk][0], qf[kk][2]);   // heads 0-7: rows 0-7 of the Q tile
          if (NT > 1) mma_bf16_16816(acc[NT - 1], ka, qf[kk][1], qf[kk][3]);
        }
      }
      // ---- online softmax per head: this thread holds keys r0 (e = 0, 1) and r0 + 8 (e = 2, 3) of heads nt*8 + cq + (e & 1) ----
      const bool ok0 = key0 + r0 < total, ok1 = key0 + r0 + 8 < total;
      uint32_t pb[NT][2];
      bool moved = false;
      float corr[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float tmax[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          acc[nt][j] = ok0 ? acc[nt][j] * scale : -INFINITY;
          acc[nt][2 + j] = ok1 ? acc[nt][2 + j] * scale : -INFINITY;
          tmax[j] = fmaxf(acc[nt][j], acc[nt][2 + j]);
          tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 4));
          tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 8));
          tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 16));
          const float mn = fmaxf(run_m[nt][j], tmax[j]);      // finite: every processed tile has a live key
          corr[nt][j] = __expf(run_m[nt][j] - mn);            // exp(-inf) = 0 on the first tile
          run_m[nt][j] = mn;
          run_l[nt][j] *= corr[nt][j];
          moved |= corr[nt][j] != 1.0f;
        }
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[e] = __expf(acc[nt][e] - run_m[nt][e & 1]);
          run_l[nt][e & 1] += pv[e];
        }
        // P^T as the B operand of O^T += V^T . P^T needs (key pair, head) per lane: transpose the two 8 x 8 blocks
        pb[nt][0] = movmatrix_trans(pack_bf16(pv[0], pv[1]));
        pb[nt][1] = movmatrix_trans(pack_bf16(pv[2], pv[3]));
      }
      // ---- O^T = O^T * corr + V^T . P^T (the scaling is skipped when no head of the warp moved its maximum) ----
      {
        const bool rescale = __any_sync(0xffffffffu, moved);
        const int vr = (m & 1) * 8 + l8;                      // key row this lane addresses
#pragma unroll
        for (int jj = 0; jj < KK; ++jj) {
          uint32_t bb[4];
          const int ch = (jj & 3) * 2 + (m >> 1);
          ldmatrix_x4_trans(bb, sv + (jj >> 2) * SM::ATOM_BYTES + vr * 128 + ((ch ^ (vr & 7)) << 4));
          // transposed 8 x 8 blocks: (dims 0-7, keys 0-7), (dims 0-7, keys 8-15), (dims 8-15, keys 0-7), (dims 8-15, keys 8-15)
          const uint32_t va[4] = {bb[0], bb[2], bb[1], bb[3]};
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float (&oo)[4] = o[jj][nt];
            if (rescale) { oo[0] *= corr[nt][0]; oo[1] *= corr[nt][1]; oo[2] *= corr[nt][0]; oo[3] *= corr[nt][1]; }
            mma_bf16_16816(oo, va, pb[nt][0], pb[nt][1]);
          }
        }
      }
      __syncwarp();                      // every lane is done with this ring slot
      issue_tile(i + DA_STAGES);
    }
    // ---- one fp32 partial (max, sum, o[hd]) per head of this item, straight from the registers ----
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        run_l[nt][j] += __shfl_xor_sync(0xffffffffu, run_l[nt][j], 4);
        run_l[nt][j] += __shfl_xor_sync(0xffffffffu, run_l[nt][j], 8);
        run_l[nt][j] += __shfl_xor_sync(0xffffffffu, run_l[nt][j], 16);
      }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int h = nt * 8 + cq + j;                        // head of this lane's columns
        if (h < G) {
          float *wh = ws + (size_t)h * ws_head;
#pragma unroll
          for (int mt = 0; mt < KK; ++mt) {
            wh[2 + mt * 16 + r0] = o[mt][nt][j];
            wh[2 + mt * 16 + r0 + 8] = o[mt][nt][2 + j];
          }
          if (r0 == 0) { wh[0] = run_m[nt][j]; wh[1] = run_l[nt][j]; }
        }
      }
    __syncwarp();                        // the Q staging rows are rewritten by the next item
  }
  if (!waited) pdl_wait();               // a worker without items still has to honour the dependency before it exits
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

// 2-D view of one layer's K (or V) cache: rows = n_cache_pages * n_kv * page_size tokens, cols = hd; box [16 x 64].
static int make_cache_map(CUtensorMap *m, const void *cache, long long rows, int hd) {
  return make_tensor_map_bf16(m, cache, rows, hd, hd, DA_TILE);
}

template <int HD, int WARPS, int STAGES, int NT>
static int launch_decode_attn(const CUtensorMap &mk, const CUtensorMap &mv, const bf16 *qkv, long long ldqkv, bf16 *kc, bf16 *vc,
                              const int32_t *bt, int max_pages, const int32_t *ctx_len, int B, int page_size, int n_q, int n_kv,
                              const bf16 *cosT, const bf16 *sinT, float scale, float *split_ws, int n_splits, int chunk,
                              cudaStream_t st) {
  constexpr size_t smem = DaSmem<HD, WARPS, STAGES>::BYTES;
  static_assert(smem <= 227 * 1024, "decode attention: shared memory budget");
  static bool attr_set = false;
  static int n_sm = 0;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(decode_attn_kernel<HD, WARPS, STAGES, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
    attr_set = true;
  }
  const long long items = (long long)B * n_kv * n_splits;
  const int grid = (int)(items < n_sm ? items : n_sm);       // persistent: one CTA of independent warps per SM
  OCRB_CUDA(launch_pdl_bit(2, decode_attn_kernel<HD, WARPS, STAGES, NT>, dim3(grid), dim3(WARPS * 32), smem, st, mk, mv, qkv, ldqkv,
                           kc, vc, bt, max_pages, ctx_len, B, page_size, n_q, n_kv, cosT, sinT, scale, split_ws, n_splits, chunk));
  return check_launch("decode_attn_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_decode_attention(const void *qkv, int64_t ldqkv, void *k_cache, void *v_cache, int32_t n_cache_pages,
                                     const int32_t *block_table, int32_t max_pages, const int32_t *ctx_len, int32_t B,
                                     int32_t page_size, int32_t n_q, int32_t n_kv, int32_t hd, const void *cosT,
                                     const void *sinT, float scale, void *out, int64_t ldo, float *split_ws,
                                     int32_t n_splits, void *stream) {
  OCRB_REQUIRE(qkv && k_cache && v_cache && block_table && ctx_len && cosT && sinT && out && split_ws,
               "decode_attention: null pointer");
  OCRB_REQUIRE(B > 0 && n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= DA_MAXG && (hd == 64 || hd == 128) && n_splits > 0,
               "decode_attention: needs n_q / n_kv <= 16 query heads per KV head and hd of 64 or 128");
  OCRB_REQUIRE(page_size > 0 && page_size % DA_TILE == 0, "decode_attention: page_size must be a multiple of 16");
  OCRB_REQUIRE(((uintptr_t)qkv & 15) == 0 && ldqkv % 8 == 0 && ((uintptr_t)cosT & 15) == 0 && ((uintptr_t)sinT & 15) == 0,
               "decode_attention: qkv rows and the rope tables must be 16-byte aligned");
  OCRB_REQUIRE(n_cache_pages > 0 && ((uintptr_t)k_cache & 15) == 0 && ((uintptr_t)v_cache & 15) == 0,
               "decode_attention: cache pointers must be 16-byte aligned, n_cache_pages > 0");
  const int max_ctx = max_pages * page_size;
  const int chunk = cdiv(cdiv(max_ctx, n_splits), DA_TILE) * DA_TILE;     // keys per work item, whole tiles
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)n_cache_pages * n_kv * page_size;
  CUtensorMap mk, mv;
  int rc = make_cache_map(&mk, k_cache, rows, hd);
  if (rc) return rc;
  rc = make_cache_map(&mv, v_cache, rows, hd);
  if (rc) return rc;
  // workers per CTA x ring depth: 6 x 4 while a worker has one or two items (small batches: latency of the ring), 8 x 2 once
  // every worker has several (B = 96: 225 vs 231 us per layer -- more warps hide the serial mma / softmax chain of a tile).
  // A sequence's arithmetic does not depend on the choice.
  static int cfg_env = -1;
  if (cfg_env < 0) {
    const char *e = getenv("OCRB_ATTN_CFG");
    cfg_env = e ? atoi(e) : 0;
  }
  const int cfg = cfg_env ? cfg_env : (((long long)B * n_kv * n_splits >= 3LL * 148 * 6) ? 82 : 64);
#define DA_LAUNCH(HD_, W_, S_)                                                                                              \
  (n_q / n_kv > 8 ? DA_LAUNCH_NT(HD_, W_, S_, 2) : DA_LAUNCH_NT(HD_, W_, S_, 1))
#define DA_LAUNCH_NT(HD_, W_, S_, NT_)                                                                                      \
  launch_decode_attn<HD_, W_, S_, NT_>(mk, mv, (const bf16 *)qkv, (long long)ldqkv, (bf16 *)k_cache, (bf16 *)v_cache, block_table, \
                                  (int)max_pages, ctx_len, (int)B, (int)page_size, (int)n_q, (int)n_kv, (const bf16 *)cosT,    \
                                  (const bf16 *)sinT, scale, split_ws, (int)n_splits, chunk, st)
  if (hd == 128)
    rc = (cfg == 82) ? DA_LAUNCH(128, 8, 2) : ((cfg == 102) ? DA_LAUNCH(128, 10, 2) : DA_LAUNCH(128, 6, 4));
  else
    rc = DA_LAUNCH(64, 8, 4);
#undef DA_LAUNCH
#undef DA_LAUNCH_NT
  if (rc) return rc;
  OCRB_CUDA(launch_pdl_bit(4, decode_attn_combine_kernel, dim3(n_q, B), dim3(hd), 0, st, (const float *)split_ws, (int)n_q,
                           (int)n_splits, (int)hd, (bf16 *)out, (long long)ldo));
  return check_launch("decode_attn_combine_kernel");
}
