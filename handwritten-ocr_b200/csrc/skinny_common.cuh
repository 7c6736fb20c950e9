// Device helpers shared by the weight-streaming decode GEMMs (skinny.cu: one linear per launch; chain.cu: a chain of
// dependent linears in one persistent launch).
#pragma once
#include "tc_common.cuh"
#include <math.h>

namespace ocrb {

constexpr int SK_BM = 128;          // weight rows per tile
constexpr int SK_BK = 64;           // k-block: 64 bf16 = one 128-byte swizzle row
#ifndef SK_STAGES_N
#define SK_STAGES_N 5
#endif
constexpr int SK_STAGES = SK_STAGES_N;           // 5 x 18 KiB (BP=16): two CTAs of consecutive kernels fit one SM under PDL
constexpr int SK_THREADS = 192;
constexpr int SK_MAX_GRID = 296;    // workspace slots (2 x 148)
constexpr int SK_MAXBP = 128;
constexpr int SK_MAX_NORM_K = 32768;  // widest row the B > 16 RMSNorm scratch holds
constexpr uint32_t SK_W_BYTES = SK_BM * SK_BK * 2;   // 16 KiB

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t *r);
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}

__device__ __forceinline__ float sk_silu(float g) { return bf16_round(g / (1.0f + expf(-g))); }
__device__ __forceinline__ float sk_gelu(float x) { return bf16_round(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f))); }

__device__ __forceinline__ void unpack8f(const uint4 &raw, float *f) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}

// Span of the (tile, k-block) iteration space owned by CTA `cta` (all int32: total units < 2^31 is checked on the host).
__device__ __forceinline__ void sk_span(int cta, int grid, int total, int &begin, int &end) {
  const int per = total / grid, rem = total % grid;
  begin = cta * per + (cta < rem ? cta : rem);
  end = begin + per + (cta < rem ? 1 : 0);
}

// Processing order inside a span: tiles in REVERSE order (k-blocks ascending inside a tile).  The tail segment --
// the only one that may end before its tile does -- is therefore computed and published first, and the head
// segment -- the only one that may need other CTAs' partials -- last: nobody waits on a CTA that itself waits.
struct SkSpan {
  int u_begin, u_end, KB, T0, T1;
  __device__ __forceinline__ void init(int cta, int grid, int num_tiles, int kb) {
    KB = kb;
    sk_span(cta, grid, num_tiles * kb, u_begin, u_end);
    T0 = u_begin / KB;
    T1 = (u_end - 1) / KB;
  }
  __device__ __forceinline__ int num_units() const { return u_end - u_begin; }
  __device__ __forceinline__ int num_segs() const { return T1 - T0 + 1; }
  // segment i in processing order -> tile, first k-block, number of k-blocks
  __device__ __forceinline__ void seg(int i, int &tile, int &kb0, int &nkb) const {
    tile = T1 - i;
    const int ts = tile * KB;
    const int a = ts > u_begin ? ts : u_begin;
    const int b = (ts + KB) < u_end ? (ts + KB) : u_end;
    kb0 = a - ts;
    nkb = b - a;
  }
};

// Walks the units of a span in processing order without divisions.
struct SkCursor {
  int seg, tile, kb, kb_end;
  bool valid;
  __device__ __forceinline__ void init(const SkSpan &sp) {
    seg = 0;
    valid = sp.num_units() > 0;
    int kb0 = 0, nkb = 0;
    tile = 0;
    if (valid) sp.seg(0, tile, kb0, nkb);
    kb = kb0;
    kb_end = kb0 + nkb;
  }
  __device__ __forceinline__ void advance(const SkSpan &sp) {
    if (++kb == kb_end) {
      if (++seg < sp.num_segs()) {
        int kb0, nkb;
        sp.seg(seg, tile, kb0, nkb);
        kb = kb0;
        kb_end = kb0 + nkb;
      } else {
        valid = false;
      }
    }
  }
};

// HF RMSNorm statistics (modeling_qwen2_5_vl.py:66-71): rstd[b] = rsqrt(mean(x[b]^2) + eps) in fp32.  Warps 2..9
// (256 threads) share the rows; one warp per row, up to 16 independent 16-byte loads in flight per lane, fixed
// summation order (lane-strided partial sums, xor-shuffle tree) so the result never depends on B.
__device__ __forceinline__ float sk_one_row_rstd(const bf16 *xr, int K, float eps, int lane) {
  const int kvec = K >> 3;
  float ss = 0.f;
  for (int v0 = lane; v0 < kvec; v0 += 32 * 16) {
    uint4 raw[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int v = v0 + i * 32;
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
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  return rsqrtf(ss / (float)K + eps);
}

}  // namespace ocrb
