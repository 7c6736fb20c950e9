// Flash attention on the 5th-generation tensor cores for the two long-sequence attentions of a read:
//   * vision tower full-attention blocks (HF modeling_qwen2_5_vl.py:231-283: non-causal, 16 heads, hd 80, S = 3 996 per image);
//   * decoder prefill (HF :718-760: causal, GQA 28 / 4 heads, hd 128, T ~ 1 036 per sequence).
// (The vision tower's windowed blocks -- sequences of <= 64 tokens -- stay on flash_varlen_kernel, attention.cu.)
//
// One CTA = one (sequence, query head, 256 consecutive queries) = TWO 128-row query tiles that share every K / V block:
//   warp 0      : TMA producer -- Q tiles once, then K_j / V_j blocks of 64 keys ([64 x 64]-column boxes, 128B swizzle)
//                 through a 3-stage ring;
//   warp 1      : single-thread tcgen05.mma issuer.  S_t = Q_t . K_j^T (M 128, N 64, K hd) into a double-buffered TMEM
//                 tile per query tile; O_t += P_t . V_j (M 128, N hd, K 64) with P_t read from shared memory (K-major) and
//                 V_j read in place as an MN-major operand (no transpose pass).  QK of block j+1 is issued BEFORE PV of
//                 block j, so the tensor pipe works on one tile while the softmax warps work on the other;
//   warps 4-7   : softmax of query tile 0 -- thread = query row: tcgen05.ld the 64 scores, running max / sum in the log2
//   warps 8-11  : softmax of query tile 1    domain, exp2, bf16 P written to shared memory in the UMMA swizzled layout.
// The accumulator O_t stays in TMEM for the whole key loop.  The reference maximum of a row is only moved when the running
// maximum exceeds it by more than 2^8 (the probabilities then stay below 256, exact in the same arithmetic), so the
// tcgen05.ld / scale / tcgen05.st correction of O is rare after the first blocks.
// TMEM (512 columns): S_0[2] | S_1[2] (4 x 64) | O_0 | O_1 (2 x 128).
#include "tc_common.cuh"
#include <math.h>

namespace ocrb {

constexpr int FT_BM = 128;           // query rows per tile
constexpr int FT_TILES = 2;          // query tiles per CTA
constexpr int FT_BN = 64;            // keys per block
constexpr int FT_STAGES = 3;         // K/V ring depth
constexpr int FT_THREADS = 384;      // warps 0-3: TMA, MMA, 2 idle; warps 4-7 / 8-11: softmax of tile 0 / 1
constexpr uint32_t FT_ATOM_Q = FT_BM * 128;     // bytes of one 64-column atom of a Q tile (128 rows x 128 B)
constexpr uint32_t FT_ATOM_KV = FT_BN * 128;    // bytes of one 64-column atom of a K or V block (64 rows x 128 B)
constexpr uint32_t FT_Q_BYTES = 2 * FT_ATOM_Q;  // 32 KiB per query tile (two atoms: columns 0-63, 64-127)
constexpr uint32_t FT_KV_BYTES = 2 * FT_ATOM_KV;   // 16 KiB per K (or V) block
constexpr uint32_t FT_P_BYTES = FT_BM * 128;    // 16 KiB: [128 rows][64 keys] bf16
constexpr uint32_t FT_STAGE_BYTES = 2 * FT_KV_BYTES;
constexpr float FT_RESCALE_THRESHOLD = 8.0f;    // log2 units

struct FlashParams {
  bf16 *O;
  long long ldo;
  const int32_t *cu_seqlens;
  int n_q, n_kv;
  float scale_log2;                  // softmax scale * log2(e)
  int causal;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major, 128B-swizzled operand (V block read in place as the N x K operand of P.V: N = head dim contiguous in
// memory, K = keys along the 128-byte rows).  Canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements:
// a 128-byte row holds 64 consecutive N elements of ONE key, 8 keys form a 1024-byte swizzle atom (SBO between
// 8-key groups), the next 64 N elements live LBO bytes further (the block's second 64-column atom).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int HD>
__global__ void __launch_bounds__(FT_THREADS, 1)
flash_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, FlashParams p) {
  constexpr int KSTEPS = HD / 16;                  // k-steps of Q.K^T (5 for hd 80: the second atom is used up to column 79)
  extern __shared__ uint8_t ft_smem_raw[];
  uint8_t *smem = ft_smem_raw + ((1024u - (smem_u32(ft_smem_raw) & 1023u)) & 1023u);   // offset into the __shared__ array: keeps the address space
  uint8_t *sQ = smem;                                            // [2 tiles][2 atoms][128 rows][128 B]
  uint8_t *sKV = sQ + FT_TILES * FT_Q_BYTES;                     // [stages][K | V][2 atoms][64 rows][128 B]
  uint8_t *sP = sKV + FT_STAGES * FT_STAGE_BYTES;                // [2 tiles][128 rows][128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + FT_TILES * FT_P_BYTES);
  uint64_t *q_full = bars;                    // [1]
  uint64_t *kv_full = q_full + 1;             // [STAGES]
  uint64_t *kv_empty = kv_full + FT_STAGES;   // [STAGES]
  uint64_t *s_full = kv_empty + FT_STAGES;    // [tile][buf]
  uint64_t *s_empty = s_full + 4;             // [tile][buf]
  uint64_t *p_full = s_empty + 4;             // [tile]
  uint64_t *o_done = p_full + 2;              // [tile]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = (int)gridDim.x - 1 - (int)blockIdx.x;         // longest (causal) work first
  const int head = blockIdx.y, seq = blockIdx.z;
  const int seq_start = p.cu_seqlens[seq];
  const int seq_len = p.cu_seqlens[seq + 1] - seq_start;
  const int q0 = pair * (FT_TILES * FT_BM);
  if (q0 >= seq_len) return;                                     // whole CTA beyond the sequence (uniform exit)
  const int kv_head = head / (p.n_q / p.n_kv);
  // key blocks each query tile needs
  int nblk_t[FT_TILES];
#pragma unroll
  for (int t = 0; t < FT_TILES; ++t) {
    const int qs = q0 + t * FT_BM;
    int hi = seq_len;
    if (p.causal) hi = min(seq_len, qs + FT_BM);
    nblk_t[t] = (qs < seq_len) ? (hi + FT_BN - 1) / FT_BN : 0;
  }
  const int nblk = max(nblk_t[0], nblk_t[1]);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < FT_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ───────────── TMA producer ─────────────
    if (lane == 0) {
      const int n_qt = (nblk_t[1] > 0) ? 2 : 1;
      mbar_expect_tx(q_full, n_qt * FT_Q_BYTES);
      for (int t = 0; t < n_qt; ++t)
        for (int a = 0; a < 2; ++a)
          tma_load_2d(sQ + t * FT_Q_BYTES + a * FT_ATOM_Q, &map_q, q_full, head * HD + a * 64, seq_start + q0 + t * FT_BM);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&kv_full[s], FT_STAGE_BYTES);
        uint8_t *sk = sKV + s * FT_STAGE_BYTES, *sv = sk + FT_KV_BYTES;
        for (int a = 0; a < 2; ++a) {
          tma_load_2d(sk + a * FT_ATOM_KV, &map_k, &kv_full[s], kv_head * HD + a * 64, seq_start + j * FT_BN);
          tma_load_2d(sv + a * FT_ATOM_KV, &map_v, &kv_full[s], kv_head * HD + a * 64, seq_start + j * FT_BN);
        }
        if (++s == FT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ───────────── MMA issuer ─────────────
    if (lane == 0) {
      // D = f32, A = B = bf16; S: both operands K-major, N = 64; O: B (= V) MN-major, N = HD
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FT_BN >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
      mbar_wait(q_full, 0);
      tcgen05_fence_after();
      auto issue_qk = [&](int j) {
        const int s = j % FT_STAGES;
        const uint32_t sk = smem_u32(sKV + s * FT_STAGE_BYTES);
        const int buf = j & 1;
        const int use = j >> 1;                       // n-th use of this S buffer
#pragma unroll
        for (int t = 0; t < FT_TILES; ++t) {
          if (j >= nblk_t[t]) continue;
          mbar_wait(&s_empty[t * 2 + buf], (use & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t sq = smem_u32(sQ + t * FT_Q_BYTES);
          const uint32_t tS = tmem_base + (uint32_t)(t * 2 + buf) * FT_BN;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint64_t adesc = make_smem_desc(sq + (k >> 2) * FT_ATOM_Q + (k & 3) * 32);
            const uint64_t bdesc = make_smem_desc(sk + (k >> 2) * FT_ATOM_KV + (k & 3) * 32);
            umma_bf16(tS, adesc, bdesc, idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[t * 2 + buf]);
        }
      };
      mbar_wait(&kv_full[0], 0);
      tcgen05_fence_after();
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % FT_STAGES;
        if (j + 1 < nblk) {
          mbar_wait(&kv_full[(j + 1) % FT_STAGES], ((j + 1) / FT_STAGES) & 1);
          tcgen05_fence_after();
          issue_qk(j + 1);
        }
        const uint32_t sv = smem_u32(sKV + s * FT_STAGE_BYTES + FT_KV_BYTES);
#pragma unroll
        for (int t = 0; t < FT_TILES; ++t) {
          if (j >= nblk_t[t]) continue;
          mbar_wait(&p_full[t], j & 1);
          tcgen05_fence_after();
          const uint32_t sp = smem_u32(sP + t * FT_P_BYTES);
          const uint32_t tO = tmem_base + 4 * FT_BN + (uint32_t)t * 128;
#pragma unroll
          for (int k = 0; k < FT_BN / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sp + k * 32);                       // 16 keys = 32 bytes along K
            const uint64_t bdesc = make_smem_desc_mn(sv + k * 2048, FT_ATOM_KV);      // 16 keys = two 8-key groups
            umma_bf16(tO, adesc, bdesc, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&o_done[t]);
        }
        umma_commit(&kv_empty[s]);                    // K_j (read by the S MMAs above) and V_j are free
      }
    }
  } else if (warp >= 4) {
    // ───────────── softmax warps: thread = one query row of tile t ─────────────
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                    // row inside the tile = TMEM lane
    const int q_row = q0 + t * FT_BM + r;              // row inside the sequence
    const int nb = nblk_t[t];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tO = lane_addr + 4 * FT_BN + (uint32_t)t * 128;
    uint8_t *prow = sP + t * FT_P_BYTES + r * 128;
    float m_ref = 0.f, l_sum = 0.f;
    for (int j = 0; j < nb; ++j) {
      const int buf = j & 1;
      mbar_wait(&s_full[t * 2 + buf], (j >> 1) & 1);
      tcgen05_fence_after();
      uint32_t sr[64];
      {
        uint32_t a[32], b[32];
        tmem_ld32(lane_addr + (uint32_t)(t * 2 + buf) * FT_BN, a);
        tmem_ld32(lane_addr + (uint32_t)(t * 2 + buf) * FT_BN + 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { sr[i] = a[i]; sr[32 + i] = b[i]; }
      }
      tcgen05_fence_before();
      mbar_arrive_cta(&s_empty[t * 2 + buf]);          // the S buffer may be overwritten by block j + 2
      // scores in the log2 domain, masked
      const int key0 = j * FT_BN;
      const bool boundary = (key0 + FT_BN > seq_len) || (p.causal && key0 + FT_BN - 1 > q0 + t * FT_BM + quad * 32);
      float bm = -INFINITY;
      float sv[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float x = __uint_as_float(sr[i]) * p.scale_log2;
        if (boundary) {
          const int key = key0 + i;
          if (key >= seq_len || (p.causal && key > q_row)) x = -INFINITY;
        }
        sv[i] = x;
        bm = fmaxf(bm, x);
      }
      // reference maximum: moved only when the block maximum exceeds it by more than the threshold
      float factor = 1.0f;
      bool need = false;
      if (j == 0) {
        m_ref = bm;                                    // finite: key 0 is visible to every row
      } else if (bm > m_ref + FT_RESCALE_THRESHOLD) {
        factor = exp2f(m_ref - bm);
        m_ref = bm;
        need = true;
      }
      float psum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = exp2f(sv[2 * i] - m_ref), p1 = exp2f(sv[2 * i + 1] - m_ref);
        psum += p0 + p1;
        __nv_bfloat162 v2 = __floats2bfloat162_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t *>(&v2);
      }
      l_sum = l_sum * factor + psum;
      // P.V of block j - 1 must be complete before P is overwritten / O is corrected
      if (j > 0) {
        mbar_wait(&o_done[t], (j - 1) & 1);
        tcgen05_fence_after();
      }
      if (__any_sync(0xffffffffu, need)) {
        // correction of the accumulator row in TMEM (rare): O *= 2^(old reference - new reference)
#pragma unroll 1
        for (int c = 0; c < HD; c += 16) {
          uint32_t o[16];
          tmem_ld16(tO + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st16(tO + c, o);
        }
        tmem_st_wait();
      }
      // P row -> shared memory, K-major 128B swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4 *>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive_cta(&p_full[t]);
    }
    if (nb > 0) {
      mbar_wait(&o_done[t], (nb - 1) & 1);
      tcgen05_fence_after();
      const float inv = 1.0f / l_sum;
      const bool row_ok = q_row < seq_len;
      bf16 *orow = p.O + (size_t)(seq_start + q_row) * p.ldo + (size_t)head * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += 16) {
        uint32_t o[16];
        tmem_ld16(tO + c, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int v8 = 0; v8 < 2; ++v8) {
            uint4 pkd;
            bf16 *pe = reinterpret_cast<bf16 *>(&pkd);
#pragma unroll
            for (int e = 0; e < 8; ++e) pe[e] = __float2bfloat16_rn(__uint_as_float(o[v8 * 8 + e]) * inv);
            *reinterpret_cast<uint4 *>(orow + c + v8 * 8) = pkd;
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int HD>
static int launch_flash_tc(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const FlashParams &p,
                           int n_seq, int max_seqlen, cudaStream_t st) {
  constexpr size_t smem = (size_t)FT_TILES * FT_Q_BYTES + (size_t)FT_STAGES * FT_STAGE_BYTES + (size_t)FT_TILES * FT_P_BYTES +
                          1024 /*align*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(flash_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const dim3 grid(cdiv(max_seqlen, FT_TILES * FT_BM), p.n_q, n_seq);
  flash_tc_kernel<HD><<<grid, FT_THREADS, smem, st>>>(mq, mk, mv, p);
  return check_launch("flash_tc_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_flash_attention_bf16(const void *q, int64_t ldq, const void *k, int64_t ldk, const void *v, int64_t ldv,
                                         void *out, int64_t ldo, const int32_t *cu_seqlens, int32_t n_seq,
                                         int32_t total_tokens, int32_t max_seqlen, int32_t n_q, int32_t n_kv, int32_t hd,
                                         float scale, int32_t causal, void *stream) {
  OCRB_REQUIRE(q && k && v && out && cu_seqlens, "flash_attention_bf16: null pointer");
  OCRB_REQUIRE(n_seq > 0 && total_tokens > 0 && max_seqlen > 0 && n_kv > 0 && n_q % n_kv == 0, "flash_attention_bf16: bad sizes");
  OCRB_REQUIRE(hd == 80 || hd == 128, "flash_attention_bf16: head_dim must be 80 or 128 (use ocrb_attention_varlen otherwise)");
  OCRB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "flash_attention_bf16: strides must be multiples of 8");
  OCRB_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "flash_attention_bf16: pointers must be 16-byte aligned");
  OCRB_REQUIRE(n_q <= 65535 && n_seq <= 65535, "flash_attention_bf16: grid too large");
  CUtensorMap mq, mk, mv;
  // 2-D views [total_tokens, heads * hd] of the q / k / v sections; boxes of 64 columns (128B swizzle).  A box that
  // reaches past the section's last column or past the last token is zero-filled by the TMA unit.
  int rc = make_tensor_map_bf16(&mq, q, total_tokens, (long long)n_q * hd, ldq, FT_BM);
  if (rc) return rc;
  rc = make_tensor_map_bf16(&mk, k, total_tokens, (long long)n_kv * hd, ldk, FT_BN);
  if (rc) return rc;
  rc = make_tensor_map_bf16(&mv, v, total_tokens, (long long)n_kv * hd, ldv, FT_BN);
  if (rc) return rc;
  FlashParams p;
  p.O = (bf16 *)out;
  p.ldo = ldo;
  p.cu_seqlens = cu_seqlens;
  p.n_q = n_q;
  p.n_kv = n_kv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  if (hd == 80) return launch_flash_tc<80>(mq, mk, mv, p, n_seq, max_seqlen, (cudaStream_t)stream);
  return launch_flash_tc<128>(mq, mk, mv, p, n_seq, max_seqlen, (cudaStream_t)stream);
}
