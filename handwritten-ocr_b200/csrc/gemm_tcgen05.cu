// Dense GEMM of the VLM on 5th-generation tensor cores: D[M,N'] = epilogue(A[M,K] . W[N,K]^T).
// Every nn.Linear / Conv3d-as-GEMM of the vision tower and of decoder prefill goes through here
// (HF modeling_qwen2_5_vl.py: patch_embed :99-111, qkv/proj :214-283, MLPs :76-88 / :607-622,
// merger :133-146, q/k/v/o :704-760).
//
// Structure (one 128 x BN output tile per CTA, 64 + 32 * GM_EPI_WARPS threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 128x64 (A) and BNx64 (W) bf16 boxes, 128B swizzle,
//              into a 4-stage shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32 in TMEM),
//              tcgen05.commit frees ring slots / signals the accumulator
//   warps 2.. : epilogue -- tcgen05.ld 32 lanes x 32 columns -> registers -> bias / residual / SwiGLU / GELU with
//              HF's bf16 rounding points -> 16-byte global stores.  Two warps per TMEM lane quadrant (warp % 4), each
//              half of the tile's columns: one warp per scheduler is issue-latency bound (~400 dependent instructions per
//              32-column chunk) and took 12-17 us per 128 x 256 tile -- longer than the 7.5 us of MMAs of a K = 1280
//              tile, so the vision tower's GEMMs waited for their own epilogue (profiles/r02_summary.md)
// M/N/K tails are handled by TMA zero fill on loads and predicated stores.
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <mutex>

namespace ocrb {

constexpr int GM_BM = 128;
constexpr int GM_BK = 64;          // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int GM_STAGES = 4;
#ifndef GM_EPI_WARPS_N
#define GM_EPI_WARPS_N 8
#endif
constexpr int GM_EPI_WARPS = GM_EPI_WARPS_N;   // epilogue warps: GM_EPI_WARPS / 4 per TMEM lane quadrant, each a share of the columns
constexpr int GM_NPART = GM_EPI_WARPS / 4;
constexpr int GM_THREADS = 64 + 32 * GM_EPI_WARPS;

__device__ __forceinline__ float silu_bf16r(float g) { return bf16_round(g / (1.0f + expf(-g))); }
__device__ __forceinline__ float gelu_bf16r(float x) { return bf16_round(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f))); }

struct GemmParams {
  bf16 *D;
  long long ldd;
  const bf16 *bias;
  const bf16 *residual;
  long long ldr;
  int M, N, K;
  int epilogue;
};

// Tile order: the 148 concurrently processed tiles should share the LARGER operand panel through L2 and stream it from HBM
// once.  N > M (prefill: weights 271 MB vs activations 22 MB): m fastest, so neighbours reuse the same weight tile;
// otherwise (vision tower: M = 12k rows) n fastest, neighbours reuse the same activation rows.
template <int BN>
__device__ __forceinline__ void tile_origin(int tile, int num_n, int num_m, int &n0, int &m0) {
  if (num_n * BN > num_m * GM_BM) {
    n0 = (tile / num_m) * BN;
    m0 = (tile % num_m) * GM_BM;
  } else {
    n0 = (tile % num_n) * BN;
    m0 = (tile / num_n) * GM_BM;
  }
}

// Persistent kernel: gridDim.x CTAs (one per SM) walk the 128 x BN output tiles (tile = m_blk * num_n + n_blk, so the
// CTAs of a wave share A panels through L2).  Two TMEM accumulators of BN columns: the MMA warp fills one while the
// epilogue warps drain the other, so bias / SwiGLU / GELU / residual arithmetic and the global stores of tile i overlap
// the tensor-core work of tile i+1; the TMA ring runs ahead across tile boundaries.
template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmParams p) {
  constexpr uint32_t A_BYTES = GM_BM * GM_BK * 2;   // 16 KiB
  constexpr uint32_t B_BYTES = BN * GM_BK * 2;      // 16 / 32 KiB
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;                 // 256 / 512 columns: two accumulators
  extern __shared__ uint8_t gm_smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t *smem = gm_smem_raw + ((1024u - (smem_u32(gm_smem_raw) & 1023u)) & 1023u);   // offset into the __shared__ array: keeps the address space
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + GM_STAGES * STAGE_BYTES);
  uint64_t *empty_bar = full_bar + GM_STAGES;
  uint64_t *tmem_full = empty_bar + GM_STAGES;      // [2]
  uint64_t *tmem_empty = tmem_full + 2;             // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + GM_BK - 1) / GM_BK;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_m = (p.M + GM_BM - 1) / GM_BM;
  const int num_tiles = num_n * num_m;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < GM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 32 * GM_EPI_WARPS);
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

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n0, m0;
        tile_origin<BN>(tile, num_n, num_m, n0, m0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t *sa = smem + s * STAGE_BYTES;
          tma_load_2d(sa, &map_a, &full_bar[s], kb * GM_BK, m0);
          tma_load_2d(sa + A_BYTES, &map_w, &full_bar[s], kb * GM_BK, n0);
          if (++s == GM_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < GM_BK / UMMA_K; ++k) {
            // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the ring slot once these MMAs have read it
          if (++s == GM_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ───────────── epilogue: warp w owns TMEM lane quadrant w % 4 and column share (w - 2) / 4 ─────────────
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int n0, m0;
      tile_origin<BN>(tile, num_n, num_m, n0, m0);
      const int acc = it & 1;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t lane_addr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      const bool row_ok = row < p.M;
      if (p.epilogue == OCRB_EPI_SWIGLU) {
        // weight rows are packed per 128 as [gate 64 | up 64]; output column block = n0/2
        constexpr int ITEMS = BN / 64;              // (128-column block, 32-column chunk of its gate / up halves)
        constexpr int PER = ITEMS / GM_NPART;
        static_assert(ITEMS % GM_NPART == 0, "epilogue warps must divide the SwiGLU chunks");
#pragma unroll 1
        for (int item = part * PER; item < (part + 1) * PER; ++item) {
          {
            const int blk = item >> 1, c = (item & 1) * 32;
            uint32_t g[32], u[32];
            tmem_ld32(lane_addr + blk * 128 + c, g);
            tmem_ld32(lane_addr + blk * 128 + 64 + c, u);
            tmem_ld_wait();
            if (item == (part + 1) * PER - 1) {     // this warp's last TMEM read of the tile: hand the accumulator back
              tcgen05_fence_before();
              mbar_arrive_cta(&tmem_empty[acc]);
            }
            const int ng = n0 + blk * 128 + c;        // packed row index of the gate columns
            const int out_col = (n0 >> 1) + blk * 64 + c;
            if (row_ok && ng < p.N) {
              bf16 *drow = p.D + (size_t)row * p.ldd + out_col;
              uint4 bg4[4], bu4[4];
#pragma unroll
              for (int v8 = 0; v8 < 4; ++v8) {
                bg4[v8] = make_uint4(0, 0, 0, 0);
                bu4[v8] = make_uint4(0, 0, 0, 0);
                if (p.bias) {
                  bg4[v8] = __ldg(reinterpret_cast<const uint4 *>(p.bias + ng + v8 * 8));
                  bu4[v8] = __ldg(reinterpret_cast<const uint4 *>(p.bias + ng + 64 + v8 * 8));
                }
              }
#pragma unroll
              for (int v8 = 0; v8 < 4; ++v8) {
                uint4 pk;
                bf16 *pe = reinterpret_cast<bf16 *>(&pk);
                const bf16 *bge = reinterpret_cast<const bf16 *>(&bg4[v8]);
                const bf16 *bue = reinterpret_cast<const bf16 *>(&bu4[v8]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int i = v8 * 8 + e;
                  float gv = __uint_as_float(g[i]), uv = __uint_as_float(u[i]);
                  if (p.bias) {
                    gv += __bfloat162float(bge[e]);
                    uv += __bfloat162float(bue[e]);
                  }
                  pe[e] = __float2bfloat16_rn(silu_bf16r(bf16_round(gv)) * bf16_round(uv));
                }
                *reinterpret_cast<uint4 *>(drow + v8 * 8) = pk;
              }
            }
          }
        }
      } else {
        constexpr int CPW = BN / GM_NPART;          // columns per epilogue warp
        static_assert(CPW % 32 == 0, "epilogue warps must divide the 32-column chunks");
#pragma unroll 1
        for (int c = part * CPW; c < (part + 1) * CPW; c += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + c, r);
          tmem_ld_wait();
          if (c == (part + 1) * CPW - 32) {          // this warp's last TMEM read of the tile
            tcgen05_fence_before();
            mbar_arrive_cta(&tmem_empty[acc]);
          }
          const int n = n0 + c;
          if (row_ok && n < p.N) {
            bf16 *drow = p.D + (size_t)row * p.ldd + n;
            const bf16 *rrow = p.residual ? p.residual + (size_t)row * p.ldr + n : nullptr;
            // all loads of the chunk first: the residual is usually updated in place (D == residual), so a load inside
            // the store loop may not be moved above the preceding store and every 16-byte piece would cost its own
            // L2 round trip (4 per chunk, 32 per tile: longer than the tile's MMAs at K = 1280)
            uint4 rk4[4], bk4[4];
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
              rk4[v8] = make_uint4(0, 0, 0, 0);
              bk4[v8] = make_uint4(0, 0, 0, 0);
              if (n + v8 * 8 < p.N) {
                if (p.epilogue == OCRB_EPI_RESIDUAL) rk4[v8] = *reinterpret_cast<const uint4 *>(rrow + v8 * 8);
                if (p.bias) bk4[v8] = __ldg(reinterpret_cast<const uint4 *>(p.bias + n + v8 * 8));
              }
            }
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
              if (n + v8 * 8 >= p.N) break;
              uint4 pk;
              const uint4 rk = rk4[v8];
              bf16 *pe = reinterpret_cast<bf16 *>(&pk);
              const bf16 *re = reinterpret_cast<const bf16 *>(&rk);
              const bf16 *be = reinterpret_cast<const bf16 *>(&bk4[v8]);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int i = v8 * 8 + e;
                float v = __uint_as_float(r[i]);
                if (p.bias) v += __bfloat162float(be[e]);
                v = bf16_round(v);
                if (p.epilogue == OCRB_EPI_RESIDUAL) v += __bfloat162float(re[e]);
                else if (p.epilogue == OCRB_EPI_GELU) v = gelu_bf16r(v);
                pe[e] = __float2bfloat16_rn(v);
              }
              *reinterpret_cast<uint4 *>(drow + v8 * 8) = pk;
            }
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ───────────── host: tensor maps ─────────────
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tensor_map_bf16(CUtensorMap *m, const void *ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled not available from the driver");
    return OCRB_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)GM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return OCRB_ECUDA;
  }
  return OCRB_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mw, const GemmParams &p, cudaStream_t st) {
  constexpr size_t smem = (size_t)GM_STAGES * (GM_BM * GM_BK * 2 + BN * GM_BK * 2) + 1024 /*align*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    OCRB_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const int tiles = cdiv(p.N, BN) * cdiv(p.M, GM_BM);
  gemm_tcgen05_kernel<BN><<<tiles < n_sm ? tiles : n_sm, GM_THREADS, smem, st>>>(ma, mw, p);
  return check_launch("gemm_tcgen05_kernel");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_gemm_bf16(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int32_t M,
                              int32_t N, int32_t K, const void *bias, const void *residual, int64_t ldr, int32_t epilogue,
                              void *stream) {
  OCRB_REQUIRE(A && W && D, "gemm_bf16: null pointer");
  OCRB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad sizes");
  OCRB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0 && N % 8 == 0,
               "gemm_bf16: K, N and row strides must be multiples of 8 (16-byte TMA / store granularity)");
  OCRB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)D & 15) == 0,
               "gemm_bf16: pointers must be 16-byte aligned");
  OCRB_REQUIRE(epilogue >= 0 && epilogue <= 3, "gemm_bf16: bad epilogue");
  OCRB_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "gemm_bf16: bias must be 16-byte aligned");
  OCRB_REQUIRE(epilogue != OCRB_EPI_RESIDUAL || (residual && ldr % 8 == 0 && ((uintptr_t)residual & 15) == 0),
               "gemm_bf16: residual epilogue needs a 16-byte aligned residual with ldr % 8 == 0");
  OCRB_REQUIRE(epilogue != OCRB_EPI_SWIGLU || N % 128 == 0, "gemm_bf16: SwiGLU needs packed N % 128 == 0");
  OCRB_REQUIRE(epilogue != OCRB_EPI_GELU || bias, "gemm_bf16: GELU epilogue expects a bias");
  GemmParams p;
  p.D = (bf16 *)D;
  p.ldd = ldd;
  p.bias = (const bf16 *)bias;
  p.residual = (epilogue == OCRB_EPI_RESIDUAL) ? (const bf16 *)residual : nullptr;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epilogue = epilogue;
  // Tile width: 256 whenever it divides N and the grid still fills the machine.  (Choosing 128 to shave the last
  // partial wave was measured slower: 128-wide tiles re-read the activation panel twice as often.)
  // (128-wide tiles for the N = 1280 shapes -- 940 half tiles = 6.4 half waves instead of 470 tiles = 3.2 waves -- re-measured
  // with eight epilogue warps: proj 46.6 -> 54.1 us, down 95.8 -> 123.5 us; the activation panel is re-read twice as often.)
  const bool wide = (N % 256 == 0) && ((long long)cdiv(N, 256) * cdiv(M, GM_BM) >= 148);
  const int BN = wide ? 256 : 128;
  CUtensorMap ma, mw;
  int rc = make_tensor_map_bf16(&ma, A, M, K, lda, GM_BM);
  if (rc) return rc;
  rc = make_tensor_map_bf16(&mw, W, N, K, ldw, BN);
  if (rc) return rc;
  if (wide) return launch_gemm<256>(ma, mw, p, (cudaStream_t)stream);
  return launch_gemm<128>(ma, mw, p, (cudaStream_t)stream);
}
