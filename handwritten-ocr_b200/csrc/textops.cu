// Candidate agreement / merge arithmetic: Levenshtein (tools.py:69-100) and LCS alignment
// (tools.py:465-493) on int32 symbol sequences.  Integer-exact.
#include "common.cuh"

namespace ocrb {

// ───────────────────────── Levenshtein: one warp per pair ─────────────────────────
// Rows are processed in bands of 32 (one DP row per lane).  At step t lane l computes cell
// (row base+l+1, col t-l): a skewed anti-diagonal wavefront.  `up` and the b symbol flow from
// lane l-1 by shuffle; the band's top boundary row lives in `rowbuf` (global, L1/L2 resident),
// read by lane 0 through 32-wide coalesced chunks and rewritten in place by lane 31 (the write
// column trails the read column by 31, so in-place is safe).
constexpr int LEV_WARPS = 4;

__global__ void __launch_bounds__(LEV_WARPS * 32)
levenshtein_kernel(const int32_t *__restrict__ seq_a, const int32_t *__restrict__ off_a,
                   const int32_t *__restrict__ seq_b, const int32_t *__restrict__ off_b, int n_pairs,
                   int ws_stride, int32_t *__restrict__ out, int32_t *__restrict__ workspace) {
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * LEV_WARPS + (threadIdx.x >> 5);
  if (pair >= n_pairs) return;
  const int32_t *a = seq_a + off_a[pair];
  const int32_t *b = seq_b + off_b[pair];
  const int n = off_a[pair + 1] - off_a[pair];
  const int m = off_b[pair + 1] - off_b[pair];
  if (n == 0 || m == 0) {
    if (lane == 0) out[pair] = n + m;
    return;
  }
  volatile int32_t *rowbuf = workspace + (size_t)pair * ws_stride;  // D[base][0..m]
  for (int j = lane; j <= m; j += 32) rowbuf[j] = j;
  __syncwarp();
  const unsigned full = 0xffffffffu;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane + 1;  // 1-based DP row of this lane
    const bool row_ok = i <= n;
    const int32_t ai = row_ok ? a[i - 1] : -1;
    int left = i;       // D[i][0]
    int diag = i - 1;   // D[i-1][0]
    int val = 0;        // value computed at the previous step (passed down as `up`)
    int32_t bch = 0;    // b symbol used at the previous step (passed down)
    const int steps = m + 31;
    for (int t0 = 1; t0 <= steps; t0 += 32) {
      // coalesced chunk: columns t0 .. t0+31 of the boundary row and of b
      const int jc = t0 + lane;
      int chunk_up = 0;
      int32_t chunk_b = 0;
      if (jc <= m) {
        chunk_up = rowbuf[jc];
        chunk_b = b[jc - 1];
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int t = t0 + k;
        int up = __shfl_up_sync(full, val, 1);
        int32_t bc = __shfl_up_sync(full, bch, 1);
        const int up0 = __shfl_sync(full, chunk_up, k);
        const int32_t b0 = __shfl_sync(full, chunk_b, k);
        if (lane == 0) {
          up = up0;
          bc = b0;
        }
        const int j = t - lane;
        if (row_ok && j >= 1 && j <= m) {
          const int cost = (ai != bc) ? 1 : 0;
          int v = min(up + 1, left + 1);
          v = min(v, diag + cost);
          diag = up;
          left = v;
          val = v;
          if (lane == 31) rowbuf[j] = v;         // becomes D[base+32][j]
          if (i == n && j == m) out[pair] = v;
        }
        bch = bc;
      }
    }
    __syncwarp();
  }
}

// ───────────────────────── LCS align: one CTA per (backbone, version) pair ─────────────────────────
// Anti-diagonal fill with three rolling diagonals in shared memory; per-cell direction byte
// (0 = match/diag, 1 = up, 2 = left) to global; single-thread backtrack.
constexpr int LCS_THREADS = 256;

__global__ void __launch_bounds__(LCS_THREADS)
lcs_align_kernel(const int32_t *__restrict__ seq_bb, const int32_t *__restrict__ off_bb,
                 const int32_t *__restrict__ seq_w, const int32_t *__restrict__ off_w,
                 int32_t *__restrict__ aligned, uint8_t *__restrict__ workspace,
                 const int64_t *__restrict__ ws_off, int diag_stride) {
  extern __shared__ uint16_t lcs_smem[];
  const int pair = blockIdx.x;
  const int32_t *bb = seq_bb + off_bb[pair];
  const int32_t *w = seq_w + off_w[pair];
  const int n = off_bb[pair + 1] - off_bb[pair];
  const int m = off_w[pair + 1] - off_w[pair];
  int32_t *al = aligned + off_bb[pair];
  for (int i = threadIdx.x; i < n; i += LCS_THREADS) al[i] = -1;
  if (n == 0 || m == 0) return;
  uint8_t *dir = workspace + ws_off[pair];
  uint16_t *d0 = lcs_smem;                 // diagonal d   (being written)
  uint16_t *d1 = lcs_smem + diag_stride;   // diagonal d-1
  uint16_t *d2 = lcs_smem + 2 * diag_stride;  // diagonal d-2
  for (int d = 2; d <= n + m; ++d) {
    const int ilo = max(1, d - m), ihi = min(n, d - 1);
    for (int i = ilo + threadIdx.x; i <= ihi; i += LCS_THREADS) {
      const int j = d - i;
      const int up = (i == 1) ? 0 : d1[i - 1];                 // dp[i-1][j]
      const int lf = (j == 1) ? 0 : d1[i];                     // dp[i][j-1]
      const int dg = (i == 1 || j == 1) ? 0 : d2[i - 1];       // dp[i-1][j-1]
      int v;
      uint8_t dr;
      if (bb[i - 1] == w[j - 1]) {
        v = dg + 1;
        dr = 0;
      } else if (up >= lf) {
        v = up;
        dr = 1;
      } else {
        v = lf;
        dr = 2;
      }
      d0[i] = (uint16_t)v;
      dir[(size_t)(i - 1) * m + (j - 1)] = dr;
    }
    __syncthreads();
    uint16_t *tmp = d2;
    d2 = d1;
    d1 = d0;
    d0 = tmp;
  }
  __threadfence_block();
  if (threadIdx.x == 0) {
    int i = n, j = m;
    while (i > 0 && j > 0) {
      const uint8_t dr = __ldcg(dir + (size_t)(i - 1) * m + (j - 1));
      if (dr == 0) {
        al[i - 1] = j - 1;
        --i;
        --j;
      } else if (dr == 1) {
        --i;
      } else {
        --j;
      }
    }
  }
}

}  // namespace ocrb

extern "C" int ocrb_levenshtein_batch(const int32_t *seq_a, const int32_t *off_a, const int32_t *seq_b,
                                      const int32_t *off_b, int32_t n_pairs, int32_t max_len_b,
                                      int32_t *out, int32_t *workspace, void *stream) {
  using namespace ocrb;
  if (n_pairs == 0) return OCRB_OK;
  OCRB_REQUIRE(n_pairs > 0 && max_len_b >= 0, "levenshtein_batch: bad sizes");
  OCRB_REQUIRE(seq_a && off_a && seq_b && off_b && out && workspace, "levenshtein_batch: null pointer");
  levenshtein_kernel<<<cdiv(n_pairs, LEV_WARPS), LEV_WARPS * 32, 0, (cudaStream_t)stream>>>(
      seq_a, off_a, seq_b, off_b, n_pairs, max_len_b + 1, out, workspace);
  return check_launch("levenshtein_kernel");
}

extern "C" int ocrb_lcs_align_batch(const int32_t *seq_bb, const int32_t *off_bb, const int32_t *seq_w,
                                    const int32_t *off_w, int32_t n_pairs, int32_t max_len_bb,
                                    int32_t *aligned, uint8_t *workspace, const int64_t *ws_off,
                                    void *stream) {
  using namespace ocrb;
  if (n_pairs == 0) return OCRB_OK;
  OCRB_REQUIRE(n_pairs > 0 && max_len_bb >= 0 && max_len_bb <= 16000, "lcs_align_batch: bad sizes");
  OCRB_REQUIRE(seq_bb && off_bb && seq_w && off_w && aligned && workspace && ws_off,
               "lcs_align_batch: null pointer");
  const int diag_stride = (max_len_bb + 2 + 7) & ~7;
  const size_t smem = (size_t)3 * diag_stride * sizeof(uint16_t);
  if (smem > 48 * 1024)
    OCRB_CUDA(cudaFuncSetAttribute(lcs_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lcs_align_kernel<<<n_pairs, LCS_THREADS, smem, (cudaStream_t)stream>>>(
      seq_bb, off_bb, seq_w, off_w, aligned, workspace, ws_off, diag_stride);
  return check_launch("lcs_align_kernel");
}
