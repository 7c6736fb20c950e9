// Kernels of the candidate agreement / merge arithmetic: Levenshtein (tools.py:69-100) and LCS alignment
// (tools.py:465-493) on int32 symbol sequences.  Integer-exact.  Kernel definitions only -- no launches -- so that
// tests/emu can compile this file for the host (OCRB_EMU) and run the kernels thread by thread against the oracle.
#pragma once
#ifndef OCRB_EMU
#include <cuda_runtime.h>
#ifndef OCRB_DYN_SMEM
#define OCRB_DYN_SMEM(T, name) extern __shared__ T name[]
#endif
#endif
#include <stdint.h>

namespace ocrb {

__device__ __forceinline__ int tmin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int tmax(int a, int b) { return a > b ? a : b; }

// ───────────────────────── Levenshtein: skewed anti-diagonal wavefront ─────────────────────────
// One CTA per pair.  Thread t owns the column strip [t*C, (t+1)*C) of the DP matrix (previous-row values and the b
// symbols of the strip live in registers); at step s it computes row i = s - t + 1 of its strip, so the active cells of
// a step form an anti-diagonal band of T x C cells.  The only inter-thread traffic is the strip's right boundary
// value, handed to thread t+1 through a double-buffered shared array, one __syncthreads per step.  Small pairs keep
// the old formulation's granularity (one warp, C = 1..) simply by instantiating fewer threads.
// Integer-exact: D[i][j] = min(D[i-1][j] + 1, D[i][j-1] + 1, D[i-1][j-1] + (a_i != b_j))   (tools.py:69-83).
template <int C, int T>
__global__ void __launch_bounds__(T)
levenshtein_kernel(const int32_t *__restrict__ seq_a, const int32_t *__restrict__ off_a,
                   const int32_t *__restrict__ seq_b, const int32_t *__restrict__ off_b, int32_t *__restrict__ out) {
  __shared__ int bnd[2][T];
  const int pair = blockIdx.x;
  const int t = threadIdx.x;
  const int32_t *a = seq_a + off_a[pair];
  const int32_t *b = seq_b + off_b[pair];
  const int n = off_a[pair + 1] - off_a[pair];
  const int m = off_b[pair + 1] - off_b[pair];
  if (n == 0 || m == 0) {
    if (t == 0) out[pair] = n + m;
    return;
  }
  if (m > C * T) return;                       // handled by a wider instantiation (host dispatch guarantees this)
  const int j0 = t * C;                        // 0-based first column of the strip (DP column j0 + 1)
  int prev[C];
  int32_t bs[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = j0 + c;
    prev[c] = j + 1;                           // D[0][j+1]
    bs[c] = (j < m) ? b[j] : (int32_t)0x7fffffff;
  }
  int diag_in = j0;                            // D[i-1][j0] for the row about to be computed; row 0: D[0][j0] = j0
  const int steps = n + T - 1;
  const int own_t = (m - 1) / C, own_c = (m - 1) % C;   // strip / register holding column m
  for (int s = 0; s < steps; ++s) {
    const int i = s - t + 1;                   // 1-based row of this thread at this step
    const bool active = (i >= 1) && (i <= n) && (j0 < m);
    if (active) {
      const int32_t ai = a[i - 1];
      int left = (t == 0) ? i : bnd[(s + 1) & 1][t - 1];       // D[i][j0], written by thread t-1 in step s-1
      int diag = (t == 0) ? i - 1 : diag_in;                   // D[i-1][j0]
      diag_in = left;                                          // becomes the diagonal of the next row
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int up = prev[c];
        const int v = tmin(tmin(up, left) + 1, diag + (ai != bs[c] ? 1 : 0));
        diag = up;
        left = v;
        prev[c] = v;
      }
      bnd[s & 1][t] = left;                                    // D[i][j0 + C]
      if (i == n && t == own_t) {
        int r = 0;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (c == own_c) r = prev[c];
        out[pair] = r;
      }
    }
    __syncthreads();
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
  OCRB_DYN_SMEM(uint16_t, lcs_smem);
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
    const int ilo = tmax(1, d - m), ihi = tmin(n, d - 1);
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
