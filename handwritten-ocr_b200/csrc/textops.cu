// Candidate agreement / merge arithmetic: Levenshtein (tools.py:69-100) and LCS alignment
// (tools.py:465-493) on int32 symbol sequences.  Integer-exact.
#include "common.cuh"
#include "textops_kernels.cuh"

namespace ocrb {

template <int C, int T>
static int launch_lev(const int32_t *seq_a, const int32_t *off_a, const int32_t *seq_b, const int32_t *off_b, int n_pairs,
                      int32_t *out, cudaStream_t st) {
  levenshtein_kernel<C, T><<<n_pairs, T, 0, st>>>(seq_a, off_a, seq_b, off_b, out);
  return check_launch("levenshtein_kernel");
}

}  // namespace ocrb

extern "C" int ocrb_levenshtein_batch(const int32_t *seq_a, const int32_t *off_a, const int32_t *seq_b,
                                      const int32_t *off_b, int32_t n_pairs, int32_t max_len_b,
                                      int32_t *out, int32_t *workspace, void *stream) {
  using namespace ocrb;
  if (n_pairs == 0) return OCRB_OK;
  OCRB_REQUIRE(n_pairs > 0 && max_len_b >= 0, "levenshtein_batch: bad sizes");
  OCRB_REQUIRE(seq_a && off_a && seq_b && off_b && out, "levenshtein_batch: null pointer");
  OCRB_REQUIRE(max_len_b <= 32 * 1024, "levenshtein_batch: sequences longer than 32768 symbols are not supported");
  (void)workspace;   // kept in the ABI for callers that size it; the wavefront needs no global scratch
  cudaStream_t st = (cudaStream_t)stream;
  // strip width C x threads T >= max_len_b: narrow strips / few threads for short texts (shorter pipeline fill)
  if (max_len_b <= 32) return launch_lev<1, 32>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  if (max_len_b <= 256) return launch_lev<4, 64>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  if (max_len_b <= 1024) return launch_lev<8, 128>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  if (max_len_b <= 4096) return launch_lev<16, 256>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  if (max_len_b <= 8192) return launch_lev<32, 256>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  if (max_len_b <= 16384) return launch_lev<16, 1024>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
  return launch_lev<32, 1024>(seq_a, off_a, seq_b, off_b, n_pairs, out, st);
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
