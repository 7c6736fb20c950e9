// TEST INFRASTRUCTURE: runs the kernels of handwritten-ocr_b200/csrc/textops_kernels.cuh on the CPU through tests/emu/cuda_emu.h
// with the dispatch of the product's C ABI (textops.cu: ocrb_levenshtein_batch, ocrb_lcs_align_batch).  Never shipped.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/textops_kernels.cuh"

using namespace ocrb;

template <int C, int T>
static void run_lev(const int32_t *sa, const int32_t *oa, const int32_t *sb, const int32_t *ob, int n_pairs, int32_t *out) {
  emu::launch(dim3(n_pairs), dim3(T), 0, [&] { levenshtein_kernel<C, T>(sa, oa, sb, ob, out); });
}

// force: 0 = the product's dispatch by max_len_b; otherwise the index (1..5) of a wider instantiation than needed
extern "C" int emu_levenshtein_batch(const int32_t *sa, const int32_t *oa, const int32_t *sb, const int32_t *ob, int n_pairs,
                                     int max_len_b, int32_t *out, int force) {
  if (n_pairs == 0) return 0;
  if (force == 0) {
    if (max_len_b <= 32) run_lev<1, 32>(sa, oa, sb, ob, n_pairs, out);
    else if (max_len_b <= 256) run_lev<4, 64>(sa, oa, sb, ob, n_pairs, out);
    else if (max_len_b <= 1024) run_lev<8, 128>(sa, oa, sb, ob, n_pairs, out);
    else if (max_len_b <= 4096) run_lev<16, 256>(sa, oa, sb, ob, n_pairs, out);
    else return -1;   // the wider instantiations (256 / 1024 threads x 32 columns) are too slow to emulate
    return 0;
  }
  if (force == 1 && max_len_b <= 256) { run_lev<4, 64>(sa, oa, sb, ob, n_pairs, out); return 0; }
  if (force == 2 && max_len_b <= 1024) { run_lev<8, 128>(sa, oa, sb, ob, n_pairs, out); return 0; }
  if (force == 3 && max_len_b <= 4096) { run_lev<16, 256>(sa, oa, sb, ob, n_pairs, out); return 0; }
  return -1;
}

extern "C" int emu_lcs_align_batch(const int32_t *sbb, const int32_t *obb, const int32_t *sw, const int32_t *ow, int n_pairs,
                                   int max_len_bb, int32_t *aligned, uint8_t *workspace, const int64_t *ws_off) {
  if (n_pairs == 0) return 0;
  const int diag_stride = (max_len_bb + 2 + 7) & ~7;
  const size_t smem = (size_t)3 * diag_stride * sizeof(uint16_t);
  emu::launch(dim3(n_pairs), dim3(LCS_THREADS), smem, [&] { lcs_align_kernel(sbb, obb, sw, ow, aligned, workspace, ws_off, diag_stride); });
  return 0;
}
