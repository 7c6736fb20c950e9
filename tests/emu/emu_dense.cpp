// TEST INFRASTRUCTURE: runs the kernels of handwritten-ocr_b200/csrc/dense_kernels.cuh on the CPU through tests/emu/cuda_emu.h with
// the launch geometry of the product's C ABI (dense.cu).  bf16 tensors travel as uint16 arrays.  Never shipped.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/dense_kernels.cuh"

using namespace ocrb;

static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// variant 0: one CTA per row (rmsnorm_kernel), 1: one warp per row (rmsnorm_warp_kernel, dim <= 4096)
extern "C" int emu_rmsnorm(const uint16_t *x, long long ldx, const uint16_t *w, uint16_t *y, long long ldy, int rows, int dim,
                           float eps, int variant) {
  const bf16 *X = (const bf16 *)x, *W = (const bf16 *)w;
  bf16 *Y = (bf16 *)y;
  if (variant == 1) emu::launch(dim3(cdivu(rows, 8)), dim3(256), 0, [&] { rmsnorm_warp_kernel(X, ldx, W, Y, ldy, rows, dim, eps); });
  else emu::launch(dim3(rows), dim3(256), 0, [&] { rmsnorm_kernel(X, ldx, W, Y, ldy, dim, eps); });
  return 0;
}

extern "C" int emu_rope_vision(uint16_t *qkv, int S, int heads, int hd, const float *cosT, const float *sinT, int variant) {
  bf16 *Q = (bf16 *)qkv;
  if (variant == 1) {
    const int total_v = S * 2 * heads * (hd / 16);
    emu::launch(dim3(cdivu(total_v, 256)), dim3(256), 0, [&] { rope_vision_vec_kernel(Q, S, heads, hd, cosT, sinT); });
  } else {
    const long long total = (long long)S * 2 * heads * (hd / 2);
    emu::launch(dim3(cdivu(total, 256)), dim3(256), 0, [&] { rope_vision_kernel(Q, S, heads, hd, cosT, sinT); });
  }
  return 0;
}

extern "C" int emu_rope_text(uint16_t *q, long long ldq, uint16_t *k, long long ldk, int T, int n_q, int n_kv, int hd,
                             const uint16_t *cosT, const uint16_t *sinT, int variant) {
  bf16 *Q = (bf16 *)q, *K = (bf16 *)k;
  const bf16 *C = (const bf16 *)cosT, *S = (const bf16 *)sinT;
  if (variant == 1) {
    const int total_v = T * (n_q + n_kv) * (hd / 16);
    emu::launch(dim3(cdivu(total_v, 256)), dim3(256), 0, [&] { rope_text_vec_kernel(Q, ldq, K, ldk, T, n_q, n_kv, hd, C, S); });
  } else {
    const long long total = (long long)T * (n_q + n_kv) * (hd / 2);
    emu::launch(dim3(cdivu(total, 256)), dim3(256), 0, [&] { rope_text_kernel(Q, ldq, K, ldk, T, n_q, n_kv, hd, C, S); });
  }
  return 0;
}

extern "C" int emu_rows_copy(const uint16_t *src, long long lds, const int32_t *src_idx, uint16_t *dst, long long ldd,
                             const int32_t *dst_idx, int n_rows, int dim) {
  const int nvec = dim / 8;
  emu::launch(dim3(cdivu((long long)n_rows * nvec, 256)), dim3(256), 0,
              [&] { rows_copy_kernel((const bf16 *)src, lds, src_idx, (bf16 *)dst, ldd, dst_idx, n_rows, nvec); });
  return 0;
}

extern "C" int emu_kv_write_prefill(const uint16_t *k, long long ldk, const uint16_t *v, long long ldv, uint16_t *k_cache,
                                    uint16_t *v_cache, const int32_t *block_table, int max_pages, const int32_t *cu_seqlens,
                                    int n_seq, int T, int page_size, int n_kv, int hd) {
  const int row_vec = n_kv * hd / 8;
  emu::launch(dim3(cdivu((long long)T * row_vec, 256)), dim3(256), 0, [&] {
    kv_write_prefill_kernel((const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)k_cache, (bf16 *)v_cache, block_table, max_pages,
                            cu_seqlens, n_seq, T, page_size, row_vec, hd / 8);
  });
  return 0;
}

extern "C" int emu_argmax_step(const uint16_t *logits, long long ldl, int B, int V, int eos, int pad, int max_new,
                               int32_t *out_tokens, int32_t *next_ids, int32_t *finished, int32_t *ctx_len, int32_t *step,
                               int advance_ctx) {
  emu::launch(dim3(B), dim3(512), 0, [&] {
    argmax_step_kernel((const bf16 *)logits, ldl, V, eos, pad, max_new, out_tokens, next_ids, finished, ctx_len, step, advance_ctx);
  });
  emu::launch(dim3(1), dim3(1), 0, [&] { step_increment_kernel(step); });
  return 0;
}

extern "C" int emu_residual_add(uint16_t *x, long long ldx, const uint16_t *y, long long ldy, int rows, int dim) {
  const int vpr = dim / 8;
  emu::launch(dim3(cdivu((long long)rows * vpr, 256)), dim3(256), 0,
              [&] { residual_add_kernel((bf16 *)x, ldx, (const bf16 *)y, ldy, rows, vpr); });
  return 0;
}
