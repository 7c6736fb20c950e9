// Small bandwidth-bound ops of the VLM (HF modeling_qwen2_5_vl.py): RMSNorm, vision RoPE, text mRoPE,
// embedding / row gathers, paged-KV prefill write, greedy argmax + step bookkeeping.  All mirror HF's
// rounding points (fp32 math, bf16 stores where HF materialises bf16 tensors).
#include "common.cuh"
#include <math.h>
#include "dense_kernels.cuh"

namespace ocrb {

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_rmsnorm_bf16(const void *x, int64_t ldx, const void *w, void *y, int64_t ldy, int32_t rows,
                                 int32_t dim, float eps, void *stream) {
  OCRB_REQUIRE(x && w && y && rows > 0 && dim > 0 && dim % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0,
               "rmsnorm_bf16: bad arguments (dim and strides must be multiples of 8)");
  if (dim <= 4096 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0) {
    rmsnorm_warp_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const bf16 *)x, ldx, (const bf16 *)w, (bf16 *)y, ldy,
                                                                         rows, dim, eps);
    return check_launch("rmsnorm_warp_kernel");
  }
  rmsnorm_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const bf16 *)x, ldx, (const bf16 *)w, (bf16 *)y, ldy, dim, eps);
  return check_launch("rmsnorm_kernel");
}

extern "C" int ocrb_rope_vision(void *qkv, int32_t S, int32_t heads, int32_t hd, const float *cosT, const float *sinT,
                                void *stream) {
  OCRB_REQUIRE(qkv && cosT && sinT && S > 0 && heads > 0 && hd > 0 && hd % 2 == 0, "rope_vision: bad arguments");
  if ((hd / 2) % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)cosT & 15) == 0 && ((uintptr_t)sinT & 15) == 0 &&
      (long long)S * 2 * heads * (hd / 16) < (1ll << 31)) {
    const int total_v = S * 2 * heads * (hd / 16);
    rope_vision_vec_kernel<<<cdiv(total_v, 256), 256, 0, (cudaStream_t)stream>>>((bf16 *)qkv, S, heads, hd, cosT, sinT);
    return check_launch("rope_vision_vec_kernel");
  }
  const long long total = (long long)S * 2 * heads * (hd / 2);
  rope_vision_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((bf16 *)qkv, S, heads, hd, cosT, sinT);
  return check_launch("rope_vision_kernel");
}

extern "C" int ocrb_rope_text(void *q, int64_t ldq, void *k, int64_t ldk, int32_t T, int32_t n_q, int32_t n_kv,
                              int32_t hd, const void *cosT, const void *sinT, void *stream) {
  OCRB_REQUIRE(q && k && cosT && sinT && T > 0 && n_q > 0 && n_kv > 0 && hd % 2 == 0, "rope_text: bad arguments");
  if ((hd / 2) % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 &&
      ((uintptr_t)cosT & 15) == 0 && ((uintptr_t)sinT & 15) == 0 && (long long)T * (n_q + n_kv) * (hd / 16) < (1ll << 31)) {
    const int total_v = T * (n_q + n_kv) * (hd / 16);
    rope_text_vec_kernel<<<cdiv(total_v, 256), 256, 0, (cudaStream_t)stream>>>((bf16 *)q, ldq, (bf16 *)k, ldk, T, n_q, n_kv, hd,
                                                                               (const bf16 *)cosT, (const bf16 *)sinT);
    return check_launch("rope_text_vec_kernel");
  }
  const long long total = (long long)T * (n_q + n_kv) * (hd / 2);
  rope_text_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((bf16 *)q, ldq, (bf16 *)k, ldk, T, n_q, n_kv, hd,
                                                                      (const bf16 *)cosT, (const bf16 *)sinT);
  return check_launch("rope_text_kernel");
}

extern "C" int ocrb_decode_rope_table(const int32_t *ctx_len, const int32_t *rope_delta, const float *inv_freq, int32_t B,
                                      int32_t hd, void *cosT, void *sinT, void *stream) {
  OCRB_REQUIRE(ctx_len && rope_delta && inv_freq && cosT && sinT && B > 0 && hd % 2 == 0, "decode_rope_table: bad arguments");
  decode_rope_table_kernel<<<cdiv(B * (hd / 2), 128), 128, 0, (cudaStream_t)stream>>>(ctx_len, rope_delta, inv_freq, B, hd,
                                                                                   (bf16 *)cosT, (bf16 *)sinT);
  return check_launch("decode_rope_table_kernel");
}

extern "C" int ocrb_residual_add_bf16(void *x, int64_t ldx, const void *y, int64_t ldy, int32_t rows, int32_t dim, void *stream) {
  if (rows == 0) return OCRB_OK;
  OCRB_REQUIRE(x && y && rows > 0 && dim > 0 && dim % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "residual_add_bf16: bad arguments");
  const long long total = (long long)rows * (dim / 8);
  residual_add_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((bf16 *)x, ldx, (const bf16 *)y, ldy, rows, dim / 8);
  return check_launch("residual_add_kernel");
}

extern "C" int ocrb_rows_copy(const void *src, int64_t lds, const int32_t *src_idx, void *dst, int64_t ldd,
                              const int32_t *dst_idx, int32_t n_rows, int32_t dim, void *stream) {
  if (n_rows == 0) return OCRB_OK;
  OCRB_REQUIRE(src && dst && n_rows > 0 && dim > 0 && dim % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0,
               "rows_copy: bad arguments");
  const long long total = (long long)n_rows * (dim / 8);
  rows_copy_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16 *)src, lds, src_idx, (bf16 *)dst, ldd,
                                                                      dst_idx, n_rows, dim / 8);
  return check_launch("rows_copy_kernel");
}

extern "C" int ocrb_embed_gather(const void *table, const int32_t *ids, void *out, int32_t T, int32_t dim, void *stream) {
  OCRB_REQUIRE(ids, "embed_gather: null ids");
  return ocrb_rows_copy(table, dim, ids, out, dim, nullptr, T, dim, stream);
}

extern "C" int ocrb_kv_write_prefill(const void *k, int64_t ldk, const void *v, int64_t ldv, void *k_cache, void *v_cache,
                                     const int32_t *block_table, int32_t max_pages, const int32_t *cu_seqlens,
                                     int32_t n_seq, int32_t T, int32_t page_size, int32_t n_kv, int32_t hd, void *stream) {
  OCRB_REQUIRE(k && v && k_cache && v_cache && block_table && cu_seqlens, "kv_write_prefill: null pointer");
  OCRB_REQUIRE(T > 0 && n_seq > 0 && page_size > 0 && hd % 8 == 0 && n_kv > 0 && ldk % 8 == 0 && ldv % 8 == 0,
               "kv_write_prefill: bad sizes");
  const int row_vec = n_kv * hd / 8;
  kv_write_prefill_kernel<<<cdiv((long long)T * row_vec, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16 *)k, ldk, (const bf16 *)v, ldv, (bf16 *)k_cache, (bf16 *)v_cache, block_table, max_pages, cu_seqlens,
      n_seq, T, page_size, row_vec, hd / 8);
  return check_launch("kv_write_prefill_kernel");
}

extern "C" int ocrb_argmax_step(const void *logits, int64_t ldl, int32_t B, int32_t V, int32_t eos, int32_t pad,
                                int32_t max_new, int32_t *out_tokens, int32_t *next_ids, int32_t *finished,
                                int32_t *ctx_len, int32_t *step, int32_t advance_ctx, void *stream) {
  OCRB_REQUIRE(logits && out_tokens && next_ids && finished && ctx_len && step, "argmax_step: null pointer");
  OCRB_REQUIRE(B > 0 && V > 0 && ldl % 8 == 0 && max_new > 0, "argmax_step: bad sizes");
  argmax_step_kernel<<<B, 512, 0, (cudaStream_t)stream>>>((const bf16 *)logits, ldl, V, eos, pad, max_new, out_tokens,
                                                          next_ids, finished, ctx_len, step, advance_ctx);
  int rc = check_launch("argmax_step_kernel");
  if (rc) return rc;
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  return check_launch("step_increment_kernel");
}
