/* libocrb200 -- C ABI of the B200-native OCR read path.
 *
 * This is the drop-in boundary for the reference's `ocr_agent.tools` hot functions
 * (/root/reference/ocr_agent/tools.py).  The reference is pure Python; its "FFI" for
 * this path would be a ctypes binding, shown in INTEGRATION.md.  Every entry point:
 *   - is `extern "C"`, takes plain pointers / sizes (no torch types);
 *   - returns 0 on success, a negative OCRB_E* code on failure (text via ocrb_last_error());
 *   - takes DEVICE pointers unless the parameter name ends in `_host`;
 *   - launches on the given `stream` (a cudaStream_t passed as void*) and does not synchronise
 *     unless stated.
 * sm_100a only.  There is no CPU fallback anywhere in this library.
 */
#ifndef OCRB200_H
#define OCRB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRB_OK 0
#define OCRB_EINVAL (-1)
#define OCRB_ECUDA (-2)
#define OCRB_EUNSUPPORTED (-3)

/* ───────────── library ───────────── */
int ocrb_version(void);              /* 100*major + minor */
const char *ocrb_last_error(void);   /* thread-local text of the last failure */
/* Number of kernel launches this library has issued in this process (bench `gpu_launches`). */
uint64_t ocrb_launch_count(void);
void ocrb_launch_count_reset(void);

/* ───────────── text ops: tools.py:69-100 (levenshtein, _levenshtein_words) ─────────────
 * One warp per pair, anti-diagonal wavefront.  seq*: concatenated int32 symbols (code points or
 * word ids); off*: n_pairs+1 int32 offsets.  out[p] = unit-cost edit distance of pair p.
 * workspace: int32[n_pairs * (max_len_b + 1)]. */
int ocrb_levenshtein_batch(const int32_t *seq_a, const int32_t *off_a, const int32_t *seq_b,
                           const int32_t *off_b, int32_t n_pairs, int32_t max_len_b,
                           int32_t *out, int32_t *workspace, void *stream);

/* tools.py:465-493 (_align_to_backbone): LCS table + backtrack with the reference tie rule
 * (`dp[i-1][j] >= dp[i][j-1]` -> up).  Symbols are ids of LOWER-CASED words.  One CTA per pair.
 * aligned: int32[total backbone symbols] (same offsets as off_bb): index into the pair's word
 * list, or -1.  workspace: uint8[sum_p n_p*m_p] direction table; ws_off: int64[n_pairs] offsets
 * into it.  max_len_bb bounds the shared-memory diagonals (<= 16000). */
int ocrb_lcs_align_batch(const int32_t *seq_bb, const int32_t *off_bb, const int32_t *seq_w,
                         const int32_t *off_w, int32_t n_pairs, int32_t max_len_bb,
                         int32_t *aligned, uint8_t *workspace, const int64_t *ws_off, void *stream);

/* ───────────── image ops: tools.py:503-573 via OpenCV 4.13 semantics ─────────────
 * Images are uint8, HWC (C = 1 or 3), n_img images of identical H x W, contiguous. */

/* tools.py:510 cv2.cvtColor(RGB2GRAY): Y = (9798R + 19235G + 3735B + 16384) >> 15 */
int ocrb_rgb2gray_u8(const uint8_t *src_rgb, uint8_t *dst_gray, int32_t n_img, int32_t H,
                     int32_t W, void *stream);

/* tools.py:511-512 cv2.createCLAHE(3.0,(8,8)).apply(gray).  lut_ws: uint8[n_img*64*256]. */
int ocrb_clahe_u8(const uint8_t *src_gray, uint8_t *dst_gray, int32_t n_img, int32_t H, int32_t W,
                  uint8_t *lut_ws, void *stream);

/* tools.py:503-516 (_apply_high_contrast) in one call: RGB (C == 3) or gray (C == 1) page -> CLAHE(3.0,(8,8)) of its gray
 * version.  For C == 3 the RGB -> gray conversion is fused into the tile-histogram pass (no pass of its own) and the gray
 * page is left in gray_ws: uint8[n_img*H*W] (ignored for C == 1).  lut_ws: uint8[n_img*64*256]. */
int ocrb_high_contrast_u8(const uint8_t *src, uint8_t *dst_gray, uint8_t *gray_ws, uint8_t *lut_ws, int32_t n_img,
                          int32_t H, int32_t W, int32_t C, void *stream);

/* tools.py:519-531 (_apply_binarize) in one call: RGB or gray page -> adaptive Gaussian threshold 21/10 of its gray version;
 * for C == 3 the gray tile is computed while it is staged into shared memory (no gray page in memory at all). */
int ocrb_binarize_u8(const uint8_t *src, uint8_t *dst_gray, int32_t n_img, int32_t H, int32_t W, int32_t C,
                     void *stream);

/* tools.py:527-529 cv2.adaptiveThreshold(gray,255,GAUSSIAN_C,BINARY,21,10) */
int ocrb_adaptive_gauss_thresh_u8(const uint8_t *src_gray, uint8_t *dst_gray, int32_t n_img,
                                  int32_t H, int32_t W, void *stream);

/* tools.py:541-542 cv2.filter2D(img,-1,[[0,-1,0],[-1,5,-1],[0,-1,0]]), reflect-101 */
int ocrb_sharpen3x3_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W,
                       int32_t C, void *stream);

/* tools.py:598-614 (_apply_remove_lines up to the inpaint): ruled-line mask = dilate_1x3(open_{W/4 x 1}(
 * adaptiveThreshold(255 - gray, 255, MEAN_C, BINARY, 15, -2))), bit-exact.  mask: uint8[n_img*H*W]; nonzero:
 * int32[n_img], set to 1 when the image's mask has any pixel; tmp: uint8[n_img*H*W]. */
int ocrb_remove_lines_mask_u8(const uint8_t *src, uint8_t *mask, int32_t *nonzero, uint8_t *tmp, int32_t n_img,
                              int32_t H, int32_t W, int32_t C, void *stream);

/* tools.py:617 cv2.inpaint(img, mask, radius, cv2.INPAINT_TELEA), bit-exact against OpenCV 4.13 at radius 3, the radius the
 * reference passes (1..7 accepted; at radii other than 2 and 3 a few flat-region pixels can differ from OpenCV by 1-4).  src, dst: uint8[n_img*H*W*C], C = 1 or 3, src != dst; mask: uint8[n_img*H*W], non-zero =
 * repaint; pages with an empty mask are copied.  H, W >= 2.  ws: ocrb_inpaint_workspace_bytes(n_img, H, W) bytes,
 * 16-byte aligned, contents irrelevant.  Row segments of the mask separated by 2*radius+2 clean rows are marched
 * concurrently (one warp each); within a segment the march order is OpenCV's. */
int64_t ocrb_inpaint_workspace_bytes(int32_t n_img, int32_t H, int32_t W);
int ocrb_inpaint_telea_u8(const uint8_t *src, const uint8_t *mask, uint8_t *dst, uint8_t *ws, int32_t n_img, int32_t H,
                          int32_t W, int32_t C, int32_t radius, void *stream);

/* tools.py:582-587 (_apply_denoise), bit-exact against OpenCV 4.13:
 *   C == 1: cv2.fastNlMeansDenoising(gray, None, 10, 7, 21) (PIL mode "L" pages);
 *   C == 3: cv2.fastNlMeansDenoisingColored(rgb, None, 10, 10, 7, 21) -- 8-bit Lab (channel 0 read as blue, as the
 *           reference's RGB array is), NLM on L and on the (a, b) pair, back.
 * ws: uint8[n_img*H*W*6], 2-byte aligned, for C == 3 (ignored for C == 1).  src != dst. */
int ocrb_nlm_denoise_u8(const uint8_t *src, uint8_t *dst, uint8_t *ws, int32_t n_img, int32_t H, int32_t W,
                        int32_t C, void *stream);

/* Host-only: the integer tables ocrb_nlm_denoise_u8 uploads (Lab cube-root table int32[3072], LabToYF int32[512],
 * forward + inverse Lab coefficients int32[18], NLM weight prefixes for 1 and 2 channels int32[2048] each), so that
 * they can be compared with the oracle on a machine without a GPU. */
int ocrb_denoise_tables_host(int32_t *cbrt_tab, int32_t *lab_yf, int32_t *coef, int32_t *w1, int32_t *w2);

/* tools.py:556-564: dark-pixel (<128) row extents -> convex hull -> min-area-rect angle ->
 * rotation matrix about (W//2, H//2).  src has C channels (gray computed on the fly for C=3).
 * out_angle: double[n_img] (NaN when <= 100 dark pixels: image must be left unchanged),
 * out_M: double[n_img*6] forward matrix of cv2.getRotationMatrix2D.
 * ext_ws: int32[n_img*H*3].  hull_ws: int32[n_img*(4*H+8)*2] (only touched for H > ~2600, where the hull tree does not
 * fit in shared memory). */
int ocrb_deskew_angle(const uint8_t *src, int32_t n_img, int32_t H, int32_t W, int32_t C,
                      double *out_angle, double *out_M, int32_t *ext_ws, int32_t *hull_ws,
                      void *stream);

/* tools.py:568-570 cv2.warpAffine(img, M, (W,H), INTER_CUBIC, BORDER_REPLICATE).  M: double[n_img*6]
 * on the device (forward matrix; inverted in-kernel as OpenCV does).  An image whose M[0] is NaN
 * is copied unchanged (the `<= 100 dark pixels` early return). */
int ocrb_warp_affine_cubic_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H,
                              int32_t W, int32_t C, const double *M, void *stream);

/* HF image_processing_qwen2_vl.py:62-88 smart_resize (host arithmetic, no device work). */
int ocrb_smart_resize_host(int32_t H, int32_t W, int32_t factor, int64_t min_pixels,
                           int64_t max_pixels, int32_t *out_H, int32_t *out_W);

/* torchvision uint8 bicubic-antialias resize (horizontal pass, uint8 intermediate, vertical pass).
 * tmp: uint8[n_img*H*outW*C]. */
int ocrb_resize_bicubic_aa_u8(const uint8_t *src, uint8_t *dst, uint8_t *tmp, int32_t n_img,
                              int32_t H, int32_t W, int32_t C, int32_t out_H, int32_t out_W,
                              void *stream);

/* HF rescale+normalize+patchify: y=(float(x)-mean*255)/(std*255), frame duplicated x2, written as
 * [n_img*gh*gw, 1176] in 2x2 merge-group order.  src: HWC with C in {1,3} (C=1 is replicated,
 * = PIL convert("RGB") of an "L" image).  group_perm: optional int32[n_img*gh*gw/4] gather map
 * (output group g reads source group group_perm[g], global index) or NULL.
 * out_dtype: 0 = fp32 (HF `pixel_values`), 1 = bf16 (what the vision tower consumes). */
int ocrb_normalize_patchify(const uint8_t *src, void *dst, int32_t n_img, int32_t H, int32_t W,
                            int32_t C, const int32_t *group_perm, int32_t out_dtype, void *stream);

/* ───────────── dense ops of the VLM (HF modeling_qwen2_5_vl.py) ───────────── */

/* Epilogues of ocrb_gemm_bf16 */
#define OCRB_EPI_NONE 0      /* D = bf16(acc [+ bias]) */
#define OCRB_EPI_RESIDUAL 1  /* D = bf16( bf16(acc [+ bias]) + residual ) */
#define OCRB_EPI_SWIGLU 2    /* weight rows interleaved per 64: [gate64|up64]; D[:, n/2] =
                                bf16( bf16(silu(bf16(g))) * bf16(u) ), g/u = acc [+ bias] */
#define OCRB_EPI_GELU 3      /* D = bf16(gelu_erf(bf16(acc + bias))) */

/* D[M,N'] = epilogue(A[M,K] * W[N,K]^T).  A, W, D, residual: bf16 row-major with row strides
 * lda/ldw/ldd/ldr (elements, multiples of 8).  bias: bf16[N] or NULL.  tcgen05 + TMEM + TMA. */
int ocrb_gemm_bf16(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd,
                   int32_t M, int32_t N, int32_t K, const void *bias, const void *residual,
                   int64_t ldr, int32_t epilogue, void *stream);

/* Weight-streaming skinny GEMM for decode on tcgen05 (swap-AB: the 128x64 weight tile is the UMMA M operand,
 * the B <= 128 activation rows are the N operand; TMA weight ring; stream-K split over all SMs with a
 * deterministic, batch-invariant fix-up).  B in 1..128, same epilogues as ocrb_gemm_bf16.
 * Optional RMSNorm prologue: if norm_w != NULL, X is first normalised row-wise the HF way
 * (fp32 normalise -> bf16 -> * weight in bf16) with `eps`.
 * workspace: ocrb_skinny_workspace_bytes() bytes, ZERO-initialised once by the caller and then owned by
 * this entry point (it holds stream-K partials and their ready flags; flags are returned to zero by every
 * launch).  One workspace must not be shared by launches that can run concurrently. */
int64_t ocrb_skinny_workspace_bytes(void);
int ocrb_skinny_gemm_bf16(const void *X, int64_t ldx, const void *W, int64_t ldw, void *D, int64_t ldd,
                          int32_t B, int32_t N, int32_t K, const void *bias, const void *residual,
                          int64_t ldr, int32_t epilogue, const void *norm_w, float eps, void *workspace,
                          void *stream);

/* A chain of up to OCRB_CHAIN_MAX dependent skinny linears in ONE persistent launch (one decode step's
 * o_proj -> [RMSNorm] gate/up+SwiGLU -> down_proj -> [RMSNorm] qkv of the next layer / lm_head; HF decoder layer
 * modeling_qwen2_5_vl.py:839-879 inside the generation loop utils.py:2743-2806).  Linear g+1 may read what linear g
 * (or any earlier one) wrote: the kernel orders them with device-side counters while the weight stream of the whole
 * chain keeps HBM busy.  Each linear has the meaning of one ocrb_skinny_gemm_bf16 call with the same arguments.
 * `residual` may equal `D` (in-place update).  norm_w != NULL: K <= 8192.
 * workspace: ocrb_chain_workspace_bytes() bytes, ZERO-initialised once by the caller, then owned by these entry points
 * (counters are returned to zero by every launch); not shared by launches that can run concurrently. */
#define OCRB_CHAIN_MAX 5
typedef struct ocrb_chain_linear {
  const void *X; int64_t ldx;          /* [B, K] activations (row stride in elements) */
  const void *W; int64_t ldw;          /* [N, K] weights */
  void *D; int64_t ldd;                /* [B, N'] output */
  int32_t N, K;
  const void *bias;                    /* [N] or NULL */
  const void *residual; int64_t ldr;   /* OCRB_EPI_RESIDUAL */
  int32_t epilogue;
  float eps;
  const void *norm_w;                  /* [K] RMSNorm weight applied to X first, or NULL */
} ocrb_chain_linear;
int64_t ocrb_chain_workspace_bytes(void);
int ocrb_skinny_chain_bf16(const ocrb_chain_linear *lin, int32_t n, int32_t B, void *workspace, void *stream);

/* A whole decode step as ONE persistent launch: a plan of up to OCRB_CHAIN_PLAN_MAX_OPS ops, each a skinny linear
 * (above) or one layer's paged decode attention with the meaning of ocrb_decode_attention (mRoPE of the new token's
 * q / k, KV append, split-KV attention, combine).  Op g+1 may read what op g wrote.  The attention ops of a plan share
 * block table, context lengths and geometry (one per layer: only the caches differ); hd must be 128.
 * ocrb_chain_plan_build writes the plan (ocrb_chain_plan_bytes(n) bytes of DEVICE memory, 64-byte aligned) and
 * synchronises `stream`; ocrb_chain_plan_run only launches (CUDA-graph capturable).  The plan bakes in every pointer. */
#define OCRB_CHAIN_PLAN_MAX_OPS 192
#define OCRB_CHAIN_OP_LINEAR 0
#define OCRB_CHAIN_OP_ATTENTION 1
typedef struct ocrb_chain_attention {
  const void *qkv; int64_t ldqkv;      /* [B, (n_q + 2 n_kv) * hd] fused q|k|v of the step (pre-RoPE) */
  void *k_cache; void *v_cache;        /* the layer's paged cache [n_cache_pages][n_kv][page_size][hd] */
  int32_t n_cache_pages;
  const int32_t *block_table; int32_t max_pages;
  const int32_t *ctx_len;              /* [B] cached tokens per sequence (the new token goes to position ctx_len) */
  int32_t page_size, n_q, n_kv, hd;
  const void *cosT; const void *sinT;  /* [B, hd] bf16 mRoPE tables of the new position */
  float scale;
  void *out; int64_t ldo;              /* [B, n_q * hd] */
  float *split_ws; int32_t n_splits;   /* as ocrb_decode_attention */
} ocrb_chain_attention;
typedef struct ocrb_chain_op {
  int32_t kind; int32_t reserved;
  ocrb_chain_linear lin;
  ocrb_chain_attention att;
} ocrb_chain_op;
int64_t ocrb_chain_plan_bytes(int32_t n_ops);
int ocrb_chain_plan_build(const ocrb_chain_op *ops, int32_t n, int32_t B, void *workspace, void *plan, void *stream);
int ocrb_chain_plan_run(const void *plan, int32_t n, int32_t B, void *workspace, void *stream);

/* HF Qwen2_5_VLRMSNorm (modeling:66-71): y = w * bf16( x_f32 * rsqrt(mean(x^2)+eps) ) */
int ocrb_rmsnorm_bf16(const void *x, int64_t ldx, const void *w, void *y, int64_t ldy, int32_t rows,
                      int32_t dim, float eps, void *stream);

/* Vision 2-D RoPE (modeling:156-167) applied in place to q and k inside a fused qkv buffer
 * [S, 3*heads*hd]: fp32 math, bf16 out.  cos/sin: fp32 [S, hd]. */
int ocrb_rope_vision(void *qkv, int32_t S, int32_t heads, int32_t hd, const float *cos,
                     const float *sin, void *stream);

/* Text mRoPE (modeling:659-669) in bf16 arithmetic as HF does (cos/sin already bf16, already
 * section-mixed per token: [T, hd]).  q: [T, n_q*hd] (stride ldq), k: [T, n_kv*hd] (stride ldk). */
int ocrb_rope_text(void *q, int64_t ldq, void *k, int64_t ldk, int32_t T, int32_t n_q, int32_t n_kv,
                   int32_t hd, const void *cos, const void *sin, void *stream);

/* Variable-length flash attention (non-causal or causal), bf16 in/out, fp32 softmax.
 * q: [T, n_q, hd] strides (ldq per token), k/v: [T, n_kv, hd]; cu_seqlens: int32[n_seq+1];
 * n_q % n_kv == 0 (GQA).  hd in {80, 128}.  out: [T, n_q*hd] stride ldo. */
int ocrb_attention_varlen(const void *q, int64_t ldq, const void *k, int64_t ldk, const void *v,
                          int64_t ldv, void *out, int64_t ldo, const int32_t *cu_seqlens,
                          int32_t n_seq, int32_t max_seqlen, int32_t n_q, int32_t n_kv, int32_t hd,
                          float scale, int32_t causal, void *stream);

/* Paged KV cache: pages of `page_size` tokens; k_cache/v_cache: bf16 [n_pages, n_kv, page_size, hd] (the tokens of a
 * (page, kv head) are one contiguous block, so a 16-key attention tile is a single 4 KiB TMA box).
 * block_table: int32[B, max_pages_per_seq].  ctx_len: int32[B] on the device (tokens already cached). */

/* Scatter T tokens of k/v (prefill) into the paged cache: token t of sequence s goes to
 * position pos0[s] + (t - cu_seqlens[s]). */
int ocrb_kv_write_prefill(const void *k, int64_t ldk, const void *v, int64_t ldv, void *k_cache,
                          void *v_cache, const int32_t *block_table, int32_t max_pages,
                          const int32_t *cu_seqlens, int32_t n_seq, int32_t T, int32_t page_size,
                          int32_t n_kv, int32_t hd, void *stream);

/* Flash attention on tcgen05 / TMEM / TMA for long sequences (vision full-attention blocks: hd 80, non-causal;
 * decoder prefill: hd 128, causal, GQA).  Same tensors as ocrb_attention_varlen plus total_tokens (rows of q/k/v/out);
 * pointers 16-byte aligned, strides multiples of 8.  Replaces the attention inside HF's
 * Qwen2_5_VLVisionAttention / Qwen2_5_VLAttention forward (modeling_qwen2_5_vl.py:231-283,718-760). */
int ocrb_flash_attention_bf16(const void *q, int64_t ldq, const void *k, int64_t ldk, const void *v,
                              int64_t ldv, void *out, int64_t ldo, const int32_t *cu_seqlens,
                              int32_t n_seq, int32_t total_tokens, int32_t max_seqlen, int32_t n_q,
                              int32_t n_kv, int32_t hd, float scale, int32_t causal, void *stream);

/* One decode step of attention for B sequences: applies mRoPE (bf16) to q,k of the new token,
 * appends k,v at position ctx_len[b], attends over ctx_len[b]+1 tokens.  qkv: [B, (n_q+2*n_kv)*hd].
 * cos/sin: bf16 [B, hd] for this step.  out: [B, n_q*hd].  hd = 64 or 128, n_q / n_kv <= 16, page_size % 16 == 0.
 * k_cache / v_cache: one layer's cache, bf16 [n_cache_pages, n_kv, page_size, hd] (read through TMA).
 * Split-KV: every sequence's max_pages*page_size key positions are cut into n_splits ranges, one work item per
 * (range, kv head, sequence); split_ws: fp32 [B * n_q * n_splits * (hd + 2)] partials (max, sum, o[hd]). */
int ocrb_decode_attention(const void *qkv, int64_t ldqkv, void *k_cache, void *v_cache,
                          int32_t n_cache_pages, const int32_t *block_table, int32_t max_pages,
                          const int32_t *ctx_len, int32_t B, int32_t page_size, int32_t n_q, int32_t n_kv,
                          int32_t hd, const void *cos, const void *sin, float scale, void *out,
                          int64_t ldo, float *split_ws, int32_t n_splits, void *stream);

/* Greedy pick (HF generation/utils.py:2793 argmax over fp32 logits, lowest index on ties) +
 * sequence bookkeeping for one decode step, all on the device:
 *   tok = argmax(logits[b]); if finished[b] tok = pad; out_tokens[b*max_new + step[0]] = tok;
 *   finished[b] |= tok == eos; ctx_len[b] += 1; next_ids[b] = tok; (step[0] += 1 by block 0)
 * logits: bf16 [B, V] (stride ldl). */
int ocrb_argmax_step(const void *logits, int64_t ldl, int32_t B, int32_t V, int32_t eos,
                     int32_t pad, int32_t max_new, int32_t *out_tokens, int32_t *next_ids,
                     int32_t *finished, int32_t *ctx_len, int32_t *step, int32_t advance_ctx,
                     void *stream);

/* Embedding gather: out[t] = table[ids[t]] (bf16 rows of `dim`). */
int ocrb_embed_gather(const void *table, const int32_t *ids, void *out, int32_t T, int32_t dim,
                      void *stream);

/* Row gather/scatter of bf16 rows: dst[dst_idx[i]] = src[src_idx[i]] (NULL index = identity). */
int ocrb_rows_copy(const void *src, int64_t lds, const int32_t *src_idx, void *dst, int64_t ldd,
                   const int32_t *dst_idx, int32_t n_rows, int32_t dim, void *stream);

/* x[r, :dim] = bf16(x + y), row strides ldx/ldy: the residual add after a tensor-parallel all-reduce of the
 * row-parallel o_proj / down_proj outputs (HF base_model_tp_plan, configuration_qwen2_5_vl.py:90-98). */
int ocrb_residual_add_bf16(void *x, int64_t ldx, const void *y, int64_t ldy, int32_t rows, int32_t dim,
                           void *stream);

/* ───────────── tensor-parallel collectives over NVLink peer memory (BASELINE configs[4]) ─────────────
 * ocrb_comm_ipc_handle / ocrb_comm_ipc_open: CUDA-IPC plumbing so that every rank (one process per GPU) can map the
 * other ranks' partial-sum buffers and flag arrays; handle64 is a 64-byte cudaIpcMemHandle_t, `offset` the position of
 * `ptr` inside its allocation.
 * ocrb_allreduce_residual_bf16: x[rows, dim] += bf16(sum_r partial_r[rows, dim]) in ONE kernel per rank: remote flag
 * stores announce the partial, peers' partials are read directly over NVLink and added in rank order (bit-identical on
 * all ranks).  data_ptrs / flag_ptrs: host arrays of `world` device pointers (this GPU's mappings); flags int32 [16][8]
 * and seq int32 [16] zero-initialised once.  Replaces ncclAllReduce + residual add after the row-parallel o_proj /
 * down_proj GEMMs of the decode step (HF configuration_qwen2_5_vl.py:90-98 "rowwise"). */
int ocrb_comm_ipc_handle(const void *ptr, void *handle64, int64_t *offset);
int ocrb_comm_ipc_open(const void *handle64, int64_t offset, void **ptr);
int ocrb_allreduce_residual_bf16(void *x, int64_t ldx, const void *const *data_ptrs, void *const *flag_ptrs,
                                 int32_t world, int32_t rank, int32_t *seq, int32_t rows, int32_t dim,
                                 int64_t ld_part, void *stream);

/* Row-parallel linear of the tensor-parallel decode step (o_proj / down_proj, HF "rowwise") WITH its all-reduce and residual
 * add: x[B, N] += bf16(sum over ranks of bf16(X_r[B, K] @ W_r[N, K]^T)), every rank calling with its K-slice.  When the
 * linear runs on the cluster kernel (few 128-row tiles, K >= 1024) the exchange is part of the GEMM's epilogue: each
 * (tile, cluster CTA) announces its partial to the peers with one remote flag store, waits for theirs, reads their
 * partials of its tile over NVLink peer memory and finishes the residual stream in place -- no second kernel, and a
 * tile's reduction overlaps the weight streaming of the others.  Otherwise (tiny shapes) the partial is written and
 * ocrb_allreduce_residual_bf16 follows; both routes produce identical bits.  data_ptrs: every rank's partial slot for
 * this call (this rank's own at [rank]), row stride ld_part; fused_flag_ptrs: every rank's int32 [512][8] flag array and
 * fused_seq: int32[512] (zero-initialised once; used by the fused route); flag_ptrs / seq: the arrays of
 * ocrb_allreduce_residual_bf16 (fallback route).  allow_fused: 0 = two-kernel route; 1 = fused, flag + pull (above); 2 = fused,
 * LL protocol: every partial travels with its flag -- two bf16 values of adjacent rows and the call index in one 8-byte cell,
 * pushed by one 8-byte store into every peer's receive buffer ll_ptrs[r] (uint64 [2][8][64][ld_part / 2], zero-initialised
 * once; ll_seq: int32[2], zero-initialised once), so the exchange is ONE one-way NVLink latency and the receiver polls its
 * own memory (B <= 64).  B <= 128. */
int ocrb_skinny_rowparallel_tp_bf16(const void *X, int64_t ldx, const void *W, int64_t ldw, int32_t B, int32_t N,
                                    int32_t K, void *x, int64_t ldx_res, const void *const *data_ptrs,
                                    int64_t ld_part, void *const *fused_flag_ptrs, int32_t *fused_seq,
                                    void *const *flag_ptrs, int32_t *seq, void *const *ll_ptrs, int32_t *ll_seq,
                                    int32_t world, int32_t rank, void *workspace, int32_t allow_fused, void *stream);
/* 1 when the last ocrb_skinny_rowparallel_tp_bf16 call ran the exchange inside the GEMM kernel. */
int ocrb_skinny_rowparallel_tp_was_fused(void);

/* Tensor-parallel greedy step for a vocab-split lm_head (HF `lm_head: colwise_rep`, configuration_qwen2_5_vl.py:90-98, reached
 * from tools.py:764-765): instead of all-gathering B x V logits, each rank reduces its [B, vl] slice (rank r = vocabulary rows
 * [r*vl, (r+1)*vl)) to one (max, global index) pair per sequence, the pairs are exchanged through peer memory, ties go to
 * the lowest index (= first-index arg max over the full vocabulary), then the bookkeeping of ocrb_argmax_step runs with the
 * winning token.  pair_ptrs / flag_ptrs: host arrays of `world` device pointers valid on this GPU (uint64 [2][max_rows] and
 * int32 [max_rows][8] per rank, zero-initialised once); seq: int32[max_rows], private, zero-initialised once. */
int ocrb_tp_argmax_step(const void *logits_local, int64_t ldl, int32_t B, int32_t vl, const void *const *pair_ptrs,
                        void *const *flag_ptrs, int32_t world, int32_t rank, int32_t *seq, int32_t max_rows,
                        int32_t eos, int32_t pad, int32_t max_new, int32_t *out_tokens, int32_t *next_ids,
                        int32_t *finished, int32_t *ctx_len, int32_t *step, int32_t advance_ctx, void *stream);

/* Per-step mRoPE tables for decode: pos[b] = ctx_len[b] + rope_delta[b]; writes bf16 cos/sin [B, hd]
 * (all three mrope sections share the position for text tokens). inv_freq: fp32[hd/2]. */
int ocrb_decode_rope_table(const int32_t *ctx_len, const int32_t *rope_delta, const float *inv_freq,
                           int32_t B, int32_t hd, void *cos, void *sin, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OCRB200_H */
