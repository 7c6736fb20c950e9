// HF image-processor stage of run_ocr (tools.py:756-762 -> HF image_processing_qwen2_vl.py:62-88,
// 166-232): smart_resize, uint8 bicubic-antialias resize (torch CPU `upsample_bicubic2d_aa` uint8
// path semantics), rescale+normalize, patchify.  Bit-exact.
#include "common.cuh"
#include <math.h>
#include <map>
#include <mutex>
#include <vector>
#include "resize_kernels.cuh"

namespace ocrb {

// ───────────── weight tables (host, double, cached per (in,out,device)) ─────────────
struct AxisTable {
  int kmax = 0, prec = 0;
  int32_t *d_xmin = nullptr;   // [out]
  int32_t *d_xsize = nullptr;  // [out]
  int16_t *d_w = nullptr;      // [out * kmax]
};

static int build_axis_table(int in_size, int out_size, AxisTable *t) {
  AxisWeights h;
  compute_axis_weights(in_size, out_size, &h);
  t->kmax = h.kmax;
  t->prec = h.prec;
  OCRB_CUDA(cudaMalloc(&t->d_xmin, sizeof(int32_t) * out_size));
  OCRB_CUDA(cudaMalloc(&t->d_xsize, sizeof(int32_t) * out_size));
  OCRB_CUDA(cudaMalloc(&t->d_w, sizeof(int16_t) * h.w.size()));
  OCRB_CUDA(cudaMemcpy(t->d_xmin, h.xmin.data(), sizeof(int32_t) * out_size, cudaMemcpyHostToDevice));
  OCRB_CUDA(cudaMemcpy(t->d_xsize, h.xsize.data(), sizeof(int32_t) * out_size, cudaMemcpyHostToDevice));
  OCRB_CUDA(cudaMemcpy(t->d_w, h.w.data(), sizeof(int16_t) * h.w.size(), cudaMemcpyHostToDevice));
  return OCRB_OK;
}

static std::mutex g_tab_mu;
static std::map<std::tuple<int, int, int>, AxisTable> g_tabs;

static int get_axis_table(int in_size, int out_size, AxisTable *out) {
  int dev = 0;
  OCRB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_tuple(dev, in_size, out_size);
  auto it = g_tabs.find(key);
  if (it == g_tabs.end()) {
    AxisTable t;
    int rc = build_axis_table(in_size, out_size, &t);
    if (rc) return rc;
    it = g_tabs.emplace(key, t).first;
  }
  *out = it->second;
  return OCRB_OK;
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_smart_resize_host(int32_t H, int32_t W, int32_t factor, int64_t min_pixels, int64_t max_pixels,
                                      int32_t *out_H, int32_t *out_W) {
  OCRB_REQUIRE(H > 0 && W > 0 && factor > 0 && out_H && out_W, "smart_resize: bad arguments");
  const double mx = H > W ? H : W, mn = H > W ? W : H;
  OCRB_REQUIRE(mx / mn <= 200.0, "smart_resize: absolute aspect ratio must be smaller than 200");
  // Python round() = round-half-even on the double quotient
  long long h_bar = (long long)nearbyint((double)H / factor) * factor;
  long long w_bar = (long long)nearbyint((double)W / factor) * factor;
  if (h_bar * w_bar > max_pixels) {
    const double beta = sqrt(((double)H * (double)W) / (double)max_pixels);
    h_bar = (long long)floor((double)H / beta / factor) * factor;
    w_bar = (long long)floor((double)W / beta / factor) * factor;
    if (h_bar < factor) h_bar = factor;
    if (w_bar < factor) w_bar = factor;
  } else if (h_bar * w_bar < min_pixels) {
    const double beta = sqrt((double)min_pixels / ((double)H * (double)W));
    h_bar = (long long)ceil((double)H * beta / factor) * factor;
    w_bar = (long long)ceil((double)W * beta / factor) * factor;
  }
  *out_H = (int32_t)h_bar;
  *out_W = (int32_t)w_bar;
  return OCRB_OK;
}

extern "C" int ocrb_resize_bicubic_aa_u8(const uint8_t *src, uint8_t *dst, uint8_t *tmp, int32_t n_img, int32_t H,
                                         int32_t W, int32_t C, int32_t out_H, int32_t out_W, void *stream) {
  OCRB_REQUIRE(src && dst && tmp && n_img > 0 && H > 0 && W > 0 && out_H > 0 && out_W > 0 && (C == 1 || C == 3),
               "resize_bicubic_aa_u8: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t *hsrc = src;
  if (out_W != W) {
    AxisTable tx;
    int rc = get_axis_table(W, out_W, &tx);
    if (rc) return rc;
    resize_h_kernel<<<dim3(cdiv((long long)out_W * C, 256), n_img * H), 256, 0, st>>>(
        src, tmp, n_img * H, W, C, out_W, tx.d_xmin, tx.d_xsize, tx.d_w, tx.kmax, tx.prec);
    rc = check_launch("resize_h_kernel");
    if (rc) return rc;
    hsrc = tmp;
  }
  if (out_H != H) {
    AxisTable ty;
    int rc = get_axis_table(H, out_H, &ty);
    if (rc) return rc;
    resize_v_kernel<<<dim3(cdiv((long long)out_W * C, 256), out_H, n_img), 256, 0, st>>>(
        hsrc, dst, H, out_W * C, out_H, ty.d_xmin, ty.d_xsize, ty.d_w, ty.kmax, ty.prec);
    return check_launch("resize_v_kernel");
  }
  OCRB_CUDA(cudaMemcpyAsync(dst, hsrc, (size_t)n_img * H * out_W * C, cudaMemcpyDeviceToDevice, st));
  return OCRB_OK;
}

extern "C" int ocrb_normalize_patchify(const uint8_t *src, void *dst, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                       const int32_t *group_perm, int32_t out_dtype, void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && (C == 1 || C == 3), "normalize_patchify: bad arguments");
  OCRB_REQUIRE(H % 28 == 0 && W % 28 == 0 && H > 0 && W > 0, "normalize_patchify: H and W must be multiples of 28");
  OCRB_REQUIRE(out_dtype == 0 || out_dtype == 1, "normalize_patchify: out_dtype must be 0 (fp32) or 1 (bf16)");
  const int gh = H / 14, gw = W / 14;
  const long long total = (long long)n_img * gh * gw * 3 * 14 * 14;
  // CLIP mean/std as float32 times 255.0f (HF image_processing_backends.py:291-331)
  volatile float m0 = 0.48145466f * 255.0f, m1 = 0.4578275f * 255.0f, m2 = 0.40821073f * 255.0f;
  volatile float s0 = 0.26862954f * 255.0f, s1 = 0.26130258f * 255.0f, s2 = 0.27577711f * 255.0f;
  const int blocks = cdiv(total, 256);
  if (out_dtype == 0)
    normalize_patchify_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (float *)dst, H, W, C, gh, gw,
                                                                               group_perm, total, m0, m1, m2, s0, s1, s2);
  else
    normalize_patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        src, (__nv_bfloat16 *)dst, H, W, C, gh, gw, group_perm, total, m0, m1, m2, s0, s1, s2);
  return check_launch("normalize_patchify_kernel");
}
