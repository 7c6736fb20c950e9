// HF image-processor stage of run_ocr (tools.py:756-762 -> HF image_processing_qwen2_vl.py:62-88,
// 166-232): smart_resize, uint8 bicubic-antialias resize (torch CPU `upsample_bicubic2d_aa` uint8
// path semantics), rescale+normalize, patchify.  Bit-exact.
#include "common.cuh"
#include <math.h>
#include <map>
#include <mutex>
#include <vector>

namespace ocrb {

// ───────────── weight tables (host, double, cached per (in,out,device)) ─────────────
struct AxisTable {
  int kmax = 0, prec = 0;
  int32_t *d_xmin = nullptr;   // [out]
  int32_t *d_xsize = nullptr;  // [out]
  int16_t *d_w = nullptr;      // [out * kmax]
};

static double cubic_aa(double x) {
  const double a = -0.5;
  x = fabs(x);
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}

static int build_axis_table(int in_size, int out_size, AxisTable *t) {
  const double scale = (double)in_size / (double)out_size;
  const double support = 2.0 * (scale > 1.0 ? scale : 1.0);
  const double inv = 1.0 / (scale > 1.0 ? scale : 1.0);
  const int kmax = (int)ceil(support) * 2 + 1;
  std::vector<int32_t> xmin(out_size), xsize(out_size);
  std::vector<double> wd((size_t)out_size * kmax, 0.0);
  double wmax = 0.0;
  for (int i = 0; i < out_size; ++i) {
    const double center = scale * (i + 0.5);
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int size = hi - lo;
    double tot = 0.0;
    for (int j = 0; j < size; ++j) {
      const double w = cubic_aa((j + lo - center + 0.5) * inv);
      wd[(size_t)i * kmax + j] = w;
      tot += w;
    }
    if (tot != 0.0)
      for (int j = 0; j < size; ++j) wd[(size_t)i * kmax + j] /= tot;
    for (int j = 0; j < size; ++j)
      if (wd[(size_t)i * kmax + j] > wmax) wmax = wd[(size_t)i * kmax + j];
    xmin[i] = lo;
    xsize[i] = size;
  }
  int prec = 0;
  while (prec < 22) {
    const int nxt = (int)(0.5 + wmax * (double)(1 << (prec + 1)));
    if (nxt >= (1 << 15)) break;
    ++prec;
  }
  std::vector<int16_t> iw((size_t)out_size * kmax, 0);
  for (size_t q = 0; q < wd.size(); ++q) {
    const double s = wd[q] * (double)(1 << prec);
    iw[q] = (int16_t)(wd[q] < 0 ? (int)(s - 0.5) : (int)(s + 0.5));
  }
  t->kmax = kmax;
  t->prec = prec;
  OCRB_CUDA(cudaMalloc(&t->d_xmin, sizeof(int32_t) * out_size));
  OCRB_CUDA(cudaMalloc(&t->d_xsize, sizeof(int32_t) * out_size));
  OCRB_CUDA(cudaMalloc(&t->d_w, sizeof(int16_t) * iw.size()));
  OCRB_CUDA(cudaMemcpy(t->d_xmin, xmin.data(), sizeof(int32_t) * out_size, cudaMemcpyHostToDevice));
  OCRB_CUDA(cudaMemcpy(t->d_xsize, xsize.data(), sizeof(int32_t) * out_size, cudaMemcpyHostToDevice));
  OCRB_CUDA(cudaMemcpy(t->d_w, iw.data(), sizeof(int16_t) * iw.size(), cudaMemcpyHostToDevice));
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

// Horizontal pass: out[y][ox][c] = clamp((sum_k w[ox][k] * in[y][xmin[ox]+k][c] + 2^(p-1)) >> p)
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int rows_total, int W, int C, int outW,
                const int32_t *__restrict__ xmin, const int32_t *__restrict__ xsize, const int16_t *__restrict__ w,
                int kmax, int prec) {
  const int ob = blockIdx.x * blockDim.x + threadIdx.x;  // output byte within the row
  const int row = blockIdx.y;
  if (ob >= outW * C) return;
  const int ox = ob / C, c = ob - ox * C;
  const uint8_t *r = src + (size_t)row * W * C;
  const int lo = xmin[ox], n = xsize[ox];
  const int16_t *wk = w + (size_t)ox * kmax;
  int acc = 1 << (prec - 1);
  for (int k = 0; k < n; ++k) acc += (int)wk[k] * (int)r[(lo + k) * C + c];
  int v = acc >> prec;
  dst[(size_t)row * outW * C + ob] = (uint8_t)min(max(v, 0), 255);
}

// Vertical pass: out[oy][x][c] = clamp((sum_k w[oy][k] * in[ymin[oy]+k][x][c] + 2^(p-1)) >> p)
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int rowb, int outH,
                const int32_t *__restrict__ ymin, const int32_t *__restrict__ ysize, const int16_t *__restrict__ w,
                int kmax, int prec) {
  const int xb = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int img = blockIdx.z;
  if (xb >= rowb) return;
  const uint8_t *im = src + (size_t)img * H * rowb;
  const int lo = ymin[oy], n = ysize[oy];
  const int16_t *wk = w + (size_t)oy * kmax;
  int acc = 1 << (prec - 1);
  for (int k = 0; k < n; ++k) acc += (int)wk[k] * (int)im[(size_t)(lo + k) * rowb + xb];
  int v = acc >> prec;
  dst[((size_t)img * outH + oy) * rowb + xb] = (uint8_t)min(max(v, 0), 255);
}

// ───────────── normalize + patchify ─────────────
// One thread per (patch, channel, py, 2 px): reads 2 source bytes (1 if gray), writes the value for
// both temporal frames.  Output row layout: feature = ((c*2 + t)*14 + py)*14 + px.
template <typename OutT>
__global__ void __launch_bounds__(256)
normalize_patchify_kernel(const uint8_t *__restrict__ src, OutT *__restrict__ dst, int H, int W, int C, int gh, int gw,
                          const int32_t *__restrict__ group_perm, long long total, float m0, float m1, float m2, float s0,
                          float s1, float s2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx -> (out_patch, c, py, px)
  int px = (int)(idx % 14);
  long long r = idx / 14;
  int py = (int)(r % 14);
  r /= 14;
  int c = (int)(r % 3);
  long long out_patch = r / 3;
  const int per_img = gh * gw;
  // source patch: optional group permutation (groups of 4 patches)
  long long src_patch = out_patch;
  if (group_perm) src_patch = (long long)group_perm[out_patch >> 2] * 4 + (out_patch & 3);
  const int img = (int)(src_patch / per_img);
  const int pin = (int)(src_patch - (long long)img * per_img);
  // merge-group order: pin = ((gy2 * (gw/2) + gx2) * 2 + my) * 2 + mx
  const int mx = pin & 1, my = (pin >> 1) & 1;
  const int g = pin >> 2;
  const int gx2 = g % (gw >> 1), gy2 = g / (gw >> 1);
  const int y = (gy2 * 2 + my) * 14 + py;
  const int x = (gx2 * 2 + mx) * 14 + px;
  const uint8_t v = src[(((size_t)img * H + y) * W + x) * C + (C == 3 ? c : 0)];
  const float m = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float s = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const float f = __fdiv_rn(__fsub_rn((float)v, m), s);
  OutT o;
  if constexpr (sizeof(OutT) == 4) o = f;
  else o = __float2bfloat16_rn(f);
  OutT *row = dst + out_patch * 1176;
  row[((c * 2 + 0) * 14 + py) * 14 + px] = o;
  row[((c * 2 + 1) * 14 + py) * 14 + px] = o;
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
