// Kernels and host weight tables of the HF image-processor stage of run_ocr (tools.py:756-762 -> HF
// image_processing_qwen2_vl.py:62-88, 166-232): uint8 bicubic-antialias resize (horizontal pass, uint8 intermediate, vertical
// pass), rescale + normalize + patchify.  Bit-exact.  Definitions only -- no launches, no CUDA runtime calls -- so that tests/emu
// can compile this file for the host (OCRB_EMU) and run the kernels against the oracle and under the host sanitizers.
#pragma once
#ifndef OCRB_EMU
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#endif
#include <math.h>
#include <stdint.h>
#include <vector>

namespace ocrb {

static inline double cubic_aa(double x) {
  const double a = -0.5;
  x = fabs(x);
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}

// Separable uint8 bicubic-antialias weights of one axis (torch `upsample_bicubic2d_aa` uint8 path): per output index the first
// source index, the tap count and the int16 fixed-point taps ([out * kmax], precision chosen so the largest tap fits int16).
// Pure host arithmetic (double), shared by the product (resize_patchify.cu uploads the result) and the emulator tests.
struct AxisWeights {
  int kmax = 0, prec = 0;
  std::vector<int32_t> xmin, xsize;
  std::vector<int16_t> w;
};

static inline void compute_axis_weights(int in_size, int out_size, AxisWeights *t) {
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
  t->xmin.swap(xmin);
  t->xsize.swap(xsize);
  t->w.swap(iw);
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
  dst[(size_t)row * outW * C + ob] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
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
  dst[((size_t)img * outH + oy) * rowb + xb] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
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
