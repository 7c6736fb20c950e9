// TEST INFRASTRUCTURE: runs the kernels of handwritten-ocr_b200/csrc/resize_kernels.cuh on the CPU through tests/emu/cuda_emu.h
// with the launch sequence of the product's C ABI (resize_patchify.cu).  Never shipped.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/resize_kernels.cuh"

#include <cstring>

using namespace ocrb;

static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

extern "C" int emu_resize_bicubic_aa(const uint8_t *src, uint8_t *dst, uint8_t *tmp, int n_img, int H, int W, int C, int out_H,
                                     int out_W) {
  const uint8_t *hsrc = src;
  if (out_W != W) {
    AxisWeights tx;
    compute_axis_weights(W, out_W, &tx);
    emu::launch(dim3(cdivu((long long)out_W * C, 256), n_img * H), dim3(256), 0, [&] {
      resize_h_kernel(src, tmp, n_img * H, W, C, out_W, tx.xmin.data(), tx.xsize.data(), tx.w.data(), tx.kmax, tx.prec);
    });
    hsrc = tmp;
  }
  if (out_H != H) {
    AxisWeights ty;
    compute_axis_weights(H, out_H, &ty);
    emu::launch(dim3(cdivu((long long)out_W * C, 256), out_H, n_img), dim3(256), 0, [&] {
      resize_v_kernel(hsrc, dst, H, out_W * C, out_H, ty.xmin.data(), ty.xsize.data(), ty.w.data(), ty.kmax, ty.prec);
    });
    return 0;
  }
  std::memcpy(dst, hsrc, (size_t)n_img * H * out_W * C);
  return 0;
}

extern "C" int emu_normalize_patchify_f32(const uint8_t *src, float *dst, int n_img, int H, int W, int C,
                                          const int32_t *group_perm) {
  if (H % 28 || W % 28) return -1;
  const int gh = H / 14, gw = W / 14;
  const long long total = (long long)n_img * gh * gw * 3 * 14 * 14;
  volatile float m0 = 0.48145466f * 255.0f, m1 = 0.4578275f * 255.0f, m2 = 0.40821073f * 255.0f;
  volatile float s0 = 0.26862954f * 255.0f, s1 = 0.26130258f * 255.0f, s2 = 0.27577711f * 255.0f;
  const float a0 = m0, a1 = m1, a2 = m2, b0 = s0, b1 = s1, b2 = s2;
  emu::launch(dim3(cdivu(total, 256)), dim3(256), 0, [&] {
    normalize_patchify_kernel<float>(src, dst, H, W, C, gh, gw, group_perm, total, a0, a1, a2, b0, b1, b2);
  });
  return 0;
}
