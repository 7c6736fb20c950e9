// TEST INFRASTRUCTURE: runs the kernels of handwritten-ocr_b200/csrc/denoise_kernels.cuh on the CPU through tests/emu/cuda_emu.h
// with the launch sequence of the product's C ABI (denoise.cu: ocrb_nlm_denoise_u8).  Never shipped.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/denoise_kernels.cuh"

using namespace ocrb;

static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

extern "C" int emu_nlm_denoise(const uint8_t *src, uint8_t *dst, uint8_t *ws, int n_img, int H, int W, int C) {
  const DenoiseHostTables &t = host_tables();
  if (!t.ok) return -2;
  const dim3 grid(cdivu(W, NLM_COLS), cdivu(H, NLM_ROWS), n_img);
  if (C == 1) {
    emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<1>(src, dst, H, W, t.w[0]); });
    return 0;
  }
  const size_t npix = (size_t)n_img * H * W;
  uint8_t *ab0 = ws, *ab1 = ws + 2 * npix, *L0 = ws + 4 * npix, *L1 = ws + 5 * npix;
  const int2 *yf = reinterpret_cast<const int2 *>(t.yf);
  emu::launch(dim3(cdivu((long long)npix, 256)), dim3(256), 0, [&] { lbgr2lab_kernel(src, L0, ab0, npix, t.cbrt_tab, t.cf); });
  emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<1>(L0, L1, H, W, t.w[0]); });
  emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<2>(ab0, ab1, H, W, t.w[1]); });
  emu::launch(dim3(cdivu((long long)npix, 256)), dim3(256), 0, [&] { lab2lbgr_kernel(L1, ab1, dst, npix, yf, t.cf); });
  return 0;
}
