// TEST INFRASTRUCTURE: runs the kernels of handwritten-ocr_b200/csrc/image_fast.cuh on the CPU through tests/emu/cuda_emu.h,
// with the launch geometry the product's C ABI uses (image.cu: clahe_run, thresh_run, ocrb_sharpen3x3_u8,
// ocrb_deskew_angle, ocrb_warp_affine_cubic_u8).  Built by tests/test_emu_image_kernels.py with g++; never shipped.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/image_fast.cuh"
#include "../../handwritten-ocr_b200/csrc/image_general.cuh"

using namespace ocrb;

static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

extern "C" int emu_sharpen(const uint8_t *src, uint8_t *dst, int n, int H, int W, int C) {
  if ((W * C) % 16 != 0 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15) || H < 2) return -1;
  const int nv = W * C / 16;
  const dim3 grid(cdivu(nv, 64), cdivu(H, 4), n);
  if (C == 3) emu::launch(grid, dim3(256), 0, [&] { sharpen_vec16_kernel<3>(src, dst, H, nv); });
  else emu::launch(grid, dim3(256), 0, [&] { sharpen_vec16_kernel<1>(src, dst, H, nv); });
  return 0;
}

// returns 0 (cells kernel ran), 1 (hist kernel ran, cells kernel not applicable: dst untouched, lut filled), -1 bad size
extern "C" int emu_high_contrast(const uint8_t *src, uint8_t *dst, uint8_t *gray_ws, uint8_t *lut_ws, int n, int H, int W,
                                 int C, int *used_vec) {
  int We = W, He = H;
  if (!(W % 8 == 0 && H % 8 == 0)) {
    We = W + (8 - W % 8);
    He = H + (8 - H % 8);
  }
  const int tw = We / 8, th = He / 8;
  if (!(tw >= 1 && th >= 1 && tw <= W && th <= H)) return -1;
  const int area = tw * th;
  int clip = (int)(3.0 * area / 256.0);
  if (clip < 1) clip = 1;
  volatile float lut_scale = 255.0f / (float)area;
  const uint8_t *gray = (C == 3) ? gray_ws : src;
  const int vec_ok = (W % 16 == 0) && (tw % 16 == 0) && ((uintptr_t)src & 15) == 0 && ((uintptr_t)gray & 15) == 0;
  *used_vec = vec_ok;
  const float ls = lut_scale;
  if (C == 3)
    emu::launch(dim3(64, n), dim3(256), 0, [&] { clahe_hist_lut_kernel<3>(src, gray_ws, lut_ws, H, W, tw, th, clip, ls, vec_ok); });
  else
    emu::launch(dim3(64, n), dim3(256), 0, [&] { clahe_hist_lut_kernel<1>(src, nullptr, lut_ws, H, W, tw, th, clip, ls, vec_ok); });
  volatile float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  ClaheCells cells;
  const int ok = clahe_cells_host(H, W, inv_tw, inv_th, &cells);
  if (!ok || ((uintptr_t)gray & 3) || ((uintptr_t)dst & 3)) return 1;
  const float itw = inv_tw, ith = inv_th;
  emu::launch(dim3(81, n), dim3(256), 0, [&] { clahe_apply_cells_kernel(gray, dst, lut_ws, H, W, itw, ith, cells); });
  return 0;
}

extern "C" int emu_binarize(const uint8_t *src, uint8_t *dst, int n, int H, int W, int C) {
  const int aligned = (W % 4 == 0) && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0;
  const dim3 grid(cdivu(W, AT2_TW), cdivu(H, AT2_TH), n);
  if (C == 3) emu::launch(grid, dim3(256), 0, [&] { adaptive_thresh_tile_kernel<3>(src, dst, H, W, aligned); });
  else emu::launch(grid, dim3(256), 0, [&] { adaptive_thresh_tile_kernel<1>(src, dst, H, W, aligned); });
  return 0;
}

extern "C" int emu_deskew_angle(const uint8_t *src, int n, int H, int W, int C, double *out_angle, double *out_M,
                                int32_t *ext_ws) {
  if (W % 16 != 0 || ((uintptr_t)src & 15)) return -1;
  const int rows = n * H;
  emu::launch(dim3(cdivu(rows, 8)), dim3(256), 0, [&] { dark_extents16_kernel(src, ext_ws, W, C, rows); });
  const size_t smem = deskew_par_smem_bytes(H);
  emu::launch(dim3(n), dim3(256), smem, [&] { deskew_angle_par_kernel(ext_ws, H, W, out_angle, out_M); });
  return 0;
}

extern "C" int emu_warp(const uint8_t *src, uint8_t *dst, int n, int H, int W, int C, const double *M) {
  static bool built = false;
  if (!built) {
    build_cubic_itab(g_cubic_itab);
    built = true;
  }
  const dim3 grid(cdivu(W, 256), H, n);
  if (C == 3) emu::launch(grid, dim3(256), 0, [&] { warp_affine_cubic_dp2a_kernel<3>(src, dst, H, W, M); });
  else emu::launch(grid, dim3(256), 0, [&] { warp_affine_cubic_dp2a_kernel<1>(src, dst, H, W, M); });
  return 0;
}

// the host-side CLAHE cell boundaries alone (image.cu: clahe_run), for property tests over many page sizes
extern "C" int emu_clahe_cells(int H, int W, int *xb, int *yb, float *inv_tw_out, float *inv_th_out) {
  int We = W, He = H;
  if (!(W % 8 == 0 && H % 8 == 0)) {
    We = W + (8 - W % 8);
    He = H + (8 - H % 8);
  }
  const int tw = We / 8, th = He / 8;
  volatile float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  ClaheCells cells;
  const int ok = clahe_cells_host(H, W, inv_tw, inv_th, &cells);
  for (int k = 0; k < 10; ++k) {
    xb[k] = cells.xb[k];
    yb[k] = cells.yb[k];
  }
  *inv_tw_out = inv_tw;
  *inv_th_out = inv_th;
  return ok;
}

// ───────────── the general-shape kernels (image_general.cuh), launch geometry of image.cu ─────────────
extern "C" int emu_rgb2gray(const uint8_t *src, uint8_t *dst, int n, int H, int W) {
  const size_t npix = (size_t)n * H * W;
  emu::launch(dim3(cdivu((long long)((npix + 15) / 16), 256)), dim3(256), 0, [&] { rgb2gray_kernel(src, dst, npix); });
  return 0;
}

// variant 0: sharpen_kernel (bytes), 1: sharpen4_kernel (W * C % 4 == 0)
extern "C" int emu_sharpen_general(const uint8_t *src, uint8_t *dst, int n, int H, int W, int C, int variant) {
  if (variant == 1) {
    if ((W * C) % 4 != 0 || W < 4 || ((uintptr_t)src & 3) || ((uintptr_t)dst & 3)) return -1;
    emu::launch(dim3(cdivu((long long)W * C / 4, 256), H, n), dim3(256), 0, [&] { sharpen4_kernel(src, dst, H, W, C); });
  } else {
    emu::launch(dim3(cdivu((long long)W * C, 256), H, n), dim3(256), 0, [&] { sharpen_kernel(src, dst, H, W, C); });
  }
  return 0;
}

// CLAHE of a gray page through the general apply kernels (variant 0: per pixel, 1: four pixels per thread)
extern "C" int emu_clahe_general(const uint8_t *src, uint8_t *dst, uint8_t *lut_ws, int n, int H, int W, int variant) {
  int We = W, He = H;
  if (!(W % 8 == 0 && H % 8 == 0)) {
    We = W + (8 - W % 8);
    He = H + (8 - H % 8);
  }
  const int tw = We / 8, th = He / 8;
  if (!(tw >= 1 && th >= 1 && tw <= W && th <= H)) return -1;
  const int area = tw * th;
  int clip = (int)(3.0 * area / 256.0);
  if (clip < 1) clip = 1;
  volatile float lut_scale = 255.0f / (float)area;
  const float ls = lut_scale;
  emu::launch(dim3(64, n), dim3(256), 0, [&] { clahe_hist_lut_kernel<1>(src, nullptr, lut_ws, H, W, tw, th, clip, ls, 0); });
  volatile float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  const float itw = inv_tw, ith = inv_th;
  if (variant == 1) {
    if (W % 4 != 0 || ((uintptr_t)src & 3) || ((uintptr_t)dst & 3)) return -1;
    emu::launch(dim3(cdivu(W / 4, 256), H, n), dim3(256), 0, [&] { clahe_apply4_kernel(src, dst, lut_ws, H, W, itw, ith); });
  } else {
    emu::launch(dim3(cdivu(W, 256), H, n), dim3(256), 0, [&] { clahe_apply_kernel(src, dst, lut_ws, H, W, itw, ith); });
  }
  return 0;
}

// deskew angle through the general extent kernel and the one-thread hull scan (the path of very tall pages)
extern "C" int emu_deskew_angle_general(const uint8_t *src, int n, int H, int W, int C, double *out_angle, double *out_M,
                                        int32_t *ext_ws, int32_t *hull_ws) {
  const int rows = n * H;
  emu::launch(dim3(cdivu(rows, 8)), dim3(256), 0, [&] { dark_extents_kernel(src, ext_ws, H, W, C, rows); });
  emu::launch(dim3(n), dim3(32), 0, [&] { deskew_angle_seq_kernel(ext_ws, H, W, out_angle, out_M, hull_ws); });
  return 0;
}

// ruled-line mask (ocrb_remove_lines_mask_u8)
extern "C" int emu_remove_lines_mask(const uint8_t *src, uint8_t *mask, int32_t *nonzero, uint8_t *tmp, int n, int H, int W,
                                     int C) {
  if (W < 4 || W > 8192) return -1;
  for (int i = 0; i < n; ++i) nonzero[i] = 0;
  emu::launch(dim3(cdivu(W, 256), H, n), dim3(256), 0, [&] { rl_thresh_kernel(src, mask, H, W, C); });
  emu::launch(dim3(n * H), dim3(256), (2 * (size_t)W + 1) * sizeof(int), [&] { rl_open_row_kernel(mask, tmp, W, W / 4); });
  emu::launch(dim3(cdivu(W, 256), H, n), dim3(256), 0, [&] { rl_dilate_v_kernel(tmp, mask, nonzero, H, W); });
  return 0;
}
