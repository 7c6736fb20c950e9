// uint8 image kernels of the preprocessing strategies (tools.py:503-573), bit-exact against
// OpenCV 4.13.0.92 semantics.  Compiled with -fmad=false; every
// floating-point operation whose rounding matters is additionally written with an explicit
// _rn intrinsic so that nothing is contracted or reassociated.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include "image_fast.cuh"
#include "image_general.cuh"

namespace ocrb {

static int ensure_itab() {
  static int state = 0;  // per process; tables are tiny
  static int dev_done[64] = {0};
  int dev = 0;
  OCRB_CUDA(cudaGetDevice(&dev));
  (void)state;
  if (dev < 64 && dev_done[dev]) return OCRB_OK;
  static int16_t host_tab[1024 * 16];
  build_cubic_itab(host_tab);
  OCRB_CUDA(cudaMemcpyToSymbol(g_cubic_itab, host_tab, sizeof(host_tab)));
  if (dev < 64) dev_done[dev] = 1;
  return OCRB_OK;
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_rgb2gray_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0, "rgb2gray_u8: bad arguments");
  OCRB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "rgb2gray_u8: pointers must be 16-byte aligned");
  const size_t npix = (size_t)n_img * H * W;
  const int threads = 256;
  const long long nthr = (long long)((npix + 15) / 16);
  rgb2gray_kernel<<<cdiv(nthr, threads), threads, 0, (cudaStream_t)stream>>>(src, dst, npix);
  return check_launch("rgb2gray_kernel");
}

// gray (C == 1) or RGB (C == 3) page -> CLAHE(3.0, 8x8) of its gray version.  gray_ws: uint8[n_img*H*W] for C == 3.
static int clahe_run(const uint8_t *src, int C, uint8_t *gray_ws, uint8_t *dst, int n_img, int H, int W, uint8_t *lut_ws,
                     cudaStream_t st) {
  int We = W, He = H;
  if (!(W % 8 == 0 && H % 8 == 0)) {
    We = W + (8 - W % 8);
    He = H + (8 - H % 8);
  }
  const int tw = We / 8, th = He / 8;
  OCRB_REQUIRE(tw >= 1 && th >= 1 && tw <= W && th <= H, "clahe_u8: image too small for an 8x8 tile grid");
  const int area = tw * th;
  int clip = (int)(3.0 * area / 256.0);
  if (clip < 1) clip = 1;
  volatile float lut_scale = 255.0f / (float)area;
  const uint8_t *gray = (C == 3) ? gray_ws : src;
  const int vec_ok = (W % 16 == 0) && (tw % 16 == 0) && ((uintptr_t)src & 15) == 0 && ((uintptr_t)gray & 15) == 0;
  if (C == 3)
    clahe_hist_lut_kernel<3><<<dim3(64, n_img), 256, 0, st>>>(src, gray_ws, lut_ws, H, W, tw, th, clip, lut_scale, vec_ok);
  else
    clahe_hist_lut_kernel<1><<<dim3(64, n_img), 256, 0, st>>>(src, nullptr, lut_ws, H, W, tw, th, clip, lut_scale, vec_ok);
  int rc = check_launch("clahe_hist_lut_kernel");
  if (rc) return rc;
  volatile float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  const int al4 = ((uintptr_t)gray & 3) == 0 && ((uintptr_t)dst & 3) == 0;
  // the cell boundaries depend on (H, W) only: cached for the last geometry
  static int c_H = 0, c_W = 0, c_ok = 0;
  static ClaheCells c_cells;
  if (c_H != H || c_W != W) {
    c_ok = clahe_cells_host(H, W, inv_tw, inv_th, &c_cells);
    c_H = H;
    c_W = W;
  }
  if (c_ok && al4)
    clahe_apply_cells_kernel<<<dim3(81, n_img), 256, 0, st>>>(gray, dst, lut_ws, H, W, inv_tw, inv_th, c_cells);
  else if (W % 4 == 0 && al4)
    clahe_apply4_kernel<<<dim3(cdiv(W / 4, 256), H, n_img), 256, 0, st>>>(gray, dst, lut_ws, H, W, inv_tw, inv_th);
  else
    clahe_apply_kernel<<<dim3(cdiv(W, 256), H, n_img), 256, 0, st>>>(gray, dst, lut_ws, H, W, inv_tw, inv_th);
  return check_launch("clahe_apply_kernel");
}

extern "C" int ocrb_clahe_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, uint8_t *lut_ws,
                             void *stream) {
  OCRB_REQUIRE(src && dst && lut_ws && n_img > 0 && H > 0 && W > 0, "clahe_u8: bad arguments");
  return clahe_run(src, 1, nullptr, dst, n_img, H, W, lut_ws, (cudaStream_t)stream);
}

extern "C" int ocrb_high_contrast_u8(const uint8_t *src, uint8_t *dst, uint8_t *gray_ws, uint8_t *lut_ws, int32_t n_img,
                                     int32_t H, int32_t W, int32_t C, void *stream) {
  OCRB_REQUIRE(src && dst && lut_ws && n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "high_contrast_u8: bad arguments");
  OCRB_REQUIRE(C == 1 || gray_ws, "high_contrast_u8: gray_ws is required for RGB pages");
  return clahe_run(src, C, gray_ws, dst, n_img, H, W, lut_ws, (cudaStream_t)stream);
}

static int thresh_run(const uint8_t *src, int C, uint8_t *dst, int n_img, int H, int W, cudaStream_t st) {
  const int aligned = (W % 4 == 0) && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0;
  const dim3 grid(cdiv(W, AT2_TW), cdiv(H, AT2_TH), n_img);
  // 4 CTAs per SM at 63 registers (default; 64 pages: 0.223 ms) or 5 at 48 with a few spilled words (0.232 ms): OCRB_AT_OCC=5
  static int occ = -1;
  if (occ < 0) {
    const char *e = getenv("OCRB_AT_OCC");
    occ = (e && atoi(e) == 5) ? 5 : 4;
  }
  if (occ == 5) {
    if (C == 3) adaptive_thresh_tile_kernel<3, 5><<<grid, 256, 0, st>>>(src, dst, H, W, aligned);
    else adaptive_thresh_tile_kernel<1, 5><<<grid, 256, 0, st>>>(src, dst, H, W, aligned);
  } else if (C == 3) adaptive_thresh_tile_kernel<3><<<grid, 256, 0, st>>>(src, dst, H, W, aligned);
  else adaptive_thresh_tile_kernel<1><<<grid, 256, 0, st>>>(src, dst, H, W, aligned);
  return check_launch("adaptive_thresh_tile_kernel");
}

extern "C" int ocrb_adaptive_gauss_thresh_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W,
                                             void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0, "adaptive_gauss_thresh_u8: bad arguments");
  return thresh_run(src, 1, dst, n_img, H, W, (cudaStream_t)stream);
}

extern "C" int ocrb_binarize_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "binarize_u8: bad arguments");
  return thresh_run(src, C, dst, n_img, H, W, (cudaStream_t)stream);
}

extern "C" int ocrb_sharpen3x3_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                  void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 1 && W > 1 && (C == 1 || C == 3), "sharpen3x3_u8: bad arguments");
  if ((W * C) % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
    const int nv = W * C / 16;
    const dim3 grid(cdiv(nv, 64), cdiv(H, 4), n_img);
    if (C == 3) sharpen_vec16_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, H, nv);
    else sharpen_vec16_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, H, nv);
  } else if ((W * C) % 4 == 0 && W >= 4 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0)
    sharpen4_kernel<<<dim3(cdiv((long long)W * C / 4, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, C);
  else
    sharpen_kernel<<<dim3(cdiv((long long)W * C, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, C);
  return check_launch("sharpen_kernel");
}

extern "C" int ocrb_remove_lines_mask_u8(const uint8_t *src, uint8_t *mask, int32_t *nonzero, uint8_t *tmp, int32_t n_img,
                                         int32_t H, int32_t W, int32_t C, void *stream) {
  OCRB_REQUIRE(src && mask && nonzero && tmp && n_img > 0 && H > 0 && W >= 4 && W <= 8192 && (C == 1 || C == 3),
               "remove_lines_mask_u8: bad arguments (W must be in 4..8192)");
  cudaStream_t st = (cudaStream_t)stream;
  OCRB_CUDA(cudaMemsetAsync(nonzero, 0, sizeof(int32_t) * n_img, st));
  rl_thresh_kernel<<<dim3(cdiv(W, 256), H, n_img), 256, 0, st>>>(src, mask, H, W, C);          // th -> mask buffer
  int rc = check_launch("rl_thresh_kernel");
  if (rc) return rc;
  rl_open_row_kernel<<<n_img * H, 256, (2 * W + 1) * sizeof(int), st>>>(mask, tmp, W, W / 4);        // open -> tmp
  rc = check_launch("rl_open_row_kernel");
  if (rc) return rc;
  rl_dilate_v_kernel<<<dim3(cdiv(W, 256), H, n_img), 256, 0, st>>>(tmp, mask, nonzero, H, W);     // mask
  return check_launch("rl_dilate_v_kernel");
}

extern "C" int ocrb_deskew_angle(const uint8_t *src, int32_t n_img, int32_t H, int32_t W, int32_t C, double *out_angle,
                                 double *out_M, int32_t *ext_ws, int32_t *hull_ws, void *stream) {
  OCRB_REQUIRE(src && out_angle && out_M && ext_ws && hull_ws, "deskew_angle: null pointer");
  OCRB_REQUIRE(n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "deskew_angle: bad sizes");
  const int rows = n_img * H;
  if (W % 16 == 0 && ((uintptr_t)src & 15) == 0)
    dark_extents16_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(src, ext_ws, W, C, rows);
  else
    dark_extents_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(src, ext_ws, H, W, C, rows);
  int rc = check_launch("dark_extents_kernel");
  if (rc) return rc;
  const size_t smem = deskew_par_smem_bytes(H);
  if (smem <= 200 * 1024) {
    static size_t attr = 48 * 1024;
    if (smem > attr) {
      OCRB_CUDA(cudaFuncSetAttribute(deskew_angle_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    deskew_angle_par_kernel<<<n_img, 256, smem, (cudaStream_t)stream>>>(ext_ws, H, W, out_angle, out_M);
  } else {
    deskew_angle_seq_kernel<<<n_img, 32, 0, (cudaStream_t)stream>>>(ext_ws, H, W, out_angle, out_M, hull_ws);
  }
  return check_launch("deskew_angle_kernel");
}

extern "C" int ocrb_warp_affine_cubic_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W,
                                         int32_t C, const double *M, void *stream) {
  OCRB_REQUIRE(src && dst && M && n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "warp_affine_cubic_u8: bad arguments");
  OCRB_REQUIRE(src != dst, "warp_affine_cubic_u8: in-place not supported");
  OCRB_REQUIRE((long long)H * W * C < (1ll << 31) - 8, "warp_affine_cubic_u8: image too large (H * W * C must be below 2^31)");
  int rc = ensure_itab();
  if (rc) return rc;
  const dim3 grid(cdiv(W, 256), H, n_img);
  // 5 CTAs per SM at 48 registers (default; 64 pages: 0.66 ms for the whole deskew against 0.72 ms) or 4 at 64: OCRB_WARP_OCC=4
  static int occ = -1;
  if (occ < 0) {
    const char *e = getenv("OCRB_WARP_OCC");
    occ = (e && atoi(e) == 4) ? 4 : 5;
  }
  if (occ == 5 && C == 3) warp_affine_cubic_dp2a_kernel<3, 5><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, M);
  else if (C == 3) warp_affine_cubic_dp2a_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, M);
  else warp_affine_cubic_dp2a_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, M);
  return check_launch("warp_affine_cubic_kernel");
}
