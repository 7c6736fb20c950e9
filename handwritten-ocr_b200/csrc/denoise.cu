// denoise strategy (tools.py:576-589): cv2.fastNlMeansDenoising(gray, None, 10, 7, 21) and
// cv2.fastNlMeansDenoisingColored(rgb, None, 10, 10, 7, 21), bit-exact against OpenCV 4.13.0.92.
// Everything is integer arithmetic (the weight table, the Lab cube-root table and the Lab coefficients are built
// once on the host with the float / double steps OpenCV uses and uploaded).
//
//   nlm_kernel<CN>:   one CTA of 4 warps per 26 x 32 output tile, each warp taking a quarter of the 441 search
//                     displacements (integer partial sums meet in shared memory: order-free, so still exact).
//                     Lane l owns column l - 3 of the tile: for each displacement it walks 38 rows, keeps the running 7-row column sum of squared
//                     differences in a register (ring of the last 7 values), gets the 7-column patch sum with three
//                     warp shuffles, looks the weight up in a shared-memory table and accumulates weight and
//                     weight * pixel in registers (32 rows x (1 + CN) accumulators).  The reflect-101 extended tile
//                     lives in shared memory; the lane's own column of it is packed into registers once.
//   lbgr2lab_kernel / lab2lbgr_kernel: the two 8-bit Lab conversions of the colored variant (the reference hands
//                     an RGB array to a function that reads channel 0 as blue; restated as called).
#include "common.cuh"
#include <math.h>
#include <string.h>
#include "denoise_kernels.cuh"

namespace ocrb {

struct DenoiseDevTables {
  uint16_t *cbrt_tab;
  int2 *yf;
  uint16_t *w[2];
};

static int device_tables(const DenoiseDevTables **out) {
  static DenoiseDevTables dev[64];
  static bool done[64] = {false};
  int d = 0;
  OCRB_CUDA(cudaGetDevice(&d));
  OCRB_REQUIRE(d >= 0 && d < 64, "denoise: device index out of range");
  if (!done[d]) {
    const DenoiseHostTables &h = host_tables();
    OCRB_REQUIRE(h.ok, "denoise: weight table does not end inside the shared-memory table");
    uint8_t *base = nullptr;
    const size_t sz = sizeof(h.cbrt_tab) + sizeof(h.yf) + sizeof(h.w);
    OCRB_CUDA(cudaMalloc(&base, sz));
    OCRB_CUDA(cudaMemcpy(base, h.yf, sizeof(h.yf), cudaMemcpyHostToDevice));
    OCRB_CUDA(cudaMemcpy(base + sizeof(h.yf), h.cbrt_tab, sizeof(h.cbrt_tab), cudaMemcpyHostToDevice));
    OCRB_CUDA(cudaMemcpy(base + sizeof(h.yf) + sizeof(h.cbrt_tab), h.w, sizeof(h.w), cudaMemcpyHostToDevice));
    dev[d].yf = reinterpret_cast<int2 *>(base);
    dev[d].cbrt_tab = reinterpret_cast<uint16_t *>(base + sizeof(h.yf));
    dev[d].w[0] = reinterpret_cast<uint16_t *>(base + sizeof(h.yf) + sizeof(h.cbrt_tab));
    dev[d].w[1] = dev[d].w[0] + NLM_LUT;
    done[d] = true;
  }
  *out = &dev[d];
  return OCRB_OK;
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_denoise_tables_host(int32_t *cbrt_tab, int32_t *lab_yf, int32_t *coef, int32_t *w1, int32_t *w2) {
  OCRB_REQUIRE(cbrt_tab && lab_yf && coef && w1 && w2, "denoise_tables_host: null pointer");
  const DenoiseHostTables &h = host_tables();
  for (int i = 0; i < 3072; ++i) cbrt_tab[i] = h.cbrt_tab[i];
  for (int i = 0; i < 512; ++i) lab_yf[i] = h.yf[i];
  for (int i = 0; i < 9; ++i) {
    coef[i] = h.cf.fwd[i];
    coef[9 + i] = h.cf.inv[i];
  }
  for (int i = 0; i < NLM_LUT; ++i) {
    w1[i] = h.w[0][i];
    w2[i] = h.w[1][i];
  }
  return h.ok ? OCRB_OK : OCRB_EINVAL;
}

extern "C" int ocrb_nlm_denoise_u8(const uint8_t *src, uint8_t *dst, uint8_t *ws, int32_t n_img, int32_t H, int32_t W,
                                   int32_t C, void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "nlm_denoise_u8: bad arguments");
  OCRB_REQUIRE(src != dst, "nlm_denoise_u8: in-place not supported");
  OCRB_REQUIRE(C == 1 || ws, "nlm_denoise_u8: RGB pages need the 6-bytes-per-pixel workspace");
  OCRB_REQUIRE(cdiv(H, NLM_ROWS) <= 65535 && n_img <= 65535, "nlm_denoise_u8: page too tall / batch too large");
  const DenoiseDevTables *t = nullptr;
  int rc = device_tables(&t);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(cdiv(W, NLM_COLS), cdiv(H, NLM_ROWS), n_img);
  if (C == 1) {
    nlm_kernel<1><<<grid, 32 * NLM_NW, 0, st>>>(src, dst, H, W, t->w[0]);
    return check_launch("nlm_kernel<1>");
  }
  const size_t npix = (size_t)n_img * H * W;
  uint8_t *ab0 = ws, *ab1 = ws + 2 * npix, *L0 = ws + 4 * npix, *L1 = ws + 5 * npix;   // (a, b) pairs first: 2-byte aligned
  OCRB_REQUIRE(((uintptr_t)ws & 1) == 0, "nlm_denoise_u8: workspace must be 2-byte aligned");
  lbgr2lab_kernel<<<cdiv((long long)npix, 256), 256, 0, st>>>(src, L0, ab0, npix, t->cbrt_tab, host_tables().cf);
  if ((rc = check_launch("lbgr2lab_kernel"))) return rc;
  nlm_kernel<1><<<grid, 32 * NLM_NW, 0, st>>>(L0, L1, H, W, t->w[0]);
  if ((rc = check_launch("nlm_kernel<1>"))) return rc;
  nlm_kernel<2><<<grid, 32 * NLM_NW, 0, st>>>(ab0, ab1, H, W, t->w[1]);
  if ((rc = check_launch("nlm_kernel<2>"))) return rc;
  lab2lbgr_kernel<<<cdiv((long long)npix, 256), 256, 0, st>>>(L1, ab1, dst, npix, t->yf, host_tables().cf);
  return check_launch("lab2lbgr_kernel");
}
