// Kernels and host tables of the denoise strategy (tools.py:576-589): cv2.fastNlMeansDenoising(gray, None, 10, 7, 21) and
// cv2.fastNlMeansDenoisingColored(rgb, None, 10, 10, 7, 21), bit-exact against OpenCV 4.13.0.92 (see denoise.cu for the
// design).  Definitions only -- no launches, no CUDA runtime calls -- so that tests/emu can compile this file for the host
// (OCRB_EMU) and run the kernels thread by thread against the oracle and under the host sanitizers.
#pragma once
#ifndef OCRB_EMU
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace ocrb {

constexpr int NLM_SH = 10;             // search half width (21 x 21)
constexpr int NLM_B = 13;              // border = search half + template half
constexpr int NLM_LUT = 2048;          // weights are zero from an index below this (host-checked)
constexpr int NLM_ROWS = 32;           // output rows per warp tile
constexpr int NLM_COLS = 26;           // output columns per warp: 32 lanes - 6 template halo lanes
constexpr int NLM_DROWS = NLM_ROWS + 6;
constexpr int NLM_ER = NLM_ROWS + 2 * NLM_B;
constexpr int NLM_EC = 32 + 2 * NLM_SH;  // 52 columns of the extended tile
constexpr int NLM_SHIFT = 6;           // 2^6 >= 7 * 7
constexpr int NLM_FPM = 2147483647 / (21 * 21 * 255);

__device__ __forceinline__ int reflect101_any(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

template <int CN> struct NlmPix;
template <> struct NlmPix<1> { typedef uint8_t T; };
template <> struct NlmPix<2> { typedef uint16_t T; };

constexpr int NLM_NW = 4;              // warps per tile: each takes a quarter of the 441 displacements

template <int CN>
__global__ void __launch_bounds__(32 * NLM_NW)
nlm_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, const uint16_t *__restrict__ lut_g) {
  typedef typename NlmPix<CN>::T T;
  constexpr int PER = 4 / CN;                         // pixels of the lane's own column per 32-bit register
  constexpr int NPACK = (NLM_DROWS + PER - 1) / PER;
  constexpr int NSEARCH = (2 * NLM_SH + 1) * (2 * NLM_SH + 1);
  __shared__ T ext[NLM_ER * NLM_EC];
  __shared__ uint16_t lut[NLM_LUT];
  __shared__ uint32_t acc[(1 + CN) * NLM_ROWS * 32];  // the warps' partial sums meet here (integer: order-free)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx0 = blockIdx.x * NLM_COLS, ty0 = blockIdx.y * NLM_ROWS;
  const T *im = reinterpret_cast<const T *>(src) + (size_t)blockIdx.z * H * W;
  T *om = reinterpret_cast<T *>(dst) + (size_t)blockIdx.z * H * W;
  for (int i = threadIdx.x; i < NLM_LUT; i += 32 * NLM_NW) lut[i] = lut_g[i];
  for (int i = threadIdx.x; i < (1 + CN) * NLM_ROWS * 32; i += 32 * NLM_NW) acc[i] = 0;
  for (int er = warp; er < NLM_ER; er += NLM_NW) {
    const int gy = reflect101_any(ty0 - NLM_B + er, H);
    for (int ec = lane; ec < NLM_EC; ec += 32)
      ext[er * NLM_EC + ec] = im[(size_t)gy * W + reflect101_any(tx0 - NLM_B + ec, W)];
  }
  __syncthreads();

  uint32_t apack[NPACK];
#pragma unroll
  for (int k = 0; k < NPACK; ++k) {
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int r = k * PER + q;
      if (r < NLM_DROWS) v |= (uint32_t)ext[(r + NLM_SH) * NLM_EC + lane + NLM_SH] << (q * 8 * CN);
    }
    apack[k] = v;
  }

  uint32_t wsum[NLM_ROWS], est0[NLM_ROWS], est1[CN == 2 ? NLM_ROWS : 1];
#pragma unroll
  for (int j = 0; j < NLM_ROWS; ++j) {
    wsum[j] = 0;
    est0[j] = 0;
    if (CN == 2) est1[j] = 0;
  }

  const int d_end = (warp + 1) * NSEARCH / NLM_NW;
#pragma unroll 1
  for (int d = warp * NSEARCH / NLM_NW; d < d_end; ++d) {
    const int dy = d / (2 * NLM_SH + 1) - NLM_SH, dx = d % (2 * NLM_SH + 1) - NLM_SH;
    const T *bp = ext + (NLM_SH + dy) * NLM_EC + lane + NLM_SH + dx;
    uint32_t col = 0, dring[7], bring[4];
#pragma unroll
    for (int r = 0; r < NLM_DROWS; ++r) {
      const uint32_t a = (apack[r / PER] >> ((r % PER) * 8 * CN)) & (CN == 1 ? 0xffu : 0xffffu);
      const uint32_t b = bp[r * NLM_EC];
      const uint32_t ad = __vabsdiffu4(a, b);
      const uint32_t dd = __dp4a(ad, ad, 0u);
      col += dd;
      if (r >= 7) col -= dring[r % 7];
      dring[r % 7] = dd;
      bring[r % 4] = b;
      if (r >= 6) {
        const int j = r - 6;
        const uint32_t t2 = col + __shfl_down_sync(0xffffffffu, col, 1);
        const uint32_t t4 = t2 + __shfl_down_sync(0xffffffffu, t2, 2);          // columns l .. l+3
        const uint32_t s7 = __shfl_up_sync(0xffffffffu, t4, 3) + t4 - col;      // columns l-3 .. l+3
        const uint32_t w = lut[min(s7 >> NLM_SHIFT, (uint32_t)(NLM_LUT - 1))];
        const uint32_t bc = bring[(r - 3) % 4];
        wsum[j] += w;
        if (CN == 1) {
          est0[j] += w * bc;
        } else {
          est0[j] += w * (bc & 0xffu);
          est1[j] += w * (bc >> 8);
        }
      }
    }
  }

#pragma unroll
  for (int j = 0; j < NLM_ROWS; ++j) {
    atomicAdd(&acc[j * 32 + lane], wsum[j]);
    atomicAdd(&acc[(NLM_ROWS + j) * 32 + lane], est0[j]);
    if (CN == 2) atomicAdd(&acc[(2 * NLM_ROWS + j) * 32 + lane], est1[j]);
  }
  __syncthreads();
  const int x = tx0 + lane - 3;
  if (lane < 3 || lane >= 3 + NLM_COLS || x >= W) return;
  for (int j = warp; j < NLM_ROWS; j += NLM_NW) {
    const int y = ty0 + j;
    if (y < H) {
      const uint32_t ws = acc[j * 32 + lane];
      uint32_t o = min((acc[(NLM_ROWS + j) * 32 + lane] + ws / 2) / ws, 255u);
      if (CN == 2) o |= min((acc[(2 * NLM_ROWS + j) * 32 + lane] + ws / 2) / ws, 255u) << 8;
      om[(size_t)y * W + x] = (T)o;
    }
  }
}

// ───────────── 8-bit Lab conversions (cvtColor COLOR_LBGR2Lab / COLOR_Lab2LBGR) ─────────────
struct LabCoef {
  int fwd[9];   // RGB2Lab_b coefficients, 12 bits, columns already swapped for "channel 0 is blue"
  int inv[9];   // Lab2RGBinteger coefficients, 12 bits
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__global__ void __launch_bounds__(256)
lbgr2lab_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ Lp, uint8_t *__restrict__ abp, size_t npix,
                const uint16_t *__restrict__ cbrt_tab, LabCoef cf) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int c0 = src[3 * p] * 8, c1 = src[3 * p + 1] * 8, c2 = src[3 * p + 2] * 8;
  const int fX = __ldg(cbrt_tab + descale(c0 * cf.fwd[0] + c1 * cf.fwd[1] + c2 * cf.fwd[2], 12));
  const int fY = __ldg(cbrt_tab + descale(c0 * cf.fwd[3] + c1 * cf.fwd[4] + c2 * cf.fwd[5], 12));
  const int fZ = __ldg(cbrt_tab + descale(c0 * cf.fwd[6] + c1 * cf.fwd[7] + c2 * cf.fwd[8], 12));
  const int L = descale(296 * fY - 1336935, 15);              // (116*255+50)/100 and ((16*255*2^15+50)/100)
  const int a = descale(500 * (fX - fY) + 128 * 32768, 15);
  const int b = descale(200 * (fY - fZ) + 128 * 32768, 15);
  Lp[p] = (uint8_t)min(max(L, 0), 255);
  reinterpret_cast<uint16_t *>(abp)[p] = (uint16_t)(min(max(a, 0), 255) | (min(max(b, 0), 255) << 8));
}

__device__ __forceinline__ int ab_to_xz(int i) {
  constexpr int B = 1 << 14;
  if (i <= 3390) return i * 108 / 841 - (B * 16 / 116 * 108 / 841);       // C division truncates toward zero
  return (int)((long long)(i * i / B) * i / B);
}

__global__ void __launch_bounds__(256)
lab2lbgr_kernel(const uint8_t *__restrict__ Lp, const uint8_t *__restrict__ abp, uint8_t *__restrict__ dst, size_t npix,
                const int2 *__restrict__ yf_tab, LabCoef cf) {
  constexpr int B = 1 << 14;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int L = Lp[p];
  const uint32_t ab = reinterpret_cast<const uint16_t *>(abp)[p];
  const int aa = ab & 0xff, bb = ab >> 8;
  const int2 yf = __ldg(yf_tab + L);
  const int y = yf.x, ify = yf.y;
  const int adiv = ((5 * aa * 53687 + 128) >> 13) - 128 * B / 500;
  const int bdiv = ((bb * 41943 + 16) >> 9) - 128 * B / 200 + 1;
  const int x = ab_to_xz(ify + adiv);
  const int z = ab_to_xz(ify - bdiv);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int v = descale(cf.inv[3 * k] * x + cf.inv[3 * k + 1] * y + cf.inv[3 * k + 2] * z, 14);
    v = min(max(v, 0), 4095);
    dst[3 * p + k] = (uint8_t)((v * 255) >> 12);
  }
}

// ───────────── host tables (built once; float / double steps as OpenCV's softfloat code takes them) ─────────────
static float cv_cbrt(float x) {   // cv::cubeRoot: exponent in thirds, quartic rational on the fraction, mantissa truncated
  if (x == 0.f) return 0.f;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  ix &= 0x7fffffffu;
  const int ex = (int)(ix >> 23) - 127;
  int shx = ex % 3;
  shx -= shx >= 0 ? 3 : 0;
  const int ex3 = (ex - shx) / 3;
  const uint32_t fbits = (ix & ((1u << 23) - 1)) | ((uint32_t)(shx + 127) << 23);
  float frf;
  memcpy(&frf, &fbits, 4);
  const double fr = frf;
  const double num = ((((45.2548339756803022511987494 * fr + 192.2798368355061050458134625) * fr +
                        119.1654824285581628956914143) * fr + 13.43250139086239872172837314) * fr +
                      0.1636161226585754240958355063);
  const double den = ((((14.80884093219134573786480845 * fr + 151.9714051044435648658557668) * fr +
                        168.5254414101568283957668343) * fr + 33.9905941350215598754191872) * fr + 1.0);
  const double q = num / den;
  uint64_t qb;
  memcpy(&qb, &q, 8);
  // double -> float by truncating the mantissa to 23 bits, then add the exponent third
  const uint32_t m = (uint32_t)((qb >> 29) & ((1u << 23) - 1));
  const int e = (int)((qb >> 52) & 0x7ff) - 1023 + 127 + ex3;
  const uint32_t ob = ((uint32_t)e << 23) | m;
  float out;
  memcpy(&out, &ob, 4);
  return out;
}

struct DenoiseHostTables {
  uint16_t cbrt_tab[3072];
  int yf[512];
  LabCoef cf;
  uint16_t w[2][NLM_LUT];
  bool ok;
};

static const DenoiseHostTables &host_tables() {
  static DenoiseHostTables t;
  static bool built = false;
  if (built) return t;
  t.ok = true;
  const float thr = 216.0f / 24389.0f, sc = 841.0f / 108.0f, off = 16.0f / 116.0f;
  for (int i = 0; i < 3072; ++i) {
    const float x = (float)i / 2040.0f;
    const float f = x < thr ? fmaf(x, sc, off) : cv_cbrt(x);
    t.cbrt_tab[i] = (uint16_t)lrint((double)(32768.0f * f));
  }
  const int B = 1 << 14;
  for (int i = 0; i < 256; ++i) {
    float y, ify;
    if (i <= 20) {
      y = (float)(i * B * 20 * 9) / (float)(17 * 29 * 29 * 29);
      ify = (float)B * (16.0f / 116.0f + (float)(i * 100) / (float)(255 * 116));
    } else {
      ify = (float)(i * 100 * B) / (float)(255 * 116) + (float)(16 * B) / 116.0f;
      y = ify * ify * ify / (float)(B * B);
    }
    t.yf[2 * i] = (int)lrint((double)y);
    t.yf[2 * i + 1] = (int)lrint((double)ify);
  }
  static const double wp[3] = {0.950456, 1.0, 1.088754};
  static const double r2x[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
  static const double x2r[9] = {3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556, 0.055648, -0.204043, 1.057311};
  for (int i = 0; i < 3; ++i) {
    t.cf.fwd[i * 3 + 2] = (int)lrint(4096 * r2x[i * 3] / wp[i]);
    t.cf.fwd[i * 3 + 1] = (int)lrint(4096 * r2x[i * 3 + 1] / wp[i]);
    t.cf.fwd[i * 3 + 0] = (int)lrint(4096 * r2x[i * 3 + 2] / wp[i]);
    t.cf.inv[i] = (int)lrint(4096 * x2r[i + 6] * wp[i]);
    t.cf.inv[i + 3] = (int)lrint(4096 * x2r[i + 3] * wp[i]);
    t.cf.inv[i + 6] = (int)lrint(4096 * x2r[i] * wp[i]);
  }
  const double mult = 64.0 / 49.0;
  for (int cn = 1; cn <= 2; ++cn) {
    const int n = (int)(255.0 * 255.0 * cn / mult + 1);
    for (int d = 0; d < n; ++d) {
      long wgt = lrint(NLM_FPM * exp(-(d * mult) / (10.0 * 10.0 * cn)));
      if ((double)wgt < 0.001 * NLM_FPM) wgt = 0;
      if (d < NLM_LUT) t.w[cn - 1][d] = (uint16_t)wgt;
      else if (wgt != 0) t.ok = false;
    }
    if (t.w[cn - 1][NLM_LUT - 1] != 0) t.ok = false;
  }
  built = true;
  return t;
}

}  // namespace ocrb
