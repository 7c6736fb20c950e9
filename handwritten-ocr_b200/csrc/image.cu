// uint8 image kernels of the preprocessing strategies (tools.py:503-573), bit-exact against
// OpenCV 4.13.0.92 semantics.  Compiled with -fmad=false; every
// floating-point operation whose rounding matters is additionally written with an explicit
// _rn intrinsic so that nothing is contracted or reassociated.
#include "common.cuh"
#include <math.h>

namespace ocrb {

// ───────────────────────── A.1 RGB -> gray ─────────────────────────
__device__ __forceinline__ uint32_t gray_px(uint32_t r, uint32_t g, uint32_t b) {
  return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15;
}

// 16 pixels per thread: three 16-byte loads, one 16-byte store.
__global__ void __launch_bounds__(256)
rgb2gray_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t npix) {
  const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (base >= npix) return;
  if (base + 16 <= npix) {
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + base * 3);
    union { uint4 v[3]; uint8_t b[48]; } in;
    in.v[0] = __ldg(s4);
    in.v[1] = __ldg(s4 + 1);
    in.v[2] = __ldg(s4 + 2);
    union { uint4 v; uint8_t b[16]; } o;
#pragma unroll
    for (int k = 0; k < 16; ++k) o.b[k] = (uint8_t)gray_px(in.b[3 * k], in.b[3 * k + 1], in.b[3 * k + 2]);
    *reinterpret_cast<uint4 *>(dst + base) = o.v;
  } else {
    for (size_t p = base; p < npix; ++p) dst[p] = (uint8_t)gray_px(src[3 * p], src[3 * p + 1], src[3 * p + 2]);
  }
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ───────────────────────── A.2 CLAHE ─────────────────────────
// Pass 1: one CTA per (tile, image): 256-bin shared histogram -> clip -> redistribute -> LUT.
__global__ void __launch_bounds__(256)
clahe_lut_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ lut, int H, int W, int tw, int th,
                 int clip, float lut_scale) {
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int tile = blockIdx.x, img = blockIdx.y;
  const int ty = tile >> 3, tx = tile & 7;
  const uint8_t *im = src + (size_t)img * H * W;
  hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_clipped = 0;
  __syncthreads();
  const int x0 = tx * tw, y0 = ty * th;
  for (int p = threadIdx.x; p < tw * th; p += 256) {
    const int yy = reflect101(y0 + p / tw, H);
    const int xx = reflect101(x0 + p % tw, W);
    atomicAdd(&hist[im[(size_t)yy * W + xx]], 1);
  }
  __syncthreads();
  int h = hist[threadIdx.x];
  const int excess = max(h - clip, 0);
  // block sum of the excess
  int e = excess;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_clipped, e);
  __syncthreads();
  const int clipped = s_clipped;
  h = min(h, clip);
  const int batch = clipped / 256;
  const int resid = clipped - batch * 256;
  h += batch;
  if (resid) {
    const int step = max(256 / resid, 1);
    const int i = threadIdx.x;
    if (i % step == 0 && i / step < resid) h += 1;
  }
  // inclusive scan over 256 bins (Hillis-Steele in shared memory)
  scan[threadIdx.x] = h;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int v = scan[threadIdx.x];
    if ((int)threadIdx.x >= o) v += scan[threadIdx.x - o];
    __syncthreads();
    scan[threadIdx.x] = v;
    __syncthreads();
  }
  const float f = __fmul_rn((float)scan[threadIdx.x], lut_scale);
  int q = __float2int_rn(f);
  q = min(max(q, 0), 255);
  lut[((size_t)img * 64 + tile) * 256 + threadIdx.x] = (uint8_t)q;
}

// Pass 2: bilinear blend of the four neighbouring tile LUTs; unfused fp32 mul/add in OpenCV's order.
__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const uint8_t *__restrict__ lut,
                   int H, int W, float inv_tw, float inv_th) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (x >= W) return;
  const float xf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
  const float yf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int tx1 = (int)floorf(xf), ty1 = (int)floorf(yf);
  int tx2 = tx1 + 1, ty2 = ty1 + 1;
  const float xa = __fsub_rn(xf, (float)tx1), ya = __fsub_rn(yf, (float)ty1);
  const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
  tx1 = max(tx1, 0);
  tx2 = min(tx2, 7);
  ty1 = max(ty1, 0);
  ty2 = min(ty2, 7);
  const size_t p = ((size_t)img * H + y) * W + x;
  const int v = src[p];
  const uint8_t *L = lut + (size_t)img * 64 * 256;
  const float l11 = (float)L[(ty1 * 8 + tx1) * 256 + v];
  const float l12 = (float)L[(ty1 * 8 + tx2) * 256 + v];
  const float l21 = (float)L[(ty2 * 8 + tx1) * 256 + v];
  const float l22 = (float)L[(ty2 * 8 + tx2) * 256 + v];
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  int q = __float2int_rn(res);
  q = min(max(q, 0), 255);
  dst[p] = (uint8_t)q;
}

// Same arithmetic, four pixels per thread (one 32-bit load and store): used when W % 4 == 0.
__global__ void __launch_bounds__(256)
clahe_apply4_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const uint8_t *__restrict__ lut,
                    int H, int W, float inv_tw, float inv_th) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (x0 >= W) return;
  const float yf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int ty1 = (int)floorf(yf);
  int ty2 = ty1 + 1;
  const float ya = __fsub_rn(yf, (float)ty1);
  const float ya1 = __fsub_rn(1.0f, ya);
  ty1 = max(ty1, 0);
  ty2 = min(ty2, 7);
  const size_t p = ((size_t)img * H + y) * W + x0;
  const uint32_t v4 = *reinterpret_cast<const uint32_t *>(src + p);
  const uint8_t *L = lut + (size_t)img * 64 * 256;
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float xf = __fsub_rn(__fmul_rn((float)(x0 + k), inv_tw), 0.5f);
    int tx1 = (int)floorf(xf);
    int tx2 = tx1 + 1;
    const float xa = __fsub_rn(xf, (float)tx1);
    const float xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0);
    tx2 = min(tx2, 7);
    const int v = (v4 >> (8 * k)) & 0xff;
    const float l11 = (float)L[(ty1 * 8 + tx1) * 256 + v];
    const float l12 = (float)L[(ty1 * 8 + tx2) * 256 + v];
    const float l21 = (float)L[(ty2 * 8 + tx1) * 256 + v];
    const float l22 = (float)L[(ty2 * 8 + tx2) * 256 + v];
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    int q = __float2int_rn(res);
    q = min(max(q, 0), 255);
    o |= (uint32_t)q << (8 * k);
  }
  *reinterpret_cast<uint32_t *>(dst + p) = o;
}

// ───────────────────────── A.3 adaptive Gaussian threshold ─────────────────────────
// cv2.getGaussianKernel(21, 0, CV_32F) bit patterns (sigma = 3.5).
__constant__ uint32_t c_gauss21[21] = {
    0x3afcd8aau, 0x3b8946cfu, 0x3c09607cu, 0x3c7d66a6u, 0x3cd7632bu, 0x3d28b99eu, 0x3d739f36u,
    0x3da21867u, 0x3dc6cb1eu, 0x3de0b045u, 0x3dea0c9bu, 0x3de0b045u, 0x3dc6cb1eu, 0x3da21867u,
    0x3d739f36u, 0x3d28b99eu, 0x3cd7632bu, 0x3c7d66a6u, 0x3c09607cu, 0x3b8946cfu, 0x3afcd8aau};

constexpr int AT_TW = 64, AT_TH = 32, AT_R = 10;
constexpr int AT_SW = AT_TW + 2 * AT_R;  // 84
constexpr int AT_SH = AT_TH + 2 * AT_R;  // 52

// One CTA = one 64x32 output tile.  Source tile (+10 halo, replicate border) staged in shared
// memory; the fp32 row pass (sequential FMA) is kept in shared memory for the column pass
// (symmetric FMA), then round-half-even and compare: one HBM read + one HBM write per pixel.
__global__ void __launch_bounds__(256)
adaptive_thresh_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W) {
  __shared__ __align__(16) uint8_t s_src[AT_SH][AT_SW + 4];
  __shared__ float s_row[AT_SH][AT_TW + 1];
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * AT_TW, y0 = blockIdx.y * AT_TH;
  const uint8_t *im = src + (size_t)img * H * W;
  for (int p = threadIdx.x; p < AT_SH * AT_SW; p += 256) {
    const int r = p / AT_SW, c = p % AT_SW;
    const int yy = min(max(y0 + r - AT_R, 0), H - 1);
    const int xx = min(max(x0 + c - AT_R, 0), W - 1);
    s_src[r][c] = im[(size_t)yy * W + xx];
  }
  __syncthreads();
  // row pass: one thread = 8 consecutive columns of one row; its 28 source bytes are loaded once (7 words) and converted
  // once, each output still accumulates its 21 taps in OpenCV's order
  for (int p = threadIdx.x; p < AT_SH * (AT_TW / 8); p += 256) {
    const int r = p / (AT_TW / 8), c0 = (p % (AT_TW / 8)) * 8;
    float v[28];
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(&s_src[r][c0]);
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const uint32_t w4 = sw[q];
      v[4 * q] = (float)(w4 & 0xff);
      v[4 * q + 1] = (float)((w4 >> 8) & 0xff);
      v[4 * q + 2] = (float)((w4 >> 16) & 0xff);
      v[4 * q + 3] = (float)(w4 >> 24);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < 21; ++j) acc = __fmaf_rn(v[o + j], __uint_as_float(c_gauss21[j]), acc);
      s_row[r][c0 + o] = acc;
    }
  }
  __syncthreads();
  // column pass: one thread = 8 consecutive rows of one column (28 row-pass values loaded once)
  {
    const int c = threadIdx.x % AT_TW, r0 = (threadIdx.x / AT_TW) * 8;
    float v[28];
#pragma unroll
    for (int q = 0; q < 28; ++q) v[q] = s_row[r0 + q][c];
    const int x = x0 + c;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int y = y0 + r0 + o;
      float acc = __fmaf_rn(v[o + AT_R], __uint_as_float(c_gauss21[10]), 0.0f);
#pragma unroll
      for (int i = 1; i <= 10; ++i)
        acc = __fmaf_rn(__fadd_rn(v[o + AT_R + i], v[o + AT_R - i]), __uint_as_float(c_gauss21[10 + i]), acc);
      int mean = __float2int_rn(acc);
      mean = min(max(mean, 0), 255);
      const int sv = s_src[r0 + o + AT_R][c + AT_R];
      if (y < H && x < W) dst[((size_t)img * H + y) * W + x] = (sv - mean > -10) ? 255 : 0;
    }
  }
}

// ───────────────────────── A.4 sharpen ─────────────────────────
__global__ void __launch_bounds__(256)
sharpen_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int C) {
  const int xb = blockIdx.x * blockDim.x + threadIdx.x;  // byte index within the row
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  const int rowb = W * C;
  if (xb >= rowb) return;
  const int x = xb / C, c = xb - x * C;
  const uint8_t *im = src + (size_t)img * H * rowb;
  const int yu = reflect101(y - 1, H), yd = reflect101(y + 1, H);
  const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
  const int ctr = im[(size_t)y * rowb + xb];
  int v = 5 * ctr - im[(size_t)yu * rowb + xb] - im[(size_t)yd * rowb + xb] - im[(size_t)y * rowb + xl * C + c] -
          im[(size_t)y * rowb + xr * C + c];
  v = min(max(v, 0), 255);
  dst[((size_t)img * H + y) * rowb + xb] = (uint8_t)v;
}

// Four bytes per thread (rows of W * C bytes with W * C % 4 == 0): aligned 32-bit loads of the rows above / below and of the
// previous / current / next word of the row, left / right neighbours (C bytes away) picked with byte permutes.  Words
// that touch the first or last pixel of the row (reflect-101) take the byte path.
__global__ void __launch_bounds__(256)
sharpen4_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int C) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  const int rowb = W * C, nw = rowb >> 2;
  if (w >= nw) return;
  const uint8_t *im = src + (size_t)img * H * rowb;
  uint8_t *om = dst + ((size_t)img * H + y) * rowb;
  const int yu = reflect101(y - 1, H), yd = reflect101(y + 1, H);
  const int xb = 4 * w;
  if (w >= 1 && w + 1 < nw && xb + 3 < rowb - C) {
    const uint32_t *rc = reinterpret_cast<const uint32_t *>(im + (size_t)y * rowb);
    const uint32_t up = reinterpret_cast<const uint32_t *>(im + (size_t)yu * rowb)[w];
    const uint32_t dn = reinterpret_cast<const uint32_t *>(im + (size_t)yd * rowb)[w];
    const uint32_t prev = rc[w - 1], cur = rc[w], next = rc[w + 1];
    const uint32_t left = C == 3 ? __byte_perm(prev, cur, 0x4321) : __byte_perm(prev, cur, 0x6543);
    const uint32_t right = C == 3 ? __byte_perm(cur, next, 0x6543) : __byte_perm(cur, next, 0x4321);
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sh = 8 * k;
      int v = 5 * (int)((cur >> sh) & 0xff) - (int)((up >> sh) & 0xff) - (int)((dn >> sh) & 0xff) -
              (int)((left >> sh) & 0xff) - (int)((right >> sh) & 0xff);
      v = min(max(v, 0), 255);
      o |= (uint32_t)v << sh;
    }
    reinterpret_cast<uint32_t *>(om)[w] = o;
  } else {
    for (int k = 0; k < 4; ++k) {
      const int b = xb + k;
      const int x = b / C, c = b - x * C;
      const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
      int v = 5 * im[(size_t)y * rowb + b] - im[(size_t)yu * rowb + b] - im[(size_t)yd * rowb + b] -
              im[(size_t)y * rowb + xl * C + c] - im[(size_t)y * rowb + xr * C + c];
      om[b] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

// ───────────────────────── A.5 deskew ─────────────────────────
// (a) per-row extents of dark (<128) pixels: one warp per row.
__global__ void __launch_bounds__(256)
dark_extents_kernel(const uint8_t *__restrict__ src, int32_t *__restrict__ ext, int H, int W, int C, int n_rows_total) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows_total) return;
  const uint8_t *r = src + (size_t)row * W * C;
  int cnt = 0, mn = W, mx = -1;
  for (int x = lane; x < W; x += 32) {
    uint32_t g;
    if (C == 3) g = gray_px(r[3 * x], r[3 * x + 1], r[3 * x + 2]);
    else g = r[x];
    if (g < 128) {
      ++cnt;
      mn = min(mn, x);
      mx = max(mx, x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    ext[(size_t)row * 3 + 0] = cnt;
    ext[(size_t)row * 3 + 1] = mn;
    ext[(size_t)row * 3 + 2] = mx;
  }
}

__device__ __forceinline__ long long cross3(int ox, int oy, int ax, int ay, int bx, int by) {
  return (long long)(ax - ox) * (by - oy) - (long long)(ay - oy) * (bx - ox);
}

// Sixteen bytes per thread (rows of W * C bytes with W * C % 16 == 0 and 16-byte aligned images): one 16-byte load of the
// row above, the row itself and the row below, plus the word before and the word after the 16 bytes; the left / right
// neighbours (C bytes away) of every word come out of byte permutes.  The first and last 16 bytes of a row (reflect-101
// at the image border) take the byte path.  Same integer arithmetic as above.
__global__ void __launch_bounds__(256)
sharpen16_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int C) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  const int rowb = W * C, nv = rowb >> 4;
  if (v >= nv) return;
  const uint8_t *im = src + (size_t)img * H * rowb;
  uint8_t *om = dst + ((size_t)img * H + y) * rowb;
  const int yu = reflect101(y - 1, H), yd = reflect101(y + 1, H);
  if (v >= 1 && v + 1 < nv) {
    const uint4 up = reinterpret_cast<const uint4 *>(im + (size_t)yu * rowb)[v];
    const uint4 dn = reinterpret_cast<const uint4 *>(im + (size_t)yd * rowb)[v];
    const uint32_t *rc = reinterpret_cast<const uint32_t *>(im + (size_t)y * rowb);
    const uint4 cu = reinterpret_cast<const uint4 *>(rc)[v];
    const uint32_t w[6] = {rc[4 * v - 1], cu.x, cu.y, cu.z, cu.w, rc[4 * v + 4]};
    const uint32_t u[4] = {up.x, up.y, up.z, up.w}, d[4] = {dn.x, dn.y, dn.z, dn.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t prev = w[j], cur = w[j + 1], next = w[j + 2];
      const uint32_t left = C == 3 ? __byte_perm(prev, cur, 0x4321) : __byte_perm(prev, cur, 0x6543);
      const uint32_t right = C == 3 ? __byte_perm(cur, next, 0x6543) : __byte_perm(cur, next, 0x4321);
      uint32_t ow = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sh = 8 * k;
        int val = 5 * (int)((cur >> sh) & 0xff) - (int)((u[j] >> sh) & 0xff) - (int)((d[j] >> sh) & 0xff) -
                  (int)((left >> sh) & 0xff) - (int)((right >> sh) & 0xff);
        val = min(max(val, 0), 255);
        ow |= (uint32_t)val << sh;
      }
      o[j] = ow;
    }
    reinterpret_cast<uint4 *>(om)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
    for (int k = 0; k < 16; ++k) {
      const int b = 16 * v + k;
      const int x = b / C, c = b - x * C;
      const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
      int val = 5 * im[(size_t)y * rowb + b] - im[(size_t)yu * rowb + b] - im[(size_t)yd * rowb + b] -
                im[(size_t)y * rowb + xl * C + c] - im[(size_t)y * rowb + xr * C + c];
      om[b] = (uint8_t)min(max(val, 0), 255);
    }
  }
}

// (b) one CTA per image: monotone-chain hull of the <= 2H extent points (thread 0; the points are
// already sorted: x = row ascending, y = min col then max col), then OpenCV's float32 rotating calipers
// restated step by step (bit-equal angle), atan2 in double, rotation matrix.
// The reference hands (row, col) to minAreaRect as (x, y) (tools.py:557-560).
__global__ void __launch_bounds__(256)
deskew_angle_kernel(const int32_t *__restrict__ ext, int H, int W, double *__restrict__ out_angle,
                    double *__restrict__ out_M, int32_t *__restrict__ hull_ws, int use_smem) {
  const int img = blockIdx.x;
  const int32_t *e = ext + (size_t)img * H * 3;
  int32_t *pts = hull_ws + (size_t)img * (4 * H + 8) * 2;  // [2H+4][2] candidate points
  int32_t *hull = pts + (2 * H + 4) * 2;                   // [2H+4][2] hull
  __shared__ int s_np, s_nh, s_total;
  // The hull scan and the calipers are one thread's sequential work.  With its arrays in global memory every step was an
  // L2 round trip (read-after-write of the hull stack): 640 us per page.  When they fit, the row extents are staged into
  // shared memory by the whole CTA and the candidate / hull arrays live there too (same algorithm, same order).
  extern __shared__ int32_t dk_smem[];
  if (use_smem) {
    int32_t *se = dk_smem;
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) se[i] = e[i];
    e = se;
    pts = dk_smem + 3 * H;
    hull = pts + (2 * H + 4) * 2;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int total = 0, np = 0;
    for (int y = 0; y < H; ++y) {
      const int c = e[y * 3];
      total += c;
      if (c > 0) {
        const int mn = e[y * 3 + 1], mx = e[y * 3 + 2];
        pts[2 * np] = y; pts[2 * np + 1] = mn; ++np;
        if (mx != mn) { pts[2 * np] = y; pts[2 * np + 1] = mx; ++np; }
      }
    }
    s_total = total;
    s_np = np;
    int k = 0;
    if (total > 100 && np >= 3) {
      // lower hull
      for (int i = 0; i < np; ++i) {
        const int qx = pts[2 * i], qy = pts[2 * i + 1];
        while (k >= 2 && cross3(hull[2 * (k - 2)], hull[2 * (k - 2) + 1], hull[2 * (k - 1)], hull[2 * (k - 1) + 1], qx, qy) <= 0) --k;
        hull[2 * k] = qx; hull[2 * k + 1] = qy; ++k;
      }
      // upper hull
      const int lo = k + 1;
      for (int i = np - 2; i >= 0; --i) {
        const int qx = pts[2 * i], qy = pts[2 * i + 1];
        while (k >= lo && cross3(hull[2 * (k - 2)], hull[2 * (k - 2) + 1], hull[2 * (k - 1)], hull[2 * (k - 1) + 1], qx, qy) <= 0) --k;
        hull[2 * k] = qx; hull[2 * k + 1] = qy; ++k;
      }
      --k;  // last point equals the first
    }
    s_nh = k;
  }
  __syncthreads();
  const int nh = s_nh;
  if (s_total <= 100 || nh < 3) {
    if (threadIdx.x == 0) {
      // <= 100 dark pixels: unchanged image (tools.py:558-559).  Degenerate hulls (all dark pixels
      // collinear) are reported as unsupported by NaN as well.
      out_angle[img] = nan("");
      for (int q = 0; q < 6; ++q) out_M[img * 6 + q] = nan("");
    }
    return;
  }
  if (threadIdx.x == 0) {
    // ---- cv::minAreaRect (OpenCV 4.13) restated: hull in cv2.convexHull(clockwise=false) order (same vertices as the
    // monotone chain, starting at the vertex with the largest x, ties -> largest y), then rotatingCalipers in float32
    // with every operation rounded separately; the advancing caliper is chosen by exact cross products between the four
    // candidate edges rotated into one frame (firstVecIsRight); `area <= minarea` keeps the LAST minimum.  Sequential
    // by nature (each step depends on the previous caliper state); the hull has a few dozen vertices.
    int start = 0;
    for (int i = 1; i < nh; ++i)
      if (hull[2 * i] > hull[2 * start] || (hull[2 * i] == hull[2 * start] && hull[2 * i + 1] > hull[2 * start + 1])) start = i;
    auto PX = [&](int i) { int q = start + i; if (q >= nh) q -= nh; return (float)hull[2 * q]; };
    auto PY = [&](int i) { int q = start + i; if (q >= nh) q -= nh; return (float)hull[2 * q + 1]; };
    auto VX = [&](int i) { return __fsub_rn(PX(i + 1 == nh ? 0 : i + 1), PX(i)); };   // exact: small integers
    auto VY = [&](int i) { return __fsub_rn(PY(i + 1 == nh ? 0 : i + 1), PY(i)); };
    auto INV = [&](int i) {
      const double dx = (double)VX(i), dy = (double)VY(i);
      return (float)(1.0 / sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    };
    int left = 0, bottom = 0, right = 0, top = 0;
    float left_x = PX(0), right_x = PX(0), top_y = PY(0), bottom_y = PY(0);
    for (int i = 0; i < nh; ++i) {
      const float x = PX(i), y = PY(i);
      if (x < left_x) { left_x = x; left = i; }
      if (x > right_x) { right_x = x; right = i; }
      if (y > top_y) { top_y = y; top = i; }
      if (y < bottom_y) { bottom_y = y; bottom = i; }
    }
    float orientation = 0.f;
    {
      double ax = (double)VX(nh - 1), ay = (double)VY(nh - 1);
      for (int i = 0; i < nh; ++i) {
        const double bx = (double)VX(i), by = (double)VY(i);
        const double convexity = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        if (convexity != 0.0) { orientation = convexity > 0.0 ? 1.f : -1.f; break; }
        ax = bx; ay = by;
      }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    float minarea = 3.402823466e+38f;
    float bA = 1.f, bB = 0.f, bW = 0.f, bH = 0.f;
    for (int k = 0; k < nh; ++k) {
      // candidate edges rotated into the frame of caliper 0: identity, 90 CW, 180, 90 CCW
      long long rx[4], ry[4];
      rx[0] = (long long)VX(seq[0]);  ry[0] = (long long)VY(seq[0]);
      rx[1] = (long long)VY(seq[1]);  ry[1] = -(long long)VX(seq[1]);
      rx[2] = -(long long)VX(seq[2]); ry[2] = -(long long)VY(seq[2]);
      rx[3] = -(long long)VY(seq[3]); ry[3] = (long long)VX(seq[3]);
      int main_el = 0;
      for (int i = 1; i < 4; ++i)
        if (ry[i] * rx[main_el] - rx[i] * ry[main_el] < 0) main_el = i;     // rotate90CW(v_i) . v_main < 0
      const int pindex = seq[main_el];
      const float inv = INV(pindex);
      const float lead_x = __fmul_rn(VX(pindex), inv), lead_y = __fmul_rn(VY(pindex), inv);
      switch (main_el) {
        case 0: base_a = lead_x;  base_b = lead_y;  break;
        case 1: base_a = lead_y;  base_b = -lead_x; break;
        case 2: base_a = -lead_x; base_b = -lead_y; break;
        default: base_a = -lead_y; base_b = lead_x; break;
      }
      seq[main_el] = (seq[main_el] + 1 == nh) ? 0 : seq[main_el] + 1;
      float dx = __fsub_rn(PX(seq[1]), PX(seq[3])), dy = __fsub_rn(PY(seq[1]), PY(seq[3]));
      const float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
      dx = __fsub_rn(PX(seq[2]), PX(seq[0]));
      dy = __fsub_rn(PY(seq[2]), PY(seq[0]));
      const float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
      const float area = __fmul_rn(width, height);
      if (area <= minarea) { minarea = area; bA = base_a; bW = width; bB = base_b; bH = height; }
    }
    (void)bH;
    // side vector out[1] = (A1 * width, B1 * width), turned by exact quarter turns into [-pi/2, 0)
    const double PI = 3.14159265358979323846;
    double x = (double)__fmul_rn(bA, bW), y = (double)__fmul_rn(bB, bW);
    double r = atan2(y, x);
    for (int it = 0; it < 4 && r >= 0.0; ++it) { const double t = x; x = y; y = -t; r = atan2(y, x); }
    for (int it = 0; it < 4 && r < -PI / 2; ++it) { const double t = x; x = -y; y = t; r = atan2(y, x); }
    const float ang = (float)(r * 180.0 / PI);
    double angle = (double)ang;
    if (angle < -45.0) angle = -(90.0 + angle);
    else angle = -angle;
    out_angle[img] = angle;
    const double th = angle * (PI / 180.0);
    const double al = cos(th), be = sin(th);
    const double ccx = (double)(W / 2), ccy = (double)(H / 2);
    double *M = out_M + img * 6;
    M[0] = al;
    M[1] = be;
    M[2] = __dsub_rn(__dmul_rn(__dsub_rn(1.0, al), ccx), __dmul_rn(be, ccy));
    M[3] = -be;
    M[4] = al;
    M[5] = __dadd_rn(__dmul_rn(be, ccx), __dmul_rn(__dsub_rn(1.0, al), ccy));
  }
}

// ───────────── remove_lines: ruled-line mask (tools.py:592-614) ─────────────
// mask = dilate_1x3( open_{W/4 x 1}( adaptiveThreshold(255 - gray, MEAN_C, BINARY, 15, -2) ) ), all integer:
//   mean  = rint(sum15x15(255 - gray, replicate border) / 225)   (cv::boxFilter normalised, round half even)
//   th    = (255 - gray) - mean > 2 ? 255 : 0
//   open  = horizontal erosion then dilation with a W/4-wide window, anchor W/8 (outside: 255 for the erosion, 0 for the
//           dilation -- cv::morphologyDefaultBorderValue), done with per-row prefix counts
//   mask  = max over rows y-1, y, y+1
// The Telea inpaint that follows in the reference is not built; callers use the mask only to prove it is empty
// (then cv2.inpaint returns its input) and refuse otherwise.
__global__ void __launch_bounds__(256)
rl_thresh_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ th, int H, int W, int C) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, img = blockIdx.z;
  if (x >= W) return;
  const uint8_t *im = src + (size_t)img * H * W * C;
  auto inv_at = [&](int yy, int xx) {
    yy = min(max(yy, 0), H - 1);
    xx = min(max(xx, 0), W - 1);
    const uint8_t *p = im + ((size_t)yy * W + xx) * C;
    const int g = (C == 3) ? gray_px(p[0], p[1], p[2]) : p[0];
    return 255 - g;
  };
  int sum = 0;
  for (int dy = -7; dy <= 7; ++dy)
    for (int dx = -7; dx <= 7; ++dx) sum += inv_at(y + dy, x + dx);
  const int mean = __double2int_rn(__dmul_rn((double)sum, 1.0 / 225.0));
  th[((size_t)img * H + y) * W + x] = (inv_at(y, x) - mean > 2) ? 255 : 0;
}

// one CTA per row: the row lives in shared memory, prefix counts by one thread (W <= 8192), window tests by all
__global__ void __launch_bounds__(256)
rl_open_row_kernel(const uint8_t *__restrict__ th, uint8_t *__restrict__ op, int W, int kw) {
  extern __shared__ int rl_pre[];               // [W + 1] prefix counts, then [W] flags
  int *flag = rl_pre + (W + 1);
  const size_t row = blockIdx.x;
  const uint8_t *r = th + row * W;
  const int anchor = kw / 2;
  for (int x = threadIdx.x; x < W; x += 256) flag[x] = (r[x] == 0);       // erosion: the window must hold no zero
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    if (threadIdx.x == 0) {
      int acc = 0;
      rl_pre[0] = 0;
      for (int x = 0; x < W; ++x) { acc += flag[x]; rl_pre[x + 1] = acc; }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += 256) {
      const int lo = max(0, x - anchor), hi = min(W, x - anchor + kw);
      const int cnt = rl_pre[hi] - rl_pre[lo];
      if (pass == 0) flag[x] = (cnt == 0);       // eroded pixel is 255; dilation: the window must hold one such pixel
      else op[row * W + x] = (cnt > 0) ? 255 : 0;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
rl_dilate_v_kernel(const uint8_t *__restrict__ op, uint8_t *__restrict__ mask, int32_t *__restrict__ nonzero, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, img = blockIdx.z;
  if (x >= W) return;
  const uint8_t *o = op + (size_t)img * H * W;
  uint8_t v = o[(size_t)y * W + x];
  if (y > 0) v = max(v, o[(size_t)(y - 1) * W + x]);
  if (y + 1 < H) v = max(v, o[(size_t)(y + 1) * W + x]);
  mask[((size_t)img * H + y) * W + x] = v;
  if (v) atomicOr(nonzero + img, 1);
}

// OpenCV fixed-point bicubic table: int16 [1024][16], index (fy*32 + fx), built on the host.
__device__ int16_t g_cubic_itab[1024 * 16];

__global__ void __launch_bounds__(256)
warp_affine_cubic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int C,
                         const double *__restrict__ Mall) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (x >= W) return;
  const size_t img_off = (size_t)img * H * W * C;
  const uint8_t *im = src + img_off;
  uint8_t *o = dst + img_off + ((size_t)y * W + x) * C;
  const double *Mf = Mall + img * 6;
  double m0 = Mf[0], m1 = Mf[1], m2 = Mf[2], m3 = Mf[3], m4 = Mf[4], m5 = Mf[5];
  if (m0 != m0) {  // NaN: leave the image unchanged
    for (int c = 0; c < C; ++c) o[c] = im[((size_t)y * W + x) * C + c];
    return;
  }
  // invertAffineTransform as in cv::warpAffine
  double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
  D = (D != 0.0) ? 1.0 / D : 0.0;
  const double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D);
  m0 = A11;
  m1 = __dmul_rn(m1, -D);
  m3 = __dmul_rn(m3, -D);
  m4 = A22;
  const double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
  const double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
  m2 = b1;
  m5 = b2;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m0, (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m3, (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)y), m2), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)y), m5), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  int sx = X >> 5, sy = Y >> 5;
  sx = min(max(sx, -32768), 32767);
  sy = min(max(sy, -32768), 32767);
  // the 16 fixed-point weights of this sub-pixel position: two 16-byte loads (the table row is 32-byte aligned)
  const uint4 *wt4 = reinterpret_cast<const uint4 *>(g_cubic_itab + (((Y & 31) * 32 + (X & 31)) << 4));
  const uint4 wa = wt4[0], wb = wt4[1];
  const uint32_t wpk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};      // wpk[i] = weights 2i (low), 2i + 1 (high)
  auto wgt = [&](int i) -> int { return (int)(short)((i & 1) ? (wpk[i >> 1] >> 16) : (wpk[i >> 1] & 0xffffu)); };
  int ys[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ys[k] = min(max(sy - 1 + k, 0), H - 1);
  // Interior columns (no clamping in x, and room for the aligned over-read): the 4 taps x C channels of a source row
  // are 4C contiguous bytes at an arbitrary alignment -- fetched as aligned 32-bit words and funnel-shifted into place
  // (C = 3: 4 loads per row instead of 12 byte loads; the same integers enter the same sums).
  const int margin = (C == 3) ? 1 : 3;
  if (sx - 1 >= 0 && sx + 2 + margin <= W - 1 && (C == 1 || C == 3)) {
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const size_t off = ((size_t)ys[ky] * W + (sx - 1)) * C + img_off;
      const uint32_t *wp = reinterpret_cast<const uint32_t *>(src + (off & ~(size_t)3));
      const uint32_t sh = (uint32_t)(off & 3) * 8;
      if (C == 3) {
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const uint32_t b[3] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh)};
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int j = kx * 3 + c;
            acc[c] += wgt(ky * 4 + kx) * (int)((b[j >> 2] >> ((j & 3) * 8)) & 0xffu);
          }
      } else {
        const uint32_t w0 = wp[0], w1 = wp[1];
        const uint32_t b0 = __funnelshift_r(w0, w1, sh);
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) acc[0] += wgt(ky * 4 + kx) * (int)((b0 >> (kx * 8)) & 0xffu);
      }
    }
    for (int c = 0; c < C; ++c) {
      const int v = (acc[c] + 16384) >> 15;
      o[c] = (uint8_t)min(max(v, 0), 255);
    }
    return;
  }
  int xs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xs[k] = min(max(sx - 1 + k, 0), W - 1);
  for (int c = 0; c < C; ++c) {
    int acc = 0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx)
        acc += wgt(ky * 4 + kx) * (int)im[((size_t)ys[ky] * W + xs[kx]) * C + c];
    int v = (acc + 16384) >> 15;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

// Host: OpenCV's interpolateCubic + initInterTab2D (fp32, unfused), incl. the ksize/2 quirk.
static void build_cubic_itab(int16_t *out) {
  const float A = -0.75f;
  float tab[32][4];
  for (int i = 0; i < 32; ++i) {
    volatile float t = (float)i / 32.0f;
    volatile float x1 = t + 1.0f;
    volatile float c0 = A * x1;
    c0 = c0 - 5.0f * A; c0 = c0 * x1; c0 = c0 + 8.0f * A; c0 = c0 * x1; c0 = c0 - 4.0f * A;
    volatile float c1 = (A + 2.0f) * t;
    c1 = c1 - (A + 3.0f); c1 = c1 * t; c1 = c1 * t; c1 = c1 + 1.0f;
    volatile float u = 1.0f - t;
    volatile float c2 = (A + 2.0f) * u;
    c2 = c2 - (A + 3.0f); c2 = c2 * u; c2 = c2 * u; c2 = c2 + 1.0f;
    volatile float c3 = 1.0f - c0;
    c3 = c3 - c1; c3 = c3 - c2;
    tab[i][0] = c0; tab[i][1] = c1; tab[i][2] = c2; tab[i][3] = c3;
  }
  for (int fy = 0; fy < 32; ++fy)
    for (int fx = 0; fx < 32; ++fx) {
      int iw[4][4];
      int sum = 0;
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
          volatile float v = tab[fy][ky] * tab[fx][kx];
          volatile float sc = v * 32768.0f;
          long r = lrintf(sc);
          if (r > 32767) r = 32767;
          if (r < -32768) r = -32768;
          iw[ky][kx] = (int)r;
          sum += (int)r;
        }
      if (sum != 32768) {
        const int diff = sum - 32768;
        int mk1 = 2, mk2 = 2, Mk1 = 2, Mk2 = 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (iw[k1][k2] < iw[mk1][mk2]) { mk1 = k1; mk2 = k2; }
            else if (iw[k1][k2] > iw[Mk1][Mk2]) { Mk1 = k1; Mk2 = k2; }
          }
        if (diff < 0) iw[Mk1][Mk2] = (int16_t)(iw[Mk1][Mk2] - diff);
        else iw[mk1][mk2] = (int16_t)(iw[mk1][mk2] - diff);
      }
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) out[(fy * 32 + fx) * 16 + ky * 4 + kx] = (int16_t)iw[ky][kx];
    }
}

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

extern "C" int ocrb_clahe_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, uint8_t *lut_ws,
                             void *stream) {
  OCRB_REQUIRE(src && dst && lut_ws && n_img > 0 && H > 0 && W > 0, "clahe_u8: bad arguments");
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
  clahe_lut_kernel<<<dim3(64, n_img), 256, 0, (cudaStream_t)stream>>>(src, lut_ws, H, W, tw, th, clip, lut_scale);
  int rc = check_launch("clahe_lut_kernel");
  if (rc) return rc;
  volatile float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  if (W % 4 == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0)
    clahe_apply4_kernel<<<dim3(cdiv(W / 4, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, lut_ws, H, W, inv_tw, inv_th);
  else
    clahe_apply_kernel<<<dim3(cdiv(W, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, lut_ws, H, W, inv_tw, inv_th);
  return check_launch("clahe_apply_kernel");
}

extern "C" int ocrb_adaptive_gauss_thresh_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W,
                                             void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0, "adaptive_gauss_thresh_u8: bad arguments");
  adaptive_thresh_kernel<<<dim3(cdiv(W, AT_TW), cdiv(H, AT_TH), n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W);
  return check_launch("adaptive_thresh_kernel");
}

extern "C" int ocrb_sharpen3x3_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                  void *stream) {
  OCRB_REQUIRE(src && dst && n_img > 0 && H > 1 && W > 1 && (C == 1 || C == 3), "sharpen3x3_u8: bad arguments");
  if ((W * C) % 16 == 0 && W * C >= 48 && ((size_t)H * W * C) % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0)
    sharpen16_kernel<<<dim3(cdiv((long long)W * C / 16, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, C);
  else if ((W * C) % 4 == 0 && W >= 4 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0)
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
  dark_extents_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(src, ext_ws, H, W, C, rows);
  int rc = check_launch("dark_extents_kernel");
  if (rc) return rc;
  // extents [3H] + candidate points and hull [2H+4][2] each, in shared memory when they fit
  const size_t smem = ((size_t)3 * H + 2 * (size_t)(2 * H + 4) * 2) * sizeof(int32_t);
  const int use_smem = smem <= 200 * 1024;
  if (use_smem && smem > 48 * 1024) {
    static size_t attr = 48 * 1024;
    if (smem > attr) {
      OCRB_CUDA(cudaFuncSetAttribute(deskew_angle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
  }
  deskew_angle_kernel<<<n_img, 256, use_smem ? smem : 0, (cudaStream_t)stream>>>(ext_ws, H, W, out_angle, out_M, hull_ws, use_smem);
  return check_launch("deskew_angle_kernel");
}

extern "C" int ocrb_warp_affine_cubic_u8(const uint8_t *src, uint8_t *dst, int32_t n_img, int32_t H, int32_t W,
                                         int32_t C, const double *M, void *stream) {
  OCRB_REQUIRE(src && dst && M && n_img > 0 && H > 0 && W > 0 && (C == 1 || C == 3), "warp_affine_cubic_u8: bad arguments");
  OCRB_REQUIRE(src != dst, "warp_affine_cubic_u8: in-place not supported");
  int rc = ensure_itab();
  if (rc) return rc;
  warp_affine_cubic_kernel<<<dim3(cdiv(W, 256), H, n_img), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, C, M);
  return check_launch("warp_affine_cubic_kernel");
}
