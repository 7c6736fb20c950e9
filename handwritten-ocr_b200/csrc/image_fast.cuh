// Throughput versions of the page-sized uint8 image kernels (tools.py:503-573), same integer / fp32
// arithmetic as the general kernels in image.cu (which stay as the path for odd widths and tiny pages):
//   sharpen_vec16_kernel        two pixels per 32-bit register (16-bit lanes), 16 bytes per thread, no border divergence
//   clahe_hist_lut_kernel<C>    RGB->gray fused into the tile histogram pass (per-warp histograms, run-aggregated atomics)
//   clahe_apply_cells_kernel    one CTA per interpolation cell: its four LUTs as one float4 table in shared memory
//   adaptive_thresh_tile_kernel<C>  64 x 96 tiles, RGB->gray fused into the staging, word loads / stores
//   dark_extents16_kernel       16 pixels per load
//   deskew_angle_par_kernel     convex hull by a 64 -> 8 -> 1 tree of monotone chains instead of one thread's scan
//   warp_affine_cubic_dp2a_kernel   16 taps x 3 channels as 24 dp2a, matrix inversion once per CTA
// Kernel definitions and the host-side geometry helpers only -- no launches -- so that tests/emu can compile this
// file for the host (OCRB_EMU) and run the kernels thread by thread against the oracle without a GPU.
#pragma once
#ifndef OCRB_EMU
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>

#ifndef OCRB_EMU
#ifndef OCRB_DYN_SMEM
#define OCRB_DYN_SMEM(T, name) extern __shared__ T name[]
#endif
#endif

namespace ocrb {

// ───────────────────────── shared scalar helpers ─────────────────────────
__device__ __forceinline__ uint32_t gray_px(uint32_t r, uint32_t g, uint32_t b) {
  return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15;
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

// gray values of the 4 RGB pixels held in three consecutive 32-bit words, packed into one word.  The 15-bit coefficients
// split into byte pairs (9798 = 38 * 256 + 70, 19235 = 75 * 256 + 35, 3735 = 14 * 256 + 151), so a pixel's weighted sum is two
// dp4a over its three bytes (the fourth byte of the register meets a zero coefficient) instead of three IMAD on extracted bytes.
__device__ __forceinline__ uint32_t gray_px_dp4a(uint32_t rgbx) {
  const uint32_t hi = __dp4a(rgbx, 0x000e4b26u, 0u);      // 38 R + 75 G + 14 B
  const uint32_t lo = __dp4a(rgbx, 0x00972346u, 16384u);  // 70 R + 35 G + 151 B + rounding
  return (hi * 256u + lo) >> 15;
}
__device__ __forceinline__ uint32_t gray4_from_rgb12(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t g0 = gray_px_dp4a(a);
  const uint32_t g1 = gray_px_dp4a(__funnelshift_r(a, b, 24));
  const uint32_t g2 = gray_px_dp4a(__funnelshift_r(b, c, 16));
  const uint32_t g3 = gray_px_dp4a(c >> 8);
  return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

// byte k of w as a float without a conversion instruction: 2^23 + b has b in its low mantissa bits
template <int K>
__device__ __forceinline__ float u8_to_f32(uint32_t w) {
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4b000000u, 0x7540 + K)), 8388608.0f);
}

#ifndef OCRB_EMU
__device__ __forceinline__ int dp2a_lo_s16u8(uint32_t w2, uint32_t b4, int acc) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(b4), "r"(acc));
  return d;
}
__device__ __forceinline__ int dp2a_hi_s16u8(uint32_t w2, uint32_t b4, int acc) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(b4), "r"(acc));
  return d;
}
#endif

// ───────────────────────── A.4 sharpen ─────────────────────────
// One 32-bit word = 4 bytes of the row; `left` / `right` hold each byte's neighbour C bytes away.  Even and odd bytes are
// spread into 16-bit lanes: lane = 5c + 1020 - (u + d + l + r) stays in [0, 2295], so plain 32-bit adds never carry
// between lanes; one VIADDMNMX.S16x2.RELU per parity subtracts the bias and clamps to [0, 255].
__device__ __forceinline__ uint32_t sharpen_word(uint32_t cur, uint32_t up, uint32_t dn, uint32_t left, uint32_t right) {
  const uint32_t M = 0x00ff00ffu, K2 = 0x03fc03fcu;
  const uint32_t se = (up & M) + (dn & M) + (left & M);
  const uint32_t xe = (cur & M) * 5u + K2 - se - (right & M);
  const uint32_t so = __byte_perm(up, 0u, 0x4341) + __byte_perm(dn, 0u, 0x4341) + __byte_perm(left, 0u, 0x4341);
  const uint32_t xo = __byte_perm(cur, 0u, 0x4341) * 5u + K2 - so - __byte_perm(right, 0u, 0x4341);
  const uint32_t ve = __viaddmin_s16x2_relu(xe, 0xfc04fc04u, M);
  const uint32_t vo = __viaddmin_s16x2_relu(xo, 0xfc04fc04u, M);
  return __byte_perm(ve, vo, 0x6240);
}

// Rows of nv 16-byte vectors (W * C % 16 == 0, 16-byte aligned images, H >= 2), one vector per thread: a CTA covers 64
// vectors x 4 rows (the rows above / below are mostly its own rows, served by L1), blockIdx.z = image.
// The first / last vector of a row build the reflect-101 neighbour word from bytes of the row itself, so every thread
// runs the same instructions.
template <int C>
__global__ void __launch_bounds__(256)
sharpen_vec16_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int nv) {
  const int v = blockIdx.x * 64 + (threadIdx.x & 63);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (v >= nv || y >= H) return;
  // 32-bit vector offsets inside the image (H * nv < 2^27 for any page), one 64-bit base per image
  const uint4 *im = reinterpret_cast<const uint4 *>(src) + (size_t)blockIdx.z * H * nv;
  const unsigned rv = (unsigned)y * (unsigned)nv + (unsigned)v;
  const unsigned ru = (y == 0) ? rv + (unsigned)nv : rv - (unsigned)nv;          // reflect-101: row -1 -> row 1
  const unsigned rd = (y == H - 1) ? rv - (unsigned)nv : rv + (unsigned)nv;      //              row H -> row H - 2
  const uint4 cu = im[rv];
  const uint4 up = im[ru];
  const uint4 dn = im[rd];
  const uint32_t *imw = reinterpret_cast<const uint32_t *>(im);
  uint32_t wl, wr;
  if (v > 0) wl = imw[4u * rv - 1u];
  else wl = (C == 3) ? __byte_perm(cu.x, cu.y, 0x5430) : __byte_perm(cu.x, 0u, 0x1000);
  if (v + 1 < nv) wr = imw[4u * rv + 4u];
  else wr = (C == 3) ? __byte_perm(cu.z, cu.w, 0x0432) : __byte_perm(cu.w, 0u, 0x0002);
  const uint32_t w[6] = {wl, cu.x, cu.y, cu.z, cu.w, wr};
  const uint32_t u[4] = {up.x, up.y, up.z, up.w}, d[4] = {dn.x, dn.y, dn.z, dn.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t prev = w[j], cur = w[j + 1], next = w[j + 2];
    const uint32_t left = (C == 3) ? __byte_perm(prev, cur, 0x4321) : __byte_perm(prev, cur, 0x6543);
    const uint32_t right = (C == 3) ? __byte_perm(cur, next, 0x6543) : __byte_perm(cur, next, 0x4321);
    o[j] = sharpen_word(cur, u[j], d[j], left, right);
  }
  (reinterpret_cast<uint4 *>(dst) + (size_t)blockIdx.z * H * nv)[rv] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ───────────────────────── A.2 CLAHE ─────────────────────────
// Pass 1: one CTA per (tile, image).  C == 3: the tile is read as RGB, the gray page is written on the way (each image
// pixel belongs to exactly one tile), so RGB -> gray costs no pass of its own.  Histogram: one 256-bin table per warp,
// equal neighbours of a thread's 16 pixels merged into one shared-memory atomic (paper background = long runs).
// fast: the tile lies inside the image, tw % 16 == 0, W % 16 == 0 and the images are 16-byte aligned.
template <int C>
__global__ void __launch_bounds__(256)
clahe_hist_lut_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ gray_out, uint8_t *__restrict__ lut, int H,
                      int W, int tw, int th, int clip, float lut_scale, int vec_ok) {
  __shared__ int whist[8][256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int t = threadIdx.x;
  const int tile = blockIdx.x, img = blockIdx.y;
  const int ty = tile >> 3, tx = tile & 7;
  const uint8_t *im = src + (size_t)img * H * W * C;
  uint8_t *go = gray_out + (size_t)img * H * W;      // C == 3 only
#pragma unroll
  for (int k = 0; k < 8; ++k) whist[k][t] = 0;
  if (t == 0) s_clipped = 0;
  __syncthreads();
  int *wh = whist[t >> 5];
  const int x0 = tx * tw, y0 = ty * th;
  if (vec_ok && x0 + tw <= W && y0 + th <= H) {
    const int upr = tw >> 4, total = upr * th;
    for (int u = t; u < total; u += 256) {
      const int r = u / upr, cu = u - r * upr;
      const size_t p = (size_t)(y0 + r) * W + x0 + 16 * cu;
      union { uint4 v; uint8_t b[16]; } g;
      if (C == 3) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(im + p * 3);
        union { uint4 v[3]; uint32_t w[12]; } in;
        in.v[0] = __ldg(s4);
        in.v[1] = __ldg(s4 + 1);
        in.v[2] = __ldg(s4 + 2);
        g.v.x = gray4_from_rgb12(in.w[0], in.w[1], in.w[2]);
        g.v.y = gray4_from_rgb12(in.w[3], in.w[4], in.w[5]);
        g.v.z = gray4_from_rgb12(in.w[6], in.w[7], in.w[8]);
        g.v.w = gray4_from_rgb12(in.w[9], in.w[10], in.w[11]);
        *reinterpret_cast<uint4 *>(go + p) = g.v;
      } else {
        g.v = __ldg(reinterpret_cast<const uint4 *>(im + p));
      }
      int cur = g.b[0], n = 1;
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const int v = g.b[k];
        if (v == cur) {
          ++n;
        } else {
          atomicAdd(&wh[cur], n);
          cur = v;
          n = 1;
        }
      }
      atomicAdd(&wh[cur], n);
    }
  } else {
    for (int p = t; p < tw * th; p += 256) {
      const int py = p / tw, px = p - py * tw;
      const int yy = reflect101(y0 + py, H);
      const int xx = reflect101(x0 + px, W);
      int g;
      if (C == 3) {
        const uint8_t *q = im + ((size_t)yy * W + xx) * 3;
        g = (int)gray_px(q[0], q[1], q[2]);
        if (y0 + py < H && x0 + px < W) go[(size_t)yy * W + xx] = (uint8_t)g;
      } else {
        g = im[(size_t)yy * W + xx];
      }
      atomicAdd(&wh[g], 1);
    }
  }
  __syncthreads();
  int h = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) h += whist[k][t];
  const int excess = imax(h - clip, 0);
  int e = excess;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((t & 31) == 0) atomicAdd(&s_clipped, e);
  __syncthreads();
  const int clipped = s_clipped;
  h = imin(h, clip);
  const int batch = clipped / 256;
  const int resid = clipped - batch * 256;
  h += batch;
  if (resid) {
    const int step = imax(256 / resid, 1);
    if (t % step == 0 && t / step < resid) h += 1;
  }
  scan[t] = h;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int v = scan[t];
    if (t >= o) v += scan[t - o];
    __syncthreads();
    scan[t] = v;
    __syncthreads();
  }
  const float f = __fmul_rn((float)scan[t], lut_scale);
  int q = __float2int_rn(f);
  q = imin(imax(q, 0), 255);
  lut[((size_t)img * 64 + tile) * 256 + t] = (uint8_t)q;
}

// Pass 2.  A pixel blends the LUTs of the four tiles around it; which four changes at the tile CENTRES, so the image
// splits into 9 x 9 cells with one LUT quadruple each.  The host finds the cell boundaries with the same fp32
// arithmetic as the per-pixel formula (clahe_cells_host), one CTA per (cell, image) keeps the quadruple as a float4
// table -- one 16-byte shared-memory load per pixel instead of four global byte loads and four conversions -- and the
// column weights of the cell in shared memory as well.  Rounding by the 1.5 * 2^23 add (round-half-even, as cvRound).
constexpr int CLAHE_MAX_CELL = 1024;   // widest / tallest cell the shared-memory weight tables hold
struct ClaheCells {
  int xb[10];
  int yb[10];
};

// cell index of coordinate i: floor(i * inv - 0.5) + 1, in fp32 exactly as clahe_apply_kernel evaluates it
static inline int clahe_cell_of(int i, float inv) {
  volatile float f = (float)i * inv;
  volatile float g = f - 0.5f;
  int c = (int)floorf(g) + 1;
  return c < 0 ? 0 : (c > 8 ? 8 : c);
}

// returns 1 when clahe_apply_cells_kernel applies (boundaries multiples of 4, cells no wider than the shared table)
static inline int clahe_cells_host(int H, int W, float inv_tw, float inv_th, ClaheCells *cells) {
  int ok = (W % 4 == 0);
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass ? H : W;
    const float inv = pass ? inv_th : inv_tw;
    int *b = pass ? cells->yb : cells->xb;
    for (int k = 0; k < 10; ++k) b[k] = n;
    b[0] = 0;
    int prev = 0;
    for (int i = 0; i < n; ++i) {
      const int c = clahe_cell_of(i, inv);
      if (c < prev) return 0;      // cannot happen (monotone); refuse rather than mis-assign
      for (int k = prev + 1; k <= c; ++k) b[k] = i;
      prev = c;
    }
    b[9] = n;
  }
  for (int k = 0; k < 9; ++k) {
    if (cells->xb[k] % 4) ok = 0;
    if (cells->xb[k + 1] - cells->xb[k] > CLAHE_MAX_CELL) ok = 0;
    if (cells->yb[k + 1] - cells->yb[k] > CLAHE_MAX_CELL) ok = 0;
  }
  return ok;
}

// one pixel: the reference's unfused blend, rounded half-to-even by the 1.5 * 2^23 add; the result's low byte IS the
// output (a convex blend of values <= 255 cannot round above 255, nor below 0), so no clamp and no conversion
__device__ __forceinline__ uint32_t clahe_blend_bits(const float4 l, float xa, float xa1, float ya, float ya1) {
  const float top = __fadd_rn(__fmul_rn(l.x, xa1), __fmul_rn(l.y, xa));
  const float bot = __fadd_rn(__fmul_rn(l.z, xa1), __fmul_rn(l.w, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  return (uint32_t)__float_as_int(__fadd_rn(res, 12582912.0f));
}

__device__ __forceinline__ uint32_t clahe_blend4(uint32_t v4, const float4 *s_lut, const float *xa, const float *xa1,
                                                 float ya, float ya1) {
  const uint32_t q0 = clahe_blend_bits(s_lut[v4 & 0xffu], xa[0], xa1[0], ya, ya1);
  const uint32_t q1 = clahe_blend_bits(s_lut[(v4 >> 8) & 0xffu], xa[1], xa1[1], ya, ya1);
  const uint32_t q2 = clahe_blend_bits(s_lut[(v4 >> 16) & 0xffu], xa[2], xa1[2], ya, ya1);
  const uint32_t q3 = clahe_blend_bits(s_lut[v4 >> 24], xa[3], xa1[3], ya, ya1);
  return __byte_perm(__byte_perm(q0, q1, 0x0040), __byte_perm(q2, q3, 0x0040), 0x5410);
}

__global__ void __launch_bounds__(256)
clahe_apply_cells_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const uint8_t *__restrict__ lut,
                         int H, int W, float inv_tw, float inv_th, ClaheCells cells) {
  __shared__ float4 s_lut[256];
  __shared__ __align__(16) float s_xa[CLAHE_MAX_CELL];
  __shared__ __align__(16) float s_xa1[CLAHE_MAX_CELL];
  __shared__ float2 s_ya[CLAHE_MAX_CELL];
  const int t = threadIdx.x;
  const int cx = blockIdx.x % 9, cy = blockIdx.x / 9, img = blockIdx.y;
  const int x_lo = cells.xb[cx], x_hi = cells.xb[cx + 1];
  const int y_lo = cells.yb[cy], y_hi = cells.yb[cy + 1];
  const int w = x_hi - x_lo, h = y_hi - y_lo;
  if (w <= 0 || h <= 0) return;
  const int tx1 = imax(cx - 1, 0), tx2 = imin(cx, 7), ty1 = imax(cy - 1, 0), ty2 = imin(cy, 7);
  const uint8_t *L = lut + (size_t)img * 64 * 256;
  s_lut[t] = make_float4((float)L[(ty1 * 8 + tx1) * 256 + t], (float)L[(ty1 * 8 + tx2) * 256 + t],
                         (float)L[(ty2 * 8 + tx1) * 256 + t], (float)L[(ty2 * 8 + tx2) * 256 + t]);
  for (int c = t; c < w; c += 256) {
    const float xf = __fsub_rn(__fmul_rn((float)(x_lo + c), inv_tw), 0.5f);
    const float xa = __fsub_rn(xf, floorf(xf));
    s_xa[c] = xa;
    s_xa1[c] = __fsub_rn(1.0f, xa);
  }
  for (int r = t; r < h; r += 256) {
    const float yf = __fsub_rn(__fmul_rn((float)(y_lo + r), inv_th), 0.5f);
    const float ya = __fsub_rn(yf, floorf(yf));
    s_ya[r] = make_float2(ya, __fsub_rn(1.0f, ya));
  }
  __syncthreads();
  const int gpr = w >> 2;
  const size_t cell0 = ((size_t)img * H + y_lo) * W + x_lo;
  if (gpr <= 256 && (gpr & (gpr - 1)) == 0) {
    // 256 / gpr rows per pass; a thread keeps its column group, so its four column weights live in registers
    const int g = t & (gpr - 1), rpp = 256 / gpr;
    const float4 a4 = *reinterpret_cast<const float4 *>(&s_xa[4 * g]);
    const float4 b4 = *reinterpret_cast<const float4 *>(&s_xa1[4 * g]);
    const float xa[4] = {a4.x, a4.y, a4.z, a4.w}, xa1[4] = {b4.x, b4.y, b4.z, b4.w};
    int r = t / gpr;
    const uint8_t *sp = src + cell0 + (size_t)r * W + 4 * g;
    uint8_t *dp = dst + cell0 + (size_t)r * W + 4 * g;
    const size_t step = (size_t)rpp * W;
    for (; r + rpp < h; r += 2 * rpp) {           // two rows in flight
      const uint32_t va = *reinterpret_cast<const uint32_t *>(sp);
      const uint32_t vb = *reinterpret_cast<const uint32_t *>(sp + step);
      const float2 ya = s_ya[r], yb = s_ya[r + rpp];
      *reinterpret_cast<uint32_t *>(dp) = clahe_blend4(va, s_lut, xa, xa1, ya.x, ya.y);
      *reinterpret_cast<uint32_t *>(dp + step) = clahe_blend4(vb, s_lut, xa, xa1, yb.x, yb.y);
      sp += 2 * step;
      dp += 2 * step;
    }
    if (r < h) {
      const float2 ya = s_ya[r];
      *reinterpret_cast<uint32_t *>(dp) = clahe_blend4(*reinterpret_cast<const uint32_t *>(sp), s_lut, xa, xa1, ya.x, ya.y);
    }
    return;
  }
  const int total = gpr * h;
  const int dr = 256 / gpr, dg = 256 - dr * gpr;
  int r = t / gpr, g = t - r * gpr;
  for (int i = t; i < total; i += 256) {
    const float2 ya = s_ya[r];
    const size_t p = cell0 + (size_t)r * W + 4 * g;
    const float4 a4 = *reinterpret_cast<const float4 *>(&s_xa[4 * g]);
    const float4 b4 = *reinterpret_cast<const float4 *>(&s_xa1[4 * g]);
    const float xa[4] = {a4.x, a4.y, a4.z, a4.w}, xa1[4] = {b4.x, b4.y, b4.z, b4.w};
    *reinterpret_cast<uint32_t *>(dst + p) = clahe_blend4(*reinterpret_cast<const uint32_t *>(src + p), s_lut, xa, xa1, ya.x, ya.y);
    r += dr;
    g += dg;
    if (g >= gpr) {
      g -= gpr;
      ++r;
    }
  }
}

// ───────────────────────── A.3 adaptive Gaussian threshold ─────────────────────────
// cv2.getGaussianKernel(21, 0, CV_32F) bit patterns (sigma = 3.5).
__constant__ uint32_t c_gauss21[21] = {
    0x3afcd8aau, 0x3b8946cfu, 0x3c09607cu, 0x3c7d66a6u, 0x3cd7632bu, 0x3d28b99eu, 0x3d739f36u,
    0x3da21867u, 0x3dc6cb1eu, 0x3de0b045u, 0x3dea0c9bu, 0x3de0b045u, 0x3dc6cb1eu, 0x3da21867u,
    0x3d739f36u, 0x3d28b99eu, 0x3cd7632bu, 0x3c7d66a6u, 0x3c09607cu, 0x3b8946cfu, 0x3afcd8aau};

constexpr int AT2_TW = 64, AT2_TH = 96, AT2_R = 10, AT2_HX = 12;
constexpr int AT2_SW = AT2_TW + 2 * AT2_HX;  // 88 staged columns: x0 - 12 .. x0 + 75 (word aligned)
constexpr int AT2_SH = AT2_TH + 2 * AT2_R;   // 116 staged rows
constexpr int AT2_WPR = AT2_SW / 4;          // 22 words per staged row

// One CTA = one 64 x 96 output tile (halo rows 116 / 96 = 1.21x instead of 52 / 32 = 1.63x).  The source tile is staged as
// gray bytes by 32-bit loads (C == 3: three words -> four gray pixels on the way, no gray page in memory); the fp32 row
// pass (sequential FMA, OpenCV's order) converts bytes with a PRMT + FSUB instead of I2F; the column pass keeps a
// column's 44 row-pass values in registers for 24 outputs; results replace the tile's own centre bytes in shared memory
// and leave as 32-bit stores.  aligned: W % 4 == 0 and 4-byte aligned images.
template <int C, int OCC = 4>
__global__ void __launch_bounds__(256, OCC)
adaptive_thresh_tile_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int aligned) {
  __shared__ __align__(16) uint8_t s_src[AT2_SH][AT2_SW];
  __shared__ float s_row[AT2_SH][AT2_TW + 1];
  const int t = threadIdx.x;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * AT2_TW, y0 = blockIdx.y * AT2_TH;
  const uint8_t *im = src + (size_t)img * H * W * C;
  for (int p = t; p < AT2_SH * AT2_WPR; p += 256) {
    const int r = p / AT2_WPR, k = p - r * AT2_WPR;
    const int yy = imin(imax(y0 + r - AT2_R, 0), H - 1);
    const int xs = x0 - AT2_HX + 4 * k;
    const uint8_t *rowp = im + (size_t)yy * W * C;
    uint32_t g4;
    if (aligned && xs >= 0 && xs + 3 < W) {
      if (C == 1) {
        g4 = *reinterpret_cast<const uint32_t *>(rowp + xs);
      } else {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(rowp + 3 * xs);
        g4 = gray4_from_rgb12(q[0], q[1], q[2]);
      }
    } else {
      g4 = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = imin(imax(xs + b, 0), W - 1);
        const uint32_t g = (C == 1) ? (uint32_t)rowp[xx] : gray_px(rowp[3 * xx], rowp[3 * xx + 1], rowp[3 * xx + 2]);
        g4 |= g << (8 * b);
      }
    }
    *reinterpret_cast<uint32_t *>(&s_src[r][4 * k]) = g4;
  }
  __syncthreads();
  // row pass: one task = 8 consecutive outputs of one staged row; output j reads staged columns j + 2 .. j + 22
  for (int p = t; p < AT2_SH * (AT2_TW / 8); p += 256) {
    const int r = p >> 3, c0 = (p & 7) * 8;
    const uint2 *sw = reinterpret_cast<const uint2 *>(&s_src[r][c0]);
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint2 d = sw[q];
      w[2 * q] = d.x;
      w[2 * q + 1] = d.y;
    }
    float v[28];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      // bytes 4q + 2 .. 4q + 5 of the 32 loaded
      v[4 * q] = u8_to_f32<2>(w[q]);
      v[4 * q + 1] = u8_to_f32<3>(w[q]);
      v[4 * q + 2] = u8_to_f32<0>(w[q + 1]);
      v[4 * q + 3] = u8_to_f32<1>(w[q + 1]);
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
  // column pass: one thread = 24 consecutive rows of one column
  {
    const int c = t & 63, r0 = (t >> 6) * 24;
    float v[44];
#pragma unroll
    for (int q = 0; q < 44; ++q) v[q] = s_row[r0 + q][c];
#pragma unroll
    for (int o = 0; o < 24; ++o) {
      float acc = __fmaf_rn(v[o + AT2_R], __uint_as_float(c_gauss21[10]), 0.0f);
#pragma unroll
      for (int i = 1; i <= 10; ++i)
        acc = __fmaf_rn(__fadd_rn(v[o + AT2_R + i], v[o + AT2_R - i]), __uint_as_float(c_gauss21[10 + i]), acc);
      int mean = __float_as_int(__fadd_rn(acc, 12582912.0f)) - 0x4b400000;   // cvRound (half to even), acc in [0, 256)
      mean = imin(imax(mean, 0), 255);
      uint8_t *ctr = &s_src[r0 + o + AT2_R][c + AT2_HX];
      const int sv = *ctr;
      *ctr = (sv - mean > -10) ? 255 : 0;
    }
  }
  __syncthreads();
  for (int p = t; p < AT2_TH * (AT2_TW / 4); p += 256) {
    const int r = p >> 4, k = p & 15;
    const int y = y0 + r, x = x0 + 4 * k;
    if (y >= H || x >= W) continue;
    const uint32_t w4 = *reinterpret_cast<const uint32_t *>(&s_src[r + AT2_R][AT2_HX + 4 * k]);
    uint8_t *o = dst + ((size_t)img * H + y) * W + x;
    if (aligned && x + 3 < W) {
      *reinterpret_cast<uint32_t *>(o) = w4;
    } else {
      for (int b = 0; b < 4 && x + b < W; ++b) o[b] = (uint8_t)(w4 >> (8 * b));
    }
  }
}

// ───────────────────────── A.5 deskew ─────────────────────────
// (a) per-row extents of dark (< 128) pixels, one warp per row, 16 pixels per 16-byte load (W % 16 == 0, aligned rows)
__global__ void __launch_bounds__(256)
dark_extents16_kernel(const uint8_t *__restrict__ src, int32_t *__restrict__ ext, int W, int C, int n_rows_total) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows_total) return;
  const uint4 *r = reinterpret_cast<const uint4 *>(src + (size_t)row * W * C);
  int cnt = 0, mn = W, mx = -1;
  const int nchunk = W >> 4;
  for (int ch = lane; ch < nchunk; ch += 32) {
    uint32_t m = 0;
    if (C == 3) {
      union { uint4 v[3]; uint32_t w[12]; } in;
      in.v[0] = __ldg(r + 3 * ch);
      in.v[1] = __ldg(r + 3 * ch + 1);
      in.v[2] = __ldg(r + 3 * ch + 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t g4 = gray4_from_rgb12(in.w[3 * q], in.w[3 * q + 1], in.w[3 * q + 2]);
        // dark <=> bit 7 clear
        const uint32_t dk = ~g4 & 0x80808080u;
        m |= (((dk >> 7) & 1u) | ((dk >> 14) & 2u) | ((dk >> 21) & 4u) | ((dk >> 28) & 8u)) << (4 * q);
      }
    } else {
      const uint4 v = __ldg(r + ch);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t dk = ~w[q] & 0x80808080u;
        m |= (((dk >> 7) & 1u) | ((dk >> 14) & 2u) | ((dk >> 21) & 4u) | ((dk >> 28) & 8u)) << (4 * q);
      }
    }
    if (m) {
      cnt += __popc(m);
      mn = imin(mn, 16 * ch + __ffs((int)m) - 1);
      mx = imax(mx, 16 * ch + 31 - __clz((int)m));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = imin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = imax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
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

// Andrew's monotone chain over n >= 2 lexicographically sorted distinct points pts[2i], pts[2i+1]; strict hull
// vertices into st (room for 2n + 2 points): st[0 .. *lower) is the lower chain first -> last, st[*lower .. k) the upper
// chain back towards the first point (not repeated).  Exactly the loops the one-thread kernel ran.
__device__ __forceinline__ int chain_hull(const int32_t *pts, int n, int32_t *st, int *lower) {
  // the two topmost stack entries are mirrored in registers (a below b): a push never waits for a shared-memory load,
  // a pop reloads one entry
  int k = 0, ax = 0, ay = 0, bx = 0, by = 0;
  for (int i = 0; i < n; ++i) {
    const int qx = pts[2 * i], qy = pts[2 * i + 1];
    while (k >= 2 && cross3(ax, ay, bx, by, qx, qy) <= 0) {
      --k;
      bx = ax; by = ay;
      if (k >= 2) { ax = st[2 * (k - 2)]; ay = st[2 * (k - 2) + 1]; }
    }
    st[2 * k] = qx;
    st[2 * k + 1] = qy;
    ++k;
    ax = bx; ay = by; bx = qx; by = qy;
  }
  *lower = k;
  const int lo = k + 1;
  for (int i = n - 2; i >= 0; --i) {
    const int qx = pts[2 * i], qy = pts[2 * i + 1];
    while (k >= lo && cross3(ax, ay, bx, by, qx, qy) <= 0) {      // k >= lo >= 3: two entries stay below the top
      --k;
      bx = ax; by = ay;
      ax = st[2 * (k - 2)]; ay = st[2 * (k - 2) + 1];
    }
    st[2 * k] = qx;
    st[2 * k + 1] = qy;
    ++k;
    ax = bx; ay = by; bx = qx; by = qy;
  }
  return k - 1;  // last point equals the first
}

// The strict hull vertices of a group of sorted points, again in sorted order, written over the group's own list
// (a vertex of the whole set's hull is a vertex of its group's hull, so nothing the final chain needs is lost).
__device__ __forceinline__ int hull_survivors(int32_t *pts, int n, int32_t *st) {
  if (n <= 2) return n;
  int a;
  const int k = chain_hull(pts, n, st, &a);
  int i = 0, j = k - 1, m = 0;
  while (i < a || j >= a) {
    bool take_i;
    if (j < a) take_i = true;
    else if (i >= a) take_i = false;
    else take_i = st[2 * i] < st[2 * j] || (st[2 * i] == st[2 * j] && st[2 * i + 1] < st[2 * j + 1]);
    const int s = take_i ? i++ : j--;
    pts[2 * m] = st[2 * s];
    pts[2 * m + 1] = st[2 * s + 1];
    ++m;
  }
  return m;
}

// cv::minAreaRect (OpenCV 4.13) restated on a hull in monotone-chain order: cv2.convexHull(clockwise=false) vertex order
// (start at the vertex with the largest x, ties -> largest y), float32 rotating calipers with every operation rounded
// separately, the advancing caliper chosen by exact cross products (firstVecIsRight), `area <= minarea` keeps the LAST
// minimum, quarter-turn normalisation of the side vector, degrees in double; then the reference's angle rule
// (tools.py:561-565) and cv2.getRotationMatrix2D about (W // 2, H // 2).  One thread; the hull has a few dozen vertices.
__device__ __forceinline__ void deskew_calipers(const int32_t *hull, int nh, int H, int W, double *out_angle, double *M) {
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
  float bA = 1.f, bB = 0.f, bW = 0.f;
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
    if (area <= minarea) { minarea = area; bA = base_a; bW = width; bB = base_b; }
  }
  // side vector out[1] = (A1 * width, B1 * width), turned by exact quarter turns into [-pi/2, 0)
  const double PI = 3.14159265358979323846;
  double x = (double)__fmul_rn(bA, bW), y = (double)__fmul_rn(bB, bW);
  double r = atan2(y, x);
  for (int it = 0; it < 4 && r >= 0.0; ++it) { const double tt = x; x = y; y = -tt; r = atan2(y, x); }
  for (int it = 0; it < 4 && r < -PI / 2; ++it) { const double tt = x; x = -y; y = tt; r = atan2(y, x); }
  const float ang = (float)(r * 180.0 / PI);
  double angle = (double)ang;
  if (angle < -45.0) angle = -(90.0 + angle);
  else angle = -angle;
  *out_angle = angle;
  const double th = angle * (PI / 180.0);
  const double al = cos(th), be = sin(th);
  const double ccx = (double)(W / 2), ccy = (double)(H / 2);
  M[0] = al;
  M[1] = be;
  M[2] = __dsub_rn(__dmul_rn(__dsub_rn(1.0, al), ccx), __dmul_rn(be, ccy));
  M[3] = -be;
  M[4] = al;
  M[5] = __dadd_rn(__dmul_rn(be, ccx), __dmul_rn(__dsub_rn(1.0, al), ccy));
}

// (b) one CTA per image, everything in shared memory.  The reference hands (row, col) to minAreaRect as (x, y)
// (tools.py:557-560); the <= 2H extent points are already sorted (row ascending, min col then max col).  The hull scan was
// ONE thread walking all of them (0.5 ms per page, whatever the batch); here 64 threads build the hulls of 64 row
// groups, 8 threads merge eight groups each, and one thread merges those eight -- every level runs the same chain over
// fewer points, so the final vertex sequence is the one the single scan produced.
// Dynamic shared memory (int32): extents [3H] | A [64][2R][2] | B [8][16R][2] | S [512R + 256], R = ceil(H / 64).
static inline size_t deskew_par_smem_bytes(int H) {
  const size_t R = ((size_t)H + 63) / 64;
  return ((size_t)3 * H + 1024 * R + 256) * sizeof(int32_t);
}

__global__ void __launch_bounds__(256)
deskew_angle_par_kernel(const int32_t *__restrict__ ext, int H, int W, double *__restrict__ out_angle,
                        double *__restrict__ out_M) {
  OCRB_DYN_SMEM(int32_t, dk_smem);
  __shared__ int s_cnt[64], s_np[64], s_dark[64], s_cnt2[8];
  const int t = threadIdx.x;
  const int img = blockIdx.x;
  const int R = (H + 63) / 64;
  int32_t *se = dk_smem;
  int32_t *A = se + 3 * H;
  int32_t *B = A + 256 * R;
  int32_t *S = B + 256 * R;
  {
    const int32_t *e = ext + (size_t)img * H * 3;
    for (int i = t; i < 3 * H; i += 256) se[i] = e[i];
  }
  __syncthreads();
  if ((t & 3) == 0) {                       // level 0: group g = rows [g R, (g + 1) R)
    const int g = t >> 2;
    int32_t *pts = A + (size_t)g * 4 * R;
    int32_t *st = S + (size_t)g * (8 * R + 4);
    int np = 0, dark = 0;
    const int y1 = imin((g + 1) * R, H);
    for (int y = g * R; y < y1; ++y) {
      const int c = se[y * 3];
      dark += c;
      if (c > 0) {
        const int mn = se[y * 3 + 1], mx = se[y * 3 + 2];
        pts[2 * np] = y; pts[2 * np + 1] = mn; ++np;
        if (mx != mn) { pts[2 * np] = y; pts[2 * np + 1] = mx; ++np; }
      }
    }
    s_np[g] = np;
    s_dark[g] = dark;
    s_cnt[g] = hull_survivors(pts, np, st);
  }
  __syncthreads();
  if ((t & 31) == 0) {                      // level 1: group j = level-0 groups 8j .. 8j + 7
    const int j = t >> 5;
    int32_t *pts = B + (size_t)j * 32 * R;
    int32_t *st = S + (size_t)j * (64 * R + 4);
    int n = 0;
    for (int g = 8 * j; g < 8 * j + 8; ++g) {
      const int32_t *q = A + (size_t)g * 4 * R;
      for (int i = 0; i < s_cnt[g]; ++i) { pts[2 * n] = q[2 * i]; pts[2 * n + 1] = q[2 * i + 1]; ++n; }
    }
    s_cnt2[j] = hull_survivors(pts, n, st);
  }
  __syncthreads();
  if (t == 0) {
    int total = 0, np = 0;
    for (int g = 0; g < 64; ++g) { total += s_dark[g]; np += s_np[g]; }
    int n = 0;
    for (int j = 0; j < 8; ++j) {
      const int32_t *q = B + (size_t)j * 32 * R;
      for (int i = 0; i < s_cnt2[j]; ++i) { A[2 * n] = q[2 * i]; A[2 * n + 1] = q[2 * i + 1]; ++n; }
    }
    int k = 0;
    if (total > 100 && np >= 3) {
      int a;
      k = chain_hull(A, n, S, &a);
    }
    if (total <= 100 || k < 3) {
      // <= 100 dark pixels: unchanged image (tools.py:558-559).  Degenerate hulls (all dark pixels collinear) are
      // reported as unsupported by NaN as well.
      out_angle[img] = nan("");
      for (int q = 0; q < 6; ++q) out_M[img * 6 + q] = nan("");
    } else {
      deskew_calipers(S, k, H, W, out_angle + img, out_M + img * 6);
    }
  }
}

// ───────────────────────── cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE) ─────────────────────────
// OpenCV fixed-point bicubic table: int16 [1024][16], built on the host.  Index (fx*32 + fy), the transpose of OpenCV's: along an
// output row of a slightly rotated page fx stays put while fy sweeps its 32 values, so the 32 lanes of a warp read 32
// CONSECUTIVE 32-byte rows (8 cache lines) instead of rows 1 KB apart (32 lines) -- the weight loads were two thirds of the
// kernel's L1 wavefronts.
__device__ int16_t g_cubic_itab[1024 * 16];

// Host: OpenCV's interpolateCubic + initInterTab2D (fp32, unfused), incl. the ksize/2 quirk.
static inline void build_cubic_itab(int16_t *out) {
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
        for (int kx = 0; kx < 4; ++kx) out[(fx * 32 + fy) * 16 + ky * 4 + kx] = (int16_t)iw[ky][kx];
    }
}

// One CTA = 256 consecutive pixels of one row.  The inverse matrix and the row's fixed-point origin are computed once per
// CTA (they were ~25 fp64 operations per PIXEL); the 4 x 4 taps of the 3 channels go through dp2a: the table row already
// holds the 16 weights as int16 pairs (taps kx, kx + 1 of one source row), a PRMT puts the two source bytes of a channel
// next to each other, and one IDP.2A adds both products -- 4 PRMT + 6 IDP.2A per source row instead of 12 byte extractions,
// 4 weight extractions and 12 IMAD.  Integer sums, so the order of the taps does not matter.
template <int C, int OCC = 4>
__global__ void __launch_bounds__(256, OCC)
warp_affine_cubic_dp2a_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W,
                              const double *__restrict__ Mall) {
  __shared__ double s_m[2];
  __shared__ int s_o[3];
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (threadIdx.x == 0) {
    const double *Mf = Mall + img * 6;
    double m0 = Mf[0], m1 = Mf[1], m2 = Mf[2], m3 = Mf[3], m4 = Mf[4], m5 = Mf[5];
    s_o[2] = (m0 != m0);  // NaN: leave the image unchanged
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
    s_m[0] = m0;
    s_m[1] = m3;
    s_o[0] = s_o[2] ? 0 : __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)y), b1), 1024.0)) + 16;
    s_o[1] = s_o[2] ? 0 : __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)y), b2), 1024.0)) + 16;
  }
  __syncthreads();
  if (x >= W) return;
  const size_t img_off = (size_t)img * H * W * C;
  const uint8_t *im = src + img_off;
  uint8_t *o = dst + img_off + ((size_t)y * W + x) * C;
  // 32-bit byte offsets inside the image (the host checks H * W * C < 2^31), from the image base rounded down to a word
  const uint32_t basemis = (uint32_t)(reinterpret_cast<uintptr_t>(im) & 3);
  const uint32_t *imw = reinterpret_cast<const uint32_t *>(im - basemis);
  const uint32_t rowb = (uint32_t)W * C;
  if (s_o[2]) {
    for (int c = 0; c < C; ++c) o[c] = im[((size_t)y * W + x) * C + c];
    return;
  }
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(s_m[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(s_m[1], (double)x), 1024.0));
  const int X = (s_o[0] + adelta) >> 5, Y = (s_o[1] + bdelta) >> 5;
  int sx = X >> 5, sy = Y >> 5;
  sx = imin(imax(sx, -32768), 32767);
  sy = imin(imax(sy, -32768), 32767);
  // the 16 fixed-point weights of this sub-pixel position: two 16-byte loads (the table row is 32-byte aligned);
  // wpk[i] = weights 2i (low half), 2i + 1 (high half)
  const uint4 *wt4 = reinterpret_cast<const uint4 *>(g_cubic_itab + (((X & 31) * 32 + (Y & 31)) << 4));
  const uint4 wa = wt4[0], wb = wt4[1];
  const uint32_t wpk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
  int ys[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ys[k] = imin(imax(sy - 1 + k, 0), H - 1);
  // Interior columns (no clamping in x, and room for the aligned over-read): the 4 taps x C channels of a source row
  // are 4C contiguous bytes at an arbitrary alignment -- aligned 32-bit words funnel-shifted into place.
  const int margin = (C == 3) ? 1 : 3;
  if (sx - 1 >= 0 && sx + 2 + margin <= W - 1) {
    int acc0 = 0, acc1 = 0, acc2 = 0;
    const uint32_t colb = basemis + (uint32_t)(sx - 1) * C;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const uint32_t off = (uint32_t)ys[ky] * rowb + colb;
      const uint32_t *wp = imw + (off >> 2);
      const uint32_t sh = (off & 3u) * 8;
      const uint32_t w01 = wpk[2 * ky], w23 = wpk[2 * ky + 1];
      if (C == 3) {
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
        // byte j = 3 kx + c of (b0 b1 b2)
        const uint32_t p0 = __byte_perm(b0, b1, 0x4130);   // [j0 j3 | j1 j4]: channels 0 and 1, taps 0 1
        const uint32_t p1 = __byte_perm(b0, b1, 0x0052);   // [j2 j5 | . .]:   channel 2, taps 0 1
        const uint32_t p2 = __byte_perm(b1, b2, 0x6352);   // [j6 j9 | j7 j10]: channels 0 and 1, taps 2 3
        const uint32_t p3 = __byte_perm(b1, b2, 0x0074);   // [j8 j11 | . .]:  channel 2, taps 2 3
        acc0 = dp2a_lo_s16u8(w01, p0, acc0);
        acc1 = dp2a_hi_s16u8(w01, p0, acc1);
        acc2 = dp2a_lo_s16u8(w01, p1, acc2);
        acc0 = dp2a_lo_s16u8(w23, p2, acc0);
        acc1 = dp2a_hi_s16u8(w23, p2, acc1);
        acc2 = dp2a_lo_s16u8(w23, p3, acc2);
      } else {
        const uint32_t w0 = wp[0], w1 = wp[1];
        const uint32_t b0 = __funnelshift_r(w0, w1, sh);
        acc0 = dp2a_lo_s16u8(w01, b0, acc0);
        acc0 = dp2a_hi_s16u8(w23, b0, acc0);
      }
    }
    o[0] = (uint8_t)imin(imax((acc0 + 16384) >> 15, 0), 255);
    if (C == 3) {
      o[1] = (uint8_t)imin(imax((acc1 + 16384) >> 15, 0), 255);
      o[2] = (uint8_t)imin(imax((acc2 + 16384) >> 15, 0), 255);
    }
    return;
  }
  int xs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xs[k] = imin(imax(sx - 1 + k, 0), W - 1);
  for (int c = 0; c < C; ++c) {
    int acc = 0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int i = ky * 4 + kx;
        const int wg = (int)(short)((i & 1) ? (wpk[i >> 1] >> 16) : (wpk[i >> 1] & 0xffffu));
        acc += wg * (int)im[((size_t)ys[ky] * W + xs[kx]) * C + c];
      }
    const int v = (acc + 16384) >> 15;
    o[c] = (uint8_t)imin(imax(v, 0), 255);
  }
}

}  // namespace ocrb
