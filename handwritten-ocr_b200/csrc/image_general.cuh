// General-shape versions of the uint8 image kernels (tools.py:503-614): any width, any alignment.  The page-sized fast
// paths are in image_fast.cuh; the host chooses (image.cu).  Kernel definitions only -- no launches -- so that tests/emu can
// compile this file for the host (OCRB_EMU) and run the kernels thread by thread against the oracle and under the host
// sanitizers.  Compiled with -fmad=false; every floating-point operation whose rounding matters is written with an explicit
// _rn intrinsic.
#pragma once
#include "image_fast.cuh"

namespace ocrb {

// ───────────────────────── A.1 RGB -> gray ─────────────────────────
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

// ───────────────────────── A.2 CLAHE ─────────────────────────
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

// (b) fallback for pages too tall for deskew_angle_par_kernel's shared-memory tree (H > ~2600): one thread's scan with
// the candidate points and the hull stack in global memory -- same chain, same calipers.
__global__ void __launch_bounds__(32)
deskew_angle_seq_kernel(const int32_t *__restrict__ ext, int H, int W, double *__restrict__ out_angle,
                        double *__restrict__ out_M, int32_t *__restrict__ hull_ws) {
  const int img = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int32_t *e = ext + (size_t)img * H * 3;
  int32_t *pts = hull_ws + (size_t)img * (4 * H + 8) * 2;  // [2H+4][2] candidate points
  int32_t *hull = pts + (2 * H + 4) * 2;                   // [2H+4][2] hull
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
  int k = 0;
  if (total > 100 && np >= 3) {
    int a;
    k = chain_hull(pts, np, hull, &a);
  }
  if (total <= 100 || k < 3) {
    out_angle[img] = nan("");
    for (int q = 0; q < 6; ++q) out_M[img * 6 + q] = nan("");
  } else {
    deskew_calipers(hull, k, H, W, out_angle + img, out_M + img * 6);
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
  OCRB_DYN_SMEM(int, rl_pre);                   // [W + 1] prefix counts, then [W] flags
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

}  // namespace ocrb
