// TEST INFRASTRUCTURE: the emulated kernels (tests/emu/cuda_emu.h) as a standalone program to be built with
//   g++ -fsanitize=address,undefined   -> out-of-bounds accesses to global buffers (exact-size heap blocks) and to
//                                         __shared__ arrays (statics carry red zones), misaligned vector accesses
//   g++ -fsanitize=thread              -> data races between the emulated CUDA threads (a missing __syncthreads, two
//                                         threads writing one shared-memory word)
// compute-sanitizer is closed on the GPU pool this repository is measured on (profiles/r02_sanitize.md); this is the
// host-side stand-in for memcheck / racecheck of the kernels that can be emulated (image_fast.cuh, textops_kernels.cuh).
// Results are not compared here (tests/test_emu_*_kernels.py do that); the sanitizers are the check.
#include "cuda_emu.h"
#include "../../handwritten-ocr_b200/csrc/image_fast.cuh"
#include "../../handwritten-ocr_b200/csrc/image_general.cuh"
#include "../../handwritten-ocr_b200/csrc/textops_kernels.cuh"
#include "../../handwritten-ocr_b200/csrc/denoise_kernels.cuh"
#include "../../handwritten-ocr_b200/csrc/resize_kernels.cuh"
#include "../../handwritten-ocr_b200/csrc/dense_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

using namespace ocrb;

static std::mt19937 rng(12345);
static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// exact-size, 64-byte aligned heap block (ASan puts red zones around it)
template <class T>
struct Buf {
  T *p;
  size_t n;
  explicit Buf(size_t n_, int fill = -1) : n(n_) {
    void *q = nullptr;
    if (posix_memalign(&q, 64, n * sizeof(T) ? n * sizeof(T) : 1) != 0) std::abort();
    p = static_cast<T *>(q);
    for (size_t i = 0; i < n; ++i) p[i] = fill >= 0 ? (T)fill : (T)(rng() & 0xff);
  }
  ~Buf() { free(p); }
};

static void paper(uint8_t *img, size_t n) {
  for (size_t i = 0; i < n; ++i) img[i] = (rng() % 23 == 0) ? (uint8_t)(rng() % 90) : (uint8_t)(230 + rng() % 20);
}

static void run_image(int n, int H, int W, int C) {
  const size_t px = (size_t)n * H * W;
  Buf<uint8_t> src(px * C), dst(px * C, 0), gray(px, 0), out1(px, 0), lut((size_t)n * 64 * 256, 0);
  paper(src.p, px * C);
  // sharpen
  if ((W * C) % 16 == 0 && H >= 2) {
    const int nv = W * C / 16;
    const dim3 grid(cdivu(nv, 64), cdivu(H, 4), n);
    if (C == 3) emu::launch(grid, dim3(256), 0, [&] { sharpen_vec16_kernel<3>(src.p, dst.p, H, nv); });
    else emu::launch(grid, dim3(256), 0, [&] { sharpen_vec16_kernel<1>(src.p, dst.p, H, nv); });
  }
  // high_contrast
  {
    int We = W, He = H;
    if (!(W % 8 == 0 && H % 8 == 0)) { We = W + (8 - W % 8); He = H + (8 - H % 8); }
    const int tw = We / 8, th = He / 8;
    if (tw >= 1 && th >= 1 && tw <= W && th <= H) {
      const int area = tw * th;
      int clip = (int)(3.0 * area / 256.0);
      if (clip < 1) clip = 1;
      const float ls = 255.0f / (float)area;
      const uint8_t *g = (C == 3) ? gray.p : src.p;
      const int vec_ok = (W % 16 == 0) && (tw % 16 == 0);
      if (C == 3) emu::launch(dim3(64, n), dim3(256), 0, [&] { clahe_hist_lut_kernel<3>(src.p, gray.p, lut.p, H, W, tw, th, clip, ls, vec_ok); });
      else emu::launch(dim3(64, n), dim3(256), 0, [&] { clahe_hist_lut_kernel<1>(src.p, nullptr, lut.p, H, W, tw, th, clip, ls, vec_ok); });
      const float itw = 1.0f / (float)tw, ith = 1.0f / (float)th;
      ClaheCells cells;
      if (clahe_cells_host(H, W, itw, ith, &cells))
        emu::launch(dim3(81, n), dim3(256), 0, [&] { clahe_apply_cells_kernel(g, out1.p, lut.p, H, W, itw, ith, cells); });
    }
  }
  // binarize
  {
    const int aligned = (W % 4 == 0);
    const dim3 grid(cdivu(W, AT2_TW), cdivu(H, AT2_TH), n);
    if (C == 3) emu::launch(grid, dim3(256), 0, [&] { adaptive_thresh_tile_kernel<3>(src.p, out1.p, H, W, aligned); });
    else emu::launch(grid, dim3(256), 0, [&] { adaptive_thresh_tile_kernel<1>(src.p, out1.p, H, W, aligned); });
  }
  // deskew: extents, hull tree + calipers, warp
  if (W % 16 == 0) {
    Buf<int32_t> ext((size_t)n * H * 3, 0);
    Buf<double> angle(n, 0), M((size_t)n * 6, 0);
    const int rows = n * H;
    emu::launch(dim3(cdivu(rows, 8)), dim3(256), 0, [&] { dark_extents16_kernel(src.p, ext.p, W, C, rows); });
    emu::launch(dim3(n), dim3(256), deskew_par_smem_bytes(H), [&] { deskew_angle_par_kernel(ext.p, H, W, angle.p, M.p); });
    static bool built = false;
    if (!built) { build_cubic_itab(g_cubic_itab); built = true; }
    const dim3 grid(cdivu(W, 256), H, n);
    if (C == 3) emu::launch(grid, dim3(256), 0, [&] { warp_affine_cubic_dp2a_kernel<3>(src.p, dst.p, H, W, M.p); });
    else emu::launch(grid, dim3(256), 0, [&] { warp_affine_cubic_dp2a_kernel<1>(src.p, dst.p, H, W, M.p); });
  }
  std::printf("image n=%d %dx%d C=%d ok\n", n, H, W, C);
}

// the general-shape kernels on sizes the fast paths refuse
static void run_general(int H, int W, int C) {
  const size_t px = (size_t)H * W;
  Buf<uint8_t> src(px * C), dst(px * C, 0), gray(px, 0), out1(px, 0), tmp(px, 0), lut(64 * 256, 0);
  paper(src.p, px * C);
  if (C == 3) {
    emu::launch(dim3(cdivu((long long)((px + 15) / 16), 256)), dim3(256), 0, [&] { rgb2gray_kernel(src.p, gray.p, px); });
  } else {
    std::copy(src.p, src.p + px, gray.p);
  }
  emu::launch(dim3(cdivu((long long)W * C, 256), H, 1), dim3(256), 0, [&] { sharpen_kernel(src.p, dst.p, H, W, C); });
  if ((W * C) % 4 == 0 && W >= 4)
    emu::launch(dim3(cdivu((long long)W * C / 4, 256), H, 1), dim3(256), 0, [&] { sharpen4_kernel(src.p, dst.p, H, W, C); });
  int We = W, He = H;
  if (!(W % 8 == 0 && H % 8 == 0)) { We = W + (8 - W % 8); He = H + (8 - H % 8); }
  const int tw = We / 8, th = He / 8;
  if (tw >= 1 && th >= 1 && tw <= W && th <= H) {
    const int area = tw * th;
    int clip = (int)(3.0 * area / 256.0);
    if (clip < 1) clip = 1;
    const float ls = 255.0f / (float)area, itw = 1.0f / (float)tw, ith = 1.0f / (float)th;
    emu::launch(dim3(64, 1), dim3(256), 0, [&] { clahe_hist_lut_kernel<1>(gray.p, nullptr, lut.p, H, W, tw, th, clip, ls, 0); });
    emu::launch(dim3(cdivu(W, 256), H, 1), dim3(256), 0, [&] { clahe_apply_kernel(gray.p, out1.p, lut.p, H, W, itw, ith); });
    if (W % 4 == 0)
      emu::launch(dim3(cdivu(W / 4, 256), H, 1), dim3(256), 0, [&] { clahe_apply4_kernel(gray.p, out1.p, lut.p, H, W, itw, ith); });
  }
  {
    Buf<int32_t> ext((size_t)H * 3, 0), hull((size_t)(4 * H + 8) * 2, 0), nz(1, 0);
    Buf<double> angle(1, 0), M(6, 0);
    emu::launch(dim3(cdivu(H, 8)), dim3(256), 0, [&] { dark_extents_kernel(src.p, ext.p, H, W, C, H); });
    emu::launch(dim3(1), dim3(32), 0, [&] { deskew_angle_seq_kernel(ext.p, H, W, angle.p, M.p, hull.p); });
    if (W >= 4) {
      emu::launch(dim3(cdivu(W, 256), H, 1), dim3(256), 0, [&] { rl_thresh_kernel(src.p, out1.p, H, W, C); });
      emu::launch(dim3(H), dim3(256), (2 * (size_t)W + 1) * sizeof(int), [&] { rl_open_row_kernel(out1.p, tmp.p, W, W / 4); });
      emu::launch(dim3(cdivu(W, 256), H, 1), dim3(256), 0, [&] { rl_dilate_v_kernel(tmp.p, out1.p, nz.p, H, W); });
    }
  }
  std::printf("general %dx%d C=%d ok\n", H, W, C);
}

static void run_resize(int H, int W, int C, int oh, int ow) {
  const size_t n = (size_t)H * W * C;
  Buf<uint8_t> src(n), tmp((size_t)H * ow * C, 0), dst((size_t)oh * ow * C, 0);
  AxisWeights tx, ty;
  compute_axis_weights(W, ow, &tx);
  compute_axis_weights(H, oh, &ty);
  Buf<int32_t> xmin(tx.xmin.size(), 0), xsize(tx.xsize.size(), 0), ymin(ty.xmin.size(), 0), ysize(ty.xsize.size(), 0);
  Buf<int16_t> wx(tx.w.size(), 0), wy(ty.w.size(), 0);
  std::copy(tx.xmin.begin(), tx.xmin.end(), xmin.p);
  std::copy(tx.xsize.begin(), tx.xsize.end(), xsize.p);
  std::copy(tx.w.begin(), tx.w.end(), wx.p);
  std::copy(ty.xmin.begin(), ty.xmin.end(), ymin.p);
  std::copy(ty.xsize.begin(), ty.xsize.end(), ysize.p);
  std::copy(ty.w.begin(), ty.w.end(), wy.p);
  emu::launch(dim3(cdivu((long long)ow * C, 256), H), dim3(256), 0,
              [&] { resize_h_kernel(src.p, tmp.p, H, W, C, ow, xmin.p, xsize.p, wx.p, tx.kmax, tx.prec); });
  emu::launch(dim3(cdivu((long long)ow * C, 256), oh, 1), dim3(256), 0,
              [&] { resize_v_kernel(tmp.p, dst.p, H, ow * C, oh, ymin.p, ysize.p, wy.p, ty.kmax, ty.prec); });
  if (oh % 28 == 0 && ow % 28 == 0) {
    const int gh = oh / 14, gw = ow / 14;
    const long long total = (long long)gh * gw * 3 * 14 * 14;
    Buf<float> pv((size_t)gh * gw * 1176, 0);
    emu::launch(dim3(cdivu(total, 256)), dim3(256), 0, [&] {
      normalize_patchify_kernel<float>(dst.p, pv.p, oh, ow, C, gh, gw, nullptr, total, 122.77f, 116.75f, 104.09f, 68.5f, 66.63f, 70.32f);
    });
  }
  std::printf("resize %dx%d C=%d -> %dx%d ok\n", H, W, C, oh, ow);
}

static void run_dense() {
  const int rows = 11, dim = 256;
  Buf<uint16_t> x((size_t)rows * dim), w(dim), y((size_t)rows * dim, 0);
  for (size_t i = 0; i < x.n; ++i) x.p[i] = (uint16_t)(0x3f00 + (rng() & 0xff));     // bf16 values around 0.5 .. 1
  for (size_t i = 0; i < w.n; ++i) w.p[i] = (uint16_t)(0x3f80 + (rng() & 0x3f));
  emu::launch(dim3(rows), dim3(256), 0, [&] { rmsnorm_kernel((const bf16 *)x.p, dim, (const bf16 *)w.p, (bf16 *)y.p, dim, dim, 1e-6f); });
  emu::launch(dim3(cdivu(rows, 8)), dim3(256), 0, [&] { rmsnorm_warp_kernel((const bf16 *)x.p, dim, (const bf16 *)w.p, (bf16 *)y.p, dim, rows, dim, 1e-6f); });
  {
    const int S = 7, heads = 2, hd = 32;
    Buf<uint16_t> qkv((size_t)S * 3 * heads * hd);
    Buf<float> c((size_t)S * hd, 0), sn((size_t)S * hd, 0);            // tables are [S, hd]: both halves of a head
    emu::launch(dim3(cdivu(S * 2 * heads * (hd / 16), 256)), dim3(256), 0, [&] { rope_vision_vec_kernel((bf16 *)qkv.p, S, heads, hd, c.p, sn.p); });
    emu::launch(dim3(cdivu((long long)S * 2 * heads * (hd / 2), 256)), dim3(256), 0, [&] { rope_vision_kernel((bf16 *)qkv.p, S, heads, hd, c.p, sn.p); });
  }
  {
    const int T = 5, nq = 4, nkv = 2, hd = 32;
    Buf<uint16_t> q((size_t)T * nq * hd), k((size_t)T * nkv * hd), c((size_t)T * hd, 0x3f80), sn((size_t)T * hd, 0);
    emu::launch(dim3(cdivu(T * (nq + nkv) * (hd / 16), 256)), dim3(256), 0,
                [&] { rope_text_vec_kernel((bf16 *)q.p, nq * hd, (bf16 *)k.p, nkv * hd, T, nq, nkv, hd, (const bf16 *)c.p, (const bf16 *)sn.p); });
    emu::launch(dim3(cdivu((long long)T * (nq + nkv) * (hd / 2), 256)), dim3(256), 0,
                [&] { rope_text_kernel((bf16 *)q.p, nq * hd, (bf16 *)k.p, nkv * hd, T, nq, nkv, hd, (const bf16 *)c.p, (const bf16 *)sn.p); });
    Buf<int32_t> ctx(3, 40), delta(3, 2);
    Buf<float> invf(hd / 2, 0);
    Buf<uint16_t> ct((size_t)3 * hd, 0), st((size_t)3 * hd, 0);
    emu::launch(dim3(cdivu(3 * (hd / 2), 128)), dim3(128), 0,
                [&] { decode_rope_table_kernel(ctx.p, delta.p, invf.p, 3, hd, (bf16 *)ct.p, (bf16 *)st.p); });
    // paged KV write: sequences of 5 and 3 tokens... T = 5 here: one sequence of 5 tokens over two pages of 4
    Buf<int32_t> bt(2, 0), cu(2, 0);
    bt.p[0] = 1; bt.p[1] = 0; cu.p[0] = 0; cu.p[1] = T;
    Buf<uint16_t> kc((size_t)2 * nkv * 4 * hd, 0), vc((size_t)2 * nkv * 4 * hd, 0);
    emu::launch(dim3(cdivu((long long)T * (nkv * hd / 8), 256)), dim3(256), 0, [&] {
      kv_write_prefill_kernel((const bf16 *)k.p, nkv * hd, (const bf16 *)k.p, nkv * hd, (bf16 *)kc.p, (bf16 *)vc.p, bt.p, 2, cu.p, 1, T, 4,
                              nkv * hd / 8, hd / 8);
    });
  }
  {
    const int B = 3, V = 1003, ld = 1008, max_new = 4;
    Buf<uint16_t> lg((size_t)B * ld);
    for (size_t i = 0; i < lg.n; ++i) lg.p[i] = (uint16_t)(0x3c00 + (rng() & 0x3ff));
    Buf<int32_t> out((size_t)B * max_new, 0), nxt(B, 0), fin(B, 0), ctx(B, 5), step(1, 1);
    emu::launch(dim3(B), dim3(512), 0, [&] { argmax_step_kernel((const bf16 *)lg.p, ld, V, 7, 7, max_new, out.p, nxt.p, fin.p, ctx.p, step.p, 1); });
    Buf<int32_t> si(2, 0), di(2, 0);
    si.p[0] = 2; si.p[1] = 0; di.p[0] = 0; di.p[1] = 1;
    Buf<uint16_t> dst((size_t)2 * 64, 0);
    emu::launch(dim3(1), dim3(256), 0, [&] { rows_copy_kernel((const bf16 *)lg.p, ld, si.p, (bf16 *)dst.p, 64, di.p, 2, 8); });
    emu::launch(dim3(1), dim3(256), 0, [&] { residual_add_kernel((bf16 *)dst.p, 64, (const bf16 *)lg.p, ld, 2, 8); });
  }
  std::printf("dense ok\n");
}

static void run_text() {
  const int lens[][2] = {{1, 1}, {33, 31}, {64, 1}, {0, 5}, {70, 100}, {200, 255}};
  const int np = 6;
  std::vector<int32_t> a, b, oa{0}, ob{0};
  int maxb = 0;
  for (int k = 0; k < np; ++k) {
    for (int i = 0; i < lens[k][0]; ++i) a.push_back((int32_t)(rng() % 5));
    for (int i = 0; i < lens[k][1]; ++i) b.push_back((int32_t)(rng() % 5));
    oa.push_back((int32_t)a.size());
    ob.push_back((int32_t)b.size());
    if (lens[k][1] > maxb) maxb = lens[k][1];
  }
  Buf<int32_t> A(a.size() + 1, 0), B(b.size() + 1, 0), OA(oa.size(), 0), OB(ob.size(), 0), out(np, 0);
  std::copy(a.begin(), a.end(), A.p);
  std::copy(b.begin(), b.end(), B.p);
  std::copy(oa.begin(), oa.end(), OA.p);
  std::copy(ob.begin(), ob.end(), OB.p);
  emu::launch(dim3(np), dim3(64), 0, [&] { levenshtein_kernel<4, 64>(A.p, OA.p, B.p, OB.p, out.p); });
  emu::launch(dim3(np), dim3(128), 0, [&] { levenshtein_kernel<8, 128>(A.p, OA.p, B.p, OB.p, out.p); });
  // LCS: backbone = a sequences, versions = b sequences
  std::vector<int64_t> wo(np);
  size_t tot = 0;
  int maxbb = 0;
  for (int k = 0; k < np; ++k) {
    wo[k] = (int64_t)tot;
    tot += (size_t)lens[k][0] * lens[k][1];
    if (lens[k][0] > maxbb) maxbb = lens[k][0];
  }
  Buf<uint8_t> work(tot + 1, 0);
  Buf<int64_t> WO(np, 0);
  std::copy(wo.begin(), wo.end(), WO.p);
  Buf<int32_t> aligned(a.size() + 1, 0);
  const int diag_stride = (maxbb + 2 + 7) & ~7;
  emu::launch(dim3(np), dim3(LCS_THREADS), (size_t)3 * diag_stride * sizeof(uint16_t),
              [&] { lcs_align_kernel(A.p, OA.p, B.p, OB.p, aligned.p, work.p, WO.p, diag_stride); });
  std::printf("text ok\n");
}

static void run_denoise(int H, int W, int C) {
  const DenoiseHostTables &t = host_tables();
  const size_t npix = (size_t)H * W;
  Buf<uint8_t> src(npix * C), dst(npix * C, 0), ws(npix * 6, 0);
  Buf<uint16_t> w0(NLM_LUT, 0), w1(NLM_LUT, 0), cb(3072, 0);
  Buf<int2> yf(256, 0);
  std::copy(t.w[0], t.w[0] + NLM_LUT, w0.p);
  std::copy(t.w[1], t.w[1] + NLM_LUT, w1.p);
  std::copy(t.cbrt_tab, t.cbrt_tab + 3072, cb.p);
  std::memcpy(yf.p, t.yf, sizeof(t.yf));
  paper(src.p, npix * C);
  const dim3 grid(cdivu(W, NLM_COLS), cdivu(H, NLM_ROWS), 1);
  if (C == 1) {
    emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<1>(src.p, dst.p, H, W, w0.p); });
  } else {
    uint8_t *ab0 = ws.p, *ab1 = ws.p + 2 * npix, *L0 = ws.p + 4 * npix, *L1 = ws.p + 5 * npix;
    emu::launch(dim3(cdivu((long long)npix, 256)), dim3(256), 0, [&] { lbgr2lab_kernel(src.p, L0, ab0, npix, cb.p, t.cf); });
    emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<1>(L0, L1, H, W, w0.p); });
    emu::launch(grid, dim3(32 * NLM_NW), 0, [&] { nlm_kernel<2>(ab0, ab1, H, W, w1.p); });
    emu::launch(dim3(cdivu((long long)npix, 256)), dim3(256), 0, [&] { lab2lbgr_kernel(L1, ab1, dst.p, npix, yf.p, t.cf); });
  }
  std::printf("denoise %dx%d C=%d ok\n", H, W, C);
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "denoise") {
    run_denoise(30, 28, 1);    // two tile columns (26 + 2), reflect-101 on every side
    run_denoise(20, 40, 3);    // colored route: Lab, NLM on L and on (a, b), back
    std::printf("emulated kernels: sanitizer run complete\n");
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "dense") {
    run_dense();
    std::printf("emulated kernels: sanitizer run complete\n");
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "resize") {
    run_resize(37, 53, 3, 28, 56);
    run_resize(30, 20, 1, 56, 28);
    std::printf("emulated kernels: sanitizer run complete\n");
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "general") {
    run_general(23, 27, 3);
    run_general(24, 36, 1);
    run_general(9, 13, 1);
    std::printf("emulated kernels: sanitizer run complete\n");
    return 0;
  }
  run_image(2, 64, 128, 3);    // vector histogram path, cell kernel, interior + border tiles
  run_image(1, 96, 256, 1);
  run_image(1, 50, 70, 3);     // reflect-101 tile extension, unaligned widths, scalar paths
  run_image(1, 37, 64, 1);
  run_image(2, 130, 160, 3);   // two tile rows of the threshold, ragged hull groups
  run_text();
  std::printf("emulated kernels: sanitizer run complete\n");
  return 0;
}
