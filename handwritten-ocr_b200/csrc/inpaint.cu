// remove_lines strategy, second half (tools.py:617): cv2.inpaint(img, mask, 3, cv2.INPAINT_TELEA), bit-exact against
// OpenCV 4.13.0.92 for uint8 pages with 1 or 3 channels.  Compiled with -fmad=false: every float / double step is the
// one OpenCV takes.
//
// Telea's method is a fast march: pixels are painted in the order a priority queue (distance to the mask boundary,
// FIFO among equal distances) hands them out, and every painted pixel feeds the ones painted after it, so the order IS
// the result.  What can run in parallel without changing a bit:
//   * row segments of the mask separated by >= 8 clean rows never read each other's data (everything the march touches
//     lies within 4 rows of the mask), so each segment is marched by its own warp with its own queue -- a ruled page
//     has one segment per ruled line;
//   * inside a segment, the four distance solves of each of the four neighbours of a popped pixel run on 16 lanes, and
//     the <= 49 disc positions of a painted pixel are evaluated by the lanes in parallel; only the float sums are then
//     accumulated in OpenCV's raster order (one lane per running sum), because float addition does not reassociate.
// The queue is a binary heap on the key (float bits of T) << 32 | insertion number, equal in order to OpenCV's sorted
// list; only lane 0 touches it.  A segment's window of the distance field, the flags and the page rows is staged in
// shared memory together with the heap (freshly written global data would come back from L2 at ~600 cycles a load,
// and the march is one long dependent chain); segments too tall for that run from global memory through the same
// generic pointers, and a heap that outgrows its shared part spills into the segment's global slice.
#include "common.cuh"
#include <math.h>

namespace ocrb {

enum : uint8_t { F_KNOWN = 0, F_BAND = 1, F_INSIDE = 2, F_CHANGE = 3, F_SEED = 4, F_SEED_DONE = 5 };
constexpr int INP_MAX_RANGE = 7;
// clean rows between independent segments: the march reads up to range + 1 rows away from the mask and writes up to
// range rows away, so 2 * range + 1 clean rows already separate two segments; one more for margin
static inline __host__ __device__ int inp_gap(int range) { return 2 * range + 2; }
constexpr int INP_CHAINS = 10;        // Ia[3], Jx[3], Jy[3], s

struct InpWs {
  uint64_t *key;   // [n * ne] heap keys
  float *t;        // [n * ne] distance field (extended image: 1-pixel frame)
  int32_t *pos;    // [n * ne] heap payload
  int32_t *seg;    // [n * (1 + 2 * maxseg)] segment count, then (first, last) extended mask rows
  uint8_t *f;      // [n * ne] KNOWN / BAND / INSIDE of the inward march
  uint8_t *rg;     // [n * ne] ring flags of the outward march (INSIDE = ring, SEED = boundary band)
  uint8_t *rowflag;  // [n * H] image row has mask pixels
};

static inline int inp_maxseg(int H) { return (H + inp_gap(1)) / (inp_gap(1) + 1) + 1; }   // bound for every radius

static size_t inp_layout(InpWs *w, uint8_t *base, int n, int H, int W) {
  const size_t ne = (size_t)(H + 2) * (W + 2) * n;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t *p = base ? base + off : nullptr;
    off += (bytes + 15) & ~(size_t)15;
    return p;
  };
  w->key = (uint64_t *)take(ne * 8);
  w->t = (float *)take(ne * 4);
  w->pos = (int32_t *)take(ne * 4);
  w->seg = (int32_t *)take((size_t)n * (1 + 2 * inp_maxseg(H)) * 4);
  w->f = take(ne);
  w->rg = take(ne);
  w->rowflag = take((size_t)n * H);
  return off;
}

// ───────────── setup: flags, distance field, band and ring of every page ─────────────
__global__ void __launch_bounds__(256)
inp_init_kernel(const uint8_t *__restrict__ mask, InpWs w, int H, int W, int range) {
  const int er = H + 2, ec = W + 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, img = blockIdx.z;
  if (j >= ec) return;
  const uint8_t *m = mask + (size_t)img * H * W;
  auto M = [&](int ii, int jj) { return ii >= 1 && jj >= 1 && ii <= H && jj <= W && m[(size_t)(ii - 1) * W + jj - 1] != 0; };
  const size_t p = (size_t)img * er * ec + (size_t)i * ec + j;
  const bool interior = i >= 1 && j >= 1 && i <= H && j <= W;
  const bool in = M(i, j);
  bool band = false, ring = false;
  if (interior && !in) {
    band = M(i - 1, j) || M(i + 1, j) || M(i, j - 1) || M(i, j + 1);
    if (!band)
      for (int k = i - range; k <= i + range && !ring; ++k)
        for (int l = j - range; l <= j + range; ++l)
          if (M(k, l)) {
            ring = true;
            break;
          }
  }
  w.f[p] = in ? F_INSIDE : F_KNOWN;
  w.rg[p] = band ? F_SEED : ring ? F_INSIDE : F_KNOWN;
  w.t[p] = band ? 0.f : 1.0e6f;
  if (in) w.rowflag[(size_t)img * H + i - 1] = 1;
}

// one warp per page: maximal runs of mask rows, runs closer than `gap` clean rows merged
__global__ void __launch_bounds__(32)
inp_segment_kernel(InpWs w, int H, int maxseg, int gap) {
  const int img = blockIdx.x, lane = threadIdx.x;
  const uint8_t *rf = w.rowflag + (size_t)img * H;
  int32_t *seg = w.seg + (size_t)img * (1 + 2 * maxseg);
  int n = 0, first = -1, last = -1;
  for (int y0 = 0; y0 < H; y0 += 32) {
    unsigned bits = __ballot_sync(0xffffffffu, y0 + lane < H && rf[y0 + lane] != 0);
    if (lane != 0) continue;
    while (bits) {
      const int y = y0 + __ffs(bits) - 1;
      bits &= bits - 1;
      if (first >= 0 && y - last - 1 >= gap) {
        seg[1 + 2 * n] = first + 1;
        seg[2 + 2 * n] = last + 1;
        ++n;
        first = -1;
      }
      if (first < 0) first = y;
      last = y;
    }
  }
  if (lane != 0) return;
  if (first >= 0) {
    seg[1 + 2 * n] = first + 1;
    seg[2 + 2 * n] = last + 1;
    ++n;
  }
  seg[0] = n;
}

// ───────────── the queue (lane 0 only) ─────────────
// A 32-ary heap handled by the whole warp: a pop loads the 32 children of a node with one access per lane and finds the
// smallest (T bits, insertion number) with two warp min-reductions, so a queue of a few thousand entries is 2-3 levels
// deep.  The first `cap` entries live in shared memory, the rest (pathological masks only) in the segment's global slice.
// Every lane keeps n / seq; only lane 0 writes entries.
struct Heap {
  uint32_t *sT, *sS;
  int32_t *sP;
  uint32_t *gT, *gS;
  int32_t *gP;
  int cap;
  int n;
  uint32_t seq;
  __device__ __forceinline__ uint32_t T(int i) const { return i < cap ? sT[i] : gT[i]; }
  __device__ __forceinline__ uint32_t S(int i) const { return i < cap ? sS[i] : gS[i]; }
  __device__ __forceinline__ int P(int i) const { return i < cap ? sP[i] : gP[i]; }
  __device__ __forceinline__ void set(int i, uint32_t t, uint32_t q, int p) {
    if (i < cap) {
      sT[i] = t;
      sS[i] = q;
      sP[i] = p;
    } else {
      gT[i] = t;
      gS[i] = q;
      gP[i] = p;
    }
  }
};

// all lanes call with the same arguments
__device__ __forceinline__ void heap_push(Heap &h, int p, float Tv, int lane) {
  const uint32_t kt = Tv == 0.f ? 0u : __float_as_uint(Tv), ks = h.seq++;
  int i = h.n++;
  if (lane == 0) {
    while (i > 0) {
      const int par = (i - 1) >> 5;
      const uint32_t pt = h.T(par), ps = h.S(par);
      if (pt < kt || (pt == kt && ps <= ks)) break;
      h.set(i, pt, ps, h.P(par));
      i = par;
    }
    h.set(i, kt, ks, p);
  }
  __syncwarp();
}

// all lanes call; every lane gets the popped pixel (or -1)
__device__ __forceinline__ int heap_pop(Heap &h, int lane) {
  if (h.n == 0) return -1;
  const int top = h.P(0);
  const int n = --h.n;
  if (n == 0) return top;
  const uint32_t kt = h.T(n), ks = h.S(n);
  const int kp = h.P(n);
  __syncwarp();
  int i = 0;
  for (;;) {
    const int c0 = 32 * i + 1;
    if (c0 >= n) break;
    const int idx = c0 + lane;
    const bool valid = idx < n;
    const uint32_t ct = valid ? h.T(idx) : 0xffffffffu, cs = valid ? h.S(idx) : 0xffffffffu;
    const uint32_t mt = __reduce_min_sync(0xffffffffu, ct);
    const uint32_t cand = ct == mt ? cs : 0xffffffffu;
    const uint32_t ms = __reduce_min_sync(0xffffffffu, cand);
    if (mt > kt || (mt == kt && ms >= ks)) break;
    const int widx = c0 + __ffs(__ballot_sync(0xffffffffu, ct == mt && cand == ms)) - 1;
    if (lane == 0) h.set(i, mt, ms, h.P(widx));
    i = widx;
  }
  if (lane == 0) h.set(i, kt, ks, kp);
  __syncwarp();
  return top;
}

// Boundary-band pixels of the segment, in raster order, as the initial queue content: equal keys T = 0 with rising
// insertion numbers form a sorted array, which is a valid heap.  All lanes take part (ballot compaction).
__device__ int heap_seed(Heap &h, const uint8_t *rg, uint8_t marker, int row0, int row1, int ec, int lane) {
  int n = 0;
  const int total = (row1 - row0 + 1) * ec;
  for (int base = 0; base < total; base += 32) {
    const int q = base + lane;
    const bool is = q < total && rg[row0 * ec + q] == marker;
    const unsigned b = __ballot_sync(0xffffffffu, is);
    if (is) {
      const int idx = n + __popc(b & ((1u << lane) - 1));
      h.set(idx, 0u, (uint32_t)idx, row0 * ec + q);
    }
    n += __popc(b);
  }
  __syncwarp();
  h.n = n;
  h.seq = (uint32_t)n;
  return n;
}

__device__ __forceinline__ float fmm_solve(const uint8_t *f, const float *t, int p1, int p2) {
  const double a11 = t[p1], a22 = t[p2], m12 = a11 < a22 ? a11 : a22;
  double sol;
  if (f[p1] != F_INSIDE) {
    if (f[p2] != F_INSIDE) {
      const double d = a11 - a22;
      if (fabs(d) >= 1.0)
        sol = 1 + m12;
      else
        sol = (a11 + a22 + sqrt(2 - d * d)) * 0.5;
    } else
      sol = 1 + a11;
  } else if (f[p2] != F_INSIDE)
    sol = 1 + a22;
  else
    sol = 1 + m12;
  return (float)sol;
}

// Lanes 0..15: lane = 4 * neighbour + solve.  Returns (for every lane of a neighbour group) whether that neighbour of p
// is still INSIDE in `f`, and its arrival time (minimum of the four solves).
__device__ __forceinline__ bool neighbour_dist(const uint8_t *f, const float *t, int p, int er, int ec, int lane, int &nb,
                                               float &dist) {
  const int q = (lane >> 2) & 3, s = lane & 3;
  const int ii = p / ec, jj = p - ii * ec;
  const int i = ii + (q == 0 ? -1 : q == 2 ? 1 : 0), j = jj + (q == 1 ? -1 : q == 3 ? 1 : 0);
  nb = i * ec + j;
  bool valid = lane < 16 && !(i <= 0 || j <= 0 || i > er - 1 || j > ec - 1);
  valid = valid && f[nb] == F_INSIDE;
  float d = 3.0e38f;
  if (valid) d = fmm_solve(f, t, nb + ((s & 1) ? ec : -ec), nb + ((s & 2) ? 1 : -1));
  d = fminf(d, __shfl_xor_sync(0xffffffffu, d, 1));
  d = fminf(d, __shfl_xor_sync(0xffffffffu, d, 2));
  dist = d;
  return valid;
}

// ───────────── one painted pixel ─────────────
template <int C>
__device__ void paint_pixel(const uint8_t *f, const float *t, uint8_t *out, int er, int ec, int range, int p, int lane,
                            const float *dst_tab, float *terms) {
  const int W = ec - 2, D = 2 * range + 1;
  const int i = p / ec, j = p - i * ec;
  const float tij = t[p];
  float gx, gy;
  if (f[p + 1] != F_INSIDE)
    gx = f[p - 1] != F_INSIDE ? (t[p + 1] - t[p - 1]) * 0.5f : t[p + 1] - tij;
  else
    gx = f[p - 1] != F_INSIDE ? tij - t[p - 1] : 0.f;
  if (f[p + ec] != F_INSIDE)
    gy = f[p - ec] != F_INSIDE ? (t[p + ec] - t[p - ec]) * 0.5f : t[p + ec] - tij;
  else
    gy = f[p - ec] != F_INSIDE ? tij - t[p - ec] : 0.f;

  for (int q = lane; q < D * D; q += 32) {
    const int dk = q / D - range, dl = q - (q / D) * D - range;
    const int k = i + dk, l = j + dl;
    float tr[INP_CHAINS];
#pragma unroll
    for (int c = 0; c < INP_CHAINS; ++c) tr[c] = 0.f;
    if (k > 0 && l > 0 && k < er - 1 && l < ec - 1 && dk * dk + dl * dl <= range * range) {
      const int kl = k * ec + l;
      if (f[kl] != F_INSIDE) {
        const float ry = (float)(-dk), rx = (float)(-dl);
        const float lev = (float)(1. / (1 + fabs((double)(t[kl] - tij))));
        float dir = rx * gx + ry * gy;
        if ((double)fabsf(dir) <= 0.01) dir = 0.000001f;
        const float w = fabsf(dst_tab[q] * lev * dir);
        const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
        const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
        const bool fr = f[kl + 1] != F_INSIDE, fl = f[kl - 1] != F_INSIDE;
        const bool fd = f[kl + ec] != F_INSIDE, fu = f[kl - ec] != F_INSIDE;
        // the two pixels of each difference and its factor (central differences are doubled, not halved, in OpenCV)
        const int xa = (km * W + (fr ? lp + 1 : lp)) * C, xb = (km * W + (fl ? lm - 1 : lm)) * C;
        const int ya = ((fd ? kp + 1 : kp) * W + lm) * C, yb = ((fu ? km - 1 : km) * W + lm) * C;
        const float sx = fr ? (fl ? 2.0f : 1.0f) : (fl ? 1.0f : 0.f), sy = fd ? (fu ? 2.0f : 1.0f) : (fu ? 1.0f : 0.f);
        const int ctr = ((k - 1) * W + (l - 1)) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float gix = (float)((int)out[xa + c] - (int)out[xb + c]) * sx;
          const float giy = (float)((int)out[ya + c] - (int)out[yb + c]) * sy;
          tr[c] = w * (float)out[ctr + c];
          tr[3 + c] = -(w * (gix * rx));       // Jx -= v  ==  Jx += -v
          tr[6 + c] = -(w * (giy * ry));
        }
        tr[9] = w;
      }
    }
#pragma unroll
    for (int c = 0; c < INP_CHAINS; ++c) terms[q * INP_CHAINS + c] = tr[c];
  }
  __syncwarp();
  // running sums in raster order: lane c < 3 -> Ia[c], 3..5 -> Jx, 6..8 -> Jy, 9 -> s
  float acc = lane == 9 ? 1.0e-20f : 0.f;
  if (lane < INP_CHAINS)
    for (int q = 0; q < D * D; ++q) acc += terms[q * INP_CHAINS + lane];
  const int c = lane < C ? lane : 0;
  const float Ia = __shfl_sync(0xffffffffu, acc, c), Jx = __shfl_sync(0xffffffffu, acc, 3 + c);
  const float Jy = __shfl_sync(0xffffffffu, acc, 6 + c), s = __shfl_sync(0xffffffffu, acc, 9);
  if (lane < C) {
    const float sat = (float)(Ia / s + (Jx + Jy) / (sqrt((double)(Jx * Jx + Jy * Jy)) + (double)1.0e-20f));
    const int v = __float2int_rn(sat + 0.5f);
    out[((i - 1) * W + (j - 1)) * C + lane] = (uint8_t)min(max(v, 0), 255);
  }
  __syncwarp();
}

// ───────────── one warp per (segment, page) ─────────────
// Dynamic shared memory: [terms | dst_tab | window of t, f, rg, page rows (when it fits) | heap].
constexpr int INP_FIXED_SMEM = ((2 * INP_MAX_RANGE + 1) * (2 * INP_MAX_RANGE + 1) * (INP_CHAINS + 1) * 4 + 15) & ~15;
constexpr int INP_MIN_HEAP = 1024;    // entries kept in shared memory at the very least

template <int C>
__global__ void __launch_bounds__(32, 1)
inp_march_kernel(uint8_t *__restrict__ dst, InpWs w, int H, int W, int range, int maxseg, int smem_bytes) {
  extern __shared__ __align__(16) uint8_t smem[];
  float *terms = reinterpret_cast<float *>(smem);
  float *dst_tab = terms + (2 * INP_MAX_RANGE + 1) * (2 * INP_MAX_RANGE + 1) * INP_CHAINS;
  const int lane = threadIdx.x, img = blockIdx.y;
  const int32_t *seg = w.seg + (size_t)img * (1 + 2 * maxseg);
  if ((int)blockIdx.x >= seg[0]) return;
  const int er = H + 2, ec = W + 2;
  const int r0 = seg[1 + 2 * blockIdx.x], r1 = seg[2 + 2 * blockIdx.x];      // extended rows holding mask pixels
  const size_t ne = (size_t)er * ec;
  uint8_t *f = w.f + img * ne, *rg = w.rg + img * ne;
  float *t = w.t + img * ne;
  uint8_t *out = dst + (size_t)img * H * W * C;
  const int lo = max(r0 - range, 0), hi = min(r1 + range, er - 1);
  // everything the march reads or writes lies in extended rows wlo..whi and page rows olo..ohi
  const int wlo = max(r0 - range - 1, 0), whi = min(r1 + range + 1, er - 1);
  const int olo = max(wlo - 1, 0), ohi = min(whi - 1, H - 1);
  const size_t wpx = (size_t)(whi - wlo + 1) * ec, obytes = (size_t)(ohi - olo + 1) * W * C;
  const size_t win_bytes = ((wpx * 4 + 15) & ~(size_t)15) + 2 * ((wpx + 15) & ~(size_t)15) + ((obytes + 15) & ~(size_t)15);
  const bool staged = INP_FIXED_SMEM + win_bytes + (size_t)INP_MIN_HEAP * 12 <= (size_t)smem_bytes;
  uint8_t *sp = smem + INP_FIXED_SMEM;
  uint8_t *out_s = nullptr;
  if (staged) {
    float *t_s = reinterpret_cast<float *>(sp);
    sp += (wpx * 4 + 15) & ~(size_t)15;
    uint8_t *f_s = sp;
    sp += (wpx + 15) & ~(size_t)15;
    uint8_t *rg_s = sp;
    sp += (wpx + 15) & ~(size_t)15;
    out_s = sp;
    sp += (obytes + 15) & ~(size_t)15;
    const size_t w0 = (size_t)wlo * ec;
    for (size_t q = lane; q < wpx; q += 32) {
      t_s[q] = t[w0 + q];
      f_s[q] = f[w0 + q];
      rg_s[q] = rg[w0 + q];
    }
    const uint8_t *og = out + (size_t)olo * W * C;
    for (size_t q = lane; q < obytes; q += 32) out_s[q] = og[q];
    t = t_s - w0;            // generic pointers biased so that extended / page coordinates index them unchanged
    f = f_s - w0;
    rg = rg_s - w0;
    out = out_s - (size_t)olo * W * C;
  }
  Heap h;
  h.cap = (int)((smem + smem_bytes - sp) / 12);
  h.sT = reinterpret_cast<uint32_t *>(sp);
  h.sS = h.sT + h.cap;
  h.sP = reinterpret_cast<int32_t *>(h.sS + h.cap);
  const size_t slice = (size_t)(hi - lo + 1) * ec;          // the segment's part of the global spill arrays
  h.gT = reinterpret_cast<uint32_t *>(w.key + img * ne + (size_t)lo * ec);
  h.gS = h.gT + slice;
  h.gP = w.pos + img * ne + (size_t)lo * ec;
  const int D = 2 * range + 1;
  for (int q = lane; q < D * D; q += 32) {
    const int dk = q / D - range, dl = q % D - range;
    const float len2 = (float)(dk * dk + dl * dl);
    dst_tab[q] = len2 > 0.f ? (float)(1. / (len2 * sqrt((double)len2))) : 0.f;
  }

  // march outwards through the ring: distances there, negated afterwards
  __syncwarp();
  heap_seed(h, rg, F_SEED, r0 - 1, r1 + 1, ec, lane);
  for (;;) {
    const int p = heap_pop(h, lane);
    if (p < 0) break;
    if (lane == 0) rg[p] = rg[p] == F_SEED ? F_SEED_DONE : F_CHANGE;
    __syncwarp();
    int nb;
    float d;
    const bool valid = neighbour_dist(rg, t, p, er, ec, lane, nb, d);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool vq = __shfl_sync(0xffffffffu, (int)valid, 4 * q) != 0;
      const float dq = __shfl_sync(0xffffffffu, d, 4 * q);
      const int nq = __shfl_sync(0xffffffffu, nb, 4 * q);
      if (vq) {
        if (lane == 0) {
          t[nq] = dq;
          rg[nq] = F_BAND;
        }
        heap_push(h, nq, dq, lane);
      }
    }
  }
  for (int q = lo * ec + lane; q < (hi + 1) * ec; q += 32)
    if (rg[q] == F_CHANGE || rg[q] == F_SEED_DONE) t[q] = -t[q];
  __syncwarp();

  // march inwards, painting every pixel when the front reaches it
  heap_seed(h, rg, F_SEED_DONE, r0 - 1, r1 + 1, ec, lane);
  for (;;) {
    const int p = heap_pop(h, lane);
    if (p < 0) break;
    if (lane == 0) f[p] = F_KNOWN;
    __syncwarp();
    int nb;
    float d;
    const bool valid = neighbour_dist(f, t, p, er, ec, lane, nb, d);
    __syncwarp();
    for (int q = 0; q < 4; ++q) {
      const bool vq = __shfl_sync(0xffffffffu, (int)valid, 4 * q) != 0;
      const float dq = __shfl_sync(0xffffffffu, d, 4 * q);
      const int nq = __shfl_sync(0xffffffffu, nb, 4 * q);
      if (!vq) continue;
      if (lane == 0) t[nq] = dq;
      __syncwarp();
      paint_pixel<C>(f, t, out, er, ec, range, nq, lane, dst_tab, terms);
      if (lane == 0) f[nq] = F_BAND;
      heap_push(h, nq, dq, lane);
    }
  }
  if (staged) {                                  // painted rows back to the page
    uint8_t *og = dst + (size_t)img * H * W * C;
    const size_t b0 = (size_t)(r0 - 1) * W * C, b1 = (size_t)r1 * W * C;
    for (size_t q = b0 + lane; q < b1; q += 32) og[q] = out[q];
  }
}

}  // namespace ocrb

using namespace ocrb;

extern "C" int64_t ocrb_inpaint_workspace_bytes(int32_t n_img, int32_t H, int32_t W) {
  if (n_img <= 0 || H <= 0 || W <= 0) return 0;
  InpWs w;
  return (int64_t)inp_layout(&w, nullptr, n_img, H, W);
}

extern "C" int ocrb_inpaint_telea_u8(const uint8_t *src, const uint8_t *mask, uint8_t *dst, uint8_t *ws, int32_t n_img,
                                     int32_t H, int32_t W, int32_t C, int32_t radius, void *stream) {
  OCRB_REQUIRE(src && mask && dst && ws && n_img > 0 && (C == 1 || C == 3), "inpaint_telea_u8: bad arguments");
  OCRB_REQUIRE(H >= 2 && W >= 2, "inpaint_telea_u8: pages one pixel high or wide are not supported (OpenCV reads outside them)");
  OCRB_REQUIRE(radius >= 1 && radius <= INP_MAX_RANGE, "inpaint_telea_u8: radius must be in 1..7");
  OCRB_REQUIRE(src != dst && ((uintptr_t)ws & 15) == 0, "inpaint_telea_u8: in-place not supported; workspace must be 16-byte aligned");
  OCRB_REQUIRE((size_t)(H + 2) * (W + 2) * 3 < ((size_t)1 << 31) && H + 2 <= 65535 && n_img <= 65535, "inpaint_telea_u8: page too large");
  cudaStream_t st = (cudaStream_t)stream;
  InpWs w;
  inp_layout(&w, ws, n_img, H, W);
  const int maxseg = inp_maxseg(H);
  OCRB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n_img * H * W * C, cudaMemcpyDeviceToDevice, st));
  OCRB_CUDA(cudaMemsetAsync(w.rowflag, 0, (size_t)n_img * H, st));
  inp_init_kernel<<<dim3(cdiv(W + 2, 256), H + 2, n_img), 256, 0, st>>>(mask, w, H, W, radius);
  int rc = check_launch("inp_init_kernel");
  if (rc) return rc;
  inp_segment_kernel<<<n_img, 32, 0, st>>>(w, H, maxseg, inp_gap(radius));
  if ((rc = check_launch("inp_segment_kernel"))) return rc;
  static int smem_bytes = 0;
  if (!smem_bytes) {
    int dev = 0, optin = 0;
    OCRB_CUDA(cudaGetDevice(&dev));
    OCRB_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    OCRB_REQUIRE(optin >= INP_FIXED_SMEM + INP_MIN_HEAP * 12, "inpaint_telea_u8: not enough shared memory per block");
    OCRB_CUDA(cudaFuncSetAttribute(inp_march_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    OCRB_CUDA(cudaFuncSetAttribute(inp_march_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    smem_bytes = optin;
  }
  if (C == 1)
    inp_march_kernel<1><<<dim3(maxseg, n_img), 32, smem_bytes, st>>>(dst, w, H, W, radius, maxseg, smem_bytes);
  else
    inp_march_kernel<3><<<dim3(maxseg, n_img), 32, smem_bytes, st>>>(dst, w, H, W, radius, maxseg, smem_bytes);
  return check_launch("inp_march_kernel");
}
