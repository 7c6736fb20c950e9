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
//     accumulated in OpenCV's raster order (one lane per running sum), because float addition does not reassociate;
//   * the two marches and the painting are decoupled (one CTA of 4 warps per segment).  The outward march (negative
//     distances in the ring) and the inward march (arrival times of the mask pixels) read disjoint data -- the eikonal
//     solve of a ring pixel never touches a mask pixel and vice versa, the band between them is T = 0 in both -- so they
//     run concurrently on two warps with two queues.  The inward march does not depend on pixel VALUES at all: it only
//     emits the list of pixels in the order OpenCV would paint them, stamping each with its position in that order.
//     Two painter warps consume the list: "known when pixel n is painted" is stamp < n, so a painter can evaluate
//     flags, distances and weights of job n while job n - 1 is still being painted by the other warp, and only the
//     part that reads pixel values waits for its predecessor (a counter in shared memory).
// The queue is a binary heap on the key (float bits of T) << 32 | insertion number, equal in order to OpenCV's sorted
// list; only lane 0 touches it.  A segment's window of the distance field, the flags and the page rows is staged in
// shared memory together with the heap (freshly written global data would come back from L2 at ~600 cycles a load,
// and the march is one long dependent chain); segments too tall for that run from global memory through the same
// generic pointers, and a heap that outgrows its shared part spills into the segment's global slice.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ocrb {

enum : uint8_t { F_KNOWN = 0, F_BAND = 1, F_INSIDE = 2, F_CHANGE = 3, F_SEED = 4, F_SEED_DONE = 5 };
constexpr int INP_MAX_RANGE = 7;
// clean rows between independent segments: the march reads up to range + 1 rows away from the mask and writes up to
// range rows away, so 2 * range + 1 clean rows already separate two segments; one more for margin
static inline __host__ __device__ int inp_gap(int range) { return 2 * range + 2; }
constexpr int INP_CHAINS = 10;        // Ia[3], Jx[3], Jy[3], s
constexpr int INP_NOTYET = 0x7fffffff;
#ifndef INP_SPIN_SLEEP
#define INP_SPIN_SLEEP 0
#endif
constexpr int INP_PAINTERS = 3;
constexpr int INP_WARPS = 2 + INP_PAINTERS;   // outward march, inward march, painters
// a waiting warp traps instead of hanging if its partner never arrives (2^36 cycles: half a minute)
__device__ __forceinline__ void inp_spin_check(int it, long long t0) {
  if (INP_SPIN_SLEEP) __nanosleep(INP_SPIN_SLEEP);
  if ((it & 1023) == 1023 && clock64() - t0 > (1ll << 36)) __trap();
}

struct InpWs {
  uint64_t *key;   // [n * ne] heap keys
  float *t;        // [n * ne] distance field (extended image: 1-pixel frame)
  int32_t *pos;    // [n * ne] heap payload
  int32_t *seg;    // [n * (1 + 2 * maxseg)] segment count, then (first, last) extended mask rows
  uint8_t *f;      // [n * ne] KNOWN / BAND / INSIDE of the inward march
  uint8_t *rg;     // [n * ne] ring flags of the outward march (INSIDE = ring, SEED = boundary band)
  uint8_t *rowflag;  // [n * H] image row has mask pixels
  uint64_t *key2;  // second queue (the two marches run concurrently)
  int32_t *pos2;
  int32_t *stamp;  // [n * ne] paint order of a mask pixel (NOTYET until the inward march reaches it), -1 elsewhere
  int32_t *job;    // [n * ne] pixels in paint order, per segment at the offset of its first mask row
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
  w->key2 = (uint64_t *)take(ne * 8);
  w->pos2 = (int32_t *)take(ne * 4);
  w->stamp = (int32_t *)take(ne * 4);
  w->job = (int32_t *)take(ne * 4);
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
  w.stamp[p] = in ? INP_NOTYET : -1;
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

// all lanes call with the same arguments and walk the same path (loads are broadcasts, only lane 0 stores): a lane-0-only
// loop would leave the warp diverged, and every later shuffle / ballot / redux would take the slow collective path
__device__ __forceinline__ void heap_push(Heap &h, int p, float Tv, int lane) {
  const uint32_t kt = Tv == 0.f ? 0u : __float_as_uint(Tv), ks = h.seq++;
  int i = h.n++;
  while (i > 0) {
    const int par = (i - 1) >> 5;
    const uint32_t pt = h.T(par), ps = h.S(par);
    if (pt < kt || (pt == kt && ps <= ks)) break;
    const int pp = h.P(par);
    if (lane == 0) h.set(i, pt, ps, pp);
    i = par;
  }
  if (lane == 0) h.set(i, kt, ks, p);
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
__device__ int heap_seed(Heap &h, const uint8_t *rg, int row0, int row1, int ec, int lane) {
  int n = 0;
  const int total = (row1 - row0 + 1) * ec;
  for (int base = 0; base < total; base += 32) {
    const int q = base + lane;
    const uint8_t v = q < total ? rg[row0 * ec + q] : (uint8_t)F_KNOWN;
    const bool is = v == F_SEED || v == F_SEED_DONE;      // the outward march flips SEED -> SEED_DONE while it runs
    const unsigned b = __ballot_sync(0xffffffffu, is);
    if (is) {
      const int idx = n + __popc(b & ((1u << lane) - 1));
      h.set(idx, 0u, (uint32_t)idx, ((row0 + q / ec) << 16) | (q % ec));
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
__device__ __forceinline__ bool neighbour_dist(const uint8_t *f, const float *t, int pk, int er, int ec, int lane, int &nb,
                                               int &nb_pk, float &dist) {
  const int q = (lane >> 2) & 3, s = lane & 3;
  const int ii = pk >> 16, jj = pk & 0xffff;             // queue entries carry (row << 16 | column): no division per pop
  const int i = ii + (q == 0 ? -1 : q == 2 ? 1 : 0), j = jj + (q == 1 ? -1 : q == 3 ? 1 : 0);
  nb = i * ec + j;
  nb_pk = (i << 16) | j;
  bool valid = lane < 16 && !(i <= 0 || j <= 0 || i > er - 1 || j > ec - 1);
  valid = valid && f[nb] == F_INSIDE;
  float d = 3.0e38f;
  if (valid) d = fmm_solve(f, t, nb + ((s & 1) ? ec : -ec), nb + ((s & 2) ? 1 : -1));
  d = fminf(d, __shfl_xor_sync(0xffffffffu, d, 1));
  d = fminf(d, __shfl_xor_sync(0xffffffffu, d, 2));
  dist = d;
  return valid;
}

// ───────────── one painted pixel, in two parts ─────────────
struct InpSync {        // shared-memory mailbox of a segment's four warps
  int jobs_ready;       // pixels emitted by the inward march so far
  int queue_total;      // -1 while the inward march runs, then the number of pixels
  int outside_done;     // ring distances final (negated)
  int painted_upto;     // pixels 0 .. painted_upto-1 carry their final value
  long long ts[6];      // OCRB_INPAINT_DEBUG=1: clock64 at start / outward done / inward done / painters done / first paint
};

struct PosState {       // what one disc position contributes, everything that does not depend on pixel values
  float w, sx, sy, rx, ry;
  int xa, xb, ya, yb, ctr;   // ctr < 0: position contributes nothing
};

__device__ __forceinline__ bool inp_known(const int32_t *stamp, int m0, int m1, int q, int n) {
  return q < m0 || q >= m1 || stamp[q] < n;
}

// disc entry e = (dk, dl) packed by the kernel prologue, raster order, centre left out
template <int C>
__device__ __forceinline__ void paint_position(PosState &st, int e, int nd, int i, int j, int n, float tij, float gx, float gy,
                                               const float *t, const int32_t *stamp, int m0, int m1, int er, int ec,
                                               const int *disc, const float *dst_tab) {
  const int W = ec - 2;
  // a position that contributes nothing gets weight 0 and reads the pixel being painted: no branch in the dependent part
  st.w = 0.f;
  st.sx = st.sy = st.rx = st.ry = 0.f;
  st.xa = st.xb = st.ya = st.yb = st.ctr = ((i - 1) * W + (j - 1)) * C;
  if (e >= nd) return;
  const int dkl = disc[e], dk = dkl >> 8, dl = (int)(int8_t)(dkl & 0xff);
  const int k = i + dk, l = j + dl;
  if (!(k > 0 && l > 0 && k < er - 1 && l < ec - 1)) return;
  const int kl = k * ec + l;
  if (!inp_known(stamp, m0, m1, kl, n)) return;
  st.ry = (float)(-dk);
  st.rx = (float)(-dl);
  const float lev = (float)(1. / (1 + fabs((double)(t[kl] - tij))));
  float dir = st.rx * gx + st.ry * gy;
  if ((double)fabsf(dir) <= 0.01) dir = 0.000001f;
  st.w = fabsf(dst_tab[e] * lev * dir);
  const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
  const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
  const bool fr = inp_known(stamp, m0, m1, kl + 1, n), fl = inp_known(stamp, m0, m1, kl - 1, n);
  const bool fd = inp_known(stamp, m0, m1, kl + ec, n), fu = inp_known(stamp, m0, m1, kl - ec, n);
  // the two pixels of each difference and its factor (central differences are doubled, not halved, in OpenCV)
  st.xa = (km * W + (fr ? lp + 1 : lp)) * C;
  st.xb = (km * W + (fl ? lm - 1 : lm)) * C;
  st.ya = ((fd ? kp + 1 : kp) * W + lm) * C;
  st.yb = ((fu ? km - 1 : km) * W + lm) * C;
  st.sx = fr ? (fl ? 2.0f : 1.0f) : (fl ? 1.0f : 0.f);
  st.sy = fd ? (fu ? 2.0f : 1.0f) : (fu ? 1.0f : 0.f);
  st.ctr = ((k - 1) * W + (l - 1)) * C;
}

// terms[chain * ndp + e]: chain-major, so that the lane owning a running sum walks consecutive words
template <int C>
__device__ __forceinline__ void paint_terms(const PosState &st, int e, int nd, int ndp, const uint8_t *out, float *terms) {
  if (e >= nd) return;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float gix = (float)((int)out[st.xa + c] - (int)out[st.xb + c]) * st.sx;
    const float giy = (float)((int)out[st.ya + c] - (int)out[st.yb + c]) * st.sy;
    terms[c * ndp + e] = st.w * (float)out[st.ctr + c];
    terms[(3 + c) * ndp + e] = -(st.w * (gix * st.rx));       // Jx -= v  ==  Jx += -v
    terms[(6 + c) * ndp + e] = -(st.w * (giy * st.ry));
  }
  terms[9 * ndp + e] = st.w;
}

// A painter warp: jobs k, k + INP_PAINTERS, ...
template <int C>
__device__ void painter_loop(int k, int lane, volatile InpSync *sy, const int32_t *job, const int32_t *stamp, int m0, int m1,
                             const float *t, uint8_t *out, int er, int ec, int nd, int ndp, const int *disc,
                             const float *dst_tab, float *terms, int debug) {
  const int W = ec - 2;
  bool outside_seen = false;
  long long ph[6] = {0, 0, 0, 0, 0, 0}, tc = clock64();   // debug: cycles per phase of this painter
#define INP_PHASE(x)                      \
  if (debug) {                            \
    const long long now_ = clock64();     \
    ph[x] += now_ - tc;                   \
    tc = now_;                            \
  }
  for (int n = k;; n += INP_PAINTERS) {
    // every lane polls (one broadcast load per poll): the warp stays converged, so the collectives below are cheap
    bool ok = true;
    {
      const long long t0 = clock64();
      for (int it = 0;; ++it) {
        if (sy->jobs_ready > n) break;
        const int tot = sy->queue_total;
        if (tot >= 0 && n >= tot) {
          ok = false;
          break;
        }
        inp_spin_check(it, t0);
      }
      if (ok && !outside_seen)
        for (int it = 0; !sy->outside_done; ++it) inp_spin_check(it, t0);
    }
    if (!ok) break;
    if (outside_seen) {
      INP_PHASE(0)                                             // 0: waiting for a job (after the outward march is done)
    } else {
      tc = clock64();
    }
    outside_seen = true;
    __threadfence_block();
    const int pk = job[n];
    const int i = pk >> 16, j = pk & 0xffff, p = i * ec + j;
    const float tij = t[p];
    float gx, gy;
    {
      const bool r = inp_known(stamp, m0, m1, p + 1, n), l = inp_known(stamp, m0, m1, p - 1, n);
      const bool d = inp_known(stamp, m0, m1, p + ec, n), u = inp_known(stamp, m0, m1, p - ec, n);
      gx = r ? (l ? (t[p + 1] - t[p - 1]) * 0.5f : t[p + 1] - tij) : (l ? tij - t[p - 1] : 0.f);
      gy = d ? (u ? (t[p + ec] - t[p - ec]) * 0.5f : t[p + ec] - tij) : (u ? tij - t[p - ec] : 0.f);
    }
    // everything that does not read pixel values: the first two rounds of disc positions (all of them for radius <= 4)
    PosState s0, s1;
    paint_position<C>(s0, lane, nd, i, j, n, tij, gx, gy, t, stamp, m0, m1, er, ec, disc, dst_tab);
    paint_position<C>(s1, lane + 32, nd, i, j, n, tij, gx, gy, t, stamp, m0, m1, er, ec, disc, dst_tab);
    INP_PHASE(1)                                               // 1: prologue
    {
      const long long t0 = clock64();
      for (int it = 0; sy->painted_upto != n; ++it) inp_spin_check(it, t0);
    }
    __syncwarp();
    __threadfence_block();
    INP_PHASE(2)                                               // 2: waiting for the predecessor
    paint_terms<C>(s0, lane, nd, ndp, out, terms);
    paint_terms<C>(s1, lane + 32, nd, ndp, out, terms);
    for (int e = lane + 64; e < nd; e += 32) {          // radius > 4 only
      PosState sx_;
      paint_position<C>(sx_, e, nd, i, j, n, tij, gx, gy, t, stamp, m0, m1, er, ec, disc, dst_tab);
      paint_terms<C>(sx_, e, nd, ndp, out, terms);
    }
    __syncwarp();
    INP_PHASE(3)                                               // 3: terms
    // running sums in raster order: lane c < 3 -> Ia[c], 3..5 -> Jx, 6..8 -> Jy, 9 -> s
    float acc = lane == 9 ? 1.0e-20f : 0.f;
    {
      const int ch = lane < INP_CHAINS && (lane == 9 || lane % 3 < C) ? lane : 9;
      const float *tp = terms + ch * ndp;
#pragma unroll 4
      for (int e = 0; e < nd; ++e) acc += tp[e];
    }
    const int c = lane < C ? lane : 0;
    const float Ia = __shfl_sync(0xffffffffu, acc, c), Jx = __shfl_sync(0xffffffffu, acc, 3 + c);
    const float Jy = __shfl_sync(0xffffffffu, acc, 6 + c), s = __shfl_sync(0xffffffffu, acc, 9);
    INP_PHASE(4)                                               // 4: ordered sums
    if (lane < C) {
      // OpenCV evaluates the whole expression in float (quotient, sum of squares, square root, quotient, sum), adds
      // 0.5f and rounds half to even; IEEE division and square root here (no fast-math), -fmad=false for the sums
      const float sat = Ia / s + (Jx + Jy) / (sqrtf(Jx * Jx + Jy * Jy) + 1.0e-20f);
      const int v = __float2int_rn(sat + 0.5f);
      out[((i - 1) * W + (j - 1)) * C + lane] = (uint8_t)min(max(v, 0), 255);
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) sy->painted_upto = n + 1;
    INP_PHASE(5)                                               // 5: final quotient, store, publish
  }
  if (debug && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0)
    printf("painter %d: wait-job %lld prologue %lld wait-pred %lld terms %lld sums %lld finish %lld cycles\n", k, ph[0], ph[1],
           ph[2], ph[3], ph[4], ph[5]);
#undef INP_PHASE
}

// ───────────── one CTA (2 march warps + painters) per (segment, page) ─────────────
// Dynamic shared memory: [mailbox | dst_tab | painters' terms | window of t, f, rg, page rows, stamps, jobs (when it
// fits) | two heaps].
constexpr int INP_DD_MAX = (2 * INP_MAX_RANGE + 1) * (2 * INP_MAX_RANGE + 1);
constexpr int INP_MIN_HEAP = 512;     // entries each queue keeps in shared memory at the very least
__host__ __device__ inline size_t inp_a16(size_t x) { return (x + 15) & ~(size_t)15; }

template <int C>
__global__ void __launch_bounds__(32 * INP_WARPS, 1)
inp_march_kernel(uint8_t *__restrict__ dst, InpWs w, int H, int W, int range, int maxseg, int smem_bytes, int stage_ok) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, img = blockIdx.y;
  const int32_t *seg = w.seg + (size_t)img * (1 + 2 * maxseg);
  if ((int)blockIdx.x >= seg[0]) return;
  const int er = H + 2, ec = W + 2, D = 2 * range + 1, DD = D * D;
  const int r0 = seg[1 + 2 * blockIdx.x], r1 = seg[2 + 2 * blockIdx.x];      // extended rows holding mask pixels
  const size_t ne = (size_t)er * ec;
  volatile InpSync *sy = reinterpret_cast<volatile InpSync *>(smem);
  // disc positions (dk, dl) with dk^2 + dl^2 <= range^2 in raster order, centre left out (the pixel being painted is
  // never known to itself); same count on every thread
  int nd = 0;
  for (int q = 0; q < DD; ++q) {
    const int dk = q / D - range, dl = q % D - range;
    nd += (dk * dk + dl * dl <= range * range && (dk | dl) != 0);
  }
  const int ndp = (nd + 3) & ~3;
  float *dst_tab = reinterpret_cast<float *>(smem + 96);
  int *disc = reinterpret_cast<int *>(dst_tab + INP_DD_MAX);
  float *terms_all = reinterpret_cast<float *>(disc + INP_DD_MAX);
  uint8_t *sp = smem + inp_a16(96 + 2 * INP_DD_MAX * 4 + (size_t)INP_PAINTERS * ndp * INP_CHAINS * 4);
  uint8_t *f = w.f + img * ne, *rg = w.rg + img * ne;
  float *t = w.t + img * ne;
  int32_t *stamp = w.stamp + img * ne;
  uint8_t *out = dst + (size_t)img * H * W * C;
  const int lo = max(r0 - range, 0), hi = min(r1 + range, er - 1);
  // everything the marches read or write lies in extended rows wlo..whi and page rows olo..ohi
  const int wlo = max(r0 - range - 1, 0), whi = min(r1 + range + 1, er - 1);
  const int olo = max(wlo - 1, 0), ohi = min(whi - 1, H - 1);
  const int m0 = r0 * ec, m1 = (r1 + 1) * ec;                 // the mask rows: stamps and jobs exist only there
  int32_t *job = w.job + img * ne + m0;
  const size_t wpx = (size_t)(whi - wlo + 1) * ec, obytes = (size_t)(ohi - olo + 1) * W * C, mpx = (size_t)(m1 - m0);
  const size_t win_bytes = inp_a16(wpx * 4) + 2 * inp_a16(wpx) + inp_a16(obytes) + 2 * inp_a16(mpx * 4);
  const bool staged = stage_ok && (size_t)(sp - smem) + win_bytes + 2 * (size_t)INP_MIN_HEAP * 12 <= (size_t)smem_bytes;
  if (staged) {
    float *t_s = reinterpret_cast<float *>(sp);
    sp += inp_a16(wpx * 4);
    uint8_t *f_s = sp;
    sp += inp_a16(wpx);
    uint8_t *rg_s = sp;
    sp += inp_a16(wpx);
    uint8_t *out_s = sp;
    sp += inp_a16(obytes);
    int32_t *stamp_s = reinterpret_cast<int32_t *>(sp);
    sp += inp_a16(mpx * 4);
    int32_t *job_s = reinterpret_cast<int32_t *>(sp);
    sp += inp_a16(mpx * 4);
    const size_t w0 = (size_t)wlo * ec;
    for (size_t q = threadIdx.x; q < wpx; q += 32 * INP_WARPS) {
      t_s[q] = t[w0 + q];
      f_s[q] = f[w0 + q];
      rg_s[q] = rg[w0 + q];
    }
    for (size_t q = threadIdx.x; q < mpx; q += 32 * INP_WARPS) stamp_s[q] = stamp[m0 + q];
    const uint8_t *og = out + (size_t)olo * W * C;
    for (size_t q = threadIdx.x; q < obytes; q += 32 * INP_WARPS) out_s[q] = og[q];
    t = t_s - w0;            // generic pointers biased so that extended / page coordinates index them unchanged
    f = f_s - w0;
    rg = rg_s - w0;
    out = out_s - (size_t)olo * W * C;
    stamp = stamp_s - m0;
    job = job_s;
  }
  if (threadIdx.x == 0) {
    int e = 0;
    for (int q = 0; q < DD; ++q) {
      const int dk = q / D - range, dl = q % D - range;
      if (dk * dk + dl * dl > range * range || (dk | dl) == 0) continue;
      const float len2 = (float)(dk * dk + dl * dl);
      dst_tab[e] = (float)(1. / (len2 * sqrt((double)len2)));
      disc[e++] = (dk << 8) | (dl & 0xff);
    }
  }
  if (threadIdx.x == 0) {
    sy->jobs_ready = 0;
    sy->queue_total = -1;
    sy->outside_done = 0;
    sy->painted_upto = 0;
    sy->ts[0] = clock64();
  }
  __syncthreads();

  if (warp < 2) {
    // the two queues split what is left of the shared memory; each spills into its own global slice
    Heap h;
    const int cap = (int)((smem + smem_bytes - sp) / 24);
    uint8_t *hp = sp + (size_t)warp * cap * 12;
    h.cap = cap;
    h.sT = reinterpret_cast<uint32_t *>(hp);
    h.sS = h.sT + cap;
    h.sP = reinterpret_cast<int32_t *>(h.sS + cap);
    const size_t slice = (size_t)(hi - lo + 1) * ec;
    uint64_t *gk = (warp == 0 ? w.key : w.key2) + img * ne + (size_t)lo * ec;
    h.gT = reinterpret_cast<uint32_t *>(gk);
    h.gS = h.gT + slice;
    h.gP = (warp == 0 ? w.pos : w.pos2) + img * ne + (size_t)lo * ec;
    heap_seed(h, rg, r0 - 1, r1 + 1, ec, lane);
    if (warp == 0) {
      // march outwards through the ring: distances there, negated afterwards
      for (;;) {
        const int pk = heap_pop(h, lane);
        if (pk < 0) break;
        const int p = (pk >> 16) * ec + (pk & 0xffff);
        if (lane == 0) rg[p] = rg[p] == F_SEED ? F_SEED_DONE : F_CHANGE;
        __syncwarp();
        int nb, nbk;
        float d;
        const bool valid = neighbour_dist(rg, t, pk, er, ec, lane, nb, nbk, d);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool vq = __shfl_sync(0xffffffffu, (int)valid, 4 * q) != 0;
          const float dq = __shfl_sync(0xffffffffu, d, 4 * q);
          const int nq = __shfl_sync(0xffffffffu, nb, 4 * q), nk = __shfl_sync(0xffffffffu, nbk, 4 * q);
          if (vq) {
            if (lane == 0) {
              t[nq] = dq;
              rg[nq] = F_BAND;
            }
            heap_push(h, nk, dq, lane);
          }
        }
      }
      for (int q = lo * ec + lane; q < (hi + 1) * ec; q += 32)
        if (rg[q] == F_CHANGE || rg[q] == F_SEED_DONE) t[q] = -t[q];
      __threadfence_block();
      __syncwarp();
      if (lane == 0) {
        sy->ts[1] = clock64();
        sy->outside_done = 1;
      }
    } else {
      // march inwards: arrival times of the mask pixels and the order in which they are painted
      int n = 0;
      for (;;) {
        const int pk = heap_pop(h, lane);
        if (pk < 0) break;
        const int p = (pk >> 16) * ec + (pk & 0xffff);
        if (lane == 0) f[p] = F_KNOWN;
        __syncwarp();
        int nb, nbk;
        float d;
        const bool valid = neighbour_dist(f, t, pk, er, ec, lane, nb, nbk, d);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool vq = __shfl_sync(0xffffffffu, (int)valid, 4 * q) != 0;
          const float dq = __shfl_sync(0xffffffffu, d, 4 * q);
          const int nq = __shfl_sync(0xffffffffu, nb, 4 * q), nk = __shfl_sync(0xffffffffu, nbk, 4 * q);
          if (vq) {
            if (lane == 0) {
              t[nq] = dq;
              f[nq] = F_BAND;
              stamp[nq] = n;
              job[n] = nk;
            }
            __threadfence_block();
            if (lane == 0) sy->jobs_ready = n + 1;
            ++n;
            heap_push(h, nk, dq, lane);
          }
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) {
        sy->ts[2] = clock64();
        sy->queue_total = n;
      }
    }
  } else {
    painter_loop<C>(warp - 2, lane, sy, job, stamp, m0, m1, t, out, er, ec, nd, ndp, disc, dst_tab,
                    terms_all + (size_t)(warp - 2) * ndp * INP_CHAINS, stage_ok > 1);
    if (lane == 0 && warp == 2) sy->ts[3] = clock64();
  }
  __syncthreads();
  if (stage_ok > 1 && threadIdx.x == 0 && blockIdx.x < 2 && img == 0)
    printf("inpaint segment %d rows %d..%d staged %d jobs %d: outward %lld inward %lld painters %lld cycles\n", (int)blockIdx.x,
           r0, r1, (int)staged, sy->queue_total, sy->ts[1] - sy->ts[0], sy->ts[2] - sy->ts[0], sy->ts[3] - sy->ts[0]);
  if (staged) {                                  // painted rows back to the page
    uint8_t *og = dst + (size_t)img * H * W * C;
    const size_t b0 = (size_t)(r0 - 1) * W * C, b1 = (size_t)r1 * W * C;
    for (size_t q = b0 + threadIdx.x; q < b1; q += 32 * INP_WARPS) og[q] = out[q];
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
  OCRB_REQUIRE((size_t)(H + 2) * (W + 2) * 3 < ((size_t)1 << 31) && H + 2 <= 32767 && W + 2 <= 65535 && n_img <= 65535, "inpaint_telea_u8: page too large");
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
    OCRB_REQUIRE(optin >= 64 * 1024, "inpaint_telea_u8: not enough shared memory per block");
    OCRB_CUDA(cudaFuncSetAttribute(inp_march_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    OCRB_CUDA(cudaFuncSetAttribute(inp_march_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    smem_bytes = optin;
  }
  static int stage_ok = -1;               // OCRB_INPAINT_STAGE=0 forces the global-memory path (tests, measurements)
  if (stage_ok < 0) {
    const char *e = getenv("OCRB_INPAINT_STAGE");
    stage_ok = e ? atoi(e) : 1;           // 2: staged + one timing line per segment (device printf)
  }
  if (C == 1)
    inp_march_kernel<1><<<dim3(maxseg, n_img), 32 * INP_WARPS, smem_bytes, st>>>(dst, w, H, W, radius, maxseg, smem_bytes, stage_ok);
  else
    inp_march_kernel<3><<<dim3(maxseg, n_img), 32 * INP_WARPS, smem_bytes, st>>>(dst, w, H, W, radius, maxseg, smem_bytes, stage_ok);
  return check_launch("inp_march_kernel");
}
