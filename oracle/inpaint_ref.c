/* ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of cv2.inpaint(img, mask, 3, cv2.INPAINT_TELEA) for uint8 images with 1 or 3 channels, the call
 * /root/reference/ocr_agent/tools.py:617 makes in _apply_remove_lines.  The algorithm lives in opencv-python 4.13.0.92
 * (poetry.lock:3175), which is not vendored in the reference; this file restates the published algorithm (Telea 2004
 * as OpenCV's photo module implements it: fast marching from the mask boundary with a FIFO-stable priority queue, a
 * first march outwards to get negative distances in a `range`-wide ring, then the inward march that paints each pixel
 * from the known pixels of its disc) and is pinned against the installed wheel on random masks, ruled pages and edge
 * cases in tests/test_remove_lines.py, and against golden outputs of the unmodified reference (tests/golden).
 * Pinned at radius 3, the only radius the reference passes (8000 random cases equal to cv2); radius 2 is equal on every
 * case tried as well; at other radii a few pixels of flat regions, where the gradient term is pure rounding residue,
 * can differ from cv2 by 1-4 grey levels.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off; the float / double steps below matter bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { KNOWN = 0, BAND = 1, INSIDE = 2, CHANGE = 3 };

/* Priority queue ordered by (T, insertion number): among equal T the element pushed first leaves first, which is what
 * OpenCV's sorted list does (a push walks back from the tail past every strictly larger T). */
typedef struct {
  uint64_t *key; /* float bits of T (T >= 0) << 32 | insertion number */
  int32_t *pos;
  int n;
  uint32_t seq;
} pq_t;

static int pq_init(pq_t *q, int cap) {
  q->key = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(cap + 1));
  q->pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
  q->n = 0;
  q->seq = 0;
  return q->key && q->pos;
}
static void pq_free(pq_t *q) {
  free(q->key);
  free(q->pos);
}
static void pq_push(pq_t *q, int pos, float T) {
  uint32_t tb;
  memcpy(&tb, &T, 4);
  if (T == 0.0f) tb = 0; /* -0.0 never occurs, but keep the order total */
  uint64_t k = ((uint64_t)tb << 32) | q->seq++;
  int i = q->n++;
  while (i > 0) {
    int p = (i - 1) >> 1;
    if (q->key[p] <= k) break;
    q->key[i] = q->key[p];
    q->pos[i] = q->pos[p];
    i = p;
  }
  q->key[i] = k;
  q->pos[i] = pos;
}
static int pq_pop(pq_t *q) {
  if (q->n == 0) return -1;
  int top = q->pos[0];
  int n = --q->n;
  uint64_t k = q->key[n];
  int pv = q->pos[n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    if (c + 1 < n && q->key[c + 1] < q->key[c]) c++;
    if (q->key[c] >= k) break;
    q->key[i] = q->key[c];
    q->pos[i] = q->pos[c];
    i = c;
  }
  q->key[i] = k;
  q->pos[i] = pv;
  return top;
}

static float fmm_solve(const uint8_t *f, const float *t, int p1, int p2) {
  double a11 = t[p1], a22 = t[p2], m12 = a11 < a22 ? a11 : a22, sol;
  if (f[p1] != INSIDE) {
    if (f[p2] != INSIDE) {
      if (fabs(a11 - a22) >= 1.0)
        sol = 1 + m12;
      else
        sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
    } else
      sol = 1 + a11;
  } else if (f[p2] != INSIDE)
    sol = 1 + a22;
  else
    sol = 1 + m12;
  return (float)sol;
}

static float fmm_dist(const uint8_t *f, const float *t, int i, int j, int ec) {
  float d0 = fmm_solve(f, t, (i - 1) * ec + j, i * ec + j - 1);
  float d1 = fmm_solve(f, t, (i + 1) * ec + j, i * ec + j - 1);
  float d2 = fmm_solve(f, t, (i - 1) * ec + j, i * ec + j + 1);
  float d3 = fmm_solve(f, t, (i + 1) * ec + j, i * ec + j + 1);
  float a = d0 < d1 ? d0 : d1, b = d2 < d3 ? d2 : d3;
  return a < b ? a : b;
}

static const int DI[4] = {-1, 0, 1, 0}, DJ[4] = {0, -1, 0, 1};

/* Paint pixel (i, j) (extended coordinates) from the known pixels of its disc. */
static void telea_pixel(const uint8_t *f, const float *t, uint8_t *out, int er, int ec, int C, int range, int i, int j) {
  const int W = ec - 2;
  float gx, gy;
  if (f[i * ec + j + 1] != INSIDE) {
    if (f[i * ec + j - 1] != INSIDE)
      gx = (t[i * ec + j + 1] - t[i * ec + j - 1]) * 0.5f;
    else
      gx = t[i * ec + j + 1] - t[i * ec + j];
  } else {
    gx = f[i * ec + j - 1] != INSIDE ? t[i * ec + j] - t[i * ec + j - 1] : 0.f;
  }
  if (f[(i + 1) * ec + j] != INSIDE) {
    if (f[(i - 1) * ec + j] != INSIDE)
      gy = (t[(i + 1) * ec + j] - t[(i - 1) * ec + j]) * 0.5f;
    else
      gy = t[(i + 1) * ec + j] - t[i * ec + j];
  } else {
    gy = f[(i - 1) * ec + j] != INSIDE ? t[i * ec + j] - t[(i - 1) * ec + j] : 0.f;
  }
  float Ia[3] = {0, 0, 0}, Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, s[3] = {1.0e-20f, 1.0e-20f, 1.0e-20f};
  for (int k = i - range; k <= i + range; k++) {
    const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
    for (int l = j - range; l <= j + range; l++) {
      const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
      if (!(k > 0 && l > 0 && k < er - 1 && l < ec - 1)) continue;
      if (f[k * ec + l] == INSIDE || (l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
      const float ry = (float)(i - k), rx = (float)(j - l);
      const float len2 = rx * rx + ry * ry;
      const float dst = (float)(1. / (len2 * sqrt((double)len2)));
      const float lev = (float)(1. / (1 + fabs(t[k * ec + l] - t[i * ec + j]))); /* float difference, double from there */
      float dir = rx * gx + ry * gy;
      if (fabs(dir) <= 0.01) dir = 0.000001f;
      const float w = (float)fabs(dst * lev * dir);
      for (int c = 0; c < C; c++) {
#define PX(y, x) ((int)out[((size_t)(y) * W + (x)) * C + c])
        float gix, giy;
        if (f[k * ec + l + 1] != INSIDE) {
          if (f[k * ec + l - 1] != INSIDE)
            gix = (float)(PX(km, lp + 1) - PX(km, lm - 1)) * 2.0f;
          else
            gix = (float)(PX(km, lp + 1) - PX(km, lm));
        } else {
          gix = f[k * ec + l - 1] != INSIDE ? (float)(PX(km, lp) - PX(km, lm - 1)) : 0.f;
        }
        if (f[(k + 1) * ec + l] != INSIDE) {
          if (f[(k - 1) * ec + l] != INSIDE)
            giy = (float)(PX(kp + 1, lm) - PX(km - 1, lm)) * 2.0f;
          else
            giy = (float)(PX(kp + 1, lm) - PX(km, lm));
        } else {
          giy = f[(k - 1) * ec + l] != INSIDE ? (float)(PX(kp, lm) - PX(km - 1, lm)) : 0.f;
        }
        Ia[c] += w * (float)PX(k - 1, l - 1);
        Jx[c] -= w * (gix * rx);
        Jy[c] -= w * (giy * ry);
        s[c] += w;
#undef PX
      }
    }
  }
  for (int c = 0; c < C; c++) {
    /* all float: quotient, sum of squares, square root, second quotient, sum -- then + 0.5f and round half to even.
     * (Pinned at radius 3 on 8000 random images / masks against cv2 4.13; a double square root here differs from cv2
     * on about 1 image in 1000.) */
    const float sat = Ia[c] / s[c] + (Jx[c] + Jy[c]) / (sqrtf(Jx[c] * Jx[c] + Jy[c] * Jy[c]) + 1.0e-20f);
    long v = lrintf(sat + 0.5f); /* OpenCV adds 0.5 and THEN rounds half to even (saturate_cast<uchar>) */
    out[((size_t)(i - 1) * W + (j - 1)) * C + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
  }
}

/* img, out: uint8 [H, W, C] (C = 1 or 3); mask: uint8 [H, W], non-zero = repaint.  Returns 0, or -1 on bad arguments. */
int oracle_inpaint_telea_u8(const uint8_t *img, const uint8_t *mask, uint8_t *out, int H, int W, int C, int range) {
  if (!img || !mask || !out || H <= 0 || W <= 0 || (C != 1 && C != 3) || range < 1) return -1;
  const int er = H + 2, ec = W + 2, ne = er * ec;
  memcpy(out, img, (size_t)H * W * C);
  uint8_t *f = (uint8_t *)calloc(ne, 1), *band = (uint8_t *)calloc(ne, 1), *ring = (uint8_t *)calloc(ne, 1);
  float *t = (float *)malloc(sizeof(float) * ne);
  pq_t heap, outq;
  if (!f || !band || !ring || !t || !pq_init(&heap, ne) || !pq_init(&outq, ne)) return -1;
  for (int p = 0; p < ne; p++) t[p] = 1.0e6f;
  /* f: INSIDE on the mask; band: 4-neighbours of the mask that are not mask (T = 0), never on the 1-pixel frame */
  int any = 0;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      if (mask[(size_t)y * W + x]) f[(y + 1) * ec + x + 1] = INSIDE, any = 1;
  if (any) {
    for (int i = 1; i < er - 1; i++)
      for (int j = 1; j < ec - 1; j++) {
        const int p = i * ec + j;
        if (f[p] != INSIDE && (f[p - ec] == INSIDE || f[p + ec] == INSIDE || f[p - 1] == INSIDE || f[p + 1] == INSIDE)) {
          band[p] = 1;
          t[p] = 0.f;
        }
      }
    for (int p = 0; p < ne; p++)
      if (band[p]) {
        pq_push(&heap, p, 0.f);
        pq_push(&outq, p, 0.f);
      }
    /* ring: not mask, not band, within Chebyshev distance `range` of the mask, not on the frame */
    for (int i = 1; i < er - 1; i++)
      for (int j = 1; j < ec - 1; j++) {
        const int p = i * ec + j;
        if (f[p] == INSIDE || band[p]) continue;
        int hit = 0;
        for (int k = i - range; k <= i + range && !hit; k++)
          for (int l = j - range; l <= j + range; l++)
            if (k >= 0 && l >= 0 && k < er && l < ec && f[k * ec + l] == INSIDE) {
              hit = 1;
              break;
            }
        if (hit) ring[p] = INSIDE;
      }
    /* march outwards through the ring; distances become negative there */
    int p;
    while ((p = pq_pop(&outq)) >= 0) {
      ring[p] = CHANGE;
      const int ii = p / ec, jj = p % ec;
      for (int q = 0; q < 4; q++) {
        const int i = ii + DI[q], j = jj + DJ[q];
        if (i <= 0 || j <= 0 || i > er - 1 || j > ec - 1) continue;
        if (ring[i * ec + j] == INSIDE) {
          const float d = fmm_dist(ring, t, i, j, ec);
          t[i * ec + j] = d;
          ring[i * ec + j] = BAND;
          pq_push(&outq, i * ec + j, d);
        }
      }
    }
    for (int q = 0; q < ne; q++)
      if (ring[q] == CHANGE) t[q] = -t[q];
    /* march inwards, painting each pixel when the front reaches it */
    for (int q = 0; q < ne; q++)
      if (band[q]) f[q] = BAND;
    while ((p = pq_pop(&heap)) >= 0) {
      f[p] = KNOWN;
      const int ii = p / ec, jj = p % ec;
      for (int q = 0; q < 4; q++) {
        const int i = ii + DI[q], j = jj + DJ[q];
        if (i <= 0 || j <= 0 || i > er - 1 || j > ec - 1) continue;
        if (f[i * ec + j] == INSIDE) {
          const float d = fmm_dist(f, t, i, j, ec);
          t[i * ec + j] = d;
          telea_pixel(f, t, out, er, ec, C, range, i, j);
          f[i * ec + j] = BAND;
          pq_push(&heap, i * ec + j, d);
        }
      }
    }
  }
  free(f);
  free(band);
  free(ring);
  free(t);
  pq_free(&heap);
  pq_free(&outq);
  return 0;
}
