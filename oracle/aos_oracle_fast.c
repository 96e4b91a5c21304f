/*
 * aos_oracle_fast.c -- see aos_oracle_fast.h.  TEST INFRASTRUCTURE ONLY.
 * Citations as in aos_oracle_seed.c ("sg" = /root/reference/src/aos_seed_gen_node.cpp).
 */
#include "aos_oracle_fast.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define OCC 100

static int g_fast = 0, g_threads = 1, g_skip_labels = 0;
void orc_set_fast(int on, int threads, int skip_labels) {
  g_fast = on != 0;
  g_threads = threads < 1 ? 1 : (threads > 256 ? 256 : threads);
  g_skip_labels = skip_labels != 0;
}
int orc_fast_enabled(void) { return g_fast; }
int orc_fast_threads(void) { return g_threads; }
int orc_fast_skip_labels(void) { return g_fast && g_skip_labels; }

/* ---- parallel for ---------------------------------------------------------------------------------------- */
typedef struct { size_t lo, hi; orc_range_fn fn; void *arg; } pf_job;
static void *pf_run(void *v) {
  pf_job *j = (pf_job *)v;
  j->fn(j->lo, j->hi, j->arg);
  return NULL;
}
void orc_parallel_for(size_t n, orc_range_fn fn, void *arg) {
  int t = g_threads;
  if ((size_t)t > n) t = (int)(n ? n : 1);
  if (t <= 1) { fn(0, n, arg); return; }
  pthread_t th[256];
  pf_job job[256];
  size_t per = (n + (size_t)t - 1) / (size_t)t;
  int started = 0;
  for (int i = 0; i < t; ++i) {
    job[i].lo = (size_t)i * per;
    job[i].hi = job[i].lo + per > n ? n : job[i].lo + per;
    job[i].fn = fn; job[i].arg = arg;
    if (job[i].lo >= job[i].hi) break;
    if (i == t - 1 || job[i].hi == n) { fn(job[i].lo, job[i].hi, arg); break; }   /* last chunk on this thread */
    if (pthread_create(&th[i], NULL, pf_run, &job[i]) != 0) { fn(job[i].lo, job[i].hi, arg); continue; }
    started = i + 1;
  }
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* ---- binning (sg:459-538, 607-619): independent points, every writer stores the same value ----------------- */
typedef struct { const orc_seed_params *p; const float *pts; size_t stride; int w, h; double ox, oy; int8_t *grid;
                 float minx, maxx, miny, maxy; } bin_arg;
static void bin_range(size_t lo, size_t hi, void *v) {
  bin_arg *a = (bin_arg *)v;
  const orc_seed_params *p = a->p;
  const float res = p->grid_resolution;
  for (size_t i = lo; i < hi; ++i) {
    const float *pt = a->pts + i * a->stride;
    float x = pt[0], y = pt[1], z = pt[2];
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (z < p->clipping_minz || z > p->clipping_maxz) continue;
    if (x < a->minx || x > a->maxx) continue;
    if (y < a->miny || y > a->maxy) continue;
    int exclude = 0;
    for (int e = 0; e < p->n_excl; ++e) {
      float dx = x - p->excl[3 * e], dy = y - p->excl[3 * e + 1];
      float dist_sq = dx * dx + dy * dy, r = p->excl[3 * e + 2];
      if (dist_sq <= r * r) { exclude = 1; break; }
    }
    if (exclude) continue;
    int gx = (int)((x - a->ox) / res);
    int gy = (int)((y - a->oy) / res);
    if (gx >= 0 && gx < a->w && gy >= 0 && gy < a->h) a->grid[(size_t)gx + (size_t)gy * (size_t)a->w] = OCC;
  }
}
void orc_fast_bin_points(const orc_seed_params *p, const float *points, size_t n, size_t stride_floats, int w, int h,
                         double ox, double oy, int8_t *grid) {
  bin_arg a = {p, points, stride_floats, w, h, ox, oy, grid, 0, 0, 0, 0};
  orc_active_bounds(p, &a.minx, &a.maxx, &a.miny, &a.maxy);
  memset(grid, 0, (size_t)w * (size_t)h);
  orc_parallel_for(n, bin_range, &a);
}

/* ---- inflation (sg:933-967).  A cell whose four edge neighbours are occupied adds nothing: for d in the disc D,
 *      d != 0, stepping one cell towards the origin along a non-zero coordinate stays inside D, so
 *      c + D is covered by the discs of c's neighbours (R >= 1).  Writers only ever store OCC. ------------------ */
typedef struct { const int8_t *in; int8_t *out; int w, h, cells; } infl_arg;
static void infl_range(size_t lo, size_t hi, void *v) {
  infl_arg *a = (infl_arg *)v;
  const int w = a->w, h = a->h, R = a->cells;
  int *half = (int *)malloc(sizeof(int) * (size_t)(2 * R + 1));
  for (int dy = -R; dy <= R; ++dy) {          /* largest |dx| with dx^2 + dy^2 <= R^2 */
    int m = 0;
    while ((m + 1) * (m + 1) + dy * dy <= R * R) ++m;
    half[dy + R] = m;
  }
  for (size_t yy = lo; yy < hi; ++yy) {
    int y = (int)yy;
    const int8_t *row = a->in + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      if (row[x] != OCC) continue;
      if (R >= 1 && x > 0 && x < w - 1 && y > 0 && y < h - 1 && row[x - 1] == OCC && row[x + 1] == OCC && row[x - w] == OCC &&
          row[x + w] == OCC)
        continue;
      for (int dy = -R; dy <= R; ++dy) {
        int ny = y + dy;
        if (ny < 0 || ny >= h) continue;
        int m = half[dy + R];
        int x0 = x - m < 0 ? 0 : x - m, x1 = x + m >= w ? w - 1 : x + m;
        memset(a->out + (size_t)ny * w + x0, OCC, (size_t)(x1 - x0 + 1));
      }
    }
  }
  free(half);
}
void orc_fast_inflate(const int8_t *in, int w, int h, int cells, int8_t *out) {
  memcpy(out, in, (size_t)w * (size_t)h);
  if (cells < 0) return;   /* dx*dx+dy*dy > cells*cells never holds for an empty range: literal loop writes nothing */
  infl_arg a = {in, out, w, h, cells};
  orc_parallel_for((size_t)h, infl_range, &a);
}

/* ---- opening (sg:678-680), rows in parallel ---------------------------------------------------------------------- */
typedef struct { const int8_t *in; int8_t *er, *out; int w, h; } open_arg;
static void erode_range(size_t lo, size_t hi, void *v) {
  open_arg *a = (open_arg *)v;
  const int w = a->w, h = a->h;
  for (size_t yy = lo; yy < hi; ++yy)
    for (int x = 0; x < w; ++x) {
      int y = (int)yy;
      size_t i = (size_t)x + (size_t)y * w;
      int c = a->in[i] == OCC;
      if (c && x > 0) c = a->in[i - 1] == OCC;
      if (c && x < w - 1) c = a->in[i + 1] == OCC;
      if (c && y > 0) c = a->in[i - w] == OCC;
      if (c && y < h - 1) c = a->in[i + w] == OCC;
      a->er[i] = (int8_t)c;
    }
}
static void dilate_range(size_t lo, size_t hi, void *v) {
  open_arg *a = (open_arg *)v;
  const int w = a->w, h = a->h;
  for (size_t yy = lo; yy < hi; ++yy)
    for (int x = 0; x < w; ++x) {
      int y = (int)yy;
      size_t i = (size_t)x + (size_t)y * w;
      int c = a->er[i];
      if (!c && x > 0) c = a->er[i - 1];
      if (!c && x < w - 1) c = a->er[i + 1];
      if (!c && y > 0) c = a->er[i - w];
      if (!c && y < h - 1) c = a->er[i + w];
      a->out[i] = c ? OCC : 0;
    }
}
void orc_fast_open_cross(const int8_t *in, int w, int h, int8_t *out) {
  open_arg a = {in, (int8_t *)malloc((size_t)w * (size_t)h), out, w, h};
  orc_parallel_for((size_t)h, erode_range, &a);
  orc_parallel_for((size_t)h, dilate_range, &a);
  free(a.er);
}

/* ---- Zhang-Suen (sg:684): the marker of a sub-iteration is a pure function of the pre-sub-iteration image, so rows
 *      are marked in parallel, then applied in parallel.  Rows whose 3-row neighbourhood did not change in the last
 *      two sub-iterations cannot change in this one (each sub-iteration's rule depends only on the 3x3 window and the
 *      sub-iteration parity) and are skipped. ------------------------------------------------------------------------ */
typedef struct { uint8_t *img, *marker; int w, h, iter; const uint8_t *active; uint8_t *changed; } thin_arg;
static void thin_mark_range(size_t lo, size_t hi, void *v) {
  thin_arg *a = (thin_arg *)v;
  const int w = a->w, h = a->h, iter = a->iter;
  for (size_t ii = lo; ii < hi; ++ii) {
    int i = (int)ii;
    a->changed[i] = 0;
    if (i < 1 || i >= h - 1 || !a->active[i]) continue;
    uint8_t *mrow = a->marker + (size_t)i * w;
    int any = 0;
    for (int j = 1; j < w - 1; ++j) {
      const uint8_t *c = a->img + (size_t)i * w + j;
      if (!*c) continue;
      int p2 = c[-w], p3 = c[-w + 1], p4 = c[1], p5 = c[w + 1];
      int p6 = c[w], p7 = c[w - 1], p8 = c[-1], p9 = c[-w - 1];
      int A = (p2 == 0 && p3 == 1) + (p3 == 0 && p4 == 1) + (p4 == 0 && p5 == 1) + (p5 == 0 && p6 == 1) +
              (p6 == 0 && p7 == 1) + (p7 == 0 && p8 == 1) + (p8 == 0 && p9 == 1) + (p9 == 0 && p2 == 1);
      int B = p2 + p3 + p4 + p5 + p6 + p7 + p8 + p9;
      int m1 = iter == 0 ? (p2 * p4 * p6) : (p2 * p4 * p8);
      int m2 = iter == 0 ? (p4 * p6 * p8) : (p2 * p6 * p8);
      if (A == 1 && (B >= 2 && B <= 6) && m1 == 0 && m2 == 0) { mrow[j] = 1; any = 1; }
    }
    a->changed[i] = (uint8_t)any;
  }
}
static void thin_apply_range(size_t lo, size_t hi, void *v) {
  thin_arg *a = (thin_arg *)v;
  const int w = a->w;
  for (size_t ii = lo; ii < hi; ++ii) {
    if (!a->changed[ii]) continue;
    uint8_t *row = a->img + ii * (size_t)w, *mrow = a->marker + ii * (size_t)w;
    for (int j = 0; j < w; ++j)
      if (mrow[j]) { row[j] = 0; mrow[j] = 0; }
  }
}
typedef struct { const int8_t *g; uint8_t *img; int8_t *out; } conv_arg;
static void to_img_range(size_t lo, size_t hi, void *v) { conv_arg *a = (conv_arg *)v; for (size_t k = lo; k < hi; ++k) a->img[k] = a->g[k] == OCC; }
static void from_img_range(size_t lo, size_t hi, void *v) { conv_arg *a = (conv_arg *)v; for (size_t k = lo; k < hi; ++k) a->out[k] = a->img[k] ? OCC : 0; }
int orc_fast_thin_zhangsuen(int8_t *grid, int w, int h) {
  size_t n = (size_t)w * (size_t)h;
  uint8_t *img = (uint8_t *)malloc(n), *marker = (uint8_t *)calloc(n, 1);
  /* chg[k]: rows changed by the sub-iteration k steps back (k = 0: just now, 1: the one before) */
  uint8_t *chg0 = (uint8_t *)malloc((size_t)h + 2), *chg1 = (uint8_t *)malloc((size_t)h + 2), *active = (uint8_t *)malloc((size_t)h + 2);
  conv_arg ca = {grid, img, grid};
  orc_parallel_for(n, to_img_range, &ca);
  memset(chg0, 1, (size_t)h + 2);
  memset(chg1, 1, (size_t)h + 2);
  int passes = 0;
  for (;;) {
    int any_pass = 0;
    for (int iter = 0; iter < 2; ++iter) {
      /* a row can change now only if a row within +-1 changed in one of the last two sub-iterations: the row's
         outcome under THIS parity was last evaluated two sub-iterations ago, on an image that differs from the
         current one only by what those two sub-iterations deleted */
      for (int i = 0; i < h; ++i) {
        int a0 = chg0[i] | chg1[i];
        if (i > 0) a0 |= chg0[i - 1] | chg1[i - 1];
        if (i < h - 1) a0 |= chg0[i + 1] | chg1[i + 1];
        active[i] = (uint8_t)a0;
      }
      uint8_t *t = chg1; chg1 = chg0; chg0 = t;   /* chg0 now receives this sub-iteration's rows */
      thin_arg ta = {img, marker, w, h, iter, active, chg0};
      orc_parallel_for((size_t)h, thin_mark_range, &ta);
      orc_parallel_for((size_t)h, thin_apply_range, &ta);
      for (int i = 0; i < h; ++i) any_pass |= chg0[i];
    }
    ++passes;
    if (!any_pass) break;
  }
  orc_parallel_for(n, from_img_range, &ca);
  free(img); free(marker); free(chg0); free(chg1); free(active);
  return passes;
}

/* ---- cluster diameter (sg:1062-1074): the maximum is attained on the convex hull ------------------------------------ */
typedef struct { int x, y; } ipt;
static int ipt_cmp(const void *a, const void *b) {
  const ipt *p = (const ipt *)a, *q = (const ipt *)b;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  return p->y < q->y ? -1 : (p->y > q->y ? 1 : 0);
}
static int64_t cross3(ipt o, ipt a, ipt b) { return (int64_t)(a.x - o.x) * (b.y - o.y) - (int64_t)(a.y - o.y) * (b.x - o.x); }
int64_t orc_fast_max_pair_d2(const int32_t *cells, int n, int w) {
  if (n < 2) return 0;
  ipt *p = (ipt *)malloc(sizeof(ipt) * (size_t)n), *hull = (ipt *)malloc(sizeof(ipt) * (size_t)(2 * n + 2));
  for (int i = 0; i < n; ++i) { p[i].x = cells[i] % w; p[i].y = cells[i] / w; }
  qsort(p, (size_t)n, sizeof(ipt), ipt_cmp);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && cross3(hull[k - 2], hull[k - 1], p[i]) <= 0) --k;
    hull[k++] = p[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && cross3(hull[k - 2], hull[k - 1], p[i]) <= 0) --k;
    hull[k++] = p[i];
  }
  if (k > 1) --k;
  int64_t best = 0;
  for (int a = 0; a < k; ++a)
    for (int b = a + 1; b < k; ++b) {
      int64_t dx = hull[a].x - hull[b].x, dy = hull[a].y - hull[b].y, d2 = dx * dx + dy * dy;
      if (d2 > best) best = d2;
    }
  free(p); free(hull);
  return best;
}

/* ---- hash grid ---------------------------------------------------------------------------------------------- */
static size_t sg_slot(const orc_sgrid *g, int64_t key) {
  uint64_t hsh = (uint64_t)key * 0x9E3779B97F4A7C15ull;
  size_t i = (size_t)(hsh >> 17) & (g->cap - 1);
  while (g->head[i] != -1 && g->key[i] != key) i = (i + 1) & (g->cap - 1);
  return i;
}
static int64_t sg_key(int64_t cx, int64_t cy) { return (cx << 32) ^ (int64_t)(uint32_t)cy; }
static void sg_alloc(orc_sgrid *g, size_t cap) {
  g->cap = cap;
  g->key = (int64_t *)malloc(sizeof(int64_t) * cap);
  g->head = (int32_t *)malloc(sizeof(int32_t) * cap);
  for (size_t i = 0; i < cap; ++i) g->head[i] = -1;
}
void orc_sgrid_init(orc_sgrid *g, double cell, int expected_points) {
  memset(g, 0, sizeof(*g));
  g->cell = cell;
  size_t cap = 64;
  while (cap < (size_t)(expected_points > 0 ? expected_points : 1) * 2) cap <<= 1;
  sg_alloc(g, cap);
  g->ncap = expected_points > 16 ? expected_points : 16;
  g->next = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->ncap);
  g->xy = (double *)malloc(sizeof(double) * 2 * (size_t)g->ncap);
}
void orc_sgrid_free(orc_sgrid *g) {
  free(g->key); free(g->head); free(g->next); free(g->xy);
  memset(g, 0, sizeof(*g));
}
static void sg_cell_of(const orc_sgrid *g, double x, double y, int64_t *cx, int64_t *cy) {
  double fx = floor(x / g->cell), fy = floor(y / g->cell);
  /* far-away or non-finite coordinates all land in one overflow cell; they are only ever compared by distance */
  if (!(fx > -1e9 && fx < 1e9)) fx = 2e9;
  if (!(fy > -1e9 && fy < 1e9)) fy = 2e9;
  *cx = (int64_t)fx;
  *cy = (int64_t)fy;
}
int orc_sgrid_add(orc_sgrid *g, double x, double y) {
  if (g->n == g->ncap) {
    g->ncap *= 2;
    g->next = (int32_t *)realloc(g->next, sizeof(int32_t) * (size_t)g->ncap);
    g->xy = (double *)realloc(g->xy, sizeof(double) * 2 * (size_t)g->ncap);
  }
  if ((size_t)(g->n + 1) * 2 > g->cap) {   /* rehash */
    int64_t *ok = g->key; int32_t *oh = g->head; size_t ocap = g->cap;
    sg_alloc(g, ocap * 2);
    for (size_t i = 0; i < ocap; ++i)
      if (oh[i] != -1) { size_t s = sg_slot(g, ok[i]); g->key[s] = ok[i]; g->head[s] = oh[i]; }
    free(ok); free(oh);
  }
  int idx = g->n++;
  g->xy[2 * idx] = x;
  g->xy[2 * idx + 1] = y;
  int64_t cx, cy;
  sg_cell_of(g, x, y, &cx, &cy);
  int64_t key = sg_key(cx, cy);
  size_t s = sg_slot(g, key);
  g->next[idx] = g->head[s];
  g->key[s] = key;
  g->head[s] = idx;
  return idx;
}
int orc_sgrid_gather(const orc_sgrid *g, double x, double y, int rings, int32_t *out, int max_out) {
  int64_t cx, cy;
  sg_cell_of(g, x, y, &cx, &cy);
  int cnt = 0;
  for (int64_t dy = -rings; dy <= rings; ++dy)
    for (int64_t dx = -rings; dx <= rings; ++dx) {
      size_t s = sg_slot(g, sg_key(cx + dx, cy + dy));
      for (int32_t i = g->head[s]; i != -1; i = g->next[i]) {
        if (out && cnt < max_out) out[cnt] = i;
        ++cnt;
      }
    }
  return cnt;
}
static int i32_cmp(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
void orc_sort_i32(int32_t *a, int n) { if (n > 1) qsort(a, (size_t)n, sizeof(int32_t), i32_cmp); }
