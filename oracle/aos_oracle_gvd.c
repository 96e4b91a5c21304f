/*
 * aos_oracle_gvd.c -- CPU restatement of aos_gvd_node's processGraph path around cv::Subdiv2D.
 * TEST INFRASTRUCTURE ONLY (see aos_oracle.h).  Compile with -O2 -ffp-contract=off.
 * Citations: /root/reference/src/aos_gvd_node.cpp ("gvd"), src/utils/voronoi_diagram.cpp ("vd").
 * The Delaunay/Voronoi arithmetic itself is OpenCV's (cv::Subdiv2D, call sites vd:63,84,94); the
 * Python glue in oracle/subdiv.py calls the real cv2.Subdiv2D and hands the facets to this file.
 */
#include "aos_oracle.h"
#include "aos_oracle_fast.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OCC 100

/* ---- voronoiSeedsCallback, gvd:84-128 ----------------------------------------------------- */
/* indexed variant: the members of seed i's group are the unused j > i within 0.5 m, visited in ascending j (the
 * summation order); a 0.5 m hash grid yields the same candidates, sorted before use */
static int merge_seeds_fast(const double *seeds, int n, double *out) {
  const double merge_distance = 0.5;
  uint8_t *used = (uint8_t *)calloc((size_t)n + 1, 1);
  orc_sgrid g;
  orc_sgrid_init(&g, 0.5005, n);
  for (int i = 0; i < n; ++i) orc_sgrid_add(&g, seeds[2 * i], seeds[2 * i + 1]);
  int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    used[i] = 1;
    double sx = seeds[2 * i], sy = seeds[2 * i + 1];
    int cnt = 1;
    if (isfinite(seeds[2 * i]) && isfinite(seeds[2 * i + 1])) { /* a non-finite seed is within 0.5 m of nothing */
      int k = orc_sgrid_gather(&g, seeds[2 * i], seeds[2 * i + 1], 1, cand, n);
      orc_sort_i32(cand, k);
      for (int c = 0; c < k; ++c) {
        int j = cand[c];
        if (j <= i || used[j]) continue;
        double dx = seeds[2 * i] - seeds[2 * j], dy = seeds[2 * i + 1] - seeds[2 * j + 1];
        double dist = sqrt(dx * dx + dy * dy);
        if (dist <= merge_distance) {
          used[j] = 1;
          sx += seeds[2 * j];
          sy += seeds[2 * j + 1];
          cnt++;
        }
      }
    }
    out[2 * m] = sx / (double)cnt;
    out[2 * m + 1] = sy / (double)cnt;
    m++;
  }
  free(cand);
  free(used);
  orc_sgrid_free(&g);
  return m;
}

int orc_gvd_merge_seeds(const double *seeds, int n, double *out) {
  if (orc_fast_enabled()) return merge_seeds_fast(seeds, n, out);
  const double merge_distance = 0.5;
  uint8_t *used = (uint8_t *)calloc((size_t)n + 1, 1);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    used[i] = 1;
    double sx = 0.0, sy = 0.0;
    int cnt = 0;
    sx += seeds[2 * i];
    sy += seeds[2 * i + 1];
    cnt = 1;
    for (int j = i + 1; j < n; ++j) {
      if (used[j]) continue;
      double dx = seeds[2 * i] - seeds[2 * j], dy = seeds[2 * i + 1] - seeds[2 * j + 1];
      double dist = sqrt(dx * dx + dy * dy);
      if (dist <= merge_distance) {
        used[j] = 1;
        sx += seeds[2 * j];
        sy += seeds[2 * j + 1];
        cnt++;
      }
    }
    out[2 * m] = sx / (double)cnt;
    out[2 * m + 1] = sy / (double)cnt;
    m++;
  }
  free(used);
  return m;
}

/* ---- VoronoiDiagram::compute up to subdiv.insert, vd:16-89 -------------------------------- */
static int cv_round(double v) { return (int)lrint(v); } /* cvRound: round-half-even */

void orc_gvd_subdiv_inputs(const double *seeds, int n, double min_x, double max_x, double min_y,
                           double max_y, int *rect_i, float *rect_f, float *pts, uint8_t *keep,
                           int *valid) {
  *valid = 0;
  if (n == 0) return;
  if (!isfinite(min_x) || !isfinite(max_x) || !isfinite(min_y) || !isfinite(max_y)) return;
  if (min_x > max_x) { double t = min_x; min_x = max_x; max_x = t; }
  if (min_y > max_y) { double t = min_y; min_y = max_y; max_y = t; }
  const double min_size = 1.0;
  if (max_x - min_x < min_size) {
    double c = (min_x + max_x) / 2.0;
    min_x = c - min_size / 2.0;
    max_x = c + min_size / 2.0;
  }
  if (max_y - min_y < min_size) {
    double c = (min_y + max_y) / 2.0;
    min_y = c - min_size / 2.0;
    max_y = c + min_size / 2.0;
  }
  float rx = (float)(min_x - 1.0), ry = (float)(min_y - 1.0);
  float rw = (float)(fabs(max_x - min_x) + 2.0), rh = (float)(fabs(max_y - min_y) + 2.0);
  rect_f[0] = rx; rect_f[1] = ry; rect_f[2] = rw; rect_f[3] = rh;
  if (rw <= 0 || rh <= 0) return;
  /* OpenCV 4.5.4: Subdiv2D(Rect) -- the Rect2f converts through saturate_cast<int> (cvRound) */
  rect_i[0] = cv_round(rx); rect_i[1] = cv_round(ry); rect_i[2] = cv_round(rw); rect_i[3] = cv_round(rh);
  const float margin = 0.1f;
  for (int i = 0; i < n; ++i) {
    double sx = seeds[2 * i], sy = seeds[2 * i + 1];
    if (!isfinite(sx) || !isfinite(sy)) {
      keep[i] = 0;
      pts[2 * i] = pts[2 * i + 1] = 0.f;
      continue;
    }
    float x = (float)sx, y = (float)sy;
    x = fmaxf(rx + margin, fminf(rx + rw - margin, x));
    y = fmaxf(ry + margin, fminf(ry + rh - margin, y));
    pts[2 * i] = x;
    pts[2 * i + 1] = y;
    keep[i] = 1;
  }
  *valid = 1;
}

/* ---- helpers -------------------------------------------------------------------------------- */
typedef struct { double x, y; } v2;
static double nrm(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

typedef struct {
  const int8_t *data;
  int w, h;
  double ox, oy;
  float res;
} gridview;

static void bounds(const gridview *g, double *minx, double *maxx, double *miny, double *maxy) {
  *minx = g->ox; /* gvd:278-281: uint32 * float -> float */
  *maxx = g->ox + (float)((float)(unsigned)g->w * g->res);
  *miny = g->oy;
  *maxy = g->oy + (float)((float)(unsigned)g->h * g->res);
}

/* ---- edgePassesThroughOccupiedPixels, gvd:320-359 ------------------------------------------ */
static int edge_hits(const gridview *g, v2 s, v2 e) {
  const double resolution = g->res;
  double ex = e.x - s.x, ey = e.y - s.y;
  double edge_length = nrm(ex, ey);
  if (edge_length < 1e-6) return 0;
  const double sample_step = resolution * 0.5;
  int num_samples = (int)(edge_length / sample_step) + 1;
  double z = ex * ex + ey * ey, dx = ex, dy = ey; /* normalized() */
  if (z > 0) {
    double q = sqrt(z);
    dx = ex / q;
    dy = ey / q;
  }
  for (int i = 0; i <= num_samples; ++i) {
    double t = (i == num_samples) ? 1.0 : ((double)i / (double)num_samples);
    /* start + t * dir * edge_length : ((t*dir) * edge_length) componentwise */
    double px = s.x + (t * dx) * edge_length, py = s.y + (t * dy) * edge_length;
    int mx = (int)((px - g->ox) / resolution);
    int my = (int)((py - g->oy) / resolution);
    if (mx >= 0 && mx < g->w && my >= 0 && my < g->h)
      if (g->data[(size_t)mx + (size_t)my * g->w] == OCC) return 1;
  }
  return 0;
}

/* tiny open-addressing set of int64 keys (std::unordered_set<int64_t> added_edges) */
typedef struct { int64_t *k; uint8_t *u; size_t cap, n; } i64set;
static void set_init(i64set *s, size_t cap) {
  size_t c = 64;
  while (c < cap * 2) c <<= 1;
  s->cap = c; s->n = 0;
  s->k = (int64_t *)malloc(sizeof(int64_t) * c);
  s->u = (uint8_t *)calloc(c, 1);
}
static size_t set_slot(const i64set *s, int64_t key) {
  uint64_t hsh = (uint64_t)key * 0x9E3779B97F4A7C15ull;
  size_t i = (size_t)(hsh >> 20) & (s->cap - 1);
  while (s->u[i] && s->k[i] != key) i = (i + 1) & (s->cap - 1);
  return i;
}
static void set_grow(i64set *s) {
  i64set t;
  set_init(&t, s->cap);
  for (size_t i = 0; i < s->cap; ++i)
    if (s->u[i]) { size_t j = set_slot(&t, s->k[i]); t.u[j] = 1; t.k[j] = s->k[i]; t.n++; }
  free(s->k); free(s->u);
  *s = t;
}
static int set_has(const i64set *s, int64_t key) { return s->u[set_slot(s, key)]; }
static void set_add(i64set *s, int64_t key) {
  if ((s->n + 1) * 2 > s->cap) set_grow(s);
  size_t i = set_slot(s, key);
  if (!s->u[i]) { s->u[i] = 1; s->k[i] = key; s->n++; }
}

typedef struct { int from, to; double length_m; float clearance; } edge_rec;

/* ---- castRay, gvd:558-684 ------------------------------------------------------------------ */
static v2 cast_ray(const gridview *g, v2 sp, v2 other, double angle_offset_deg, double min_distance) {
  double ex = other.x - sp.x, ey = other.y - sp.y;
  double d = nrm(ex, ey);
  if (d < 1e-6) { ex = 1.0; ey = 0.0; }
  else { double z = ex * ex + ey * ey; if (z > 0) { double q = sqrt(z); ex /= q; ey /= q; } }
  double outx = -ex, outy = -ey, perpx = -ey, perpy = ex;
  double a = angle_offset_deg * M_PI / 180.0;
  double rdx, rdy;
  if (angle_offset_deg > 0) { rdx = cos(a) * outx + sin(a) * perpx; rdy = cos(a) * outy + sin(a) * perpy; }
  else { rdx = cos(-a) * outx + sin(-a) * (-perpx); rdy = cos(-a) * outy + sin(-a) * (-perpy); }
  { double z = rdx * rdx + rdy * rdy; if (z > 0) { double q = sqrt(z); rdx /= q; rdy /= q; } }
  double minx, maxx, miny, maxy;
  bounds(g, &minx, &maxx, &miny, &maxy);
  const double resolution = g->res;
  double step_size = g->res * 0.5; /* float * double -> double, gvd:624 */
  if (step_size < 0.01) step_size = 0.01;
  double current = min_distance;
  double gw = (float)((float)(unsigned)g->w * g->res), gh = (float)((float)(unsigned)g->h * g->res);
  double abs_max = sqrt(gw * gw + gh * gh) * 3.0;
  v2 r;
  while (current <= abs_max) {
    double px = sp.x + rdx * current, py = sp.y + rdy * current;
    if (!(px >= minx && px <= maxx && py >= miny && py <= maxy)) {
      r.x = fmax(minx, fmin(maxx, px));
      r.y = fmax(miny, fmin(maxy, py));
      return r;
    }
    int mx = (int)((px - g->ox) / resolution), my = (int)((py - g->oy) / resolution);
    if (mx >= 0 && mx < g->w && my >= 0 && my < g->h && g->data[(size_t)mx + (size_t)my * g->w] == OCC) {
      r.x = px; r.y = py;
      return r;
    }
    current += step_size;
  }
  r.x = sp.x + rdx * abs_max;
  r.y = sp.y + rdy * abs_max;
  if (!(r.x >= minx && r.x <= maxx && r.y >= miny && r.y <= maxy)) {
    r.x = fmax(minx, fmin(maxx, r.x));
    r.y = fmax(miny, fmin(maxy, r.y));
  }
  return r;
}

/* ---- findVoronoiBoundaryPointNearEndpoint, gvd:686-790 -------------------------------------- */
static const orc_sgrid *g_corner_grid = NULL; /* fast mode: 1 m hash grid over the cropped nodes, in node order */

static v2 find_corner(const gridview *g, const v2 *nodes, int M, v2 endpoint, v2 other,
                      double target_angle_deg, double min_distance, double max_distance) {
  double mx = other.x - endpoint.x, my = other.y - endpoint.y;
  double ml = nrm(mx, my);
  if (ml < 1e-6) { mx = 1.0; my = 0.0; }
  else { double z = mx * mx + my * my; if (z > 0) { double q = sqrt(z); mx /= q; my /= q; } }
  double outx = -mx, outy = -my, perpx = -my, perpy = mx;
  int neg = fabs(target_angle_deg - (-90.0)) < 1e-6, pos = fabs(target_angle_deg - 90.0) < 1e-6;
  double gw = (float)((float)(unsigned)g->w * g->res), gh = (float)((float)(unsigned)g->h * g->res);
  double radii[4] = {max_distance, 7.0, 9.0, sqrt(gw * gw + gh * gh) * 2.0};
  for (int ri = 0; ri < 4; ++ri) {
    double search_radius = radii[ri];
    double best = DBL_MAX;
    int best_i = -1;
    /* fast mode: only nodes in the grid cells the search disc touches can pass `dist <= search_radius`; the winner is
       the smallest distance, the lowest index among equals -- the same as the first strict minimum in node order */
    int32_t *cand = NULL;
    int ncand = -1;
    if (g_corner_grid && search_radius <= 16.0 && isfinite(endpoint.x) && isfinite(endpoint.y)) {
      int rings = (int)ceil(search_radius / g_corner_grid->cell) + 1;
      ncand = orc_sgrid_gather(g_corner_grid, endpoint.x, endpoint.y, rings, NULL, 0);
      cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncand + 1));
      orc_sgrid_gather(g_corner_grid, endpoint.x, endpoint.y, rings, cand, ncand);
      orc_sort_i32(cand, ncand);
    }
    for (int ii = 0; ii < (ncand >= 0 ? ncand : M); ++ii) {
      int i = ncand >= 0 ? cand[ii] : ii;
      double dx = nodes[i].x - endpoint.x, dy = nodes[i].y - endpoint.y;
      double dist = nrm(dx, dy);
      if (dist < min_distance || dist > search_radius) continue;
      { double z = dx * dx + dy * dy; if (z > 0) { double q = sqrt(z); dx /= q; dy /= q; } }
      double dot_out = outx * dx + outy * dy;
      if (dot_out < 0.0) continue;
      double dot_perp = perpx * dx + perpy * dy;
      if (neg) { if (dot_perp > 0.0) continue; }
      else if (pos) { if (dot_perp < 0.0) continue; }
      /* candidate; "closest among candidates", strict <, candidate order = node order */
      if (dist < best) { best = dist; best_i = i; }
    }
    free(cand);
    if (best_i >= 0) return nodes[best_i];
  }
  return cast_ray(g, endpoint, other, target_angle_deg, min_distance);
}

/* nearest node of q by (distance, index): ring r of 5 cm cells is searched while a closer node could still be there
 * (everything outside the (2r+1)^2 block is farther than r cells) */
static int nearest_node_fast(const orc_sgrid *g, const v2 *bp, int M, v2 q) {
  if (!isfinite(q.x) || !isfinite(q.y) || fabs(q.x) > 1e8 || fabs(q.y) > 1e8) {
    int si = -1;
    double md = DBL_MAX;
    for (int i = 0; i < M; ++i) { double d = nrm(bp[i].x - q.x, bp[i].y - q.y); if (d < md) { md = d; si = i; } }
    return si;
  }
  for (int rings = 1; rings <= 64; rings *= 2) {
    int32_t cand[512];
    int m = orc_sgrid_gather(g, q.x, q.y, rings, cand, 512);
    if (m > 512) break;
    double md = DBL_MAX;
    int si = -1;
    for (int k = 0; k < m; ++k) {
      int i = cand[k];
      double d = nrm(bp[i].x - q.x, bp[i].y - q.y);
      if (d < md || (d == md && i < si)) { md = d; si = i; }
    }
    if (si >= 0 && md < (double)rings * 0.05) return si;   /* nothing outside the block can be closer or equal */
  }
  int si = -1;
  double md = DBL_MAX;
  for (int i = 0; i < M; ++i) { double d = nrm(bp[i].x - q.x, bp[i].y - q.y); if (d < md) { md = d; si = i; } }
  return si;
}

int orc_gvd_graph(const float *facet_xy, const int32_t *facet_off, int n_facets,
                  const int8_t *skel_framed, int w, int h, double origin_x, double origin_y,
                  float res, const double *rows_info, int n_rows, orc_graph *out) {
  memset(out, 0, sizeof(*out));
  gridview g = {skel_framed, w, h, origin_x, origin_y, res};

  /* facets -> edges, vd:97-114 */
  int n_edges = 0;
  for (int f = 0; f < n_facets; ++f) {
    int k = facet_off[f + 1] - facet_off[f];
    if (k >= 2) n_edges += k;
  }
  v2 *es = (v2 *)malloc(sizeof(v2) * (size_t)(n_edges + 1));
  v2 *ee = (v2 *)malloc(sizeof(v2) * (size_t)(n_edges + 1));
  int ne = 0;
  for (int f = 0; f < n_facets; ++f) {
    int b = facet_off[f], k = facet_off[f + 1] - b;
    if (k < 2) continue;
    for (int i = 0; i < k; ++i) {
      int j = (i + 1) % k;
      es[ne].x = facet_xy[2 * (b + i)]; es[ne].y = facet_xy[2 * (b + i) + 1];
      ee[ne].x = facet_xy[2 * (b + j)]; ee[ne].y = facet_xy[2 * (b + j) + 1];
      ne++;
    }
  }
  out->n_voro_edges = ne;

  /* extractBoundaryPoints, vd:149-207 (first-come, int key + 5 cm) */
  v2 *bp = (v2 *)malloc(sizeof(v2) * (size_t)(2 * ne + 1));
  int M = 0;
  i64set keys;
  set_init(&keys, (size_t)ne + 16);
  const double threshold = 0.05;
  const int fast = orc_fast_enabled();
  orc_sgrid bpgrid;
  memset(&bpgrid, 0, sizeof(bpgrid));
  if (fast) orc_sgrid_init(&bpgrid, 0.05005, ne + 16);
  for (int e = 0; e < ne; ++e) {
    for (int side = 0; side < 2; ++side) {
      v2 q = side == 0 ? es[e] : ee[e];
      int ix = (int)(q.x * 100), iy = (int)(q.y * 100);
      int64_t key = ((int64_t)ix << 32) ^ (uint32_t)iy;
      if (set_has(&keys, key)) continue;
      int too_close = 0;
      if (fast) { /* existence of an accepted point closer than 5 cm: the 3x3 block of 5 cm cells holds them all */
        int32_t cand[64];
        int m = orc_sgrid_gather(&bpgrid, q.x, q.y, 1, cand, 64);
        if (m > 64) m = -1;
        for (int k = 0; k < m; ++k) {
          double dx = bp[cand[k]].x - q.x, dy = bp[cand[k]].y - q.y;
          if (dx * dx + dy * dy < threshold * threshold) { too_close = 1; break; }
        }
        if (m < 0)
          for (int i = 0; i < M; ++i) {
            double dx = bp[i].x - q.x, dy = bp[i].y - q.y;
            if (dx * dx + dy * dy < threshold * threshold) { too_close = 1; break; }
          }
      } else
      for (int i = 0; i < M; ++i) {
        double dx = bp[i].x - q.x, dy = bp[i].y - q.y;
        if (dx * dx + dy * dy < threshold * threshold) { too_close = 1; break; }
      }
      if (!too_close) { set_add(&keys, key); bp[M++] = q; if (fast) orc_sgrid_add(&bpgrid, q.x, q.y); }
    }
  }
  free(keys.k); free(keys.u);
  out->n_boundary_points_precrop = M;

  /* buildGraphFromBoundaryPoints, gvd:794-895 */
  edge_rec *rec = (edge_rec *)malloc(sizeof(edge_rec) * 16);
  int nrec = 0, reccap = 16;
  i64set added;
  set_init(&added, (size_t)ne + 16);
  if (M > 0 && ne > 0) {
    for (int e = 0; e < ne; ++e) {
      int si = -1, ei = -1;
      double md = DBL_MAX;
      if (fast) {
        si = nearest_node_fast(&bpgrid, bp, M, es[e]);
        ei = nearest_node_fast(&bpgrid, bp, M, ee[e]);
      } else {
      for (int i = 0; i < M; ++i) { double d = nrm(bp[i].x - es[e].x, bp[i].y - es[e].y); if (d < md) { md = d; si = i; } }
      md = DBL_MAX;
      for (int i = 0; i < M; ++i) { double d = nrm(bp[i].x - ee[e].x, bp[i].y - ee[e].y); if (d < md) { md = d; ei = i; } }
      }
      if (si >= 0 && ei >= 0 && si != ei) {
        int a = si, b = ei;
        if (a > b) { int t = a; a = b; b = t; }
        int64_t key = ((int64_t)a << 32) ^ (uint32_t)b;
        if (!set_has(&added, key)) {
          if (edge_hits(&g, bp[si], bp[ei])) continue;
          set_add(&added, key);
          if (nrec == reccap) { reccap *= 2; rec = (edge_rec *)realloc(rec, sizeof(edge_rec) * reccap); }
          rec[nrec].from = a; rec[nrec].to = b;
          rec[nrec].length_m = nrm(bp[ei].x - bp[si].x, bp[ei].y - bp[si].y);
          rec[nrec].clearance = 0.0f;
          nrec++;
        }
      }
    }
    const double nearby = 0.5;
    orc_sgrid pgrid;
    memset(&pgrid, 0, sizeof(pgrid));
    int32_t *pc = NULL;
    if (fast) {
      orc_sgrid_init(&pgrid, 0.5005, M + 16);
      for (int i = 0; i < M; ++i) orc_sgrid_add(&pgrid, bp[i].x, bp[i].y);
      pc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M + 1));
    }
    for (int i = 0; i < M; ++i) {
      /* fast mode: the j > i within 0.5 m all sit in the 3x3 block of 0.5 m cells; visited in ascending j as the
         double loop does (the order decides the order of the appended edges) */
      int nj = fast ? orc_sgrid_gather(&pgrid, bp[i].x, bp[i].y, 1, pc, M) : M - (i + 1);
      if (fast) orc_sort_i32(pc, nj);
      for (int jj = 0; jj < nj; ++jj) {
        int j = fast ? pc[jj] : i + 1 + jj;
        if (j <= i) continue;
        double dist = nrm(bp[i].x - bp[j].x, bp[i].y - bp[j].y);
        if (dist <= nearby && dist > 1e-6) {
          int64_t key = ((int64_t)i << 32) ^ (uint32_t)j;
          if (!set_has(&added, key)) {
            if (edge_hits(&g, bp[i], bp[j])) continue;
            set_add(&added, key);
            if (nrec == reccap) { reccap *= 2; rec = (edge_rec *)realloc(rec, sizeof(edge_rec) * reccap); }
            rec[nrec].from = i; rec[nrec].to = j; rec[nrec].length_m = dist; rec[nrec].clearance = 0.0f;
            nrec++;
          }
        }
      }
    }
    if (fast) { orc_sgrid_free(&pgrid); free(pc); }
  }
  if (fast) orc_sgrid_free(&bpgrid);
  free(added.k); free(added.u);
  free(es); free(ee);

  /* filterNodesAndEdgesOutsideGrid, gvd:420-483 */
  double minx, maxx, miny, maxy;
  bounds(&g, &minx, &maxx, &miny, &maxy);
  int *remap = (int *)malloc(sizeof(int) * (size_t)(M + 1));
  v2 *nodes = (v2 *)malloc(sizeof(v2) * (size_t)(M + 1));
  int N = 0;
  for (int i = 0; i < M; ++i) {
    if (bp[i].x >= minx && bp[i].x <= maxx && bp[i].y >= miny && bp[i].y <= maxy) { nodes[N] = bp[i]; remap[i] = N++; }
    else remap[i] = -1;
  }
  edge_rec *frec = (edge_rec *)malloc(sizeof(edge_rec) * (size_t)(nrec + 1));
  int nf = 0;
  for (int k = 0; k < nrec; ++k) {
    int a = remap[rec[k].from], b = remap[rec[k].to];
    if (a >= 0 && b >= 0 && a != b) {
      v2 fp = nodes[a], tp = nodes[b];
      if (a > b) { int t = a; a = b; b = t; }
      frec[nf].from = a; frec[nf].to = b;
      frec[nf].length_m = nrm(tp.x - fp.x, tp.y - fp.y);
      frec[nf].clearance = rec[k].clearance;
      nf++;
    }
  }
  free(rec); free(remap); free(bp);

  /* findClusterEndpointVoronoiBoundaryPoints, gvd:485-556; rows arrive as {start,end} pairs and are
   * swapped so that ep1.x <= ep2.x (gvd:135-146) */
  out->corner_points = (double *)malloc(sizeof(double) * 8 * (size_t)(n_rows + 1));
  int have_corners = N > 0;
  orc_sgrid cgrid;
  memset(&cgrid, 0, sizeof(cgrid));
  if (fast && N > 0) {
    orc_sgrid_init(&cgrid, 1.0, N + 16);
    for (int i = 0; i < N; ++i) orc_sgrid_add(&cgrid, nodes[i].x, nodes[i].y);
    g_corner_grid = &cgrid;
  }
  for (int r = 0; r < n_rows && have_corners; ++r) {
    v2 s = {rows_info[4 * r], rows_info[4 * r + 1]}, e = {rows_info[4 * r + 2], rows_info[4 * r + 3]};
    if (s.x > e.x) { v2 t = s; s = e; e = t; }
    v2 c[4];
    c[0] = find_corner(&g, nodes, N, s, e, -90.0, 0.5, 5.0); /* TL */
    c[1] = find_corner(&g, nodes, N, s, e, 90.0, 0.5, 5.0);  /* TR */
    c[2] = find_corner(&g, nodes, N, e, s, -90.0, 0.5, 5.0); /* BL */
    c[3] = find_corner(&g, nodes, N, e, s, 90.0, 0.5, 5.0);  /* BR */
    for (int k = 0; k < 4; ++k) { out->corner_points[8 * r + 2 * k] = c[k].x; out->corner_points[8 * r + 2 * k + 1] = c[k].y; }
  }
  int n_corner_rows = have_corners ? n_rows : 0;
  if (g_corner_grid) { g_corner_grid = NULL; orc_sgrid_free(&cgrid); }
  /* fast mode: corner points in a 0.1 m hash grid, index = 4 * row + corner, so ascending index = the (row, corner)
     order of the double loop below */
  orc_sgrid lgrid;
  memset(&lgrid, 0, sizeof(lgrid));
  if (fast) {
    orc_sgrid_init(&lgrid, 0.1001, 4 * n_corner_rows + 16);
    for (int r = 0; r < n_corner_rows; ++r)
      for (int k = 0; k < 4; ++k) orc_sgrid_add(&lgrid, out->corner_points[8 * r + 2 * k], out->corner_points[8 * r + 2 * k + 1]);
  }

  /* publishGraph, gvd:897-1010 */
  out->n_nodes = N;
  out->nodes = (double *)malloc(sizeof(double) * 2 * (size_t)(N + 1));
  out->node_labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
  out->node_cluster_indices = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
  out->node_label_counts = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
  int lcap = 64, nl = 0;
  out->node_label_clusters = (int32_t *)malloc(sizeof(int32_t) * lcap);
  out->node_label_types = (int32_t *)malloc(sizeof(int32_t) * lcap);
  const double tol = 0.1;
  for (int i = 0; i < N; ++i) {
    out->nodes[2 * i] = nodes[i].x;
    out->nodes[2 * i + 1] = nodes[i].y;
    int mask = 0, cidx = -1, cnt = 0;
    int32_t lc[64];
    int nlc = fast ? orc_sgrid_gather(&lgrid, nodes[i].x, nodes[i].y, 1, lc, 64) : 4 * n_corner_rows;
    int use_lc = fast && nlc <= 64;
    if (use_lc) orc_sort_i32(lc, nlc);
    else nlc = 4 * n_corner_rows;
    for (int q = 0; q < nlc; ++q) {
      {
        int r = (use_lc ? lc[q] : q) / 4, k = (use_lc ? lc[q] : q) % 4;
        double d = nrm(nodes[i].x - out->corner_points[8 * r + 2 * k], nodes[i].y - out->corner_points[8 * r + 2 * k + 1]);
        if (d < tol) {
          mask |= 1 << k;
          if (nl == lcap) {
            lcap *= 2;
            out->node_label_clusters = (int32_t *)realloc(out->node_label_clusters, sizeof(int32_t) * lcap);
            out->node_label_types = (int32_t *)realloc(out->node_label_types, sizeof(int32_t) * lcap);
          }
          out->node_label_clusters[nl] = r;
          out->node_label_types[nl] = k;
          nl++;
          cnt++;
          if (cidx == -1) cidx = r;
        }
      }
    }
    out->node_labels[i] = mask;
    out->node_cluster_indices[i] = cidx;
    out->node_label_counts[i] = cnt;
  }
  if (fast) orc_sgrid_free(&lgrid);
  out->n_label_entries = nl;
  out->n_edges = nf;
  out->edges = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(nf + 1));
  out->edge_lengths = (float *)malloc(sizeof(float) * (size_t)(nf + 1));
  out->edge_clearances = (float *)malloc(sizeof(float) * (size_t)(nf + 1));
  for (int k = 0; k < nf; ++k) {
    out->edges[2 * k] = frec[k].from;
    out->edges[2 * k + 1] = frec[k].to;
    out->edge_lengths[k] = (float)frec[k].length_m;
    out->edge_clearances[k] = frec[k].clearance;
  }
  free(frec); free(nodes);
  return 0;
}

void orc_graph_free(orc_graph *g) {
  free(g->nodes); free(g->node_labels); free(g->node_cluster_indices); free(g->node_label_counts);
  free(g->node_label_clusters); free(g->node_label_types); free(g->edges); free(g->edge_lengths);
  free(g->edge_clearances); free(g->corner_points);
  memset(g, 0, sizeof(*g));
}


/* trimPathNearOccupiedRegions, src/aos_path_gen_node.cpp:1570-1630: the first pose
 * i > 0 with a skeleton cell == 100 inside the 0.2 m stencil around it truncates the path to i poses; pose 0 is
 * tested but never trims (":1620  if (too_close && i > 0)").  Returns the new number of poses.
 * grid: the published (framed) skeleton, int8 row-major; resolution is the message's float widened to double. */
int orc_trim_path(const double *path_xy, int n, const int8_t *grid, int width, int height, double origin_x,
                  double origin_y, float res, double safety_distance) {
  const double resolution = (double)res;
  if (!grid || n <= 0) return n;
  for (int i = 0; i < n; ++i) {
    const double px = path_xy[2 * i], py = path_xy[2 * i + 1];
    int too_close = 0;
    const int check_radius_cells = (int)ceil(safety_distance / resolution);
    for (int dx = -check_radius_cells; dx <= check_radius_cells && !too_close; ++dx) {
      for (int dy = -check_radius_cells; dy <= check_radius_cells && !too_close; ++dy) {
        const double check_x = px + dx * resolution;
        const double check_y = py + dy * resolution;
        const double dist = sqrt((double)(dx * dx + dy * dy)) * resolution;
        if (dist > safety_distance) continue;
        const int mx = (int)((check_x - origin_x) / resolution);
        const int my = (int)((check_y - origin_y) / resolution);
        if (mx >= 0 && mx < width && my >= 0 && my < height) {
          const int index = mx + my * width;
          if (index >= 0 && (size_t)index < (size_t)width * (size_t)height) {
            if (grid[index] == 100) {
              too_close = 1;
              break;
            }
          }
        }
      }
    }
    if (too_close && i > 0) return i;
  }
  return n;
}
