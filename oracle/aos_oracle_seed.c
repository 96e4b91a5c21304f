/*
 * aos_oracle_seed.c -- CPU restatement of aos_seed_gen_node's processPointCloud path.
 * TEST INFRASTRUCTURE ONLY (see aos_oracle.h).  Compile with -O2 -ffp-contract=off: the reference
 * is x86-64 without FMA contraction, and the float32/float64 mix below is kept literally.
 * Citations: /root/reference/src/aos_seed_gen_node.cpp (abbreviated "sg").
 */
#include "aos_oracle.h"
#include "aos_oracle_fast.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OCC 100

/* ---- getActiveBounds, sg:874-890 --------------------------------------------------------- */
void orc_active_bounds(const orc_seed_params *p, float *minx, float *maxx, float *miny, float *maxy) {
  if (p->n_poly > 0) {
    /* bbox as computed at sg:202-212 */
    double bx0 = p->poly[0], bx1 = p->poly[0], by0 = p->poly[1], by1 = p->poly[1];
    for (int i = 0; i < p->n_poly; ++i) {
      double x = p->poly[2 * i], y = p->poly[2 * i + 1];
      if (x < bx0) bx0 = x;
      if (x > bx1) bx1 = x;
      if (y < by0) by0 = y;
      if (y > by1) by1 = y;
    }
    const double margin = 2.5;
    *minx = (float)(bx0 - margin);
    *maxx = (float)(bx1 + margin);
    *miny = (float)(by0 - margin);
    *maxy = (float)(by1 + margin);
  } else {
    *minx = p->clipping_minx;
    *maxx = p->clipping_maxx;
    *miny = p->clipping_miny;
    *maxy = p->clipping_maxy;
  }
}

/* ---- generateOccupancyGrid dims, sg:587-596 ---------------------------------------------- */
void orc_grid_dims(float minx, float maxx, float miny, float maxy, float res, int *w, int *h) {
  float width = fmaxf(0.0f, maxx - minx);
  float height = fmaxf(0.0f, maxy - miny);
  unsigned int wc = (unsigned int)ceilf(width / res);
  unsigned int hc = (unsigned int)ceilf(height / res);
  if (wc == 0) wc = 1;
  if (hc == 0) hc = 1;
  *w = (int)wc;
  *h = (int)hc;
}

/* ---- processPointCloud steps 1-3: PassThrough z,x,y (sg:459-477), exclusion discs (sg:481-525),
 *      z:=0 (sg:528-538), generateOccupancyGrid scatter (sg:607-619) -------------------------- */
void orc_bin_points(const orc_seed_params *p, const float *points, size_t n, size_t stride_floats,
                    int w, int h, double ox, double oy, int8_t *grid) {
  float minx, maxx, miny, maxy;
  orc_active_bounds(p, &minx, &maxx, &miny, &maxy);
  const float res = p->grid_resolution;
  memset(grid, 0, (size_t)w * (size_t)h);
  for (size_t i = 0; i < n; ++i) {
    const float *pt = points + i * stride_floats;
    float x = pt[0], y = pt[1], z = pt[2];
    /* pcl::PassThrough: non-finite points removed, limits inclusive, float compare */
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (z < p->clipping_minz || z > p->clipping_maxz) continue;
    if (x < minx || x > maxx) continue;
    if (y < miny || y > maxy) continue;
    int exclude = 0;
    for (int e = 0; e < p->n_excl; ++e) {
      float dx = x - p->excl[3 * e];
      float dy = y - p->excl[3 * e + 1];
      float dist_sq = dx * dx + dy * dy;
      float r = p->excl[3 * e + 2];
      if (dist_sq <= r * r) {
        exclude = 1;
        break;
      }
    }
    if (exclude) continue;
    /* sg:609-610: float - double -> double; / float -> double; truncation */
    int gx = (int)((x - ox) / res);
    int gy = (int)((y - oy) / res);
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) grid[(size_t)gx + (size_t)gy * (size_t)w] = OCC;
  }
}

/* ---- applyInflation, sg:933-967 ---------------------------------------------------------- */
void orc_inflate(const int8_t *in, int w, int h, int cells, int8_t *out) {
  memcpy(out, in, (size_t)w * (size_t)h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (in[(size_t)x + (size_t)y * w] != OCC) continue;
      for (int dy = -cells; dy <= cells; ++dy)
        for (int dx = -cells; dx <= cells; ++dx) {
          if (dx * dx + dy * dy > cells * cells) continue;
          int nx = x + dx, ny = y + dy;
          if (nx >= 0 && nx < w && ny >= 0 && ny < h) out[(size_t)nx + (size_t)ny * w] = OCC;
        }
    }
}

/* ---- markBoundariesAsOccupied, sg:708-757 ------------------------------------------------ */
void orc_mark_borders(const int8_t *in, int w, int h, int8_t *out) {
  const int t = 5;
  memcpy(out, in, (size_t)w * (size_t)h);
  for (int y = 0; y < t && y < h; ++y)
    for (int x = 0; x < w; ++x) out[(size_t)x + (size_t)y * w] = OCC;
  for (int y = (h - t > 0 ? h - t : 0); y < h; ++y)
    for (int x = 0; x < w; ++x) out[(size_t)x + (size_t)y * w] = OCC;
  for (int x = 0; x < t && x < w; ++x)
    for (int y = 0; y < h; ++y) out[(size_t)x + (size_t)y * w] = OCC;
  for (int x = (w - t > 0 ? w - t : 0); x < w; ++x)
    for (int y = 0; y < h; ++y) out[(size_t)x + (size_t)y * w] = OCC;
}

/* ---- cv::morphologyEx(MORPH_OPEN, getStructuringElement(MORPH_ELLIPSE,3x3)), sg:678-680.
 *      The 3x3 ellipse is the 4-connected cross.  Default border: out-of-image pixels never
 *      constrain the erosion and never add in the dilation.  Pinned against cv2 in tests. ------ */
void orc_open_cross(const int8_t *in, int w, int h, int8_t *out) {
  size_t n = (size_t)w * (size_t)h;
  int8_t *er = (int8_t *)malloc(n);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      size_t i = (size_t)x + (size_t)y * w;
      int v = in[i] == OCC;
      if (v && x > 0) v = in[i - 1] == OCC;
      if (v && x < w - 1) v = in[i + 1] == OCC;
      if (v && y > 0) v = in[i - w] == OCC;
      if (v && y < h - 1) v = in[i + w] == OCC;
      er[i] = (int8_t)v;
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      size_t i = (size_t)x + (size_t)y * w;
      int v = er[i];
      if (!v && x > 0) v = er[i - 1];
      if (!v && x < w - 1) v = er[i + 1];
      if (!v && y > 0) v = er[i - w];
      if (!v && y < h - 1) v = er[i + w];
      out[i] = v ? OCC : 0;
    }
  free(er);
}

/* ---- cv::ximgproc::thinning(THINNING_ZHANGSUEN), call site sg:684 (opencv_contrib 4.5.4:
 *      modules/ximgproc/src/thinning.cpp thinningIteration/thinning).  Restated, un-pinned. ---- */
static int thin_subiter(uint8_t *img, uint8_t *marker, int w, int h, int iter) {
  int changed = 0;
  memset(marker, 0, (size_t)w * (size_t)h);
  for (int i = 1; i < h - 1; ++i)
    for (int j = 1; j < w - 1; ++j) {
      const uint8_t *c = img + (size_t)i * w + j;
      if (!*c) continue; /* deleting a 0 pixel is a no-op (img &= ~marker) */
      int p2 = c[-w], p3 = c[-w + 1], p4 = c[1], p5 = c[w + 1];
      int p6 = c[w], p7 = c[w - 1], p8 = c[-1], p9 = c[-w - 1];
      int A = (p2 == 0 && p3 == 1) + (p3 == 0 && p4 == 1) + (p4 == 0 && p5 == 1) +
              (p5 == 0 && p6 == 1) + (p6 == 0 && p7 == 1) + (p7 == 0 && p8 == 1) +
              (p8 == 0 && p9 == 1) + (p9 == 0 && p2 == 1);
      int B = p2 + p3 + p4 + p5 + p6 + p7 + p8 + p9;
      int m1 = iter == 0 ? (p2 * p4 * p6) : (p2 * p4 * p8);
      int m2 = iter == 0 ? (p4 * p6 * p8) : (p2 * p6 * p8);
      if (A == 1 && (B >= 2 && B <= 6) && m1 == 0 && m2 == 0) marker[(size_t)i * w + j] = 1;
    }
  size_t n = (size_t)w * (size_t)h;
  for (size_t k = 0; k < n; ++k)
    if (marker[k]) {
      img[k] = 0;
      changed = 1;
    }
  return changed;
}

int orc_thin_zhangsuen(int8_t *grid, int w, int h) {
  size_t n = (size_t)w * (size_t)h;
  uint8_t *img = (uint8_t *)malloc(n), *marker = (uint8_t *)malloc(n);
  for (size_t k = 0; k < n; ++k) img[k] = grid[k] == OCC; /* occupancyGridToMat + /=255 */
  int passes = 0;
  for (;;) {
    int c0 = thin_subiter(img, marker, w, h, 0);
    int c1 = thin_subiter(img, marker, w, h, 1);
    ++passes;
    if (!c0 && !c1) break; /* absdiff(processed, prev) has no non-zero */
  }
  for (size_t k = 0; k < n; ++k) grid[k] = img[k] ? OCC : 0; /* *=255, matToOccupancyGrid */
  free(img);
  free(marker);
  return passes;
}

/* ---- worldToGrid, sg:760-769 ------------------------------------------------------------- */
static void world_to_grid(double ox, double oy, float res, int w, int h, float wx, float wy, int *gx,
                          int *gy) {
  float rel_x = (float)((wx - ox) / res);
  float rel_y = (float)((wy - oy) / res);
  *gx = (int)floorf(rel_x);
  *gy = (int)floorf(rel_y);
  if (*gx < 0) *gx = 0; else if (*gx >= w) *gx = w - 1;
  if (*gy < 0) *gy = 0; else if (*gy >= h) *gy = h - 1;
}

/* ---- drawLineInGrid (Bresenham), sg:828-870 ---------------------------------------------- */
static void draw_line(int8_t *g, int w, int h, int x0, int y0, int x1, int y1) {
  x0 = x0 < 0 ? 0 : (x0 > w - 1 ? w - 1 : x0);
  y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0);
  x1 = x1 < 0 ? 0 : (x1 > w - 1 ? w - 1 : x1);
  y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
  int dx = abs(x1 - x0), dy = abs(y1 - y0);
  int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
  int err = dx - dy, x = x0, y = y0;
  for (;;) {
    g[(size_t)x + (size_t)y * w] = OCC;
    if (x == x1 && y == y1) break;
    int e2 = 2 * err;
    if (e2 > -dy) { err -= dy; x += sx; }
    if (e2 < dx) { err += dx; y += sy; }
  }
}

/* ---- markPolygonBoundaryAsOccupied, sg:772-825 (polygon non-empty branch; else grid frame) - */
void orc_frame_polygon_bbox(const orc_seed_params *p, const int8_t *in, int w, int h, double ox,
                            double oy, int8_t *out) {
  if (p->n_poly == 0) { /* sg:799-801 */
    orc_mark_borders(in, w, h, out);
    return;
  }
  memcpy(out, in, (size_t)w * (size_t)h);
  double bx0 = p->poly[0], bx1 = p->poly[0], by0 = p->poly[1], by1 = p->poly[1];
  for (int i = 0; i < p->n_poly; ++i) {
    double x = p->poly[2 * i], y = p->poly[2 * i + 1];
    if (x < bx0) bx0 = x;
    if (x > bx1) bx1 = x;
    if (y < by0) by0 = y;
    if (y > by1) by1 = y;
  }
  const double margin = 2.5;
  int gx0, gy0, gx1, gy1;
  world_to_grid(ox, oy, p->grid_resolution, w, h, (float)(bx0 - margin), (float)(by0 - margin), &gx0, &gy0);
  world_to_grid(ox, oy, p->grid_resolution, w, h, (float)(bx1 + margin), (float)(by1 + margin), &gx1, &gy1);
  draw_line(out, w, h, gx0, gy0, gx1, gy0);
  draw_line(out, w, h, gx0, gy1, gx1, gy1);
  draw_line(out, w, h, gx0, gy0, gx0, gy1);
  draw_line(out, w, h, gx1, gy0, gx1, gy1);
}

/* ---- isPointInPolygon, sg:1231-1255 ------------------------------------------------------ */
int orc_point_in_polygon(double px, double py, const double *poly, int n) {
  if (n < 3) return 0;
  int inside = 0;
  int j = n - 1;
  for (int i = 0; i < n; ++i) {
    double pix = poly[2 * i], piy = poly[2 * i + 1];
    double pjx = poly[2 * j], pjy = poly[2 * j + 1];
    double dy = pjy - piy;
    if (fabs(dy) > 1e-9) {
      if (((piy > py) != (pjy > py)) && (px < (pjx - pix) * (py - piy) / dy + pix)) inside = !inside;
    }
    j = i;
  }
  return inside;
}

/* ---- clusterOccupiedCells, sg:970-1083 ---------------------------------------------------- */
typedef struct {
  int n, cap;
  int32_t *first, *size;
  int64_t *sumx, *sumy, *maxd2;
  float *cx, *cy, *len;
  int32_t *off;   /* n+1 */
  int32_t *cells; /* BFS order */
  int ncells, cellcap;
} cluster_set;

static void cs_push_cell(cluster_set *cs, int32_t idx) {
  if (cs->ncells == cs->cellcap) {
    cs->cellcap = cs->cellcap ? cs->cellcap * 2 : 1024;
    cs->cells = (int32_t *)realloc(cs->cells, sizeof(int32_t) * (size_t)cs->cellcap);
  }
  cs->cells[cs->ncells++] = idx;
}

static void cs_grow(cluster_set *cs) {
  if (cs->n + 1 >= cs->cap) {
    cs->cap = cs->cap ? cs->cap * 2 : 64;
    cs->first = (int32_t *)realloc(cs->first, sizeof(int32_t) * cs->cap);
    cs->size = (int32_t *)realloc(cs->size, sizeof(int32_t) * cs->cap);
    cs->sumx = (int64_t *)realloc(cs->sumx, sizeof(int64_t) * cs->cap);
    cs->sumy = (int64_t *)realloc(cs->sumy, sizeof(int64_t) * cs->cap);
    cs->maxd2 = (int64_t *)realloc(cs->maxd2, sizeof(int64_t) * cs->cap);
    cs->cx = (float *)realloc(cs->cx, sizeof(float) * cs->cap);
    cs->cy = (float *)realloc(cs->cy, sizeof(float) * cs->cap);
    cs->len = (float *)realloc(cs->len, sizeof(float) * cs->cap);
    cs->off = (int32_t *)realloc(cs->off, sizeof(int32_t) * (cs->cap + 1));
  }
}

static void cluster_cells(const orc_seed_params *p, const int8_t *grid, int w, int h, double ox,
                          double oy, cluster_set *cs, int32_t *labels) {
  const float res = p->grid_resolution;
  const int use_poly = p->n_poly > 0;
  size_t n = (size_t)w * (size_t)h;
  uint8_t *visited = (uint8_t *)calloc(n, 1);
  if (labels)
    for (size_t k = 0; k < n; ++k) labels[k] = -1;
  static const int ddx[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  static const int ddy[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
  memset(cs, 0, sizeof(*cs));
  cs_grow(cs);
  cs->off[0] = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      size_t index = (size_t)x + (size_t)y * w;
      if (grid[index] != OCC || visited[index]) continue;
      if (use_poly) {
        /* sg:998-999: double + float(float*float) -> double -> float */
        float wx = (float)(ox + (float)x * res);
        float wy = (float)(oy + (float)y * res);
        if (!orc_point_in_polygon(wx, wy, p->poly, p->n_poly)) {
          visited[index] = 1;
          continue;
        }
      }
      cs_grow(cs);
      int c = cs->n;
      int qhead = cs->ncells; /* the BFS queue is the cells array itself (FIFO) */
      cs_push_cell(cs, (int32_t)index);
      visited[index] = 1;
      while (qhead < cs->ncells) {
        int32_t cur = cs->cells[qhead++];
        int cx = cur % w, cy = cur / w;
        for (int i = 0; i < 8; ++i) {
          int nx = cx + ddx[i], ny = cy + ddy[i];
          if (nx < 0 || nx >= w || ny < 0 || ny >= h) continue;
          size_t ni = (size_t)nx + (size_t)ny * w;
          if (visited[ni] || grid[ni] != OCC) continue;
          if (use_poly) {
            float nwx = (float)(ox + (float)nx * res);
            float nwy = (float)(oy + (float)ny * res);
            if (!orc_point_in_polygon(nwx, nwy, p->poly, p->n_poly)) {
              visited[ni] = 1;
              continue;
            }
          }
          visited[ni] = 1;
          cs_push_cell(cs, (int32_t)ni);
        }
      }
      int beg = cs->off[c], end = cs->ncells;
      /* sg:1053-1060: float32 running sums in BFS order */
      float sum_x = 0.0f, sum_y = 0.0f;
      int64_t isx = 0, isy = 0;
      for (int k = beg; k < end; ++k) {
        int gx = cs->cells[k] % w, gy = cs->cells[k] / w;
        sum_x += gx;
        sum_y += gy;
        isx += gx;
        isy += gy;
        if (labels) labels[cs->cells[k]] = (int32_t)index;
      }
      size_t cnt = (size_t)(end - beg);
      cs->first[c] = (int32_t)index;
      cs->size[c] = (int32_t)cnt;
      cs->sumx[c] = isx;
      cs->sumy[c] = isy;
      cs->cx[c] = sum_x / cnt; /* float / size_t -> float division */
      cs->cy[c] = sum_y / cnt;
      /* sg:1062-1074: max over pairs of float(sqrt(int d2) * res); monotone in d2 */
      int64_t maxd2 = 0;
      float max_distance = 0.0f;
      if (orc_fast_enabled()) maxd2 = orc_fast_max_pair_d2(cs->cells + beg, end - beg, w); /* same maximum, via the hull */
      else for (int a = beg; a < end; ++a) {
        int ax = cs->cells[a] % w, ay = cs->cells[a] / w;
        for (int b = a + 1; b < end; ++b) {
          int dx = ax - cs->cells[b] % w, dy = ay - cs->cells[b] / w;
          int d2 = dx * dx + dy * dy;
          if (d2 > maxd2) maxd2 = d2;
        }
      }
      if (maxd2 > 0) {
        float distance = (float)(sqrt((double)maxd2) * res);
        if (distance > max_distance) max_distance = distance;
      }
      cs->maxd2[c] = maxd2;
      cs->len[c] = max_distance;
      cs->n++;
      cs->off[cs->n] = cs->ncells;
    }
  free(visited);
}

/* ---- tiny growable point list --------------------------------------------------------------- */
typedef struct {
  double *xy;
  int n, cap;
  orc_sgrid *grid; /* fast mode: hash grid over the points pushed so far (cell slightly above the largest radius asked) */
} ptlist;
static void pl_push(ptlist *l, double x, double y) {
  if (orc_fast_enabled()) {
    if (!l->grid) {
      l->grid = (orc_sgrid *)malloc(sizeof(orc_sgrid));
      orc_sgrid_init(l->grid, 0.5005, 1024);
    }
    orc_sgrid_add(l->grid, x, y);
  }
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 256;
    l->xy = (double *)realloc(l->xy, sizeof(double) * 2 * (size_t)l->cap);
  }
  l->xy[2 * l->n] = x;
  l->xy[2 * l->n + 1] = y;
  l->n++;
}
/* first-come duplicate test used all over sg (dist < 0.5) */
static int pl_has_within(const ptlist *l, double x, double y, double r) {
  if (l->grid && r <= 0.5) { /* existence test: the order of the candidates does not matter */
    int32_t cand[256];
    int m = orc_sgrid_gather(l->grid, x, y, 1, cand, 256);
    if (m <= 256) {
      for (int k = 0; k < m; ++k) {
        double ddx = l->xy[2 * cand[k]] - x, ddy = l->xy[2 * cand[k] + 1] - y;
        if (sqrt(ddx * ddx + ddy * ddy) < r) return 1;
      }
      return 0;
    }
  }
  for (int i = 0; i < l->n; ++i) {
    double ddx = l->xy[2 * i] - x, ddy = l->xy[2 * i + 1] - y;
    double dist = sqrt(ddx * ddx + ddy * ddy); /* std::pow(.,2) == exact square */
    if (dist < r) return 1;
  }
  return 0;
}

typedef struct {
  const int8_t *data;
  int w, h;
  double ox, oy;
  float res;
} gridview;

/* ---- raycastToOccupiedCell, sg:1730-1771 (uses node member grid_resolution + worldToGrid) -- */
static int raycast_to_occupied(const gridview *g, double sx, double sy, double dx, double dy,
                               double max_distance, double *hx, double *hy) {
  const double step_size = g->res * 0.5;
  const int max_steps = (int)(max_distance / step_size);
  const double min_distance = 1.0;
  double cx = sx, cy = sy;
  for (int i = 0; i < max_steps; ++i) {
    cx += dx * step_size;
    cy += dy * step_size;
    double ex = cx - sx, ey = cy - sy;
    double distance = sqrt(ex * ex + ey * ey);
    if (distance < min_distance) continue;
    int gx, gy;
    world_to_grid(g->ox, g->oy, g->res, g->w, g->h, (float)cx, (float)cy, &gx, &gy);
    if (g->data[(size_t)gx + (size_t)gy * g->w] == OCC) {
      *hx = cx;
      *hy = cy;
      return 1;
    }
  }
  return 0;
}

static void grid_world_bounds(const gridview *g, double *minx, double *maxx, double *miny, double *maxy) {
  /* info.width * info.resolution is uint32 * float -> float (sg:1808,1810) */
  *minx = g->ox;
  *maxx = g->ox + (float)((float)(unsigned)g->w * g->res);
  *miny = g->oy;
  *maxy = g->oy + (float)((float)(unsigned)g->h * g->res);
}

/* ---- castRayFromEndpoint, sg:1774-1891 ---------------------------------------------------- */
static void cast_ray_from_endpoint(const gridview *g, double spx, double spy, double opx, double opy,
                                   double angle_offset_deg, double min_distance, double *rx, double *ry) {
  double ex = opx - spx, ey = opy - spy;
  double dist_to_other = sqrt(ex * ex + ey * ey);
  if (dist_to_other < 1e-6) {
    ex = 1.0;
    ey = 0.0;
  } else { /* Eigen normalize(): z = squaredNorm; if (z > 0) v /= sqrt(z) */
    double z = ex * ex + ey * ey;
    if (z > 0) {
      double s = sqrt(z);
      ex /= s;
      ey /= s;
    }
  }
  double outx = -ex, outy = -ey;
  double perpx = -ey, perpy = ex;
  double a = angle_offset_deg * M_PI / 180.0;
  double rdx, rdy;
  if (angle_offset_deg > 0) {
    rdx = cos(a) * outx + sin(a) * perpx;
    rdy = cos(a) * outy + sin(a) * perpy;
  } else {
    rdx = cos(-a) * outx + sin(-a) * (-perpx);
    rdy = cos(-a) * outy + sin(-a) * (-perpy);
  }
  {
    double z = rdx * rdx + rdy * rdy;
    if (z > 0) {
      double s = sqrt(z);
      rdx /= s;
      rdy /= s;
    }
  }
  double minx, maxx, miny, maxy;
  grid_world_bounds(g, &minx, &maxx, &miny, &maxy);
  const double resolution = g->res;
  const double step_size = 0.1;
  double current = min_distance;
  double gw = (float)((float)(unsigned)g->w * g->res), gh = (float)((float)(unsigned)g->h * g->res);
  double abs_max = sqrt(gw * gw + gh * gh) * 3.0;
  while (current <= abs_max) {
    double px = spx + rdx * current, py = spy + rdy * current;
    if (!(px >= minx && px <= maxx && py >= miny && py <= maxy)) {
      *rx = fmax(minx, fmin(maxx, px));
      *ry = fmax(miny, fmin(maxy, py));
      return;
    }
    int mx = (int)((px - g->ox) / resolution);
    int my = (int)((py - g->oy) / resolution);
    if (mx >= 0 && mx < g->w && my >= 0 && my < g->h && g->data[(size_t)mx + (size_t)my * g->w] == OCC) {
      *rx = px;
      *ry = py;
      return;
    }
    current += step_size;
  }
  double fx = spx + rdx * abs_max, fy = spy + rdy * abs_max;
  if (!(fx >= minx && fx <= maxx && fy >= miny && fy <= maxy)) {
    fx = fmax(minx, fmin(maxx, fx));
    fy = fmax(miny, fmin(maxy, fy));
  }
  *rx = fx;
  *ry = fy;
}

/* stable insertion sort with the reference's comparator (sg:2554-2560) */
static int row_less(const double *a, const double *b) { /* a,b -> 7-double rows: cx,cy,... */
  if (fabs(a[1] - b[1]) < 1e-6) return a[0] < b[0];
  return a[1] < b[1];
}

int orc_seed_stage(const orc_seed_params *p, const float *points, size_t n_points,
                   size_t stride_floats, orc_seed_result *out) {
  memset(out, 0, sizeof(*out));
  float minx, maxx, miny, maxy;
  orc_active_bounds(p, &minx, &maxx, &miny, &maxy);
  int w, h;
  orc_grid_dims(minx, maxx, miny, maxy, p->grid_resolution, &w, &h);
  size_t n = (size_t)w * (size_t)h;
  out->w = w;
  out->h = h;
  out->origin_x = minx; /* sg:597-598 float -> double */
  out->origin_y = miny;
  out->res = p->grid_resolution;
  out->occ_raw = (int8_t *)malloc(n);
  out->occ_inflated = (int8_t *)malloc(n);
  out->occ_border = (int8_t *)malloc(n);
  out->opened = (int8_t *)malloc(n);
  out->skel = (int8_t *)malloc(n);
  out->skel_framed = (int8_t *)malloc(n);
  out->labels = orc_fast_skip_labels() ? NULL : (int32_t *)malloc(sizeof(int32_t) * n);

  const int fast = orc_fast_enabled(); /* indexed / threaded variants with identical results (aos_oracle_fast.h) */
  int inflation_cells = (int)(p->inflation_radius / p->grid_resolution); /* sg:936 float division */
  if (fast) {
    orc_fast_bin_points(p, points, n_points, stride_floats, w, h, out->origin_x, out->origin_y, out->occ_raw);
    orc_fast_inflate(out->occ_raw, w, h, inflation_cells, out->occ_inflated);
  } else {
    orc_bin_points(p, points, n_points, stride_floats, w, h, out->origin_x, out->origin_y, out->occ_raw);
    orc_inflate(out->occ_raw, w, h, inflation_cells, out->occ_inflated);
  }
  orc_mark_borders(out->occ_inflated, w, h, out->occ_border);
  if (fast) orc_fast_open_cross(out->occ_inflated, w, h, out->opened);
  else orc_open_cross(out->occ_inflated, w, h, out->opened); /* skeleton of the UN-bordered grid, sg:560 */
  memcpy(out->skel, out->opened, n);
  if (fast) orc_fast_thin_zhangsuen(out->skel, w, h);
  else orc_thin_zhangsuen(out->skel, w, h);

  cluster_set cs;
  cluster_cells(p, out->skel, w, h, out->origin_x, out->origin_y, &cs, out->labels);
  out->n_clusters = cs.n;
  out->cl_first = cs.first;
  out->cl_size = cs.size;
  out->cl_sumx = cs.sumx;
  out->cl_sumy = cs.sumy;
  out->cl_cx = cs.cx;
  out->cl_cy = cs.cy;
  out->cl_maxd2 = cs.maxd2;
  out->cl_len = cs.len;
  out->cl_cell_off = cs.off;
  out->cl_cells = cs.cells;

  /* clusterAndVisualizeSkeletonizedGrid length filter (sg:1262-1270) then
   * convertClustersToTreeRows (sg:1309-1406) */
  const float res = p->grid_resolution;
  const float min_length = (float)p->cluster_min_length;
  const int use_poly = p->n_poly > 0;
  out->row_cluster = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cs.n + 1));
  out->rows = (double *)malloc(sizeof(double) * 7 * (size_t)(cs.n + 1));
  int nr = 0;
  for (int c = 0; c < cs.n; ++c) {
    if (!(cs.len[c] >= min_length)) continue;
    float center_x = (float)(out->origin_x + cs.cx[c] * res); /* sg:1333 */
    float center_y = (float)(out->origin_y + cs.cy[c] * res);
    if (use_poly && !orc_point_in_polygon(center_x, center_y, p->poly, p->n_poly)) continue;
    double rcx = center_x, rcy = center_y;
    int beg = cs.off[c], end = cs.off[c + 1], cnt = end - beg;
    double *wp = (double *)malloc(sizeof(double) * 2 * (size_t)cnt);
    for (int k = 0; k < cnt; ++k) {
      int gx = cs.cells[beg + k] % w, gy = cs.cells[beg + k] / w;
      float wx = (float)(out->origin_x + gx * res); /* sg:1349: int*float -> float */
      float wy = (float)(out->origin_y + gy * res);
      wp[2 * k] = wx;
      wp[2 * k + 1] = wy;
    }
    double max_dist_sq = 0.0;
    int first_idx = 0;
    double fdx = 0.0, fdy = 0.0; /* first_direction (uninitialised in the reference if no point differs) */
    for (int k = 0; k < cnt; ++k) {
      double dx = wp[2 * k] - rcx, dy = wp[2 * k + 1] - rcy;
      double d2 = dx * dx + dy * dy;
      if (d2 > max_dist_sq) {
        max_dist_sq = d2;
        first_idx = k;
        double s = sqrt(d2); /* normalized(): d2 > 0 here */
        fdx = dx / s;
        fdy = dy / s;
      }
    }
    double max_opp = 0.0;
    int second_idx = 0;
    for (int k = 0; k < cnt; ++k) {
      if (k == first_idx) continue;
      double dx = wp[2 * k] - rcx, dy = wp[2 * k + 1] - rcy;
      double d2 = dx * dx + dy * dy;
      double nx = dx, ny = dy;
      if (d2 > 0) {
        double s = sqrt(d2);
        nx = dx / s;
        ny = dy / s;
      }
      double dot = nx * fdx + ny * fdy;
      if (dot < 0.0 && d2 > max_opp) {
        max_opp = d2;
        second_idx = k;
      }
    }
    if (max_opp == 0.0) {
      for (int k = 0; k < cnt; ++k) {
        if (k == first_idx) continue;
        double dx = wp[2 * k] - wp[2 * first_idx], dy = wp[2 * k + 1] - wp[2 * first_idx + 1];
        double d2 = dx * dx + dy * dy;
        if (d2 > max_opp) {
          max_opp = d2;
          second_idx = k;
        }
      }
    }
    double *r = out->rows + 7 * nr;
    r[0] = rcx;
    r[1] = rcy;
    r[2] = wp[2 * first_idx];
    r[3] = wp[2 * first_idx + 1];
    r[4] = wp[2 * second_idx];
    r[5] = wp[2 * second_idx + 1];
    r[6] = cs.len[c];
    out->row_cluster[nr] = c;
    nr++;
    free(wp);
  }
  out->n_rows = nr;

  /* rows_info: stable sort copy by (cy, cx) with the 1e-6 rule */
  out->rows_info = (double *)malloc(sizeof(double) * 4 * (size_t)(nr + 1));
  {
    int *ord = (int *)malloc(sizeof(int) * (size_t)(nr + 1));
    for (int i = 0; i < nr; ++i) {
      int j = i;
      while (j > 0 && row_less(out->rows + 7 * i, out->rows + 7 * ord[j - 1])) {
        ord[j] = ord[j - 1];
        --j;
      }
      ord[j] = i;
    }
    for (int i = 0; i < nr; ++i) memcpy(out->rows_info + 4 * i, out->rows + 7 * ord[i] + 2, sizeof(double) * 4);
    free(ord);
  }

  /* ---- seeds (host-side "seed selection"), on the UN-framed skeleton (sg:563-566,1436-1441) -- */
  gridview g = {out->skel, w, h, out->origin_x, out->origin_y, res};
  ptlist virt = {0}, ray = {0}, endp = {0};
  const double virtual_seed_interval = 1.0; /* sg:2666 */
  /* generateVirtualSeeds, sg:1987-2268 (real_seeds_ is always empty: confirmed trees were removed) */
  for (int ri = 0; ri < nr; ++ri) {
    const double *r = out->rows + 7 * ri;
    if (use_poly && !orc_point_in_polygon(r[0], r[1], p->poly, p->n_poly)) continue;
    double dx = r[4] - r[2], dy = r[5] - r[3];
    double distance = sqrt(dx * dx + dy * dy);
    if (distance < virtual_seed_interval) continue;
    double nrm = sqrt(dx * dx + dy * dy);
    if (nrm < 1e-6) continue;
    double rdx = dx / nrm, rdy = dy / nrm;
    double p1x = -rdy, p1y = rdx, p2x = rdy, p2y = -rdx;
    int num_seeds = (int)floor(distance / virtual_seed_interval);
    for (int i = 1; i <= num_seeds; ++i) {
      double t = (double)i / (num_seeds + 1);
      double bx = r[2] + t * dx, by = r[3] + t * dy;
      if (!pl_has_within(&virt, bx, by, 0.5)) pl_push(&virt, bx, by);
      const double max_raycast_distance = 4.0;
      for (int side = 0; side < 2; ++side) {
        double pdx = side == 0 ? p1x : p2x, pdy = side == 0 ? p1y : p2y;
        double hx, hy, sx, sy;
        if (raycast_to_occupied(&g, bx, by, pdx, pdy, max_raycast_distance, &hx, &hy)) {
          sx = hx;
          sy = hy;
        } else {
          sx = bx + pdx * max_raycast_distance;
          sy = by + pdy * max_raycast_distance;
        }
        if (use_poly && orc_point_in_polygon(sx, sy, p->poly, p->n_poly)) continue;
        if (!pl_has_within(&virt, sx, sy, 0.5)) pl_push(&virt, sx, sy);
      }
    }
  }
  /* generateRayPointsFromEndpoints, sg:1894-1982 (rows in all_tree_rows order) */
  {
    double gminx, gmaxx, gminy, gmaxy;
    grid_world_bounds(&g, &gminx, &gmaxx, &gminy, &gmaxy);
    for (int ri = 0; ri < nr; ++ri) {
      const double *r = out->rows + 7 * ri;
      double e1x = r[2], e1y = r[3], e2x = r[4], e2y = r[5];
      double rp[12];
      cast_ray_from_endpoint(&g, e1x, e1y, e2x, e2y, 0.0, 1.0, &rp[0], &rp[1]);
      cast_ray_from_endpoint(&g, e1x, e1y, e2x, e2y, -90.0, 1.0, &rp[2], &rp[3]);
      cast_ray_from_endpoint(&g, e1x, e1y, e2x, e2y, 90.0, 1.0, &rp[4], &rp[5]);
      cast_ray_from_endpoint(&g, e2x, e2y, e1x, e1y, 0.0, 1.0, &rp[6], &rp[7]);
      cast_ray_from_endpoint(&g, e2x, e2y, e1x, e1y, -90.0, 1.0, &rp[8], &rp[9]);
      cast_ray_from_endpoint(&g, e2x, e2y, e1x, e1y, 90.0, 1.0, &rp[10], &rp[11]);
      for (int k = 0; k < 6; ++k) {
        double x = rp[2 * k], y = rp[2 * k + 1];
        if (!isfinite(x) || !isfinite(y)) continue;
        if (!(x >= gminx && x <= gmaxx && y >= gminy && y <= gmaxy)) continue;
        if (use_poly && orc_point_in_polygon(x, y, p->poly, p->n_poly)) continue;
        if (!pl_has_within(&ray, x, y, 0.5)) pl_push(&ray, x, y);
      }
    }
  }
  /* endpoint seeds, sg:1450-1496 */
  for (int ri = 0; ri < nr; ++ri) {
    const double *r = out->rows + 7 * ri;
    if (!pl_has_within(&endp, r[2], r[3], 0.5)) pl_push(&endp, r[2], r[3]);
    if (!pl_has_within(&endp, r[4], r[5], 0.5)) pl_push(&endp, r[4], r[5]);
  }
  out->n_virtual = virt.n;
  out->n_ray = ray.n;
  out->n_endpoint = endp.n;
  out->n_seeds = virt.n + ray.n + endp.n;
  out->seeds = (double *)malloc(sizeof(double) * 2 * (size_t)(out->n_seeds + 1));
  memcpy(out->seeds, virt.xy, sizeof(double) * 2 * (size_t)virt.n);
  memcpy(out->seeds + 2 * virt.n, ray.xy, sizeof(double) * 2 * (size_t)ray.n);
  memcpy(out->seeds + 2 * (virt.n + ray.n), endp.xy, sizeof(double) * 2 * (size_t)endp.n);
  free(virt.xy);
  free(ray.xy);
  free(endp.xy);
  {
    ptlist *ls[3] = {&virt, &ray, &endp};
    for (int k = 0; k < 3; ++k)
      if (ls[k]->grid) { orc_sgrid_free(ls[k]->grid); free(ls[k]->grid); }
  }

  /* Step 9: frame AFTER clustering, sg:572 */
  orc_frame_polygon_bbox(p, out->skel, w, h, out->origin_x, out->origin_y, out->skel_framed);
  return 0;
}

void orc_seed_result_free(orc_seed_result *r) {
  free(r->occ_raw); free(r->occ_inflated); free(r->occ_border); free(r->opened);
  free(r->skel); free(r->skel_framed); free(r->labels);
  free(r->cl_first); free(r->cl_size); free(r->cl_sumx); free(r->cl_sumy);
  free(r->cl_cx); free(r->cl_cy); free(r->cl_maxd2); free(r->cl_len);
  free(r->cl_cell_off); free(r->cl_cells);
  free(r->row_cluster); free(r->rows); free(r->rows_info); free(r->seeds);
  memset(r, 0, sizeof(*r));
}
