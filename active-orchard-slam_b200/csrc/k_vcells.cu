// k_vcells.cu -- the OPT-IN device Voronoi (aos_set_voronoi_mode(ctx, AOS_VORONOI_DEVICE)): VoronoiDiagram::compute
// (src/utils/voronoi_diagram.cpp:16-114) without the sequential Subdiv2D insertion replay.
//
// What the reference takes from cv::Subdiv2D is, per seed, the polygon of its Voronoi cell in the diagram of all seeds
// plus Subdiv2D's three far outer vertices.  That diagram is a property of the point set, not of the insertion order, so
// it can be built in parallel: ONE THREAD PER SEED clips a polygon by the bisectors of the seed's neighbours, which it
// finds ring by ring in a uniform grid over the seeds, nearest first, until no unseen seed can cut the cell any more
// (a seed at distance d cuts only if d / 2 is smaller than the cell's current radius).  Cell vertices are kept
// combinatorially (the two neighbours whose bisectors meet there) and are emitted as the circumcentre of that site
// triple evaluated with OpenCV's computeVoronoiPoint formula (float32 differences and sums, double solve) on the triple
// in ascending site order, so the three cells that share a Voronoi vertex emit the same float32 bits.
//
// What this path does NOT reproduce, and why it is not the default: which of a triangle's three quad-edge pairs
// Subdiv2D computes the circumcentre from (float32 ulps), and at which vertex each facet starts -- both depend on
// Subdiv2D's flip history and decide, through the first-come 5 cm merge of extractBoundaryPoints, which of two nearby
// Voronoi vertices survives as a graph node.  The graph this path yields is the reference's graph up to those
// choices; tests/test_vcells_gpu.py measures the difference (nodes matched within 1e-4 m, edges under that matching).
// Cells are clipped to the seed rectangle grown by kClipMargin: everything outside the grid is cropped by
// filterNodesAndEdgesOutsideGrid (gvd:420-483) anyway, and bounded cells are what makes the neighbour search local.
#include <float.h>
#include <math.h>

#include <cmath>
#include <utility>

#include "aos_common.cuh"
#include "host_subdiv.h"

namespace aos {

namespace {
constexpr int kMaxPoly = 32;        // vertices of a cell while it is being clipped
constexpr int kMaxOut = 24;         // vertices of a finished cell (slot stride of the scratch polygons)
constexpr double kGridCell = 2.0;   // metres; merged seeds are at least 0.5 m apart
constexpr double kClipMargin = 3.0;   // beyond Subdiv2D's rectangle (= grid + 1 m): outside the grid either way
constexpr int kOuter = 3;           // sites 0..2 are Subdiv2D's outer triangle

struct VcParams {
  int n;                 // sites incl. the three outer ones
  float rx, ry, rw, rh;  // cv::Rect2f bounding_rect (vd:51-56)
  double cx0, cy0, cx1, cy1;  // clip rectangle
  double gx0, gy0;       // grid origin
  int gnx, gny;
};

// computeVoronoiPoint on the triple (a, b, c): bisectors of a->b and b->c
__device__ __forceinline__ float2 circumcentre_cv(double2 a, double2 b, double2 c) {
  const float ax = (float)a.x, ay = (float)a.y, bx = (float)b.x, by = (float)b.y, cx = (float)c.x, cy = (float)c.y;
  double a0 = __fsub_rn(bx, ax);
  double b0 = __fsub_rn(by, ay);
  double c0 = -0.5 * (a0 * (double)__fadd_rn(bx, ax) + b0 * (double)__fadd_rn(by, ay));
  double a1 = __fsub_rn(cx, bx);
  double b1 = __fsub_rn(cy, by);
  double c1 = -0.5 * (a1 * (double)__fadd_rn(cx, bx) + b1 * (double)__fadd_rn(cy, by));
  double det = a0 * b1 - a1 * b0;
  if (det == 0) return make_float2(FLT_MAX, FLT_MAX);
  det = 1. / det;
  return make_float2((float)((b0 * c1 - b1 * c0) * det), (float)((a1 * c0 - a0 * c1) * det));
}

// sites: the three outer vertices, then the seeds as VoronoiDiagram::compute inserts them (float32, clipped, vd:66-80)
__global__ void vc_sites_kernel(const double *__restrict__ seeds, VcParams P, float big, float irx, float iry,
                                double2 *__restrict__ sites, int *__restrict__ site_cell, uint32_t *__restrict__ cell_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  if (i < kOuter) {  // Subdiv2D::initDelaunay: (rx + big, ry), (rx, ry + big), (rx - big, ry - big) on the INT rect
    const float x = i == 0 ? __fadd_rn(irx, big) : i == 1 ? irx : __fsub_rn(irx, big);
    const float y = i == 0 ? iry : i == 1 ? __fadd_rn(iry, big) : __fsub_rn(iry, big);
    sites[i] = make_double2((double)x, (double)y);
    site_cell[i] = -1;
    return;
  }
  const double sx = seeds[2 * (i - kOuter)], sy = seeds[2 * (i - kOuter) + 1];
  if (!isfinite(sx) || !isfinite(sy)) {
    sites[i] = make_double2(NAN, NAN);
    site_cell[i] = -1;
    return;
  }
  const float margin = 0.1f;
  float x = (float)sx, y = (float)sy;
  x = fmaxf(__fadd_rn(P.rx, margin), fminf(__fsub_rn(__fadd_rn(P.rx, P.rw), margin), x));
  y = fmaxf(__fadd_rn(P.ry, margin), fminf(__fsub_rn(__fadd_rn(P.ry, P.rh), margin), y));
  sites[i] = make_double2((double)x, (double)y);
  int gx = (int)floor(((double)x - P.gx0) / kGridCell), gy = (int)floor(((double)y - P.gy0) / kGridCell);
  gx = min(max(gx, 0), P.gnx - 1);
  gy = min(max(gy, 0), P.gny - 1);
  const int cell = gy * P.gnx + gx;
  site_cell[i] = cell;
  atomicAdd(&cell_count[cell], 1u);
}

__global__ void vc_bin_kernel(const int *__restrict__ site_cell, int n, const uint32_t *__restrict__ cell_off,
                              uint32_t *__restrict__ cursor, int *__restrict__ items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || site_cell[i] < 0) return;
  const int cell = site_cell[i];
  items[cell_off[cell] + atomicAdd(&cursor[cell], 1u)] = i;
}

// ascending site index inside every grid cell: the clipping order, hence every tie, is the same on every run
__global__ void vc_sort_kernel(const uint32_t *__restrict__ cell_off, int ncell, int *__restrict__ items) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int b = (int)cell_off[c], e = (int)cell_off[c + 1];
  for (int i = b + 1; i < e; ++i) {
    const int v = items[i];
    int j = i - 1;
    while (j >= b && items[j] > v) {
      items[j + 1] = items[j];
      --j;
    }
    items[j + 1] = v;
  }
}

// A cell polygon lives in SHARED memory, one column per thread (element k of thread t at [k * kCellThreads + t]: no bank
// conflicts), and is clipped in place: as thread-local arrays the polygons spilled to local memory and the kernel moved
// 9 GB through DRAM per 240 k seeds (ncu, profiles/r02_c_*): 2.2 ms; in shared memory it is arithmetic only.
constexpr int kCellThreads = 64;
struct PolyRef {
  double2 *v;  // vertex k = start of edge k
  int *id;     // edge k lies on the bisector of (site, id[k]); -1..-4 = clip rectangle sides
  int m;
  __device__ __forceinline__ double2 &V(int k) { return v[k * kCellThreads]; }
  __device__ __forceinline__ int &I(int k) { return id[k * kCellThreads]; }
};

// clip by the half plane of points at least as close to p as to q; returns false on overflow.  The outside vertices of a
// convex polygon form one cyclic run [a, b]: it is replaced by the two crossing points (the first keeps edge a-1's
// label... i.e. the crossing on edge a-1 starts the new edge along the bisector of q, the crossing on edge b continues edge b).
__device__ bool clip(PolyRef &poly, double2 p, double2 q, int qid) {
  const double nx = q.x - p.x, ny = q.y - p.y;
  const double mx = 0.5 * (p.x + q.x), my = 0.5 * (p.y + q.y);
  const int m = poly.m;
  auto side = [&](int k) {  // > 0: outside (closer to q)
    const double2 v = poly.V(k);
    return (v.x - mx) * nx + (v.y - my) * ny;
  };
  // first outside vertex whose predecessor is inside, and the length of the outside run
  int a = -1, n_out = 0;
  bool prev_in = side(m - 1) <= 0;
  for (int k = 0; k < m; ++k) {
    const bool in = side(k) <= 0;
    if (!in) {
      ++n_out;
      if (prev_in) a = k;
    }
    prev_in = in;
  }
  if (n_out == 0) return true;
  if (n_out == m || a < 0) return true;  // cannot happen for a site's own cell (p is inside every half plane); keep the cell
  const int b = (a + n_out - 1) % m;     // last outside vertex of the run
  const int ap = a == 0 ? m - 1 : a - 1, bn = b + 1 == m ? 0 : b + 1;
  // crossing on edge ap -> a (leaving) and on edge b -> bn (entering)
  const double s0 = side(ap), s1 = side(a), s2 = side(b), s3 = side(bn);
  const double t0 = s0 / (s0 - s1), t1 = s2 / (s2 - s3);
  const double2 va = poly.V(ap), vb = poly.V(a), vc = poly.V(b), vd = poly.V(bn);
  const double2 x0 = make_double2(va.x + t0 * (vb.x - va.x), va.y + t0 * (vb.y - va.y));
  const double2 x1 = make_double2(vc.x + t1 * (vd.x - vc.x), vc.y + t1 * (vd.y - vc.y));
  const int id_b = poly.I(b);  // the edge that continues after the entering crossing
  const int new_m = m - n_out + 2;
  if (new_m > kMaxPoly) return false;
  // rotate so that the run starts at index a with a <= b (no wrap): if it wraps, rotate the kept part to the front
  if (a + n_out > m) {
    // outside run wraps: kept vertices are (b+1 .. a-1), contiguous.  Move them to the front, then append x0, x1.
    const int keep = m - n_out, first = bn;
    for (int k = 0; k < keep; ++k) {
      poly.V(k) = poly.V(first + k);
      poly.I(k) = poly.I(first + k);
    }
    poly.V(keep) = x0;
    poly.I(keep) = qid;
    poly.V(keep + 1) = x1;
    poly.I(keep + 1) = id_b;
  } else {
    // [0 .. a-1] kept, [a .. b] replaced by x0, x1, [b+1 .. m-1] kept and shifted
    const int shift = 2 - n_out;
    if (shift > 0) {
      for (int k = m - 1; k > b; --k) {
        poly.V(k + shift) = poly.V(k);
        poly.I(k + shift) = poly.I(k);
      }
    } else if (shift < 0) {
      for (int k = b + 1; k < m; ++k) {
        poly.V(k + shift) = poly.V(k);
        poly.I(k + shift) = poly.I(k);
      }
    }
    poly.V(a) = x0;
    poly.I(a) = qid;
    poly.V(a + 1) = x1;
    poly.I(a + 1) = id_b;
  }
  poly.m = new_m;
  return true;
}

// One thread per site: the cell polygon, written as up to kMaxOut float2 at scratch[i * kMaxOut], count in cnt[i]
__global__ void __launch_bounds__(kCellThreads) vc_cell_kernel(VcParams P, const double2 *__restrict__ sites,
                                                               const int *__restrict__ site_cell,
                                                               const uint32_t *__restrict__ cell_off, const int *__restrict__ items,
                                                               float2 *__restrict__ scratch, uint32_t *__restrict__ cnt,
                                                               int *__restrict__ err) {
  __shared__ double2 s_v[kMaxPoly * kCellThreads];
  __shared__ int s_id[kMaxPoly * kCellThreads];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > P.n) return;
  if (i == P.n) {
    cnt[i] = 0;  // closes the scan
    return;
  }
  cnt[i] = 0;
  if (i < kOuter || site_cell[i] < 0) return;
  const double2 p = sites[i];
  const int cell = site_cell[i];
  const int gx = cell % P.gnx, gy = cell / P.gnx;
  // an earlier seed with the same float32 coordinates: Subdiv2D::insert returns the existing vertex (LOC_VERTEX), the
  // later seed has no vertex and no facet
  for (uint32_t k = cell_off[cell]; k < cell_off[cell + 1]; ++k) {
    const int j = items[k];
    if (j < i && sites[j].x == p.x && sites[j].y == p.y) return;
  }
  PolyRef poly{s_v + threadIdx.x, s_id + threadIdx.x, 4};  // counter-clockwise, as getVoronoiFacetList's polygons
  poly.V(0) = make_double2(P.cx0, P.cy0);
  poly.V(1) = make_double2(P.cx1, P.cy0);
  poly.V(2) = make_double2(P.cx1, P.cy1);
  poly.V(3) = make_double2(P.cx0, P.cy1);
  poly.I(0) = -1;
  poly.I(1) = -2;
  poly.I(2) = -3;
  poly.I(3) = -4;
  bool ok = true;
  for (int o = 0; o < kOuter && ok; ++o) ok = clip(poly, p, sites[o], o);
  const int max_ring = max(P.gnx, P.gny);
  for (int r = 0; r <= max_ring && ok; ++r) {
    // ring r of grid cells (r = 0: the site's own cell)
    for (int dy = -r; dy <= r && ok; ++dy) {
      const int yy = gy + dy;
      if (yy < 0 || yy >= P.gny) continue;
      const int step = (dy == -r || dy == r) ? 1 : 2 * r;
      for (int dx = -r; dx <= r && ok; dx += (step > 0 ? step : 1)) {
        const int xx = gx + dx;
        if (xx < 0 || xx >= P.gnx) continue;
        const int c2 = yy * P.gnx + xx;
        for (uint32_t k = cell_off[c2]; k < cell_off[c2 + 1] && ok; ++k) {
          const int j = items[k];
          if (j == i) continue;
          const double2 q = sites[j];
          if (q.x == p.x && q.y == p.y) continue;  // a later duplicate of this seed
          ok = clip(poly, p, q, j);
        }
      }
    }
    // every unseen seed is farther than r * kGridCell: it can only cut a cell whose radius exceeds half of that
    double rad2 = 0;
    for (int k = 0; k < poly.m; ++k) {
      const double2 v = poly.V(k);
      const double ddx = v.x - p.x, ddy = v.y - p.y;
      rad2 = fmax(rad2, ddx * ddx + ddy * ddy);
    }
    const double reach = (double)r * kGridCell;
    if (reach * reach >= 4.0 * rad2) break;
  }
  if (!ok || poly.m > kMaxOut) {
    atomicExch(err, 1);
    return;
  }
  if (poly.m < 2) return;
  // Where Subdiv2D starts the facet: at rot(vtx.firstEdge), and firstEdge is the edge most recently (re)assigned at the
  // vertex.  An edge to a LATER-inserted neighbour w is assigned when w is inserted and survives (flipping it away would
  // need a still later neighbour), so firstEdge = the edge to the highest-index neighbour whenever that index exceeds the
  // seed's own, and the facet starts at the vertex that begins the bisector edge of that neighbour (verified against
  // cv2 on every facet of the C2 maps).  A seed whose neighbours are all earlier (about 3 %) keeps the edge last assigned
  // during its own insertion, which depends on the flip order: those start at the smallest neighbour id instead.
  int start = 0, hi = 0;
  for (int k = 1; k < poly.m; ++k) {
    if (poly.I(k) < poly.I(start)) start = k;
    if (poly.I(k) > poly.I(hi)) hi = k;
  }
  if (poly.I(hi) > i) start = hi;
  for (int t = 0; t < poly.m; ++t) {
    const int k = (start + t) % poly.m, kp = k == 0 ? poly.m - 1 : k - 1;
    const int a = poly.I(kp), b = poly.I(k);  // vertex k is where the bisectors of (i, a) and (i, b) meet
    const double2 vk = poly.V(k);
    float2 out;
    if (a >= 0 && b >= 0 && a != b) {
      int t0 = i, t1 = a, t2 = b;  // ascending site order: the same bits from all three cells
      if (t0 > t1) { int s = t0; t0 = t1; t1 = s; }
      if (t1 > t2) { int s = t1; t1 = t2; t2 = s; }
      if (t0 > t1) { int s = t0; t0 = t1; t1 = s; }
      out = circumcentre_cv(sites[t0], sites[t1], sites[t2]);
      if (!(fabsf(out.x) < FLT_MAX * 0.5f) || !(fabsf(out.y) < FLT_MAX * 0.5f)) out = make_float2((float)vk.x, (float)vk.y);
    } else {
      out = make_float2((float)vk.x, (float)vk.y);  // on the clip rectangle: outside the grid, cropped later
    }
    scratch[(size_t)i * kMaxOut + t] = out;
  }
  cnt[i] = (uint32_t)poly.m;
}

__global__ void vc_fill_kernel(int n, const float2 *__restrict__ scratch, const uint32_t *__restrict__ base,
                               float2 *__restrict__ fxy, int *__restrict__ enext) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t b = base[i], m = base[i + 1] - b;
  for (uint32_t j = 0; j < m; ++j) {
    fxy[b + j] = scratch[(size_t)i * kMaxOut + j];
    enext[b + j] = (int)(j + 1 == m ? b : b + j + 1);
  }
}
}  // namespace

// seeds: merged seeds (host, page-locked or not), as VoronoiDiagram::compute receives them; bounds as processGraph passes
// them.  *n_slots = total facet vertices, or -1 when a cell overflowed the fixed-size polygons (caller falls back).
aos_status vcells_prepare(Ctx *c, const double *seeds, int n_seeds, double min_x, double max_x, double min_y, double max_y,
                          int *n_slots) {
  *n_slots = 0;
  if (n_seeds <= 0) return AOS_OK;
  if (!std::isfinite(min_x) || !std::isfinite(max_x) || !std::isfinite(min_y) || !std::isfinite(max_y)) return AOS_OK;
  if (min_x > max_x) std::swap(min_x, max_x);
  if (min_y > max_y) std::swap(min_y, max_y);
  const double min_size = 1.0;  // vd:36-47
  if (max_x - min_x < min_size) {
    double m = (min_x + max_x) / 2.0;
    min_x = m - min_size / 2.0;
    max_x = m + min_size / 2.0;
  }
  if (max_y - min_y < min_size) {
    double m = (min_y + max_y) / 2.0;
    min_y = m - min_size / 2.0;
    max_y = m + min_size / 2.0;
  }
  VcParams P;
  P.n = n_seeds + kOuter;
  P.rx = (float)(min_x - 1.0);
  P.ry = (float)(min_y - 1.0);
  P.rw = (float)(fabs(max_x - min_x) + 2.0);
  P.rh = (float)(fabs(max_y - min_y) + 2.0);
  if (P.rw <= 0 || P.rh <= 0) return AOS_OK;
  const int irx = (int)lrint((double)P.rx), iry = (int)lrint((double)P.ry), irw = (int)lrint((double)P.rw), irh = (int)lrint((double)P.rh);
  const float big = g_outer_factor * (float)(irw > irh ? irw : irh);
  P.cx0 = (double)P.rx - kClipMargin;
  P.cy0 = (double)P.ry - kClipMargin;
  P.cx1 = (double)P.rx + (double)P.rw + kClipMargin;
  P.cy1 = (double)P.ry + (double)P.rh + kClipMargin;
  P.gx0 = (double)P.rx;
  P.gy0 = (double)P.ry;
  P.gnx = (int)ceil((double)P.rw / kGridCell) + 1;
  P.gny = (int)ceil((double)P.rh / kGridCell) + 1;
  const size_t ncell = (size_t)P.gnx * P.gny;
  if (ncell > ((size_t)1 << 28)) {
    set_error(c, "device Voronoi: seed rectangle too large for the neighbour grid");
    return AOS_ERR_CAPACITY;
  }
  cudaStream_t st = c->stream;
  AOS_CUDA_OK(c, c->sd_verts.reserve(sizeof(double2) * (size_t)P.n + sizeof(double) * 2 * (size_t)n_seeds));
  AOS_CUDA_OK(c, c->sd_quads.reserve(sizeof(int) * 2 * (size_t)P.n));
  AOS_CUDA_OK(c, c->vc_cells.reserve(sizeof(uint32_t) * 2 * (ncell + 1)));
  AOS_CUDA_OK(c, c->sd_vor.reserve(sizeof(float2) * kMaxOut * (size_t)P.n));
  AOS_CUDA_OK(c, c->sd_base.reserve(sizeof(uint32_t) * ((size_t)P.n + 1)));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  double2 *sites = c->sd_verts.as<double2>();
  double *d_seeds = reinterpret_cast<double *>(sites + P.n);
  int *site_cell = c->sd_quads.as<int>(), *items = site_cell + P.n;
  uint32_t *cell_off = c->vc_cells.as<uint32_t>(), *cursor = cell_off + ncell + 1;
  int *d_err = c->misc.as<int>() + 96;
  uint32_t *d_tot = reinterpret_cast<uint32_t *>(c->misc.as<int>() + 97);
  aos_status hs = h2d_small(c, d_seeds, seeds, sizeof(double) * 2 * (size_t)n_seeds, seeds == c->h_merged.data());
  if (hs != AOS_OK) return hs;
  AOS_CUDA_OK(c, cudaMemsetAsync(cell_off, 0, sizeof(uint32_t) * 2 * (ncell + 1), st));
  AOS_CUDA_OK(c, cudaMemsetAsync(d_err, 0, 8, st));
  const int tb = 256;
  vc_sites_kernel<<<(P.n + tb - 1) / tb, tb, 0, st>>>(d_seeds, P, big, (float)irx, (float)iry, sites, site_cell, cell_off);
  ++c->launches;
  aos_status s = exclusive_scan_u32(c, cell_off, ncell + 1, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  vc_bin_kernel<<<(P.n + tb - 1) / tb, tb, 0, st>>>(site_cell, P.n, cell_off, cursor, items);
  ++c->launches;
  vc_sort_kernel<<<(unsigned)((ncell + tb - 1) / tb), tb, 0, st>>>(cell_off, (int)ncell, items);
  ++c->launches;
  vc_cell_kernel<<<(P.n + 1 + kCellThreads - 1) / kCellThreads, kCellThreads, 0, st>>>(P, sites, site_cell, cell_off, items, c->sd_vor.as<float2>(),
                                                     c->sd_base.as<uint32_t>(), d_err);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  s = exclusive_scan_u32(c, c->sd_base.as<uint32_t>(), (size_t)P.n + 1, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_err, 8, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  if (c->h_flag[0] != 0) {
    *n_slots = -1;
    return AOS_OK;
  }
  *n_slots = c->h_flag[1];
  c->vc_n = P.n;
  return AOS_OK;
}

aos_status vcells_fill(Ctx *c, float2 *d_fxy, int *d_enext) {
  vc_fill_kernel<<<(c->vc_n + 255) / 256, 256, 0, c->stream>>>(c->vc_n, c->sd_vor.as<float2>(), c->sd_base.as<uint32_t>(), d_fxy,
                                                               d_enext);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

}  // namespace aos
