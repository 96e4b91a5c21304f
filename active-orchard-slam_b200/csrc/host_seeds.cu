// host_seeds.cu -- the part of the path that stays on the HOST (north star: "seed selection ... stay on
// the host"): generateVirtualSeeds / raycastToOccupiedCell / generateRayPointsFromEndpoints /
// castRayFromEndpoint / endpoint seeds / publish order of aos_seed_gen_node
// (src/aos_seed_gen_node.cpp:1434-1511, 1670-1710, 1730-1771, 1774-1891, 1894-1982, 1987-2268, 2546-2582)
// and the greedy 0.5 m merge of aos_gvd_node::voronoiSeedsCallback (src/aos_gvd_node.cpp:84-128).
//
// Same arithmetic as the reference (double everywhere the reference uses double, float where it uses
// float, glibc cos/sin), different data structures: the skeleton is read from the bit-packed grid, and
// the reference's O(S^2) first-come "is there an earlier seed within 0.5 m" scans run on a uniform hash
// grid, which returns the same answers.
#include <math.h>

#include <algorithm>
#include <unordered_map>

#include "aos_common.cuh"

namespace aos {

namespace {

struct HostGrid {
  const uint32_t *bits;
  int w, h, pitch;
  double ox, oy;
  float res;
  bool occ(int x, int y) const { return (bits[(size_t)y * pitch + (x >> 5)] >> (x & 31)) & 1u; }
};

// First-come duplicate filter: "some earlier accepted point is closer than r" (strict <, sqrt distance as
// in the reference's std::sqrt(std::pow(dx,2)+std::pow(dy,2)) < 0.5).
class FirstComeSet {
 public:
  explicit FirstComeSet(double r) : r_(r), inv_(1.0 / r) {}
  bool has_within(double x, double y) const {
    long long cx = (long long)floor(x * inv_), cy = (long long)floor(y * inv_);
    for (long long dy = -1; dy <= 1; ++dy)
      for (long long dx = -1; dx <= 1; ++dx) {
        auto it = cells_.find(key(cx + dx, cy + dy));
        if (it == cells_.end()) continue;
        for (int i : it->second) {
          double ex = pts_[2 * i] - x, ey = pts_[2 * i + 1] - y;
          if (sqrt(ex * ex + ey * ey) < r_) return true;
        }
      }
    return false;
  }
  void add(double x, double y) {
    int i = (int)(pts_.size() / 2);
    pts_.push_back(x);
    pts_.push_back(y);
    cells_[key((long long)floor(x * inv_), (long long)floor(y * inv_))].push_back(i);
  }
  const std::vector<double> &points() const { return pts_; }

 private:
  static unsigned long long key(long long cx, long long cy) {
    return ((unsigned long long)(cx + (1ll << 30)) << 32) ^ (unsigned long long)(cy + (1ll << 30));
  }
  double r_, inv_;
  std::vector<double> pts_;
  std::unordered_map<unsigned long long, std::vector<int>> cells_;
};

// isPointInPolygon, seed_gen:1231-1255
bool in_polygon(double px, double py, const double *poly, int n) {
  if (n < 3) return false;
  bool inside = false;
  int j = n - 1;
  for (int i = 0; i < n; ++i) {
    double pix = poly[2 * i], piy = poly[2 * i + 1], pjx = poly[2 * j], pjy = poly[2 * j + 1];
    double dy = pjy - piy;
    if (fabs(dy) > 1e-9) {
      if (((piy > py) != (pjy > py)) && (px < (pjx - pix) * (py - piy) / dy + pix)) inside = !inside;
    }
    j = i;
  }
  return inside;
}

// worldToGrid, seed_gen:760-769
void world_to_grid(const HostGrid &g, float wx, float wy, int *gx, int *gy) {
  float rel_x = (float)((wx - g.ox) / g.res);
  float rel_y = (float)((wy - g.oy) / g.res);
  *gx = std::max(0, std::min(g.w - 1, (int)floorf(rel_x)));
  *gy = std::max(0, std::min(g.h - 1, (int)floorf(rel_y)));
}

// raycastToOccupiedCell, seed_gen:1730-1771
bool raycast_to_occupied(const HostGrid &g, double sx, double sy, double dx, double dy, double max_distance,
                         double *hx, double *hy) {
  const double step = g.res * 0.5;
  const int max_steps = (int)(max_distance / step);
  double cx = sx, cy = sy;
  for (int i = 0; i < max_steps; ++i) {
    cx += dx * step;
    cy += dy * step;
    double ex = cx - sx, ey = cy - sy;
    if (sqrt(ex * ex + ey * ey) < 1.0) continue;
    int gx, gy;
    world_to_grid(g, (float)cx, (float)cy, &gx, &gy);
    if (g.occ(gx, gy)) {
      *hx = cx;
      *hy = cy;
      return true;
    }
  }
  return false;
}

void normalize2(double &x, double &y) {  // Eigen normalize(): z = squaredNorm; if (z > 0) v /= sqrt(z)
  double z = x * x + y * y;
  if (z > 0) {
    double s = sqrt(z);
    x /= s;
    y /= s;
  }
}

// castRayFromEndpoint, seed_gen:1774-1891 (fixed 0.1 m steps from min_distance)
void cast_ray_from_endpoint(const HostGrid &g, double spx, double spy, double opx, double opy, double angle_deg,
                            double min_distance, double *rx, double *ry) {
  double ex = opx - spx, ey = opy - spy;
  if (sqrt(ex * ex + ey * ey) < 1e-6) {
    ex = 1.0;
    ey = 0.0;
  } else {
    normalize2(ex, ey);
  }
  const double outx = -ex, outy = -ey, perpx = -ey, perpy = ex;
  const double a = angle_deg * M_PI / 180.0;
  double rdx, rdy;
  if (angle_deg > 0) {
    rdx = cos(a) * outx + sin(a) * perpx;
    rdy = cos(a) * outy + sin(a) * perpy;
  } else {
    rdx = cos(-a) * outx + sin(-a) * (-perpx);
    rdy = cos(-a) * outy + sin(-a) * (-perpy);
  }
  normalize2(rdx, rdy);
  // info.width * info.resolution is uint32 * float -> float (seed_gen:1808-1810)
  const double gw = (float)((float)(unsigned)g.w * g.res), gh = (float)((float)(unsigned)g.h * g.res);
  const double minx = g.ox, maxx = g.ox + gw, miny = g.oy, maxy = g.oy + gh;
  const double resolution = g.res;
  const double abs_max = sqrt(gw * gw + gh * gh) * 3.0;
  double cur = min_distance;
  while (cur <= abs_max) {
    double px = spx + rdx * cur, py = spy + rdy * cur;
    if (!(px >= minx && px <= maxx && py >= miny && py <= maxy)) {
      *rx = std::max(minx, std::min(maxx, px));
      *ry = std::max(miny, std::min(maxy, py));
      return;
    }
    int mx = (int)((px - g.ox) / resolution), my = (int)((py - g.oy) / resolution);
    if (mx >= 0 && mx < g.w && my >= 0 && my < g.h && g.occ(mx, my)) {
      *rx = px;
      *ry = py;
      return;
    }
    cur += 0.1;
  }
  double fx = spx + rdx * abs_max, fy = spy + rdy * abs_max;
  if (!(fx >= minx && fx <= maxx && fy >= miny && fy <= maxy)) {
    fx = std::max(minx, std::min(maxx, fx));
    fy = std::max(miny, std::min(maxy, fy));
  }
  *rx = fx;
  *ry = fy;
}

}  // namespace

// rows: all_tree_rows in cluster order.  Produces /voronoi_seeds (virtual, ray, endpoint seeds in publish
// order, seed_gen:1670-1710) and /exploration_tree_rows_info (rows sorted by centre y then x, :2546-2582).
void host_select_seeds(const uint32_t *skel_bits, int w, int h, int pitch, double ox, double oy, float res,
                       const std::vector<aos_tree_row> &rows, const double *poly, int n_poly,
                       std::vector<double> *seeds, int counts[3], std::vector<double> *rows_info) {
  HostGrid g{skel_bits, w, h, pitch, ox, oy, res};
  const bool use_poly = n_poly > 0;
  FirstComeSet virt(0.5), ray(0.5), endp(0.5);
  const double interval = 1.0;  // virtual_seed_interval_, seed_gen:2666
  // generateVirtualSeeds, seed_gen:1987-2268 (real_seeds_ is always empty)
  for (const aos_tree_row &r : rows) {
    if (use_poly && !in_polygon(r.center_x, r.center_y, poly, n_poly)) continue;
    double dx = r.end_x - r.start_x, dy = r.end_y - r.start_y;
    double distance = sqrt(dx * dx + dy * dy);
    if (distance < interval) continue;
    double nrm = sqrt(dx * dx + dy * dy);
    if (nrm < 1e-6) continue;
    double rdx = dx / nrm, rdy = dy / nrm;
    const double pdx[2] = {-rdy, rdy}, pdy[2] = {rdx, -rdx};
    int num = (int)floor(distance / interval);
    for (int i = 1; i <= num; ++i) {
      double t = (double)i / (num + 1);
      double bx = r.start_x + t * dx, by = r.start_y + t * dy;
      if (!virt.has_within(bx, by)) virt.add(bx, by);
      for (int side = 0; side < 2; ++side) {
        double hx, hy, sx, sy;
        if (raycast_to_occupied(g, bx, by, pdx[side], pdy[side], 4.0, &hx, &hy)) {
          sx = hx;
          sy = hy;
        } else {
          sx = bx + pdx[side] * 4.0;
          sy = by + pdy[side] * 4.0;
        }
        if (use_poly && in_polygon(sx, sy, poly, n_poly)) continue;
        if (!virt.has_within(sx, sy)) virt.add(sx, sy);
      }
    }
  }
  // generateRayPointsFromEndpoints, seed_gen:1894-1982
  {
    const double gw = (float)((float)(unsigned)w * res), gh = (float)((float)(unsigned)h * res);
    const double minx = ox, maxx = ox + gw, miny = oy, maxy = oy + gh;
    for (const aos_tree_row &r : rows) {
      double rp[12];
      const double ang[3] = {0.0, -90.0, 90.0};
      for (int k = 0; k < 3; ++k) cast_ray_from_endpoint(g, r.start_x, r.start_y, r.end_x, r.end_y, ang[k], 1.0, &rp[2 * k], &rp[2 * k + 1]);
      for (int k = 0; k < 3; ++k) cast_ray_from_endpoint(g, r.end_x, r.end_y, r.start_x, r.start_y, ang[k], 1.0, &rp[6 + 2 * k], &rp[7 + 2 * k]);
      for (int k = 0; k < 6; ++k) {
        double x = rp[2 * k], y = rp[2 * k + 1];
        if (!std::isfinite(x) || !std::isfinite(y)) continue;
        if (!(x >= minx && x <= maxx && y >= miny && y <= maxy)) continue;
        if (use_poly && in_polygon(x, y, poly, n_poly)) continue;
        if (!ray.has_within(x, y)) ray.add(x, y);
      }
    }
  }
  // endpoint seeds, seed_gen:1450-1496
  for (const aos_tree_row &r : rows) {
    if (!endp.has_within(r.start_x, r.start_y)) endp.add(r.start_x, r.start_y);
    if (!endp.has_within(r.end_x, r.end_y)) endp.add(r.end_x, r.end_y);
  }
  seeds->clear();
  seeds->insert(seeds->end(), virt.points().begin(), virt.points().end());
  seeds->insert(seeds->end(), ray.points().begin(), ray.points().end());
  seeds->insert(seeds->end(), endp.points().begin(), endp.points().end());
  counts[0] = (int)(virt.points().size() / 2);
  counts[1] = (int)(ray.points().size() / 2);
  counts[2] = (int)(endp.points().size() / 2);

  host_rows_info(rows, rows_info);
}

// publishExplorationTreeRowsInfoFromClusters (seed_gen:2546-2582): sort by centre (y, then x if |dy| < 1e-6);
// stable, so rows with equal keys keep cluster order
void host_rows_info(const std::vector<aos_tree_row> &rows, std::vector<double> *rows_info) {
  std::vector<int> ord(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) ord[i] = (int)i;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
    if (fabs(rows[a].center_y - rows[b].center_y) < 1e-6) return rows[a].center_x < rows[b].center_x;
    return rows[a].center_y < rows[b].center_y;
  });
  rows_info->clear();
  for (int i : ord) {
    rows_info->push_back(rows[i].start_x);
    rows_info->push_back(rows[i].start_y);
    rows_info->push_back(rows[i].end_x);
    rows_info->push_back(rows[i].end_y);
  }
}

// voronoiSeedsCallback, gvd:84-128: greedy leader clustering at 0.5 m (<=), centroid in index order.
// Same answers as the reference's O(S^2) scan; the neighbour candidates come from a flat open-addressing grid
// (cell = merge distance) instead.
void host_merge_seeds(const double *seeds, int n, std::vector<double> *out) {
  const double merge_distance = 0.5;
  out->clear();
  if (n <= 0) return;
  size_t cap = 64;
  while (cap < (size_t)n * 2) cap <<= 1;
  const size_t mask = cap - 1;
  std::vector<unsigned long long> keys(cap, ~0ull);
  std::vector<int> head(cap, -1), next((size_t)n, -1);
  std::vector<char> used((size_t)n, 0), finite((size_t)n, 1);
  auto cell = [&](double v) { return (long long)floor(v / merge_distance); };
  auto key = [](long long cx, long long cy) {
    return ((unsigned long long)(cx + (1ll << 30)) << 32) ^ (unsigned long long)(cy + (1ll << 30));
  };
  auto slot_of = [&](unsigned long long k, bool insert) -> long long {
    unsigned long long h = k;
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    size_t i = (size_t)h & mask;
    for (;;) {
      if (keys[i] == k) return (long long)i;
      if (keys[i] == ~0ull) {
        if (!insert) return -1;
        keys[i] = k;
        return (long long)i;
      }
      i = (i + 1) & mask;
    }
  };
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(seeds[2 * i]) || !std::isfinite(seeds[2 * i + 1])) {
      finite[i] = 0;  // a non-finite seed never merges with anything (norm is NaN/inf) and is dropped at gvd:266
      continue;
    }
    long long s = slot_of(key(cell(seeds[2 * i]), cell(seeds[2 * i + 1])), true);
    next[i] = head[s];
    head[s] = i;
  }
  std::vector<int> members;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    used[i] = 1;
    if (!finite[i]) continue;  // its own cluster, filtered out by processGraph (gvd:266-270)
    members.clear();
    const long long cx = cell(seeds[2 * i]), cy = cell(seeds[2 * i + 1]);
    for (long long dy = -1; dy <= 1; ++dy)
      for (long long dx = -1; dx <= 1; ++dx) {
        long long s = slot_of(key(cx + dx, cy + dy), false);
        if (s < 0) continue;
        for (int j = head[s]; j >= 0; j = next[j]) {
          if (j <= i || used[j]) continue;
          double ex = seeds[2 * i] - seeds[2 * j], ey = seeds[2 * i + 1] - seeds[2 * j + 1];
          if (sqrt(ex * ex + ey * ey) <= merge_distance) members.push_back(j);
        }
      }
    std::sort(members.begin(), members.end());
    double sx = seeds[2 * i], sy = seeds[2 * i + 1];
    for (int j : members) {
      used[j] = 1;
      sx += seeds[2 * j];
      sy += seeds[2 * j + 1];
    }
    double cnt = (double)(members.size() + 1);
    out->push_back(sx / cnt);
    out->push_back(sy / cnt);
  }
}

}  // namespace aos
