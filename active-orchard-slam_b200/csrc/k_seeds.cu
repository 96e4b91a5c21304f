// k_seeds.cu -- seed selection on the device: the ray casts and first-come
// 0.5 m de-duplications of aos_seed_gen_node
//   generateVirtualSeeds             src/aos_seed_gen_node.cpp:1987-2268   (+ raycastToOccupiedCell :1730-1771)
//   generateRayPointsFromEndpoints   :1894-1982                            (+ castRayFromEndpoint :1774-1891)
//   endpoint seeds                   :1450-1496
// Every candidate (row, sample, side) is independent: one thread walks one ray over the un-framed skeleton bit
// grid with the reference's arithmetic (double steps, float worldToGrid, glibc cos/sin of 0 and pi/2 evaluated
// on the host).  The reference then filters each candidate list first-come ("is an earlier accepted seed
// closer than 0.5 m"): that greedy rule is resolved exactly by monotone rounds over a 0.5 m hash grid, as for
// the graph nodes (k_graph.cu), and the survivors are emitted in order by a prefix sum.
#include <math.h>

#include <algorithm>

#include <float.h>

#include "aos_common.cuh"
#include "dev_hash.cuh"

namespace aos {

struct RowDev {
  double cx, cy, sx, sy, ex, ey;
};

struct SeedGrid {
  const uint32_t *bits;  // un-framed skeleton
  int w, h, pitch;
  double ox, oy;
  float res;
};

__device__ __forceinline__ bool sg_occ(const SeedGrid &g, int x, int y) {
  return (__ldg(g.bits + (size_t)y * g.pitch + (x >> 5)) >> (x & 31)) & 1u;
}

// isPointInPolygon, seed_gen:1231-1255
__device__ __forceinline__ bool in_polygon_dev(const SeedDeviceParams &P, double px, double py) {
  if (P.n_poly < 3) return false;
  bool inside = false;
  int j = P.n_poly - 1;
  for (int i = 0; i < P.n_poly; ++i) {
    double pix = P.poly[2 * i], piy = P.poly[2 * i + 1], pjx = P.poly[2 * j], pjy = P.poly[2 * j + 1];
    double dy = pjy - piy;
    if (fabs(dy) > 1e-9) {
      if (((piy > py) != (pjy > py)) && (px < (pjx - pix) * (py - piy) / dy + pix)) inside = !inside;
    }
    j = i;
  }
  return inside;
}

// worldToGrid, seed_gen:760-769
__device__ __forceinline__ void world_to_grid_dev(const SeedGrid &g, float wx, float wy, int *gx, int *gy) {
  float rel_x = (float)(((double)wx - g.ox) / (double)g.res);
  float rel_y = (float)(((double)wy - g.oy) / (double)g.res);
  *gx = max(0, min(g.w - 1, (int)floorf(rel_x)));
  *gy = max(0, min(g.h - 1, (int)floorf(rel_y)));
}

// raycastToOccupiedCell, seed_gen:1730-1771
// The same cell without the two double divisions per ray step: rel = float(q) with q = (w - o) / res, cell = floorf(rel).
// With q placed by a reciprocal multiplication (relative error of a few 2^-53) the cell can differ from the literal one
// only when q is within that error of an integer from above, or within half a float32 ulp (<= |q| 2^-24) of one from below
// (where the float rounding reaches the integer): those steps take the literal expression.
__device__ __forceinline__ void world_to_grid_fast(const SeedGrid &g, double inv_res, float wx, float wy, int *gx, int *gy) {
  const double qx = ((double)wx - g.ox) * inv_res, qy = ((double)wy - g.oy) * inv_res;
  const double fx = qx - floor(qx), fy = qy - floor(qy);
  const double hx = fabs(qx) * 6.0e-8 + 1e-6, hy = fabs(qy) * 6.0e-8 + 1e-6;
  if (!(fx > 1e-6 && fx < 1.0 - hx && fy > 1e-6 && fy < 1.0 - hy)) {
    world_to_grid_dev(g, wx, wy, gx, gy);
    return;
  }
  *gx = max(0, min(g.w - 1, (int)floorf((float)qx)));
  *gy = max(0, min(g.h - 1, (int)floorf((float)qy)));
}

__device__ bool raycast_to_occupied_dev(const SeedGrid &g, double sx, double sy, double dx, double dy, double max_distance,
                                        double *hx, double *hy) {
  const double step = (double)g.res * 0.5;
  const int max_steps = (int)(max_distance / step);
  const double inv_res = 1.0 / (double)g.res;
  double cx = sx, cy = sy;
  for (int i = 0; i < max_steps; ++i) {
    cx += dx * step;
    cy += dy * step;
    double ex = cx - sx, ey = cy - sy;
    // the reference tests sqrt(z) < 1.0; a correctly rounded square root is below 1 exactly when z is (the largest
    // double below 1 has a root below the midpoint to 1), so the root itself is not needed
    if (ex * ex + ey * ey < 1.0) continue;
    int gx, gy;
    world_to_grid_fast(g, inv_res, (float)cx, (float)cy, &gx, &gy);
    if (sg_occ(g, gx, gy)) {
      *hx = cx;
      *hy = cy;
      return true;
    }
  }
  return false;
}

__device__ __forceinline__ void normalize2_dev(double &x, double &y) {
  double z = x * x + y * y;
  if (z > 0) {
    double s = sqrt(z);
    x /= s;
    y /= s;
  }
}

struct RayConsts {
  double c[3], s[3];  // cos / sin as castRayFromEndpoint evaluates them for angle 0, -90, +90 (glibc, host)
  double gw, gh;      // float(width * resolution), float(height * resolution)
};

// castRayFromEndpoint, seed_gen:1774-1891, one WARP per ray.  The reference advances `cur += 0.1` step by step
// (a sequential floating-point accumulation) until the sample leaves the grid or lands on a skeleton cell; rays
// along a row can run for 10^4 steps.  Every ray starts at min_distance = 1.0, so the accumulated values are the same
// sequence for all of them: T[k] = 1.0 + 0.1 + ... (k additions, rounded one by one), built once on the host
// (ray_steps_table) and shared by every ray of every map.  Lane l evaluates step 32 i + l; the first lane whose step
// terminates the loop wins.  Entries past the table are above abs_max by construction.
__device__ double2 cast_ray_from_endpoint_warp(const SeedGrid &g, const RayConsts &K, double spx, double spy, double opx,
                                               double opy, int ang /* 0: 0 deg, 1: -90, 2: +90 */,
                                               const double *__restrict__ T, int nT, int lane) {
  double ex = opx - spx, ey = opy - spy;
  if (sqrt(ex * ex + ey * ey) < 1e-6) {
    ex = 1.0;
    ey = 0.0;
  } else {
    normalize2_dev(ex, ey);
  }
  const double outx = -ex, outy = -ey, perpx = -ey, perpy = ex;
  double rdx, rdy;
  if (ang == 2) {  // angle > 0
    rdx = K.c[2] * outx + K.s[2] * perpx;
    rdy = K.c[2] * outy + K.s[2] * perpy;
  } else {
    rdx = K.c[ang] * outx + K.s[ang] * (-perpx);
    rdy = K.c[ang] * outy + K.s[ang] * (-perpy);
  }
  normalize2_dev(rdx, rdy);
  const double minx = g.ox, maxx = g.ox + K.gw, miny = g.oy, maxy = g.oy + K.gh;
  const double resolution = (double)g.res;
  const double abs_max = sqrt(K.gw * K.gw + K.gh * K.gh) * 3.0;
  for (int k = lane;; k += 32) {
    const double mine = k < nT ? T[k] : DBL_MAX;
    // status of this lane's step: 0 continue, 1 left the grid, 2 hit a cell, 3 past abs_max (loop condition fails)
    int status = 0;
    double px = 0.0, py = 0.0;
    if (!(mine <= abs_max)) {
      status = 3;
    } else {
      px = spx + rdx * mine;
      py = spy + rdy * mine;
      if (!(px >= minx && px <= maxx && py >= miny && py <= maxy)) {
        status = 1;
      } else {
        int mx = (int)((px - g.ox) / resolution), my = (int)((py - g.oy) / resolution);
        if (mx >= 0 && mx < g.w && my >= 0 && my < g.h && sg_occ(g, mx, my)) status = 2;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, status != 0);
    if (m) {
      const int f = __ffs(m) - 1;
      const int st = __shfl_sync(0xffffffffu, status, f);
      px = __shfl_sync(0xffffffffu, px, f);
      py = __shfl_sync(0xffffffffu, py, f);
      if (st == 1) return make_double2(fmax(minx, fmin(maxx, px)), fmax(miny, fmin(maxy, py)));
      if (st == 2) return make_double2(px, py);
      break;
    }
  }
  double fx = spx + rdx * abs_max, fy = spy + rdy * abs_max;
  if (!(fx >= minx && fx <= maxx && fy >= miny && fy <= maxy)) {
    fx = fmax(minx, fmin(maxx, fx));
    fy = fmax(miny, fmin(maxy, fy));
  }
  return make_double2(fx, fy);
}

enum : unsigned char { kSeedUndecided = 0, kSeedAccept = 1, kSeedReject = 2 };

// ---- generateVirtualSeeds ------------------------------------------------------------------------------
__global__ void vs_count_kernel(const __grid_constant__ SeedDeviceParams P, const RowDev *__restrict__ rows, int n_rows,
                                uint32_t *__restrict__ counts) {
  const double interval = 1.0;  // virtual_seed_interval_, seed_gen:2666
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += gridDim.x * blockDim.x) {
    uint32_t c = 0;
    if (r < n_rows) {
      const RowDev R = rows[r];
      bool ok = !(P.n_poly > 0 && !in_polygon_dev(P, R.cx, R.cy));
      double dx = R.ex - R.sx, dy = R.ey - R.sy;
      double distance = sqrt(dx * dx + dy * dy);
      if (distance < interval) ok = false;
      if (distance < 1e-6) ok = false;
      if (ok) c = 3u * (uint32_t)(int)floor(distance / interval);
    }
    counts[r] = c;
  }
}

__global__ void vs_generate_kernel(const __grid_constant__ SeedDeviceParams P, SeedGrid g, const RowDev *__restrict__ rows,
                                   int n_rows, const uint32_t *__restrict__ offs, int n_cand, double2 *__restrict__ pts,
                                   unsigned char *__restrict__ state) {
  const double interval = 1.0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_cand; idx += gridDim.x * blockDim.x) {
    // row of this candidate: last r with offs[r] <= idx
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (offs[mid] <= (uint32_t)idx) lo = mid;
      else hi = mid - 1;
    }
    const RowDev R = rows[lo];
    const int local = idx - (int)offs[lo];
    const int i = local / 3 + 1, which = local % 3;
    const double dx = R.ex - R.sx, dy = R.ey - R.sy;
    const double distance = sqrt(dx * dx + dy * dy);
    const double rdx = dx / distance, rdy = dy / distance;
    const int num = (int)floor(distance / interval);
    const double t = (double)i / (double)(num + 1);
    const double bx = R.sx + t * dx, by = R.sy + t * dy;
    double2 out;
    bool valid = true;
    if (which == 0) {
      out = make_double2(bx, by);
    } else {
      const double pdx = which == 1 ? -rdy : rdy, pdy = which == 1 ? rdx : -rdx;
      double hx, hy;
      if (raycast_to_occupied_dev(g, bx, by, pdx, pdy, 4.0, &hx, &hy)) out = make_double2(hx, hy);
      else out = make_double2(bx + pdx * 4.0, by + pdy * 4.0);
      if (P.n_poly > 0 && in_polygon_dev(P, out.x, out.y)) valid = false;
    }
    pts[idx] = out;
    state[idx] = valid ? kSeedUndecided : kSeedReject;
  }
}

// ---- generateRayPointsFromEndpoints + endpoint seeds ---------------------------------------------------------
__global__ void ray_points_kernel(const __grid_constant__ SeedDeviceParams P, SeedGrid g, RayConsts K,
                                  const RowDev *__restrict__ rows, int n_rows, const double *__restrict__ T, int nT,
                                  double2 *__restrict__ ray_pts,
                                  unsigned char *__restrict__ ray_state, double2 *__restrict__ end_pts,
                                  unsigned char *__restrict__ end_state) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < 6 * n_rows; idx += warps) {
    const int r = idx / 6, k = idx % 6;
    const RowDev R = rows[r];
    const bool from_start = k < 3;
    const double2 p = cast_ray_from_endpoint_warp(g, K, from_start ? R.sx : R.ex, from_start ? R.sy : R.ey,
                                                  from_start ? R.ex : R.sx, from_start ? R.ey : R.sy, k % 3, T, nT, lane);
    if (lane != 0) continue;
    const double minx = g.ox, maxx = g.ox + K.gw, miny = g.oy, maxy = g.oy + K.gh;
    bool valid = isfinite(p.x) && isfinite(p.y) && (p.x >= minx && p.x <= maxx && p.y >= miny && p.y <= maxy);
    if (valid && P.n_poly > 0 && in_polygon_dev(P, p.x, p.y)) valid = false;
    ray_pts[idx] = p;
    ray_state[idx] = valid ? kSeedUndecided : kSeedReject;
    if (k < 2) {  // endpoint seeds: start, end of every row
      end_pts[2 * r + k] = k == 0 ? make_double2(R.sx, R.sy) : make_double2(R.ex, R.ey);
      end_state[2 * r + k] = kSeedUndecided;
    }
  }
}

// ---- first-come de-duplication: "an earlier accepted point is closer than r" (sqrt distance, strict <) ----
__global__ void seed_grid_build_kernel(const double2 *__restrict__ pts, const unsigned char *__restrict__ state, int n,
                                       PointGrid g) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (state[i] == kSeedReject) continue;  // filtered candidates never enter the reference's lists
    double2 p = pts[i];
    int slot = hash_insert(g.h, cell_key(cell_coord(p.x, g.inv), cell_coord(p.y, g.inv)));
    g.next[i] = atomicExch(&g.h.val[slot], i);
  }
}

// The three candidate lists (virtual seeds, ray points, end points) are filtered independently by the reference, each
// against its own earlier points; they lie one behind the other in pts (boundaries b1, b2), so one set of rounds serves
// all three: a point only looks at earlier points of its own list.  prev_flag (may be null): the previous round of the
// same batch; a round whose predecessor left nothing pending exits at once, so the host synchronises once per batch.
__global__ void seed_dedup_round_kernel(const double2 *__restrict__ pts, int n, int b1, int b2, PointGrid g,
                                        volatile unsigned char *state, double radius, const int *prev_flag, int *pending_flag) {
  if (prev_flag && *prev_flag == 0) return;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] != kSeedUndecided) continue;
    const int first = v >= b2 ? b2 : v >= b1 ? b1 : 0;  // start of v's own list
    const double2 p = pts[v];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    bool reject = false, pending = false;
    for (int oy = -1; oy <= 1 && !reject; ++oy)
      for (int ox = -1; ox <= 1 && !reject; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
          if (u >= v || u < first) continue;
          unsigned char su = state[u];
          if (su == kSeedReject) continue;
          double ex = pts[u].x - p.x, ey = pts[u].y - p.y;
          if (!(sqrt(ex * ex + ey * ey) < radius)) continue;
          if (su == kSeedAccept) {
            reject = true;
            break;
          }
          pending = true;
        }
      }
    if (reject) state[v] = kSeedReject;
    else if (!pending) state[v] = kSeedAccept;
    else *pending_flag = 1;
  }
}

__global__ void seed_flags_kernel(const unsigned char *__restrict__ state, int n, uint32_t *__restrict__ flags) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flags[i] = state[i] == kSeedAccept;
}
__global__ void seed_emit_kernel(const double2 *__restrict__ pts, const unsigned char *__restrict__ state,
                                 const uint32_t *__restrict__ pos, int n, double2 *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (state[i] == kSeedAccept) out[pos[i]] = pts[i];
}

namespace {
inline int blocks_for(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  b = std::min<size_t>(std::max<size_t>(b, 1), (size_t)kNumSMs * 16);
  return (int)b;
}
inline unsigned pow2_at_least(size_t n) {
  unsigned c = 64;
  while (c < n) c <<= 1;
  return c;
}

// De-duplicates the three candidate lists pts[0..b1), [b1..b2), [b2..n) in place of the reference's FirstComeSets; the
// survivors go to `out` in order, n_out[k] of list k.  d_flags: kRoundBatch ints; d_tot: 3 uint32.
constexpr int kRoundBatch = 4;
aos_status dedup_and_fetch(Ctx *c, double2 *d_pts, unsigned char *d_state, int n, int b1, int b2, PointGrid g, unsigned cap,
                           uint32_t *d_scan, double2 *d_out, int *d_flags, uint32_t *d_tot, int n_out[3]) {
  cudaStream_t st = c->stream;
  n_out[0] = n_out[1] = n_out[2] = 0;
  if (n == 0) return AOS_OK;
  AOS_CUDA_OK(c, cudaMemsetAsync(g.h.keys, 0xff, sizeof(unsigned long long) * cap, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(g.h.val, 0xff, sizeof(int) * cap, st));
  seed_grid_build_kernel<<<blocks_for(n), 256, 0, st>>>(d_pts, d_state, n, g);
  ++c->launches;
  for (int batch = 0; batch < 100000; ++batch) {
    AOS_CUDA_OK(c, cudaMemsetAsync(d_flags, 0, sizeof(int) * kRoundBatch, st));
    for (int j = 0; j < kRoundBatch; ++j)
      seed_dedup_round_kernel<<<blocks_for(n), 256, 0, st>>>(d_pts, n, b1, b2, g, d_state, 0.5, j ? d_flags + j - 1 : nullptr,
                                                             d_flags + j);
    c->launches += kRoundBatch;
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_flags, sizeof(int) * kRoundBatch, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    bool done = false;
    for (int j = 0; j < kRoundBatch; ++j) done |= c->h_flag[j] == 0;  // a round that left nothing pending
    if (done) break;
  }
  seed_flags_kernel<<<blocks_for(n), 256, 0, st>>>(d_state, n, d_scan);
  ++c->launches;
  aos_status s = exclusive_scan_u32(c, d_scan, (size_t)n, c->cc_blocksum, d_tot + 2);  // total of all three
  if (s != AOS_OK) return s;
  seed_emit_kernel<<<blocks_for(n), 256, 0, st>>>(d_pts, d_state, d_scan, n, d_out);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  // survivors before b1, before b2, in all: the exclusive scan at the list boundaries
  if (b1 < n) AOS_CUDA_OK(c, cudaMemcpyAsync(d_tot, d_scan + b1, 4, cudaMemcpyDeviceToDevice, st));
  if (b2 < n) AOS_CUDA_OK(c, cudaMemcpyAsync(d_tot + 1, d_scan + b2, 4, cudaMemcpyDeviceToDevice, st));
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 12, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int total = c->h_flag[2];
  const int upto1 = b1 < n ? c->h_flag[0] : total, upto2 = b2 < n ? c->h_flag[1] : total;
  n_out[0] = upto1;
  n_out[1] = upto2 - upto1;
  n_out[2] = total - upto2;
  return AOS_OK;
}
}  // namespace

// ---- voronoiSeedsCallback's greedy 0.5 m merge (src/aos_gvd_node.cpp:84-128) -------------------------------------
// The reference walks the seeds in order; an unused seed i becomes a leader and absorbs every later unused seed j
// with |s_i - s_j| <= 0.5 (distance to the LEADER, sqrt of the squared norm, inclusive), and the cluster's point is
// the mean taken in index order.  So: j is a leader iff no earlier leader lies within 0.5 m -- the same monotone
// rounds as the first-come filters above -- a member belongs to the EARLIEST such leader, and each leader adds up
// its members in ascending index.  Non-finite seeds never merge (their norm is NaN/inf) and are dropped afterwards
// (gvd:266-270), so they simply stay out of the grid.
__global__ void merge_init_kernel(const double2 *__restrict__ pts, int n, unsigned char *__restrict__ state) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 p = pts[i];
    state[i] = (isfinite(p.x) && isfinite(p.y)) ? kSeedUndecided : kSeedReject;
  }
}

__global__ void merge_round_kernel(const double2 *__restrict__ pts, int n, PointGrid g, volatile unsigned char *state,
                                   double radius, const int *prev_flag, int *pending_flag) {
  if (prev_flag && *prev_flag == 0) return;  // the previous round of this batch left nothing pending
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] != kSeedUndecided) continue;
    const double2 p = pts[v];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    bool member = false, pending = false;
    for (int oy = -1; oy <= 1 && !member; ++oy)
      for (int ox = -1; ox <= 1 && !member; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
          if (u >= v) continue;
          unsigned char su = state[u];
          if (su == kSeedReject) continue;  // a member absorbs nobody
          double ex = pts[u].x - p.x, ey = pts[u].y - p.y;
          if (!(sqrt(ex * ex + ey * ey) <= radius)) continue;
          if (su == kSeedAccept) {
            member = true;
            break;
          }
          pending = true;
        }
      }
    if (member) state[v] = kSeedReject;
    else if (!pending) state[v] = kSeedAccept;
    else *pending_flag = 1;
  }
}

// member -> its leader: the earliest leader within the radius
__global__ void merge_owner_kernel(const double2 *__restrict__ pts, int n, PointGrid g, const unsigned char *__restrict__ state,
                                   double radius, int *__restrict__ owner) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    int best = -1;
    const double2 p = pts[v];
    if (state[v] == kSeedReject && isfinite(p.x) && isfinite(p.y)) {
      const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
          if (slot < 0) continue;
          for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
            if (u >= v || state[u] != kSeedAccept || (best >= 0 && u > best)) continue;
            double ex = pts[u].x - p.x, ey = pts[u].y - p.y;
            if (sqrt(ex * ex + ey * ey) <= radius) best = u;
          }
        }
    }
    owner[v] = best;
  }
}

// leader -> mean of (itself, its members in ascending index), written at its rank among the leaders
__global__ void merge_emit_kernel(const double2 *__restrict__ pts, int n, PointGrid g, const unsigned char *__restrict__ state,
                                  const int *__restrict__ owner, const uint32_t *__restrict__ pos, double2 *__restrict__ out) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] != kSeedAccept) continue;
    const double2 p = pts[v];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    double sx = p.x, sy = p.y;
    int last = v, count = 1;
    for (;;) {  // next member in index order; clusters hold a handful of seeds
      int nxt = 0x7fffffff;
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
          if (slot < 0) continue;
          for (int u = g.h.val[slot]; u >= 0; u = g.next[u])
            if (u > last && u < nxt && owner[u] == v) nxt = u;
        }
      if (nxt == 0x7fffffff) break;
      sx += pts[nxt].x;
      sy += pts[nxt].y;
      ++count;
      last = nxt;
    }
    const double cnt = (double)count;
    out[pos[v]] = make_double2(sx / cnt, sy / cnt);
  }
}

// rows: all_tree_rows in cluster order (c->h_rows).  Fills c->h_seeds / c->seed_counts.
aos_status device_select_seeds(Ctx *c) {
  cudaStream_t st = c->stream;
  const SeedDeviceParams &P = c->P;
  const int n_rows = (int)c->h_rows.size();
  c->h_seeds.resize(0);
  c->seed_counts[0] = c->seed_counts[1] = c->seed_counts[2] = 0;
  if (n_rows == 0) return AOS_OK;
  if (!c->pin_a.resize(sizeof(RowDev) * (size_t)n_rows)) {
    set_error(c, "cudaHostAlloc failed (rows staging)");
    return AOS_ERR_CUDA;
  }
  RowDev *hr = reinterpret_cast<RowDev *>(c->pin_a.data());
  for (int i = 0; i < n_rows; ++i) {
    const aos_tree_row &r = c->h_rows[i];
    hr[i] = RowDev{r.center_x, r.center_y, r.start_x, r.start_y, r.end_x, r.end_y};
  }
  // device memory, part 1: rows + counts
  AOS_CUDA_OK(c, c->seed_buf.reserve(sizeof(RowDev) * (size_t)n_rows + sizeof(uint32_t) * ((size_t)n_rows + 16) + 4096));
  RowDev *d_rows = c->seed_buf.as<RowDev>();
  uint32_t *d_offs = reinterpret_cast<uint32_t *>(d_rows + n_rows);
  uint32_t *d_tot = d_offs + n_rows + 2;  // [0..3] totals, [4] pending flag
  {
    aos_status hs = h2d_small(c, d_rows, hr, sizeof(RowDev) * (size_t)n_rows, true);
    if (hs != AOS_OK) return hs;
  }
  vs_count_kernel<<<blocks_for((size_t)n_rows + 1), 256, 0, st>>>(P, d_rows, n_rows, d_offs);
  ++c->launches;
  aos_status s = exclusive_scan_u32(c, d_offs, (size_t)n_rows + 1, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int n_virt = c->h_flag[0], n_ray = 6 * n_rows, n_end = 2 * n_rows;
  const int n_max = n_virt + n_ray + n_end;  // the three lists are filtered in one pass

  // part 2: candidates
  const unsigned cap = pow2_at_least((size_t)n_max * 2);
  size_t need = (sizeof(double2) * 2 + 1 + sizeof(int) + sizeof(uint32_t)) * ((size_t)n_virt + n_ray + n_end + 64) +
                (sizeof(unsigned long long) + sizeof(int)) * (size_t)cap + 8192;
  AOS_CUDA_OK(c, c->seed_buf2.reserve(need));
  char *base = c->seed_buf2.as<char>();
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    char *p = base + off;
    off += bytes;
    return p;
  };
  const size_t n_all = (size_t)n_virt + n_ray + n_end;
  double2 *d_pts = reinterpret_cast<double2 *>(take(sizeof(double2) * (n_all + 1)));
  double2 *d_out = reinterpret_cast<double2 *>(take(sizeof(double2) * (n_all + 1)));
  unsigned char *d_state = reinterpret_cast<unsigned char *>(take(n_all + 1));
  uint32_t *d_scan = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)n_max + 1)));
  PointGrid g;
  g.h.keys = reinterpret_cast<unsigned long long *>(take(sizeof(unsigned long long) * cap));
  g.h.val = reinterpret_cast<int *>(take(sizeof(int) * cap));
  g.h.mask = cap - 1;
  g.next = reinterpret_cast<int *>(take(sizeof(int) * ((size_t)n_max + 1)));
  g.inv = 1.0 / (0.5 * (1.0 + 1e-9));

  SeedGrid sg{c->g_skel.as<uint32_t>(), P.w, P.h, P.pitch, P.ox, P.oy, P.res};
  RayConsts K;
  const double ang[3] = {0.0, -90.0, 90.0};
  for (int k = 0; k < 3; ++k) {  // castRayFromEndpoint: cos(a), sin(a) for a > 0, cos(-a), sin(-a) otherwise
    double a = ang[k] * M_PI / 180.0;
    K.c[k] = ang[k] > 0 ? cos(a) : cos(-a);
    K.s[k] = ang[k] > 0 ? sin(a) : sin(-a);
  }
  K.gw = (double)(float)((float)(unsigned)P.w * P.res);  // info.width * info.resolution is uint32 * float -> float
  K.gh = (double)(float)((float)(unsigned)P.h * P.res);

  double2 *d_vpts = d_pts, *d_rpts = d_pts + n_virt, *d_epts = d_rpts + n_ray;
  unsigned char *d_vst = d_state, *d_rst = d_state + n_virt, *d_est = d_rst + n_ray;
  // the perpendicular rays of the virtual seeds and the endpoint rays are independent and both latency-bound: side by side
  AOS_CUDA_OK(c, cudaEventRecord(c->ev_fork, st));
  if (n_virt > 0) {
    AOS_CUDA_OK(c, cudaStreamWaitEvent(c->aux[0], c->ev_fork, 0));
    vs_generate_kernel<<<blocks_for(n_virt, 128), 128, 0, c->aux[0]>>>(P, sg, d_rows, n_rows, d_offs, n_virt, d_vpts, d_vst);
    ++c->launches;
    AOS_CUDA_OK(c, cudaEventRecord(c->ev_join[0], c->aux[0]));
  }
  // the accumulated ray parameter 1.0 + 0.1 + 0.1 + ... of castRayFromEndpoint (seed_gen:1840-1880), one addition at a time
  // as the reference performs them; the same for every ray, so it is built once and only ever extended
  {
    const double abs_max = sqrt(K.gw * K.gw + K.gh * K.gh) * 3.0;
    if (!(abs_max < 1e7)) {
      set_error(c, "grid too large for the ray step table");
      return AOS_ERR_CAPACITY;
    }
    std::vector<double> &T = c->ray_steps;
    if (T.empty()) T.push_back(1.0);
    const size_t before = T.size();
    while (T.back() <= abs_max) T.push_back(T.back() + 0.1);
    if (T.size() != before || c->ray_steps_dev != T.size()) {
      AOS_CUDA_OK(c, c->ray_table.reserve(sizeof(double) * T.size()));
      AOS_CUDA_OK(c, cudaMemcpyAsync(c->ray_table.as<double>(), T.data(), sizeof(double) * T.size(), cudaMemcpyHostToDevice, st));
      AOS_CUDA_OK(c, cudaStreamSynchronize(st));  // pageable source: once per context (and per larger grid)
      c->ray_steps_dev = T.size();
    }
  }
  ray_points_kernel<<<blocks_for((size_t)n_ray * 32, 128), 128, 0, st>>>(P, sg, K, d_rows, n_rows, c->ray_table.as<double>(),
                                                                         (int)c->ray_steps_dev, d_rpts, d_rst, d_epts, d_est);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  if (n_virt > 0) AOS_CUDA_OK(c, cudaStreamWaitEvent(st, c->ev_join[0], 0));

  int counts[3] = {0, 0, 0};
  int *d_flags = reinterpret_cast<int *>(d_tot + 4);  // kRoundBatch ints
  s = dedup_and_fetch(c, d_pts, d_state, (int)n_all, n_virt, n_virt + n_ray, g, cap, d_scan, d_out, d_flags, d_tot + 1, counts);
  if (s != AOS_OK) return s;
  const int total = counts[0] + counts[1] + counts[2];
  if (!c->h_seeds.resize(2 * (size_t)total)) {
    set_error(c, "cudaHostAlloc failed (seeds)");
    return AOS_ERR_CUDA;
  }
  if (total > 0)
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_seeds.data(), d_out, sizeof(double2) * (size_t)total, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  for (int k = 0; k < 3; ++k) c->seed_counts[k] = counts[k];
  return AOS_OK;
}

// seeds (host, n x,y pairs) -> c->h_merged (merged seeds in leader order, non-finite dropped)
aos_status device_merge_seeds(Ctx *c, const double *seeds, int n) {
  cudaStream_t st = c->stream;
  c->h_merged.resize(0);
  if (n <= 0) return AOS_OK;
  const unsigned cap = pow2_at_least((size_t)n * 2);
  const size_t N = (size_t)n;
  size_t need = sizeof(double2) * 2 * (N + 1) + (N + 1) + (sizeof(int) * 2 + sizeof(uint32_t)) * (N + 1) +
                (sizeof(unsigned long long) + sizeof(int)) * (size_t)cap + 8192;
  AOS_CUDA_OK(c, c->seed_buf2.reserve(need));
  char *base = c->seed_buf2.as<char>();
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    char *p = base + off;
    off += bytes;
    return p;
  };
  double2 *d_pts = reinterpret_cast<double2 *>(take(sizeof(double2) * (N + 1)));
  double2 *d_out = reinterpret_cast<double2 *>(take(sizeof(double2) * (N + 1)));
  unsigned char *d_state = reinterpret_cast<unsigned char *>(take(N + 1));
  uint32_t *d_scan = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (N + 1)));
  int *d_owner = reinterpret_cast<int *>(take(sizeof(int) * (N + 1)));
  PointGrid g;
  g.h.keys = reinterpret_cast<unsigned long long *>(take(sizeof(unsigned long long) * cap));
  g.h.val = reinterpret_cast<int *>(take(sizeof(int) * cap));
  g.h.mask = cap - 1;
  g.next = reinterpret_cast<int *>(take(sizeof(int) * (N + 1)));
  g.inv = 1.0 / (0.5 * (1.0 + 1e-9));
  uint32_t *d_tot = reinterpret_cast<uint32_t *>(take(64));
  int *d_flag = reinterpret_cast<int *>(d_tot + 4);

  if (seeds != c->h_seeds.data()) {  // caller-owned (pageable) memory: stage it, the copy below must be a plain DMA
    if (!c->pin_seed_in.resize(2 * N)) {
      set_error(c, "cudaHostAlloc failed (seed staging)");
      return AOS_ERR_CUDA;
    }
    memcpy(c->pin_seed_in.data(), seeds, sizeof(double2) * N);
    seeds = c->pin_seed_in.data();
  }
  {
    aos_status hs = h2d_small(c, d_pts, seeds, sizeof(double2) * N, true);
    if (hs != AOS_OK) return hs;
  }
  AOS_CUDA_OK(c, cudaMemsetAsync(g.h.keys, 0xff, sizeof(unsigned long long) * cap, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(g.h.val, 0xff, sizeof(int) * cap, st));
  merge_init_kernel<<<blocks_for(N), 256, 0, st>>>(d_pts, n, d_state);
  ++c->launches;
  seed_grid_build_kernel<<<blocks_for(N), 256, 0, st>>>(d_pts, d_state, n, g);
  ++c->launches;
  for (int batch = 0; batch < 100000; ++batch) {  // kRoundBatch rounds per host synchronisation
    AOS_CUDA_OK(c, cudaMemsetAsync(d_flag, 0, sizeof(int) * kRoundBatch, st));
    for (int j = 0; j < kRoundBatch; ++j)
      merge_round_kernel<<<blocks_for(N), 256, 0, st>>>(d_pts, n, g, d_state, 0.5, j ? d_flag + j - 1 : nullptr, d_flag + j);
    c->launches += kRoundBatch;
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_flag, sizeof(int) * kRoundBatch, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    bool done = false;
    for (int j = 0; j < kRoundBatch; ++j) done |= c->h_flag[j] == 0;
    if (done) break;
  }
  merge_owner_kernel<<<blocks_for(N), 256, 0, st>>>(d_pts, n, g, d_state, 0.5, d_owner);
  ++c->launches;
  seed_flags_kernel<<<blocks_for(N), 256, 0, st>>>(d_state, n, d_scan);
  ++c->launches;
  aos_status s = exclusive_scan_u32(c, d_scan, N, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  merge_emit_kernel<<<blocks_for(N), 256, 0, st>>>(d_pts, n, g, d_state, d_owner, d_scan, d_out);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int m = c->h_flag[0];
  if (!c->h_merged.resize(2 * (size_t)m)) {
    set_error(c, "cudaHostAlloc failed (merged seeds)");
    return AOS_ERR_CUDA;
  }
  if (m > 0) {
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_merged.data(), d_out, sizeof(double2) * (size_t)m, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  }
  return AOS_OK;
}

}  // namespace aos
