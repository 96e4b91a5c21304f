// k_graph.cu -- the gvd half on the device: everything of AosGvdNode::processGraph after the Voronoi facets
// exist (src/aos_gvd_node.cpp:255-318):
//   extractBoundaryPoints            src/utils/voronoi_diagram.cpp:149-207   first-come 5 cm merge -> graph nodes
//   buildGraphFromBoundaryPoints     gvd:794-895   nearest node per edge end, (min,max) de-duplication,
//                                                  skeleton crossing test (gvd:320-359), 0.5 m proximity edges
//   filterNodesAndEdgesOutsideGrid   gvd:420-483   crop + reindex
//   findClusterEndpointVoronoiBoundaryPoints  gvd:485-556, 686-790, castRay gvd:558-684
//   publishGraph's label arrays      gvd:897-1010
// The reference's loops are O(E*M) / O(M^2) linear scans; here every "is there an earlier / a nearest point"
// question goes through a uniform hash grid, and every order-dependent rule (first-come merge, first
// occurrence of an edge key, (i,j) ordering of the proximity edges) is reproduced exactly: the first-come
// merge is resolved by rounds of a monotone fixed point (a point is accepted once every earlier point within
// the threshold is rejected, rejected as soon as one of them is accepted), ordered outputs are produced by
// prefix-sum stream compaction.  Arithmetic is the reference's (double, no FMA contraction: -fmad=false).
#include <float.h>
#include <limits.h>
#include <math.h>

#include <algorithm>

#include "aos_common.cuh"
#include "dev_hash.cuh"

namespace aos {

__global__ void grid_build_kernel(const double2 *__restrict__ pts, int n, PointGrid g) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double2 p = pts[i];
    int slot = hash_insert(g.h, cell_key(cell_coord(p.x, g.inv), cell_coord(p.y, g.inv)));
    g.next[i] = atomicExch(&g.h.val[slot], i);
  }
}

// ---------------------------------------------------------------------------------------------------
// extractBoundaryPoints, vd:149-207
// ---------------------------------------------------------------------------------------------------
// static_cast<int>(double) as x86-64 evaluates it (cvttsd2si: out-of-range and NaN give INT_MIN)
__device__ __forceinline__ int x86_int(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int)v;
}
__device__ __forceinline__ bool key_overflows(double2 p) {
  double a = p.x * 100, b = p.y * 100;
  return !(a > -2147483649.0 && a < 2147483648.0) || !(b > -2147483649.0 && b < 2147483648.0);
}
__device__ __forceinline__ bool same_int_key(double2 a, double2 b) {
  return x86_int(a.x * 100) == x86_int(b.x * 100) && x86_int(a.y * 100) == x86_int(b.y * 100);
}

__global__ void facet_points_kernel(const float2 *__restrict__ fxy, int n, double2 *__restrict__ pts, int *ovf_list,
                                    int *ovf_count, int ovf_cap) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float2 f = fxy[i];
    double2 p = make_double2((double)f.x, (double)f.y);
    pts[i] = p;
    if (key_overflows(p)) {
      int k = atomicAdd(ovf_count, 1);
      if (k < ovf_cap) ovf_list[k] = i;
    }
  }
}

enum : unsigned char { kUndecided = 0, kAccept = 1, kReject = 2 };

// One round of the first-come merge over the facet-vertex slots (slot order == the order in which the
// reference meets the points: edge e = (slot e, next slot) contributes its start, then its end).
__global__ void boundary_round_kernel(const double2 *__restrict__ pts, int n, PointGrid g, volatile unsigned char *state,
                                      const int *__restrict__ ovf_list, int n_ovf, const int *prev_flag, int *pending_flag) {
  if (prev_flag && *prev_flag == 0) return;  // the previous round of this batch left nothing pending
  const double thr = 0.05;
  const double thr2 = thr * thr;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] != kUndecided) continue;
    const double2 p = pts[v];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    bool reject = false, pending = false;
    for (int oy = -1; oy <= 1 && !reject; ++oy)
      for (int ox = -1; ox <= 1 && !reject; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
          if (u >= v) continue;
          unsigned char su = state[u];
          if (su == kReject) continue;
          double2 q = pts[u];
          double dx = q.x - p.x, dy = q.y - p.y;
          if (!(dx * dx + dy * dy < thr2) && !same_int_key(p, q)) continue;
          if (su == kAccept) {
            reject = true;
            break;
          }
          pending = true;
        }
      }
    if (!reject && n_ovf > 0 && key_overflows(p)) {  // integer keys that collide far apart (overflowed casts)
      for (int k = 0; k < n_ovf; ++k) {
        int u = ovf_list[k];
        if (u >= v) continue;
        unsigned char su = state[u];
        if (su == kReject || !same_int_key(p, pts[u])) continue;
        if (su == kAccept) {
          reject = true;
          break;
        }
        pending = true;
      }
    }
    if (reject) state[v] = kReject;
    else if (!pending) state[v] = kAccept;
    else *pending_flag = 1;
  }
}

__global__ void flags_from_state_kernel(const unsigned char *__restrict__ state, int n, uint32_t *__restrict__ flags) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flags[i] = state[i] == kAccept;
}

__global__ void compact_nodes_kernel(const double2 *__restrict__ pts, const unsigned char *__restrict__ state,
                                     const uint32_t *__restrict__ rank, int n, double2 *__restrict__ nodes) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (state[i] == kAccept) nodes[rank[i]] = pts[i];
}

// nearest node of every facet-vertex slot (gvd:812-830: linear scan with strict <, i.e. lowest index among
// equally near nodes).  A rejected slot has an accepted one within 5 cm, so its nearest node is in the 3x3
// block of 5 cm cells; slots rejected only by a far-away integer-key twin fall back to the full scan.
__global__ void nearest_node_kernel(const double2 *__restrict__ pts, const unsigned char *__restrict__ state,
                                    const uint32_t *__restrict__ rank, int n, PointGrid g, const double2 *__restrict__ nodes,
                                    int n_nodes, int *__restrict__ nn) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] == kAccept) {
      nn[v] = (int)rank[v];
      continue;
    }
    const double2 p = pts[v];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    double best = DBL_MAX;
    int best_i = -1;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
          if (state[u] != kAccept) continue;
          double dx = pts[u].x - p.x, dy = pts[u].y - p.y;
          double d = sqrt(dx * dx + dy * dy);
          int id = (int)rank[u];
          if (d < best || (d == best && id < best_i)) {
            best = d;
            best_i = id;
          }
        }
      }
    if (best_i < 0 || best >= 0.05) {  // not guaranteed to be the global nearest: scan everything
      best = DBL_MAX;
      best_i = -1;
      for (int i = 0; i < n_nodes; ++i) {
        double dx = nodes[i].x - p.x, dy = nodes[i].y - p.y;
        double d = sqrt(dx * dx + dy * dy);
        if (d < best) {
          best = d;
          best_i = i;
        }
      }
    }
    nn[v] = best_i;
  }
}

// ---------------------------------------------------------------------------------------------------
// edgePassesThroughOccupiedPixels, gvd:320-359 -- one warp per segment, lanes stride over the samples
// ---------------------------------------------------------------------------------------------------
struct GridView {
  const uint32_t *bits;  // framed skeleton, 32 cells per word
  int w, h, pitch;
  double ox, oy;
  float res;
};

__device__ __forceinline__ bool grid_occ(const GridView &g, int mx, int my) {
  return (__ldg(g.bits + (size_t)my * g.pitch + (mx >> 5)) >> (mx & 31)) & 1u;
}

__device__ bool warp_segment_hits(const GridView &g, double2 s, double2 e, int lane) {
  const double resolution = (double)g.res;
  const double ex = e.x - s.x, ey = e.y - s.y;
  const double z = ex * ex + ey * ey;
  const double edge_length = sqrt(z);
  if (edge_length < 1e-6) return false;
  const double sample_step = resolution * 0.5;
  const double ratio = edge_length / sample_step;
  if (!(ratio < 2147483648.0)) return false;  // (int) overflows to INT_MIN: num_samples < 0, the loop body never runs
  const int num_samples = (int)ratio + 1;
  double dx = ex, dy = ey;  // normalized()
  if (z > 0) {
    double q = sqrt(z);
    dx = ex / q;
    dy = ey / q;
  }
  // samples outside the grid never hit: for long segments (far-away Voronoi vertices) restrict the index range
  // to a conservative superset of the samples that can fall inside the grid
  int i_lo = 0, i_hi = num_samples;
  if (num_samples > 2048) {
    const double x0 = g.ox - 4 * resolution, x1 = g.ox + (g.w + 4) * resolution;
    const double y0 = g.oy - 4 * resolution, y1 = g.oy + (g.h + 4) * resolution;
    double t0 = 0.0, t1 = 1.0;
    if (ex != 0.0) {
      double a = (x0 - s.x) / ex, b = (x1 - s.x) / ex;
      t0 = fmax(t0, fmin(a, b));
      t1 = fmin(t1, fmax(a, b));
    } else if (s.x < x0 || s.x > x1) {
      t1 = -1.0;
    }
    if (ey != 0.0) {
      double a = (y0 - s.y) / ey, b = (y1 - s.y) / ey;
      t0 = fmax(t0, fmin(a, b));
      t1 = fmin(t1, fmax(a, b));
    } else if (s.y < y0 || s.y > y1) {
      t1 = -1.0;
    }
    if (t0 > t1) return false;
    double lo = floor(t0 * num_samples) - 4.0, hi = ceil(t1 * num_samples) + 4.0;
    i_lo = (int)fmax(lo, 0.0);
    i_hi = (int)fmin(hi, (double)num_samples);
  }
  // Only the CELL a sample falls into matters, and three double divisions per sample (i / num_samples, two by the
  // resolution) are most of this kernel's instructions.  The sample is therefore placed with reciprocal multiplications
  // first (relative error of a few 2^-53, i.e. < 1e-9 cells at these magnitudes) and evaluated with the reference's
  // divisions only when either coordinate is within 1e-6 of a cell boundary -- the only case where the two can disagree.
  const double inv_n = 1.0 / (double)num_samples, inv_res = 1.0 / resolution;
  for (int base = i_lo; base <= i_hi; base += 32) {
    int i = base + lane;
    bool hit = false;
    if (i <= i_hi) {
      double t = (i == num_samples) ? 1.0 : ((double)i * inv_n);
      double px = s.x + (t * dx) * edge_length, py = s.y + (t * dy) * edge_length;
      double fx = (px - g.ox) * inv_res, fy = (py - g.oy) * inv_res;
      const double rx = fx - floor(fx), ry = fy - floor(fy);
      if (!(rx > 1e-6 && rx < 1.0 - 1e-6 && ry > 1e-6 && ry < 1.0 - 1e-6)) {  // near a boundary (or not finite): literally
        t = (i == num_samples) ? 1.0 : ((double)i / (double)num_samples);
        px = s.x + (t * dx) * edge_length, py = s.y + (t * dy) * edge_length;
        fx = (px - g.ox) / resolution, fy = (py - g.oy) / resolution;
      }
      if (fx > -1.0 && fx < (double)g.w && fy > -1.0 && fy < (double)g.h) {  // (int) truncates toward zero
        int mx = (int)fx, my = (int)fy;
        if (mx >= 0 && mx < g.w && my >= 0 && my < g.h) hit = grid_occ(g, mx, my);
      }
    }
    if (__any_sync(0xffffffffu, hit)) return true;
  }
  return false;
}

// Opt-in edge clearance (the reference publishes 0.0f, gvd:856,890): minimum over the same
// samples of the exact distance to the nearest skeleton cell (k_edt.cu), in metres.  Samples outside the grid
// are skipped; an edge without any sample inside the grid gets FLT_MAX.
__global__ void edge_clearance_kernel(const int32_t *__restrict__ edges, int n_edges, const double2 *__restrict__ nodes,
                                      GridView g, const int32_t *__restrict__ dist2, float *__restrict__ clearances) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n_edges; e += warps) {
    const double2 s = nodes[edges[2 * (size_t)e]], t = nodes[edges[2 * (size_t)e + 1]];
    const double resolution = (double)g.res;
    const double ex = t.x - s.x, ey = t.y - s.y;
    const double z = ex * ex + ey * ey;
    const double edge_length = sqrt(z);
    int best = 0x7fffffff;
    if (edge_length >= 1e-6 && edge_length / (resolution * 0.5) < 2147483648.0) {
      const int num_samples = (int)(edge_length / (resolution * 0.5)) + 1;
      const double q = sqrt(z), dx = ex / q, dy = ey / q;
      for (int i = lane; i <= num_samples; i += 32) {
        double tt = (i == num_samples) ? 1.0 : ((double)i / (double)num_samples);
        double px = s.x + (tt * dx) * edge_length, py = s.y + (tt * dy) * edge_length;
        double fx = (px - g.ox) / resolution, fy = (py - g.oy) / resolution;
        if (fx > -1.0 && fx < (double)g.w && fy > -1.0 && fy < (double)g.h) {
          int mx = (int)fx, my = (int)fy;
          if (mx >= 0 && mx < g.w && my >= 0 && my < g.h) best = min(best, dist2[(size_t)my * g.w + mx]);
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) clearances[e] = best == 0x7fffffff ? FLT_MAX : (float)(sqrt((double)best) * resolution);
  }
}

// buildGraphFromBoundaryPoints part 1 (gvd:806-859): every Voronoi edge -> (nearest node of start, of end)
__global__ void edge_key_kernel(const int *__restrict__ nn, const int *__restrict__ enext, int n_edges, DevHash H) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    int a = nn[e], b = nn[enext[e]];
    if (a < 0 || b < 0 || a == b) continue;
    unsigned long long key = ((unsigned long long)(unsigned)min(a, b) << 32) | (unsigned)max(a, b);
    int slot = hash_insert(H, key);
    atomicMin(&H.val[slot], e);
  }
}

// first occurrence of each key is tested against the skeleton; keep[e] = 1 for the edges the reference adds.
// Two steps: one THREAD per facet-vertex slot finds the slots that are the first occurrence of their node pair (a third of
// them) and lists them; one WARP per listed slot samples its segment.  (A warp per slot spent most of its warps on the two
// loads and the hash probe that say "not first".)
__global__ void edge_first_kernel(const int *__restrict__ nn, const int *__restrict__ enext, int n_edges, DevHash H,
                                  uint32_t *__restrict__ keep, int4 *__restrict__ list, int *__restrict__ n_list) {
  const int lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n_edges; base += gridDim.x * blockDim.x) {
    const int e = base + lane;
    bool first = false;
    int a = -1, b = -1, slot = -1;
    if (e < n_edges) {
      keep[e] = 0u;
      a = nn[e];
      b = nn[enext[e]];
      if (a >= 0 && b >= 0 && a != b) {
        unsigned long long key = ((unsigned long long)(unsigned)min(a, b) << 32) | (unsigned)max(a, b);
        slot = hash_find(H, key);
        first = slot >= 0 && H.val[slot] == e;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (!m) continue;
    int at = 0;
    if (lane == 0) at = atomicAdd(n_list, __popc(m));
    at = __shfl_sync(0xffffffffu, at, 0);
    if (first) list[at + __popc(m & ((1u << lane) - 1u))] = make_int4(e, a, b, slot);
  }
}

__global__ void edge_test_kernel(const int4 *__restrict__ list, const int *__restrict__ n_list, int *__restrict__ accepted,
                                 const double2 *__restrict__ nodes, GridView g, uint32_t *__restrict__ keep) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n = *n_list;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int4 it = list[i];
    const bool k = !warp_segment_hits(g, nodes[it.y], nodes[it.z], lane);
    if (lane == 0 && k) {
      keep[it.x] = 1u;
      accepted[it.w] = 1;
    }
  }
}

struct EdgeRec {
  int from, to;
};

__global__ void edge_emit_kernel(const int *__restrict__ nn, const int *__restrict__ enext, int n_edges,
                                 const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos, EdgeRec *__restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    if (!keep[e]) continue;
    int a = nn[e], b = nn[enext[e]];
    out[pos[e]] = EdgeRec{min(a, b), max(a, b)};
  }
}

// proximity edges (gvd:861-894): node pairs i < j with 1e-6 < dist <= 0.5 that are not yet edges
template <bool FILL>
__global__ void proximity_pairs_kernel(const double2 *__restrict__ nodes, int n_nodes, PointGrid g, DevHash H,
                                       const int *__restrict__ accepted, uint32_t *__restrict__ counts /* !FILL: out */,
                                       const uint32_t *__restrict__ offsets /* FILL */, int *__restrict__ pair_j /* FILL */) {
  const double nearby = 0.5;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
    const double2 p = nodes[i];
    const long long cx = cell_coord(p.x, g.inv), cy = cell_coord(p.y, g.inv);
    uint32_t cnt = 0;
    const uint32_t base = FILL ? offsets[i] : 0;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int j = g.h.val[slot]; j >= 0; j = g.next[j]) {
          if (j <= i) continue;
          double dx = p.x - nodes[j].x, dy = p.y - nodes[j].y;
          double dist = sqrt(dx * dx + dy * dy);
          if (!(dist <= nearby && dist > 1e-6)) continue;
          unsigned long long key = ((unsigned long long)(unsigned)i << 32) | (unsigned)j;
          int s = hash_find(H, key);
          if (s >= 0 && accepted[s]) continue;
          if (FILL) pair_j[base + cnt] = j;
          ++cnt;
        }
      }
    if (!FILL) {
      counts[i] = cnt;
    } else {
      // (i, j) pairs are visited in increasing j by the reference: order this node's list
      int *seg = pair_j + base;
      for (uint32_t a = 1; a < cnt; ++a) {
        int key = seg[a];
        int b = (int)a - 1;
        while (b >= 0 && seg[b] > key) {
          seg[b + 1] = seg[b];
          --b;
        }
        seg[b + 1] = key;
      }
    }
  }
}

__global__ void pair_owner_kernel(const uint32_t *__restrict__ offsets, int n_nodes, int *__restrict__ pair_i) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x)
    for (uint32_t k = offsets[i]; k < offsets[i + 1]; ++k) pair_i[k] = i;
}

__global__ void pair_test_kernel(const int *__restrict__ pair_i, const int *__restrict__ pair_j, int n_pairs,
                                 const double2 *__restrict__ nodes, GridView g, uint32_t *__restrict__ keep) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_pairs; k += warps) {
    bool hit = warp_segment_hits(g, nodes[pair_i[k]], nodes[pair_j[k]], lane);
    if (lane == 0) keep[k] = hit ? 0u : 1u;
  }
}

__global__ void pair_emit_kernel(const int *__restrict__ pair_i, const int *__restrict__ pair_j, int n_pairs,
                                 const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos, EdgeRec *__restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pairs; k += gridDim.x * blockDim.x)
    if (keep[k]) out[pos[k]] = EdgeRec{pair_i[k], pair_j[k]};
}

// ---------------------------------------------------------------------------------------------------
// filterNodesAndEdgesOutsideGrid, gvd:420-483
// ---------------------------------------------------------------------------------------------------
struct Bounds {
  double minx, maxx, miny, maxy;
};

__global__ void crop_flag_kernel(const double2 *__restrict__ nodes, int n, Bounds b, uint32_t *__restrict__ flags) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double2 p = nodes[i];
    flags[i] = (p.x >= b.minx && p.x <= b.maxx && p.y >= b.miny && p.y <= b.maxy) ? 1u : 0u;
  }
}

__global__ void crop_nodes_kernel(const double2 *__restrict__ nodes, int n, Bounds b, const uint32_t *__restrict__ remap,
                                  double2 *__restrict__ out, double *__restrict__ out_xyz) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double2 p = nodes[i];
    if (p.x >= b.minx && p.x <= b.maxx && p.y >= b.miny && p.y <= b.maxy) {
      uint32_t k = remap[i];
      out[k] = p;
      out_xyz[3 * (size_t)k] = p.x;
      out_xyz[3 * (size_t)k + 1] = p.y;
      out_xyz[3 * (size_t)k + 2] = 0.0;
    }
  }
}

__device__ __forceinline__ bool in_bounds(double2 p, const Bounds &b) {
  return p.x >= b.minx && p.x <= b.maxx && p.y >= b.miny && p.y <= b.maxy;
}

__global__ void crop_edge_flag_kernel(const EdgeRec *__restrict__ rec, int n, const double2 *__restrict__ nodes, Bounds b,
                                      uint32_t *__restrict__ flags) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    flags[k] = (in_bounds(nodes[rec[k].from], b) && in_bounds(nodes[rec[k].to], b)) ? 1u : 0u;
}

__global__ void crop_edges_kernel(const EdgeRec *__restrict__ rec, int n, const double2 *__restrict__ nodes, Bounds b,
                                  const uint32_t *__restrict__ node_remap, const uint32_t *__restrict__ pos,
                                  int32_t *__restrict__ edges, float *__restrict__ lengths, float *__restrict__ clearances) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double2 f = nodes[rec[k].from], t = nodes[rec[k].to];
    if (!(in_bounds(f, b) && in_bounds(t, b))) continue;
    uint32_t o = pos[k];
    edges[2 * (size_t)o] = (int32_t)node_remap[rec[k].from];  // from < to survives the monotone remap
    edges[2 * (size_t)o + 1] = (int32_t)node_remap[rec[k].to];
    double dx = t.x - f.x, dy = t.y - f.y;
    lengths[o] = (float)sqrt(dx * dx + dy * dy);
    clearances[o] = 0.0f;  // gvd:856,890
  }
}

// ---------------------------------------------------------------------------------------------------
// findVoronoiBoundaryPointNearEndpoint (gvd:686-790) + castRay (gvd:558-684): one warp per (row, corner)
// ---------------------------------------------------------------------------------------------------
struct CornerParams {
  GridView g;
  Bounds b;
  double cos90, sin90;   // glibc cos/sin of pi/2 evaluated on the host
  double gw, gh;         // float(width * resolution), float(height * resolution) widened
  int n_rows, n_nodes;
};

__device__ __forceinline__ void normalize2(double &x, double &y) {
  double z = x * x + y * y;
  if (z > 0) {
    double q = sqrt(z);
    x /= q;
    y /= q;
  }
}

// is node p a candidate for this corner?  returns its distance (or a negative value)
__device__ __forceinline__ double corner_candidate(double2 p, double2 endpoint, double outx, double outy, double perpx,
                                                   double perpy, bool neg, bool pos, double min_distance, double max_radius) {
  double dx = p.x - endpoint.x, dy = p.y - endpoint.y;
  double dist = sqrt(dx * dx + dy * dy);
  if (dist < min_distance || dist > max_radius) return -1.0;
  normalize2(dx, dy);
  double dot_out = outx * dx + outy * dy;
  if (dot_out < 0.0) return -1.0;
  double dot_perp = perpx * dx + perpy * dy;
  if (neg) {
    if (dot_perp > 0.0) return -1.0;
  } else if (pos) {
    if (dot_perp < 0.0) return -1.0;
  }
  return dist;
}

__device__ double2 cast_ray_dev(const CornerParams &P, double2 sp, double2 other, double angle_deg, double min_distance) {
  const GridView &g = P.g;
  double ex = other.x - sp.x, ey = other.y - sp.y;
  if (sqrt(ex * ex + ey * ey) < 1e-6) {
    ex = 1.0;
    ey = 0.0;
  } else {
    normalize2(ex, ey);
  }
  const double outx = -ex, outy = -ey, perpx = -ey, perpy = ex;
  double rdx, rdy;
  // cos(+-a) with a = +-pi/2: cos(a), sin(a) for the positive angle; cos(-a), sin(-a) with -a = pi/2 otherwise
  if (angle_deg > 0) {
    rdx = P.cos90 * outx + P.sin90 * perpx;
    rdy = P.cos90 * outy + P.sin90 * perpy;
  } else {
    rdx = P.cos90 * outx + P.sin90 * (-perpx);
    rdy = P.cos90 * outy + P.sin90 * (-perpy);
  }
  normalize2(rdx, rdy);
  const double resolution = (double)g.res;
  double step = (double)g.res * 0.5;
  if (step < 0.01) step = 0.01;
  const double abs_max = sqrt(P.gw * P.gw + P.gh * P.gh) * 3.0;
  double cur = min_distance;
  while (cur <= abs_max) {
    double px = sp.x + rdx * cur, py = sp.y + rdy * cur;
    if (!(px >= P.b.minx && px <= P.b.maxx && py >= P.b.miny && py <= P.b.maxy))
      return make_double2(fmax(P.b.minx, fmin(P.b.maxx, px)), fmax(P.b.miny, fmin(P.b.maxy, py)));
    double fx = (px - g.ox) / resolution, fy = (py - g.oy) / resolution;
    int mx = (int)fx, my = (int)fy;
    if (mx >= 0 && mx < g.w && my >= 0 && my < g.h && grid_occ(g, mx, my)) return make_double2(px, py);
    cur += step;
  }
  double fx = sp.x + rdx * abs_max, fy = sp.y + rdy * abs_max;
  if (!(fx >= P.b.minx && fx <= P.b.maxx && fy >= P.b.miny && fy <= P.b.maxy)) {
    fx = fmax(P.b.minx, fmin(P.b.maxx, fx));
    fy = fmax(P.b.miny, fmin(P.b.maxy, fy));
  }
  return make_double2(fx, fy);
}

__device__ __forceinline__ void warp_argmin(double &d, int &i) {
  for (int o = 16; o > 0; o >>= 1) {
    double od = __shfl_xor_sync(0xffffffffu, d, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (oi >= 0 && (i < 0 || od < d || (od == d && oi < i))) {
      d = od;
      i = oi;
    }
  }
}

// rows: start x,y,end x,y per row (message order); corners out: TL,TR,BL,BR x,y per row.
// The radius ladder 5, 7, 9, 2*diagonal (gvd:716-722) returns the nearest admissible node of the first
// non-empty rung, which is the nearest admissible node overall: searched here as a 9 m block of 0.5 m cells
// first and the whole node array if that block has none.
struct CornerJob {
  double2 endpoint, other;
  double outx, outy, perpx, perpy, angle;
  bool neg, pos;
};

__device__ __forceinline__ CornerJob corner_job(const double *__restrict__ rows, int job) {
  const int r = job >> 2, k = job & 3;
  double2 s = make_double2(rows[4 * r], rows[4 * r + 1]), e = make_double2(rows[4 * r + 2], rows[4 * r + 3]);
  if (s.x > e.x) {  // gvd:140-145
    double2 t = s;
    s = e;
    e = t;
  }
  CornerJob J;
  J.endpoint = k < 2 ? s : e;
  J.other = k < 2 ? e : s;
  J.angle = (k & 1) ? 90.0 : -90.0;  // TL -90, TR +90, BL -90, BR +90
  double mx = J.other.x - J.endpoint.x, my = J.other.y - J.endpoint.y;
  if (sqrt(mx * mx + my * my) < 1e-6) {
    mx = 1.0;
    my = 0.0;
  } else {
    normalize2(mx, my);
  }
  J.outx = -mx, J.outy = -my, J.perpx = -my, J.perpy = mx;
  J.neg = fabs(J.angle - (-90.0)) < 1e-6;
  J.pos = fabs(J.angle - 90.0) < 1e-6;
  return J;
}

// fb[0] counts the corners without an admissible node within 9 m, fb[1 ...] lists them: corner_far_kernel finishes those
// (a whole-array scan by one warp took as long as all other corners together)
__global__ void corner_kernel(const double *__restrict__ rows, CornerParams P, const double2 *__restrict__ nodes, PointGrid g,
                              double *__restrict__ corners, int *__restrict__ fb) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const double min_distance = 0.5;
  const double last_radius = sqrt(P.gw * P.gw + P.gh * P.gh) * 2.0;
  for (int job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; job < 4 * P.n_rows; job += warps) {
    const int r = job >> 2, k = job & 3;
    const CornerJob J = corner_job(rows, job);
    double best = DBL_MAX;
    int best_i = -1;
    // block of cells covering radius 9 m around the endpoint (rungs 5, 7, 9 apply even on grids whose diagonal is shorter),
    // searched in growing square shells of the 0.5 m grid: a node in a cell k or more cells away (in either axis) from
    // the endpoint's cell is farther than (k - 1) cells, so once the best admissible node is nearer than that the shells
    // still to come cannot beat it (nor tie with it) -- usually after 3-4 m instead of the whole 18 m x 18 m block
    {
      const double R = 9.0;
      const double cs = 1.0 / g.inv;  // cell size
      const long long cx0 = cell_coord(J.endpoint.x - R, g.inv), cx1 = cell_coord(J.endpoint.x + R, g.inv);
      const long long cy0 = cell_coord(J.endpoint.y - R, g.inv), cy1 = cell_coord(J.endpoint.y + R, g.inv);
      const long long ex = cell_coord(J.endpoint.x, g.inv), ey = cell_coord(J.endpoint.y, g.inv);
      const int kmax = (int)max(max(ex - cx0, cx1 - ex), max(ey - cy0, cy1 - ey));
      int k_done = -1;  // shells 0 .. k_done are searched
      while (k_done < kmax) {
        const int k_lo = k_done + 1, k_hi = min(kmax, k_done + 4);  // four shells per pass
        // cells of the square of half-width k_hi that are not in the square of half-width k_done
        const int side = 2 * k_hi + 1;
        for (int c = lane; c < side * side; c += 32) {
          const int ox = c % side - k_hi, oy = c / side - k_hi;
          if (max(abs(ox), abs(oy)) < k_lo) continue;
          const long long gx = ex + ox, gy = ey + oy;
          if (gx < cx0 || gx > cx1 || gy < cy0 || gy > cy1) continue;
          int slot = hash_find(g.h, cell_key(gx, gy));
          if (slot < 0) continue;
          for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
            double d = corner_candidate(nodes[u], J.endpoint, J.outx, J.outy, J.perpx, J.perpy, J.neg, J.pos, min_distance, R);
            if (d >= 0.0 && (d < best || (d == best && u < best_i))) {
              best = d;
              best_i = u;
            }
          }
        }
        warp_argmin(best, best_i);
        best = __shfl_sync(0xffffffffu, best, 0);
        best_i = __shfl_sync(0xffffffffu, best_i, 0);
        k_done = k_hi;
        if (best_i >= 0 && best < (double)(k_done) * cs * 0.999) break;
      }
    }
    if (lane != 0) continue;
    if (best_i < 0 && last_radius > 9.0) {  // last rung of the ladder (2 x diagonal): all nodes, by a whole CTA
      fb[1 + atomicAdd(fb, 1)] = job;
      continue;
    }
    double2 c = best_i >= 0 ? nodes[best_i] : cast_ray_dev(P, J.endpoint, J.other, J.angle, min_distance);
    corners[8 * (size_t)r + 2 * k] = c.x;
    corners[8 * (size_t)r + 2 * k + 1] = c.y;
  }
}

constexpr int kFarThreads = 1024;
__global__ void __launch_bounds__(kFarThreads) corner_far_kernel(const double *__restrict__ rows, CornerParams P,
                                                                 const double2 *__restrict__ nodes, double *__restrict__ corners,
                                                                 const int *__restrict__ fb) {
  __shared__ double s_d[kFarThreads / 32];
  __shared__ int s_i[kFarThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double min_distance = 0.5;
  const double last_radius = sqrt(P.gw * P.gw + P.gh * P.gh) * 2.0;
  const int n_jobs = fb[0];
  for (int q = blockIdx.x; q < n_jobs; q += gridDim.x) {
    const int job = fb[1 + q];
    const int r = job >> 2, k = job & 3;
    const CornerJob J = corner_job(rows, job);
    double best = DBL_MAX;
    int best_i = -1;
    for (int u = threadIdx.x; u < P.n_nodes; u += kFarThreads) {
      double d = corner_candidate(nodes[u], J.endpoint, J.outx, J.outy, J.perpx, J.perpy, J.neg, J.pos, min_distance, last_radius);
      if (d >= 0.0 && (d < best || (d == best && u < best_i))) {
        best = d;
        best_i = u;
      }
    }
    warp_argmin(best, best_i);
    __syncthreads();  // the previous job's readers are done with the scratch
    if (lane == 0) {
      s_d[warp] = best;
      s_i[warp] = best_i;
    }
    __syncthreads();
    if (warp == 0) {
      best = s_d[lane];
      best_i = s_i[lane];
      warp_argmin(best, best_i);
      if (lane == 0) {
        double2 c = best_i >= 0 ? nodes[best_i] : cast_ray_dev(P, J.endpoint, J.other, J.angle, min_distance);
        corners[8 * (size_t)r + 2 * k] = c.x;
        corners[8 * (size_t)r + 2 * k + 1] = c.y;
      }
    }
  }
}

// publishGraph labels (gvd:925-985): node within 0.1 m of a corner point gets its bit; entries ordered by
// (node, row, corner).  PASS 0 counts, PASS 1 fills codes (row * 4 + corner) per node.
template <int PASS>
__global__ void label_scatter_kernel(const double *__restrict__ corners, int n_rows, const double2 *__restrict__ nodes,
                                     PointGrid g, uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                                     uint32_t *__restrict__ cursor, int *__restrict__ codes) {
  const double tol = 0.1;
  for (int job = blockIdx.x * blockDim.x + threadIdx.x; job < 4 * n_rows; job += gridDim.x * blockDim.x) {
    const double2 c = make_double2(corners[2 * (size_t)job], corners[2 * (size_t)job + 1]);
    const long long cx = cell_coord(c.x, g.inv), cy = cell_coord(c.y, g.inv);
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        int slot = hash_find(g.h, cell_key(cx + ox, cy + oy));
        if (slot < 0) continue;
        for (int u = g.h.val[slot]; u >= 0; u = g.next[u]) {
          double dx = nodes[u].x - c.x, dy = nodes[u].y - c.y;
          if (!(sqrt(dx * dx + dy * dy) < tol)) continue;
          if (PASS == 0) atomicAdd(&counts[u], 1u);
          else codes[offsets[u] + atomicAdd(&cursor[u], 1u)] = job;
        }
      }
  }
}

__global__ void label_finish_kernel(int n_nodes, const uint32_t *__restrict__ offsets, int *__restrict__ codes,
                                    int32_t *__restrict__ labels, int32_t *__restrict__ cluster_idx, int32_t *__restrict__ label_counts,
                                    int32_t *__restrict__ label_clusters, int32_t *__restrict__ label_types) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
    const uint32_t b = offsets[i], cnt = offsets[i + 1] - b;
    int *seg = codes + b;
    for (uint32_t a = 1; a < cnt; ++a) {
      int key = seg[a];
      int j = (int)a - 1;
      while (j >= 0 && seg[j] > key) {
        seg[j + 1] = seg[j];
        --j;
      }
      seg[j + 1] = key;
    }
    int mask = 0;
    for (uint32_t a = 0; a < cnt; ++a) {
      mask |= 1 << (seg[a] & 3);
      label_clusters[b + a] = seg[a] >> 2;
      label_types[b + a] = seg[a] & 3;
    }
    labels[i] = mask;
    cluster_idx[i] = cnt ? (seg[0] >> 2) : -1;
    label_counts[i] = (int32_t)cnt;
  }
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
namespace {

struct Arena {
  char *base = nullptr;
  size_t off = 0, cap = 0;
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(base + off);
    off += sizeof(T) * n;
    return p;
  }
};

inline unsigned pow2_at_least(size_t n) {
  unsigned c = 64;
  while (c < n) c <<= 1;
  return c;
}
inline int blocks_for(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  b = std::min<size_t>(std::max<size_t>(b, 1), (size_t)kNumSMs * 16);
  return (int)b;
}

}  // namespace

aos_status run_graph(Ctx *c, const GraphInputs &in) {
  cudaStream_t st = c->stream;
  GraphHost &G = c->graph;
  G.clear();
  G.resolution = in.res;
  G.origin_x = in.ox;
  G.origin_y = in.oy;
  G.n_rows = in.n_rows;
  const int K = in.n_slots;  // facet-vertex slots == Voronoi edges (vd:97-114)
  G.n_voronoi_edges = K;
  const int n_rows = in.n_rows;
  // bounds, gvd:278-281 / :422-431: origin + float(width * resolution)
  const double gw = (double)(float)((float)(unsigned)in.w * in.res), gh = (double)(float)((float)(unsigned)in.h * in.res);
  const Bounds B{in.ox, in.ox + gw, in.oy, in.oy + gh};
  GridView gv{in.skel_bits, in.w, in.h, in.pitch, in.ox, in.oy, in.res};
  if (K == 0) {
    if (!G.corner_points.resize((size_t)8 * n_rows)) return AOS_ERR_CUDA;
    memset(G.corner_points.data(), 0, sizeof(double) * 8 * (size_t)n_rows);
    return AOS_OK;
  }

  // ---- device memory (one arena, sized for the worst case M == K) ---------------------------------
  const unsigned cap_pts = pow2_at_least((size_t)K * 2), cap_edges = pow2_at_least((size_t)K * 2);
  const int ovf_cap = 4096;
  size_t need = 0;
  auto add = [&](size_t bytes) { need += ((bytes + 255) & ~(size_t)255) + 256; };
  add(sizeof(float2) * K);                       // facet xy
  add(sizeof(int) * K);                          // enext
  add(sizeof(double2) * K);                      // pts
  add(K);                                        // state
  add(sizeof(unsigned long long) * cap_pts);     // grid5 keys
  add(sizeof(int) * cap_pts);                    // grid5 heads
  add(sizeof(int) * K);                          // grid5 next
  add(sizeof(int) * (ovf_cap + 64));             // overflow list + counters
  add(sizeof(uint32_t) * (K + 1));               // scan buffer A
  add(sizeof(uint32_t) * (K + 1));               // scan buffer B
  add(sizeof(double2) * K);                      // nodes (pre-crop)
  add(sizeof(int) * K);                          // nn
  add(sizeof(unsigned long long) * cap_edges);   // edge hash keys
  add(sizeof(int) * cap_edges);                  // edge hash first index
  add(sizeof(int) * cap_edges);                  // edge hash accepted
  add(sizeof(unsigned long long) * cap_pts);     // node grid keys
  add(sizeof(int) * cap_pts);                    // node grid heads
  add(sizeof(int) * K);                          // node grid next
  add(sizeof(double2) * K);                      // cropped nodes
  add(sizeof(double) * 3 * K);                   // nodes xyz
  add(sizeof(int32_t) * 3 * K);                  // labels, cluster idx, counts
  add(sizeof(double) * 8 * (n_rows + 1));        // corners
  add(sizeof(double) * 4 * (n_rows + 1));        // rows
  add(sizeof(uint32_t) * (K + 1) * 2);           // label offsets, cursor
  add(sizeof(int) * 16 * (n_rows + 1) * 3);      // label codes / clusters / types (upper bound grows below)
  AOS_CUDA_OK(c, c->gvd_buf.reserve(need));
  Arena A{c->gvd_buf.as<char>(), 0, c->gvd_buf.cap};
  float2 *d_fxy = A.take<float2>(K);
  int *d_enext = A.take<int>(K);
  double2 *d_pts = A.take<double2>(K);
  unsigned char *d_state = A.take<unsigned char>(K);
  PointGrid g5;
  g5.h.keys = A.take<unsigned long long>(cap_pts);
  g5.h.val = A.take<int>(cap_pts);
  g5.h.mask = cap_pts - 1;
  g5.next = A.take<int>(K);
  g5.inv = 1.0 / (0.05 * (1.0 + 1e-9));
  int *d_ovf = A.take<int>(ovf_cap + 64);
  int *d_cnt = d_ovf + ovf_cap;  // [0] overflow count, [1] pending flag
  uint32_t *d_scanA = A.take<uint32_t>(K + 1);
  uint32_t *d_scanB = A.take<uint32_t>(K + 1);
  double2 *d_nodes = A.take<double2>(K);
  int *d_nn = A.take<int>(K);
  DevHash H;
  H.keys = A.take<unsigned long long>(cap_edges);
  H.val = A.take<int>(cap_edges);
  H.mask = cap_edges - 1;
  int *d_acc = A.take<int>(cap_edges);
  PointGrid gn;
  gn.h.keys = A.take<unsigned long long>(cap_pts);
  gn.h.val = A.take<int>(cap_pts);
  gn.h.mask = cap_pts - 1;
  gn.next = A.take<int>(K);
  gn.inv = 1.0 / (0.5 * (1.0 + 1e-9));
  double2 *d_cnodes = A.take<double2>(K);
  double *d_xyz = A.take<double>(3 * (size_t)K);
  int32_t *d_labels = A.take<int32_t>(K), *d_cidx = A.take<int32_t>(K), *d_lcnt = A.take<int32_t>(K);
  double *d_corners = A.take<double>(8 * (size_t)(n_rows + 1));
  double *d_rows = A.take<double>(4 * (size_t)(n_rows + 1));
  uint32_t *d_loff = A.take<uint32_t>(K + 1), *d_lcur = A.take<uint32_t>(K + 1);
  uint32_t *d_tot = reinterpret_cast<uint32_t *>(d_cnt + 8);  // scan totals

  if (in.device_facets) {
    aos_status fs = in.device_cells ? vcells_fill(c, d_fxy, d_enext) : facets_fill(c, d_fxy, d_enext);
    if (fs != AOS_OK) return fs;
  } else {
    AOS_CUDA_OK(c, cudaMemcpyAsync(d_fxy, in.facet_xy, sizeof(float2) * K, cudaMemcpyHostToDevice, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(d_enext, in.enext, sizeof(int) * K, cudaMemcpyHostToDevice, st));
  }
  if (n_rows) {
    aos_status hs = h2d_small(c, d_rows, in.rows_info, sizeof(double) * 4 * n_rows, true);  // pin_rows
    if (hs != AOS_OK) return hs;
  }
  AOS_CUDA_OK(c, cudaMemsetAsync(d_state, 0, K, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(g5.h.keys, 0xff, sizeof(unsigned long long) * cap_pts, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(g5.h.val, 0xff, sizeof(int) * cap_pts, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(d_cnt, 0, sizeof(int) * 64, st));
  c->mark("gvd_upload");

  // ---- extractBoundaryPoints ------------------------------------------------------------------------
  facet_points_kernel<<<blocks_for(K), 256, 0, st>>>(d_fxy, K, d_pts, d_ovf, d_cnt, ovf_cap);
  ++c->launches;
  grid_build_kernel<<<blocks_for(K), 256, 0, st>>>(d_pts, K, g5);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_cnt, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int n_ovf = c->h_flag[0];
  if (n_ovf > ovf_cap) {
    set_error(c, "more than 4096 Voronoi vertices beyond +-2.1e7 m (integer-key overflow list full)");
    return AOS_ERR_CAPACITY;
  }
  constexpr int kBoundaryBatch = 6;  // rounds per host synchronisation (a round exits at once when its predecessor was the last)
  int *d_rflags = d_cnt + 32;
  for (int batch = 0; batch < 100000; ++batch) {
    AOS_CUDA_OK(c, cudaMemsetAsync(d_rflags, 0, sizeof(int) * kBoundaryBatch, st));
    for (int j = 0; j < kBoundaryBatch; ++j)
      boundary_round_kernel<<<blocks_for(K), 256, 0, st>>>(d_pts, K, g5, d_state, d_ovf, n_ovf, j ? d_rflags + j - 1 : nullptr,
                                                           d_rflags + j);
    c->launches += kBoundaryBatch;
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_rflags, sizeof(int) * kBoundaryBatch, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    bool done = false;
    for (int j = 0; j < kBoundaryBatch; ++j) done |= c->h_flag[j] == 0;
    if (done) break;
  }
  flags_from_state_kernel<<<blocks_for(K), 256, 0, st>>>(d_state, K, d_scanA);
  ++c->launches;
  aos_status s = exclusive_scan_u32(c, d_scanA, (size_t)K, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  compact_nodes_kernel<<<blocks_for(K), 256, 0, st>>>(d_pts, d_state, d_scanA, K, d_nodes);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int M = c->h_flag[0];
  G.n_boundary_points = M;
  c->mark("gvd_boundary_points");

  // ---- buildGraphFromBoundaryPoints: Voronoi edges ---------------------------------------------------
  nearest_node_kernel<<<blocks_for(K), 256, 0, st>>>(d_pts, d_state, d_scanA, K, g5, d_nodes, M, d_nn);
  ++c->launches;
  AOS_CUDA_OK(c, cudaMemsetAsync(H.keys, 0xff, sizeof(unsigned long long) * cap_edges, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(H.val, 0x7f, sizeof(int) * cap_edges, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(d_acc, 0, sizeof(int) * cap_edges, st));
  edge_key_kernel<<<blocks_for(K), 256, 0, st>>>(d_nn, d_enext, K, H);
  ++c->launches;
  uint32_t *d_keep = d_scanB;  // K entries
  {
    // list of first-occurrence slots: int4 per entry, at most K; lives in gvd_buf3 until the labels need that buffer
    AOS_CUDA_OK(c, c->gvd_buf3.reserve(sizeof(int4) * ((size_t)K + 1) + 64));
    int4 *d_list = c->gvd_buf3.as<int4>();
    int *d_nlist = d_cnt + 40;
    AOS_CUDA_OK(c, cudaMemsetAsync(d_nlist, 0, sizeof(int), st));
    edge_first_kernel<<<blocks_for(K), 256, 0, st>>>(d_nn, d_enext, K, H, d_keep, d_list, d_nlist);
    edge_test_kernel<<<blocks_for((size_t)K * 32 / 3), 256, 0, st>>>(d_list, d_nlist, d_acc, d_nodes, gv, d_keep);
    c->launches += 2;
  }
  AOS_CUDA_OK(c, cudaGetLastError());
  // positions of the kept edges: scan a copy (keep flags are still needed by the emit kernel)
  uint32_t *d_pos = reinterpret_cast<uint32_t *>(d_pts);  // pts are no longer needed after nearest_node: reuse (K * 16 B)
  AOS_CUDA_OK(c, cudaMemcpyAsync(d_pos, d_keep, sizeof(uint32_t) * K, cudaMemcpyDeviceToDevice, st));
  s = exclusive_scan_u32(c, d_pos, (size_t)K, c->cc_blocksum, d_tot + 1);
  if (s != AOS_OK) return s;

  // ---- proximity edges ---------------------------------------------------------------------------------
  AOS_CUDA_OK(c, cudaMemsetAsync(gn.h.keys, 0xff, sizeof(unsigned long long) * cap_pts, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(gn.h.val, 0xff, sizeof(int) * cap_pts, st));
  uint32_t *d_pcount = d_pos + K;  // second quarter of the old pts buffer: M + 1 entries
  if (M > 0) {
    grid_build_kernel<<<blocks_for(M), 256, 0, st>>>(d_nodes, M, gn);
  ++c->launches;
    proximity_pairs_kernel<false><<<blocks_for(M), 256, 0, st>>>(d_nodes, M, gn, H, d_acc, d_pcount, nullptr, nullptr);
  ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
  }
  AOS_CUDA_OK(c, cudaMemsetAsync(d_pcount + M, 0, 4, st));
  s = exclusive_scan_u32(c, d_pcount, (size_t)M + 1, c->cc_blocksum, d_tot + 2);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 16, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  const int n_vor = c->h_flag[1], n_pairs = c->h_flag[2];
  c->mark("gvd_voronoi_edges");

  // pair buffers + records (second arena: sizes are only known now)
  size_t need2 = sizeof(int) * 2 * ((size_t)n_pairs + 64) + sizeof(uint32_t) * 2 * ((size_t)n_pairs + 64) +
                 sizeof(EdgeRec) * ((size_t)n_vor + n_pairs + 64) + sizeof(uint32_t) * 2 * ((size_t)n_vor + n_pairs + 64) +
                 (sizeof(int32_t) * 2 + sizeof(float) * 2) * ((size_t)n_vor + n_pairs + 64) + 4096;
  AOS_CUDA_OK(c, c->gvd_buf2.reserve(need2));
  Arena A2{c->gvd_buf2.as<char>(), 0, c->gvd_buf2.cap};
  int *d_pi = A2.take<int>((size_t)n_pairs + 1), *d_pj = A2.take<int>((size_t)n_pairs + 1);
  uint32_t *d_pkeep = A2.take<uint32_t>((size_t)n_pairs + 1), *d_ppos = A2.take<uint32_t>((size_t)n_pairs + 1);
  EdgeRec *d_rec = A2.take<EdgeRec>((size_t)n_vor + n_pairs + 1);
  uint32_t *d_eflag = A2.take<uint32_t>((size_t)n_vor + n_pairs + 1);
  int32_t *d_edges = A2.take<int32_t>(2 * ((size_t)n_vor + n_pairs + 1));
  float *d_len = A2.take<float>((size_t)n_vor + n_pairs + 1), *d_clr = A2.take<float>((size_t)n_vor + n_pairs + 1);

  edge_emit_kernel<<<blocks_for(K), 256, 0, st>>>(d_nn, d_enext, K, d_keep, d_pos, d_rec);
  ++c->launches;
  int n_prox = 0;
  if (n_pairs > 0) {
    proximity_pairs_kernel<true><<<blocks_for(M), 256, 0, st>>>(d_nodes, M, gn, H, d_acc, nullptr, d_pcount, d_pj);
  ++c->launches;
    pair_owner_kernel<<<blocks_for(M), 256, 0, st>>>(d_pcount, M, d_pi);
  ++c->launches;
    pair_test_kernel<<<blocks_for((size_t)n_pairs * 32), 256, 0, st>>>(d_pi, d_pj, n_pairs, d_nodes, gv, d_pkeep);
  ++c->launches;
    AOS_CUDA_OK(c, cudaMemcpyAsync(d_ppos, d_pkeep, sizeof(uint32_t) * n_pairs, cudaMemcpyDeviceToDevice, st));
    s = exclusive_scan_u32(c, d_ppos, (size_t)n_pairs, c->cc_blocksum, d_tot + 3);
    if (s != AOS_OK) return s;
    pair_emit_kernel<<<blocks_for(n_pairs), 256, 0, st>>>(d_pi, d_pj, n_pairs, d_pkeep, d_ppos, d_rec + n_vor);
  ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot + 3, 4, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    n_prox = c->h_flag[0];
  }
  const int n_rec = n_vor + n_prox;
  c->mark("gvd_proximity_edges");

  // ---- filterNodesAndEdgesOutsideGrid --------------------------------------------------------------
  uint32_t *d_nremap = d_scanA;  // rank array no longer needed
  int N = 0, NE = 0;
  if (M > 0) {
    crop_flag_kernel<<<blocks_for(M), 256, 0, st>>>(d_nodes, M, B, d_nremap);
  ++c->launches;
    s = exclusive_scan_u32(c, d_nremap, (size_t)M, c->cc_blocksum, d_tot + 4);
    if (s != AOS_OK) return s;
    crop_nodes_kernel<<<blocks_for(M), 256, 0, st>>>(d_nodes, M, B, d_nremap, d_cnodes, d_xyz);
  ++c->launches;
    if (n_rec > 0) {
      crop_edge_flag_kernel<<<blocks_for(n_rec), 256, 0, st>>>(d_rec, n_rec, d_nodes, B, d_eflag);
  ++c->launches;
      s = exclusive_scan_u32(c, d_eflag, (size_t)n_rec, c->cc_blocksum, d_tot + 5);
      if (s != AOS_OK) return s;
      crop_edges_kernel<<<blocks_for(n_rec), 256, 0, st>>>(d_rec, n_rec, d_nodes, B, d_nremap, d_eflag, d_edges, d_len, d_clr);
  ++c->launches;
    }
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot + 4, 8, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    N = c->h_flag[0];
    NE = n_rec > 0 ? c->h_flag[1] : 0;
  }
  if (c->clearance && NE > 0) {  // opt-in: exact EDT of the framed skeleton, then one warp per published edge
    const size_t cells = (size_t)in.w * in.h;
    AOS_CUDA_OK(c, c->edt_out.reserve(cells * 8 + 1024));
    uint32_t *d_near = c->edt_out.as<uint32_t>();
    int32_t *d_d2 = reinterpret_cast<int32_t *>(d_near + cells);
    s = launch_edt(c, in.skel_bits, in.w, in.h, d_near, d_d2);
    if (s != AOS_OK) return s;
    edge_clearance_kernel<<<blocks_for((size_t)NE * 32), 256, 0, st>>>(d_edges, NE, d_cnodes, gv, d_d2, d_clr);
    ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
  }
  c->mark("gvd_crop");
  // node coordinates and edges are final here: their way back to the host (18 of the 23 MB of a config-3 graph) starts now,
  // on a side stream, beside the corner search and the labels
  bool early_d2h = false;
  if (!G.nodes_xyz.resize(3 * (size_t)N) || !G.edges.resize(2 * (size_t)NE) || !G.edge_lengths.resize(NE) ||
      !G.edge_clearances.resize(NE)) {
    set_error(c, "cudaHostAlloc failed for the graph result buffers");
    return AOS_ERR_CUDA;
  }
  if (N > 0) {
    cudaStream_t cs = c->aux[1];
    AOS_CUDA_OK(c, cudaEventRecord(c->ev_fork, st));
    AOS_CUDA_OK(c, cudaStreamWaitEvent(cs, c->ev_fork, 0));
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.nodes_xyz.data(), d_xyz, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, cs));
    if (NE > 0) {
      AOS_CUDA_OK(c, cudaMemcpyAsync(G.edges.data(), d_edges, sizeof(int32_t) * 2 * NE, cudaMemcpyDeviceToHost, cs));
      AOS_CUDA_OK(c, cudaMemcpyAsync(G.edge_lengths.data(), d_len, sizeof(float) * NE, cudaMemcpyDeviceToHost, cs));
      AOS_CUDA_OK(c, cudaMemcpyAsync(G.edge_clearances.data(), d_clr, sizeof(float) * NE, cudaMemcpyDeviceToHost, cs));
    }
    AOS_CUDA_OK(c, cudaEventRecord(c->ev_join[1], cs));
    early_d2h = true;
  }

  // ---- TL/TR/BL/BR corner nodes + labels ---------------------------------------------------------------
  if (!G.corner_points.resize((size_t)8 * n_rows)) return AOS_ERR_CUDA;
  if (n_rows) memset(G.corner_points.data(), 0, sizeof(double) * 8 * (size_t)n_rows);
  int n_label_entries = 0;
  int *d_codes = nullptr, *d_lcl = nullptr, *d_lty = nullptr;
  AOS_CUDA_OK(c, cudaMemsetAsync(d_loff, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
  if (N > 0 && n_rows > 0) {
    // node grid over the cropped nodes (rebuild: indices changed)
    AOS_CUDA_OK(c, cudaMemsetAsync(gn.h.keys, 0xff, sizeof(unsigned long long) * cap_pts, st));
    AOS_CUDA_OK(c, cudaMemsetAsync(gn.h.val, 0xff, sizeof(int) * cap_pts, st));
    grid_build_kernel<<<blocks_for(N), 256, 0, st>>>(d_cnodes, N, gn);
  ++c->launches;
    CornerParams CP;
    CP.g = gv;
    CP.b = B;
    CP.cos90 = cos(90.0 * M_PI / 180.0);
    CP.sin90 = sin(90.0 * M_PI / 180.0);
    CP.gw = gw;
    CP.gh = gh;
    CP.n_rows = n_rows;
    CP.n_nodes = N;
    AOS_CUDA_OK(c, c->corner_fb.reserve(sizeof(int) * (4 * (size_t)n_rows + 4)));
    int *d_fb = c->corner_fb.as<int>();
    AOS_CUDA_OK(c, cudaMemsetAsync(d_fb, 0, sizeof(int), st));
    corner_kernel<<<blocks_for((size_t)n_rows * 4 * 32), 256, 0, st>>>(d_rows, CP, d_cnodes, gn, d_corners, d_fb);
    corner_far_kernel<<<std::min(4 * n_rows, 4 * kNumSMs), kFarThreads, 0, st>>>(d_rows, CP, d_cnodes, d_corners, d_fb);
    c->launches += 2;
    label_scatter_kernel<0><<<blocks_for((size_t)n_rows * 4), 256, 0, st>>>(d_corners, n_rows, d_cnodes, gn, d_loff, nullptr,
                                                                           nullptr, nullptr);
  ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    s = exclusive_scan_u32(c, d_loff, (size_t)N + 1, c->cc_blocksum, d_tot + 6);
    if (s != AOS_OK) return s;
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot + 6, 4, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.corner_points.data(), d_corners, sizeof(double) * 8 * n_rows, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    n_label_entries = c->h_flag[0];
    AOS_CUDA_OK(c, c->gvd_buf3.reserve(sizeof(int) * 3 * ((size_t)n_label_entries + 64)));
    d_codes = c->gvd_buf3.as<int>();
    d_lcl = d_codes + n_label_entries + 16;
    d_lty = d_lcl + n_label_entries + 16;
    AOS_CUDA_OK(c, cudaMemsetAsync(d_lcur, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
    if (n_label_entries > 0) {
      label_scatter_kernel<1><<<blocks_for((size_t)n_rows * 4), 256, 0, st>>>(d_corners, n_rows, d_cnodes, gn, nullptr, d_loff,
                                                                             d_lcur, d_codes);
  ++c->launches;
    }
  } else {
    AOS_CUDA_OK(c, c->gvd_buf3.reserve(1024));
    d_codes = c->gvd_buf3.as<int>();
    d_lcl = d_codes + 16;
    d_lty = d_lcl + 16;
  }
  if (N > 0) {
    label_finish_kernel<<<blocks_for(N), 256, 0, st>>>(N, d_loff, d_codes, d_labels, d_cidx, d_lcnt, d_lcl, d_lty);
  ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
  }
  c->mark("gvd_corners_labels");

  // ---- results to the host (GvdGraph.msg arrays) -------------------------------------------------------
  if (!G.node_labels.resize(N) || !G.node_cluster_indices.resize(N) || !G.node_label_counts.resize(N) ||
      !G.node_label_clusters.resize(n_label_entries) || !G.node_label_types.resize(n_label_entries)) {
    set_error(c, "cudaHostAlloc failed for the graph result buffers");
    return AOS_ERR_CUDA;
  }
  if (N > 0) {
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.node_labels.data(), d_labels, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.node_cluster_indices.data(), d_cidx, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.node_label_counts.data(), d_lcnt, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
  }
  if (n_label_entries > 0) {
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.node_label_clusters.data(), d_lcl, sizeof(int32_t) * n_label_entries, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(G.node_label_types.data(), d_lty, sizeof(int32_t) * n_label_entries, cudaMemcpyDeviceToHost, st));
  }
  if (early_d2h) AOS_CUDA_OK(c, cudaStreamWaitEvent(st, c->ev_join[1], 0));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->mark("gvd_d2h");
  return AOS_OK;
}

}  // namespace aos
