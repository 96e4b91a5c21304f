// k_cluster.cu -- clusterOccupiedCells + the row extraction of convertClustersToTreeRows
// (src/aos_seed_gen_node.cpp:970-1083, 1231-1255, 1258-1306, 1309-1406) on the GPU.
//
//  1. mask   : skeleton bits AND "cell centre inside the exploration polygon" (even-odd ray cast in
//              double, the reference's literal formula), popcount per word.
//  2. scan   : exclusive prefix over the word popcounts -> every skeleton cell gets a compact index
//              in raster order (the reference's discovery order).
//  3. link   : lock-free union-find over compact indices (CAS hooking, larger root under smaller, so a
//              component's root is its raster-first cell == the reference's BFS start == the canonical
//              label "min linear index").  Only the 4 raster-earlier neighbours are hooked.
//  4. rank   : roots numbered in raster order -> cluster ordinal; sizes, exact int64 coordinate sums and
//              8 directional extreme cells by atomics; cells grouped per cluster.
//  5. per-cluster CTA: exact max pairwise d^2 (extreme-point lower bound + bounding-box pruning, then
//              brute force over the survivors), float32 centre, length, cluster_min_length and polygon
//              filters, farthest / opposite-farthest cells (row start / end).
// Order-dependent details of the reference (float32 running sums once they pass 2^24, first-in-BFS-order
// tie breaks) are resolved by an exact BFS-order replay of the flagged clusters (bfs_replay_kernel).
#include <stdlib.h>

#include <algorithm>

#include "aos_common.cuh"

namespace aos {

// ---------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three kernels; n up to 2^31)
// ---------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanBlock = kScanThreads * kScanPer;  // 2048 elements per block

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? wsum[lane] : 0;
    uint32_t wi = warp_incl_scan(w, lane);
    if (lane < kScanThreads / 32) wsum[lane] = wi - w;  // exclusive warp offsets
    if (lane == kScanThreads / 32 - 1) *total = wi;
  }
  __syncthreads();
  uint32_t r = inc - v + wsum[warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ in, size_t n,
                                                                   uint32_t *__restrict__ blocksum) {
  __shared__ uint32_t tot;
  size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPer;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k)
    if (base + k < n) s += in[base + k];
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_blocksums_kernel(uint32_t *blocksum, int nb, uint32_t *total_out) {
  __shared__ uint32_t tot;
  uint32_t carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    int i = base + threadIdx.x;
    uint32_t v = i < nb ? blocksum[i] : 0;
    uint32_t ex = block_excl_scan(v, &tot);
    if (i < nb) blocksum[i] = ex + carry;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *__restrict__ data, size_t n,
                                                                  const uint32_t *__restrict__ blocksum) {
  __shared__ uint32_t tot;
  size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPer;
  uint32_t v[kScanPer], s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    v[k] = base + k < n ? data[base + k] : 0;
    s += v[k];
  }
  uint32_t ex = block_excl_scan(s, &tot) + blocksum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    if (base + k < n) data[base + k] = ex;
    ex += v[k];
  }
}

// in-place exclusive scan; total written to d_total (device)
aos_status exclusive_scan_u32(Ctx *c, uint32_t *data, size_t n, DevBuf &blocksum_buf, uint32_t *d_total) {
  int nb = (int)((n + kScanBlock - 1) / kScanBlock);
  if (nb < 1) nb = 1;
  AOS_CUDA_OK(c, blocksum_buf.reserve(sizeof(uint32_t) * (size_t)nb));
  uint32_t *bs = blocksum_buf.as<uint32_t>();
  scan_reduce_kernel<<<nb, kScanThreads, 0, c->stream>>>(data, n, bs);
  ++c->launches;
  scan_blocksums_kernel<<<1, kScanThreads, 0, c->stream>>>(bs, nb, d_total);
  ++c->launches;
  scan_apply_kernel<<<nb, kScanThreads, 0, c->stream>>>(data, n, bs);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

// ---------------------------------------------------------------------------------------------------
// 1. polygon mask + popcount
// ---------------------------------------------------------------------------------------------------
// isPointInPolygon, seed_gen:1231-1255 (double, |dy| > 1e-9 guard); compiled with -fmad=false
__device__ __forceinline__ bool point_in_polygon(const SeedDeviceParams &P, double px, double py) {
  if (P.n_poly < 3) return false;
  bool inside = false;
  int j = P.n_poly - 1;
  for (int i = 0; i < P.n_poly; ++i) {
    double pix = P.poly[2 * i], piy = P.poly[2 * i + 1];
    double pjx = P.poly[2 * j], pjy = P.poly[2 * j + 1];
    double dy = pjy - piy;
    if (fabs(dy) > 1e-9) {
      if (((piy > py) != (pjy > py)) && (px < (pjx - pix) * (py - piy) / dy + pix)) inside = !inside;
    }
    j = i;
  }
  return inside;
}

// world coordinate of a cell as the reference computes it (seed_gen:998-999, 1349-1350):
// float(double(origin) + double(float(idx) * res_f32))
__device__ __forceinline__ float cell_world(double origin, int idx, float res) {
  return (float)(origin + (double)__fmul_rn((float)idx, res));
}

// The polygon test (seed_gen:979-982) runs only on set bits, and set bits are rare (a skeleton): a warp pools the set bits
// of its 32 words and tests them one per lane, instead of every lane walking its own word's bits while the others wait.
__global__ void __launch_bounds__(256) mask_count_kernel(const __grid_constant__ SeedDeviceParams P,
                                                         const uint32_t *__restrict__ skel, uint32_t *__restrict__ mask,
                                                         uint32_t *__restrict__ counts) {
  __shared__ uint32_t s_keep[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t total = (size_t)P.pitch * P.h;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t base = (size_t)blockIdx.x * blockDim.x + warp * 32; base < total; base += stride) {
    const size_t i = base + lane;
    uint32_t v = i < total ? skel[i] : 0u;
    if (P.n_poly > 0 && __any_sync(0xffffffffu, v != 0u)) {  // use_polygon_filter
      const uint32_t cnt = __popc(v);
      const uint32_t incl = warp_incl_scan(cnt, lane);
      const uint32_t all = __shfl_sync(0xffffffffu, incl, 31);
      s_keep[warp][lane] = 0u;
      __syncwarp();
      for (uint32_t t0 = 0; t0 < all; t0 += 32) {
        const uint32_t t = t0 + lane;
        // owner = first lane whose inclusive count exceeds t
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const uint32_t probe = __shfl_sync(0xffffffffu, incl, (owner + step - 1) & 31);
          if (owner + step <= 31 && probe <= t) owner += step;
        }
        const uint32_t ov = __shfl_sync(0xffffffffu, v, owner), oincl = __shfl_sync(0xffffffffu, incl, owner);
        if (t < all) {
          const uint32_t nth = t - (oincl - __popc(ov));  // 0-based rank of the bit inside the owner's word
          const int b = __fns(ov, 0, (int)nth + 1);
          const size_t wi = base + owner;
          const int y = (int)(wi / (size_t)P.pitch), cw = (int)(wi - (size_t)y * P.pitch);
          const float wy = cell_world(P.oy, y, P.res), wx = cell_world(P.ox, (cw << 5) + b, P.res);
          if (point_in_polygon(P, (double)wx, (double)wy)) atomicOr(&s_keep[warp][owner], 1u << b);
        }
      }
      __syncwarp();
      v = s_keep[warp][lane];
      __syncwarp();
    }
    if (i < total) {
      mask[i] = v;
      counts[i] = __popc(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 3. union-find
// ---------------------------------------------------------------------------------------------------
__global__ void cc_init_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ prefix, int pitch, int h,
                               int w, int *__restrict__ parent, int *__restrict__ cellpos) {
  size_t total = (size_t)pitch * h;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    uint32_t v = mask[i];
    if (!v) continue;
    int y = (int)(i / (size_t)pitch), cw = (int)(i - (size_t)y * pitch);
    const int idx0 = (int)prefix[i];
    int idx = idx0;
    // consecutive cells of a word start out under the first cell of their run (its index is the smaller one, as the
    // hooking rule wants): the west links inside a word, most of all links on a thinned row, are then already made
    const uint32_t v0 = v, starts = v & ~(v << 1);
    while (v) {
      int b = __ffs(v) - 1;
      v &= v - 1;
      const int sb = 31 - __clz(starts & (0xffffffffu >> (31 - b)));  // first bit of the run that holds bit b
      parent[idx] = idx0 + __popc(v0 & ((1u << sb) - 1u));
      cellpos[idx] = y * w + (cw << 5) + b;
      ++idx;
    }
  }
}

__device__ __forceinline__ int uf_find(int *parent, int i) {
  int cur = parent[i];
  if (cur != i) {
    int prev = i, next;
    while (cur > (next = parent[cur])) {
      parent[prev] = next;  // pointer jumping; parents only ever decrease, so this is benign
      prev = cur;
      cur = next;
    }
  }
  return cur;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b) {
  int ra = uf_find(parent, a), rb = uf_find(parent, b);
  bool repeat;
  do {
    repeat = false;
    if (ra != rb) {
      int ret;
      if (ra < rb) {
        if ((ret = atomicCAS(&parent[rb], rb, ra)) != rb) {
          rb = ret;
          repeat = true;
        }
      } else {
        if ((ret = atomicCAS(&parent[ra], ra, rb)) != ra) {
          ra = ret;
          repeat = true;
        }
      }
    }
  } while (repeat);
}

__device__ __forceinline__ int compact_index(const uint32_t *mask, const uint32_t *prefix, int pitch, int x, int y) {
  size_t wi = (size_t)y * pitch + (x >> 5);
  uint32_t m = mask[wi];
  uint32_t bit = 1u << (x & 31);
  if (!(m & bit)) return -1;
  return (int)prefix[wi] + __popc(m & (bit - 1u));
}

// one thread per skeleton cell (dense: every lane has a cell), hooking it to the raster-earlier half of its neighbourhood
__global__ void cc_link_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ prefix, int pitch, int w,
                               const int *__restrict__ cellpos, int n, int *parent) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int pos = cellpos[idx];
    const int y = pos / w, x = pos - y * w;
    // W, NW, N, NE  (N = row y-1)
    if (x > 0 && (x & 31) == 0) {
      // west: made by cc_init_kernel inside a word; across a word boundary the neighbour is the previous compact index
      if (mask[(size_t)y * pitch + ((x - 1) >> 5)] >> 31) uf_union(parent, idx, idx - 1);
    }
    if (y > 0) {
      if (x > 0) {
        int n = compact_index(mask, prefix, pitch, x - 1, y - 1);
        if (n >= 0) uf_union(parent, idx, n);
      }
      int n = compact_index(mask, prefix, pitch, x, y - 1);
      if (n >= 0) uf_union(parent, idx, n);
      if (x + 1 < w) {
        n = compact_index(mask, prefix, pitch, x + 1, y - 1);
        if (n >= 0) uf_union(parent, idx, n);
      }
    }
  }
}

__global__ void cc_flatten_kernel(int *parent, uint32_t *__restrict__ is_root, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = i, p;
    while ((p = parent[r]) != r) r = p;
    parent[i] = r;  // racing writers all store a value on the path to the same root
    is_root[i] = (r == i) ? 1u : 0u;
  }
}

// ---------------------------------------------------------------------------------------------------
// 4. per-cluster accumulators
// ---------------------------------------------------------------------------------------------------
struct ClusterAcc {          // one per cluster, zero/identity-initialised
  unsigned int size;
  unsigned int cursor;       // fill cursor for grouping
  unsigned long long sumx, sumy;
  // directional extremes: key = (value + bias) << 32 | compact cell pos ; [0..3] max of x, y, x+y, x-y ; [4..7] min
  unsigned long long ext[8];
};
constexpr long long kExtBias = 1ll << 30;

__global__ void acc_init_kernel(ClusterAcc *acc, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ClusterAcc a;
  a.size = 0;
  a.cursor = 0;
  a.sumx = a.sumy = 0;
  for (int k = 0; k < 4; ++k) a.ext[k] = 0ull;
  for (int k = 4; k < 8; ++k) a.ext[k] = ~0ull;
  acc[i] = a;
}

__global__ void cc_accumulate_kernel(const int *__restrict__ parent, const uint32_t *__restrict__ rootrank,
                                     const int *__restrict__ cellpos, int n, int w, int *__restrict__ cell_cluster,
                                     ClusterAcc *acc) {
  // Compact indices follow the raster, so a warp's 32 cells usually are one horizontal run of one cluster: reduce
  // inside the warp (REDUX) and issue one set of atomics per warp instead of ten atomics per cell.
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += stride) {
    const int i = base + lane;
    const bool valid = i < n;
    int c = -1, pos = 0, x = 0, y = 0;
    if (valid) {
      c = (int)rootrank[parent[i]];
      cell_cluster[i] = c;
      pos = cellpos[i];
      y = pos / w;
      x = pos - y * w;
    }
    int uniform = 0;
    __match_all_sync(0xffffffffu, c, &uniform);
    const int vals[4] = {x, y, x + y, x - y};
    if (uniform && c >= 0) {
      const unsigned sx = __reduce_add_sync(0xffffffffu, (unsigned)x), sy = __reduce_add_sync(0xffffffffu, (unsigned)y);
      unsigned long long kmax[4], kmin[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int vmax = __reduce_max_sync(0xffffffffu, vals[k]), vmin = __reduce_min_sync(0xffffffffu, vals[k]);
        const unsigned pmax = __reduce_max_sync(0xffffffffu, vals[k] == vmax ? (unsigned)pos : 0u);
        const unsigned pmin = __reduce_min_sync(0xffffffffu, vals[k] == vmin ? (unsigned)pos : 0xffffffffu);
        kmax[k] = ((unsigned long long)((long long)vmax + kExtBias) << 32) | pmax;
        kmin[k] = ((unsigned long long)((long long)vmin + kExtBias) << 32) | pmin;
      }
      if (lane == 0) {
        ClusterAcc *a = acc + c;
        atomicAdd(&a->size, 32u);
        atomicAdd(&a->sumx, (unsigned long long)sx);
        atomicAdd(&a->sumy, (unsigned long long)sy);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          atomicMax(&a->ext[k], kmax[k]);
          atomicMin(&a->ext[4 + k], kmin[k]);
        }
      }
    } else if (valid) {
      ClusterAcc *a = acc + c;
      atomicAdd(&a->size, 1u);
      atomicAdd(&a->sumx, (unsigned long long)x);
      atomicAdd(&a->sumy, (unsigned long long)y);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned long long key = ((unsigned long long)((long long)vals[k] + kExtBias) << 32) | (unsigned int)pos;
        atomicMax(&a->ext[k], key);
        atomicMin(&a->ext[4 + k], key);
      }
    }
  }
}

__global__ void acc_sizes_kernel(const ClusterAcc *acc, int n, uint32_t *sizes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sizes[i] = acc[i].size;
}

__global__ void cc_group_kernel(const int *__restrict__ cell_cluster, const int *__restrict__ cellpos, int n,
                                const uint32_t *__restrict__ offsets, ClusterAcc *acc, int *__restrict__ grouped,
                                int *__restrict__ grouped_id) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = cell_cluster[i];
    unsigned int k = atomicAdd(&acc[c].cursor, 1u);
    grouped[offsets[c] + k] = cellpos[i];
    grouped_id[offsets[c] + k] = i;  // the same list as compact indices (bfs_chain_kernel stages a cluster from it)
  }
}

// ---------------------------------------------------------------------------------------------------
// 5. one CTA per cluster
// ---------------------------------------------------------------------------------------------------
constexpr int kClThreads = 128;
constexpr int kMaxCand = 2048;  // pruned diameter candidates kept in shared memory per cluster
enum : int { kFlagNeedsOrder = 1, kFlagRow = 2, kFlagTie = 4 };

struct RowOut {  // device-side mirror of aos_tree_row + book-keeping
  aos_tree_row row;
  int valid;
  int flags;
};

template <typename T>
__device__ __forceinline__ T block_reduce_max(T v, T *scratch) {
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = scratch[0];
  for (int k = 1; k < kClThreads / 32; ++k) r = scratch[k] > r ? scratch[k] : r;
  return r;
}
template <typename T>
__device__ __forceinline__ T block_reduce_min(T v, T *scratch) {
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = scratch[0];
  for (int k = 1; k < kClThreads / 32; ++k) r = scratch[k] < r ? scratch[k] : r;
  return r;
}
__device__ __forceinline__ int block_reduce_sum(int v, int *scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
  for (int k = 0; k < kClThreads / 32; ++k) r += scratch[k];
  return r;
}

__device__ __forceinline__ long long d2ll(int ax, int ay, int bx, int by) {
  long long dx = ax - bx, dy = ay - by;
  return dx * dx + dy * dy;
}

// Exact diameter^2, centre, length, filters, row endpoints.  `cells` = this cluster's cell positions
// (any order).  order_rank (may be null) = BFS rank of each grouped cell when a replay was needed.
__global__ void __launch_bounds__(kClThreads) cluster_finalize_kernel(
    const __grid_constant__ SeedDeviceParams P, const ClusterAcc *__restrict__ acc, const uint32_t *__restrict__ offsets,
    const int *__restrict__ grouped, const int *__restrict__ root_cellpos, float min_length,
    const int *__restrict__ flagged /* null: all clusters; else the replayed ones */,
    const int *__restrict__ skip /* per entry of flagged: 1 = the replay gave up on it, leave its rows as they are */,
    const int *__restrict__ bfs_cells /* BFS-ordered cells of replayed clusters (same offsets) */,
    const float *__restrict__ replay_centre /* 2 per cluster, valid for replayed clusters */,
    aos_cluster *__restrict__ out_clusters, RowOut *__restrict__ out_rows) {
  __shared__ unsigned long long s_u64[kClThreads / 32];
  __shared__ long long s_i64[kClThreads / 32];
  __shared__ int s_int[kClThreads / 32];
  const bool replayed = flagged != nullptr;
  if (skip && skip[blockIdx.x]) return;
  const int c = replayed ? flagged[blockIdx.x] : (int)blockIdx.x;
  const ClusterAcc a = acc[c];
  const int n = (int)a.size;
  // a replayed cluster's cell list is in BFS order, so the list index is the BFS rank
  const int *cells = (replayed ? bfs_cells : grouped) + offsets[c];
  const int w = P.w;

  // ---- exact max pairwise squared distance -------------------------------------------------------
  int ex[8], ey[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int pos = (int)(a.ext[k] & 0xffffffffull);
    ey[k] = pos / w;
    ex[k] = pos - ey[k] * w;
  }
  long long lb = 0;  // lower bound from the 8 directional extreme cells
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i + 1; j < 8; ++j) {
      long long d = d2ll(ex[i], ey[i], ex[j], ey[j]);
      lb = d > lb ? d : lb;
    }
  const int bx0 = ex[4], bx1 = ex[0], by0 = ey[5], by1 = ey[1];  // min x, max x, min y, max y
  // a cell can end a longer pair only if its farthest bounding-box corner is at least lb away
  auto is_cand = [&](int x, int y) -> bool {
    long long dx = max(x - bx0, bx1 - x), dy = max(y - by0, by1 - y);
    return dx * dx + dy * dy >= lb;
  };
  // survivors of the pruning are few (the two tips of a row): gather them once, then all pairs among them
  __shared__ int s_cand[kMaxCand];
  __shared__ int s_ncand;
  if (threadIdx.x == 0) s_ncand = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kClThreads) {
    int pi = cells[i];
    int yi = pi / w, xi = pi - yi * w;
    if (is_cand(xi, yi)) {
      int k = atomicAdd(&s_ncand, 1);
      if (k < kMaxCand) s_cand[k] = pi;
    }
  }
  __syncthreads();
  const int nc = s_ncand;
  long long best = lb;
  if (nc <= kMaxCand) {
    for (int i = threadIdx.x; i < nc; i += kClThreads) {
      int pi = s_cand[i];
      int yi = pi / w, xi = pi - yi * w;
      for (int j = 0; j < nc; ++j) {
        int pj = s_cand[j];
        int yj = pj / w, xj = pj - yj * w;
        long long d = d2ll(xi, yi, xj, yj);
        best = d > best ? d : best;
      }
    }
  } else {  // too many candidates for the shared list (a blob rather than a row): filter on the fly
    for (int i = threadIdx.x; i < n; i += kClThreads) {
      int pi = cells[i];
      int yi = pi / w, xi = pi - yi * w;
      if (!is_cand(xi, yi)) continue;
      for (int j = 0; j < n; ++j) {
        int pj = cells[j];
        int yj = pj / w, xj = pj - yj * w;
        if (!is_cand(xj, yj)) continue;
        long long d = d2ll(xi, yi, xj, yj);
        best = d > best ? d : best;
      }
    }
  }
  const long long maxd2 = block_reduce_max<long long>(best, s_i64);

  // ---- centre (seed_gen:1053-1059) ---------------------------------------------------------------
  // float32 running sums are exact (order-independent) while every partial sum is below 2^24
  int flags = 0;
  float cx, cy;
  const bool exact_sum = a.sumx < (1ull << 24) && a.sumy < (1ull << 24);
  if (replayed) {
    cx = replay_centre[2 * c];
    cy = replay_centre[2 * c + 1];
  } else {
    cx = __fdiv_rn((float)a.sumx, (float)(unsigned long long)n);
    cy = __fdiv_rn((float)a.sumy, (float)(unsigned long long)n);
    if (!exact_sum) flags |= kFlagNeedsOrder;
  }
  // length (seed_gen:1068,1074): float( sqrt(int d2) [double] * res [float->double] )
  float length = maxd2 > 0 ? (float)(sqrt((double)maxd2) * (double)P.res) : 0.0f;

  // ---- filters (seed_gen:1262-1270, 1333-1341) ------------------------------------------------------
  const float centre_wx = (float)(P.ox + (double)__fmul_rn(cx, P.res));
  const float centre_wy = (float)(P.oy + (double)__fmul_rn(cy, P.res));
  bool is_row = length >= min_length;
  if (is_row && P.n_poly > 0) is_row = point_in_polygon(P, (double)centre_wx, (double)centre_wy);

  RowOut ro;
  memset(&ro, 0, sizeof(ro));
  if (is_row) {
    flags |= kFlagRow;
    const double rcx = (double)centre_wx, rcy = (double)centre_wy;
    // -- farthest cell from the centre (strict >, first in BFS order wins ties) --
    unsigned long long bestk = 0;
    for (int i = threadIdx.x; i < n; i += kClThreads) {
      int p = cells[i];
      int y = p / w, x = p - y * w;
      double dx = (double)cell_world(P.ox, x, P.res) - rcx, dy = (double)cell_world(P.oy, y, P.res) - rcy;
      double d2 = dx * dx + dy * dy;
      unsigned long long k = (unsigned long long)__double_as_longlong(d2);
      bestk = k > bestk ? k : bestk;
    }
    const unsigned long long max1 = block_reduce_max<unsigned long long>(bestk, s_u64);
    // winner among ties: lowest BFS rank if known, else lowest raster position (+ tie flag)
    long long pick = 0x7fffffffffffffffll;
    int ties = 0;
    for (int i = threadIdx.x; i < n; i += kClThreads) {
      int p = cells[i];
      int y = p / w, x = p - y * w;
      double dx = (double)cell_world(P.ox, x, P.res) - rcx, dy = (double)cell_world(P.oy, y, P.res) - rcy;
      double d2 = dx * dx + dy * dy;
      if ((unsigned long long)__double_as_longlong(d2) == max1 && max1 != 0ull) {
        ++ties;
        long long key = ((long long)(replayed ? i : 0) << 32) | (unsigned int)p;
        pick = key < pick ? key : pick;
      }
    }
    const int nties1 = block_reduce_sum(ties, s_int);
    const long long pick1 = block_reduce_min<long long>(pick, s_i64);
    int first_pos;
    double fdx = 0.0, fdy = 0.0;
    if (max1 == 0ull) {
      // no cell differs from the centre: first_idx stays 0 == the BFS start == the root cell
      first_pos = root_cellpos[c];
    } else {
      first_pos = (int)(pick1 & 0xffffffffll);
      if (nties1 > 1 && !replayed) flags |= kFlagTie;
      int y = first_pos / w, x = first_pos - y * w;
      double dx = (double)cell_world(P.ox, x, P.res) - rcx, dy = (double)cell_world(P.oy, y, P.res) - rcy;
      double s = sqrt(dx * dx + dy * dy);
      fdx = dx / s;
      fdy = dy / s;
    }
    // -- farthest cell with negative dot to the first direction --
    bestk = 0;
    for (int i = threadIdx.x; i < n; i += kClThreads) {
      int p = cells[i];
      if (p == first_pos) continue;
      int y = p / w, x = p - y * w;
      double dx = (double)cell_world(P.ox, x, P.res) - rcx, dy = (double)cell_world(P.oy, y, P.res) - rcy;
      double d2 = dx * dx + dy * dy;
      double nx = dx, ny = dy;
      if (d2 > 0.0) {
        double s = sqrt(d2);
        nx = dx / s;
        ny = dy / s;
      }
      double dot = nx * fdx + ny * fdy;
      if (dot < 0.0) {
        unsigned long long k = (unsigned long long)__double_as_longlong(d2);
        bestk = k > bestk ? k : bestk;
      }
    }
    const unsigned long long max2 = block_reduce_max<unsigned long long>(bestk, s_u64);
    int second_pos;
    if (max2 != 0ull) {
      pick = 0x7fffffffffffffffll;
      ties = 0;
      for (int i = threadIdx.x; i < n; i += kClThreads) {
        int p = cells[i];
        if (p == first_pos) continue;
        int y = p / w, x = p - y * w;
        double dx = (double)cell_world(P.ox, x, P.res) - rcx, dy = (double)cell_world(P.oy, y, P.res) - rcy;
        double d2 = dx * dx + dy * dy;
        if ((unsigned long long)__double_as_longlong(d2) != max2) continue;
        double s = sqrt(d2);
        double dot = (dx / s) * fdx + (dy / s) * fdy;
        if (dot < 0.0) {
          ++ties;
          long long key = ((long long)(replayed ? i : 0) << 32) | (unsigned int)p;
          pick = key < pick ? key : pick;
        }
      }
      const int nties2 = block_reduce_sum(ties, s_int);
      const long long pick2 = block_reduce_min<long long>(pick, s_i64);
      second_pos = (int)(pick2 & 0xffffffffll);
      if (nties2 > 1 && !replayed) flags |= kFlagTie;
    } else {
      // seed_gen:1388-1399: no opposite cell -> farthest from the first cell (second_idx starts at 0)
      const int fy = first_pos / w, fx = first_pos - fy * w;
      const double fwx = (double)cell_world(P.ox, fx, P.res), fwy = (double)cell_world(P.oy, fy, P.res);
      bestk = 0;
      for (int i = threadIdx.x; i < n; i += kClThreads) {
        int p = cells[i];
        if (p == first_pos) continue;
        int y = p / w, x = p - y * w;
        double dx = (double)cell_world(P.ox, x, P.res) - fwx, dy = (double)cell_world(P.oy, y, P.res) - fwy;
        unsigned long long k = (unsigned long long)__double_as_longlong(dx * dx + dy * dy);
        bestk = k > bestk ? k : bestk;
      }
      const unsigned long long max3 = block_reduce_max<unsigned long long>(bestk, s_u64);
      pick = 0x7fffffffffffffffll;
      ties = 0;
      for (int i = threadIdx.x; i < n; i += kClThreads) {
        int p = cells[i];
        if (p == first_pos) continue;
        int y = p / w, x = p - y * w;
        double dx = (double)cell_world(P.ox, x, P.res) - fwx, dy = (double)cell_world(P.oy, y, P.res) - fwy;
        if ((unsigned long long)__double_as_longlong(dx * dx + dy * dy) == max3 && max3 != 0ull) {
          ++ties;
          long long key = ((long long)(replayed ? i : 0) << 32) | (unsigned int)p;
          pick = key < pick ? key : pick;
        }
      }
      const int nties3 = block_reduce_sum(ties, s_int);
      const long long pick3 = block_reduce_min<long long>(pick, s_i64);
      if (max3 == 0ull) second_pos = root_cellpos[c];  // second_idx stays 0
      else {
        second_pos = (int)(pick3 & 0xffffffffll);
        if (nties3 > 1 && !replayed) flags |= kFlagTie;
      }
    }
    if (threadIdx.x == 0) {
      int y1 = first_pos / w, x1 = first_pos - y1 * w, y2 = second_pos / w, x2 = second_pos - y2 * w;
      ro.row.center_x = rcx;
      ro.row.center_y = rcy;
      ro.row.start_x = (double)cell_world(P.ox, x1, P.res);
      ro.row.start_y = (double)cell_world(P.oy, y1, P.res);
      ro.row.end_x = (double)cell_world(P.ox, x2, P.res);
      ro.row.end_y = (double)cell_world(P.oy, y2, P.res);
      ro.row.length = (double)length;
      ro.row.cluster = c;
      ro.valid = 1;
    }
  }
  if (threadIdx.x == 0) {
    aos_cluster oc;
    oc.label = root_cellpos[c];
    oc.size = n;
    oc.center_x = cx;
    oc.center_y = cy;
    oc.length = length;
    oc.reserved = flags;
    oc.sum_x = (int64_t)a.sumx;
    oc.sum_y = (int64_t)a.sumy;
    oc.max_d2 = maxd2;
    out_clusters[c] = oc;
    ro.flags = flags;
    out_rows[c] = ro;
  }
}

__global__ void root_cellpos_kernel(const uint32_t *__restrict__ is_root_rank, const int *__restrict__ parent,
                                    const int *__restrict__ cellpos, int n, int *__restrict__ root_cellpos) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (parent[i] == i) root_cellpos[is_root_rank[i]] = cellpos[i];
}

// ---------------------------------------------------------------------------------------------------
// BFS-order replay (seed_gen:1008-1049): one warp per flagged cluster walks the component exactly as the
// reference's FIFO does (neighbour order dx,dy = (-1,-1),(-1,0),(-1,1),(0,-1),(0,1),(1,-1),(1,0),(1,1)),
// accumulates the float32 sums in that order and emits the cells in BFS order.  The not-yet-visited bitmap
// of the cluster's bounding box lives in shared memory (staged from the masked grid), the FIFO head is
// served from a shared-memory ring, so a step costs shared-memory latency, not L2 latency.  Up to four
// queue entries are expanded per step: lanes 8k..8k+7 own the k-th popped cell's neighbours and a lane
// yields to any lower lane that claims the same cell, which is what sequential expansion would do.
// Clusters whose bounding box does not fit the launch's shared memory use the global bitmap instead.
// ---------------------------------------------------------------------------------------------------
constexpr int kRingN = 512;  // FIFO ring entries per warp (the frontier of a skeleton is a handful of cells)

struct ReplayJob {
  int cluster;
  int wx0, y0, ww, hh;  // bounding box in words / rows
};

// Cells travel through the ring as bounding-box-relative packed coordinates  rx | ry << 20.
// SMEM = true (the default path): the CTA builds, from the cluster's OWN cell list, a BANDED bitmap in shared memory that
// keeps for every 32-cell word column only the rows between that column's lowest and highest cell (colmin = first row,
// coloff = word offset, heights by difference): at most one word per cell for a thin connected curve whatever its
// direction, where the full bounding-box bitmap of a 1 km row is 12 k words for a few thousand cells and limits an SM to
// one replay at a time.  With the band, every flagged cluster of a map is resident at once and the stage costs what its
// longest cluster costs.  A cluster whose band does not fit the launch's budget raises its fallback flag and is replayed
// by the SMEM = false instantiation on the global bitmap (launched over all jobs; the others exit at once).
template <bool SMEM>
__global__ void __launch_bounds__(32) bfs_replay_kernel(const __grid_constant__ SeedDeviceParams P,
                                                        const ReplayJob *__restrict__ jobs,
                                                        const ClusterAcc *__restrict__ acc,
                                                        const uint32_t *__restrict__ offsets,
                                                        const int *__restrict__ grouped,
                                                        const int *__restrict__ root_cellpos, int budget_words,
                                                        uint32_t *gvisited, int *__restrict__ fallback,
                                                        int *__restrict__ queue, float *__restrict__ centre_out) {
  extern __shared__ uint32_t sm[];  // ring[kRingN] | colmin[ww] | coloff[ww + 1] | band words   (the last three: SMEM)
  if (!SMEM && !fallback[blockIdx.x]) return;  // the banded instantiation already replayed this cluster
  const int lane = threadIdx.x;
  const ReplayJob job = jobs[blockIdx.x];
  const int c = job.cluster;
  const int n = (int)acc[c].size;
  int *q = queue + offsets[c];
  const int w = P.w, pitch = P.pitch;
  uint32_t *ring = sm;
  const int ww = job.ww, hh = job.hh, bx0 = job.wx0 << 5, by0 = job.y0;
  const unsigned wlim = min(ww << 5, w - bx0);  // cells of the box that exist in the image
  int *colmin = reinterpret_cast<int *>(sm + kRingN);
  int *coloff = colmin + ww;  // first used as the per-column maximum
  uint32_t *band = reinterpret_cast<uint32_t *>(coloff + ww + 1);
  if (SMEM) {
    const int *cells = grouped + offsets[c];
    for (int i = lane; i < ww; i += 32) {
      colmin[i] = 0x7fffffff;
      coloff[i] = -1;
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      int pos = cells[k];
      int y = pos / w, x = pos - y * w;
      int col = (x - bx0) >> 5;
      atomicMin(&colmin[col], y - by0);
      atomicMax(&coloff[col], y - by0);
    }
    __syncwarp();
    int running = 0;
    for (int base = 0; base < ww; base += 32) {
      int i = base + lane;
      int h = 0;
      if (i < ww && coloff[i] >= 0) h = coloff[i] - colmin[i] + 1;
      int inc = (int)warp_incl_scan((uint32_t)h, lane);
      if (i < ww) coloff[i] = running + inc - h;
      running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) coloff[ww] = running;
    __syncwarp();
    if (running > budget_words - kRingN - 2 * ww - 1) {  // does not fit this launch's shared memory
      if (lane == 0) fallback[blockIdx.x] = 1;
      return;
    }
    for (int i = lane; i < running; i += 32) band[i] = 0u;
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      int pos = cells[k];
      int y = pos / w, x = pos - y * w;
      int col = (x - bx0) >> 5;
      atomicOr(&band[coloff[col] + (y - by0) - colmin[col]], 1u << (x & 31));
    }
    __syncwarp();
  }
  // the not-yet-visited bitmap word of box cell (rx, ry): the band (null where the column has no such row), or the
  // global scratch copy of the whole mask
  auto word_ptr = [&](int rx, int ry) -> uint32_t * {
    if (SMEM) {
      const int col = rx >> 5;
      const int o = coloff[col], r = ry - colmin[col];
      return (unsigned)r < (unsigned)(coloff[col + 1] - o) ? band + o + r : nullptr;
    }
    return gvisited + (size_t)(ry + by0) * pitch + job.wx0 + (rx >> 5);
  };
  // neighbour k of the reference's tables dx = {-1,-1,-1,0,0,1,1,1}, dy = {-1,0,1,-1,1,-1,0,1}
  const int sub = lane >> 3, k8 = lane & 7;
  const int kk = k8 + (k8 >= 4);  // 0..8 without the centre (4)
  const int mydx = kk / 3 - 1, mydy = kk % 3 - 1;
  const unsigned lt = (1u << lane) - 1u;
  int head = 0, tail = 1;
  float sum_x = 0.f, sum_y = 0.f;
  {
    int start = root_cellpos[c];
    int y = start / w, x = start - y * w;
    if (lane == 0) {
      q[0] = start;
      ring[0] = (uint32_t)(x - bx0) | ((uint32_t)(y - by0) << 20);
      uint32_t *wp = word_ptr(x - bx0, y - by0);
      if (wp) atomicAnd(wp, ~(1u << (x & 31)));
    }
    __syncwarp();
  }
  while (head < tail) {
    const int avail = min(tail - head, 4);
    // pop up to four cells in FIFO order; lanes 8k..8k+7 expand the k-th one
    uint32_t cur = 0xffffffffu;
    if (sub < avail) {
      int idx = head + sub;
      if (tail - idx <= kRingN) cur = ring[idx & (kRingN - 1)];
      else {  // fell out of the ring (very wide frontier): re-read from the global queue
        int pos = __ldcg(q + idx);
        int y = pos / w;
        cur = (uint32_t)(pos - y * w - bx0) | ((uint32_t)(y - by0) << 20);
      }
    }
    const int rx = (int)(cur & 0xfffffu) + mydx, ry = (int)(cur >> 20) + mydy;
    bool take = false;
    uint32_t *wp = nullptr;
    if (sub < avail && (unsigned)rx < wlim && (unsigned)ry < (unsigned)hh) {
      wp = word_ptr(rx, ry);
      uint32_t word = SMEM ? (wp ? *wp : 0u) : __ldcg(wp);
      take = (word >> (rx & 31)) & 1u;
    }
    // cells popped earlier in this step claim shared neighbours first (what sequential expansion does)
    const uint32_t c0 = __shfl_sync(0xffffffffu, cur, 0), c1 = __shfl_sync(0xffffffffu, cur, 8),
                   c2 = __shfl_sync(0xffffffffu, cur, 16), c3 = __shfl_sync(0xffffffffu, cur, 24);
    auto near = [&](uint32_t p) { return abs(rx - (int)(p & 0xfffffu)) <= 1 && abs(ry - (int)(p >> 20)) <= 1; };
    if (sub >= 1 && near(c0)) take = false;
    if (sub >= 2 && near(c1)) take = false;
    if (sub >= 3 && near(c2)) take = false;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (take) {
      int slot = tail + __popc(m & lt);
      ring[slot & (kRingN - 1)] = (uint32_t)rx | ((uint32_t)ry << 20);
      q[slot] = (ry + by0) * w + rx + bx0;
      atomicAnd(wp, ~(1u << (rx & 31)));
    }
    head += avail;
    tail += __popc(m);
    __syncwarp();
  }
  // float32 running sums in FIFO order (seed_gen:1053-1057): the queue now holds the BFS order; lanes fetch 32
  // cells at a time, the additions themselves stay strictly sequential (every lane keeps the same copy)
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    int pos = base + lane < n ? __ldcg(q + base + lane) : 0;
    int y = pos / w;
    float xf = (float)(pos - y * w), yf = (float)y;
    const int cnt = min(32, n - base);
    for (int j = 0; j < cnt; ++j) {
      sum_x = __fadd_rn(sum_x, __shfl_sync(0xffffffffu, xf, j));
      sum_y = __fadd_rn(sum_y, __shfl_sync(0xffffffffu, yf, j));
    }
  }
  if (lane == 0) {
    centre_out[2 * c] = __fdiv_rn(sum_x, (float)(unsigned long long)n);
    centre_out[2 * c + 1] = __fdiv_rn(sum_y, (float)(unsigned long long)n);
  }
}

// ---------------------------------------------------------------------------------------------------
// BFS order without visiting the cells one by one (the default path; bfs_replay_kernel above is its fallback).
// A thinned skeleton is almost everywhere a CHAIN: 97 % of its cells have exactly two neighbours that do not touch
// each other.  Once the FIFO has entered a maximal run of such cells from one end it can only walk it cell by cell, one
// cell per BFS level, and nothing else can interfere before the run's other end (a chain cell has no other neighbours).
// So while every cell of the current level is such a walker the next levels are known in advance: level l + t holds each
// walker's t-th successor, in the same order.  The replay therefore
//   1. finds every cell's neighbours and the chain cells (bfs_prepare_kernel; the BFS root is never a chain cell),
//   2. ranks the runs by pointer jumping along both directions at once (bfs_link_init / bfs_jump: log2(longest run)
//      rounds), which gives each chain cell its run, its position in it and the run's length, and lays the runs out as
//      arrays (bfs_runs / bfs_scatter),
//   3. walks the component with one warp (bfs_chain_kernel): a level that contains an irregular cell (junction, corner,
//      end, blob) or a walker at the end of its run is expanded literally -- eight lanes per cell test the neighbours in
//      the reference's order, four cells at a time, equal claims resolved in lane order -- and a level of walkers is
//      fast-forwarded by as many levels as the nearest event allows (end of a run; two walkers that entered one run from
//      both ends meet in its middle), the skipped cells written by all lanes at once.
// For the longest row of a config-3 map (9823 cells) that is ~300 sequential steps instead of 9823.  A visited chain cell
// is not flagged: of a run only the prefix [0, lo) and the suffix [hi, L) can have been visited.
// scripts/dev/chain_bfs_proto.py is the same algorithm in Python, checked against the plain FIFO on random images.
// ---------------------------------------------------------------------------------------------------
constexpr int kItemCap = 64;   // cells of one BFS level the warp keeps in shared memory; wider levels -> fallback kernel
constexpr int kRunHdr = 8;     // ints in front of a run's cells: lo, hi, L, out0, out1 (irregular-table slots beyond either end)
constexpr unsigned kDone = 0x80000000u;

struct IrrRec {    // one per irregular cell of a flagged cluster (junctions, corners, ends, blobs, the root): 80 bytes
  int2 nb[8];      // neighbour k in the reference's order: (run base, position) of a chain cell, (-1 - slot, 0) of an
                   // irregular one, (INT_MIN, 0) = none.  While the table is being built nb[k].x is the compact index.
  int cellpos;
  int vis;
  int pad[2];
};

struct BfsBufs {
  int2 *link;      // chain cell: its two neighbours (compact indices, reference order); irregular: (-1, slot in irr);
                   // cell of a cluster that needs no order: (-2, -2)
  uint2 *state;    // 2 per cell (one per link), pointer jumping: x = entry index of the farthest known chain cell that way,
                   // y = steps | kDone once x is the end of the run
  int4 *cinfo;     // ALIASES state (same 16 bytes per cell) once the runs are ranked: x = run base, y = position, z = L
  int *runs;       // per run: kRunHdr header ints, then cellpos of its cells in run order
  IrrRec *irr;
  int *ctr;        // [0] ints of `runs` handed out, [1] clusters left to the fallback, [2] ranking broken / out of space,
                   // [3] irregular cells, [4 + r] round r changed something
  int run_cap, irr_cap;
};

__global__ void bfs_prepare_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ prefix, int pitch, int h, int w,
                                   const int *__restrict__ cellpos, const int *__restrict__ cell_cluster,
                                   const int *__restrict__ root_cellpos, const RowOut *__restrict__ rows, int n, BfsBufs B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = cell_cluster[i];
    if (!(rows[c].flags & (kFlagNeedsOrder | kFlagTie))) {
      B.link[i] = make_int2(-2, -2);
      continue;
    }
    const int pos = cellpos[i];
    const int y = pos / w, x = pos - y * w;
    int ids[8];
    int deg = 0, n0 = -1, n1 = -1, k0 = 0, k1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int kk = k + (k >= 4);
      const int nx = x + kk / 3 - 1, ny = y + kk % 3 - 1;
      int id = -1;
      if ((unsigned)nx < (unsigned)w && (unsigned)ny < (unsigned)h) id = compact_index(mask, prefix, pitch, nx, ny);
      ids[k] = id;
      if (id >= 0) {
        if (deg == 0) n0 = id, k0 = kk;
        else if (deg == 1) n1 = id, k1 = kk;
        ++deg;
      }
    }
    // a chain cell: two neighbours that are not neighbours of each other; the root starts the FIFO and stays irregular
    const bool apart = abs(k0 / 3 - k1 / 3) > 1 || abs(k0 % 3 - k1 % 3) > 1;
    if (deg == 2 && apart && pos != root_cellpos[c]) {
      B.link[i] = make_int2(n0, n1);
      continue;
    }
    const int slot = atomicAdd(&B.ctr[3], 1);
    if (slot >= B.irr_cap) {
      B.ctr[2] = 1;
      B.link[i] = make_int2(-1, 0);
      continue;
    }
    B.link[i] = make_int2(-1, slot);
    IrrRec *r = B.irr + slot;
#pragma unroll
    for (int k = 0; k < 8; ++k) r->nb[k] = make_int2(ids[k], 0);
    r->cellpos = pos;
    r->vis = 0;
  }
}

__global__ void bfs_link_init_kernel(int n, BfsBufs B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 l = B.link[i];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      uint2 st = make_uint2(2u * (unsigned)i + s, kDone);  // not a chain cell: nothing to jump
      if (l.x >= 0) {
        const int t = s ? l.y : l.x;
        const int2 lt = B.link[t];
        // the neighbour continues the run through its OTHER link; an irregular neighbour ends the run at this cell
        if (lt.x >= 0) st = make_uint2(2u * (unsigned)t + (lt.x == i ? 1u : 0u), 1u);
      }
      B.state[2 * (size_t)i + s] = st;
    }
  }
}

// Pointer jumping, in place: an entry that reads its target before or after the target's own update composes two true
// statements either way (8-byte entries are loaded and stored whole).  Round 0 visits every entry and hops up to eight
// times; runs are short on average (one irregular cell per ~16 chain cells), so most entries arrive at once.  The ones
// that do not (entries of long runs) are listed -- in B.runs, which the ranking does not need yet -- and the following
// rounds only walk that list, four hops each.
__device__ __forceinline__ bool bfs_hop(BfsBufs &B, unsigned idx, int hops) {
  uint2 st = __ldcg(B.state + idx);
  if (st.y & kDone) return true;
#pragma unroll 1
  for (int hop = 0; hop < hops && !(st.y & kDone); ++hop) {
    const uint2 st2 = __ldcg(B.state + st.x);
    st = make_uint2(st2.x, st.y + st2.y);  // the target's kDone bit carries over: its x is the run's end
  }
  __stcg(B.state + idx, st);
  return (st.y & kDone) != 0;
}

__global__ void bfs_jump_first_kernel(int n, BfsBufs B) {
  const unsigned total = 2u * (unsigned)n;
  const int lane = threadIdx.x & 31;
  unsigned *list = reinterpret_cast<unsigned *>(B.runs);
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < total; base += gridDim.x * blockDim.x) {
    const unsigned idx = base + lane;
    const bool open = idx < total && !bfs_hop(B, idx, 8);
    const unsigned m = __ballot_sync(0xffffffffu, open);
    if (!m) continue;
    int at = 0;
    if (lane == 0) at = atomicAdd(&B.ctr[40], __popc(m));
    at = __shfl_sync(0xffffffffu, at, 0);
    if (open) list[at + __popc(m & ((1u << lane) - 1u))] = idx;  // at most 2 n entries: B.runs holds 3 n ints
  }
}

__global__ void bfs_jump_kernel(BfsBufs B, int round) {
  if (round > 0 && B.ctr[4 + round - 1] == 0) return;  // the previous round left nothing open
  const unsigned *list = reinterpret_cast<const unsigned *>(B.runs);
  const unsigned total = (unsigned)B.ctr[40];
  bool unfinished = false;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    unfinished |= !bfs_hop(B, list[t], 4);
  if (__syncthreads_or(unfinished) && threadIdx.x == 0) B.ctr[4 + round] = 1;  // one store per CTA, not per entry
}

__global__ void bfs_runs_kernel(int n, BfsBufs B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 l = B.link[i];
    if (l.x < 0) continue;
    const uint2 e0 = B.state[2 * (size_t)i], e1 = B.state[2 * (size_t)i + 1];
    if (!(e0.y & e1.y & kDone)) {  // too few rounds (cannot happen: the host sizes them by the largest cluster)
      B.ctr[2] = 1;
      continue;
    }
    const int d0 = (int)(e0.y & ~kDone), d1 = (int)(e1.y & ~kDone);
    const int E0 = (int)(e0.x >> 1), E1 = (int)(e1.x >> 1), L = d0 + d1 + 1;
    int head, pos, rb = -1;
    if (E0 == E1) {  // a run of one cell (a closed ring of chain cells cannot exist: every component has its root)
      head = i;
      pos = 0;
      if (L != 1) B.ctr[2] = 1;
    } else if (E0 < E1) {
      head = E0;
      pos = d0;
    } else {
      head = E1;
      pos = d1;
    }
    if (head == i) {
      const int ints = (L + kRunHdr + 3) & ~3;  // headers are read as int4
      rb = atomicAdd(&B.ctr[0], ints);
      if (rb + ints > B.run_cap) {
        B.ctr[2] = 1;
        rb = -1;
      } else {
        // the irregular cells beyond either end: this cell's link on the side where the run ends at once, and the far
        // end cell's link on the side its entry names
        int out0, out1;
        if (L == 1) {
          out0 = l.x;
          out1 = l.y;
        } else {
          const bool side0_ends_here = (int)(e0.x >> 1) == i;
          out0 = side0_ends_here ? l.x : l.y;
          const uint2 far = side0_ends_here ? e1 : e0;
          const int2 lf = B.link[far.x >> 1];
          out1 = (far.x & 1u) ? lf.y : lf.x;
        }
        int *hdr = B.runs + rb;
        hdr[0] = 0;
        hdr[1] = L;
        hdr[2] = L;
        hdr[3] = B.link[out0].y;
        hdr[4] = B.link[out1].y;
      }
    }
    B.cinfo[i] = make_int4(head, pos, L, rb);  // over this cell's own two state entries: nobody else reads them here
  }
}

__global__ void bfs_scatter_kernel(int n, const int *__restrict__ cellpos, BfsBufs B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 l = B.link[i];
    if (l.x >= 0) {
      const int4 ci = B.cinfo[i];
      const int rb = B.cinfo[ci.x].w;  // the head's slot; only .x of a record changes in this kernel
      B.cinfo[i].x = rb;
      if (rb >= 0) B.runs[rb + kRunHdr + ci.y] = cellpos[i];
    }
  }
}

// neighbour ids of the irregular cells -> what the walk needs to know about each neighbour
__global__ void bfs_irr_kernel(int n, BfsBufs B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 l = B.link[i];
    if (l.x != -1) continue;
    IrrRec *r = B.irr + l.y;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = r->nb[k].x;
      int2 v = make_int2(INT_MIN, 0);
      if (j >= 0) {
        const int2 lj = B.link[j];
        if (lj.x < 0) v = make_int2(-1 - lj.y, 0);
        else {
          const int4 cj = B.cinfo[j];
          v = make_int2(cj.x, cj.y);
        }
      }
      r->nb[k] = v;
    }
  }
}

// float32 running sums in FIFO order (seed_gen:1053-1057) over a finished BFS-order cell list: lanes fetch 32 cells at a
// time, the additions themselves stay strictly sequential (every lane keeps the same copy)
__device__ __forceinline__ void ordered_centre(const int *q, int n, int w, int lane, float *cx, float *cy) {
  float sum_x = 0.f, sum_y = 0.f;
  int nxt = lane < n ? __ldcg(q + lane) : 0;
  for (int base = 0; base < n; base += 32) {
    const int pos = nxt;
    nxt = base + 32 + lane < n ? __ldcg(q + base + 32 + lane) : 0;  // in flight while this batch is added up
    int y = pos / w;
    float xf = (float)(pos - y * w), yf = (float)y;
    if (n - base >= 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sum_x = __fadd_rn(sum_x, __shfl_sync(0xffffffffu, xf, j));
        sum_y = __fadd_rn(sum_y, __shfl_sync(0xffffffffu, yf, j));
      }
    } else {
      const int cnt = n - base;
      for (int j = 0; j < cnt; ++j) {
        sum_x = __fadd_rn(sum_x, __shfl_sync(0xffffffffu, xf, j));
        sum_y = __fadd_rn(sum_y, __shfl_sync(0xffffffffu, yf, j));
      }
    }
  }
  *cx = __fdiv_rn(sum_x, (float)(unsigned long long)n);
  *cy = __fdiv_rn(sum_y, (float)(unsigned long long)n);
}

// ---- where the walk keeps what it reads and writes at every level ------------------------------------------------
// GlobalStore: the tables as the ranking kernels left them (a run is named by its base in B.runs, an irregular cell by its
// slot in B.irr).  SharedStore: a copy of THIS cluster's irregular records and run headers in shared memory, indexed locally;
// a level then costs shared-memory latency instead of two L2 round trips (0.3-0.4 us each on B200), which is what the walk
// of a long row consists of.  The cells of the runs stay in B.runs either way (they are only read to be written out).
struct GlobalStore {
  BfsBufs B;
  __device__ __forceinline__ int2 nb(int irr, int k) const { return __ldg(&B.irr[irr].nb[k]); }
  __device__ __forceinline__ int2 cellpos_vis(int irr) const { return __ldcg(reinterpret_cast<const int2 *>(&B.irr[irr].cellpos)); }
  __device__ __forceinline__ void set_vis(int irr) const { B.irr[irr].vis = 1; }
  __device__ __forceinline__ int4 hdr(int run) const { return __ldcg(reinterpret_cast<const int4 *>(B.runs + run)); }  // lo, hi, L, out0
  __device__ __forceinline__ int out1(int run) const { return __ldg(B.runs + run + 4); }
  __device__ __forceinline__ int base(int run) const { return run; }
  __device__ __forceinline__ void set_lo(int run, int v) const { __stcg(B.runs + run, v); }
  __device__ __forceinline__ void set_hi(int run, int v) const { __stcg(B.runs + run + 1, v); }
};

struct SIrr {  // 72 bytes
  int2 nb[8];
  int cellpos, vis;
};
struct SHdr {  // 24 bytes
  int lo, hi, L, out0, out1, rb;
};
struct SharedStore {
  SIrr *irr;
  SHdr *hd;
  __device__ __forceinline__ int2 nb(int i, int k) const { return irr[i].nb[k]; }
  __device__ __forceinline__ int2 cellpos_vis(int i) const { return make_int2(irr[i].cellpos, irr[i].vis); }
  __device__ __forceinline__ void set_vis(int i) const { irr[i].vis = 1; }
  __device__ __forceinline__ int4 hdr(int r) const { return make_int4(hd[r].lo, hd[r].hi, hd[r].L, hd[r].out0); }
  __device__ __forceinline__ int out1(int r) const { return hd[r].out1; }
  __device__ __forceinline__ int base(int r) const { return hd[r].rb; }
  __device__ __forceinline__ void set_lo(int r, int v) const { hd[r].lo = v; }
  __device__ __forceinline__ void set_hi(int r, int v) const { hd[r].hi = v; }
};

// The walk itself, by ONE warp.  items: x = run (walker) or -1 (irregular), y = position / irregular index, z = direction.
// It writes REFERENCES into q -- the index of a chain cell in B.runs, or -1 - (irregular index) -- so that no global load
// sits on its path; the caller turns them into cell positions afterwards, all threads at once.
// Returns the number of cells written to q, or -1 when a level did not fit the list.
template <typename Store>
__device__ int bfs_chain_walk(const Store S, int4 (*s_items)[kItemCap], int item_cap, int root_irr, int n, int *q, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  int cur = 0, cnt = 1, out = 1;
  if (lane == 0) {
    q[0] = -1 - root_irr;
    S.set_vis(root_irr);
    s_items[0][0] = make_int4(-1, root_irr, 0, 0);
  }
  __syncwarp();
  const int sub = lane >> 3, k8 = lane & 7;
  while (cnt > 0) {
    // ---- how many levels can the whole list be fast-forwarded? (only a list of walkers can) ----
    bool walkers_only = true;
    for (int base = 0; base < cnt; base += 32)
      if (base + lane < cnt && s_items[cur][base + lane].x < 0) walkers_only = false;
    int f = 0;
    if (__all_sync(0xffffffffu, walkers_only)) {
      f = 0x7fffffff;
      for (int base = 0; base < cnt; base += 32) {
        if (base + lane < cnt) {
          const int4 it = s_items[cur][base + lane];
          const int4 hd = S.hdr(it.x);
          const int gap = hd.y - hd.x;
          f = min(f, (hd.x > 0 && hd.y < hd.z) ? gap >> 1 : gap);  // entered from both ends: the walkers share what is left
        }
      }
      f = __reduce_min_sync(0xffffffffu, f);
    }
    if (f >= 1) {
      if (out + (long long)f * cnt > n) return out;  // cannot happen; the caller's check reports it
      const int total = f * cnt;
      for (int idx = lane; idx < total; idx += 32) {
        const int t = idx / cnt, i = idx - t * cnt;
        const int4 it = s_items[cur][i];
        q[out + idx] = S.base(it.x) + kRunHdr + it.y + it.z * (t + 1);
      }
      __syncwarp();
      for (int i = lane; i < cnt; i += 32) {
        int4 it = s_items[cur][i];
        it.y += it.z * f;
        if (it.z > 0) S.set_lo(it.x, it.y + 1);
        else S.set_hi(it.x, it.y);
        s_items[cur][i] = it;
      }
      out += total;
      __syncwarp();
      continue;
    }
    // ---- one literal level: four cells per step.  Lanes 8 s .. 8 s + 7 test the neighbours of the s-th one if it is
    // irregular; a walker has one way to go (lane 8 s): the next cell of its run, or the irregular cell beyond its end ----
    int ncnt = 0;
    for (int base = 0; base < cnt; base += 4) {
      const bool valid = base + sub < cnt;
      const int4 it = valid ? s_items[cur][base + sub] : make_int4(-1, -1, 0, 0);
      // target: tx >= 0: chain cell (run tx, position ty), direction tz; tx < 0: irregular cell -1 - tx; tkey = its reference
      int tx = INT_MIN, ty = 0, tz = 0, tkey = 0;
      bool take = false;
      if (valid && it.x < 0) {
        const int2 nb = S.nb(it.y, k8);
        tx = nb.x;
        ty = nb.y;
        if (tx >= 0) {
          const int4 hd = S.hdr(tx);
          tkey = S.base(tx) + kRunHdr + ty;
          take = hd.x <= ty && ty < hd.y;
          // entered from outside at one of the run's ends; a run of one cell is entered "forwards" from its out0 side
          tz = hd.z == 1 ? (hd.w == it.y ? 1 : -1) : (ty == 0 ? 1 : -1);
        } else if (tx != INT_MIN) {
          take = S.cellpos_vis(-1 - tx).y == 0;
          tkey = tx;
        }
      } else if (valid && k8 == 0) {
        const int4 hd = S.hdr(it.x);
        const int o1 = S.out1(it.x);
        const int gap = hd.y - hd.x;
        if (gap >= 1) {  // the next cell of the run
          tx = it.x;
          ty = it.y + it.z;
          tz = it.z;
          tkey = S.base(it.x) + kRunHdr + ty;
          take = true;
        } else if (!(hd.x > 0 && hd.y < hd.z)) {  // walked to the end: the irregular cell beyond it
          const int o = it.z > 0 ? o1 : hd.w;
          tx = -1 - o;
          take = S.cellpos_vis(o).y == 0;
          tkey = tx;
        }
      }
      // a cell claimed by two lanes of this step goes to the lower lane: the cell popped earlier, as in the FIFO
      const unsigned same = __match_any_sync(0xffffffffu, take ? tkey : INT_MIN + 1 + lane);
      take = take && (__ffs(same) - 1 == lane);
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (ncnt + __popc(m) > item_cap) return -1;
      if (out + ncnt + __popc(m) > n) return out + ncnt + __popc(m);
      if (take) {
        const int slot = ncnt + __popc(m & lt);
        if (tx >= 0) {
          if (tz > 0) S.set_lo(tx, ty + 1);
          else S.set_hi(tx, ty);
          s_items[cur ^ 1][slot] = make_int4(tx, ty, tz, 0);
        } else {
          S.set_vis(-1 - tx);
          s_items[cur ^ 1][slot] = make_int4(-1, -1 - tx, 0, 0);
        }
        q[out + slot] = tkey;
      }
      ncnt += __popc(m);
      __syncwarp();
    }
    out += ncnt;
    cnt = ncnt;
    cur ^= 1;
    __syncwarp();
  }
  return out;
}

constexpr int kChainThreads = 256;   // threads per cluster (1024 for the longest rows: staging is a latency-bound gather)
// One CTA per flagged cluster.  All of its threads stage the cluster's irregular records and run headers in shared memory
// (if they fit `budget` bytes; otherwise the walk reads the global tables), warp 0 walks, all threads turn the references
// it wrote into cell positions, warp 0 adds up the float32 centre in that order.
__global__ void __launch_bounds__(1024) bfs_chain_kernel(const int *__restrict__ flagged, const ClusterAcc *__restrict__ acc,
                                                                  const uint32_t *__restrict__ offsets,
                                                                  const int *__restrict__ grouped_id,
                                                                  const int *__restrict__ root_cellpos,
                                                                  const uint32_t *__restrict__ mask,
                                                                  const uint32_t *__restrict__ prefix, int pitch, int w, BfsBufs B,
                                                                  int item_cap, int budget, int *__restrict__ fallback,
                                                                  int *__restrict__ queue, float *__restrict__ centre_out) {
  __shared__ int4 s_items[2][kItemCap];
  __shared__ int s_nirr, s_nrun, s_out;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int c = flagged[blockIdx.x];
  const int n = (int)acc[c].size;
  int *q = queue + offsets[c];
  if (tid == 0) {
    fallback[blockIdx.x] = 0;
    s_nirr = 0;
    s_nrun = 0;
  }
  const bool broken = __ldcg(&B.ctr[2]) != 0;  // the ranking found something it does not understand
  __syncthreads();
  const int rpos = root_cellpos[c];
  const int root = compact_index(mask, prefix, pitch, rpos - (rpos / w) * w, rpos / w);
  // ---- stage this cluster's tables: count its irregular cells and runs, give them local numbers ----
  bool shared_ok = !broken && budget > 0;
  const int *cells = grouped_id + offsets[c];
  if (shared_ok) {
    for (int k = tid; k < n; k += (int)blockDim.x) {
      const int id = cells[k];
      const int2 l = B.link[id];
      if (l.x == -1) {
        B.irr[l.y].pad[0] = atomicAdd(&s_nirr, 1);
      } else if (l.x >= 0) {
        const int4 ci = B.cinfo[id];
        if (ci.y == 0) B.runs[ci.x + 5] = atomicAdd(&s_nrun, 1);  // the head cell numbers its run
      }
    }
    __syncthreads();
    shared_ok = (size_t)s_nirr * sizeof(SIrr) + (size_t)s_nrun * sizeof(SHdr) <= (size_t)budget;
  }
  SharedStore SS;
  SS.irr = reinterpret_cast<SIrr *>(s_dyn);
  SS.hd = reinterpret_cast<SHdr *>(s_dyn + (size_t)s_nirr * sizeof(SIrr));
  if (shared_ok) {
    for (int k = tid; k < n; k += (int)blockDim.x) {
      const int id = cells[k];
      const int2 l = B.link[id];
      if (l.x == -1) {
        const IrrRec *r = B.irr + l.y;
        SIrr o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int2 v = r->nb[j];
          if (v.x >= 0) v.x = __ldcg(B.runs + v.x + 5);                         // run base -> local run number
          else if (v.x != INT_MIN) v.x = -1 - __ldcg(&B.irr[-1 - v.x].pad[0]);  // slot -> local irregular number
          o.nb[j] = v;
        }
        o.cellpos = r->cellpos;
        o.vis = 0;
        SS.irr[__ldcg(&r->pad[0])] = o;
      } else if (l.x >= 0) {
        const int4 ci = B.cinfo[id];
        if (ci.y == 0) {
          const int *h = B.runs + ci.x;
          SHdr o;
          o.lo = 0;
          o.hi = o.L = h[2];
          o.out0 = __ldcg(&B.irr[h[3]].pad[0]);
          o.out1 = __ldcg(&B.irr[h[4]].pad[0]);
          o.rb = ci.x;
          SS.hd[__ldcg(h + 5)] = o;
        }
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    int out;
    if (broken) out = -1;
    else if (shared_ok) out = bfs_chain_walk(SS, s_items, item_cap, __ldcg(&B.irr[B.link[root].y].pad[0]), n, q, lane);
    else out = bfs_chain_walk(GlobalStore{B}, s_items, item_cap, B.link[root].y, n, q, lane);
    if (lane == 0) s_out = out;
  }
  __syncthreads();
  if (s_out != n) {  // a level wider than the list (or, defensively, a count that does not add up): the literal replay takes it
    if (tid == 0) {
      fallback[blockIdx.x] = 1;
      atomicAdd(&B.ctr[1], 1);
    }
    return;
  }
  // references -> cell positions
  for (int i = tid; i < n; i += (int)blockDim.x) {
    const int ref = q[i];
    q[i] = ref >= 0 ? B.runs[ref] : shared_ok ? SS.irr[-1 - ref].cellpos : B.irr[-1 - ref].cellpos;
  }
  __syncthreads();
  if (tid >= 32) return;
  float cx, cy;
  ordered_centre(q, n, w, lane, &cx, &cy);
  if (lane == 0) {
    centre_out[2 * c] = cx;
    centre_out[2 * c + 1] = cy;
  }
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
static inline int grid_for(size_t n, int threads, int cap_mult = 16) {
  size_t b = (n + threads - 1) / threads;
  size_t cap = (size_t)kNumSMs * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

aos_status run_clusters(Ctx *c, const SeedDeviceParams &P, const uint32_t *skel, float min_length) {
  const size_t words = (size_t)P.pitch * P.h;
  cudaStream_t st = c->stream;
  c->n_clusters = 0;
  c->n_skel_cells = 0;
  c->h_clusters.clear();
  c->h_rows.clear();
  AOS_CUDA_OK(c, c->cc_mask.reserve(words * 4));
  AOS_CUDA_OK(c, c->cc_prefix.reserve(words * 4));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  uint32_t *mask = c->cc_mask.as<uint32_t>();
  uint32_t *prefix = c->cc_prefix.as<uint32_t>();
  uint32_t *d_tot = c->misc.as<uint32_t>();  // [0] skeleton cells, [1] clusters

  mask_count_kernel<<<grid_for(words, 256), 256, 0, st>>>(P, skel, mask, prefix);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  aos_status s = exclusive_scan_u32(c, prefix, words, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->mark("cc_mask_scan");
  const int n = c->h_flag[0];
  c->n_skel_cells = n;
  if (n == 0) return AOS_OK;

  AOS_CUDA_OK(c, c->cc_parent.reserve(sizeof(int) * (size_t)n));
  AOS_CUDA_OK(c, c->cc_cellpos.reserve(sizeof(int) * (size_t)n));
  AOS_CUDA_OK(c, c->cc_rootrank.reserve(sizeof(uint32_t) * (size_t)n));
  int *parent = c->cc_parent.as<int>();
  int *cellpos = c->cc_cellpos.as<int>();
  uint32_t *rootrank = c->cc_rootrank.as<uint32_t>();
  cc_init_kernel<<<grid_for(words, 256), 256, 0, st>>>(mask, prefix, P.pitch, P.h, P.w, parent, cellpos);
  ++c->launches;
  cc_link_kernel<<<grid_for(n, 256), 256, 0, st>>>(mask, prefix, P.pitch, P.w, cellpos, n, parent);
  ++c->launches;
  cc_flatten_kernel<<<grid_for(n, 256), 256, 0, st>>>(parent, rootrank, n);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  s = exclusive_scan_u32(c, rootrank, (size_t)n, c->cc_blocksum, d_tot + 1);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot + 1, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->mark("cc_link_rank");
  const int nc = c->h_flag[0];
  c->n_clusters = nc;

  // layout of cl_aux: ClusterAcc[nc] | sizes/offsets u32[nc+1] | root_cellpos int[nc] | centre f32[2nc] | flagged int[nc]
  size_t off_acc = 0;
  size_t off_offsets = off_acc + sizeof(ClusterAcc) * (size_t)nc;
  size_t off_rootpos = off_offsets + sizeof(uint32_t) * (size_t)(nc + 4);
  size_t off_centre = off_rootpos + sizeof(int) * (size_t)(nc + 4);
  size_t off_flagged = off_centre + sizeof(float) * 2 * (size_t)(nc + 4);
  size_t aux_bytes = off_flagged + sizeof(int) * (size_t)(nc + 4);
  AOS_CUDA_OK(c, c->cl_aux.reserve(aux_bytes));
  char *aux = c->cl_aux.as<char>();
  ClusterAcc *acc = reinterpret_cast<ClusterAcc *>(aux + off_acc);
  uint32_t *offsets = reinterpret_cast<uint32_t *>(aux + off_offsets);
  int *root_cellpos = reinterpret_cast<int *>(aux + off_rootpos);
  float *centre = reinterpret_cast<float *>(aux + off_centre);

  // cell_cluster int[n] | grouped int[n] | BFS-order queue int[n] | grouped_id int[n]
  AOS_CUDA_OK(c, c->cand_buf.reserve(sizeof(int) * 4 * (size_t)n));
  int *cell_cluster = c->cand_buf.as<int>();
  int *grouped = cell_cluster + n;
  int *queue = grouped + n;
  int *grouped_id = queue + n;
  c->d_cell_cluster = cell_cluster;
  c->d_root_cellpos = root_cellpos;

  acc_init_kernel<<<(nc + 127) / 128, 128, 0, st>>>(acc, nc);
  ++c->launches;
  root_cellpos_kernel<<<grid_for(n, 256), 256, 0, st>>>(rootrank, parent, cellpos, n, root_cellpos);
  ++c->launches;
  cc_accumulate_kernel<<<grid_for(n, 256), 256, 0, st>>>(parent, rootrank, cellpos, n, P.w, cell_cluster, acc);
  ++c->launches;
  acc_sizes_kernel<<<(nc + 127) / 128, 128, 0, st>>>(acc, nc, offsets);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  s = exclusive_scan_u32(c, offsets, (size_t)nc, c->cc_blocksum, d_tot + 2);
  if (s != AOS_OK) return s;
  cc_group_kernel<<<grid_for(n, 256), 256, 0, st>>>(cell_cluster, cellpos, n, offsets, acc, grouped, grouped_id);
  ++c->launches;

  c->mark("cc_stats_group");
  AOS_CUDA_OK(c, c->cl_table.reserve(sizeof(aos_cluster) * (size_t)nc + sizeof(RowOut) * (size_t)nc));
  aos_cluster *d_clusters = c->cl_table.as<aos_cluster>();
  RowOut *d_rows = reinterpret_cast<RowOut *>(d_clusters + nc);
  cluster_finalize_kernel<<<nc, kClThreads, 0, st>>>(P, acc, offsets, grouped, root_cellpos, min_length, nullptr, nullptr,
                                                     nullptr, nullptr, d_clusters, d_rows);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());

  // the per-cluster tables come back through page-locked staging (plain DMA; see Ctx::h_merged)
  c->h_clusters.resize(nc);
  if (!c->pin_a.resize(sizeof(aos_cluster) * (size_t)nc) || !c->pin_b.resize(sizeof(RowOut) * (size_t)nc)) {
    set_error(c, "cudaHostAlloc failed (cluster tables)");
    return AOS_ERR_CUDA;
  }
  RowOut *h_rows = reinterpret_cast<RowOut *>(c->pin_b.data());
  auto fetch_tables = [&]() -> aos_status {
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->pin_a.data(), d_clusters, sizeof(aos_cluster) * (size_t)nc, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaMemcpyAsync(h_rows, d_rows, sizeof(RowOut) * (size_t)nc, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    memcpy(c->h_clusters.data(), c->pin_a.data(), sizeof(aos_cluster) * (size_t)nc);
    return AOS_OK;
  };
  s = fetch_tables();
  if (s != AOS_OK) return s;

  c->mark("cc_finalize");
  // ---- order-dependent clusters: the reference's BFS order, then finalise again ------------------------
  // needed when a float32 partial sum can round (sum >= 2^24) or an arg-max tie must be broken by BFS order
  std::vector<int> need;
  unsigned max_need = 0;
  for (int i = 0; i < nc; ++i)
    if (h_rows[i].flags & (kFlagNeedsOrder | kFlagTie)) {
      need.push_back(i);
      max_need = std::max(max_need, (unsigned)c->h_clusters[i].size);
    }
  if (!need.empty() && !getenv("AOS_LITERAL_BFS")) {
    const size_t nn = (size_t)n, total_jobs = need.size();
    // link int2[n] | state uint2[2n] (later cinfo int4[n]) | runs int[3n] | irr IrrRec[n/4] | flagged, fallback int[jobs] | ctr
    const size_t jobs4 = (total_jobs + 3) & ~(size_t)3;
    const size_t run_cap = 3 * nn + 64, irr_cap = nn / 4 + 1024;
    AOS_CUDA_OK(c, c->bfs_buf.reserve(nn * (8 + 16) + run_cap * 4 + irr_cap * sizeof(IrrRec) + (2 * jobs4 + 64) * 4 + 16 * 8));
    BfsBufs B;
    char *p = c->bfs_buf.as<char>();
    auto carve = [&p](size_t bytes) {
      char *r = p;
      p += (bytes + 15) & ~(size_t)15;
      return r;
    };
    B.link = reinterpret_cast<int2 *>(carve(nn * 8));
    B.state = reinterpret_cast<uint2 *>(carve(nn * 16));
    B.cinfo = reinterpret_cast<int4 *>(B.state);
    B.runs = reinterpret_cast<int *>(carve(run_cap * 4));
    B.irr = reinterpret_cast<IrrRec *>(carve(irr_cap * sizeof(IrrRec)));
    int *d_flagged = reinterpret_cast<int *>(carve(jobs4 * 4));
    int *d_fallback = reinterpret_cast<int *>(carve(jobs4 * 4));
    B.ctr = reinterpret_cast<int *>(carve(64 * 4));
    B.run_cap = (int)std::min<size_t>(run_cap, 0x7fffffff);
    B.irr_cap = (int)std::min<size_t>(irr_cap, 0x7fffffff);
    if (!c->pin_a.resize(sizeof(int) * jobs4)) {  // pin_a is free again: h_clusters holds its copy
      set_error(c, "cudaHostAlloc failed (replay jobs)");
      return AOS_ERR_CUDA;
    }
    // longest first: CTAs are dispatched in index order
    std::sort(need.begin(), need.end(), [&](int a, int b) {
      return c->h_clusters[a].size != c->h_clusters[b].size ? c->h_clusters[a].size > c->h_clusters[b].size : a < b;
    });
    memcpy(c->pin_a.data(), need.data(), sizeof(int) * total_jobs);
    s = h2d_small(c, d_flagged, c->pin_a.data(), sizeof(int) * jobs4, true);
    if (s != AOS_OK) return s;
    AOS_CUDA_OK(c, cudaMemsetAsync(B.ctr, 0, 64 * 4, st));
    const int gb = grid_for(nn, 256);
    bfs_prepare_kernel<<<gb, 256, 0, st>>>(mask, prefix, P.pitch, P.h, P.w, cellpos, cell_cluster, root_cellpos, d_rows, n, B);
    bfs_link_init_kernel<<<gb, 256, 0, st>>>(n, B);
    c->launches += 2;
    // a run is shorter than its cluster; a round of four hops covers at least five times the distance of the one before
    // (bfs_runs_kernel checks that every entry arrived and hands the map to the literal replay otherwise)
    int rounds = 1;
    for (unsigned long long reach = 5; reach < max_need; reach *= 5) ++rounds;
    rounds += 2;
    bfs_jump_first_kernel<<<grid_for(2 * nn, 256), 256, 0, st>>>(n, B);
    for (int r = 0; r < rounds; ++r) bfs_jump_kernel<<<4 * kNumSMs, 256, 0, st>>>(B, r);
    bfs_runs_kernel<<<gb, 256, 0, st>>>(n, B);
    bfs_scatter_kernel<<<gb, 256, 0, st>>>(n, cellpos, B);
    bfs_irr_kernel<<<gb, 256, 0, st>>>(n, B);
    c->launches += rounds + 4;
    AOS_CUDA_OK(c, cudaGetLastError());
    c->mark("replay_prep");
    int item_cap = kItemCap;  // AOS_BFS_ITEM_CAP (tests): a narrower level list, so that some clusters take the fallback
    if (const char *e = getenv("AOS_BFS_ITEM_CAP")) item_cap = std::max(1, std::min(kItemCap, atoi(e)));
    // the walk stages a cluster's tables in shared memory (about 6 bytes per cell of a thinned row: 6 % irregular cells of 72
    // bytes, as many run headers of 24): launches by size class, longest rows (sorted to the front) first; a cluster that needs
    // more than its class provides walks on the global tables.  AOS_BFS_GLOBAL=1 (tests) switches the staging off.
    const bool staged = getenv("AOS_BFS_GLOBAL") == nullptr;
    const unsigned class_cells[3] = {16000u, 6000u, 1500u};  // clusters larger than this many cells ...
    const int class_budget[3] = {160 << 10, 96 << 10, 48 << 10};  // ... get this much; the rest 16 KB
    AOS_CUDA_OK(c, cudaFuncSetAttribute(bfs_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, class_budget[0]));
    // the classes are independent launches: the few longest rows would otherwise serialise in front of the many short ones
    AOS_CUDA_OK(c, cudaEventRecord(c->ev_fork, st));
    bool forked[3] = {false, false, false};
    size_t first = 0;
    for (int k = 0; k < 4; ++k) {
      size_t last = first;
      while (last < total_jobs && (k == 3 || c->h_clusters[need[last]].size > (long long)class_cells[k])) ++last;
      if (last > first) {
        const int budget = staged ? (k == 3 ? 16 << 10 : class_budget[k]) : 0;
        cudaStream_t ls = st;
        if (k < 3) {
          ls = c->aux[k];
          AOS_CUDA_OK(c, cudaStreamWaitEvent(ls, c->ev_fork, 0));
          forked[k] = true;
        }
        bfs_chain_kernel<<<(unsigned)(last - first), k < 2 ? 1024 : kChainThreads, budget, ls>>>(
            d_flagged + first, acc, offsets, grouped_id, root_cellpos, mask, prefix, P.pitch, P.w, B, item_cap, budget,
            d_fallback + first, queue, centre);
        ++c->launches;
      }
      first = last;
    }
    for (int k = 0; k < 3; ++k)
      if (forked[k]) {
        AOS_CUDA_OK(c, cudaEventRecord(c->ev_join[k], c->aux[k]));
        AOS_CUDA_OK(c, cudaStreamWaitEvent(st, c->ev_join[k], 0));
      }
    AOS_CUDA_OK(c, cudaGetLastError());
    c->mark("replay_bfs");
    cluster_finalize_kernel<<<(unsigned)total_jobs, kClThreads, 0, st>>>(P, acc, offsets, grouped, root_cellpos, min_length, d_flagged,
                                                                         d_fallback, queue, centre, d_clusters, d_rows);
    ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, B.ctr, 16, cudaMemcpyDeviceToHost, st));
    s = fetch_tables();  // one synchronisation for the tables and the fallback counter
    if (s != AOS_OK) return s;
    c->bfs_fallbacks = c->h_flag[1];
    if (getenv("AOS_DEBUG"))
      fprintf(stderr, "[aos] BFS order: %zu clusters (longest %u cells), %d jump rounds, %d run ints, %d irregular cells, %d left to the literal replay%s\n",
              total_jobs, max_need, rounds, c->h_flag[0], c->h_flag[3], c->h_flag[1], c->h_flag[2] ? " (ranking broken / out of space)" : "");
    // clusters the warp gave up on (a level wider than its list) kept their flags: the literal replay below takes them
    need.clear();
    if (c->h_flag[1] > 0)
      for (int i = 0; i < nc; ++i)
        if (h_rows[i].flags & (kFlagNeedsOrder | kFlagTie)) need.push_back(i);
  }
  if (!need.empty()) {
    if (!c->pin_c.resize(sizeof(ClusterAcc) * (size_t)nc)) {
      set_error(c, "cudaHostAlloc failed (cluster accumulators)");
      return AOS_ERR_CUDA;
    }
    const ClusterAcc *h_acc = reinterpret_cast<const ClusterAcc *>(c->pin_c.data());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->pin_c.data(), acc, sizeof(ClusterAcc) * (size_t)nc, cudaMemcpyDeviceToHost, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    // shared-memory classes of the banded replay: ring + column tables + band bitmap (at most one word per cell for a
    // thin connected curve) <= 4 K words (16 KB), <= 12 K (48 KB), <= 54 K (216 KB); larger ones go straight to the
    // global-bitmap kernel, as does any cluster whose band turns out not to fit its class
    const size_t class_words[3] = {4096, 12288, 55296};
    std::vector<ReplayJob> jobs[4];
    for (int i : need) {
      const ClusterAcc &a = h_acc[i];
      auto px = [&](int k) { int pos = (int)(a.ext[k] & 0xffffffffull); return pos % P.w; };
      auto py = [&](int k) { int pos = (int)(a.ext[k] & 0xffffffffull); return pos / P.w; };
      ReplayJob j;
      j.cluster = i;
      int x0 = px(4), x1 = px(0), y0 = py(5), y1 = py(1);
      j.wx0 = x0 >> 5;
      j.ww = (x1 >> 5) - j.wx0 + 1;
      j.y0 = y0;
      j.hh = y1 - y0 + 1;
      size_t band = std::min((size_t)j.ww * j.hh, (size_t)a.size);
      size_t words = (size_t)kRingN + 2 * (size_t)j.ww + 1 + band;
      int cls = words <= class_words[0] ? 0 : words <= class_words[1] ? 1 : words <= class_words[2] ? 2 : 3;
      if (j.ww >= 32768 || j.hh >= 4096) {
        set_error(c, "cluster bounding box exceeds the replay packing limits");
        return AOS_ERR_CAPACITY;
      }
      jobs[cls].push_back(j);
    }
    for (int k = 0; k < 4; ++k)  // longest first: CTAs are dispatched in index order
      std::sort(jobs[k].begin(), jobs[k].end(), [&](const ReplayJob &a, const ReplayJob &b) {
        return h_acc[a.cluster].size > h_acc[b.cluster].size;
      });
    size_t total_jobs = need.size();
    if (getenv("AOS_DEBUG")) {
      size_t mx[4] = {0, 0, 0, 0};
      for (int k = 0; k < 4; ++k)
        for (const ReplayJob &j : jobs[k]) mx[k] = std::max(mx[k], (size_t)h_acc[j.cluster].size);
      fprintf(stderr, "[aos] literal replay classes: %zu/%zu/%zu/%zu jobs, max cells %zu/%zu/%zu/%zu\n", jobs[0].size(),
              jobs[1].size(), jobs[2].size(), jobs[3].size(), mx[0], mx[1], mx[2], mx[3]);
    }
    AOS_CUDA_OK(c, c->cl_stats.reserve((sizeof(ReplayJob) + 2 * sizeof(int)) * total_jobs));
    ReplayJob *d_jobs = c->cl_stats.as<ReplayJob>();
    int *d_flagged = reinterpret_cast<int *>(d_jobs + total_jobs);
    int *d_fallback = d_flagged + total_jobs;
    // jobs | cluster ids | fallback flags, staged in page-locked memory (pin_a is free again: h_clusters holds its copy)
    if (!c->pin_a.resize((sizeof(ReplayJob) + 2 * sizeof(int)) * total_jobs)) {
      set_error(c, "cudaHostAlloc failed (replay jobs)");
      return AOS_ERR_CUDA;
    }
    ReplayJob *all = reinterpret_cast<ReplayJob *>(c->pin_a.data());
    int *order = reinterpret_cast<int *>(all + total_jobs);
    int *fb = order + total_jobs;
    {
      size_t q = 0;
      for (int k = 0; k < 4; ++k)
        for (const ReplayJob &j : jobs[k]) {
          all[q] = j;
          fb[q] = k == 3;
          order[q++] = j.cluster;
        }
    }
    s = h2d_small(c, d_jobs, all, (sizeof(ReplayJob) + 2 * sizeof(int)) * total_jobs, true);  // contiguous on both sides
    if (s != AOS_OK) return s;
    uint32_t *gvisited = prefix;  // compact indices are no longer needed: reuse as the global "unvisited" bitmap
    AOS_CUDA_OK(c, cudaMemcpyAsync(gvisited, mask, words * 4, cudaMemcpyDeviceToDevice, st));
    c->mark("literal_replay_prep");
    // the size classes are independent launches: class 0 stays on the context stream, the others fork onto side
    // streams so that a few very long rows do not serialise behind the many short ones
    AOS_CUDA_OK(c, cudaEventRecord(c->ev_fork, st));
    size_t done = 0;
    int side = 0;
    bool forked[3] = {false, false, false};
    AOS_CUDA_OK(c, cudaFuncSetAttribute(bfs_replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(class_words[2] * 4)));
    for (int k = 0; k < 3; ++k) {
      if (jobs[k].empty()) continue;
      cudaStream_t ls = st;
      if (k > 0 && side < 3) {
        ls = c->aux[side];
        AOS_CUDA_OK(c, cudaStreamWaitEvent(ls, c->ev_fork, 0));
        forked[side++] = true;
      }
      bfs_replay_kernel<true><<<(unsigned)jobs[k].size(), 32, class_words[k] * 4, ls>>>(
          P, d_jobs + done, acc, offsets, grouped, root_cellpos, (int)class_words[k], nullptr, d_fallback + done, queue, centre);
      ++c->launches;
      done += jobs[k].size();
    }
    AOS_CUDA_OK(c, cudaGetLastError());
    for (int k = 0; k < 3; ++k)
      if (forked[k]) {
        AOS_CUDA_OK(c, cudaEventRecord(c->ev_join[k], c->aux[k]));
        AOS_CUDA_OK(c, cudaStreamWaitEvent(st, c->ev_join[k], 0));
      }
    // whatever did not fit (and class 3): the warp-per-cluster kernel on the global bitmap; everyone else exits at once
    bfs_replay_kernel<false><<<(unsigned)total_jobs, 32, kRingN * 4, st>>>(P, d_jobs, acc, offsets, grouped, root_cellpos, 0,
                                                                           gvisited, d_fallback, queue, centre);
    ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    c->mark("literal_replay_bfs");
    cluster_finalize_kernel<<<(unsigned)total_jobs, kClThreads, 0, st>>>(P, acc, offsets, grouped, root_cellpos, min_length,
                                                                         d_flagged, nullptr, queue, centre, d_clusters, d_rows);
    ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    s = fetch_tables();  // after the replay kernels consumed the job list: the copy engine has read pin_a by now
    if (s != AOS_OK) return s;
  }
  c->mark("replay_finalize");
  for (int i = 0; i < nc; ++i)
    if (h_rows[i].valid) c->h_rows.push_back(h_rows[i].row);
  return AOS_OK;
}

// per-cell canonical label map: -1 everywhere, min linear index of the component at clustered cells
__global__ void labels_kernel(const int *__restrict__ cell_cluster, const int *__restrict__ cellpos,
                              const int *__restrict__ root_cellpos, int n, int32_t *__restrict__ labels) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    labels[cellpos[i]] = root_cellpos[cell_cluster[i]];
}

aos_status launch_labels(Ctx *c, int32_t *dst) {
  size_t cells = (size_t)c->P.w * c->P.h;
  AOS_CUDA_OK(c, cudaMemsetAsync(dst, 0xff, cells * 4, c->stream));
  if (c->n_skel_cells > 0) {
    labels_kernel<<<grid_for((size_t)c->n_skel_cells, 256), 256, 0, c->stream>>>(
        c->d_cell_cluster, c->cc_cellpos.as<int>(), c->d_root_cellpos, c->n_skel_cells, dst);
  ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
  }
  return AOS_OK;
}

}  // namespace aos
