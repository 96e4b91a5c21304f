// k_ror.cu -- pcl::RadiusOutlierRemoval ahead of the seam: globalMapCallback
// (src/aos_seed_gen_node.cpp:229-248) keeps a point of the global map iff at least `min_neighbors` OTHER points lie
// within `radius` in 3-D before it hands the cloud to processPointCloud.
//
// PCL is not in the reference tree nor in this image; the rule below restates the dense-cloud branch of
// pcl::RadiusOutlierRemoval::applyFilterIndices (PCL 1.12: k-NN with k = min_neighbors + 1, the query point counts as
// its own nearest neighbour, a point is kept iff the k-th squared distance <= radius * radius), with FLANN's
// L2_Simple<float> distance, i.e. ((dx*dx + dy*dy) + dz*dz) accumulated in float32 and compared in double.
// PARITY UNPINNED against the real PCL; pinned only against the brute-force restatement in oracle/oracle.py.
//
// The kd-tree becomes a hash of radius-sized voxels (64-bit cell key -> linked list of point indices): a point's
// neighbours lie in its 27 surrounding voxels, the walk stops as soon as enough neighbours are seen -- in a tree
// crown that is the point's own voxel.  Survivors are compacted in input order (prefix sum) into PointXYZ records.
#include <math.h>

#include <algorithm>

#include "aos_common.cuh"
#include "dev_hash.cuh"

namespace aos {

struct RorPoints {
  const uint8_t *base;
  uint32_t step, ox, oy, oz;
  bool fast;  // 16-byte records x y z pad, aligned
};

__device__ __forceinline__ float3 ror_load(const RorPoints &p, size_t i) {
  if (p.fast) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(p.base) + i);
    return make_float3(v.x, v.y, v.z);
  }
  const uint8_t *rec = p.base + i * p.step;
  uint32_t a = 0, b = 0, c = 0;
  for (int k = 3; k >= 0; --k) {
    a = (a << 8) | rec[p.ox + k];
    b = (b << 8) | rec[p.oy + k];
    c = (c << 8) | rec[p.oz + k];
  }
  return make_float3(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c));
}

__device__ __forceinline__ long long ror_cell(float v, float inv) {
  float f = floorf(v * inv);
  f = fminf(fmaxf(f, -1000000.f), 1000000.f);  // 21 bits per axis
  return (long long)f;
}
__device__ __forceinline__ unsigned long long ror_key(long long cx, long long cy, long long cz) {
  return ((unsigned long long)(cx + (1 << 20)) << 42) | ((unsigned long long)(cy + (1 << 20)) << 21) |
         (unsigned long long)(cz + (1 << 20));
}

__global__ void ror_build_kernel(RorPoints pts, size_t n, float inv_cell, DevHash h, int *__restrict__ next) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float3 p = ror_load(pts, i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      next[i] = -2;  // never a neighbour, never kept
      continue;
    }
    int slot = hash_insert(h, ror_key(ror_cell(p.x, inv_cell), ror_cell(p.y, inv_cell), ror_cell(p.z, inv_cell)));
    next[i] = atomicExch(&h.val[slot], (int)i);
  }
}

__global__ void ror_count_kernel(RorPoints pts, size_t n, float inv_cell, double r2, int min_neighbors, DevHash h,
                                 const int *__restrict__ next, uint32_t *__restrict__ keep) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (next[i] == -2) {
      keep[i] = 0;
      continue;
    }
    const float3 p = ror_load(pts, i);
    const long long cx = ror_cell(p.x, inv_cell), cy = ror_cell(p.y, inv_cell), cz = ror_cell(p.z, inv_cell);
    int found = 0;
    // own voxel first: in dense regions it already holds enough neighbours
    for (int k = 0; k < 27 && found < min_neighbors; ++k) {
      const int o = k == 0 ? 13 : (k <= 13 ? k - 1 : k);  // 13 = (0,0,0)
      const int dz = o / 9 - 1, dy = (o / 3) % 3 - 1, dx = o % 3 - 1;
      const int slot = hash_find(h, ror_key(cx + dx, cy + dy, cz + dz));
      if (slot < 0) continue;
      for (int j = h.val[slot]; j >= 0 && found < min_neighbors; j = next[j]) {
        if ((size_t)j == i) continue;
        const float3 q = ror_load(pts, (size_t)j);
        const float ex = __fsub_rn(p.x, q.x), ey = __fsub_rn(p.y, q.y), ez = __fsub_rn(p.z, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        if ((double)d2 <= r2) ++found;
      }
    }
    keep[i] = found >= min_neighbors ? 1u : 0u;
  }
}

__global__ void ror_emit_kernel(RorPoints pts, size_t n, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos,
                                float4 *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!keep[i]) continue;
    const float3 p = ror_load(pts, i);
    out[pos[i]] = make_float4(p.x, p.y, p.z, 1.0f);
  }
}

// keep flags live in `pos` after the scan as positions; `keep` is a copy made before it
aos_status run_ror(Ctx *c, const void *dpoints, size_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, float radius,
                   int min_neighbors, size_t *n_out) {
  cudaStream_t st = c->stream;
  *n_out = 0;
  if (n == 0) return AOS_OK;
  AOS_REQUIRE(c, n < (size_t)2000000000, "more than 2e9 points");
  AOS_REQUIRE(c, radius > 0.f && std::isfinite(radius) && min_neighbors >= 0, "bad radius / min_neighbors");
  size_t cap = 1024;
  while (cap < n * 2) cap <<= 1;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_keys = up(cap * 8), b_val = up(cap * 4), b_next = up(n * 4), b_flag = up((n + 1) * 4);
  AOS_CUDA_OK(c, c->ror_buf.reserve(b_keys + b_val + b_next + 2 * b_flag + 1024));
  AOS_CUDA_OK(c, c->ror_out.reserve(n * sizeof(float4)));
  char *base = c->ror_buf.as<char>();
  DevHash h;
  h.keys = reinterpret_cast<unsigned long long *>(base);
  h.val = reinterpret_cast<int *>(base + b_keys);
  h.mask = (unsigned)(cap - 1);
  int *next = reinterpret_cast<int *>(base + b_keys + b_val);
  uint32_t *keep = reinterpret_cast<uint32_t *>(base + b_keys + b_val + b_next);
  uint32_t *pos = reinterpret_cast<uint32_t *>(base + b_keys + b_val + b_next + b_flag);
  uint32_t *d_tot = pos + n;  // last element of the second flag plane (n + 1 entries)
  AOS_CUDA_OK(c, cudaMemsetAsync(h.keys, 0xff, cap * 8, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(h.val, 0xff, cap * 4, st));
  RorPoints P{static_cast<const uint8_t *>(dpoints), step, ox, oy, oz,
              step == 16 && ox == 0 && oy == 4 && oz == 8 && (((uintptr_t)dpoints) & 15u) == 0};
  const float inv_cell = 1.0f / (radius * 1.0001f);  // voxels a hair larger than the radius: float rounding of the cell index can never hide a neighbour
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 32);
  ror_build_kernel<<<grid, 256, 0, st>>>(P, n, inv_cell, h, next);
  ++c->launches;
  c->mark("ror_build");
  ror_count_kernel<<<grid, 256, 0, st>>>(P, n, inv_cell, (double)radius * (double)radius, min_neighbors, h, next, keep);
  ++c->launches;
  c->mark("ror_count");
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(pos, keep, n * 4, cudaMemcpyDeviceToDevice, st));
  aos_status s = exclusive_scan_u32(c, pos, n, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  ror_emit_kernel<<<grid, 256, 0, st>>>(P, n, keep, pos, c->ror_out.as<float4>());
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_tot, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->mark("ror_compact");
  *n_out = (size_t)(unsigned)c->h_flag[0];
  return AOS_OK;
}

}  // namespace aos
