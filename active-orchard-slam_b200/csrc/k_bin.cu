// k_bin.cu -- point binning: PassThrough z/x/y + exclusion discs + generateOccupancyGrid scatter,
// fused into one pass over the cloud (reference: src/aos_seed_gen_node.cpp:459-538 filters,
// :581-622 scatter).  HBM-bound: 16 B per point read once, bits OR-ed into the L2-resident grid.
#include "aos_common.cuh"

namespace aos {

__device__ __forceinline__ bool keep_point(const SeedDeviceParams &P, float x, float y, float z) {
  // pcl::PassThrough: non-finite removed, limits inclusive, float compares (seed_gen:459-477)
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) return false;
  if (z < P.minz || z > P.maxz) return false;
  if (x < P.minx || x > P.maxx) return false;
  if (y < P.miny || y > P.maxy) return false;
  // exclusion discs, float32 arithmetic without FMA contraction (seed_gen:504-517)
  for (int e = 0; e < P.n_excl; ++e) {
    float dx = __fsub_rn(x, P.excl[3 * e]);
    float dy = __fsub_rn(y, P.excl[3 * e + 1]);
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    float r = P.excl[3 * e + 2];
    if (d2 <= __fmul_rn(r, r)) return false;
  }
  return true;
}

// seed_gen:609-610: (float - double) / float evaluated in double, truncated toward zero.  The quotient is
// formed as a multiplication by the rounded reciprocal (error below 1e-11 cells at 4e4 cells) and redone as the
// reference's exact division only when it lands within 1e-7 of an integer, so the truncation is always the
// reference's.
__device__ __forceinline__ int cell_index(double q, double res, double inv_res) {
  double t = q * inv_res;
  if (fabs(t - rint(t)) < 1e-7) t = q / res;
  return (int)t;
}

__device__ __forceinline__ void scatter_point(const SeedDeviceParams &P, double inv_res, float x, float y, uint32_t *bits) {
  int gx = cell_index((double)x - P.ox, (double)P.res, inv_res);
  int gy = cell_index((double)y - P.oy, (double)P.res, inv_res);
  if (gx >= 0 && gx < P.w && gy >= 0 && gy < P.h) {
    uint32_t *wp = bits + (size_t)gy * P.pitch + (gx >> 5);
    uint32_t m = 1u << (gx & 31);
    // most hits land on cells that are already set: test first, RED only when needed
    if (!(__ldcg(wp) & m)) atomicOr(wp, m);
  }
}

constexpr int kBinThreads = 256;
constexpr int kBinUnroll = 8;

// Fast path: PointXYZ layout (16-byte records, x y z pad), 16-byte aligned base.  A warp streams 256
// consecutive points per step (8 coalesced 512-byte loads in flight per warp), filters them in registers and
// packs the survivors (~1 in 5 passes the z window) into a per-warp shared-memory queue, so the index
// arithmetic and the atomics run on dense lanes instead of 8 sparsely populated passes.
__global__ void __launch_bounds__(kBinThreads) bin_points_xyz16(const __grid_constant__ SeedDeviceParams P,
                                                                const float4 *__restrict__ pts, size_t n,
                                                                uint32_t *__restrict__ bits,
                                                                unsigned long long *__restrict__ n_kept) {
  __shared__ float2 queue[kBinThreads / 32][kBinUnroll * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const double inv_res = 1.0 / (double)P.res;
  float2 *q = queue[warp];
  unsigned long long kept = 0;
  const size_t warps_total = (size_t)gridDim.x * (kBinThreads / 32);
  const size_t per_warp = (size_t)kBinUnroll * 32;
  for (size_t base = ((size_t)blockIdx.x * (kBinThreads / 32) + warp) * per_warp; base < n; base += warps_total * per_warp) {
    float4 v[kBinUnroll];
#pragma unroll
    for (int u = 0; u < kBinUnroll; ++u) {
      size_t i = base + (size_t)u * 32 + lane;
      if (i < n) v[u] = ld_stream_f4(pts + i);
      else v[u] = make_float4(0.f, 0.f, __int_as_float(0x7fc00000), 0.f);  // NaN z -> dropped
    }
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kBinUnroll; ++u) {
      const bool k = keep_point(P, v[u].x, v[u].y, v[u].z);
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (k) q[cnt + __popc(m & lt)] = make_float2(v[u].x, v[u].y);
      cnt += __popc(m);
    }
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) scatter_point(P, inv_res, q[i].x, q[i].y, bits);
    __syncwarp();
    kept += (unsigned)cnt;
  }
  if (lane == 0 && kept) atomicAdd(n_kept, kept);  // diagnostic count
}

// General path: arbitrary point_step / field offsets (e.g. LIO-SAM's 32-byte XYZI records).
__global__ void __launch_bounds__(kBinThreads) bin_points_generic(const __grid_constant__ SeedDeviceParams P,
                                                                  const uint8_t *__restrict__ pts, size_t n,
                                                                  uint32_t step, uint32_t offx, uint32_t offy,
                                                                  uint32_t offz, uint32_t *__restrict__ bits,
                                                                  unsigned long long *__restrict__ n_kept) {
  unsigned kept = 0;
  const double inv_res = 1.0 / (double)P.res;
  const bool aligned4 = ((step | offx | offy | offz) & 3u) == 0 && (((uintptr_t)pts) & 3u) == 0;
  for (size_t i = (size_t)blockIdx.x * kBinThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kBinThreads) {
    const uint8_t *rec = pts + i * step;
    float x, y, z;
    if (aligned4) {
      x = __ldg(reinterpret_cast<const float *>(rec + offx));
      y = __ldg(reinterpret_cast<const float *>(rec + offy));
      z = __ldg(reinterpret_cast<const float *>(rec + offz));
    } else {
      uint32_t a = 0, b = 0, c = 0;
      for (int k = 3; k >= 0; --k) {
        a = (a << 8) | rec[offx + k];
        b = (b << 8) | rec[offy + k];
        c = (c << 8) | rec[offz + k];
      }
      x = __uint_as_float(a);
      y = __uint_as_float(b);
      z = __uint_as_float(c);
    }
    if (keep_point(P, x, y, z)) {
      ++kept;
      scatter_point(P, inv_res, x, y, bits);
    }
  }
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_kept, (unsigned long long)kept);
}

aos_status launch_bin(Ctx *c, const SeedDeviceParams &P, const void *points, size_t n, uint32_t step, uint32_t offx,
                      uint32_t offy, uint32_t offz, uint32_t *bits, unsigned long long *n_kept) {
  if (n == 0) return AOS_OK;
  const bool fast = step == 16 && offx == 0 && offy == 4 && offz == 8 && (((uintptr_t)points) & 15u) == 0;
  if (fast) {
    size_t per_block = (size_t)kBinThreads * kBinUnroll;
    size_t want = (n + per_block - 1) / per_block;
    int grid = (int)(want < (size_t)kNumSMs * 8 ? want : (size_t)kNumSMs * 8);
    bin_points_xyz16<<<grid, kBinThreads, 0, c->stream>>>(P, reinterpret_cast<const float4 *>(points), n, bits, n_kept);
  ++c->launches;
  } else {
    size_t want = (n + kBinThreads - 1) / kBinThreads;
    int grid = (int)(want < (size_t)kNumSMs * 16 ? want : (size_t)kNumSMs * 16);
    bin_points_generic<<<grid, kBinThreads, 0, c->stream>>>(P, reinterpret_cast<const uint8_t *>(points), n, step, offx,
                                                            offy, offz, bits, n_kept);
  ++c->launches;
  }
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

}  // namespace aos
