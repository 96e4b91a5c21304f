// k_bin.cu -- point binning: PassThrough z/x/y + exclusion discs + generateOccupancyGrid scatter,
// fused into one pass over the cloud (reference: src/aos_seed_gen_node.cpp:459-538 filters,
// :581-622 scatter).  HBM-bound: 16 B per point read once, bits OR-ed into the L2-resident grid.
#include <stdlib.h>

#include "aos_common.cuh"

namespace aos {

__device__ __forceinline__ bool keep_point(const SeedDeviceParams &P, float x, float y, float z) {
  // pcl::PassThrough: non-finite removed, limits inclusive, float compares (seed_gen:459-477).  With finite
  // limits the six inclusive compares already reject NaN and +-inf; infinite limits add the explicit test.
  if (!(z >= P.minz && z <= P.maxz && x >= P.minx && x <= P.maxx && y >= P.miny && y <= P.maxy)) return false;
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) return false;
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
  if (gx >= 0 && gx < P.w && gy >= 0 && gy < P.gh) {
    gy -= P.y_off;  // row-band mode: rows outside this band's local grid belong to another GPU
    if (gy < 0 || gy >= P.h) return;
    uint32_t *wp = bits + (size_t)gy * P.pitch + (gx >> 5);
    uint32_t m = 1u << (gx & 31);
    // fire-and-forget reduction at the L2 (SASS RED.OR): no dependent read on the warp's critical path; with a
    // shuffled cloud almost every kept point lands in a different 32-byte sector anyway
    atomicOr(wp, m);
  }
}

constexpr int kBinThreads = 256;       // generic path
constexpr int kBinStagePts = 1024;     // points per pipeline stage (16 KB)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Fast path: PointXYZ layout (16-byte records, x y z pad), 16-byte aligned base.  Persistent CTAs, three per SM.
// A producer warp streams the cloud into a 4-stage shared-memory ring with bulk async copies (192 KB in flight
// per SM, independent of register pressure); eight consumer warps filter the staged points in registers, pack
// the survivors (~1 in 5 passes the z window) into a per-warp queue so that the double-precision index
// arithmetic and the atomics run on dense lanes, and hand the stage back.  HBM traffic: 16 B per point, once.
template <int kBinStages, int kBinConsumers>
__global__ void __launch_bounds__((kBinConsumers + 1) * 32) bin_points_xyz16(const __grid_constant__ SeedDeviceParams P,
                                                                   const float4 *__restrict__ pts, size_t n,
                                                                   uint32_t *__restrict__ bits,
                                                                   unsigned long long *__restrict__ n_kept) {
  constexpr int kBinPerWarp = kBinStagePts / kBinConsumers;  // points per consumer warp and stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *stages = reinterpret_cast<float4 *>(smem_raw);
  float2 *queues = reinterpret_cast<float2 *>(smem_raw + (size_t)kBinStages * kBinStagePts * sizeof(float4));
  __shared__ __align__(8) uint64_t full_bar[kBinStages], empty_bar[kBinStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBinStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kBinConsumers);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const size_t n_chunks = (n + kBinStagePts - 1) / kBinStagePts;

  if (warp == kBinConsumers) {  // ---- producer ----
    if (lane == 0) {
      int it = 0;
      for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const int s = it % kBinStages;
        if (it >= kBinStages) mbar_wait(&empty_bar[s], ((it / kBinStages) - 1) & 1);
        const size_t first = c * kBinStagePts;
        const uint32_t bytes = (uint32_t)((n - first < (size_t)kBinStagePts ? n - first : (size_t)kBinStagePts) * sizeof(float4));
        mbar_expect_tx(&full_bar[s], bytes);
        bulk_load_1d(stages + (size_t)s * kBinStagePts, pts + first, bytes, &full_bar[s]);
      }
    }
    return;
  }

  // ---- consumers ----
  const unsigned lt = (1u << lane) - 1u;
  const double inv_res = 1.0 / (double)P.res;
  float2 *q = queues + warp * kBinPerWarp;
  unsigned long long kept = 0;
  int it = 0;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
    const int s = it % kBinStages;
    mbar_wait(&full_bar[s], (it / kBinStages) & 1);
    const size_t first = c * kBinStagePts;
    const int npts = (int)(n - first < (size_t)kBinStagePts ? n - first : (size_t)kBinStagePts);
    const float4 *st = stages + (size_t)s * kBinStagePts + warp * kBinPerWarp;
    const int mine = max(0, min(kBinPerWarp, npts - warp * kBinPerWarp));
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kBinPerWarp / 32; ++u) {
      const int i = u * 32 + lane;
      bool k = false;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < mine) {
        v = st[i];
        k = keep_point(P, v.x, v.y, v.z);
      }
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (k) q[cnt + __popc(m & lt)] = make_float2(v.x, v.y);
      cnt += __popc(m);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);  // every lane has its points in registers / in the queue
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
    size_t want = (n + kBinStagePts - 1) / kBinStagePts;
    // 4 stages x 16 KB + queues = 72 KB per CTA, three CTAs per SM: 192 KB of point data in flight per SM and 24
    // consumer warps.  Measured on B200 at 200 M points (GB/s of point data): 6 stages/2 CTAs 4680, 4 stages/3 CTAs
    // 5410, 3 stages/4 CTAs 4750, 16 consumer warps/2 CTAs 4370.
    constexpr int kStages = 4, kConsumers = 8, kPerSM = 3;
    int grid = (int)(want < (size_t)kNumSMs * kPerSM ? want : (size_t)kNumSMs * kPerSM);
    size_t smem = (size_t)kStages * kBinStagePts * sizeof(float4) + (size_t)kBinStagePts * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(bin_points_xyz16<kStages, kConsumers>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      bin_points_xyz16<kStages, kConsumers><<<grid, (kConsumers + 1) * 32, smem, c->stream>>>(
          P, reinterpret_cast<const float4 *>(points), n, bits, n_kept);
      e = cudaGetLastError();
    }
    AOS_CUDA_OK(c, e);
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
