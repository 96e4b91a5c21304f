// k_misc.cu -- ABI-edge conversions (occupancyGridToMat / matToOccupancyGrid, src/aos_seed_gen_node.cpp:
// 626-669: 100 <-> set bit, anything else is free), the 1-px rectangle of markPolygonBoundaryAsOccupied
// (:772-825; its four Bresenham lines are axis-aligned), TMA descriptor creation and error plumbing.
#include <stdarg.h>

#include <algorithm>

#include "aos_common.cuh"

namespace aos {

void set_error(Ctx *c, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

bool make_bitgrid_tmap(CUtensorMap *map, const uint32_t *base, int pitch_words, int rows, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)pitch_words, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_words * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// ---- int8 {0,100} -> bits: one warp packs 32 words from 1024 bytes with coalesced 4-byte loads -----
__global__ void pack_kernel(const int8_t *__restrict__ src, uint32_t *__restrict__ dst, int w, int h, int pitch) {
  int wordcol = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (wordcol >= pitch) return;
  uint32_t v = 0;
  int x0 = wordcol << 5;
  const int8_t *row = src + (size_t)y * w;
  for (int b = 0; b < 32; ++b) {
    int x = x0 + b;
    if (x < w && row[x] == 100) v |= 1u << b;
  }
  dst[(size_t)y * pitch + wordcol] = v;
}

// ---- bits -> int8 {0,100}: thread per 4 cells, 4-byte stores when the row is 4-aligned -------------
__global__ void unpack_kernel(const uint32_t *__restrict__ src, int8_t *__restrict__ dst, int w, int h, int pitch) {
  size_t total = (size_t)w * h;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int y = (int)(i / (size_t)w), x = (int)(i - (size_t)y * w);
    uint32_t word = __ldg(src + (size_t)y * pitch + (x >> 5));
    dst[i] = ((word >> (x & 31)) & 1u) ? 100 : 0;
  }
}

__global__ void unpack4_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst4, int w4, int h, int pitch) {
  // w is a multiple of 4 and dst is 4-byte aligned: every thread emits 4 cells as one 32-bit store
  size_t total = (size_t)w4 * h;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int y = (int)(i / (size_t)w4), q = (int)(i - (size_t)y * w4);
    int x = q << 2;
    uint32_t nib = (__ldg(src + (size_t)y * pitch + (x >> 5)) >> (x & 31)) & 0xfu;
    uint32_t v = ((nib & 1u) ? 100u : 0u) | ((nib & 2u) ? 100u << 8 : 0u) | ((nib & 4u) ? 100u << 16 : 0u) |
                 ((nib & 8u) ? 100u << 24 : 0u);
    dst4[i] = v;
  }
}

// bits of the word starting at cell x0 that fall in the cell range [lo, hi]
__device__ __forceinline__ uint32_t range_mask(int lo, int hi, int x0) {
  lo = max(lo - x0, 0);
  hi = min(hi - x0, 31);
  if (lo > hi) return 0u;
  return (hi - lo == 31) ? 0xffffffffu : (((1u << (hi - lo + 1)) - 1u) << lo);
}

// ring of thickness t along the rectangle [gx0,gx1] x [gy0,gy1] OR-ed into the grid
__global__ void frame_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int w, int h, int pitch,
                             int gx0, int gy0, int gx1, int gy1, int t) {
  size_t total = (size_t)pitch * h;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int y = (int)(i / (size_t)pitch), cw = (int)(i - (size_t)y * pitch);
    uint32_t v = in[i];
    int x0 = cw << 5;
    if (y >= gy0 && y <= gy1) {
      if (y < gy0 + t || y > gy1 - t) v |= range_mask(gx0, gx1, x0);
      else v |= range_mask(gx0, min(gx0 + t - 1, gx1), x0) | range_mask(max(gx1 - t + 1, gx0), gx1, x0);
    }
    out[i] = v;
  }
}

aos_status launch_pack(Ctx *c, const int8_t *src, uint32_t *dst, int w, int h) {
  int pitch = pitch_words_for(w);
  dim3 grid((pitch + 127) / 128, h);
  pack_kernel<<<grid, 128, 0, c->stream>>>(src, dst, w, h, pitch);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

aos_status launch_unpack(Ctx *c, const uint32_t *src, int8_t *dst, int w, int h) {
  int pitch = pitch_words_for(w);
  size_t total = (size_t)w * h;
  if ((w & 3) == 0 && (((uintptr_t)dst) & 3u) == 0) {
    size_t t4 = total / 4;
    int grid = (int)((t4 + 255) / 256 < (size_t)kNumSMs * 32 ? (t4 + 255) / 256 : (size_t)kNumSMs * 32);
    if (grid < 1) grid = 1;
    unpack4_kernel<<<grid, 256, 0, c->stream>>>(src, reinterpret_cast<uint32_t *>(dst), w / 4, h, pitch);
  ++c->launches;
  } else {
    int grid = (int)((total + 255) / 256 < (size_t)kNumSMs * 32 ? (total + 255) / 256 : (size_t)kNumSMs * 32);
    if (grid < 1) grid = 1;
    unpack_kernel<<<grid, 256, 0, c->stream>>>(src, dst, w, h, pitch);
  ++c->launches;
  }
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

aos_status launch_frame(Ctx *c, const uint32_t *in, uint32_t *out, int w, int h, int gx0, int gy0, int gx1, int gy1,
                        int t) {
  int pitch = pitch_words_for(w);
  size_t total = (size_t)pitch * h;
  int grid = (int)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
  if (grid < 1) grid = 1;
  frame_kernel<<<grid, 256, 0, c->stream>>>(in, out, w, h, pitch, gx0, gy0, gx1, gy1, t);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

// ---- trimPathNearOccupiedRegions (src/aos_path_gen_node.cpp:1570-1630) ----------
// One thread per pose: is a skeleton cell set inside the +-ceil(d/res) stencil (cells whose offset length
// sqrt(dx^2+dy^2)*res <= d)?  The reference stops at the first pose i > 0 for which that holds (pose 0 is tested
// but never trims), so the answer is the minimum such i.  Cell indices with the reference's expression order in
// double; (int) of an out-of-range or NaN value is INT_MIN on x86, i.e. outside the grid.
__global__ void trim_path_kernel(const double2 *__restrict__ path, int n, const uint32_t *__restrict__ bits, int w, int h,
                                 int pitch, double ox, double oy, double resolution, double safety, int radius_cells,
                                 int *__restrict__ first_hit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1 || i >= n) return;
  const double2 p = path[i];
  for (int dx = -radius_cells; dx <= radius_cells; ++dx)
    for (int dy = -radius_cells; dy <= radius_cells; ++dy) {
      const double dist = sqrt((double)(dx * dx + dy * dy)) * resolution;
      if (dist > safety) continue;
      const double fx = ((p.x + dx * resolution) - ox) / resolution, fy = ((p.y + dy * resolution) - oy) / resolution;
      if (!(fx > -2147483649.0 && fx < 2147483648.0 && fy > -2147483649.0 && fy < 2147483648.0)) continue;
      const int mx = (int)fx, my = (int)fy;
      if (mx < 0 || mx >= w || my < 0 || my >= h) continue;
      if ((__ldg(bits + (size_t)my * pitch + (mx >> 5)) >> (mx & 31)) & 1u) {
        atomicMin(first_hit, i);
        return;
      }
    }
}

// *n_kept = poses that remain.  `bits`: device, framed skeleton, pitch_words_for(w) words per row.
aos_status launch_trim_path(Ctx *c, const double *path_xy_host, int n, const uint32_t *bits, int w, int h, double ox, double oy,
                            float res, double safety, int *n_kept) {
  *n_kept = n;
  if (n <= 1) return AOS_OK;  // pose 0 never trims
  const double resolution = (double)res;
  const double rc = ceil(safety / resolution);
  if (!(rc >= 0 && rc < 4096)) {
    set_error(c, "safety_distance / resolution out of range");
    return AOS_ERR_INVALID;
  }
  AOS_CUDA_OK(c, c->seed_buf.reserve(sizeof(double2) * (size_t)n + 256));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  int *d_first = c->misc.as<int>() + 100;
  double2 *d_path = c->seed_buf.as<double2>();
  cudaStream_t st = c->stream;
  c->h_flag[0] = n;
  AOS_CUDA_OK(c, cudaMemcpyAsync(d_first, c->h_flag, 4, cudaMemcpyHostToDevice, st));
  AOS_CUDA_OK(c, cudaMemcpyAsync(d_path, path_xy_host, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, st));
  trim_path_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_path, n, bits, w, h, pitch_words_for(w), ox, oy, resolution, safety,
                                                      (int)rc, d_first);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_first, 4, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  *n_kept = c->h_flag[0];
  return AOS_OK;
}

// ---- small host -> device transfers without the copy engine --------------------------------------------------
// While another map's cloud is uploading, the host->device copy engine stays with that stream until it runs dry: a
// 4 MB cudaMemcpyAsync from a second stream was measured at 25 ms median / 56 ms worst (0.1 ms on an idle link), however
// the big upload was cut into pieces, while kernel launches and device->host copies were not delayed at all.  So the
// small per-map inputs (rows, seeds, replay jobs, the Subdiv2D arrays) are read by a kernel straight from page-locked
// host memory (unified addressing: pinned allocations and registered ranges are mapped into the device's address
// space); the loads share the PCIe link with the upload's DMA instead of queuing behind it.
__global__ void host_read_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16, uint32_t *__restrict__ dst_tail,
                                 const uint32_t *__restrict__ src_tail, int n_tail) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

// src must be page-locked (cudaHostAlloc / cudaHostRegister) and, like dst, 16-byte aligned; bytes a multiple of 4.
// Anything else takes the ordinary copy.
aos_status h2d_small(Ctx *c, void *dst, const void *src, size_t bytes, bool src_is_pinned) {
  if (bytes == 0) return AOS_OK;
  const void *dsrc = nullptr;
  if (src_is_pinned && (bytes & 3) == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)src & 15) == 0 &&
      cudaHostGetDevicePointer(const_cast<void **>(&dsrc), const_cast<void *>(src), 0) == cudaSuccess && dsrc) {
    const size_t n16 = bytes / 16;
    const int n_tail = (int)((bytes - n16 * 16) / 4);
    size_t blocks = (n16 + 255) / 256;
    blocks = std::min<size_t>(std::max<size_t>(blocks, 1), (size_t)kNumSMs * 8);
    host_read_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(static_cast<uint4 *>(dst), static_cast<const uint4 *>(dsrc), n16,
                                                                reinterpret_cast<uint32_t *>(static_cast<uint4 *>(dst) + n16),
                                                                reinterpret_cast<const uint32_t *>(static_cast<const uint4 *>(dsrc) + n16),
                                                                n_tail);
    ++c->launches;
    AOS_CUDA_OK(c, cudaGetLastError());
    return AOS_OK;
  }
  cudaGetLastError();
  AOS_CUDA_OK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return AOS_OK;
}

}  // namespace aos
