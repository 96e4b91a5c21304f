// aos_common.cuh -- shared declarations of libaos_gpu (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "aos_gpu.h"
#include "host_subdiv.h"

#ifndef __CUDA_ARCH_LIST__
#endif

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects itself

namespace aos {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Tile geometry shared by the stencil kernels.  TMA needs the innermost start coordinate and box width
// to be multiples of 16 bytes (4 words), so a CTA that owns kTileOwnW words per row stages the aligned
// superset [own_start-4, own_start+32).  Lane l of a warp works on column own_start - 2 + l, i.e. box word
// l + kTileLane0; lanes kTileLane0 .. kTileLane0+kTileOwnW-1 own output, two halo lanes on each side.
constexpr int kTileBoxW = 36;
constexpr int kTileOwnW = 28;
constexpr int kTileLane0 = 2;
constexpr int kMaxStencilRadius = 64;  // inflate_kernel's disc radius limit (cells); larger radii go through the EDT
constexpr int kThinSubIters = 8;  // sub-iterations per thinning launch == halo rows a row band refreshes per launch

// ---- error plumbing: nothing throws across the C boundary -------------------------------------
struct Ctx;
void set_error(Ctx *c, const char *fmt, ...);

#define AOS_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      aos::set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return AOS_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define AOS_REQUIRE(ctx, cond, msg)                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      aos::set_error((ctx), "%s:%d %s", __FILE__, __LINE__, (msg));  \
      return AOS_ERR_INVALID;                                        \
    }                                                                \
  } while (0)

// ---- bit-packed grid: 32 cells per word, LSB = lowest x ----------------------------------------
// Row pitch is a multiple of 4 words (16 B) so TMA tensor maps and uint4 accesses are legal.
__host__ __device__ inline int pitch_words_for(int width) { return (((width + 31) >> 5) + 3) & ~3; }

struct BitGrid {
  uint32_t *bits = nullptr;
  int w = 0, h = 0, pitch = 0;  // pitch in words
  size_t words() const { return (size_t)pitch * (size_t)h; }
  size_t bytes() const { return words() * 4; }
};

// A device buffer that only ever grows (contexts are reused map after map).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    // growing frees and re-allocates, and cudaFree / cudaFreeHost wait for the whole device: with many maps of varying
    // size in flight every growth step stalls all of them.  Small buffers therefore start at 1 MB and double (a
    // handful of steps to the steady state); only the big ones (point clouds, full-size planes) grow by an eighth.
    size_t want = bytes < ((size_t)64 << 20) ? std::max<size_t>(2 * bytes, (size_t)1 << 20) : bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

// Pinned (page-locked) host array that only ever grows: staging for the small host<->device transfers of the gvd
// half, so they run as true asynchronous DMA instead of pageable copies.
template <typename T>
struct PinVec {
  T *p = nullptr;
  size_t n = 0, cap = 0;
  bool resize(size_t count) {  // contents are NOT preserved when growing
    if (count > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      size_t want = std::max<size_t>(2 * count, ((size_t)64 << 10) / sizeof(T));  // see DevBuf::reserve
      if (cudaHostAlloc(reinterpret_cast<void **>(&p), want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) {
        n = 0;
        return false;
      }
      cap = want;
    }
    n = count;
    return true;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = cap = 0;
  }
  T *data() { return p; }
  const T *data() const { return p; }
  size_t size() const { return n; }
};

// ---- TMA (cp.async.bulk.tensor) helpers --------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();
// 2-D map over a bit grid (uint32 elements): dim0 = pitch words, dim1 = rows; OOB reads give 0.
bool make_bitgrid_tmap(CUtensorMap *map, const uint32_t *base, int pitch_words, int rows, int box_w, int box_h);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared tile; coordinates may be negative / past the end: TMA zero-fills.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
#endif

// ---- host-visible parameter block copied to the kernels -----------------------------------------
constexpr int kMaxPoly = 64;
constexpr int kMaxExcl = 64;

struct SeedDeviceParams {
  float minz, maxz, minx, maxx, miny, maxy;  // active bounds (float, inclusive)
  float res;
  double ox, oy;  // origin = (double)minx, (double)miny
  int w, h, pitch;
  int y_off, gh;  // row-band mode: global row of local row 0 and global height (0, h otherwise)
  int n_excl;
  float excl[kMaxExcl * 3];
  int n_poly;
  double poly[kMaxPoly * 2];
};

// ---- launchers (one per kernel file) -------------------------------------------------------------
struct Ctx;
aos_status launch_bin(Ctx *c, const SeedDeviceParams &P, const void *points, size_t n, uint32_t step, uint32_t ox,
                      uint32_t oy, uint32_t oz, uint32_t *bits, unsigned long long *n_kept);
aos_status launch_inflate(Ctx *c, const uint32_t *in, uint32_t *out, uint32_t *out_border, int w, int h, int R);
aos_status launch_open(Ctx *c, const uint32_t *in, uint32_t *out, int w, int h);
aos_status launch_thin(Ctx *c, uint32_t *img, uint32_t *scratch, int w, int h, int *launches, int *subiters);
struct ThinHalo {  // fused halo exchange of a row band, see ThinParams in k_thin.cu
  uint32_t *peer_lo = nullptr, *peer_hi = nullptr;
  int push_lo_r0 = 0, push_lo_shift = 0, push_hi_r0 = 0, push_hi_shift = 0;
  int skip_lo_r0 = -1, skip_hi_r0 = -1;
};
aos_status launch_thin_once(Ctx *c, const uint32_t *src, uint32_t *dst, int w, int h, int y_off, int gh, int cnt_r0,
                            int cnt_r1, int *d_count, const ThinHalo *halo);
aos_status run_ror(Ctx *c, const void *dpoints, size_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, float radius,
                   int min_neighbors, size_t *n_out);
aos_status facets_prepare(Ctx *c, const Subdiv &sd, int *n_slots);  // *n_slots = -1: structure is not a triangulation
aos_status facets_fill(Ctx *c, float2 *d_fxy, int *d_enext);
// k_vcells.cu: the opt-in device Voronoi (one thread per seed clips its cell); *n_slots = -1: a cell overflowed
aos_status vcells_prepare(Ctx *c, const double *seeds, int n_seeds, double min_x, double max_x, double min_y, double max_y,
                          int *n_slots);
aos_status vcells_fill(Ctx *c, float2 *d_fxy, int *d_enext);
void subdiv_release_pins(Ctx *c);  // host_gvd.cu

// RAII hold of one slot of the per-device kernel-phase gate (aos_api.cu; a no-op unless aos_set_device_gate(n > 0))
// NVTX range over one C-ABI call (tracing): "aos_seed_stage", "aos_gvd_stage", ...
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

struct DeviceGate {
  int dev = -1;
  explicit DeviceGate(const Ctx *c);
  ~DeviceGate() { release(); }
  void release();
  DeviceGate(const DeviceGate &) = delete;
  DeviceGate &operator=(const DeviceGate &) = delete;
};
aos_status launch_edt(Ctx *c, const uint32_t *bits, int w, int h, uint32_t *nearest, int32_t *dist2);
aos_status launch_edt_threshold(Ctx *c, const int32_t *dist2, int w, int h, int r2, uint32_t *out);
aos_status launch_frame(Ctx *c, const uint32_t *in, uint32_t *out, int w, int h, int gx0, int gy0, int gx1, int gy1,
                        int thickness);
aos_status launch_pack(Ctx *c, const int8_t *src, uint32_t *dst, int w, int h);
aos_status launch_unpack(Ctx *c, const uint32_t *src, int8_t *dst, int w, int h);
aos_status run_clusters(Ctx *c, const SeedDeviceParams &P, const uint32_t *skel, float min_length);
aos_status launch_labels(Ctx *c, int32_t *dst);
void host_rows_info(const std::vector<aos_tree_row> &rows, std::vector<double> *rows_info);
aos_status launch_trim_path(Ctx *c, const double *path_xy_host, int n, const uint32_t *bits, int w, int h, double ox, double oy,
                            float res, double safety, int *n_kept);
// host -> device without the copy engine when src is page-locked (k_misc.cu)
aos_status h2d_small(Ctx *c, void *dst, const void *src, size_t bytes, bool src_is_pinned);
aos_status device_select_seeds(Ctx *c);
aos_status device_merge_seeds(Ctx *c, const double *seeds, int n);  // -> c->h_merged
void host_merge_seeds(const double *seeds, int n, std::vector<double> *out);
aos_status exclusive_scan_u32(Ctx *c, uint32_t *data, size_t n, DevBuf &blocksum_buf, uint32_t *d_total);

// ---- gvd half --------------------------------------------------------------------------------------
// VoronoiDiagram::compute (vd:16-94) on the host: facets as flat float32 x,y + offsets; false when the
// reference returns early (no seeds / invalid bounds).
bool host_voronoi_facets(const double *seeds, int n, double min_x, double max_x, double min_y, double max_y,
                         std::vector<float> *facet_xy, std::vector<int32_t> *facet_off);

// device mirror of Subdiv::Vertex (same bytes, uploaded as they are; x, y are float32 values widened to double)
struct SdVertex {
  double x, y, n2;
  int first_edge, type;
};

struct GraphInputs {
  const float *facet_xy = nullptr;  // pinned host: one x,y per facet-vertex slot; slot e also is Voronoi edge e (vd:97-114)
  const int *enext = nullptr;       // pinned host: slot of the edge's end point (next vertex of the same facet)
  bool device_facets = false;       // slots and links come from facets_fill (k_facets.cu) instead of the two host arrays
  bool device_cells = false;        // ... or from vcells_fill (k_vcells.cu, aos_set_voronoi_mode(AOS_VORONOI_DEVICE))
  int n_slots = 0;
  const double *rows_info = nullptr;  // host, 4 per row
  int n_rows = 0;
  const uint32_t *skel_bits = nullptr;  // device, framed skeleton
  int w = 0, h = 0, pitch = 0;
  double ox = 0, oy = 0;
  float res = 0;
};

struct GraphHost {  // GvdGraph.msg arrays, host side
  float resolution = 0;
  double origin_x = 0, origin_y = 0;
  PinVec<double> nodes_xyz;
  PinVec<int32_t> node_labels, node_cluster_indices, node_label_counts, node_label_clusters, node_label_types, edges;
  PinVec<float> edge_lengths, edge_clearances;
  PinVec<double> corner_points;
  int n_merged_seeds = 0, n_voronoi_edges = 0, n_boundary_points = 0, n_rows = 0;
  void clear() {
    PinVec<int32_t> *iv[] = {&node_labels, &node_cluster_indices, &node_label_counts, &node_label_clusters, &node_label_types, &edges};
    for (auto *v : iv) v->n = 0;
    nodes_xyz.n = edge_lengths.n = edge_clearances.n = corner_points.n = 0;
    n_voronoi_edges = n_boundary_points = n_rows = 0;
  }
  void release() {
    PinVec<int32_t> *iv[] = {&node_labels, &node_cluster_indices, &node_label_counts, &node_label_clusters, &node_label_types, &edges};
    for (auto *v : iv) v->release();
    nodes_xyz.release(); edge_lengths.release(); edge_clearances.release(); corner_points.release();
  }
};
aos_status run_graph(Ctx *c, const GraphInputs &in);

struct Ctx {
  int device = 0;
  long long launches = 0;  // kernels launched by this context (aos_get_launch_count)
  // per-stage CUDA-event timers (aos_set_profiling)
  bool profile = false;
  bool composite = false;  // inside aos_map_to_graph: stages append to one timer list
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<std::string, int>> marks;  // (stage that ENDS at this event, event index)
  void mark(const char *name) {
    nvtxMarkA(name);  // NVTX marker where each stage's launches end on the host timeline (ranges: NvtxRange per C-ABI call)
    if (!profile) return;
    int i = (int)marks.size();
    if (i >= (int)ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    cudaEventRecord(ev_pool[i], stream);
    marks.emplace_back(name, i);
  }

  cudaStream_t stream = nullptr;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};  // side streams: independent launches run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  bool own_stream = false;
  std::string err;

  // row-band mode (aos_band_*): global row of local row 0 / global height; band_gh == 0 means not banded
  int band_y_off = 0, band_gh = 0, band_cnt_r0 = 0, band_cnt_r1 = 0;
  aos_band band{};
  bool have_band = false;
  SeedDeviceParams band_P{};
  aos_grid_info band_gi{};
  int band_thin_launches = 0;
  DevBuf band_thin[2];                    // p2p mode: dedicated ping-pong planes (exported over CUDA IPC, so no other
                                          // stage may ever re-allocate them)
  int band_cur = 0;                       // which of band_thin[] holds the current thinning image
  bool band_p2p = false;                  // a p2p launch has run since aos_band_raster
  void *band_peer[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [lo/hi neighbour][buffer 0/1], IPC-mapped
  bool band_exported = false;  // handles of band_thin[] were given out: the planes must not be re-allocated
  int band_peer_first_row[2] = {0, 0};    // neighbour's global row of its local row 0
  bool partial_grids = false, have_occ = false;  // after aos_seed_stage_tail only some grids exist

  // seed stage state
  bool have_seed = false;
  SeedDeviceParams P{};
  aos_seed_summary summary{};
  DevBuf g_raw, g_infl, g_occ, g_open, g_skel, g_framed, g_scratch;
  DevBuf points_stage;  // device copy of host point clouds
  DevBuf misc;          // small counters / flags
  int *h_flag = nullptr;  // pinned host word(s) for convergence flags and counters

  // clustering state (k_cluster.cu)
  DevBuf cc_mask, cc_prefix, cc_blocksum, cc_parent, cc_cellpos, cc_rootrank;
  DevBuf cl_stats;   // per-root accumulators
  DevBuf cl_table;   // aos_cluster[n_clusters]
  DevBuf cl_aux;     // per-cluster extreme points, candidate lists ...
  DevBuf cand_buf;
  DevBuf corner_fb;       // k_graph.cu: corners left to corner_far_kernel
  DevBuf bfs_buf;         // chain-compressed BFS order (k_cluster.cu, BfsBufs)
  int bfs_fallbacks = 0;  // clusters of the last map that took the literal replay instead
  int n_skel_cells = 0;
  int *d_cell_cluster = nullptr;  // compact cell -> cluster ordinal (inside cand_buf)
  int *d_root_cellpos = nullptr;  // cluster ordinal -> canonical label (inside cl_aux)
  int n_clusters = 0;
  std::vector<aos_cluster> h_clusters;
  std::vector<aos_tree_row> h_rows;
  std::vector<int32_t> h_cluster_root;  // compact index of each cluster's root

  // gvd half (host_gvd.cu, k_graph.cu)
  bool have_graph = false;
  GraphHost graph;
  DevBuf gvd_buf, gvd_buf2, gvd_buf3, gvd_skel, seed_buf, seed_buf2, edt_buf, edt_out, ror_buf, ror_out;
  bool clearance = false;  // aos_set_clearance
  PinVec<double> h_merged;       // page-locked: pageable transfers above 64 KB are staged by the driver and queue behind
                                 // the cloud uploads of the other maps in flight (seed selection 2.5 -> 50 ms measured)
  PinVec<double> pin_seed_in;    // caller-owned seeds staged for the merge
  PinVec<char> pin_a, pin_b, pin_c;  // staging of the small per-map tables (rows, replay jobs, cluster tables)
  Subdiv subdiv;                                    // lives in the context so its arrays are allocated (and pinned) once
  DevBuf sd_quads, sd_verts, sd_vor, sd_base;       // k_facets.cu (and k_vcells.cu in device-Voronoi mode)
  DevBuf vc_cells;                                  // k_vcells.cu: neighbour grid offsets + cursors
  int vc_n = 0;
  int voronoi_mode = 0;                             // aos_set_voronoi_mode: 0 = Subdiv2D replay (bit-exact), 1 = device cells
  int sd_nv = 0, sd_nq = 0;
  void *sd_pinned[3] = {nullptr, nullptr, nullptr};  // cudaHostRegister'ed storage of subdiv's three arrays
  size_t sd_pinned_bytes[3] = {0, 0, 0};
  PinVec<float> pin_facet_xy;
  PinVec<int> pin_enext;
  PinVec<double> pin_rows;

  // host seed selection (host_seeds.cu)
  bool have_seeds = false;
  PinVec<double> h_seeds;
  int seed_counts[3] = {0, 0, 0};     // virtual, ray, endpoint
  std::vector<double> h_rows_info;    // /exploration_tree_rows_info: start x,y,end x,y per row, sorted
  std::vector<double> ray_steps;      // k_seeds.cu: castRayFromEndpoint's accumulated ray parameter, step by step
  DevBuf ray_table;                   // ... its device copy (ray_steps_dev entries)
  size_t ray_steps_dev = 0;
};

}  // namespace aos
