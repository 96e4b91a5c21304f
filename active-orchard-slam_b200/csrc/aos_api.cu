// aos_api.cu -- the C-ABI of libaos_gpu (include/aos_gpu.h): context, the seed-gen stage pipeline
// (processPointCloud, src/aos_seed_gen_node.cpp:452-579) and the getters.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>

#include "aos_common.cuh"

using namespace aos;

struct aos_ctx : public aos::Ctx {};

namespace {

// Host -> device upload of a point cloud.  Large uploads from different contexts (maps in flight on several host
// threads) are serialised through one process-wide gate: concurrent copies would only share the PCIe link and
// finish together, which keeps the threads in lockstep (all uploading, then all computing); one copy at a time runs
// at the full link rate and staggers the threads, so the upload of one map overlaps the host/GPU stages of the others.
std::mutex g_upload_gates[64];  // one per device: uploads to different GPUs use different links

// Kernel-phase gate (aos_set_device_gate): a counting semaphore per device.  Identical maps in flight on one GPU (one
// context + host thread each) run in lockstep: all of them in their kernel phase at once -- the GPU time-slices 16
// seed stages and every one of them finishes late -- then all in their host phase with the GPU idle.  Admitting only
// a few maps at a time to the kernel phase staggers the threads after the first round, so later kernel phases meet a
// mostly free GPU.  (One at a time is too strict: the phases are latency-bound and overlap each other well.)
struct GateSlot {  // first come, first served: ticket t is admitted once fewer than `cap` earlier tickets are still inside
  std::mutex m;
  std::condition_variable cv;
  unsigned long long next_ticket = 0, released = 0;
};
GateSlot g_device_gates[64];
std::atomic<int> g_device_gate_cap{0};
}  // namespace

namespace aos {
DeviceGate::DeviceGate(const Ctx *c) {
  const int cap = g_device_gate_cap.load(std::memory_order_relaxed);
  if (cap <= 0 || !c || c->device < 0 || c->device >= 64) return;
  GateSlot &g = g_device_gates[c->device];
  std::unique_lock<std::mutex> lk(g.m);
  const unsigned long long t = g.next_ticket++;
  g.cv.wait(lk, [&] {
    const int k = g_device_gate_cap.load(std::memory_order_relaxed);
    return k <= 0 || t < g.released + (unsigned long long)k;
  });
  dev = c->device;
}
void DeviceGate::release() {
  if (dev < 0) return;
  GateSlot &g = g_device_gates[dev];
  {
    std::lock_guard<std::mutex> lk(g.m);
    ++g.released;
  }
  g.cv.notify_all();
  dev = -1;
}
}  // namespace aos

namespace {
aos_status upload_points(aos_ctx *c, const void *points, size_t bytes, const void **dpoints) {
  AOS_CUDA_OK(c, c->points_stage.reserve(bytes));
  if (bytes >= ((size_t)64 << 20)) {
    static const bool dbg = getenv("AOS_DEBUG") != nullptr;
    const auto t_wait = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lock(g_upload_gates[c->device & 63]);
    const auto t0 = std::chrono::steady_clock::now();
    // One copy: cutting it into pieces (even with only two queued at a time) neither frees the copy engine for other
    // streams' transfers -- it stays with this stream until it runs dry -- nor is it free (59.9 vs 57.7 ms per 3.2 GB).
    // The other maps' small inputs therefore do not use the copy engine at all (h2d_small, k_misc.cu).
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->points_stage.p, points, bytes, cudaMemcpyHostToDevice, c->stream));
    AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (dbg) {
      const auto t1 = std::chrono::steady_clock::now();
      const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      fprintf(stderr, "[aos] upload %.2f GB in %.1f ms (%.1f GB/s), waited %.1f ms at the gate\n", bytes / 1e9, ms,
              bytes / 1e6 / ms, std::chrono::duration<double, std::milli>(t0 - t_wait).count());
    }
  } else {
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->points_stage.p, points, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  *dpoints = c->points_stage.p;
  return AOS_OK;
}

// getActiveBounds, seed_gen:874-890
void active_bounds(const aos_seed_params *p, float *minx, float *maxx, float *miny, float *maxy) {
  if (p->n_polygon > 0) {
    double bx0 = p->polygon[0], bx1 = p->polygon[0], by0 = p->polygon[1], by1 = p->polygon[1];
    for (int i = 0; i < p->n_polygon; ++i) {
      bx0 = std::min(bx0, p->polygon[2 * i]);
      bx1 = std::max(bx1, p->polygon[2 * i]);
      by0 = std::min(by0, p->polygon[2 * i + 1]);
      by1 = std::max(by1, p->polygon[2 * i + 1]);
    }
    const double margin = 2.5;
    *minx = static_cast<float>(bx0 - margin);
    *maxx = static_cast<float>(bx1 + margin);
    *miny = static_cast<float>(by0 - margin);
    *maxy = static_cast<float>(by1 + margin);
  } else {
    *minx = p->clipping_minx;
    *maxx = p->clipping_maxx;
    *miny = p->clipping_miny;
    *maxy = p->clipping_maxy;
  }
}

// worldToGrid, seed_gen:760-769 (float result of a double expression, floor, clamp)
void world_to_grid(double ox, double oy, float res, int w, int h, float wx, float wy, int *gx, int *gy) {
  float rel_x = static_cast<float>((wx - ox) / res);
  float rel_y = static_cast<float>((wy - oy) / res);
  *gx = static_cast<int>(floorf(rel_x));
  *gy = static_cast<int>(floorf(rel_y));
  *gx = std::max(0, std::min(w - 1, *gx));
  *gy = std::max(0, std::min(h - 1, *gy));
}

aos_status check_params(aos_ctx *ctx, const aos_seed_params *p) {
  AOS_REQUIRE(ctx, p != nullptr, "params is null");
  AOS_REQUIRE(ctx, p->grid_resolution > 0.f && std::isfinite(p->grid_resolution), "grid_resolution must be > 0");
  AOS_REQUIRE(ctx, p->n_polygon >= 0 && p->n_polygon <= kMaxPoly, "polygon has more than 64 vertices");
  AOS_REQUIRE(ctx, p->n_exclusion >= 0 && p->n_exclusion <= kMaxExcl, "more than 64 exclusion discs");
  AOS_REQUIRE(ctx, p->n_polygon == 0 || p->polygon != nullptr, "polygon pointer is null");
  AOS_REQUIRE(ctx, p->n_exclusion == 0 || p->exclusion != nullptr, "exclusion pointer is null");
  return AOS_OK;
}

DevBuf *grid_buf(aos_ctx *c, aos_grid_id which) {
  switch (which) {
    case AOS_GRID_RAW: return &c->g_raw;
    case AOS_GRID_INFLATED: return &c->g_infl;
    case AOS_GRID_OCCUPANCY: return &c->g_occ;
    case AOS_GRID_OPENED: return &c->g_open;
    case AOS_GRID_SKELETON: return &c->g_skel;
    case AOS_GRID_SKELETON_FRAMED: return &c->g_framed;
  }
  return nullptr;
}

}  // namespace

extern "C" {

const char *aos_version(void) { return "libaos_gpu 0.1 (sm_100a)"; }

int32_t aos_bits_pitch_words(int32_t width) { return pitch_words_for(width); }

aos_status aos_create(int device, aos_ctx **out) {
  if (!out) return AOS_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return AOS_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return AOS_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return AOS_ERR_CUDA;
  aos_ctx *c = new (std::nothrow) aos_ctx();
  if (!c) return AOS_ERR_CUDA;
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void **>(&c->h_flag), 64 * sizeof(int)) != cudaSuccess) {
    delete c;
    return AOS_ERR_CUDA;
  }
  c->own_stream = true;
  bool ok = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < 3 && ok; ++k)
    ok = cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    aos_destroy(c);
    return AOS_ERR_CUDA;
  }
  *out = c;
  return AOS_OK;
}

void aos_destroy(aos_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf *bufs[] = {&c->g_raw, &c->g_infl, &c->g_occ, &c->g_open, &c->g_skel, &c->g_framed, &c->g_scratch,
                    &c->points_stage, &c->misc, &c->cc_mask, &c->cc_prefix, &c->cc_blocksum, &c->cc_parent,
                    &c->cc_cellpos, &c->cc_rootrank, &c->cl_stats, &c->cl_table, &c->cl_aux, &c->cand_buf, &c->bfs_buf, &c->ray_table, &c->corner_fb,
                    &c->gvd_buf, &c->gvd_buf2, &c->gvd_buf3, &c->gvd_skel, &c->seed_buf, &c->seed_buf2, &c->edt_buf, &c->edt_out, &c->ror_buf, &c->ror_out};
  for (DevBuf *b : bufs) b->release();
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b)
      if (c->band_peer[s][b]) cudaIpcCloseMemHandle(c->band_peer[s][b]);
  c->band_thin[0].release();
  c->band_thin[1].release();
  DevBuf *sdb[] = {&c->sd_quads, &c->sd_verts, &c->sd_vor, &c->sd_base, &c->vc_cells};
  for (DevBuf *b : sdb) b->release();
  subdiv_release_pins(c);
  c->graph.release();
  c->pin_facet_xy.release();
  c->pin_enext.release();
  c->pin_rows.release();
  c->h_seeds.release();
  c->h_merged.release();
  c->pin_seed_in.release();
  c->pin_a.release();
  c->pin_b.release();
  c->pin_c.release();
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  for (int k = 0; k < 3; ++k) {
    if (c->aux[k]) cudaStreamDestroy(c->aux[k]);
    if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *aos_last_error(const aos_ctx *c) { return c ? c->err.c_str() : "null context"; }

aos_status aos_set_stream(aos_ctx *c, void *cuda_stream) {
  if (!c) return AOS_ERR_INVALID;
  if (c->own_stream && c->stream) {
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
  }
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  c->own_stream = false;
  return AOS_OK;
}

aos_status aos_set_host_wait(int device, int32_t blocking) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return AOS_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return AOS_ERR_INVALID;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return AOS_ERR_CUDA;
  unsigned flags = 0;
  cudaError_t e = cudaGetDeviceFlags(&flags);
  if (e == cudaSuccess) {
    flags = (flags & ~(unsigned)cudaDeviceScheduleMask) | (blocking ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto);
    e = cudaSetDeviceFlags(flags);
  }
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return AOS_ERR_CUDA;
  }
  return AOS_OK;
}

aos_status aos_set_device_gate(int32_t max_concurrent) {
  g_device_gate_cap.store(max_concurrent > 0 ? max_concurrent : 0);
  for (auto &g : g_device_gates) g.cv.notify_all();
  return AOS_OK;
}

int32_t aos_get_device_gate(void) { return g_device_gate_cap.load(); }

aos_status aos_set_voronoi_mode(aos_ctx *c, int32_t mode) {
  if (!c || (mode != AOS_VORONOI_REPLAY && mode != AOS_VORONOI_DEVICE)) return AOS_ERR_INVALID;
  c->voronoi_mode = mode;
  return AOS_OK;
}

aos_status aos_set_profiling(aos_ctx *c, int enabled) {
  if (!c) return AOS_ERR_INVALID;
  c->profile = enabled != 0;
  return AOS_OK;
}

aos_status aos_get_stage_times(aos_ctx *c, aos_stage_time *dst, int32_t capacity, int32_t *n_out) {
  if (!c) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  int n = c->marks.empty() ? 0 : (int)c->marks.size() - 1;
  if (n_out) *n_out = n;
  if (!dst) return AOS_OK;
  if (capacity < n) return AOS_ERR_CAPACITY;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_pool[c->marks[i].second], c->ev_pool[c->marks[i + 1].second]);
    memset(dst[i].name, 0, sizeof(dst[i].name));
    strncpy(dst[i].name, c->marks[i + 1].first.c_str(), sizeof(dst[i].name) - 1);
    dst[i].ms = ms;
  }
  return AOS_OK;
}

aos_status aos_synchronize(aos_ctx *c) {
  if (!c) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_grid_geometry(const aos_seed_params *p, aos_grid_info *info) {
  if (!p || !info || !(p->grid_resolution > 0.f)) return AOS_ERR_INVALID;
  float minx, maxx, miny, maxy;
  active_bounds(p, &minx, &maxx, &miny, &maxy);
  // generateOccupancyGrid, seed_gen:587-600: float arithmetic
  float width = std::max(0.0f, maxx - minx);
  float height = std::max(0.0f, maxy - miny);
  unsigned int w = static_cast<unsigned int>(std::ceil(width / p->grid_resolution));
  unsigned int h = static_cast<unsigned int>(std::ceil(height / p->grid_resolution));
  if (w == 0) w = 1;
  if (h == 0) h = 1;
  info->width = static_cast<int32_t>(w);
  info->height = static_cast<int32_t>(h);
  info->resolution = p->grid_resolution;
  info->origin_x = minx;
  info->origin_y = miny;
  return AOS_OK;
}

}  // extern "C"

namespace {

// grid geometry + device parameter block of the FULL grid (generateOccupancyGrid header, seed_gen:587-600)
aos_status fill_device_params(aos_ctx *c, const aos_seed_params *p, aos_grid_info *gi, SeedDeviceParams *Pp) {
  aos_grid_geometry(p, gi);
  AOS_REQUIRE(c, (double)gi->width * (double)gi->height < 2.0e9, "grid has more than 2^31 cells");
  SeedDeviceParams &P = *Pp;
  memset(&P, 0, sizeof(P));
  active_bounds(p, &P.minx, &P.maxx, &P.miny, &P.maxy);
  P.minz = p->clipping_minz;
  P.maxz = p->clipping_maxz;
  P.res = p->grid_resolution;
  P.ox = gi->origin_x;
  P.oy = gi->origin_y;
  P.w = gi->width;
  P.h = gi->height;
  P.pitch = pitch_words_for(gi->width);
  P.y_off = 0;
  P.gh = gi->height;
  P.n_excl = p->n_exclusion;
  for (int i = 0; i < 3 * p->n_exclusion; ++i) P.excl[i] = p->exclusion[i];
  P.n_poly = p->n_polygon;
  for (int i = 0; i < 2 * p->n_polygon; ++i) P.poly[i] = p->polygon[i];
  return AOS_OK;
}

// markPolygonBoundaryAsOccupied (seed_gen:772-825) + clusterOccupiedCells / rows on the full-size skeleton in
// c->g_skel; fills the summary.  Shared by aos_seed_stage and aos_seed_stage_tail.
aos_status seed_tail(aos_ctx *c, const aos_seed_params *p, const aos_grid_info &gi, int launches, int subiters,
                     const unsigned long long *d_kept) {
  const SeedDeviceParams &P = c->P;
  cudaStream_t st = c->stream;
  aos_status s;
  if (p->n_polygon > 0) {
    double bx0 = p->polygon[0], bx1 = bx0, by0 = p->polygon[1], by1 = by0;
    for (int i = 0; i < p->n_polygon; ++i) {
      bx0 = std::min(bx0, p->polygon[2 * i]);
      bx1 = std::max(bx1, p->polygon[2 * i]);
      by0 = std::min(by0, p->polygon[2 * i + 1]);
      by1 = std::max(by1, p->polygon[2 * i + 1]);
    }
    const double margin = 2.5;
    int gx0, gy0, gx1, gy1;
    world_to_grid(P.ox, P.oy, P.res, P.w, P.h, static_cast<float>(bx0 - margin), static_cast<float>(by0 - margin), &gx0, &gy0);
    world_to_grid(P.ox, P.oy, P.res, P.w, P.h, static_cast<float>(bx1 + margin), static_cast<float>(by1 + margin), &gx1, &gy1);
    s = launch_frame(c, c->g_skel.as<uint32_t>(), c->g_framed.as<uint32_t>(), P.w, P.h, std::min(gx0, gx1),
                     std::min(gy0, gy1), std::max(gx0, gx1), std::max(gy0, gy1), 1);
  } else {
    s = launch_frame(c, c->g_skel.as<uint32_t>(), c->g_framed.as<uint32_t>(), P.w, P.h, 0, 0, P.w - 1, P.h - 1, 5);
  }
  if (s != AOS_OK) return s;

  c->mark("frame");
  s = run_clusters(c, P, c->g_skel.as<uint32_t>(), static_cast<float>(p->cluster_min_length));
  if (s != AOS_OK) return s;
  c->mark("cluster_tail");

  if (d_kept) AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag + 32, d_kept, 8, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->summary.info = gi;
  c->summary.n_clusters = c->n_clusters;
  c->summary.n_rows = static_cast<int32_t>(c->h_rows.size());
  c->summary.thinning_launches = launches;
  c->summary.thinning_subiters = subiters;
  c->summary.n_points_in = 0;
  if (d_kept) memcpy(&c->summary.n_points_in, c->h_flag + 32, 8);
  c->have_seed = true;
  return AOS_OK;
}

aos_status check_points(aos_ctx *c, const void *points, size_t n_points, uint32_t point_step, uint32_t off_x,
                        uint32_t off_y, uint32_t off_z) {
  AOS_REQUIRE(c, n_points == 0 || points != nullptr, "points is null");
  AOS_REQUIRE(c, n_points == 0 || (point_step >= 12 && off_x + 4 <= point_step && off_y + 4 <= point_step &&
                                   off_z + 4 <= point_step),
              "point_step / field offsets do not describe float32 x,y,z inside a record");
  return AOS_OK;
}

}  // namespace

extern "C" {

aos_status aos_seed_stage(aos_ctx *c, const aos_seed_params *p, const void *points, size_t n_points,
                          uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z, aos_mem points_mem) {
  NvtxRange nvtx_range("aos_seed_stage");
  if (!c) return AOS_ERR_INVALID;
  aos_status s = check_params(c, p);
  if (s != AOS_OK) return s;
  s = check_points(c, points, n_points, point_step, off_x, off_y, off_z);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->have_seed = false;
  c->have_seeds = false;
  c->have_band = false;
  c->band_gh = 0;
  c->partial_grids = false;

  aos_grid_info gi;
  SeedDeviceParams &P = c->P;
  s = fill_device_params(c, p, &gi, &P);
  if (s != AOS_OK) return s;

  const size_t gbytes = (size_t)P.pitch * P.h * 4;
  DevBuf *grids[] = {&c->g_raw, &c->g_infl, &c->g_occ, &c->g_open, &c->g_skel, &c->g_framed, &c->g_scratch};
  for (DevBuf *g : grids) AOS_CUDA_OK(c, g->reserve(gbytes));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  cudaStream_t st = c->stream;
  AOS_CUDA_OK(c, cudaMemsetAsync(c->g_raw.p, 0, gbytes, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(c->g_scratch.p, 0, gbytes, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(c->misc.p, 0, 4096, st));

  c->marks.clear();
  c->mark("start");
  const void *dpoints = points;
  if (points_mem == AOS_MEM_HOST && n_points) {
    s = upload_points(c, points, n_points * (size_t)point_step, &dpoints);
    if (s != AOS_OK) return s;
  }
  c->mark("h2d_points");
  DeviceGate gate(c);  // kernel phase of this map: held to the end of the stage
  c->mark("gate");
  unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(c->misc.as<char>() + 1024);
  s = launch_bin(c, P, dpoints, n_points, point_step, off_x, off_y, off_z, c->g_raw.as<uint32_t>(), d_kept);
  if (s != AOS_OK) return s;

  c->mark("bin");
  // applyInflation: int(inflation_radius / grid_resolution) in float (seed_gen:936)
  const int R = static_cast<int>(p->inflation_radius / p->grid_resolution);
  AOS_REQUIRE(c, R >= 0, "negative inflation radius");
  if (R <= kMaxStencilRadius) {
    s = launch_inflate(c, c->g_raw.as<uint32_t>(), c->g_infl.as<uint32_t>(), c->g_occ.as<uint32_t>(), P.w, P.h, R);
  } else {
    // beyond the stencil's reach: the same set as the threshold d^2 <= R^2 of the exact EDT (k_edt.cu), whose cost
    // does not depend on R; the 5-cell frame of markBoundariesAsOccupied is OR-ed in separately
    const size_t cells = (size_t)P.w * P.h;
    AOS_CUDA_OK(c, c->edt_out.reserve(cells * 8 + 1024));
    uint32_t *nearest = c->edt_out.as<uint32_t>();
    int32_t *d2 = reinterpret_cast<int32_t *>(nearest + cells);
    s = launch_edt(c, c->g_raw.as<uint32_t>(), P.w, P.h, nearest, d2);
    if (s == AOS_OK) s = launch_edt_threshold(c, d2, P.w, P.h, R * R, c->g_infl.as<uint32_t>());
    if (s == AOS_OK) s = launch_frame(c, c->g_infl.as<uint32_t>(), c->g_occ.as<uint32_t>(), P.w, P.h, 0, 0, P.w - 1, P.h - 1, 5);
  }
  if (s != AOS_OK) return s;
  c->mark("inflate");
  // skeletonizeOccupancyGrid runs on the inflated grid WITHOUT the frame (seed_gen:560)
  s = launch_open(c, c->g_infl.as<uint32_t>(), c->g_open.as<uint32_t>(), P.w, P.h);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_skel.p, c->g_open.p, gbytes, cudaMemcpyDeviceToDevice, st));
  c->mark("open");
  int launches = 0, subiters = 0;
  s = launch_thin(c, c->g_skel.as<uint32_t>(), c->g_scratch.as<uint32_t>(), P.w, P.h, &launches, &subiters);
  if (s != AOS_OK) return s;

  c->mark("thin");
  return seed_tail(c, p, gi, launches, subiters, d_kept);
}

// ---- row-band sharding ------------------------------------------------------------------------------------
int32_t aos_band_halo_rows(const aos_seed_params *p) {
  if (!p || !(p->grid_resolution > 0.f)) return -1;
  // inflation reaches R rows, the opening 2, one thinning launch kThinSubIters
  return static_cast<int>(p->inflation_radius / p->grid_resolution) + 2 + kThinSubIters;
}

aos_status aos_band_raster(aos_ctx *c, const aos_seed_params *p, const aos_band *band, const void *points, size_t n_points,
                           uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z, aos_mem points_mem) {
  NvtxRange nvtx_range("aos_band_raster");
  if (!c) return AOS_ERR_INVALID;
  aos_status s = check_params(c, p);
  if (s != AOS_OK) return s;
  s = check_points(c, points, n_points, point_step, off_x, off_y, off_z);
  if (s != AOS_OK) return s;
  AOS_REQUIRE(c, band != nullptr, "band is null");
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->have_seed = false;
  c->have_seeds = false;
  c->have_band = false;

  aos_grid_info gi;
  SeedDeviceParams P;
  s = fill_device_params(c, p, &gi, &P);
  if (s != AOS_OK) return s;
  const int need = aos_band_halo_rows(p);
  AOS_REQUIRE(c, band->rows > 0 && band->row0 >= 0 && band->row0 + band->rows <= gi.height, "band outside the grid");
  AOS_REQUIRE(c, band->halo_lo >= 0 && band->halo_hi >= 0 && band->halo_lo <= band->row0 &&
                     band->row0 + band->rows + band->halo_hi <= gi.height,
              "band halo outside the grid");
  AOS_REQUIRE(c, (band->halo_lo >= need || band->halo_lo == band->row0) &&
                     (band->halo_hi >= need || band->row0 + band->rows + band->halo_hi == gi.height),
              "band halo smaller than aos_band_halo_rows() away from the global border");
  // local grid = halo_lo + rows + halo_hi rows; local row r is global row y_off + r
  P.y_off = band->row0 - band->halo_lo;
  P.h = band->halo_lo + band->rows + band->halo_hi;
  c->band = *band;
  c->band_y_off = P.y_off;
  c->band_gh = gi.height;
  c->band_cnt_r0 = band->halo_lo;
  c->band_cnt_r1 = band->halo_lo + band->rows;
  c->band_P = P;
  c->band_gi = gi;

  const size_t gbytes = (size_t)P.pitch * P.h * 4;
  DevBuf *grids[] = {&c->g_raw, &c->g_infl, &c->g_occ, &c->g_open, &c->g_skel, &c->g_scratch};
  for (DevBuf *g : grids) AOS_CUDA_OK(c, g->reserve(gbytes));
  // the two thinning planes may be mapped by the neighbouring ranks (CUDA IPC): growing them would free memory a peer
  // still has open, so a larger map needs aos_band_ipc_release on every rank (and a barrier) first
  if (c->band_exported && (gbytes > c->band_thin[0].cap || gbytes > c->band_thin[1].cap)) {
    set_error(c, "band planes are exported over CUDA IPC and too small for this map: call aos_band_ipc_release on every rank "
                 "(then a barrier) before a larger map");
    return AOS_ERR_STATE;
  }
  AOS_CUDA_OK(c, c->band_thin[0].reserve(gbytes));
  AOS_CUDA_OK(c, c->band_thin[1].reserve(gbytes));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  cudaStream_t st = c->stream;
  AOS_CUDA_OK(c, cudaMemsetAsync(c->g_raw.p, 0, gbytes, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(c->g_scratch.p, 0, gbytes, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(c->misc.p, 0, 4096, st));
  c->marks.clear();
  c->mark("start");
  const void *dpoints = points;
  if (points_mem == AOS_MEM_HOST && n_points) {
    s = upload_points(c, points, n_points * (size_t)point_step, &dpoints);
    if (s != AOS_OK) return s;
  }
  c->mark("h2d_points");
  unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(c->misc.as<char>() + 1024);
  s = launch_bin(c, P, dpoints, n_points, point_step, off_x, off_y, off_z, c->g_raw.as<uint32_t>(), d_kept);
  c->mark("bin");
  const int R = static_cast<int>(p->inflation_radius / p->grid_resolution);
  if (s == AOS_OK) s = launch_inflate(c, c->g_raw.as<uint32_t>(), c->g_infl.as<uint32_t>(), c->g_occ.as<uint32_t>(), P.w, P.h, R);
  c->mark("inflate");
  if (s == AOS_OK) s = launch_open(c, c->g_infl.as<uint32_t>(), c->g_open.as<uint32_t>(), P.w, P.h);
  c->band_gh = 0;  // the launchers read the band geometry only while a band call is running
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_skel.p, c->g_open.p, gbytes, cudaMemcpyDeviceToDevice, st));
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->band_thin[0].p, c->g_open.p, gbytes, cudaMemcpyDeviceToDevice, st));
  c->mark("open");
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  c->band_thin_launches = 0;
  c->band_cur = 0;
  c->band_p2p = false;
  c->have_band = true;
  return AOS_OK;
}

aos_status aos_band_thin_launch(aos_ctx *c, int32_t *deleted) {
  NvtxRange nvtx_range("aos_band_thin_launch");
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_band) {
    set_error(c, "aos_band_raster has not completed");
    return AOS_ERR_STATE;
  }
  AOS_REQUIRE(c, !c->band_p2p, "this band is being thinned with aos_band_thin_launch_p2p");
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  const SeedDeviceParams &P = c->band_P;
  int *d_count = c->misc.as<int>() + 64;
  aos_status s = launch_thin_once(c, c->g_skel.as<uint32_t>(), c->g_scratch.as<uint32_t>(), P.w, P.h, c->band_y_off,
                                  c->band_gi.height, c->band_cnt_r0, c->band_cnt_r1, d_count, nullptr);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_skel.p, c->g_scratch.p, (size_t)P.pitch * P.h * 4, cudaMemcpyDeviceToDevice, c->stream));
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  ++c->band_thin_launches;
  if (deleted) *deleted = c->h_flag[0] != 0;
  return AOS_OK;
}

// ---- fused halo exchange over peer memory -------------------------------------------------------------------
aos_status aos_band_ipc_export(aos_ctx *c, int32_t buffer, unsigned char *handle) {
  if (!c || !handle || (buffer != 0 && buffer != 1)) return AOS_ERR_INVALID;
  if (!c->have_band) return AOS_ERR_STATE;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == AOS_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  AOS_CUDA_OK(c, cudaIpcGetMemHandle(&h, c->band_thin[buffer].p));
  memcpy(handle, &h, sizeof(h));
  c->band_exported = true;
  return AOS_OK;
}

aos_status aos_band_ipc_release(aos_ctx *c) {
  if (!c) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b)
      if (c->band_peer[s][b]) {
        cudaIpcCloseMemHandle(c->band_peer[s][b]);
        c->band_peer[s][b] = nullptr;
      }
  c->band_exported = false;  // the caller's barrier guarantees that the peers have closed their mappings too
  return AOS_OK;
}

aos_status aos_band_ipc_import(aos_ctx *c, int32_t side, int32_t buffer, const unsigned char *handle,
                               int32_t peer_first_global_row) {
  if (!c || !handle || (side != 0 && side != 1) || (buffer != 0 && buffer != 1)) return AOS_ERR_INVALID;
  if (!c->have_band) return AOS_ERR_STATE;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  if (c->band_peer[side][buffer]) {
    cudaIpcCloseMemHandle(c->band_peer[side][buffer]);
    c->band_peer[side][buffer] = nullptr;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  AOS_CUDA_OK(c, cudaIpcOpenMemHandle(&c->band_peer[side][buffer], h, cudaIpcMemLazyEnablePeerAccess));
  c->band_peer_first_row[side] = peer_first_global_row;
  return AOS_OK;
}

aos_status aos_band_thin_launch_p2p(aos_ctx *c, int32_t *deleted) {
  NvtxRange nvtx_range("aos_band_thin_launch_p2p");
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_band) {
    set_error(c, "aos_band_raster has not completed");
    return AOS_ERR_STATE;
  }
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  const SeedDeviceParams &P = c->band_P;
  const aos_band &B = c->band;
  const bool has_lo = B.halo_lo > 0, has_hi = B.halo_hi > 0;
  AOS_REQUIRE(c, c->band_p2p || c->band_thin_launches == 0, "this band is being thinned with aos_band_thin_launch");
  const int nxt = c->band_cur ^ 1;  // every rank flips in lockstep, so the neighbours' destination has the same index
  AOS_REQUIRE(c, (!has_lo || c->band_peer[0][nxt]) && (!has_hi || c->band_peer[1][nxt]),
              "neighbour buffers not imported (aos_band_ipc_import)");
  AOS_REQUIRE(c, B.rows >= kThinSubIters, "band lower than the thinning halo");
  ThinHalo H;
  if (has_lo) {
    H.peer_lo = static_cast<uint32_t *>(c->band_peer[0][nxt]);
    H.push_lo_r0 = B.halo_lo;
    H.push_lo_shift = c->band_y_off - c->band_peer_first_row[0];
    H.skip_lo_r0 = B.halo_lo - kThinSubIters;
  }
  if (has_hi) {
    H.peer_hi = static_cast<uint32_t *>(c->band_peer[1][nxt]);
    H.push_hi_r0 = B.halo_lo + B.rows - kThinSubIters;
    H.push_hi_shift = c->band_y_off - c->band_peer_first_row[1];
    H.skip_hi_r0 = B.halo_lo + B.rows;
  }
  uint32_t *bufs[2] = {c->band_thin[0].as<uint32_t>(), c->band_thin[1].as<uint32_t>()};
  int *d_count = c->misc.as<int>() + 64;
  aos_status s = launch_thin_once(c, bufs[c->band_cur], bufs[nxt], P.w, P.h, c->band_y_off, c->band_gi.height, c->band_cnt_r0,
                                  c->band_cnt_r1, d_count, &H);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  c->band_cur = nxt;
  c->band_p2p = true;
  ++c->band_thin_launches;
  if (deleted) *deleted = c->h_flag[0] != 0;
  return AOS_OK;
}

aos_status aos_band_grid_device(aos_ctx *c, aos_grid_id which, uint32_t **bits, int32_t *pitch_words, int32_t *local_rows) {
  if (!c || !bits) return AOS_ERR_INVALID;
  if (!c->have_band) return AOS_ERR_STATE;
  AOS_REQUIRE(c, which >= AOS_GRID_RAW && which <= AOS_GRID_SKELETON, "grid not available in band mode");
  *bits = (which == AOS_GRID_SKELETON && c->band_p2p) ? c->band_thin[c->band_cur].as<uint32_t>() : grid_buf(c, which)->as<uint32_t>();
  if (pitch_words) *pitch_words = c->band_P.pitch;
  if (local_rows) *local_rows = c->band_P.h;
  return AOS_OK;
}

aos_status aos_seed_stage_tail(aos_ctx *c, const aos_seed_params *p, const uint32_t *skeleton_bits, const uint32_t *occupancy_bits) {
  NvtxRange nvtx_range("aos_seed_stage_tail");
  if (!c || !skeleton_bits) return AOS_ERR_INVALID;
  aos_status s = check_params(c, p);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->have_seed = false;
  c->have_seeds = false;
  const int launches = c->have_band ? c->band_thin_launches : 0;
  c->have_band = false;
  c->band_gh = 0;
  aos_grid_info gi;
  s = fill_device_params(c, p, &gi, &c->P);
  if (s != AOS_OK) return s;
  const SeedDeviceParams &P = c->P;
  const size_t gbytes = (size_t)P.pitch * P.h * 4;
  // the inputs may alias this context's own (smaller, local) buffers: stage them before the buffers are regrown
  DevBuf *outs[] = {&c->g_framed, &c->g_scratch, &c->cc_mask};
  for (DevBuf *g : outs) AOS_CUDA_OK(c, g->reserve(gbytes));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  cudaStream_t st = c->stream;
  c->marks.clear();
  c->mark("start");
  if (skeleton_bits != c->g_skel.as<uint32_t>()) {
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_scratch.p, skeleton_bits, gbytes, cudaMemcpyDeviceToDevice, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    AOS_CUDA_OK(c, c->g_skel.reserve(gbytes));
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_skel.p, c->g_scratch.p, gbytes, cudaMemcpyDeviceToDevice, st));
  }
  if (occupancy_bits && occupancy_bits != c->g_occ.as<uint32_t>()) {
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_scratch.p, occupancy_bits, gbytes, cudaMemcpyDeviceToDevice, st));
    AOS_CUDA_OK(c, cudaStreamSynchronize(st));
    AOS_CUDA_OK(c, c->g_occ.reserve(gbytes));
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->g_occ.p, c->g_scratch.p, gbytes, cudaMemcpyDeviceToDevice, st));
  }
  c->partial_grids = true;
  c->have_occ = occupancy_bits != nullptr;
  c->mark("grids_in");
  return seed_tail(c, p, gi, launches, launches * kThinSubIters, nullptr);
}

aos_status aos_seed_summary_get(aos_ctx *c, aos_seed_summary *out) {
  if (!c || !out) return AOS_ERR_INVALID;
  if (!c->have_seed) {
    set_error(c, "aos_seed_stage has not completed");
    return AOS_ERR_STATE;
  }
  *out = c->summary;
  return AOS_OK;
}

aos_status aos_get_grid(aos_ctx *c, aos_grid_id which, aos_grid_fmt fmt, void *dst, size_t dst_bytes, aos_mem dst_mem) {
  if (!c || !dst) return AOS_ERR_INVALID;
  if (!c->have_seed) {
    set_error(c, "aos_seed_stage has not completed");
    return AOS_ERR_STATE;
  }
  DevBuf *g = grid_buf(c, which);
  AOS_REQUIRE(c, g != nullptr, "unknown grid id");
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  if (c->partial_grids && !(which == AOS_GRID_SKELETON || which == AOS_GRID_SKELETON_FRAMED ||
                            (which == AOS_GRID_OCCUPANCY && c->have_occ))) {
    set_error(c, "this grid was not produced on this context (aos_seed_stage_tail)");
    return AOS_ERR_STATE;
  }
  const SeedDeviceParams &P = c->P;
  cudaStream_t st = c->stream;
  if (fmt == AOS_FMT_BITS) {
    size_t need = (size_t)P.pitch * P.h * 4;
    if (dst_bytes < need) return AOS_ERR_CAPACITY;
    AOS_CUDA_OK(c, cudaMemcpyAsync(dst, g->p, need, dst_mem == AOS_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  } else {
    size_t need = (size_t)P.w * P.h;
    if (dst_bytes < need) return AOS_ERR_CAPACITY;
    if (dst_mem == AOS_MEM_DEVICE) {
      aos_status s = launch_unpack(c, g->as<uint32_t>(), static_cast<int8_t *>(dst), P.w, P.h);
      if (s != AOS_OK) return s;
    } else {
      AOS_CUDA_OK(c, c->points_stage.reserve(need));  // reuse the staging buffer for the byte image
      aos_status s = launch_unpack(c, g->as<uint32_t>(), c->points_stage.as<int8_t>(), P.w, P.h);
      if (s != AOS_OK) return s;
      AOS_CUDA_OK(c, cudaMemcpyAsync(dst, c->points_stage.p, need, cudaMemcpyDeviceToHost, st));
    }
  }
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  return AOS_OK;
}

aos_status aos_grid_device_bits(aos_ctx *c, aos_grid_id which, const uint32_t **bits, int32_t *pitch_words) {
  if (!c || !bits) return AOS_ERR_INVALID;
  if (!c->have_seed) return AOS_ERR_STATE;
  DevBuf *g = grid_buf(c, which);
  AOS_REQUIRE(c, g != nullptr, "unknown grid id");
  *bits = g->as<uint32_t>();
  if (pitch_words) *pitch_words = c->P.pitch;
  return AOS_OK;
}

aos_status aos_get_labels(aos_ctx *c, int32_t *dst, size_t dst_count, aos_mem dst_mem) {
  if (!c || !dst) return AOS_ERR_INVALID;
  if (!c->have_seed) return AOS_ERR_STATE;
  size_t cells = (size_t)c->P.w * c->P.h;
  if (dst_count < cells) return AOS_ERR_CAPACITY;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  if (dst_mem == AOS_MEM_DEVICE) {
    aos_status s = launch_labels(c, dst);
    if (s != AOS_OK) return s;
  } else {
    AOS_CUDA_OK(c, c->points_stage.reserve(cells * 4));
    aos_status s = launch_labels(c, c->points_stage.as<int32_t>());
    if (s != AOS_OK) return s;
    AOS_CUDA_OK(c, cudaMemcpyAsync(dst, c->points_stage.p, cells * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_get_clusters(aos_ctx *c, aos_cluster *dst, int32_t capacity, int32_t *n_out) {
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_seed) return AOS_ERR_STATE;
  int n = static_cast<int>(c->h_clusters.size());
  if (n_out) *n_out = n;
  if (!dst) return AOS_OK;
  if (capacity < n) return AOS_ERR_CAPACITY;
  if (n) memcpy(dst, c->h_clusters.data(), sizeof(aos_cluster) * (size_t)n);
  return AOS_OK;
}

aos_status aos_get_tree_rows(aos_ctx *c, aos_tree_row *dst, int32_t capacity, int32_t *n_out) {
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_seed) return AOS_ERR_STATE;
  int n = static_cast<int>(c->h_rows.size());
  if (n_out) *n_out = n;
  if (!dst) return AOS_OK;
  if (capacity < n) return AOS_ERR_CAPACITY;
  if (n) memcpy(dst, c->h_rows.data(), sizeof(aos_tree_row) * (size_t)n);
  return AOS_OK;
}

aos_status aos_get_launch_count(aos_ctx *c, int64_t *out) {
  if (!c || !out) return AOS_ERR_INVALID;
  *out = c->launches;
  return AOS_OK;
}

aos_status aos_select_seeds(aos_ctx *c, int32_t *n_seeds, int32_t counts[3]) {
  NvtxRange nvtx_range("aos_select_seeds");
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_seed) {
    set_error(c, "aos_seed_stage has not completed");
    return AOS_ERR_STATE;
  }
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  aos_status s = device_select_seeds(c);  // ray casts + first-come filters: k_seeds.cu
  if (s != AOS_OK) return s;
  host_rows_info(c->h_rows, &c->h_rows_info);
  c->have_seeds = true;
  c->mark("select_seeds");
  if (n_seeds) *n_seeds = (int32_t)(c->h_seeds.size() / 2);
  if (counts)
    for (int k = 0; k < 3; ++k) counts[k] = c->seed_counts[k];
  return AOS_OK;
}

aos_status aos_get_seeds(aos_ctx *c, double *dst, int32_t capacity, int32_t *n_out) {
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_seeds) return AOS_ERR_STATE;
  int n = (int)(c->h_seeds.size() / 2);
  if (n_out) *n_out = n;
  if (!dst) return AOS_OK;
  if (capacity < n) return AOS_ERR_CAPACITY;
  if (n) memcpy(dst, c->h_seeds.data(), sizeof(double) * 2 * (size_t)n);
  return AOS_OK;
}

aos_status aos_get_rows_info(aos_ctx *c, double *dst, int32_t capacity_rows, int32_t *n_out) {
  if (!c) return AOS_ERR_INVALID;
  if (!c->have_seeds) return AOS_ERR_STATE;
  int n = (int)(c->h_rows_info.size() / 4);
  if (n_out) *n_out = n;
  if (!dst) return AOS_OK;
  if (capacity_rows < n) return AOS_ERR_CAPACITY;
  if (n) memcpy(dst, c->h_rows_info.data(), sizeof(double) * 4 * (size_t)n);
  return AOS_OK;
}

// ---- stand-alone steps ---------------------------------------------------------------------------
aos_status aos_inflate_bits(aos_ctx *c, const uint32_t *in, uint32_t *out, uint32_t *out_border, int32_t w, int32_t h,
                            int32_t radius_cells) {
  if (!c || !in || !out) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  return launch_inflate(c, in, out, out_border, w, h, radius_cells);
}

aos_status aos_open_bits(aos_ctx *c, const uint32_t *in, uint32_t *out, int32_t w, int32_t h) {
  if (!c || !in || !out) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  return launch_open(c, in, out, w, h);
}

aos_status aos_thin_bits(aos_ctx *c, uint32_t *inout, int32_t w, int32_t h, int32_t *launches, int32_t *subiters) {
  if (!c || !inout) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  size_t gbytes = (size_t)pitch_words_for(w) * h * 4;
  AOS_CUDA_OK(c, c->g_scratch.reserve(gbytes));
  AOS_CUDA_OK(c, cudaMemsetAsync(c->g_scratch.p, 0, gbytes, c->stream));
  int l = 0, s = 0;
  aos_status r = launch_thin(c, inout, c->g_scratch.as<uint32_t>(), w, h, &l, &s);
  if (launches) *launches = l;
  if (subiters) *subiters = s;
  if (r != AOS_OK) return r;
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_radius_outlier_removal(aos_ctx *c, const void *points, size_t n_points, uint32_t point_step, uint32_t off_x,
                                      uint32_t off_y, uint32_t off_z, aos_mem points_mem, float radius, int32_t min_neighbors,
                                      const void **out_points, size_t *n_out) {
  NvtxRange nvtx_range("aos_radius_outlier_removal");
  if (!c || !out_points || !n_out) return AOS_ERR_INVALID;
  aos_status s = check_points(c, points, n_points, point_step, off_x, off_y, off_z);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->marks.clear();
  c->mark("start");
  const void *dpoints = points;
  if (points_mem == AOS_MEM_HOST && n_points) {
    s = upload_points(c, points, n_points * (size_t)point_step, &dpoints);
    if (s != AOS_OK) return s;
  }
  c->mark("h2d_points");
  s = run_ror(c, dpoints, n_points, point_step, off_x, off_y, off_z, radius, min_neighbors, n_out);
  if (s != AOS_OK) return s;
  *out_points = c->ror_out.p;
  return AOS_OK;
}

aos_status aos_set_clearance(aos_ctx *c, int enabled) {
  if (!c) return AOS_ERR_INVALID;
  c->clearance = enabled != 0;
  return AOS_OK;
}

// trimPathNearOccupiedRegions (src/aos_path_gen_node.cpp:1570-1630), SURVEY section 8(f) row F3
aos_status aos_trim_path(aos_ctx *c, const double *path_xy, int32_t n_poses, double safety_distance,
                         const uint32_t *skeleton_bits, aos_mem skeleton_mem, const aos_grid_info *info, int32_t *n_kept) {
  if (!c || !n_kept || n_poses < 0 || (n_poses > 0 && !path_xy)) return AOS_ERR_INVALID;
  AOS_REQUIRE(c, (skeleton_bits == nullptr) == (info == nullptr), "skeleton and info must be given together");
  AOS_REQUIRE(c, safety_distance >= 0 && std::isfinite(safety_distance), "bad safety distance");
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  *n_kept = n_poses;
  const uint32_t *bits = nullptr;
  int w, h;
  double ox, oy;
  float res;
  if (skeleton_bits) {
    AOS_REQUIRE(c, info->width > 0 && info->height > 0 && info->resolution > 0.f, "bad grid info");
    w = info->width;
    h = info->height;
    ox = info->origin_x;
    oy = info->origin_y;
    res = info->resolution;
    if (skeleton_mem == AOS_MEM_DEVICE) {
      bits = skeleton_bits;
    } else {
      const size_t bytes = (size_t)pitch_words_for(w) * h * 4;
      AOS_CUDA_OK(c, c->gvd_skel.reserve(bytes));
      AOS_CUDA_OK(c, cudaMemcpyAsync(c->gvd_skel.p, skeleton_bits, bytes, cudaMemcpyHostToDevice, c->stream));
      bits = c->gvd_skel.as<uint32_t>();
    }
  } else {
    if (!c->have_seed) {  // path_gen:1571: no skeleton yet -> the path is left as it is
      set_error(c, "no skeleton: pass one or run aos_seed_stage on this context first");
      return AOS_ERR_STATE;
    }
    w = c->P.w;
    h = c->P.h;
    ox = c->P.ox;
    oy = c->P.oy;
    res = c->P.res;
    bits = c->g_framed.as<uint32_t>();
  }
  int kept = n_poses;
  aos_status s = launch_trim_path(c, path_xy, n_poses, bits, w, h, ox, oy, res, safety_distance, &kept);
  if (s != AOS_OK) return s;
  *n_kept = kept;
  return AOS_OK;
}

aos_status aos_edt_bits(aos_ctx *c, const uint32_t *bits, int32_t w, int32_t h, uint32_t *nearest_xy, int32_t *dist2) {
  if (!c || !bits || !nearest_xy) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->marks.clear();
  c->mark("start");
  aos_status r = launch_edt(c, bits, w, h, nearest_xy, dist2);
  if (r != AOS_OK) return r;
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_inflate_bits_edt(aos_ctx *c, const uint32_t *in, uint32_t *out, int32_t w, int32_t h, int32_t radius_cells) {
  if (!c || !in || !out || radius_cells < 0) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  const size_t cells = (size_t)w * h;
  AOS_CUDA_OK(c, c->edt_out.reserve(cells * 8 + 1024));
  uint32_t *nearest = c->edt_out.as<uint32_t>();
  int32_t *d2 = reinterpret_cast<int32_t *>(nearest + cells);
  aos_status r = launch_edt(c, in, w, h, nearest, d2);
  if (r != AOS_OK) return r;
  r = launch_edt_threshold(c, d2, w, h, radius_cells * radius_cells, out);
  if (r != AOS_OK) return r;
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_pack_int8(aos_ctx *c, const int8_t *src, aos_mem src_mem, uint32_t *dst_bits, int32_t w, int32_t h) {
  if (!c || !src || !dst_bits) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  const int8_t *d = src;
  if (src_mem == AOS_MEM_HOST) {
    AOS_CUDA_OK(c, c->points_stage.reserve((size_t)w * h));
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->points_stage.p, src, (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
    d = c->points_stage.as<int8_t>();
  }
  aos_status r = launch_pack(c, d, dst_bits, w, h);
  if (r != AOS_OK) return r;
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

aos_status aos_unpack_int8(aos_ctx *c, const uint32_t *src_bits, int8_t *dst, aos_mem dst_mem, int32_t w, int32_t h) {
  if (!c || !src_bits || !dst) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  if (dst_mem == AOS_MEM_DEVICE) {
    aos_status r = launch_unpack(c, src_bits, dst, w, h);
    if (r != AOS_OK) return r;
  } else {
    AOS_CUDA_OK(c, c->points_stage.reserve((size_t)w * h));
    aos_status r = launch_unpack(c, src_bits, c->points_stage.as<int8_t>(), w, h);
    if (r != AOS_OK) return r;
    AOS_CUDA_OK(c, cudaMemcpyAsync(dst, c->points_stage.p, (size_t)w * h, cudaMemcpyDeviceToHost, c->stream));
  }
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

}  // extern "C"
