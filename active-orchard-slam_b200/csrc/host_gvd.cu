// host_gvd.cu -- C-ABI of the gvd half (include/aos_gpu.h): aos_gvd_stage replaces the body of
// AosGvdNode::processGraph (src/aos_gvd_node.cpp:255-318) plus the seed merge of voronoiSeedsCallback
// (gvd:84-128).  Host side: seed merge, VoronoiDiagram::compute (src/utils/voronoi_diagram.cpp:16-114, the
// Subdiv2D replay of host_subdiv.cu); device side: k_graph.cu.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>

#include "aos_common.cuh"
#include "host_subdiv.h"

namespace aos {

// VoronoiDiagram::compute, vd:16-94.
// VoronoiDiagram::compute up to the last insert (vd:16-92): bounds, Subdiv2D(rect), one insert per seed in order.
// `before_reserve(n)` runs ahead of the only (re)allocation of sd's arrays (the context un-pins them there).
template <typename F>
static bool host_subdiv_build(Subdiv &sd, const double *seeds, int n, double min_x, double max_x, double min_y, double max_y,
                              F before_reserve) {
  if (n <= 0) return false;
  if (!std::isfinite(min_x) || !std::isfinite(max_x) || !std::isfinite(min_y) || !std::isfinite(max_y)) return false;
  if (min_x > max_x) std::swap(min_x, max_x);
  if (min_y > max_y) std::swap(min_y, max_y);
  const double min_size = 1.0;
  if (max_x - min_x < min_size) {
    double c = (min_x + max_x) / 2.0;
    min_x = c - min_size / 2.0;
    max_x = c + min_size / 2.0;
  }
  if (max_y - min_y < min_size) {
    double c = (min_y + max_y) / 2.0;
    min_y = c - min_size / 2.0;
    max_y = c + min_size / 2.0;
  }
  // cv::Rect2f bounding_rect (vd:51-56)
  const float rx = (float)(min_x - 1.0), ry = (float)(min_y - 1.0);
  const float rw = (float)(fabs(max_x - min_x) + 2.0), rh = (float)(fabs(max_y - min_y) + 2.0);
  if (rw <= 0 || rh <= 0) return false;
  // cv::Subdiv2D(Rect): the Rect2f converts to the int Rect through saturate_cast<int> == cvRound
  // (round-half-even), OpenCV 4.5.4 as shipped with ROS 2 Humble (package.xml:48)
  const bool dbg = getenv("AOS_DEBUG") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  // room for twice the seeds when the arrays have to grow at all (re-allocation also means un-pinning and pinning
  // again, which stalls every map in flight), and never less than 4096 seeds
  size_t want = std::max<size_t>((size_t)n, 4096);
  if (3 * want + 16 > sd.quad_capacity() || 3 * want + 16 > sd.vertex_capacity()) want *= 2;
  before_reserve(want);
  sd.reserve(want);
  sd.init((int)lrint((double)rx), (int)lrint((double)ry), (int)lrint((double)rw), (int)lrint((double)rh));
  const float margin = 0.1f;
  for (int i = 0; i < n; ++i) {
    double sx = seeds[2 * i], sy = seeds[2 * i + 1];
    if (!std::isfinite(sx) || !std::isfinite(sy)) continue;
    float x = (float)sx, y = (float)sy;
    x = std::max(rx + margin, std::min(rx + rw - margin, x));
    y = std::max(ry + margin, std::min(ry + rh - margin, y));
    sd.insert(x, y);  // -1 where cv::Subdiv2D::insert throws: the reference skips the seed (vd:83-88)
  }
  if (dbg) {
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[aos] subdiv: %d seeds, insert %.1f ms\n", n, std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  return true;
}

// ... and getVoronoiFacetList on the host (vd:94): the stand-alone aos_voronoi_facets and the fallback of the device walk
bool host_voronoi_facets(const double *seeds, int n, double min_x, double max_x, double min_y, double max_y,
                         std::vector<float> *facet_xy, std::vector<int32_t> *facet_off) {
  facet_xy->clear();
  facet_off->assign(1, 0);
  Subdiv sd;
  if (!host_subdiv_build(sd, seeds, n, min_x, max_x, min_y, max_y, [](size_t) {})) return false;
  sd.voronoi_facets(facet_xy, facet_off);
  return true;
}

// Page-lock the storage of the context's Subdiv arrays so their upload is one DMA each (cudaHostRegister; a failure
// only means a staged copy).  Called with the arrays' current storage after reserve(); sd_unpin_if_growing runs
// before a reserve() that would re-allocate, because registered memory must not be freed.
static void sd_unpin(Ctx *c, int i) {
  if (c->sd_pinned[i]) cudaHostUnregister(c->sd_pinned[i]);
  c->sd_pinned[i] = nullptr;
  c->sd_pinned_bytes[i] = 0;
}
static void sd_pin(Ctx *c, int i, const void *p, size_t bytes) {
  if (c->sd_pinned[i] == p && c->sd_pinned_bytes[i] == bytes) return;
  sd_unpin(c, i);
  if (p && bytes && cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterMapped) == cudaSuccess) {
    c->sd_pinned[i] = const_cast<void *>(p);
    c->sd_pinned_bytes[i] = bytes;
  } else {
    cudaGetLastError();  // not fatal
  }
}
void subdiv_release_pins(Ctx *c) {
  for (int i = 0; i < 3; ++i) sd_unpin(c, i);
}
static void sd_unpin_if_growing(Ctx *c, const Subdiv &sd, size_t n) {
  if (3 * n + 16 > sd.quad_capacity()) {
    sd_unpin(c, 0);
    sd_unpin(c, 1);
  }
  if (3 * n + 16 > sd.vertex_capacity()) sd_unpin(c, 2);
}
static void sd_pin_all(Ctx *c, const Subdiv &sd) {
  sd_pin(c, 0, sd.edge_next(), sd.quad_capacity() * 4 * sizeof(int));
  sd_pin(c, 1, sd.edge_pt(), sd.quad_capacity() * 4 * sizeof(int));
  sd_pin(c, 2, sd.vertices(), sd.vertex_capacity() * sizeof(Subdiv::Vertex));
}

}  // namespace aos

using namespace aos;
struct aos_ctx : public aos::Ctx {};

extern "C" {

aos_status aos_set_subdiv_outer_factor(float factor) {
  if (!(factor >= 1.f) || !std::isfinite(factor)) return AOS_ERR_INVALID;
  aos::g_outer_factor = factor;
  return AOS_OK;
}

aos_status aos_set_subdiv_literal_splices(int32_t on) {
  aos::g_literal_splices = on != 0;
  return AOS_OK;
}
aos_status aos_set_subdiv_simd(int32_t mode) {
  if (mode < -1 || mode > 1) return AOS_ERR_INVALID;
  if (mode == 1 && !aos::subdiv_simd_available()) return AOS_ERR_INVALID;
  aos::g_subdiv_simd = mode;
  return AOS_OK;
}


aos_status aos_merge_seeds(const double *seeds_xy, int32_t n, double *out_xy, int32_t *n_out) {
  if (n < 0 || (n > 0 && (!seeds_xy || !out_xy))) return AOS_ERR_INVALID;
  std::vector<double> out;
  host_merge_seeds(seeds_xy, n, &out);
  if (!out.empty()) memcpy(out_xy, out.data(), sizeof(double) * out.size());
  if (n_out) *n_out = (int32_t)(out.size() / 2);
  return AOS_OK;
}

aos_status aos_merge_seeds_device(aos_ctx *c, const double *seeds_xy, int32_t n, double *out_xy, int32_t *n_out) {
  if (!c || n < 0 || (n > 0 && (!seeds_xy || !out_xy))) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  aos_status s = device_merge_seeds(c, seeds_xy, n);
  if (s != AOS_OK) return s;
  if (c->h_merged.size()) memcpy(out_xy, c->h_merged.data(), sizeof(double) * c->h_merged.size());
  if (n_out) *n_out = (int32_t)(c->h_merged.size() / 2);
  return AOS_OK;
}

aos_status aos_voronoi_facets(const double *seeds_xy, int32_t n_seeds, double min_x, double max_x, double min_y,
                              double max_y, float *facet_xy, int32_t xy_capacity_points, int32_t *facet_off,
                              int32_t off_capacity, int32_t *n_facets, int32_t *n_points) {
  if (n_seeds < 0 || (n_seeds > 0 && !seeds_xy)) return AOS_ERR_INVALID;
  std::vector<float> xy;
  std::vector<int32_t> off;
  host_voronoi_facets(seeds_xy, n_seeds, min_x, max_x, min_y, max_y, &xy, &off);
  if (n_facets) *n_facets = (int32_t)off.size() - 1;
  if (n_points) *n_points = (int32_t)(xy.size() / 2);
  if (!facet_xy && !facet_off) return AOS_OK;
  if (!facet_xy || !facet_off || xy_capacity_points < (int32_t)(xy.size() / 2) || off_capacity < (int32_t)off.size())
    return AOS_ERR_CAPACITY;
  if (!xy.empty()) memcpy(facet_xy, xy.data(), sizeof(float) * xy.size());
  memcpy(facet_off, off.data(), sizeof(int32_t) * off.size());
  return AOS_OK;
}

// The device half of VoronoiDiagram::compute on its own (parity tests against aos_voronoi_facets): insertions on the
// host, circumcentres + facet walks on the device; returns the facet-vertex slots (facets with fewer than 2 vertices
// dropped, as vd:97-114 does) and, per slot, the slot of the next vertex of the same facet.
aos_status aos_voronoi_facets_device(aos_ctx *c, const double *seeds_xy, int32_t n_seeds, double min_x, double max_x,
                                     double min_y, double max_y, float *slot_xy, int32_t *slot_next, int32_t capacity_slots,
                                     int32_t *n_slots) {
  if (!c || !n_slots) return AOS_ERR_INVALID;
  if (n_seeds < 0 || (n_seeds > 0 && !seeds_xy)) return AOS_ERR_INVALID;
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  *n_slots = 0;
  Subdiv &sd = c->subdiv;
  if (!host_subdiv_build(sd, seeds_xy, n_seeds, min_x, max_x, min_y, max_y, [c, &sd](size_t n) { sd_unpin_if_growing(c, sd, n); }))
    return AOS_OK;
  sd_pin_all(c, sd);
  int K = -1;
  aos_status s = facets_prepare(c, sd, &K);
  if (s != AOS_OK) return s;
  if (K < 0) {
    set_error(c, "subdivision is not a triangulation");
    return AOS_ERR_STATE;
  }
  *n_slots = K;
  if (!slot_xy && !slot_next) return AOS_OK;
  if (!slot_xy || !slot_next || capacity_slots < K) return AOS_ERR_CAPACITY;
  if (K == 0) return AOS_OK;
  AOS_CUDA_OK(c, c->gvd_buf.reserve(12 * (size_t)K + 512));
  float2 *d_fxy = c->gvd_buf.as<float2>();
  int *d_enext = reinterpret_cast<int *>(c->gvd_buf.as<char>() + ((8 * (size_t)K + 255) & ~(size_t)255));
  s = facets_fill(c, d_fxy, d_enext);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(slot_xy, d_fxy, 8 * (size_t)K, cudaMemcpyDeviceToHost, c->stream));
  AOS_CUDA_OK(c, cudaMemcpyAsync(slot_next, d_enext, 4 * (size_t)K, cudaMemcpyDeviceToHost, c->stream));
  AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return AOS_OK;
}

static aos_status gvd_stage_impl(aos_ctx *c, const double *seeds_xy, int32_t n_seeds, const double *rows_info, int32_t n_rows,
                                 const void *skeleton, bool skeleton_is_bits, aos_mem skeleton_mem, const aos_grid_info *info);

aos_status aos_gvd_stage(aos_ctx *c, const double *seeds_xy, int32_t n_seeds, const double *rows_info, int32_t n_rows,
                         const int8_t *skeleton, const aos_grid_info *info) {
  NvtxRange nvtx_range("aos_gvd_stage");
  return gvd_stage_impl(c, seeds_xy, n_seeds, rows_info, n_rows, skeleton, false, AOS_MEM_HOST, info);
}

aos_status aos_gvd_stage_bits(aos_ctx *c, const double *seeds_xy, int32_t n_seeds, const double *rows_info, int32_t n_rows,
                              const uint32_t *skeleton_bits, aos_mem skeleton_mem, const aos_grid_info *info) {
  NvtxRange nvtx_range("aos_gvd_stage_bits");
  if (!skeleton_bits || !info) return AOS_ERR_INVALID;
  return gvd_stage_impl(c, seeds_xy, n_seeds, rows_info, n_rows, skeleton_bits, true, skeleton_mem, info);
}

static aos_status gvd_stage_impl(aos_ctx *c, const double *seeds_xy, int32_t n_seeds, const double *rows_info, int32_t n_rows,
                                 const void *skeleton, bool skeleton_is_bits, aos_mem skeleton_mem, const aos_grid_info *info) {
  if (!c) return AOS_ERR_INVALID;
  AOS_REQUIRE(c, n_seeds >= 0 && n_rows >= 0, "negative count");
  AOS_REQUIRE(c, n_seeds == 0 || seeds_xy != nullptr, "seeds pointer is null");
  AOS_REQUIRE(c, n_rows == 0 || rows_info != nullptr, "rows pointer is null");
  AOS_REQUIRE(c, (skeleton == nullptr) == (info == nullptr), "skeleton and info must be given together");
  AOS_CUDA_OK(c, cudaSetDevice(c->device));
  c->have_graph = false;
  if (!c->composite) {
    c->marks.clear();
    c->mark("start");
  }

  GraphInputs in;
  if (skeleton) {
    AOS_REQUIRE(c, info->width > 0 && info->height > 0 && info->resolution > 0.f, "bad grid info");
    in.w = info->width;
    in.h = info->height;
    in.pitch = pitch_words_for(in.w);
    in.ox = info->origin_x;
    in.oy = info->origin_y;
    in.res = info->resolution;
    const size_t cells = (size_t)in.w * in.h, bit_bytes = (size_t)in.pitch * in.h * 4;
    if (skeleton_is_bits && skeleton_mem == AOS_MEM_DEVICE) {
      in.skel_bits = static_cast<const uint32_t *>(skeleton);  // shared device handle: no copy at all
    } else {
      AOS_CUDA_OK(c, c->gvd_skel.reserve(bit_bytes));
      if (skeleton_is_bits) {  // bit-packed side channel: W*H/8 bytes instead of W*H
        AOS_CUDA_OK(c, cudaMemcpyAsync(c->gvd_skel.p, skeleton, bit_bytes, cudaMemcpyHostToDevice, c->stream));
      } else {
        AOS_CUDA_OK(c, c->points_stage.reserve(cells));
        AOS_CUDA_OK(c, cudaMemcpyAsync(c->points_stage.p, skeleton, cells, cudaMemcpyHostToDevice, c->stream));
        aos_status s = launch_pack(c, c->points_stage.as<int8_t>(), c->gvd_skel.as<uint32_t>(), in.w, in.h);
        if (s != AOS_OK) return s;
      }
      in.skel_bits = c->gvd_skel.as<uint32_t>();
    }
  } else {
    if (!c->have_seed) {  // gvd:257: no skeleton yet
      set_error(c, "no skeleton: pass one or run aos_seed_stage on this context first");
      return AOS_ERR_STATE;
    }
    in.w = c->P.w;
    in.h = c->P.h;
    in.pitch = c->P.pitch;
    in.ox = c->P.ox;
    in.oy = c->P.oy;
    in.res = c->P.res;
    in.skel_bits = c->g_framed.as<uint32_t>();
  }
  c->mark("gvd_skeleton_in");

  // voronoiSeedsCallback merge (gvd:93-125); processGraph drops non-finite seeds (gvd:266-270)
  static const bool dbg = getenv("AOS_DEBUG") != nullptr;
  auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tg0 = dbg ? wall() : 0;
  {
    aos_status ms = device_merge_seeds(c, seeds_xy, n_seeds);
    if (ms != AOS_OK) return ms;
  }
  const double tg1 = dbg ? wall() : 0;
  c->mark("gvd_merge_seeds");
  c->graph.n_merged_seeds = (int)(c->h_merged.size() / 2);
  if (c->h_merged.size() == 0) {  // gvd:257 / :273: nothing to do
    set_error(c, "no valid seeds");
    return AOS_ERR_STATE;
  }
  // bounds, gvd:278-281: uint32 * float -> float, widened
  const double minx = in.ox, maxx = in.ox + (double)(float)((float)(unsigned)in.w * in.res);
  const double miny = in.oy, maxy = in.oy + (double)(float)((float)(unsigned)in.h * in.res);
  // VoronoiDiagram::compute: the Delaunay insertions are a sequential replay on the host (host_subdiv.cu); the
  // circumcentres and the facet walks are data-parallel and run on the device (k_facets.cu)
  if (!c->pin_rows.resize(4 * (size_t)n_rows)) {
    set_error(c, "cudaHostAlloc failed for the row staging buffer");
    return AOS_ERR_CUDA;
  }
  if (n_rows) memcpy(c->pin_rows.data(), rows_info, sizeof(double) * 4 * (size_t)n_rows);
  in.rows_info = c->pin_rows.data();
  in.n_rows = n_rows;
  if (c->voronoi_mode == AOS_VORONOI_DEVICE) {
    // opt-in: the Voronoi cells of the same point set, one thread per seed, no insertion replay (k_vcells.cu).  Same
    // diagram; vertex low bits and facet starts are not Subdiv2D's, so the graph is not bit-identical to the reference's.
    int cell_slots = 0;
    aos_status vs = vcells_prepare(c, c->h_merged.data(), (int)(c->h_merged.size() / 2), minx, maxx, miny, maxy, &cell_slots);
    if (vs != AOS_OK) return vs;
    c->mark("gvd_device_voronoi");
    if (cell_slots >= 0) {
      in.device_facets = true;
      in.device_cells = true;
      in.n_slots = cell_slots;
      aos_status s = run_graph(c, in);
      if (s != AOS_OK) return s;
      c->have_graph = true;
      return AOS_OK;
    }
    // a cell overflowed the fixed-size polygons (never seen): fall through to the replay
  }
  Subdiv &sd = c->subdiv;
  const bool built = host_subdiv_build(sd, c->h_merged.data(), (int)(c->h_merged.size() / 2), minx, maxx, miny, maxy,
                                       [c, &sd](size_t n) { sd_unpin_if_growing(c, sd, n); });
  c->mark("gvd_host_voronoi");
  const double tg2 = dbg ? wall() : 0;
  int dev_slots = -1;
  if (built) {
    sd_pin_all(c, sd);
    aos_status fs = facets_prepare(c, sd, &dev_slots);
    if (fs != AOS_OK) return fs;
  } else {
    dev_slots = 0;
  }
  c->mark("gvd_facets");
  const double tg3 = dbg ? wall() : 0;
  if (dev_slots >= 0) {
    in.device_facets = true;
    in.n_slots = dev_slots;
    aos_status s = run_graph(c, in);
    if (s != AOS_OK) return s;
    if (dbg) fprintf(stderr, "[aos] gvd split: merge %.1f facets %.1f graph %.1f\n", tg1 - tg0, tg3 - tg2, wall() - tg3);
    c->have_graph = true;
    return AOS_OK;
  }
  // the structure is not a triangulation (never seen; Subdiv2D would have walked it all the same): host walk
  std::vector<float> fxy;
  std::vector<int32_t> foff;
  fxy.clear();
  foff.assign(1, 0);
  sd.voronoi_facets(&fxy, &foff);
  // facets -> edge slots (vd:97-114); facets with fewer than 2 vertices contribute nothing
  const int nf = (int)foff.size() - 1;
  if (!c->pin_facet_xy.resize(fxy.size()) || !c->pin_enext.resize(fxy.size() / 2)) {
    set_error(c, "cudaHostAlloc failed for the facet staging buffers");
    return AOS_ERR_CUDA;
  }
  int slots = 0;
  for (int f = 0; f < nf; ++f) {
    const int b = foff[f], k = foff[f + 1] - b;
    if (k < 2) continue;
    const int base = slots;
    memcpy(c->pin_facet_xy.data() + 2 * (size_t)base, fxy.data() + 2 * (size_t)b, sizeof(float) * 2 * (size_t)k);
    int *en = c->pin_enext.data() + base;
    for (int j = 0; j < k; ++j) en[j] = base + j + 1;
    en[k - 1] = base;
    slots += k;
  }
  in.facet_xy = c->pin_facet_xy.data();
  in.enext = c->pin_enext.data();
  in.n_slots = slots;
  aos_status s = run_graph(c, in);
  if (s != AOS_OK) return s;
  c->have_graph = true;
  return AOS_OK;
}

aos_status aos_get_graph(aos_ctx *c, aos_gvd_graph *out) {
  if (!c || !out) return AOS_ERR_INVALID;
  if (!c->have_graph) {
    set_error(c, "aos_gvd_stage has not completed");
    return AOS_ERR_STATE;
  }
  const GraphHost &G = c->graph;
  memset(out, 0, sizeof(*out));
  out->resolution = G.resolution;
  out->origin_x = G.origin_x;
  out->origin_y = G.origin_y;
  out->n_nodes = (int32_t)G.node_labels.size();
  out->nodes_xyz = G.nodes_xyz.data();
  out->node_labels = G.node_labels.data();
  out->node_cluster_indices = G.node_cluster_indices.data();
  out->node_label_counts = G.node_label_counts.data();
  out->n_label_entries = (int32_t)G.node_label_clusters.size();
  out->node_label_clusters = G.node_label_clusters.data();
  out->node_label_types = G.node_label_types.data();
  out->n_edges = (int32_t)G.edge_lengths.size();
  out->edges = G.edges.data();
  out->edge_lengths = G.edge_lengths.data();
  out->edge_clearances = G.edge_clearances.data();
  out->n_merged_seeds = G.n_merged_seeds;
  out->n_voronoi_edges = G.n_voronoi_edges;
  out->n_boundary_points = G.n_boundary_points;
  out->corner_points = G.corner_points.data();
  out->n_rows = G.n_rows;
  return AOS_OK;
}

aos_status aos_map_to_graph(aos_ctx *c, const aos_seed_params *p, const void *points, size_t n_points, uint32_t point_step,
                            uint32_t off_x, uint32_t off_y, uint32_t off_z, aos_mem points_mem) {
  NvtxRange nvtx_range("aos_map_to_graph");
  if (!c) return AOS_ERR_INVALID;
  static const bool dbg = getenv("AOS_DEBUG") != nullptr;
  static const auto t_proc = std::chrono::steady_clock::now();
  auto now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_proc).count(); };
  const double t0 = dbg ? now() : 0;
  aos_status s = aos_seed_stage(c, p, points, n_points, point_step, off_x, off_y, off_z, points_mem);
  if (s != AOS_OK) return s;
  const double t1 = dbg ? now() : 0;
  c->composite = true;  // keep one list of stage timers for the whole call
  s = aos_select_seeds(c, nullptr, nullptr);
  const double t2 = dbg ? now() : 0;
  if (s == AOS_OK)
    s = aos_gvd_stage(c, c->h_seeds.data(), (int32_t)(c->h_seeds.size() / 2), c->h_rows_info.data(),
                      (int32_t)(c->h_rows_info.size() / 4), nullptr, nullptr);
  c->composite = false;
  if (dbg)
    fprintf(stderr, "[aos] map ctx %p start %.1f seed_stage %.1f select %.1f gvd %.1f end %.1f\n", (void *)c, t0, t1 - t0, t2 - t1,
            now() - t2, now());
  return s;
}


// BASELINE config 5 (sweeps of independent maps): one host thread per map in flight, each on its own context --
// the GPU stages of one map overlap the host stages (Subdiv2D replay) of the others.
aos_status aos_map_to_graph_batch(aos_batch_item *items, int32_t n_items, int32_t max_threads) {
  if (n_items < 0 || (n_items > 0 && !items)) return AOS_ERR_INVALID;
  for (int i = 0; i < n_items; ++i) {
    if (!items[i].ctx) return AOS_ERR_INVALID;
    for (int j = 0; j < i; ++j)
      if (items[j].ctx == items[i].ctx && max_threads != 1) return AOS_ERR_INVALID;  // a context is not thread-safe
  }
  int nt = max_threads > 0 ? std::min(max_threads, n_items) : n_items;
  // stagger the kernel phases of the maps in flight (see aos_set_device_gate) for the duration of this call only
  const int32_t gate_before = aos_get_device_gate();
  if (nt >= 4 && gate_before == 0) aos_set_device_gate(2);
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n_items; i = next.fetch_add(1)) {
      aos_batch_item &it = items[i];
      it.status = aos_map_to_graph(it.ctx, it.params, it.points, it.n_points, it.point_step, it.off_x, it.off_y, it.off_z,
                                   it.points_mem);
    }
  };
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  aos_set_device_gate(gate_before);
  for (int i = 0; i < n_items; ++i)
    if (items[i].status != AOS_OK && items[i].status != AOS_ERR_STATE) return items[i].status;
  return AOS_OK;
}

}  // extern "C"
