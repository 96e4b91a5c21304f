// k_facets.cu -- cv::Subdiv2D::calcVoronoi + getVoronoiFacetList on the device.
//
// The incremental Delaunay insertion (host_subdiv.cu) is sequential by definition of the reference's result; what
// follows it is not: VoronoiDiagram::compute (src/utils/voronoi_diagram.cpp:94-114) asks Subdiv2D for one polygon
// per seed, and Subdiv2D derives those from the finished quad-edge structure -- a circumcentre per triangle
// (calcVoronoi) and a walk around every vertex (getVoronoiFacetList).  The quad-edge and vertex arrays are uploaded
// as they are (two int arrays indexed by edge id, 32 B per vertex) and three kernels produce the facet-vertex slots + cycle links
// that k_graph.cu consumes, so the 1.4 M facet vertices of a C3 map are never materialised on the host.
//
// Sequential semantics reproduced exactly (host_subdiv.cu Subdiv::calc_voronoi is the restatement of OpenCV's loop):
//  * quads are visited in index order from 4, the LEFT face of edge 4i (slot 3) before its RIGHT face (slot 1); a face
//    gets its point from the FIRST visit for which computeVoronoiPoint succeeds, using that quad's edge and the next
//    one around the face, in float32 differences / sums and a double solve -- so the float32 value of a Voronoi
//    vertex depends on which of the triangle's three quads has the lowest index.  Here one thread per (quad, side)
//    finds the three visits of its face and only the earliest one acts, trying the visits in order;
//  * a face whose three visits all fail (det == 0) keeps "no point": Subdiv2D then reads vertex 0, i.e. (0, 0);
//  * the facet of vertex k starts at rot(firstEdge(k)) and follows NEXT_AROUND_LEFT.
#include <float.h>

#include "aos_common.cuh"

namespace aos {

namespace {
constexpr int kNextAroundLeft = 0x13, kNextAroundRight = 0x31;

// the subdivision as host_subdiv.cu keeps it: next / pt per edge id (4 per quad), 32-byte vertices
struct SdEdges {
  const int *next, *pt;
};
__device__ __forceinline__ int d_get_edge(const SdEdges &q, int edge, int type) {
  edge = q.next[(edge & ~3) + ((edge + type) & 3)];
  return (edge & ~3) + ((edge + (type >> 4)) & 3);
}
__device__ __forceinline__ int d_org(const SdEdges &q, int e) { return q.pt[e]; }
__device__ __forceinline__ int d_dst(const SdEdges &q, int e) { return q.pt[e ^ 2]; }

// computeVoronoiPoint (host_subdiv.cu Subdiv::voronoi_point): float differences and sums, double solve, no FMA
__device__ bool d_voronoi_point(const SdVertex &vo0, const SdVertex &vd0, const SdVertex &vo1, const SdVertex &vd1, float *x,
                                float *y) {
  const float2 o0 = make_float2((float)vo0.x, (float)vo0.y), d0 = make_float2((float)vd0.x, (float)vd0.y);
  const float2 o1 = make_float2((float)vo1.x, (float)vo1.y), d1 = make_float2((float)vd1.x, (float)vd1.y);
  double a0 = __fsub_rn(d0.x, o0.x);
  double b0 = __fsub_rn(d0.y, o0.y);
  double c0 = -0.5 * (a0 * (double)__fadd_rn(d0.x, o0.x) + b0 * (double)__fadd_rn(d0.y, o0.y));
  double a1 = __fsub_rn(d1.x, o1.x);
  double b1 = __fsub_rn(d1.y, o1.y);
  double c1 = -0.5 * (a1 * (double)__fadd_rn(d1.x, o1.x) + b1 * (double)__fadd_rn(d1.y, o1.y));
  double det = a0 * b1 - a1 * b0;
  if (det != 0) {
    det = 1. / det;
    *x = (float)((b0 * c1 - b1 * c0) * det);
    *y = (float)((a1 * c0 - a0 * c1) * det);
  } else {
    *x = FLT_MAX;
    *y = FLT_MAX;
  }
  return fabsf(*x) < FLT_MAX * 0.5f && fabsf(*y) < FLT_MAX * 0.5f;
}

// face-slot id of the face LEFT of edge e (e primal: rot 0 -> slot 3 of its quad, rot 2 -> slot 1)
__device__ __forceinline__ int left_face_slot(int e) { return 2 * (e >> 2) + ((e & 2) ? 1 : 0); }
// ... and RIGHT of e (rot 0 -> slot 1, rot 2 -> slot 3)
__device__ __forceinline__ int right_face_slot(int e) { return 2 * (e >> 2) + ((e & 2) ? 0 : 1); }

// one thread per face slot fs = 2 * quad + side (side 0: left of edge 4*quad, side 1: right of it)
__global__ void vor_points_kernel(const SdEdges q, const SdVertex *__restrict__ vtx, int n_quads,
                                  float2 *__restrict__ vor, int *__restrict__ err) {
  const int fs = blockIdx.x * blockDim.x + threadIdx.x;
  if (fs >= 2 * n_quads || fs < 8) return;  // the loop starts at quad 4
  const int i = fs >> 1, side = fs & 1;
  if (q.next[4 * i] <= 0) return;  // free quad-edge
  const int type = side ? kNextAroundRight : kNextAroundLeft;
  const int e0 = 4 * i;
  const int e1 = d_get_edge(q, e0, type), e2 = d_get_edge(q, e1, type);
  if (d_get_edge(q, e2, type) != e0 || (e1 & 1) || (e2 & 1)) {  // not a triangle: the derivation above does not hold
    atomicExch(err, 1);
    return;
  }
  // the three visits of this face, as (face slot, first edge, walk type): a quad sees the face on the left of its
  // edge 4j (walks NEXT_AROUND_LEFT) or on its right (NEXT_AROUND_RIGHT)
  int slot[3], walk[3];
  slot[0] = fs;
  walk[0] = type;
  const int s1 = side ? right_face_slot(e1) : left_face_slot(e1), s2 = side ? right_face_slot(e2) : left_face_slot(e2);
  slot[1] = s1;
  slot[2] = s2;
  // seen from quad j the face is on the LEFT of 4j iff the slot is even
  walk[1] = (s1 & 1) ? kNextAroundRight : kNextAroundLeft;
  walk[2] = (s2 & 1) ? kNextAroundRight : kNextAroundLeft;
  // only visits from quads >= 4 that are not free act; the earliest of them owns the face
  bool can[3];
  for (int k = 0; k < 3; ++k) can[k] = slot[k] >= 8 && q.next[4 * (slot[k] >> 1)] > 0;
  for (int k = 1; k < 3; ++k)
    if (can[k] && slot[k] < fs) return;
  // try the visits in ascending order (at most three, tiny insertion sort)
  int ord[3] = {0, 1, 2};
  if (slot[ord[1]] > slot[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
  float x = 0.f, y = 0.f;
  bool ok = false;
  for (int n = 0; n < 3 && !ok; ++n) {
    const int k = ord[n];
    if (!can[k]) continue;
    const int f0 = 4 * (slot[k] >> 1), f1 = d_get_edge(q, f0, walk[k]);
    ok = d_voronoi_point(vtx[d_org(q, f0)], vtx[d_dst(q, f0)], vtx[d_org(q, f1)], vtx[d_dst(q, f1)], &x, &y);
  }
  if (ok) vor[slot[0]] = vor[slot[1]] = vor[slot[2]] = make_float2(x, y);
  // else: "no point" == vertex 0 == (0, 0), which the memset of `vor` provides
}

// facet sizes: one thread per vertex; facets with fewer than 2 vertices contribute no edge (vd:97-114)
__global__ void facet_count_kernel(const SdEdges q, const SdVertex *__restrict__ vtx, int n_vtx,
                                   uint32_t *__restrict__ count, int *__restrict__ err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_vtx) return;
  uint32_t n = 0;
  if (k >= 4 && k < n_vtx && vtx[k].type == 0) {
    const int fe = vtx[k].first_edge;
    const int edge = (fe & ~3) + ((fe + 1) & 3);
    int t = edge;
    do {
      ++n;
      t = d_get_edge(q, t, kNextAroundLeft);
    } while (t != edge && n < (1u << 20));
    if (t != edge) atomicExch(err, 2);
    if (n < 2) n = 0;
  }
  count[k] = n;  // count[n_vtx] = 0 closes the scan
}

__global__ void facet_fill_kernel(const SdEdges q, const SdVertex *__restrict__ vtx, int n_vtx,
                                  const float2 *__restrict__ vor, const uint32_t *__restrict__ base,
                                  float2 *__restrict__ fxy, int *__restrict__ enext) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 4 || k >= n_vtx) return;
  const uint32_t b = base[k], n = base[k + 1] - b;
  if (n == 0) return;
  const int fe = vtx[k].first_edge;
  int t = (fe & ~3) + ((fe + 1) & 3);
  for (uint32_t j = 0; j < n; ++j) {
    // org of the dual edge t = the Voronoi point of slot (t & 3) of its quad: rot 1 -> slot 1, rot 3 -> slot 3
    fxy[b + j] = vor[2 * (t >> 2) + ((t & 3) == 1 ? 1 : 0)];
    enext[b + j] = (int)(j + 1 == n ? b : b + j + 1);
    t = d_get_edge(q, t, kNextAroundLeft);
  }
}
}  // namespace

// Upload the finished subdivision and size the facets; *n_slots = total facet vertices (one Voronoi edge each).
aos_status facets_prepare(Ctx *c, const Subdiv &sd, int *n_slots) {
  cudaStream_t st = c->stream;
  const int nq = (int)sd.n_quads(), nv = (int)sd.n_vertices();
  static_assert(sizeof(SdVertex) == sizeof(Subdiv::Vertex), "layout");
  const size_t ebytes = sizeof(int) * 4 * (size_t)nq;
  AOS_CUDA_OK(c, c->sd_quads.reserve(2 * ebytes));  // next | pt
  AOS_CUDA_OK(c, c->sd_verts.reserve(sizeof(SdVertex) * (size_t)nv));
  AOS_CUDA_OK(c, c->sd_vor.reserve(sizeof(float2) * 2 * (size_t)nq));
  AOS_CUDA_OK(c, c->sd_base.reserve(sizeof(uint32_t) * ((size_t)nv + 1)));
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  int *d_err = c->misc.as<int>() + 96;
  uint32_t *d_tot = reinterpret_cast<uint32_t *>(c->misc.as<int>() + 97);
  int *d_next = c->sd_quads.as<int>(), *d_pt = d_next + 4 * (size_t)nq;
  {  // registered in place by the caller (sd_pinned[i] says whether that worked)
    aos_status hs = h2d_small(c, d_next, sd.edge_next(), ebytes, c->sd_pinned[0] != nullptr);
    if (hs == AOS_OK) hs = h2d_small(c, d_pt, sd.edge_pt(), ebytes, c->sd_pinned[1] != nullptr);
    if (hs == AOS_OK) hs = h2d_small(c, c->sd_verts.p, sd.vertices(), sizeof(SdVertex) * (size_t)nv, c->sd_pinned[2] != nullptr);
    if (hs != AOS_OK) return hs;
  }
  AOS_CUDA_OK(c, cudaMemsetAsync(c->sd_vor.p, 0, sizeof(float2) * 2 * (size_t)nq, st));
  AOS_CUDA_OK(c, cudaMemsetAsync(d_err, 0, 8, st));
  const SdEdges q{d_next, d_pt};
  const SdVertex *v = c->sd_verts.as<SdVertex>();
  vor_points_kernel<<<(2 * nq + 255) / 256, 256, 0, st>>>(q, v, nq, c->sd_vor.as<float2>(), d_err);
  ++c->launches;
  facet_count_kernel<<<(nv + 1 + 255) / 256, 256, 0, st>>>(q, v, nv, c->sd_base.as<uint32_t>(), d_err);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  aos_status s = exclusive_scan_u32(c, c->sd_base.as<uint32_t>(), (size_t)nv + 1, c->cc_blocksum, d_tot);
  if (s != AOS_OK) return s;
  AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_err, 8, cudaMemcpyDeviceToHost, st));
  AOS_CUDA_OK(c, cudaStreamSynchronize(st));
  if (c->h_flag[0] != 0) {  // non-triangular face / facet walk that does not close: the caller walks it on the host
    *n_slots = -1;
    return AOS_OK;
  }
  *n_slots = c->h_flag[1];
  c->sd_nv = nv;
  c->sd_nq = nq;
  return AOS_OK;
}

// Write the facet-vertex slots and their cycle links (what run_graph otherwise receives from the host).
aos_status facets_fill(Ctx *c, float2 *d_fxy, int *d_enext) {
  const SdEdges q{c->sd_quads.as<int>(), c->sd_quads.as<int>() + 4 * (size_t)c->sd_nq};
  facet_fill_kernel<<<(c->sd_nv + 255) / 256, 256, 0, c->stream>>>(q, c->sd_verts.as<SdVertex>(), c->sd_nv,
                                                                   c->sd_vor.as<float2>(), c->sd_base.as<uint32_t>(), d_fxy,
                                                                   d_enext);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

}  // namespace aos
