// host_subdiv.h -- quad-edge incremental Delaunay / Voronoi facets (see host_subdiv.cu).
#pragma once
#include <stdint.h>

#include <vector>

namespace aos {

extern float g_outer_factor;
extern bool g_literal_splices;
extern int g_subdiv_simd;
bool subdiv_simd_available();

class Subdiv {
 public:
  // integer rectangle, as cv::Subdiv2D(Rect) receives it
  void init(int rx, int ry, int rw, int rh);
  void reserve(size_t n_points) {
    next_.reserve(4 * (3 * n_points + 16));
    pt_.reserve(4 * (3 * n_points + 16));
    vtx_.reserve(3 * n_points + 16);
  }
  // vertex id, or -1 where cv::Subdiv2D::insert would throw (point outside the rectangle / walk failure)
  int insert(float x, float y);
  // getVoronoiFacetList(idx = {}): one polygon per inserted vertex in insertion order, flat x,y + offsets
  void voronoi_facets(std::vector<float> *xy, std::vector<int32_t> *off);

  // Edges live in two flat arrays indexed by the edge id itself (id = 4 * quad + rot): next_[e] and pt_[e].  The
  // flip loop touches about 25 `next` slots per flip; with an array of {next[4], pt[4]} records every one of them
  // cost a shift/mask/scale to address, which was a quarter of the replay's time.
  struct alignas(32) Vertex {  // 32 bytes, never straddles a cache line
    double x, y;      // the float32 coordinates, widened once (the predicates are evaluated in double)
    double n2;        // x*x + y*y in double, the in-circle test's weight
    int first_edge;
    int type;         // -1 free, 0 real, 1 virtual (Voronoi vertex)
  };
  // the finished structure as it lies in memory, for the device facet kernels (k_facets.cu): next / pt per edge id, 4 per quad
  const int *edge_next() const { return next_.data(); }
  const int *edge_pt() const { return pt_.data(); }
  size_t n_quads() const { return next_.size() / 4; }
  size_t quad_capacity() const { return next_.capacity() / 4; }
  const Vertex *vertices() const { return vtx_.data(); }
  size_t n_vertices() const { return vtx_.size(); }
  size_t vertex_capacity() const { return vtx_.capacity(); }

 private:
  int org(int e) const { return pt_[e]; }
  int dst(int e) const { return pt_[e ^ 2]; }
  int get_edge(int edge, int type) const;
  int new_edge();
  void delete_edge(int edge);
  int new_point(float x, float y, bool is_virtual);
  void splice(int a, int b);
  void set_edge_points(int edge, int org, int dst);
  int connect_edges(int a, int b);
  int connect_inside_triangle(int e0, float px, float py);
  void flip_around(int curr_edge, int first_point, int curr_point, float px, float py);
  void flip_around_literal(int curr_edge, int first_point, int curr_point, float px, float py);
  void flip_around_avx2(int curr_edge, int first_point, int curr_point, float px, float py);
  int locate(float px, float py, int *edge, int *vertex);
  static bool voronoi_point(const Vertex &o0, const Vertex &d0, const Vertex &o1, const Vertex &d1, float *x, float *y);
  void calc_voronoi();

  std::vector<int> next_, pt_;  // SoA: next_[e], pt_[e] for edge id e = 4 * quad + rot
  std::vector<Vertex> vtx_;
  int free_q_ = 0, free_pt_ = 0, recent_ = 0;
  bool valid_geometry_ = false;
  float tlx_ = 0, tly_ = 0, brx_ = 0, bry_ = 0;
};

}  // namespace aos
