// host_subdiv.h -- quad-edge incremental Delaunay / Voronoi facets (see host_subdiv.cu).
#pragma once
#include <stdint.h>

#include <vector>

namespace aos {

extern float g_outer_factor;

class Subdiv {
 public:
  // integer rectangle, as cv::Subdiv2D(Rect) receives it
  void init(int rx, int ry, int rw, int rh);
  void reserve(size_t n_points) {
    q_.reserve(3 * n_points + 16);
    vtx_.reserve(3 * n_points + 16);  // real vertices + one Voronoi vertex per triangle
  }
  // vertex id, or -1 where cv::Subdiv2D::insert would throw (point outside the rectangle / walk failure)
  int insert(float x, float y);
  // getVoronoiFacetList(idx = {}): one polygon per inserted vertex in insertion order, flat x,y + offsets
  void voronoi_facets(std::vector<float> *xy, std::vector<int32_t> *off);

  struct alignas(32) QuadEdge {  // 32 bytes, never straddles a cache line
    int next[4];
    int pt[4];
  };
  struct alignas(16) Vertex {
    int first_edge;
    int type;  // -1 free, 0 real, 1 virtual (Voronoi vertex)
    float x, y;
  };
  // the finished structure as it lies in memory, for the device facet kernels (k_facets.cu)
  const QuadEdge *quads() const { return q_.data(); }
  size_t n_quads() const { return q_.size(); }
  size_t quad_capacity() const { return q_.capacity(); }
  const Vertex *vertices() const { return vtx_.data(); }
  size_t n_vertices() const { return vtx_.size(); }
  size_t vertex_capacity() const { return vtx_.capacity(); }

 private:
  int org(int e) const { return q_[e >> 2].pt[e & 3]; }
  int dst(int e) const { return q_[e >> 2].pt[(e + 2) & 3]; }
  int get_edge(int edge, int type) const;
  int new_edge();
  void delete_edge(int edge);
  int new_point(float x, float y, bool is_virtual);
  void splice(int a, int b);
  void set_edge_points(int edge, int org, int dst);
  int connect_edges(int a, int b);
  void flip_around(int curr_edge, int first_point, float px, float py);
  int locate(float px, float py, int *edge, int *vertex);
  static bool voronoi_point(const Vertex &o0, const Vertex &d0, const Vertex &o1, const Vertex &d1, float *x, float *y);
  void calc_voronoi();

  std::vector<QuadEdge> q_;
  std::vector<Vertex> vtx_;
  int free_q_ = 0, free_pt_ = 0, recent_ = 0;
  bool valid_geometry_ = false;
  float tlx_ = 0, tly_ = 0, brx_ = 0, bry_ = 0;
};

}  // namespace aos
