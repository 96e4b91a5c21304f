// ref_gvd.cpp -- oracle/_ref, TEST INFRASTRUCTURE ONLY.
// Compiles the reference's aos_gvd_node UNMODIFIED from /root/reference (see ref_seed.cpp for the method) and drives
// its subscription callbacks in-process; the last /gvd/graph message it publishes is handed back as flat arrays.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <Eigen/Dense>
#include "aos/msg/gvd_graph.hpp"

#define main aos_gvd_node_main
#define private public
#define protected public
#include "aos/voronoi_diagram.hpp"
#include "src/aos_gvd_node.cpp"
#undef private
#undef protected
#undef main

#include "ref_api.h"

template <class T, class V>
static T *dup_vec(const V &v) {
  T *p = (T *)malloc(sizeof(T) * (v.size() + 1));
  for (size_t i = 0; i < v.size(); ++i) p[i] = (T)v[i];
  return p;
}

extern "C" int ref_gvd_run(const double *seeds_xy, int n_seeds, const int8_t *skel_framed, int w, int h, double origin_x,
                           double origin_y, float res, const double *rows_info, int n_rows, ref_graph *out) {
  memset(out, 0, sizeof(*out));
  auto &ov = ref_shim::ParamOverrides::get();
  ov.num.clear();
  ov.num["max_graph_publish_rate"] = 1e12;   // the 10 Hz cap (gvd:306-314) would drop messages of back-to-back calls
  auto &board = ref_shim::Board::get();
  board.last.clear();
  board.count.clear();
  AosGvdNode node;
  auto grid = std::make_shared<nav_msgs::msg::OccupancyGrid>();
  grid->header.frame_id = "map";
  grid->info.width = (uint32_t)w;
  grid->info.height = (uint32_t)h;
  grid->info.resolution = res;
  grid->info.origin.position.x = origin_x;
  grid->info.origin.position.y = origin_y;
  grid->info.origin.orientation.w = 1.0;
  grid->data.assign(skel_framed, skel_framed + (size_t)w * h);
  node.skeletonizedGridCallback(grid);
  auto rows = std::make_shared<geometry_msgs::msg::PoseArray>();
  for (int r = 0; r < n_rows; ++r)
    for (int k = 0; k < 2; ++k) {
      geometry_msgs::msg::Pose p;
      p.position.x = rows_info[4 * r + 2 * k];
      p.position.y = rows_info[4 * r + 2 * k + 1];
      rows->poses.push_back(p);
    }
  node.explorationTreeRowsInfoCallback(rows);
  auto seeds = std::make_shared<geometry_msgs::msg::PoseArray>();
  for (int i = 0; i < n_seeds; ++i) {
    geometry_msgs::msg::Pose p;
    p.position.x = seeds_xy[2 * i];
    p.position.y = seeds_xy[2 * i + 1];
    seeds->poses.push_back(p);
  }
  node.voronoiSeedsCallback(seeds);
  out->n_merged_seeds = (int)node.voronoi_seeds_.size();
  out->merged_seeds = (double *)malloc(sizeof(double) * 2 * (node.voronoi_seeds_.size() + 1));
  for (size_t i = 0; i < node.voronoi_seeds_.size(); ++i) {
    out->merged_seeds[2 * i] = node.voronoi_seeds_[i].x();
    out->merged_seeds[2 * i + 1] = node.voronoi_seeds_[i].y();
  }
  out->n_voro_edges = (int)node.voronoi_diagram_.getEdges().size();
  out->published = board.count.count("/gvd/graph") ? board.count["/gvd/graph"] : 0;
  auto it = board.last.find("/gvd/graph");
  if (it == board.last.end()) return 0;
  auto g = std::static_pointer_cast<aos::msg::GvdGraph>(it->second);
  out->resolution = g->resolution;
  out->origin_x = g->origin_x;
  out->origin_y = g->origin_y;
  out->n_nodes = g->num_nodes;
  out->n_edges = g->num_edges;
  out->nodes_xyz = (double *)malloc(sizeof(double) * 3 * (g->nodes.size() + 1));
  for (size_t i = 0; i < g->nodes.size(); ++i) {
    out->nodes_xyz[3 * i] = g->nodes[i].x;
    out->nodes_xyz[3 * i + 1] = g->nodes[i].y;
    out->nodes_xyz[3 * i + 2] = g->nodes[i].z;
  }
  out->node_labels = dup_vec<int32_t>(g->node_labels);
  out->node_cluster_indices = dup_vec<int32_t>(g->node_cluster_indices);
  out->node_label_counts = dup_vec<int32_t>(g->node_label_counts);
  out->n_label_entries = (int)g->node_label_clusters.size();
  out->node_label_clusters = dup_vec<int32_t>(g->node_label_clusters);
  out->node_label_types = dup_vec<int32_t>(g->node_label_types);
  out->edges = dup_vec<int32_t>(g->edges);
  out->edge_lengths = dup_vec<float>(g->edge_lengths);
  out->edge_clearances = dup_vec<float>(g->edge_clearances);
  if ((int)g->nodes.size() != g->num_nodes || (int)g->edges.size() != 2 * g->num_edges) return -2;
  return 0;
}

extern "C" void ref_graph_free(ref_graph *g) {
  free(g->nodes_xyz); free(g->node_labels); free(g->node_cluster_indices); free(g->node_label_counts);
  free(g->node_label_clusters); free(g->node_label_types); free(g->edges); free(g->edge_lengths); free(g->edge_clearances);
  free(g->merged_seeds);
  memset(g, 0, sizeof(*g));
}
