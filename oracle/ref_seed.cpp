// ref_seed.cpp -- oracle/_ref, TEST INFRASTRUCTURE ONLY.
// Compiles the reference's aos_seed_gen_node UNMODIFIED, from /root/reference where it lies, against the stand-in
// headers of oracle/ref_shim/ (ROS 2 messages, rclcpp, PCL, OpenCV, Eigen are absent from this image), and exposes the
// node's own member functions to the tests through a small C interface.  No reference source is copied into this
// repository: the #include below reads it at build time (oracle/Makefile, target _ref).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <string>
#include <unordered_set>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <pcl/point_types.h>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#define main aos_seed_gen_node_main
#define private public
#define protected public
#include "src/aos_seed_gen_node.cpp"
#undef private
#undef protected
#undef main

#include "ref_api.h"

namespace {
template <class M>
std::shared_ptr<M> last_msg(const std::string &topic) {
  auto &b = ref_shim::Board::get();
  auto it = b.last.find(topic);
  if (it == b.last.end()) return nullptr;
  return std::static_pointer_cast<M>(it->second);
}
int8_t *dup_grid(const nav_msgs::msg::OccupancyGrid &g) {
  int8_t *p = (int8_t *)malloc(g.data.size() + 1);
  memcpy(p, g.data.data(), g.data.size());
  return p;
}
double *dup_poses(const geometry_msgs::msg::PoseArray &a, int *n) {
  *n = (int)a.poses.size();
  double *p = (double *)malloc(sizeof(double) * 2 * (a.poses.size() + 1));
  for (size_t i = 0; i < a.poses.size(); ++i) { p[2 * i] = a.poses[i].position.x; p[2 * i + 1] = a.poses[i].position.y; }
  return p;
}
}  // namespace

extern "C" int ref_seed_run(const ref_seed_params *prm, const float *points, size_t n_points, size_t stride_floats,
                            int use_global_map_callback, ref_seed_result *out) {
  memset(out, 0, sizeof(*out));
  auto &ov = ref_shim::ParamOverrides::get();
  ov.num.clear();
  ov.num["clipping_minz"] = prm->clipping_minz;
  ov.num["clipping_maxz"] = prm->clipping_maxz;
  ov.num["grid_resolution"] = prm->grid_resolution;
  ov.num["inflation_radius"] = prm->inflation_radius;
  ov.num["cluster_min_length"] = prm->cluster_min_length;
  ref_shim::Board::get().last.clear();
  AosSeedGenNode node;
  if (prm->n_poly > 0) {   // /aos_planner/exploration_area (PolygonStamped, float32 points)
    auto poly = std::make_shared<geometry_msgs::msg::PolygonStamped>();
    for (int i = 0; i < prm->n_poly; ++i) {
      geometry_msgs::msg::Point32 p;
      p.x = (float)prm->poly[2 * i];
      p.y = (float)prm->poly[2 * i + 1];
      poly->polygon.points.push_back(p);
    }
    node.explorationAreaCallback(poly);
  }
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>);
  cloud->points.resize(n_points);
  for (size_t i = 0; i < n_points; ++i) {
    cloud->points[i].x = points[i * stride_floats];
    cloud->points[i].y = points[i * stride_floats + 1];
    cloud->points[i].z = points[i * stride_floats + 2];
  }
  cloud->width = (uint32_t)n_points;
  cloud->height = 1;
  if (use_global_map_callback) {   // the whole callback incl. pcl::RadiusOutlierRemoval (seed_gen:229-248)
    auto msg = std::make_shared<sensor_msgs::msg::PointCloud2>();
    msg->width = (uint32_t)n_points;
    msg->height = 1;
    msg->point_step = 16;
    msg->row_step = (uint32_t)(16 * n_points);
    msg->is_dense = true;
    const char *names[3] = {"x", "y", "z"};
    for (int k = 0; k < 3; ++k) { sensor_msgs::msg::PointField f; f.name = names[k]; f.offset = 4 * k; f.datatype = 7; f.count = 1; msg->fields.push_back(f); }
    msg->data.resize(16 * n_points);
    if (n_points) memcpy(msg->data.data(), cloud->points.data(), 16 * n_points);
    node.globalMapCallback(msg);
    out->n_after_ror = node.last_cloud ? (int)node.last_cloud->points.size() : 0;
  } else {
    node.processPointCloud(cloud);   // the seam (seed_gen:452), as globalMapCallback / explorationAreaCallback call it
  }
  auto occ = last_msg<nav_msgs::msg::OccupancyGrid>("occupancy_grid");
  auto skf = last_msg<nav_msgs::msg::OccupancyGrid>("skeletonized_occupancy_grid");
  if (!occ || !skf) return -1;
  out->w = (int)occ->info.width;
  out->h = (int)occ->info.height;
  out->origin_x = occ->info.origin.position.x;
  out->origin_y = occ->info.origin.position.y;
  out->res = occ->info.resolution;
  out->occ_border = dup_grid(*occ);
  out->skel_framed = dup_grid(*skf);
  {
    std::lock_guard<std::mutex> lock(node.skeletonized_grid_mutex_);
    out->skel = dup_grid(node.last_skeletonized_grid_);
  }
  if (auto s = last_msg<geometry_msgs::msg::PoseArray>("voronoi_seeds")) out->seeds = dup_poses(*s, &out->n_seeds);
  if (auto r = last_msg<geometry_msgs::msg::PoseArray>("exploration_tree_rows_info")) {
    int n2 = 0;
    out->rows_info = dup_poses(*r, &n2);
    out->n_rows_info = n2 / 2;
  }
  if (auto c = last_msg<geometry_msgs::msg::PoseArray>("cluster_info")) {
    // publishClusterInfo: one pose per cluster kept by the length filter (position = centre, z = size or length)
    out->n_cluster_info = (int)c->poses.size();
    out->cluster_info = (double *)malloc(sizeof(double) * 3 * (c->poses.size() + 1));
    for (size_t i = 0; i < c->poses.size(); ++i) {
      out->cluster_info[3 * i] = c->poses[i].position.x;
      out->cluster_info[3 * i + 1] = c->poses[i].position.y;
      out->cluster_info[3 * i + 2] = c->poses[i].position.z;
    }
  }
  // intermediate artefacts through the node's own member functions (the same calls processPointCloud makes)
  {
    float minx, maxx, miny, maxy;
    node.getActiveBounds(minx, maxx, miny, maxy);
    pcl::PointCloud<pcl::PointXYZ>::Ptr f(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::PassThrough<pcl::PointXYZ> pass;
    pcl::PointCloud<pcl::PointXYZ>::Ptr src = use_global_map_callback ? node.last_cloud : cloud;
    pass.setInputCloud(src); pass.setFilterFieldName("z"); pass.setFilterLimits(node.clipping_minz, node.clipping_maxz); pass.filter(*f);
    pass.setInputCloud(f); pass.setFilterFieldName("x"); pass.setFilterLimits(minx, maxx); pass.filter(*f);
    pass.setInputCloud(f); pass.setFilterFieldName("y"); pass.setFilterLimits(miny, maxy); pass.filter(*f);
    // NOTE: the exclusion discs (seed_gen:481-525) are applied inside processPointCloud only; the raw grid exported
    // here is therefore the PRE-disc grid and is compared only on maps whose points lie outside every disc.
    nav_msgs::msg::OccupancyGrid raw = node.generateOccupancyGrid(f);
    out->occ_raw_nodisc = dup_grid(raw);
  }
  {
    // clusterOccupiedCells on the un-framed skeleton (seed_gen:970-1083), as clusterAndVisualizeSkeletonizedGrid calls it
    std::vector<Cluster> cl = node.clusterOccupiedCells(node.last_skeletonized_grid_);
    out->n_clusters = (int)cl.size();
    out->cl_size = (int32_t *)malloc(sizeof(int32_t) * (cl.size() + 1));
    out->cl_first = (int32_t *)malloc(sizeof(int32_t) * (cl.size() + 1));
    out->cl_cx = (float *)malloc(sizeof(float) * (cl.size() + 1));
    out->cl_cy = (float *)malloc(sizeof(float) * (cl.size() + 1));
    out->cl_len = (float *)malloc(sizeof(float) * (cl.size() + 1));
    out->cl_cell_off = (int32_t *)malloc(sizeof(int32_t) * (cl.size() + 2));
    size_t total = 0;
    for (auto &c : cl) total += c.cells.size();
    out->cl_cells = (int32_t *)malloc(sizeof(int32_t) * (total + 1));
    size_t k = 0;
    for (size_t i = 0; i < cl.size(); ++i) {
      out->cl_size[i] = cl[i].size;
      out->cl_cx[i] = cl[i].center_x;
      out->cl_cy[i] = cl[i].center_y;
      out->cl_len[i] = cl[i].length;
      out->cl_cell_off[i] = (int32_t)k;
      out->cl_first[i] = cl[i].cells.empty() ? -1 : cl[i].cells[0].first + cl[i].cells[0].second * out->w;
      for (auto &c : cl[i].cells) out->cl_cells[k++] = c.first + c.second * out->w;
    }
    out->cl_cell_off[cl.size()] = (int32_t)k;
  }
  return 0;
}

extern "C" void ref_seed_result_free(ref_seed_result *r) {
  free(r->occ_border); free(r->skel); free(r->skel_framed); free(r->occ_raw_nodisc); free(r->seeds); free(r->rows_info);
  free(r->cluster_info); free(r->cl_size); free(r->cl_first); free(r->cl_cx); free(r->cl_cy); free(r->cl_len);
  free(r->cl_cell_off); free(r->cl_cells);
  memset(r, 0, sizeof(*r));
}

// individual member functions on caller-supplied grids (unit comparisons against the oracle's steps)
namespace {
nav_msgs::msg::OccupancyGrid wrap_grid(const int8_t *g, int w, int h, float res, double ox, double oy) {
  nav_msgs::msg::OccupancyGrid m;
  m.info.width = (uint32_t)w; m.info.height = (uint32_t)h; m.info.resolution = res;
  m.info.origin.position.x = ox; m.info.origin.position.y = oy; m.info.origin.orientation.w = 1.0;
  m.data.assign(g, g + (size_t)w * h);
  return m;
}
AosSeedGenNode *step_node(float res, float inflation) {
  auto &ov = ref_shim::ParamOverrides::get();
  ov.num.clear();
  ov.num["grid_resolution"] = res;
  ov.num["inflation_radius"] = inflation;
  return new AosSeedGenNode();
}
}  // namespace

extern "C" void ref_step_inflate(const int8_t *in, int w, int h, float res, float inflation_radius, int8_t *out) {
  std::unique_ptr<AosSeedGenNode> n(step_node(res, inflation_radius));
  auto r = n->applyInflation(wrap_grid(in, w, h, res, 0, 0));
  memcpy(out, r.data.data(), (size_t)w * h);
}
extern "C" void ref_step_mark_borders(const int8_t *in, int w, int h, int8_t *out) {
  std::unique_ptr<AosSeedGenNode> n(step_node(0.05f, 0.8f));
  auto r = n->markBoundariesAsOccupied(wrap_grid(in, w, h, 0.05f, 0, 0));
  memcpy(out, r.data.data(), (size_t)w * h);
}
extern "C" void ref_step_skeletonize(const int8_t *in, int w, int h, int8_t *out) {
  std::unique_ptr<AosSeedGenNode> n(step_node(0.05f, 0.8f));
  auto r = n->skeletonizeOccupancyGrid(wrap_grid(in, w, h, 0.05f, 0, 0));
  memcpy(out, r.data.data(), (size_t)w * h);
}
extern "C" int ref_step_point_in_polygon(double x, double y, const double *poly, int n) {
  std::unique_ptr<AosSeedGenNode> nd(step_node(0.05f, 0.8f));
  std::vector<std::pair<double, double>> pg;
  for (int i = 0; i < n; ++i) pg.push_back({poly[2 * i], poly[2 * i + 1]});
  return nd->isPointInPolygon(x, y, pg) ? 1 : 0;
}
