// ref_path.cpp -- oracle/_ref, TEST INFRASTRUCTURE ONLY: the reference's aos_path_gen_node compiled unmodified from
// /root/reference, only for AosPathGenNode::trimPathNearOccupiedRegions (path_gen:1570-1630).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <optional>
#include <queue>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include "aos/msg/gvd_graph.hpp"

#define main aos_path_gen_node_main
#define private public
#define protected public
#include "src/aos_path_gen_node.cpp"
#undef private
#undef protected
#undef main

#include "ref_api.h"

extern "C" int ref_trim_path(const double *path_xy, int n, const int8_t *grid, int w, int h, double origin_x, double origin_y,
                             float res) {
  ref_shim::ParamOverrides::get().num.clear();
  AosPathGenNode node;
  nav_msgs::msg::OccupancyGrid g;
  g.info.width = (uint32_t)w;
  g.info.height = (uint32_t)h;
  g.info.resolution = res;
  g.info.origin.position.x = origin_x;
  g.info.origin.position.y = origin_y;
  g.data.assign(grid, grid + (size_t)w * h);
  node.skeletonized_grid_ = g;
  nav_msgs::msg::Path path;
  path.poses.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    path.poses[i].pose.position.x = path_xy[2 * i];
    path.poses[i].pose.position.y = path_xy[2 * i + 1];
  }
  node.trimPathNearOccupiedRegions(path);
  return (int)path.poses.size();
}
