// Stand-in for the rosidl-generated header of /root/reference/msg/GvdGraph.msg (generated code, absent here):
// same field names, order and widths as the .msg file (GvdGraph.msg:4-58).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "ref_shim_msgs.hpp"
namespace aos { namespace msg {
struct GvdGraph {
  std_msgs::msg::Header header;
  double resolution = 0, origin_x = 0, origin_y = 0;
  int32_t num_nodes = 0, num_edges = 0;
  std::vector<geometry_msgs::msg::Point> nodes;
  std::vector<int32_t> node_labels, node_cluster_indices, node_label_clusters, node_label_types, node_label_counts, edges;
  std::vector<float> edge_lengths, edge_clearances;
  REF_SHIM_MSG_PTRS(GvdGraph)
};
}}
