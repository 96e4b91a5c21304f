// examples/map_to_graph.cpp -- the C-ABI from C++, as the two ROS 2 nodes would call it, without
// ROS: reads a raw PointXYZ cloud (16-byte x,y,z,pad records), runs the whole path and prints what a node would
// publish.   g++ -std=c++17 -Iinclude examples/map_to_graph.cpp -Lactive-orchard-slam_b200/lib -laos_gpu
//   usage: map_to_graph cloud.bin x0 y0 x1 y1 [resolution] [inflation]     (polygon = that rectangle)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "aos_gpu.h"

int main(int argc, char **argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s cloud.bin x0 y0 x1 y1 [resolution] [inflation_radius]\n", argv[0]);
    return 2;
  }
  std::FILE *f = std::fopen(argv[1], "rb");
  if (!f) {
    std::perror(argv[1]);
    return 2;
  }
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<float> cloud(static_cast<size_t>(bytes) / 4);
  if (std::fread(cloud.data(), 1, static_cast<size_t>(bytes), f) != static_cast<size_t>(bytes)) return 2;
  std::fclose(f);
  const size_t n = cloud.size() / 4;

  const double x0 = std::atof(argv[2]), y0 = std::atof(argv[3]), x1 = std::atof(argv[4]), y1 = std::atof(argv[5]);
  const double polygon[8] = {x0, y0, x1, y0, x1, y1, x0, y1};
  aos_seed_params p{};
  p.clipping_minz = -0.4f;  // config/aos_planner_params.yaml defaults
  p.clipping_maxz = 0.5f;
  p.grid_resolution = argc > 6 ? static_cast<float>(std::atof(argv[6])) : 0.05f;
  p.inflation_radius = argc > 7 ? static_cast<float>(std::atof(argv[7])) : 0.8f;
  p.cluster_min_length = 2.0;
  p.n_polygon = 4;
  p.polygon = polygon;

  aos_ctx *ctx = nullptr;
  if (aos_create(0, &ctx) != AOS_OK) {
    std::fprintf(stderr, "aos_create failed: no CUDA device (libaos_gpu has no CPU path)\n");
    return 1;
  }
  aos_status st = aos_map_to_graph(ctx, &p, cloud.data(), n, 16, 0, 4, 8, AOS_MEM_HOST);
  if (st != AOS_OK && st != AOS_ERR_STATE) {
    std::fprintf(stderr, "aos_map_to_graph: %s\n", aos_last_error(ctx));
    aos_destroy(ctx);
    return 1;
  }
  aos_seed_summary s;
  aos_seed_summary_get(ctx, &s);
  std::printf("%s\ngrid %d x %d @ %g m, origin (%g, %g); %lld points in the window; %d clusters, %d tree rows\n", aos_version(),
              s.info.width, s.info.height, s.info.resolution, s.info.origin_x, s.info.origin_y,
              static_cast<long long>(s.n_points_in), s.n_clusters, s.n_rows);
  if (st == AOS_OK) {
    aos_gvd_graph g;
    aos_get_graph(ctx, &g);
    int labelled = 0;
    for (int i = 0; i < g.n_nodes; ++i) labelled += g.node_labels[i] != 0;
    std::printf("GvdGraph: %d nodes (%d labelled TL/TR/BL/BR), %d edges, %d merged seeds\n", g.n_nodes, labelled, g.n_edges,
                g.n_merged_seeds);
  } else {
    std::printf("no rows on this map: no graph published (aos_gvd_node returns early, gvd:257)\n");
  }
  aos_destroy(ctx);
  return 0;
}
