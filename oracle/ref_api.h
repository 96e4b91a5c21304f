/* ref_api.h -- C interface of oracle/_ref/libaos_ref.so: the reference's OWN node classes (compiled from
 * /root/reference against the stand-in headers of oracle/ref_shim/) driven in-process.  TEST INFRASTRUCTURE ONLY:
 * it pins the C oracle (oracle/aos_oracle_*.c) to the reference's compiled code; nothing under
 * active-orchard-slam_b200/ may include or link it. */
#ifndef AOS_REF_API_H
#define AOS_REF_API_H
#include <stddef.h>
#include <stdint.h>
#include "ref_shim/ref_shim_hooks.hpp"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float clipping_minz, clipping_maxz, grid_resolution, inflation_radius;
  double cluster_min_length;
  int n_poly;          /* 0: keep the node's built-in default polygon (seed_gen:193-215) */
  const double *poly;  /* x,y pairs; sent as geometry_msgs/PolygonStamped (float32 points) */
} ref_seed_params;

typedef struct {
  int w, h;
  double origin_x, origin_y;
  float res;
  int8_t *occ_border;      /* published /occupancy_grid */
  int8_t *skel;            /* last_skeletonized_grid_ (un-framed, what the seed ray casts walk) */
  int8_t *skel_framed;     /* published /skeletonized_occupancy_grid */
  int8_t *occ_raw_nodisc;  /* generateOccupancyGrid of the cropped cloud BEFORE the exclusion discs */
  int n_seeds;     double *seeds;        /* published /voronoi_seeds positions, x,y */
  int n_rows_info; double *rows_info;    /* published /exploration_tree_rows_info: start x,y,end x,y per row */
  int n_cluster_info; double *cluster_info; /* published cluster_info poses x,y,z */
  int n_clusters;          /* clusterOccupiedCells(last_skeletonized_grid_), all clusters in discovery order */
  int32_t *cl_size, *cl_first;
  float *cl_cx, *cl_cy, *cl_len;
  int32_t *cl_cell_off, *cl_cells;  /* BFS order, linear indices */
  int n_after_ror;
} ref_seed_result;

int ref_seed_run(const ref_seed_params *prm, const float *points, size_t n_points, size_t stride_floats,
                 int use_global_map_callback, ref_seed_result *out);
void ref_seed_result_free(ref_seed_result *r);
void ref_step_inflate(const int8_t *in, int w, int h, float res, float inflation_radius, int8_t *out);
void ref_step_mark_borders(const int8_t *in, int w, int h, int8_t *out);
void ref_step_skeletonize(const int8_t *in, int w, int h, int8_t *out);
int ref_step_point_in_polygon(double x, double y, const double *poly, int n);

typedef struct {
  int published;       /* number of /gvd/graph messages published by the call sequence */
  double resolution, origin_x, origin_y;
  int n_nodes, n_edges;
  double *nodes_xyz;   /* 3 per node */
  int32_t *node_labels, *node_cluster_indices, *node_label_counts;
  int n_label_entries;
  int32_t *node_label_clusters, *node_label_types;
  int32_t *edges;
  float *edge_lengths, *edge_clearances;
  int n_merged_seeds;  double *merged_seeds;      /* voronoi_seeds_ after voronoiSeedsCallback's merge */
  int n_voro_edges;    /* VoronoiDiagram::getEdges().size() */
  int n_boundary_points_precrop;
} ref_graph;

/* AosGvdNode: skeletonizedGridCallback, explorationTreeRowsInfoCallback, voronoiSeedsCallback (in that order), then the
 * last /gvd/graph message.  seeds: the /voronoi_seeds positions; rows_info: 4 doubles per row (start, end). */
int ref_gvd_run(const double *seeds_xy, int n_seeds, const int8_t *skel_framed, int w, int h, double origin_x,
                double origin_y, float res, const double *rows_info, int n_rows, ref_graph *out);
void ref_graph_free(ref_graph *g);

/* AosPathGenNode::trimPathNearOccupiedRegions (path_gen:1570-1630): returns the new pose count */
int ref_trim_path(const double *path_xy, int n, const int8_t *grid, int w, int h, double origin_x, double origin_y,
                  float res);

void ref_set_hooks(ref_morph_hook morph, ref_thin_hook thin, ref_subdiv_hook subdiv, ref_ror_hook ror);

#ifdef __cplusplus
}
#endif
#endif
