/*
 * aos_oracle.h -- CPU restatement of the Active-orchard-slam map->GvdGraph hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under active-orchard-slam_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker / the timed CPU arm.
 *
 * PARITY PINNING: the reference ships no tests, fixtures or golden vectors (SURVEY.md section 4,
 * section 8c) and cannot be compiled here (needs ROS 2, PCL, Eigen, OpenCV C++).  This oracle is a
 * line-by-line restatement of the reference's own loops (cited per function) and is pinned
 *   - for cv::morphologyEx(MORPH_OPEN, 3x3 MORPH_ELLIPSE): against cv2 4.13 run in-container
 *     (tests/test_oracle_cpu.py::test_open_matches_cv2),
 *   - for cv::Subdiv2D: by calling the real cv2.Subdiv2D (oracle/subdiv.py), and
 *   - for cv::ximgproc::thinning(ZHANGSUEN) (opencv_contrib, absent): PARITY UNPINNED -- restated
 *     from the published Zhang-Suen algorithm as implemented by opencv_contrib 4.5.4.
 * PCL PassThrough/RadiusOutlierRemoval are absent too: PassThrough is restated (inclusive float
 * limits, non-finite removed); ROR sits before the processPointCloud seam and is out of the path.
 *
 * All file:line citations are into /root/reference/.
 */
#ifndef AOS_ORACLE_H
#define AOS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters of aos_seed_gen_node that reach the path (src/aos_seed_gen_node.cpp:69-100,2605-2619).
 * Members are float in the reference (declared <float>, read as_double() into float members). */
typedef struct {
  float clipping_minz, clipping_maxz;
  float clipping_minx, clipping_maxx, clipping_miny, clipping_maxy; /* used only when n_poly==0 */
  float grid_resolution;
  float inflation_radius;
  double cluster_min_length;
  int n_poly;             /* hardcoded_polygon_points_ (seed_gen:193-215 / :250-277) */
  const double *poly;     /* x,y pairs */
  int n_excl;             /* exclusion discs (seed_gen:487-499) */
  const float *excl;      /* x,y,r triples */
} orc_seed_params;

typedef struct {
  /* grid geometry (generateOccupancyGrid, seed_gen:581-600) */
  int w, h;
  double origin_x, origin_y;
  float res;
  /* all grids are nav_msgs/OccupancyGrid data: int8, row-major x + y*w, values {0,100} */
  int8_t *occ_raw;        /* generateOccupancyGrid                     seed_gen:581-622 */
  int8_t *occ_inflated;   /* applyInflation                            seed_gen:933-967 */
  int8_t *occ_border;     /* markBoundariesAsOccupied -> /occupancy_grid  seed_gen:708-757 */
  int8_t *opened;         /* morphologyEx(OPEN) stage of skeletonize    seed_gen:678-680 */
  int8_t *skel;           /* skeletonizeOccupancyGrid (un-framed)       seed_gen:672-705 */
  int8_t *skel_framed;    /* markPolygonBoundaryAsOccupied -> /skeletonized_occupancy_grid seed_gen:772-825 */
  /* clusterOccupiedCells, seed_gen:970-1083 (ALL clusters, discovery = raster order of first cell) */
  int n_clusters;
  int32_t *labels;        /* w*h: canonical label = min linear index of the component, -1 elsewhere */
  int32_t *cl_first;      /* linear index of the first (raster-min) cell == canonical label */
  int32_t *cl_size;
  int64_t *cl_sumx, *cl_sumy;     /* exact integer sums (diagnostic) */
  float *cl_cx, *cl_cy;   /* float32 running-sum centre in BFS order   seed_gen:1053-1059 */
  int64_t *cl_maxd2;      /* max pairwise integer squared distance     seed_gen:1062-1073 */
  float *cl_len;          /* length in metres                          seed_gen:1068,1074 */
  int32_t *cl_cell_off;   /* n_clusters+1 offsets into cl_cells */
  int32_t *cl_cells;      /* linear cell indices in BFS order */
  /* convertClustersToTreeRows, seed_gen:1309-1406: rows in cluster order (all_tree_rows) */
  int n_rows;
  int32_t *row_cluster;   /* index into cluster arrays */
  double *rows;           /* 7 per row: cx, cy, sx, sy, ex, ey, length */
  /* publishExplorationTreeRowsInfoFromClusters, seed_gen:2546-2582: sorted by (cy, cx); 4 per row */
  double *rows_info;      /* sx, sy, ex, ey  (n_rows entries) */
  /* /voronoi_seeds in publish order (seed_gen:1670-1710): virtual, ray, endpoint seeds */
  int n_seeds, n_virtual, n_ray, n_endpoint;
  double *seeds;          /* x,y pairs */
} orc_seed_result;

int orc_seed_stage(const orc_seed_params *p, const float *points, size_t n_points,
                   size_t stride_floats, orc_seed_result *out);
void orc_seed_result_free(orc_seed_result *r);

/* Fast mode (aos_oracle_fast.h): hash-grid / convex-hull / multi-threaded variants of the quadratic and full-image
 * loops with IDENTICAL results (tests/test_oracle_fast_cpu.py), so that config 3 finishes on a CPU.  Process-wide.
 * skip_labels: do not materialise the per-cell label map (1.6 GB at config 3); result->labels is NULL then. */
void orc_set_fast(int on, int threads, int skip_labels);

/* individual steps, exposed for unit tests and for the bounded-sample CPU timing */
void orc_active_bounds(const orc_seed_params *p, float *minx, float *maxx, float *miny, float *maxy);
void orc_grid_dims(float minx, float maxx, float miny, float maxy, float res, int *w, int *h);
void orc_bin_points(const orc_seed_params *p, const float *points, size_t n, size_t stride_floats,
                    int w, int h, double ox, double oy, int8_t *grid);
void orc_inflate(const int8_t *in, int w, int h, int cells, int8_t *out);
void orc_mark_borders(const int8_t *in, int w, int h, int8_t *out);
void orc_open_cross(const int8_t *in, int w, int h, int8_t *out);
int  orc_thin_zhangsuen(int8_t *img, int w, int h); /* in place; returns number of full passes */
void orc_frame_polygon_bbox(const orc_seed_params *p, const int8_t *in, int w, int h,
                            double ox, double oy, int8_t *out);
int  orc_point_in_polygon(double px, double py, const double *poly, int n_poly);

/* ---- aos_gvd_node half ------------------------------------------------------------------ */

/* voronoiSeedsCallback greedy 0.5 m merge, gvd:84-128.  out must hold 2*n doubles. */
int orc_gvd_merge_seeds(const double *seeds, int n, double *out);
/* trimPathNearOccupiedRegions (src/aos_path_gen_node.cpp:1570-1630): new pose count */
int orc_trim_path(const double *path_xy, int n, const int8_t *grid, int width, int height, double origin_x,
                  double origin_y, float res, double safety_distance);

/* VoronoiDiagram::compute up to the Subdiv2D call, vd:16-89: bounding rect (as the int Rect that
 * OpenCV 4.5.4's Subdiv2D(Rect) receives from the Rect2f) and the clipped float32 points.
 * rect_i[4] = x,y,w,h ; rect_f[4] = the Rect2f ; pts = 2*n floats; keep[n] = 1 if finite. */
void orc_gvd_subdiv_inputs(const double *seeds, int n, double min_x, double max_x, double min_y,
                           double max_y, int *rect_i, float *rect_f, float *pts, uint8_t *keep,
                           int *valid);

typedef struct {
  int n_nodes, n_edges;
  double *nodes;                 /* x,y pairs (z = 0) */
  int32_t *node_labels;          /* bitmask 1/2/4/8 */
  int32_t *node_cluster_indices;
  int32_t *node_label_counts;
  int n_label_entries;
  int32_t *node_label_clusters, *node_label_types;
  int32_t *edges;                /* flat pairs */
  float *edge_lengths, *edge_clearances;
  /* diagnostics */
  int n_voro_edges, n_boundary_points_precrop;
  double *corner_points;         /* per row: TL,TR,BL,BR x,y = 8 doubles */
} orc_graph;

/* facets: the output of Subdiv2D::getVoronoiFacetList as flat float32 xy + offsets (n_facets+1).
 * Everything after it: vd:97-114 (edges), vd:149-207 (extractBoundaryPoints), gvd:794-895,
 * gvd:420-483, gvd:485-556/686-790 (+castRay gvd:558-684), gvd:897-1010 (message arrays). */
int orc_gvd_graph(const float *facet_xy, const int32_t *facet_off, int n_facets,
                  const int8_t *skel_framed, int w, int h, double origin_x, double origin_y,
                  float res, const double *rows_info, int n_rows, orc_graph *out);
void orc_graph_free(orc_graph *g);

#ifdef __cplusplus
}
#endif
#endif
