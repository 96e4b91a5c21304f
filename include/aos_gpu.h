/*
 * aos_gpu.h -- C-ABI of libaos_gpu: the B200-native map -> GvdGraph hot path of Active-orchard-slam.
 *
 * The reference has no plugin seam on this path: everything is a private member of a ROS 2 node
 * class.  The two cut lines are the bodies of
 *     AosSeedGenNode::processPointCloud   (src/aos_seed_gen_node.cpp:452-579, callers :247, :284)
 *     AosGvdNode::processGraph            (src/aos_gvd_node.cpp:255-318, callers :127,149,170,176,182)
 * and these entry points are what a C++ node would call in their place (INTEGRATION.md shows the
 * shim).  Plain pointers and sizes only; no exceptions cross the boundary; every call returns an
 * aos_status and aos_last_error() gives the text.  A context is re-entrant per handle, not
 * thread-safe per handle (the reference runs one callback at a time per process:
 * rclcpp::spin single-threaded executor, seed_gen:2676, gvd:1648).
 *
 * All file:line citations are into the reference repository.
 */
#ifndef AOS_GPU_H
#define AOS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AOS_API __attribute__((visibility("default")))

typedef struct aos_ctx aos_ctx;

typedef enum {
  AOS_OK = 0,
  AOS_ERR_INVALID = -1,   /* bad argument */
  AOS_ERR_CUDA = -2,      /* CUDA runtime/driver error; see aos_last_error */
  AOS_ERR_CAPACITY = -3,  /* a caller-provided buffer is too small */
  AOS_ERR_STATE = -4,     /* stage called before its inputs exist (reference: silent early return, gvd:257) */
  AOS_ERR_NO_DEVICE = -5
} aos_status;

/* Where a caller-provided buffer lives. */
typedef enum { AOS_MEM_HOST = 0, AOS_MEM_DEVICE = 1 } aos_mem;

/* Parameters of aos_seed_gen_node that reach the path.  Names and defaults follow
 * config/aos_planner_params.yaml and the declare_parameter calls at seed_gen:69-83; the node keeps
 * them in float members (seed_gen:2605-2612), hence float here. */
typedef struct {
  float clipping_minz, clipping_maxz;                 /* -0.4, 0.5  (yaml :86-89) */
  float clipping_minx, clipping_maxx;                 /* -5, 72   used only when the polygon is empty */
  float clipping_miny, clipping_maxy;                 /* -10, 20 */
  float grid_resolution;                              /* 0.05 */
  float inflation_radius;                             /* 0.8 */
  double cluster_min_length;                          /* 2.0 (seed_gen:83,100) */
  int32_t n_polygon;                                  /* exploration polygon vertices (seed_gen:193-215, :250-277) */
  const double *polygon;                              /* host, x,y pairs */
  int32_t n_exclusion;                                /* exclusion discs (seed_gen:487-499 hard-codes 11) */
  const float *exclusion;                             /* host, x,y,r triples */
} aos_seed_params;

/* nav_msgs/OccupancyGrid.info subset (seed_gen:590-600). */
typedef struct {
  int32_t width, height;
  float resolution;
  double origin_x, origin_y;
} aos_grid_info;

/* Which grid of the seed stage (all W*H, row-major x + y*W, y up). */
typedef enum {
  AOS_GRID_RAW = 0,             /* generateOccupancyGrid             seed_gen:581-622 */
  AOS_GRID_INFLATED = 1,        /* applyInflation                    seed_gen:933-967 */
  AOS_GRID_OCCUPANCY = 2,       /* + markBoundariesAsOccupied -> /occupancy_grid   seed_gen:708-757 */
  AOS_GRID_OPENED = 3,          /* morphologyEx(OPEN) inside skeletonizeOccupancyGrid seed_gen:678-680 */
  AOS_GRID_SKELETON = 4,        /* skeletonizeOccupancyGrid, un-framed (ray casts use this) seed_gen:560-566 */
  AOS_GRID_SKELETON_FRAMED = 5  /* + markPolygonBoundaryAsOccupied -> /skeletonized_occupancy_grid seed_gen:772-825 */
} aos_grid_id;

typedef enum {
  AOS_FMT_INT8 = 0,  /* nav_msgs/OccupancyGrid data: 0 / 100, one byte per cell */
  AOS_FMT_BITS = 1   /* 32 cells per uint32 word, LSB = lowest x, row pitch = aos_bits_pitch_words() words */
} aos_grid_fmt;

/* Cluster (seed_gen:35-41) in discovery order == raster order of the first cell. */
typedef struct {
  int32_t label;        /* canonical label: min linear index (x + y*W) of the component */
  int32_t size;
  float center_x, center_y; /* grid units, float32 as in seed_gen:1053-1059 */
  float length;         /* metres, seed_gen:1062-1074 */
  int32_t reserved;
  int64_t sum_x, sum_y; /* exact integer sums */
  int64_t max_d2;       /* exact max pairwise squared cell distance */
} aos_cluster;

/* TreeRowFromCluster (seed_gen:44-49) plus the index of its cluster. */
typedef struct {
  double center_x, center_y;
  double start_x, start_y;
  double end_x, end_y;
  double length;
  int32_t cluster;
  int32_t reserved;
} aos_tree_row;

typedef struct {
  aos_grid_info info;
  int32_t n_clusters;   /* all components (before the cluster_min_length filter) */
  int32_t n_rows;       /* all_tree_rows == exploration_tree_rows (seed_gen:1405-1419) */
  int32_t thinning_launches, thinning_subiters; /* diagnostics */
  int64_t n_points_in;  /* points that survived the filters (diagnostic) */
} aos_seed_summary;

/* ---- lifetime ---------------------------------------------------------------------------- */
AOS_API aos_status aos_create(int device, aos_ctx **out);
AOS_API void aos_destroy(aos_ctx *ctx);
AOS_API const char *aos_last_error(const aos_ctx *ctx);
AOS_API const char *aos_version(void);
/* Run on a caller-owned CUDA stream (cudaStream_t as void*); default is a context-owned stream. */
AOS_API aos_status aos_set_stream(aos_ctx *ctx, void *cuda_stream);
AOS_API aos_status aos_synchronize(aos_ctx *ctx);
/* Words per row of an AOS_FMT_BITS grid of this width. */
AOS_API int32_t aos_bits_pitch_words(int32_t width);

/* ---- per-stage device timing (CUDA events on the context stream; SURVEY.md section 5 "tracing") ------
 * When enabled, every stage call records events between its kernels; aos_get_stage_times returns the
 * elapsed milliseconds of the last stage call, in execution order. */
typedef struct {
  char name[32];
  float ms;
} aos_stage_time;
AOS_API aos_status aos_set_profiling(aos_ctx *ctx, int enabled);
AOS_API aos_status aos_get_stage_times(aos_ctx *ctx, aos_stage_time *dst, int32_t capacity, int32_t *n_out);

/* ---- grid geometry: getActiveBounds + generateOccupancyGrid header (seed_gen:874-890, 587-600) - */
AOS_API aos_status aos_grid_geometry(const aos_seed_params *p, aos_grid_info *info);

/* ---- ahead of the seam: pcl::RadiusOutlierRemoval of globalMapCallback (seed_gen:229-248; radius 0.2 m, 2
 *      neighbours).  A point is kept iff at least min_neighbors OTHER points lie within `radius` in 3-D (squared
 *      distance accumulated in float32 as FLANN's L2_Simple does, compared with radius*radius in double; the dense-
 *      cloud branch of PCL 1.12, restated -- PCL is absent here, parity unpinned).  Non-finite points are dropped.
 *      The survivors come back compacted in input order as 16-byte x,y,z,1 records in context-owned DEVICE memory
 *      (valid until the next call of this function), ready to be passed to aos_seed_stage with AOS_MEM_DEVICE. -- */
AOS_API aos_status aos_radius_outlier_removal(aos_ctx *ctx, const void *points, size_t n_points, uint32_t point_step,
                                              uint32_t off_x, uint32_t off_y, uint32_t off_z, aos_mem points_mem,
                                              float radius, int32_t min_neighbors, const void **out_points,
                                              size_t *n_out);

/* ---- the seed-gen half: replaces the body of processPointCloud (seed_gen:452-579) up to and
 *      including clusterOccupiedCells / convertClustersToTreeRows' row extraction (:1309-1406).
 *      `points` is the post-RadiusOutlierRemoval cloud as PointCloud2 bytes: n_points records of
 *      point_step bytes with float32 x,y,z at byte offsets off_x/off_y/off_z (pcl::fromROSMsg,
 *      seed_gen:232-233).  point_step==16 with offsets 0/4/8 and a 16-byte aligned base is the
 *      vectorised fast path.  Asynchronous on the context stream; results stay on the device until
 *      fetched with the getters below (each getter synchronises). ------------------------------- */
AOS_API aos_status aos_seed_stage(aos_ctx *ctx, const aos_seed_params *p, const void *points,
                                  size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                                  uint32_t off_z, aos_mem points_mem);
AOS_API aos_status aos_seed_summary_get(aos_ctx *ctx, aos_seed_summary *out);
AOS_API aos_status aos_get_grid(aos_ctx *ctx, aos_grid_id which, aos_grid_fmt fmt, void *dst,
                                size_t dst_bytes, aos_mem dst_mem);
/* Device pointer of the context-owned bit grid (valid until the next aos_seed_stage). */
AOS_API aos_status aos_grid_device_bits(aos_ctx *ctx, aos_grid_id which, const uint32_t **bits,
                                        int32_t *pitch_words);
/* Per-cell canonical cluster label (min linear index of the component; -1 elsewhere), int32 W*H. */
AOS_API aos_status aos_get_labels(aos_ctx *ctx, int32_t *dst, size_t dst_count, aos_mem dst_mem);
AOS_API aos_status aos_get_clusters(aos_ctx *ctx, aos_cluster *dst, int32_t capacity, int32_t *n_out);
AOS_API aos_status aos_get_tree_rows(aos_ctx *ctx, aos_tree_row *dst, int32_t capacity, int32_t *n_out);

/* ---- seed selection: generateVirtualSeeds / raycastToOccupiedCell / generateRayPointsFromEndpoints /
 *      castRayFromEndpoint / endpoint seeds (seed_gen:1434-1511, 1730-1891, 1894-1982, 1987-2268) and the
 *      sorted /exploration_tree_rows_info (seed_gen:2546-2582).  The ray casts walk the un-framed skeleton bit
 *      grid on the device (one thread per ray) and the first-come 0.5 m filters are resolved there as well
 *      (k_seeds.cu); only the row sort is host code.  Seeds come out in /voronoi_seeds publish order: virtual,
 *      ray, endpoint (seed_gen:1670-1710); counts[3] = those sizes. ------------------------------------------ */
AOS_API aos_status aos_select_seeds(aos_ctx *ctx, int32_t *n_seeds, int32_t counts[3]);
AOS_API aos_status aos_get_seeds(aos_ctx *ctx, double *dst_xy, int32_t capacity, int32_t *n_out);
/* 4 doubles per row: start x, y, end x, y; rows sorted by centre (y, then x). */
AOS_API aos_status aos_get_rows_info(aos_ctx *ctx, double *dst, int32_t capacity_rows, int32_t *n_out);

/* ---- the gvd half: replaces the body of AosGvdNode::processGraph (gvd:255-318) together with the seed
 *      merge of voronoiSeedsCallback (gvd:84-128):
 *        merge seeds (0.5 m greedy, centroid)                          gvd:93-125
 *        VoronoiDiagram::compute  (Subdiv2D replay on the host, see csrc/host_subdiv.cu)   vd:16-114
 *        extractBoundaryPoints    (first-come 5 cm merge -> nodes)      vd:149-207
 *        buildGraphFromBoundaryPoints (edges, skeleton crossing test, 0.5 m proximity edges)  gvd:794-895, 320-359
 *        filterNodesAndEdgesOutsideGrid                                 gvd:420-483
 *        findClusterEndpointVoronoiBoundaryPoints (TL/TR/BL/BR)         gvd:485-556, 686-790, 558-684
 *        publishGraph's arrays                                          gvd:897-1010
 *      seeds_xy  : /voronoi_seeds positions (un-merged), host, x,y pairs
 *      rows_info : /exploration_tree_rows_info, host, 4 doubles per row (start x,y, end x,y) in message order
 *      skeleton  : /skeletonized_occupancy_grid data (int8, 100 = occupied), host, with its `info`; pass NULL
 *                  (and info NULL) to use the framed skeleton this context produced in aos_seed_stage.
 *      Returns AOS_ERR_STATE when seeds or skeleton are missing (the reference returns silently, gvd:257). - */
typedef struct {
  /* aos/msg/GvdGraph.msg field for field (msg/GvdGraph.msg:4-58); pointers are context-owned host memory,
   * valid until the next aos_gvd_stage / aos_destroy */
  float resolution;
  double origin_x, origin_y;
  int32_t n_nodes;
  const double *nodes_xyz;             /* geometry_msgs/Point[]: x, y, z (= 0) */
  const int32_t *node_labels;          /* bitmask TL=1 TR=2 BL=4 BR=8 */
  const int32_t *node_cluster_indices; /* first row that labelled the node, -1 if none */
  const int32_t *node_label_counts;
  int32_t n_label_entries;
  const int32_t *node_label_clusters, *node_label_types;
  int32_t n_edges;
  const int32_t *edges;                /* flat (from, to) pairs, from < to */
  const float *edge_lengths;
  const float *edge_clearances;        /* 0.0f, as the reference publishes (gvd:856,890) */
  /* diagnostics */
  int32_t n_merged_seeds, n_voronoi_edges, n_boundary_points;
  const double *corner_points;         /* per row TL,TR,BL,BR x,y (8 doubles) */
  int32_t n_rows;
} aos_gvd_graph;

AOS_API aos_status aos_gvd_stage(aos_ctx *ctx, const double *seeds_xy, int32_t n_seeds, const double *rows_info,
                                 int32_t n_rows, const int8_t *skeleton, const aos_grid_info *info);
/* Same, with the skeleton as an AOS_FMT_BITS grid (the 1 byte/cell OccupancyGrid is 400 MB at
 * 20000^2 cells; the bit-packed grid is 50 MB in host memory, or a device pointer shared by the seed-gen side --
 * e.g. aos_grid_device_bits(AOS_GRID_SKELETON_FRAMED) in the same process, a CUDA IPC handle across processes). */
AOS_API aos_status aos_gvd_stage_bits(aos_ctx *ctx, const double *seeds_xy, int32_t n_seeds, const double *rows_info,
                                      int32_t n_rows, const uint32_t *skeleton_bits, aos_mem skeleton_mem,
                                      const aos_grid_info *info);
AOS_API aos_status aos_get_graph(aos_ctx *ctx, aos_gvd_graph *out);

/* Opt-in (default off): fill GvdGraph.edge_clearances with the minimum, over the samples of
 * edgePassesThroughOccupiedPixels (gvd:320-359), of the exact Euclidean distance to the nearest occupied cell of
 * the framed skeleton, in metres.  The reference declares the field (msg/GvdGraph.msg:58) but publishes 0.0f
 * (gvd:856,890), so with this switched on the array is no longer bit-identical to the reference's. */
AOS_API aos_status aos_set_clearance(aos_ctx *ctx, int enabled);

/* The whole path in one call: aos_seed_stage, aos_select_seeds, aos_gvd_stage on the context's own grids. */
AOS_API aos_status aos_map_to_graph(aos_ctx *ctx, const aos_seed_params *p, const void *points, size_t n_points,
                                    uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                                    aos_mem points_mem);

/* cv::Subdiv2D::initDelaunay places its three outer vertices big_coord = factor * max(rect.width, rect.height)
 * away: factor 3 up to OpenCV 4.5.x (ROS 2 Humble's libopencv-dev = 4.5.4, the reference's platform, package.xml:48),
 * 6 in OpenCV 4.13 (the cv2 this library is validated against bit for bit, tests/test_subdiv_cpu.py).
 * The factor is NOT cosmetic: the far vertices change the flip history, hence which quad-edge pair each Voronoi
 * vertex is computed from; on orchard seed sets the graph topology stays the same but about 5 % of the node
 * coordinates move by a few float32 ulps (up to 2.3e-5 m at 60 m; tests/test_subdiv_cpu.py::test_outer_factor_...).
 * A node that wants the graph of ITS OpenCV bit for bit sets the factor it links against -- one line, version-agnostic:
 *     cv::Subdiv2D probe(cv::Rect(0, 0, 100, 100));
 *     aos_set_subdiv_outer_factor(probe.getVertex(1).x / 100.f);
 * (INTEGRATION.md section 3).  Process-wide; default 3 = the reference's platform. */
AOS_API aos_status aos_set_subdiv_outer_factor(float factor);
/* How VoronoiDiagram::compute (vd:16-114) is carried out on this context.
 *   AOS_VORONOI_REPLAY (default): cv::Subdiv2D's incremental insertion replayed step for step on the host
 *       (csrc/host_subdiv.cu), circumcentres and facet walks on the device: the graph is the reference's bit for bit.
 *   AOS_VORONOI_DEVICE (opt-in): the Voronoi cells of the same point set (seeds + Subdiv2D's outer triangle) built in
 *       parallel, one thread per seed clipping its cell by its neighbours' bisectors (csrc/k_vcells.cu).  Same diagram,
 *       no sequential step (0.27 s -> about 1 ms per 240 k seeds); but which edge pair a circumcentre is computed from
 *       and where a facet starts are products of Subdiv2D's flip history, so vertex low bits and the winners of the
 *       5 cm first-come merge (vd:149-207) can differ: the graph equals the reference's up to those choices (measured
 *       in tests/test_vcells_gpu.py and DESIGN.md), not bit for bit. */
#define AOS_VORONOI_REPLAY 0
#define AOS_VORONOI_DEVICE 1
AOS_API aos_status aos_set_voronoi_mode(aos_ctx *ctx, int32_t mode);
/* Current kernel-phase gate (aos_set_device_gate); aos_map_to_graph_batch restores it when it returns. */
AOS_API int32_t aos_get_device_gate(void);
/* Test switch (process-wide): run the replay's insert() in its literal form -- new_edge / splice / connect_edges for
 * the new vertex, swapEdges' four splices for every Lawson flip, every operand re-read from the structure -- instead of
 * the closed-form read-once / write-once slot updates.  Same result; off by default. */
AOS_API aos_status aos_set_subdiv_literal_splices(int32_t on);

/* Test switch (process-wide): which copy of the replay's flip loop runs.  -1 (default) = the AVX2 one where the CPU has
 * AVX2 (the four triangle areas of the in-circle test in the four lanes of one register, every lane with the scalar
 * operations in their order), 0 = the scalar one, 1 = AVX2 (AOS_ERR_INVALID on a CPU without it).  Same result, bit
 * for bit (tests/test_subdiv_cpu.py). */
AOS_API aos_status aos_set_subdiv_simd(int32_t mode);

/* How host threads of this process wait for `device` (process-wide, per device): 0 = the CUDA default (a waiting thread
 * spins on its core: lowest latency, right for a node that processes one map at a time), 1 = blocking waits
 * (cudaDeviceScheduleBlockingSync: a thread waiting for the GPU sleeps and leaves its core to the Subdiv2D replays of the
 * other maps in flight: right for batches with as many or more maps in flight than host cores).  Results do not change. */
AOS_API aos_status aos_set_host_wait(int device, int32_t blocking);

/* Kernel-phase gate, process-wide, off (0) by default: at most `max_concurrent` maps per device are admitted to the
 * seed stage's kernel phase (after the upload of their cloud) at a time; the others wait.  Identical maps in flight
 * on one GPU otherwise run in lockstep -- all in their kernel phase, time-slicing the GPU, then all in their host
 * phase (the Subdiv2D insertion replay) with the GPU idle; the gate staggers them after the first round.  Results
 * do not change.  aos_map_to_graph_batch sets it to 2 when it runs four or more maps at a time (measured best at 16
 * maps in flight; 1 serialises latency-bound phases that overlap well, larger values stagger less). */
AOS_API aos_status aos_set_device_gate(int32_t max_concurrent);

/* Independent maps in flight (BASELINE.json config 5: sweeps over maps / parameters): item i runs aos_map_to_graph on
 * its own context (contexts may sit on different devices) from a pool of at most max_threads host threads
 * (0 = one per item).  Two items may share a context only with max_threads == 1.  items[i].status receives each
 * map's status (AOS_ERR_STATE = a map without rows, as in aos_map_to_graph); the call returns the first real error. */
typedef struct {
  aos_ctx *ctx;
  const aos_seed_params *params;
  const void *points;
  size_t n_points;
  uint32_t point_step, off_x, off_y, off_z;
  aos_mem points_mem;
  aos_status status;
} aos_batch_item;
AOS_API aos_status aos_map_to_graph_batch(aos_batch_item *items, int32_t n_items, int32_t max_threads);

/* Stand-alone host steps of the gvd half (unit tests; no device needed). */
/* voronoiSeedsCallback's merge; out_xy must hold 2*n doubles; *n_out = merged count. */
AOS_API aos_status aos_merge_seeds(const double *seeds_xy, int32_t n, double *out_xy, int32_t *n_out);
/* VoronoiDiagram::compute up to getVoronoiFacetList: float32 polygons, flat x,y + offsets[n_facets+1].
 * Call with facet_xy == NULL to size the buffers (*n_facets, *n_points). */
AOS_API aos_status aos_voronoi_facets(const double *seeds_xy, int32_t n_seeds, double min_x, double max_x,
                                      double min_y, double max_y, float *facet_xy, int32_t xy_capacity_points,
                                      int32_t *facet_off, int32_t off_capacity, int32_t *n_facets,
                                      int32_t *n_points);

/* ---- consumer of the skeleton next to the path --------------------------------------
 * trimPathNearOccupiedRegions of aos_path_gen_node (src/aos_path_gen_node.cpp:1570-1630; call sites :1018, :1270,
 * :1552): a path is cut at the first pose i > 0 that has a cell == 100 of /skeletonized_occupancy_grid within
 * safety_distance (0.2 m in the reference; stencil of +-ceil(d/res) cells, offsets with sqrt(dx^2+dy^2)*res <= d).
 * *n_kept = number of poses that remain (path.poses.resize(i)).  skeleton_bits == NULL (and info == NULL): the
 * framed skeleton of this context's last seed stage; otherwise a bit-packed grid (aos_bits_pitch_words) in host or
 * device memory.  AOS_ERR_STATE when there is no skeleton (the reference returns without trimming, :1571). */
AOS_API aos_status aos_trim_path(aos_ctx *ctx, const double *path_xy, int32_t n_poses, double safety_distance,
                                 const uint32_t *skeleton_bits, aos_mem skeleton_mem, const aos_grid_info *info,
                                 int32_t *n_kept);

/* The merge as the gvd stage runs it (k_seeds.cu: leaders by monotone rounds over a 0.5 m hash grid, members summed
 * in index order); non-finite seeds are dropped, as processGraph does right after (gvd:266-270). */
AOS_API aos_status aos_merge_seeds_device(aos_ctx *ctx, const double *seeds_xy, int32_t n, double *out_xy, int32_t *n_out);
/* The same step as the gvd stage runs it: Delaunay insertions replayed on the host, cv::Subdiv2D::calcVoronoi and
 * getVoronoiFacetList (voronoi_diagram.cpp:94) on the device.  Returns the facet-vertex slots in facet order, facets
 * with fewer than 2 vertices dropped (voronoi_diagram.cpp:97-114 makes one edge per slot), and for every slot the
 * slot of the facet's next vertex.  Call with slot_xy == slot_next == NULL to size the buffers. */
AOS_API aos_status aos_voronoi_facets_device(aos_ctx *ctx, const double *seeds_xy, int32_t n_seeds, double min_x,
                                             double max_x, double min_y, double max_y, float *slot_xy /* 2 per slot */,
                                             int32_t *slot_next, int32_t capacity_slots, int32_t *n_slots);

/* ---- row-band sharding of the raster stages (BASELINE.json config 4: one huge grid over several GPUs) --------
 * Each GPU (one process, one context) owns the image rows [row0, row0 + rows) of the global grid and keeps
 * halo_lo / halo_hi extra rows below / above them (0 at the global border).  The halo must cover the stencil
 * reach of inflation + opening + one thinning launch: aos_band_halo_rows(params).  The caller buckets the cloud so
 * that every point whose cell row lies inside the local rows is given to this GPU (points outside are ignored, so
 * handing over a superset -- even the whole cloud -- is correct), runs aos_band_raster, then loops
 *     exchange the AOS_BAND_THIN_HALO rows next to each band edge with the neighbour (NVLink P2P / NCCL send-recv)
 *     aos_band_thin_launch  -> all-reduce(OR) of `deleted`
 * until no band deleted anything; the bands of AOS_GRID_SKELETON / AOS_GRID_OCCUPANCY are then gathered on one
 * GPU and aos_seed_stage_tail finishes the seed stage there (clusters, rows; then seeds and graph as usual).
 * Results are bit-identical to aos_seed_stage on one GPU (tests/test_bands_*.py). */
typedef struct {
  int32_t row0, rows;        /* owned rows of the global grid */
  int32_t halo_lo, halo_hi;  /* extra rows kept below row0 / above row0+rows */
} aos_band;
#define AOS_BAND_THIN_HALO 8  /* rows refreshed from the neighbour before every aos_band_thin_launch */
AOS_API int32_t aos_band_halo_rows(const aos_seed_params *p);
AOS_API aos_status aos_band_raster(aos_ctx *ctx, const aos_seed_params *p, const aos_band *band, const void *points,
                                   size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                                   uint32_t off_z, aos_mem points_mem);
AOS_API aos_status aos_band_thin_launch(aos_ctx *ctx, int32_t *deleted);
/* Fused halo exchange over peer memory (NVLink P2P), the variant without NCCL send/recv: every rank exports its two
 * thinning buffers as CUDA IPC handles, imports those of its lower (side 0) and upper (side 1) neighbour, and
 * aos_band_thin_launch_p2p then runs the same launch, except that the kernel stores the band's first / last 8 rows
 * straight into the neighbour's destination buffer (its halo rows) and leaves its own halo rows to the neighbours.
 * The buffers alternate launch by launch on every rank in lockstep; the caller only has to all-reduce the
 * `deleted` flags between launches (that also orders the peer stores before the next launch reads them).
 * The two buffers are dedicated planes of the context (allocated by aos_band_raster, re-allocated only when a
 * LARGER band is rasterised -- export and import again then; handles of an unchanged allocation stay valid from
 * map to map).  Export after this rank's aos_band_raster, launch after EVERY rank's aos_band_raster has returned
 * (e.g. exchange the handles with an all-gather every map).  One band uses one of the two launch calls, not both. */
#define AOS_IPC_HANDLE_BYTES 64
AOS_API aos_status aos_band_ipc_export(aos_ctx *ctx, int32_t buffer, unsigned char *handle /* [AOS_IPC_HANDLE_BYTES] */);
AOS_API aos_status aos_band_ipc_import(aos_ctx *ctx, int32_t side, int32_t buffer, const unsigned char *handle,
                                       int32_t peer_first_global_row);
AOS_API aos_status aos_band_thin_launch_p2p(aos_ctx *ctx, int32_t *deleted);
/* Close this context's mappings of its neighbours' planes and forget that its own were exported.  The exported planes
 * are never re-allocated while a peer may have them open: aos_band_raster refuses (AOS_ERR_STATE) a map that needs
 * larger planes until EVERY rank has called this and passed a barrier; the next map then exports / imports again. */
AOS_API aos_status aos_band_ipc_release(aos_ctx *ctx);
/* Device pointer of a LOCAL grid (AOS_GRID_RAW .. AOS_GRID_SKELETON; the skeleton is the current thinning image): local row r is global row row0 - halo_lo + r. */
AOS_API aos_status aos_band_grid_device(aos_ctx *ctx, aos_grid_id which, uint32_t **bits, int32_t *pitch_words,
                                        int32_t *local_rows);
/* Finish the seed stage from gathered full-size bit grids in device memory (pitch aos_bits_pitch_words(width)):
 * skeleton = thinned image (un-framed); occupancy = inflated + frame (may be NULL).  Afterwards the context
 * behaves as after aos_seed_stage, except that AOS_GRID_RAW / INFLATED / OPENED are not available. */
AOS_API aos_status aos_seed_stage_tail(aos_ctx *ctx, const aos_seed_params *p, const uint32_t *skeleton_bits,
                                       const uint32_t *occupancy_bits);

/* Kernels launched by this context since aos_create (bench.py's gpu_launches). */
AOS_API aos_status aos_get_launch_count(aos_ctx *ctx, int64_t *out);

/* ---- stand-alone steps (unit tests, parameter sweeps; same kernels as aos_seed_stage) -------- */
/* Each takes/returns AOS_FMT_BITS grids in DEVICE memory with pitch aos_bits_pitch_words(width). */
AOS_API aos_status aos_inflate_bits(aos_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t *out_border,
                                    int32_t width, int32_t height, int32_t radius_cells);
AOS_API aos_status aos_open_bits(aos_ctx *ctx, const uint32_t *in, uint32_t *out, int32_t width, int32_t height);
AOS_API aos_status aos_thin_bits(aos_ctx *ctx, uint32_t *inout, int32_t width, int32_t height,
                                 int32_t *launches, int32_t *subiters);
/* Exact Euclidean distance transform of a device bit grid (k_edt.cu): for every cell the nearest set cell as
 * x | y << 16 (0xFFFFFFFF when the grid is empty) and, if dist2 != NULL, the exact squared distance in cells.
 * nearest_xy / dist2: device, width*height entries, row-major.  Width and height below 65535. */
AOS_API aos_status aos_edt_bits(aos_ctx *ctx, const uint32_t *bits, int32_t width, int32_t height, uint32_t *nearest_xy,
                                int32_t *dist2);
/* applyInflation (seed_gen:933-967) as the threshold d^2 <= R^2 of that EDT: same result as aos_inflate_bits for
 * any radius (no 64-cell limit), cost independent of the radius. */
AOS_API aos_status aos_inflate_bits_edt(aos_ctx *ctx, const uint32_t *in, uint32_t *out, int32_t width, int32_t height,
                                        int32_t radius_cells);
AOS_API aos_status aos_pack_int8(aos_ctx *ctx, const int8_t *src, aos_mem src_mem, uint32_t *dst_bits,
                                 int32_t width, int32_t height);
AOS_API aos_status aos_unpack_int8(aos_ctx *ctx, const uint32_t *src_bits, int8_t *dst, aos_mem dst_mem,
                                   int32_t width, int32_t height);

#ifdef __cplusplus
}
#endif
#endif /* AOS_GPU_H */
