// Third-party arithmetic the reference calls on this path, routed to hooks the Python side of the oracle installs
// (oracle/refbuild.py): the REAL OpenCV of this image through cv2 for cv::morphologyEx / cv::getStructuringElement /
// cv::Subdiv2D, and the oracle's restatements where the library is absent from the image (cv::ximgproc::thinning,
// pcl::RadiusOutlierRemoval).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
extern "C" {
// dst = cv2.morphologyEx(src, op, cv2.getStructuringElement(shape, (kw, kh)))   (uint8 rows x cols, contiguous)
typedef void (*ref_morph_hook)(const uint8_t *src, int rows, int cols, int op, int shape, int kw, int kh, uint8_t *dst);
// dst = cv::ximgproc::thinning(src, THINNING_ZHANGSUEN)
typedef void (*ref_thin_hook)(const uint8_t *src, int rows, int cols, int type, uint8_t *dst);
// cv2.Subdiv2D(rect).insert(points...) (per-point try/except) + getVoronoiFacetList([]):
// first call with facet_xy == nullptr returns the number of facets and total vertices, second call fills them.
typedef void (*ref_subdiv_hook)(const int *rect_xywh, const float *pts_xy, int n_pts, int *n_facets, int *n_vertices,
                                int *facet_sizes, float *facet_xy, float *centers_xy);
// keep[i] = 1 if pcl::RadiusOutlierRemoval(radius, min_neighbors) keeps point i
typedef void (*ref_ror_hook)(const float *xyz_pad, int n, double radius, int min_neighbors, uint8_t *keep);
struct ref_hooks_t { ref_morph_hook morph; ref_thin_hook thin; ref_subdiv_hook subdiv; ref_ror_hook ror; };
ref_hooks_t *ref_hooks();
}
