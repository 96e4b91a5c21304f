// ref_vd.cpp -- oracle/_ref, TEST INFRASTRUCTURE ONLY: the reference's aos::VoronoiDiagram compiled unmodified from
// /root/reference/src/utils/voronoi_diagram.cpp against oracle/ref_shim (cv::Subdiv2D -> the real cv2.Subdiv2D hook).
#include "src/utils/voronoi_diagram.cpp"
#include "ref_api.h"
static ref_hooks_t g_hooks = {nullptr, nullptr, nullptr, nullptr};
extern "C" ref_hooks_t *ref_hooks() { return &g_hooks; }
extern "C" void ref_set_hooks(ref_morph_hook morph, ref_thin_hook thin, ref_subdiv_hook subdiv, ref_ror_hook ror) {
  g_hooks.morph = morph; g_hooks.thin = thin; g_hooks.subdiv = subdiv; g_hooks.ror = ror;
}
