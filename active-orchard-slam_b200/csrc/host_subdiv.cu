// host_subdiv.cu -- incremental Delaunay triangulation / Voronoi facets with the arithmetic and the
// bookkeeping of cv::Subdiv2D, which aos::VoronoiDiagram::compute drives
// (src/utils/voronoi_diagram.cpp:51-94: Subdiv2D(rect), insert() per seed in order, getVoronoiFacetList).
//
// Why this runs on the host, sequentially: the reference's graph is a function of Subdiv2D's *history*, not
// only of the Delaunay triangulation.  (1) Every facet is emitted starting at vtx.firstEdge, which is the
// last quad-edge whose end points were (re)assigned at that vertex; extractBoundaryPoints
// (voronoi_diagram.cpp:149-207) then merges Voronoi vertices first-come within 5 cm, so the start of each
// cycle decides which of two nearby vertices survives.  (2) A Voronoi vertex is the intersection of the
// bisectors of the first two edges of its triangle in quad-edge index order, evaluated from float32
// differences and sums, so its float32 value depends on edge numbering (ulp of float32 at 1 km is 6e-5 m,
// above the 1e-5 m position tolerance).  Both are products of the insertion order, the walking point
// location and the free lists.  A parallel Delaunay construction produces the same triangles but neither
// the same vertex values nor the same merge winners, so this file replays the published Guibas-Stolfi
// quad-edge algorithm step for step (same predicates in double, same walk, same edge/vertex numbering).
// tests/test_subdiv_cpu.py pins it bit-for-bit against the real cv2.Subdiv2D.
//
// Quad-edge encoding: edge id = 4 * quad + rot (rot 0..3); sym = id ^ 2; rot +1 = dual edge.  next_[id] / pt_[id] are
// flat arrays indexed by the id itself; vertices carry their float32 coordinates widened to double plus |v|^2.
#include <float.h>
#include <math.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <utility>

#include "host_subdiv.h"

namespace aos {

namespace {
enum {
  NEXT_AROUND_ORG = 0x00,
  NEXT_AROUND_DST = 0x22,
  PREV_AROUND_ORG = 0x11,
  PREV_AROUND_DST = 0x33,
  NEXT_AROUND_LEFT = 0x13,
  NEXT_AROUND_RIGHT = 0x31,
  PREV_AROUND_LEFT = 0x20,
  PREV_AROUND_RIGHT = 0x02
};
enum { LOC_ERROR = -2, LOC_INSIDE = 0, LOC_VERTEX = 1, LOC_ON_EDGE = 2 };

inline double tri_aread(double ax, double ay, double bx, double by, double cx, double cy) {
  return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
inline double tri_area(float ax, float ay, float bx, float by, float cx, float cy) {
  return ((double)bx - ax) * ((double)cy - ay) - ((double)by - ay) * ((double)cx - ax);
}

// isPtInCircle3(pt, a, b, c) = sign of  |a|^2 A(b,c,pt) - |b|^2 A(a,c,pt) + |c|^2 A(a,b,pt) - |pt|^2 A(a,b,c)  with A =
// tri_area and a dead zone of FLT_EPSILON / 8; evaluated inline in flip_around (its only caller).
}  // namespace

float g_outer_factor = 3.f;  // OpenCV <= 4.5.x (the reference's platform); see aos_set_subdiv_outer_factor
bool g_literal_splices = false;  // see aos_set_subdiv_literal_splices (tests: take swapEdges' literal splice sequence)
int g_subdiv_simd = -1;  // see aos_set_subdiv_simd: -1 = AVX2 flip loop where the CPU has it, 0 = scalar, 1 = AVX2

bool subdiv_simd_available() {
#if defined(__x86_64__)
  static const bool have = __builtin_cpu_supports("avx2") != 0;
  return have;
#else
  return false;
#endif
}
static inline bool subdiv_simd_on() { return g_subdiv_simd != 0 && subdiv_simd_available(); }

void Subdiv::init(int rx_i, int ry_i, int rw, int rh) {
  vtx_.clear();
  next_.clear();
  pt_.clear();
  valid_geometry_ = false;
  // initDelaunay: the three outer vertices sit big_coord away (3 x max side up to OpenCV 4.5.x -- the default --, 6 x in 4.13)
  const float big = g_outer_factor * (float)(rw > rh ? rw : rh);
  const float rx = (float)rx_i, ry = (float)ry_i;
  tlx_ = rx;
  tly_ = ry;
  brx_ = rx + rw;
  bry_ = ry + rh;
  vtx_.push_back(Vertex{0, 0, 0, 0, -1});
  next_.resize(4, 0);
  pt_.resize(4, 0);
  free_q_ = 0;
  free_pt_ = 0;
  int pA = new_point(rx + big, ry, false);
  int pB = new_point(rx, ry + big, false);
  int pC = new_point(rx - big, ry - big, false);
  int eAB = new_edge(), eBC = new_edge(), eCA = new_edge();
  set_edge_points(eAB, pA, pB);
  set_edge_points(eBC, pB, pC);
  set_edge_points(eCA, pC, pA);
  splice(eAB, eCA ^ 2);
  splice(eBC, eAB ^ 2);
  splice(eCA, eBC ^ 2);
  recent_ = eAB;
}

int Subdiv::get_edge(int edge, int type) const {
  edge = next_[(edge & ~3) + ((edge + type) & 3)];
  return (edge & ~3) + ((edge + (type >> 4)) & 3);
}

int Subdiv::new_edge() {
  if (free_q_ <= 0) {
    next_.resize(next_.size() + 4, 0);
    pt_.resize(pt_.size() + 4, 0);
    free_q_ = (int)(next_.size() / 4) - 1;
  }
  int edge = free_q_ * 4;
  free_q_ = next_[edge + 1];
  next_[edge] = edge;
  next_[edge + 1] = edge + 3;
  next_[edge + 2] = edge + 2;
  next_[edge + 3] = edge + 1;
  pt_[edge] = pt_[edge + 1] = pt_[edge + 2] = pt_[edge + 3] = 0;
  return edge;
}

void Subdiv::delete_edge(int edge) {
  splice(edge, get_edge(edge, PREV_AROUND_ORG));
  int sedge = edge ^ 2;
  splice(sedge, get_edge(sedge, PREV_AROUND_ORG));
  edge >>= 2;
  next_[4 * edge] = 0;
  next_[4 * edge + 1] = free_q_;
  free_q_ = edge;
}

int Subdiv::new_point(float x, float y, bool is_virtual) {
  if (free_pt_ == 0) {
    vtx_.push_back(Vertex{0, 0, 0, 0, -1});
    free_pt_ = (int)vtx_.size() - 1;
  }
  int v = free_pt_;
  free_pt_ = vtx_[v].first_edge;
  vtx_[v] = Vertex{(double)x, (double)y, (double)x * x + (double)y * y, 0, is_virtual ? 1 : 0};
  return v;
}

void Subdiv::splice(int a, int b) {
  int &a_next = next_[a];
  int &b_next = next_[b];
  int a_rot = (a_next & ~3) + ((a_next + 1) & 3);
  int b_rot = (b_next & ~3) + ((b_next + 1) & 3);
  int &a_rot_next = next_[a_rot];
  int &b_rot_next = next_[b_rot];
  std::swap(a_next, b_next);
  std::swap(a_rot_next, b_rot_next);
}

void Subdiv::set_edge_points(int edge, int org, int dst) {
  pt_[edge] = org;
  pt_[edge ^ 2] = dst;
  vtx_[org].first_edge = edge;
  vtx_[dst].first_edge = edge ^ 2;
}

int Subdiv::connect_edges(int a, int b) {
  int edge = new_edge();
  splice(edge, get_edge(a, NEXT_AROUND_LEFT));
  splice(edge ^ 2, b);
  set_edge_points(edge, dst(a), org(b));
  return edge;
}

int Subdiv::locate(float px, float py, int *out_edge, int *out_vertex) {
  int vertex = 0;
  const int max_edges = (int)next_.size();
  if (px < tlx_ || py < tly_ || px >= brx_ || py >= bry_) return LOC_ERROR;  // cv::Exception(StsOutOfRange)
  typedef unsigned U32;
  U32 edge = (U32)recent_;
  int location = LOC_ERROR;
  const U32 *const nx = reinterpret_cast<const U32 *>(next_.data());
  const U32 *const pt = reinterpret_cast<const U32 *>(pt_.data());
  const Vertex *const vd = vtx_.data();
  auto right_of = [nx, pt, vd](double x, double y, U32 e) {  // isRightOf
    const Vertex &o = vd[pt[e]], &d = vd[pt[e ^ 2]];
    const double cw = tri_aread(x, y, d.x, d.y, o.x, o.y);
    return (cw > 0) - (cw < 0);
  };
  // The walk keeps the end points of `edge` as offsets from the query point: Onext(edge) shares its origin and
  // Dprev(edge) its destination (quad-edge ring invariants), so a step loads one or two new vertices, and isRightOf
  // (sign of triangleArea(p, dst, org) = (dst - p) x (org - p), the same subtractions and products) needs no others.
  const double pxd = px, pyd = py;
  double eox, eoy, edx, edy;
  {
    const Vertex &o = vd[pt[edge]], &d = vd[pt[edge ^ 2]];
    eox = o.x - pxd;
    eoy = o.y - pyd;
    edx = d.x - pxd;
    edy = d.y - pyd;
  }
  // isRightOf(p, edge) enters the walk's decisions only through "is it zero"; its sign is fixed by the walk (p is never
  // to the right of the current edge)
  bool on_curr;  // right_of_curr == 0
  {
    const double cw = edx * eoy - edy * eox;
    on_curr = cw == 0;
    if (cw > 0) {
      edge ^= 2u;
      std::swap(eox, edx);
      std::swap(eoy, edy);
    }
  }
  // One step of cv::Subdiv2D::locate with the tests in the order that needs the fewest loads: with p not to the right
  // of Onext(edge) and off the current edge the walk moves on to Onext whatever Dprev says (more than half of the
  // steps), so Dprev(edge) = InvRot(next[InvRot edge]) and its origin are only fetched when p is to the right of Onext
  // (or lies on the current edge).  Same moves, same stop, same number of steps as the literal if-chain.
  for (int i = 0; i < max_edges; ++i) {
    const U32 onext = nx[edge];
    const Vertex &c1 = vd[pt[onext ^ 2]];
    const double e1x = c1.x - pxd, e1y = c1.y - pyd;
    const double cw_onext = e1x * eoy - e1y * eox;  // sign = isRightOf(p, onext)
    if (cw_onext > 0 || on_curr) {
      U32 dprev = nx[edge ^ (3u - ((edge & 1u) << 1))];  // InvRot: r -> r + 3 mod 4 = xor with 11 (r even) or 01 (r odd)
      dprev ^= 3u - ((dprev & 1u) << 1);
      const Vertex &c2 = vd[pt[dprev]];
      const double e2x = c2.x - pxd, e2y = c2.y - pyd;
      const double cw_dprev = edx * e2y - edy * e2x;  // sign = isRightOf(p, dprev)
      if (cw_onext > 0) {
        if (cw_dprev > 0 || (cw_dprev == 0 && on_curr)) {
          location = LOC_INSIDE;
          break;
        }
        on_curr = cw_dprev == 0;
        edge = dprev;
        eox = e2x;
        eoy = e2y;
        continue;
      }
      // p lies on the current edge and is not to the right of Onext
      if (cw_dprev > 0) {
        if (cw_onext == 0) {
          location = LOC_INSIDE;
          break;
        }
      } else if (right_of(c1.x, c1.y, edge) >= 0) {
        edge ^= 2u;
        std::swap(eox, edx);
        std::swap(eoy, edy);
        continue;
      }
    }
    on_curr = cw_onext == 0;
    edge = onext;
    edx = e1x;
    edy = e1y;
  }
  recent_ = (int)edge;
  int edge_out = (int)edge;
  if (location == LOC_INSIDE) {
    const Vertex &vo = vtx_[org(edge_out)], &vd2 = vtx_[dst(edge_out)];
    struct { float x, y; } o{(float)vo.x, (float)vo.y}, d{(float)vd2.x, (float)vd2.y};  // the float32 coordinates
    double t1 = fabs(px - o.x);  // float differences, as cv::Point2f arithmetic
    t1 += fabs(py - o.y);
    double t2 = fabs(px - d.x);
    t2 += fabs(py - d.y);
    double t3 = fabs(o.x - d.x);
    t3 += fabs(o.y - d.y);
    if (t1 < FLT_EPSILON) {
      location = LOC_VERTEX;
      vertex = org(edge_out);
      edge_out = 0;
    } else if (t2 < FLT_EPSILON) {
      location = LOC_VERTEX;
      vertex = dst(edge_out);
      edge_out = 0;
    } else if ((t1 < t3 || t2 < t3) && fabs(tri_area(px, py, o.x, o.y, d.x, d.y)) < FLT_EPSILON) {
      location = LOC_ON_EDGE;
      vertex = 0;
    }
  }
  if (location == LOC_ERROR) {
    edge_out = 0;
    vertex = 0;
  }
  *out_edge = edge_out;
  *out_vertex = vertex;
  return location;
}

int Subdiv::insert(float px, float py) {
  int curr_point = 0, curr_edge = 0;
  int location = locate(px, py, &curr_edge, &curr_point);
  if (location == LOC_ERROR) return -1;  // cv::Exception; the reference skips the seed (voronoi_diagram.cpp:83-88)
  if (location == LOC_VERTEX) return curr_point;
  if (location == LOC_INSIDE && free_q_ <= 0 && !g_literal_splices) {
    const int p = connect_inside_triangle(curr_edge, px, py);
    if (p > 0) return p;
  }
  if (location == LOC_ON_EDGE) {
    int deleted = curr_edge;
    recent_ = curr_edge = get_edge(curr_edge, PREV_AROUND_ORG);
    delete_edge(deleted);
  }
  valid_geometry_ = false;
  curr_point = new_point(px, py, false);
  int base_edge = new_edge();
  int first_point = org(curr_edge);
  set_edge_points(base_edge, first_point, curr_point);
  splice(base_edge, curr_edge);
  do {
    base_edge = connect_edges(curr_edge, base_edge ^ 2);
    curr_edge = get_edge(base_edge, PREV_AROUND_ORG);
  } while (dst(curr_edge) != first_point);
  curr_edge = get_edge(base_edge, PREV_AROUND_ORG);
  flip_around(curr_edge, first_point, curr_point, px, py);
  return curr_point;
}

// insert() between locate() and the flip loop for a point strictly inside a triangle, with the three new quads taken from
// the end of the arrays (nothing on the free list): the new vertex p and its edges B0 = (o -> p), B1 = (d -> p),
// B2 = (c -> p) in the triangle e0 = (o -> d), e1 = Lnext(e0) = (d -> c), e2 = Lnext(e1) = (c -> o).  new_edge,
// splice(B0, e0) and the two connect_edges of the literal sequence leave 18 `next` slots with values that follow from
// e0, e1, e2 alone, written here once:
//   rings of o, d, c:   next[ei] = Bi,  next[Bi] = what next[ei] was = Sym(e(i-1))   (the face is a triangle)
//   ring of p:          Sym B0 -> Sym B1 -> Sym B2 -> Sym B0
//   faces (dual rings): next[InvRot x] = InvRot(Lnext x) around (e0, B1, Sym B0), (e1, B2, Sym B1), (e2, B0, Sym B2)
// and first_edge of o, d, c, p as the three set_edge_points calls leave them.  Returns p, or 0 with nothing changed when
// the face is not a triangle (never the case for a subdivision this class built; the literal sequence takes over).
int Subdiv::connect_inside_triangle(int e0_i, float px, float py) {
  typedef unsigned U32;
  auto rot = [](U32 e) -> U32 { return (e & ~3u) | ((e + 1u) & 3u); };
  auto invrot = [](U32 e) -> U32 { return (e & ~3u) | ((e + 3u) & 3u); };
  const U32 e0 = (U32)e0_i, ir0 = invrot(e0);
  {
    const U32 *const nx = reinterpret_cast<const U32 *>(next_.data());
    const U32 e1 = rot(nx[ir0]), e2 = rot(nx[invrot(e1)]);
    if (rot(nx[invrot(e2)]) != e0) return 0;
  }
  valid_geometry_ = false;
  const U32 p = (U32)new_point(px, py, false);
  const U32 B0 = (U32)next_.size(), B1 = B0 + 4, B2 = B0 + 8;
  next_.resize(B0 + 12, 0);
  pt_.resize(B0 + 12, 0);  // pt of the dual edges stays 0, as new_edge leaves it
  U32 *const nx = reinterpret_cast<U32 *>(next_.data());
  U32 *const pt = reinterpret_cast<U32 *>(pt_.data());
  Vertex *const vd = vtx_.data();
  const U32 e1 = rot(nx[ir0]), ir1 = invrot(e1), e2 = rot(nx[ir1]), ir2 = invrot(e2);
  const U32 o = pt[e0], d = pt[e1], c = pt[e2];
  nx[e0] = B0, nx[e1] = B1, nx[e2] = B2;
  nx[B0] = e2 ^ 2u, nx[B1] = e0 ^ 2u, nx[B2] = e1 ^ 2u;
  nx[B0 + 2] = B1 + 2, nx[B1 + 2] = B2 + 2, nx[B2 + 2] = B0 + 2;
  nx[ir0] = B1 + 3, nx[B1 + 3] = B0 + 1, nx[B0 + 1] = ir0;
  nx[ir1] = B2 + 3, nx[B2 + 3] = B1 + 1, nx[B1 + 1] = ir1;
  nx[ir2] = B0 + 3, nx[B0 + 3] = B2 + 1, nx[B2 + 1] = ir2;
  pt[B0] = o, pt[B1] = d, pt[B2] = c;
  pt[B0 + 2] = pt[B1 + 2] = pt[B2 + 2] = p;
  vd[o].first_edge = (int)B0;
  vd[d].first_edge = (int)B1;
  vd[c].first_edge = (int)B2;
  vd[p].first_edge = (int)(B2 + 2);
  flip_around((int)e2, (int)o, (int)p, px, py);
  return (int)p;
}

// The Lawson flips around the new point (the second loop of cv::Subdiv2D::insert), literally: the reference form of the
// loop below, taken when aos_set_subdiv_literal_splices is on (tests compare the two).
void Subdiv::flip_around_literal(int curr_edge, int first_point, int curr_point, float px, float py) {
  const int max_edges = (int)next_.size();
  const double eps = FLT_EPSILON * 0.125;
  const Vertex &p = vtx_[curr_point];
  (void)px, (void)py;
  for (int i = 0; i < max_edges; ++i) {
    const int temp_edge = get_edge(curr_edge, PREV_AROUND_ORG);
    const int temp_dst = dst(temp_edge), curr_org = org(curr_edge), curr_dst = dst(curr_edge);
    const Vertex &t = vtx_[temp_dst], &o = vtx_[curr_org], &d = vtx_[curr_dst];
    // isPtInCircle3(pt = org, a = t, b = dst, c = p)
    double val = t.n2 * tri_aread(d.x, d.y, p.x, p.y, o.x, o.y);
    val -= d.n2 * tri_aread(t.x, t.y, p.x, p.y, o.x, o.y);
    val += p.n2 * tri_aread(t.x, t.y, d.x, d.y, o.x, o.y);
    val -= o.n2 * tri_aread(t.x, t.y, d.x, d.y, p.x, p.y);
    if (tri_aread(t.x, t.y, d.x, d.y, o.x, o.y) > 0 && val < -eps) {  // isRightOf(t, curr_edge) > 0 && inside
      // swapEdges(curr_edge)
      const int sedge = curr_edge ^ 2;
      const int a = get_edge(curr_edge, PREV_AROUND_ORG), b = get_edge(sedge, PREV_AROUND_ORG);
      splice(curr_edge, a);
      splice(sedge, b);
      set_edge_points(curr_edge, dst(a), dst(b));
      splice(curr_edge, get_edge(a, NEXT_AROUND_LEFT));
      splice(sedge, get_edge(b, NEXT_AROUND_LEFT));
      curr_edge = get_edge(curr_edge, PREV_AROUND_ORG);
    } else if (curr_org == first_point) {
      break;
    } else {
      curr_edge = get_edge(next_[curr_edge], PREV_AROUND_LEFT);
    }
  }
}

// The same loop as the replay runs it: same tests, same order, same final stores.  Hot loop of the whole gvd half
// (seeds sorted along rows: about 46 flips and 95 iterations per seed, 70 % of the replay's time), and bound by the
// number of instructions it retires, so everything an iteration can know without a load is carried in registers:
//   * e = (o -> d) always has the new point p to its left (loop invariant of cv::Subdiv2D::insert), so
//     Oprev(Sym e) = (d -> p): the flipped edge's new destination is p itself, no load;
//   * after a flip the loop continues with Oprev(e) = (t -> d): the origin becomes the apex t, the destination
//     and its coordinates stay; after a step without flip with Lprev(Onext(e)), which ends at the old origin;
//   * rot e is carried (Rot Rot = Sym spares the mask arithmetic after a flip).
// swapEdges' four splices touch twelve `next` slots whose final values follow from the initial ones (DESIGN.md
// section 4, "Host hot loop"); every face of the subdivision is a triangle whenever this loop runs (insert() has
// connected p to every corner of the face it fell into, and a flip maps two triangles to two triangles), which gives
// the ring identities Onext(e) = Sym Lnext(b), Onext(Sym e) = Sym Lnext(a) used below: five loads and sixteen stores
// per flip.  With Q = next[rot e], Q2 = next[rot Sym e], U = next[Q], U2 = next[Q2]:
//   a = Oprev(e) = rot Q, b = Oprev(Sym e) = rot Q2, InvRot a = Q, InvRot b = Q2, la = Lnext(a) = rot U,
//   lb = Lnext(b) = rot U2, rot Onext(Sym e) = U, rot Onext(e) = U2.
#define AOS_GEOM_DECL double ox, oy, on2, dx, dy, dn2
#define AOS_GEOM_INIT(o, d) (ox = (o).x, oy = (o).y, on2 = (o).n2, dx = (d).x, dy = (d).y, dn2 = (d).n2)
// isRightOf is the sign of triangleArea(t, dst, org); the same determinant is the third term of
// isPtInCircle3(pt = org, a = t, b = dst, c = p), evaluated once
#define AOS_GEOM_TEST(t)                                               \
  const double tx = (t).x, ty = (t).y, tn2 = (t).n2;                   \
  const double area_tdo = tri_aread(tx, ty, dx, dy, ox, oy);           \
  double val = tn2 * tri_aread(dx, dy, pxd, pyd, ox, oy);              \
  val -= dn2 * tri_aread(tx, ty, pxd, pyd, ox, oy);                    \
  val += pxx * area_tdo;                                               \
  val -= on2 * tri_aread(tx, ty, dx, dy, pxd, pyd);                    \
  const bool do_flip = (area_tdo > 0) & (val < -eps)
#define AOS_GEOM_FLIPPED (ox = tx, oy = ty, on2 = tn2)
#define AOS_GEOM_ADVANCED(o) (dx = ox, dy = oy, dn2 = on2, ox = (o).x, oy = (o).y, on2 = (o).n2)
void Subdiv::flip_around(int curr_edge_i, int first_point, int curr_point, float px, float py) {
  if (g_literal_splices) return flip_around_literal(curr_edge_i, first_point, curr_point, px, py);
#if defined(__x86_64__)
  if (subdiv_simd_on()) return flip_around_avx2(curr_edge_i, first_point, curr_point, px, py);
#endif
#include "host_subdiv_flip.inc"
}
#undef AOS_GEOM_DECL
#undef AOS_GEOM_INIT
#undef AOS_GEOM_TEST
#undef AOS_GEOM_FLIPPED
#undef AOS_GEOM_ADVANCED

#if defined(__x86_64__)
// The same loop with the four triangle areas of the in-circle test in the four lanes of one AVX2 register.  Lane k
// evaluates (M - B) x (N - B) with  M = [p, p, d, d],  B = [d, t, t, t],  N = [o, o, o, p]:
//   A0 = area(d, p, o), A1 = area(t, p, o), A2 = area(t, d, o) (= isRightOf's determinant), A3 = area(t, d, p),
// every lane with the operations of tri_aread in their order, then W = [|t|^2, |d|^2, |p|^2, |o|^2] times A and the sum
// ((W0 A0 - W1 A1) + W2 A2) - W3 A3 in the scalar order: bit for bit the value of the scalar loop.  M, N and W are
// kept across iterations and re-blended where an end point changes (a flip changes o, a step changes both).
#define AOS_GEOM_DECL __m256d PXv, PYv, Mx, My, Nx, Ny, Dx, Dy, Wb
#define AOS_GEOM_INIT(o, d)                                                                                   \
  (PXv = _mm256_set1_pd(pxd), PYv = _mm256_set1_pd(pyd), Dx = _mm256_broadcast_sd(&(d).x),                    \
   Dy = _mm256_broadcast_sd(&(d).y), Mx = _mm256_blend_pd(PXv, Dx, 12), My = _mm256_blend_pd(PYv, Dy, 12),    \
   Nx = _mm256_blend_pd(_mm256_broadcast_sd(&(o).x), PXv, 8), Ny = _mm256_blend_pd(_mm256_broadcast_sd(&(o).y), PYv, 8), \
   Wb = _mm256_blend_pd(_mm256_blend_pd(_mm256_set1_pd(pxx), _mm256_broadcast_sd(&(d).n2), 2), _mm256_broadcast_sd(&(o).n2), 8))
#define AOS_GEOM_TEST(t)                                                                                      \
  const __m256d Tx = _mm256_broadcast_sd(&(t).x), Ty = _mm256_broadcast_sd(&(t).y), Tn = _mm256_broadcast_sd(&(t).n2); \
  const __m256d Bx = _mm256_blend_pd(Tx, Dx, 1), By = _mm256_blend_pd(Ty, Dy, 1);                             \
  const __m256d Ux = _mm256_sub_pd(Mx, Bx), Vy = _mm256_sub_pd(Ny, By);                                       \
  const __m256d Uy = _mm256_sub_pd(My, By), Vx = _mm256_sub_pd(Nx, Bx);                                       \
  const __m256d Ar = _mm256_sub_pd(_mm256_mul_pd(Ux, Vy), _mm256_mul_pd(Uy, Vx));                             \
  const __m256d Wa = _mm256_mul_pd(_mm256_blend_pd(Wb, Tn, 1), Ar);                                           \
  const __m128d w_lo = _mm256_castpd256_pd128(Wa), w_hi = _mm256_extractf128_pd(Wa, 1);                       \
  __m128d val = _mm_sub_sd(w_lo, _mm_unpackhi_pd(w_lo, w_lo));                                                \
  val = _mm_add_sd(val, w_hi);                                                                                \
  val = _mm_sub_sd(val, _mm_unpackhi_pd(w_hi, w_hi));                                                         \
  const bool do_flip = (_mm_cvtsd_f64(_mm256_extractf128_pd(Ar, 1)) > 0) & (_mm_cvtsd_f64(val) < -eps)
#define AOS_GEOM_FLIPPED (Nx = _mm256_blend_pd(Tx, PXv, 8), Ny = _mm256_blend_pd(Ty, PYv, 8), Wb = _mm256_blend_pd(Wb, Tn, 8))
// d := o (lane 0 of N), o := the record; W's |d|^2 lane takes the old |o|^2 (lane 3 -> lane 1)
#define AOS_GEOM_ADVANCED(o)                                                                                  \
  (Dx = _mm256_broadcastsd_pd(_mm256_castpd256_pd128(Nx)), Dy = _mm256_broadcastsd_pd(_mm256_castpd256_pd128(Ny)), \
   Mx = _mm256_blend_pd(PXv, Dx, 12), My = _mm256_blend_pd(PYv, Dy, 12),                                      \
   Nx = _mm256_blend_pd(_mm256_broadcast_sd(&(o).x), PXv, 8), Ny = _mm256_blend_pd(_mm256_broadcast_sd(&(o).y), PYv, 8), \
   Wb = _mm256_blend_pd(_mm256_permute4x64_pd(Wb, 0xEC), _mm256_broadcast_sd(&(o).n2), 8))
__attribute__((target("avx2"))) void Subdiv::flip_around_avx2(int curr_edge_i, int first_point, int curr_point, float px,
                                                              float py) {
#include "host_subdiv_flip.inc"
}
#undef AOS_GEOM_DECL
#undef AOS_GEOM_INIT
#undef AOS_GEOM_TEST
#undef AOS_GEOM_FLIPPED
#undef AOS_GEOM_ADVANCED
#endif

// intersection of the bisectors of (org0,dst0) and (org1,dst1): float differences and sums, double solve
bool Subdiv::voronoi_point(const Vertex &vo0, const Vertex &vd0, const Vertex &vo1, const Vertex &vd1, float *x, float *y) {
  struct P2f { float x, y; };  // Point2f arithmetic: the differences and sums below are float32 operations
  const P2f o0{(float)vo0.x, (float)vo0.y}, d0{(float)vd0.x, (float)vd0.y}, o1{(float)vo1.x, (float)vo1.y}, d1{(float)vd1.x, (float)vd1.y};
  double a0 = d0.x - o0.x;
  double b0 = d0.y - o0.y;
  double c0 = -0.5 * (a0 * (d0.x + o0.x) + b0 * (d0.y + o0.y));
  double a1 = d1.x - o1.x;
  double b1 = d1.y - o1.y;
  double c1 = -0.5 * (a1 * (d1.x + o1.x) + b1 * (d1.y + o1.y));
  double det = a0 * b1 - a1 * b0;
  if (det != 0) {
    det = 1. / det;
    *x = (float)((b0 * c1 - b1 * c0) * det);
    *y = (float)((a1 * c0 - a0 * c1) * det);
  } else {
    *x = FLT_MAX;
    *y = FLT_MAX;
  }
  return fabs(*x) < FLT_MAX * 0.5 && fabs(*y) < FLT_MAX * 0.5;
}

void Subdiv::calc_voronoi() {
  if (valid_geometry_) return;
  // clearVoronoi: nothing to clear, the structure is built once per compute()
  const int total = (int)(next_.size() / 4);
  for (int i = 4; i < total; ++i) {
    if (next_[4 * i] <= 0) continue;  // free quad-edge
    const int edge0 = i * 4;
    if (!pt_[4 * i + 3]) {
      int edge1 = get_edge(edge0, NEXT_AROUND_LEFT);
      int edge2 = get_edge(edge1, NEXT_AROUND_LEFT);
      float x, y;
      if (voronoi_point(vtx_[org(edge0)], vtx_[dst(edge0)], vtx_[org(edge1)], vtx_[dst(edge1)], &x, &y)) {
        int v = new_point(x, y, true);
        pt_[4 * i + 3] = pt_[(edge1 & ~3) + 3 - (edge1 & 2)] = pt_[(edge2 & ~3) + 3 - (edge2 & 2)] = v;
      }
    }
    if (!pt_[4 * i + 1]) {
      int edge1 = get_edge(edge0, NEXT_AROUND_RIGHT);
      int edge2 = get_edge(edge1, NEXT_AROUND_RIGHT);
      float x, y;
      if (voronoi_point(vtx_[org(edge0)], vtx_[dst(edge0)], vtx_[org(edge1)], vtx_[dst(edge1)], &x, &y)) {
        int v = new_point(x, y, true);
        pt_[4 * i + 1] = pt_[(edge1 & ~3) + 1 + (edge1 & 2)] = pt_[(edge2 & ~3) + 1 + (edge2 & 2)] = v;
      }
    }
  }
  valid_geometry_ = true;
}

void Subdiv::voronoi_facets(std::vector<float> *xy, std::vector<int32_t> *off) {
  calc_voronoi();
  xy->clear();
  off->clear();
  off->push_back(0);
  const size_t total = vtx_.size();
  for (size_t k = 4; k < total; ++k) {
    if (vtx_[k].type != 0) continue;  // free or virtual
    int edge = (vtx_[k].first_edge & ~3) + ((vtx_[k].first_edge + 1) & 3), t = edge;
    do {
      const Vertex &v = vtx_[org(t)];
      xy->push_back((float)v.x);
      xy->push_back((float)v.y);
      t = get_edge(t, NEXT_AROUND_LEFT);
    } while (t != edge);
    off->push_back((int32_t)(xy->size() / 2));
  }
}

}  // namespace aos
