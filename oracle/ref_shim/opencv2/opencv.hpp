// OpenCV stand-in for oracle/_ref (TEST INFRASTRUCTURE ONLY).  cv::Mat is a plain uint8 matrix; the image
// operations the reference calls (/root/reference/src/aos_seed_gen_node.cpp:678-684, src/utils/voronoi_diagram.cpp:
// 51-94) are forwarded to the real cv2 of this image through ref_shim_hooks.hpp.  Calls that sit on branches the
// reference can never take (seed_gen:687-699, 802-821) abort loudly instead of pretending.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "ref_shim_hooks.hpp"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {
struct Exception : public std::exception {
  std::string msg;
  explicit Exception(const std::string &m = "cv::Exception") : msg(m) {}
  const char *what() const noexcept override { return msg.c_str(); }
};
template <class T> struct Point_ {
  T x = 0, y = 0;
  Point_() = default;
  Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <class T> struct Size_ {
  T width = 0, height = 0;
  Size_() = default;
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
inline int cvRound(double v) { return (int)std::lrint(v); }  // round half to even, as OpenCV's SSE2 path
template <class T> struct Rect_ {
  T x = 0, y = 0, width = 0, height = 0;
  Rect_() = default;
  Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
  template <class U> static U sat(T v) {  // cv::saturate_cast<U>: float/double -> int rounds (cvRound)
    if constexpr (std::is_integral<U>::value && std::is_floating_point<T>::value) return (U)cvRound((double)v);
    else return static_cast<U>(v);
  }
  template <class U> operator Rect_<U>() const { return Rect_<U>(sat<U>(x), sat<U>(y), sat<U>(width), sat<U>(height)); }
};
typedef Rect_<int> Rect;
typedef Rect_<float> Rect2f;
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };

struct Mat {
  int rows = 0, cols = 0;
  std::shared_ptr<std::vector<uchar>> buf;
  int se_shape = -1, se_w = 0, se_h = 0;  // set when the Mat is a structuring element
  Mat() = default;
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), buf(std::make_shared<std::vector<uchar>>((size_t)r * c)) {}
  bool empty() const { return rows == 0 || cols == 0 || !buf; }
  template <class T> T &at(int y, int x) { return reinterpret_cast<T &>((*buf)[(size_t)y * cols + x]); }
  template <class T> const T &at(int y, int x) const { return reinterpret_cast<const T &>((*buf)[(size_t)y * cols + x]); }
  uchar *data() { return buf->data(); }
  const uchar *data() const { return buf->data(); }
};

enum MorphShapes { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum { DIST_L2 = 2, NORM_MINMAX = 32, THRESH_BINARY = 0, LINE_8 = 8 };

inline Mat getStructuringElement(int shape, Size k) {
  Mat m(k.height, k.width, CV_8UC1);
  m.se_shape = shape; m.se_w = k.width; m.se_h = k.height;
  return m;
}
inline void morphologyEx(const Mat &src, Mat &dst, int op, const Mat &kernel) {
  if (!ref_hooks()->morph || kernel.se_shape < 0) { std::fprintf(stderr, "ref_shim: morphologyEx hook missing\n"); std::abort(); }
  Mat out(src.rows, src.cols, CV_8UC1);
  ref_hooks()->morph(src.data(), src.rows, src.cols, op, kernel.se_shape, kernel.se_w, kernel.se_h, out.data());
  dst = out;
}
[[noreturn]] inline void ref_shim_dead(const char *what) {
  std::fprintf(stderr, "ref_shim: %s reached -- a branch the reference never takes\n", what);
  std::abort();
}
inline void distanceTransform(const Mat &, Mat &, int, int) { ref_shim_dead("cv::distanceTransform"); }
inline void normalize(const Mat &, Mat &, double, double, int, int) { ref_shim_dead("cv::normalize"); }
inline double threshold(const Mat &, Mat &, double, double, int) { ref_shim_dead("cv::threshold"); }
inline void polylines(Mat &, const Point *const *, const int *, int, bool, const Scalar &, int, int) { ref_shim_dead("cv::polylines"); }

namespace ximgproc {
enum ThinningTypes { THINNING_ZHANGSUEN = 0, THINNING_GUOHALL = 1 };
inline void thinning(const Mat &src, Mat &dst, int type) {
  if (!ref_hooks()->thin) { std::fprintf(stderr, "ref_shim: thinning hook missing\n"); std::abort(); }
  Mat out(src.rows, src.cols, CV_8UC1);
  ref_hooks()->thin(src.data(), src.rows, src.cols, type, out.data());
  dst = out;
}
}  // namespace ximgproc

// cv::Subdiv2D: points are collected and the whole insertion sequence is replayed by the real cv2.Subdiv2D when the
// facets are requested (an insert that throws inside OpenCV is skipped there exactly as vd:83-88 skips it).
class Subdiv2D {
 public:
  explicit Subdiv2D(Rect r) : rect_(r) {}
  int insert(Point2f p) { pts_.push_back(p.x); pts_.push_back(p.y); return (int)(pts_.size() / 2) + 3; }
  void getVoronoiFacetList(const std::vector<int> &idx, std::vector<std::vector<Point2f>> &facets, std::vector<Point2f> &centers) {
    if (!idx.empty()) ref_shim_dead("Subdiv2D::getVoronoiFacetList(idx != {})");
    if (!ref_hooks()->subdiv) { std::fprintf(stderr, "ref_shim: subdiv hook missing\n"); std::abort(); }
    int rect[4] = {rect_.x, rect_.y, rect_.width, rect_.height};
    int nf = 0, nv = 0;
    ref_hooks()->subdiv(rect, pts_.data(), (int)(pts_.size() / 2), &nf, &nv, nullptr, nullptr, nullptr);
    std::vector<int> sizes((size_t)nf + 1);
    std::vector<float> xy((size_t)2 * nv + 2), cen((size_t)2 * nf + 2);
    ref_hooks()->subdiv(rect, pts_.data(), (int)(pts_.size() / 2), &nf, &nv, sizes.data(), xy.data(), cen.data());
    facets.assign((size_t)nf, {});
    centers.assign((size_t)nf, Point2f());
    size_t k = 0;
    for (int f = 0; f < nf; ++f) {
      facets[f].reserve((size_t)sizes[f]);
      for (int i = 0; i < sizes[f]; ++i, ++k) facets[f].emplace_back(xy[2 * k], xy[2 * k + 1]);
      centers[f] = Point2f(cen[2 * f], cen[2 * f + 1]);
    }
  }
 private:
  Rect rect_;
  std::vector<float> pts_;
};
}  // namespace cv
