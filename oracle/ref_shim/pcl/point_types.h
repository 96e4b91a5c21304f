// PCL stand-in for oracle/_ref (TEST INFRASTRUCTURE ONLY).  PCL is absent from the image, so the three PCL pieces the
// reference calls are RESTATED here (parity with the real PCL 1.12 is unpinned; see oracle/aos_oracle.h):
//   pcl::PointXYZ / PointCloud (16-byte x y z pad), pcl::fromROSMsg (field-offset copy),
//   pcl::PassThrough (seed_gen:459-477; keeps finite points with min <= field <= max, limits held as float),
//   pcl::RadiusOutlierRemoval (seed_gen:236-242; forwarded to the hook = the oracle's brute-force restatement).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ref_shim_hooks.hpp"
#include "ref_shim_msgs.hpp"

namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0, y = 0, z = 0, pad = 1.f;
  PointXYZ() = default;
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {}
};
template <class P>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<P>>;
  using ConstPtr = std::shared_ptr<const PointCloud<P>>;
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const P &p) { points.push_back(p); }
  void clear() { points.clear(); width = height = 0; }
  typename std::vector<P>::const_iterator begin() const { return points.begin(); }
  typename std::vector<P>::const_iterator end() const { return points.end(); }
};

template <class P>
void fromROSMsg(const sensor_msgs::msg::PointCloud2 &msg, PointCloud<P> &out) {
  uint32_t ox = 0, oy = 4, oz = 8;
  for (const auto &f : msg.fields) {
    if (f.name == "x") ox = f.offset;
    else if (f.name == "y") oy = f.offset;
    else if (f.name == "z") oz = f.offset;
  }
  size_t n = (size_t)msg.width * msg.height;
  out.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const uint8_t *b = msg.data.data() + i * msg.point_step;
    std::memcpy(&out.points[i].x, b + ox, 4);
    std::memcpy(&out.points[i].y, b + oy, 4);
    std::memcpy(&out.points[i].z, b + oz, 4);
  }
  out.width = msg.width; out.height = msg.height; out.is_dense = msg.is_dense;
}

template <class P>
class PassThrough {
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr &c) { in_ = c; }
  void setFilterFieldName(const std::string &f) { field_ = f; }
  void setFilterLimits(const float &lo, const float &hi) { lo_ = lo; hi_ = hi; }
  void filter(PointCloud<P> &out) {
    // the output may be the input cloud (seed_gen:466-477 filters cloud_filtered into itself): build, then swap
    std::vector<P> kept;
    kept.reserve(in_->points.size());
    for (const P &p : in_->points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      float v = field_ == "x" ? p.x : field_ == "y" ? p.y : p.z;
      if (!std::isfinite(v)) continue;
      if (v < lo_ || v > hi_) continue;
      kept.push_back(p);
    }
    out.points.swap(kept);
    out.width = (uint32_t)out.points.size();
    out.height = 1;
    out.is_dense = true;
  }
 private:
  typename PointCloud<P>::ConstPtr in_;
  std::string field_;
  float lo_ = -3.4028235e38f, hi_ = 3.4028235e38f;
};

template <class P>
class RadiusOutlierRemoval {
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr &c) { in_ = c; }
  void setRadiusSearch(double r) { radius_ = r; }
  void setMinNeighborsInRadius(int k) { min_pts_ = k; }
  void filter(PointCloud<P> &out) {
    size_t n = in_->points.size();
    std::vector<uint8_t> keep(n, 1);
    if (ref_hooks()->ror && n) ref_hooks()->ror(&in_->points[0].x, (int)n, radius_, min_pts_, keep.data());
    std::vector<P> kept;
    for (size_t i = 0; i < n; ++i) if (keep[i]) kept.push_back(in_->points[i]);
    out.points.swap(kept);
    out.width = (uint32_t)out.points.size();
    out.height = 1;
    out.is_dense = true;
  }
 private:
  typename PointCloud<P>::ConstPtr in_;
  double radius_ = 0;
  int min_pts_ = 1;
};
}  // namespace pcl
