// ref_shim_msgs.hpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref).
// Plain-struct stand-ins for the ROS 2 message types the reference's two hot-path nodes use
// (/root/reference/src/aos_seed_gen_node.cpp:7-16, src/aos_gvd_node.cpp:10-19 include lists).  Field names and
// C++ types are those rosidl generates for ROS 2 Humble (std_msgs/Header, geometry_msgs/Point = 3 x float64,
// Point32 = 3 x float32, nav_msgs/OccupancyGrid.data = int8[], MapMetaData.resolution = float32,
// width/height = uint32): the reference's arithmetic depends on exactly those widths.  Nothing here is copied
// from the reference; the reference sources are compiled from /root/reference where they lie (oracle/Makefile).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace builtin_interfaces { namespace msg {
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
struct Duration { int32_t sec = 0; uint32_t nanosec = 0; };
}}

#define REF_SHIM_MSG_PTRS(T)                 \
  using SharedPtr = std::shared_ptr<T>;      \
  using ConstSharedPtr = std::shared_ptr<const T>; \
  using UniquePtr = std::unique_ptr<T>;

namespace std_msgs { namespace msg {
struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; REF_SHIM_MSG_PTRS(Header) };
struct MultiArrayDimension { std::string label; uint32_t size = 0, stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; uint32_t data_offset = 0; };
struct Float64MultiArray { MultiArrayLayout layout; std::vector<double> data; REF_SHIM_MSG_PTRS(Float64MultiArray) };
struct Bool { bool data = false; REF_SHIM_MSG_PTRS(Bool) };
struct Int32 { int32_t data = 0; REF_SHIM_MSG_PTRS(Int32) };
struct String { std::string data; REF_SHIM_MSG_PTRS(String) };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}}

namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; REF_SHIM_MSG_PTRS(Point) };
struct Point32 { float x = 0, y = 0, z = 0; REF_SHIM_MSG_PTRS(Point32) };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; REF_SHIM_MSG_PTRS(Pose) };
struct PoseArray { std_msgs::msg::Header header; std::vector<Pose> poses; REF_SHIM_MSG_PTRS(PoseArray) };
struct PoseStamped { std_msgs::msg::Header header; Pose pose; REF_SHIM_MSG_PTRS(PoseStamped) };
struct Polygon { std::vector<Point32> points; };
struct PolygonStamped { std_msgs::msg::Header header; Polygon polygon; REF_SHIM_MSG_PTRS(PolygonStamped) };
}}

namespace nav_msgs { namespace msg {
struct MapMetaData {
  builtin_interfaces::msg::Time map_load_time;
  float resolution = 0.f;
  uint32_t width = 0, height = 0;
  geometry_msgs::msg::Pose origin;
};
struct OccupancyGrid {
  std_msgs::msg::Header header;
  MapMetaData info;
  std::vector<int8_t> data;
  REF_SHIM_MSG_PTRS(OccupancyGrid)
};
}}

namespace nav_msgs { namespace msg {
struct Path { std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses; REF_SHIM_MSG_PTRS(Path) };
}}

namespace std_srvs { namespace srv {
struct Empty {
  struct Request { REF_SHIM_MSG_PTRS(Request) };
  struct Response { REF_SHIM_MSG_PTRS(Response) };
};
}}
namespace lio_sam_wo { namespace srv {   // external package (package.xml:38), only its type name is needed
struct SaveMap {
  struct Request { float resolution = 0; std::string destination; REF_SHIM_MSG_PTRS(Request) };
  struct Response { bool success = false; REF_SHIM_MSG_PTRS(Response) };
};
}}

namespace sensor_msgs { namespace msg {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0; };
struct PointCloud2 {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
  REF_SHIM_MSG_PTRS(PointCloud2)
};
}}

namespace visualization_msgs { namespace msg {
struct Marker {
  enum : int32_t { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6,
                   SPHERE_LIST = 7, POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
  enum : int32_t { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::msg::Header header;
  std::string ns;
  int32_t id = 0;
  int32_t type = 0;
  int32_t action = 0;
  geometry_msgs::msg::Pose pose;
  geometry_msgs::msg::Vector3 scale;
  std_msgs::msg::ColorRGBA color;
  builtin_interfaces::msg::Duration lifetime;
  bool frame_locked = false;
  std::vector<geometry_msgs::msg::Point> points;
  std::vector<std_msgs::msg::ColorRGBA> colors;
  std::string text;
  REF_SHIM_MSG_PTRS(Marker)
};
struct MarkerArray { std::vector<Marker> markers; REF_SHIM_MSG_PTRS(MarkerArray) };
}}
