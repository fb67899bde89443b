// TEST INFRASTRUCTURE ONLY — plain-struct stand-ins for the ROS message types named in vofod/voxel_map.h.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace std_msgs
{
struct Header { std::uint32_t seq = 0; double stamp = 0; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}
namespace geometry_msgs
{
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
}
namespace visualization_msgs
{
struct Marker
{
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6 };
  std_msgs::Header header;
  std::string ns;
  std::int32_t id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
};
}
