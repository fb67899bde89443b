// TEST INFRASTRUCTURE ONLY — shadows the reference's include/vofod/point_types.h (which pulls in a dozen PCL template
// implementation headers and ouster_ros): declares just the point types the voxel-grid sources touch, with the field
// layout of the reference (include/vofod/point_types.h:51-56; ouster_ros::Point fields x,y,z,intensity,range).
#pragma once
#include <pcl/common/common.h>
namespace ouster_ros
{
struct alignas(16) Point
{
  PCL_ADD_POINT4D;
  float intensity = 0.f;
  std::uint32_t t = 0;
  std::uint16_t reflectivity = 0;
  std::uint8_t ring = 0;
  std::uint16_t ambient = 0;
  std::uint32_t range = 0;
};
}
namespace vofod
{
struct alignas(16) PointXYZR
{
  PCL_ADD_POINT4D;
  std::uint32_t range = 0;
};
struct alignas(16) PointXYZRI
{
  PCL_ADD_POINT4D;
  float intensity = 0.f;
  std::uint32_t range = 0;
};
}
