// Drop-in for include/vofod/voxel_grid_weighted.h (see vofod/voxel_map.h in this directory): default-constructible like the
// reference's class (vofod_nodelet.cpp:661), working on the process-wide library context.
#pragma once
#include <vofod_b200/voxel_grids.hpp>

#include "vofod/point_types.h"

namespace vofod
{
class VoxelGridWeighted : public vofod_b200::VoxelGridWeighted<ouster_ros::Point, vofod::PointXYZR>
{
  using Base = vofod_b200::VoxelGridWeighted<ouster_ros::Point, vofod::PointXYZR>;

public:
  using PointT = ouster_ros::Point;
  using PointCloudOut = pcl::PointCloud<vofod::PointXYZR>;
  VoxelGridWeighted() : Base(vofod_b200::shared_ctx()) {}
};
}  // namespace vofod
