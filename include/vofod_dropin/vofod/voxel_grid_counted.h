// Drop-in for include/vofod/voxel_grid_counted.h (see vofod/voxel_map.h in this directory).
#pragma once
#include <vofod_b200/voxel_grids.hpp>

#include "vofod/point_types.h"

namespace vofod
{
class VoxelGridCounted : public vofod_b200::VoxelGridCounted<pcl::PointXYZI, vofod::PointXYZR>
{
  using Base = vofod_b200::VoxelGridCounted<pcl::PointXYZI, vofod::PointXYZR>;

public:
  using PointT = pcl::PointXYZI;
  using PointCloudOut = pcl::PointCloud<vofod::PointXYZR>;
  explicit VoxelGridCounted(const float threshold) : Base(vofod_b200::shared_ctx(), threshold) {}
};
}  // namespace vofod
