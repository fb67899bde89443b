// ouster_ros::Point cloud (the nodelet's input, include/vofod/types.h:7) -> packed vofod_pt records for vofod_process_scan.
// Field by field, so it follows whatever layout the caller's ouster_ros build has.  SURVEY.md §8f N1.
#pragma once
#include <vofod_cuda.h>

#include <vector>

namespace vofod_b200
{
template <class Cloud>
inline void pack_scan(const Cloud& cloud, std::vector<vofod_pt>& out)
{
  out.resize(cloud.points.size());
  for (size_t i = 0; i < out.size(); i++)
  {
    const auto& p = cloud.points[i];
    out[i] = vofod_pt{p.x, p.y, p.z, p.intensity, static_cast<uint32_t>(p.range)};
  }
}
}  // namespace vofod_b200
