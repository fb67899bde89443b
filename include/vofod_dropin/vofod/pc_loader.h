// Drop-in for include/vofod/pc_loader.h: load_cloud over vofod_load_cloud (same tokenizer, nullptr when the file cannot be opened).
#pragma once
#include <pcl/common/common.h>
#include <vofod_cuda.h>

#include <string>
#include <vector>

using pt_t = pcl::PointXYZ;
using pc_t = pcl::PointCloud<pt_t>;

inline pc_t::Ptr load_cloud(const std::string& filename)
{
  size_t n = 0;
  int rc = vofod_load_cloud(filename.c_str(), nullptr, 0, &n);
  if (rc == VOFOD_E_IO)
    return nullptr;
  std::vector<float> xyz(3 * n);
  if (n)
    rc = vofod_load_cloud(filename.c_str(), xyz.data(), n, &n);
  pc_t::Ptr cloud = boost::make_shared<pc_t>();
  cloud->reserve(n);
  for (size_t i = 0; i < n; i++)
    cloud->push_back(pt_t(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
  cloud->width = static_cast<std::uint32_t>(cloud->size());
  cloud->height = 1;
  cloud->is_dense = true;
  return cloud;
}
