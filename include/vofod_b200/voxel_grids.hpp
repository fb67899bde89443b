// vofod_b200::VoxelGridWeighted / VoxelGridCounted — the call interface of the reference's two voxel filters
// (include/vofod/voxel_grid_weighted.h:8-20, voxel_grid_counted.h:8-24: setInputCloud, setLeafSize, setVoxelAlign, filter)
// over libvofod_cuda (radix sort by voxel key + run-length reduce on the GPU).  Row N1 of SURVEY.md §8f.
// Input clouds are treated as PCL treats dense clouds plus the finite check: non-finite points are dropped.
#pragma once
#include <pcl/common/common.h>
#include <vofod_cuda.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace vofod_b200
{
// one library context per process for the filters that the nodelet default-constructs (`VoxelGridWeighted vgw;`, vofod_nodelet.cpp:661):
// they hold no state between calls, only workspace
inline vofod_ctx* shared_ctx(int device = 0)
{
  static vofod_ctx* ctx = nullptr;
  if (!ctx && vofod_create(device, &ctx) != VOFOD_OK)
    throw std::runtime_error(std::string("vofod_create: ") + vofod_last_error(nullptr));
  return ctx;
}

template <class PointIn, class PointOut>
class VoxelGridBase
{
public:
  using PointCloudIn = pcl::PointCloud<PointIn>;
  using PointCloudOut = pcl::PointCloud<PointOut>;
  explicit VoxelGridBase(vofod_ctx* ctx) : m_ctx(ctx) {}
  void setInputCloud(const typename PointCloudIn::ConstPtr& cloud) { m_input = cloud; }
  void setLeafSize(float lx, float ly, float lz)
  {
    if (lx != ly || ly != lz)
      throw std::invalid_argument("libvofod_cuda voxel grids take one leaf size (the nodelet only ever sets lx = ly = lz)");
    m_leaf = lx;
  }
  void setVoxelAlign(const Eigen::Vector4f& c)
  {
    m_align = true;
    m_center[0] = c[0];
    m_center[1] = c[1];
    m_center[2] = c[2];
  }

protected:
  void emit(const std::vector<vofod_vox>& v, size_t m, PointCloudOut& output) const
  {
    output.header = m_input->header;
    output.height = 1;
    output.is_dense = true;
    output.points.resize(m);
    for (size_t i = 0; i < m; i++)
    {
      output.points[i].x = v[i].x;
      output.points[i].y = v[i].y;
      output.points[i].z = v[i].z;
      output.points[i].range = v[i].count;
    }
    output.width = static_cast<std::uint32_t>(m);
  }
  void ck(int rc) const
  {
    if (rc < 0)
      throw std::runtime_error(std::string("libvofod_cuda: ") + vofod_last_error(m_ctx));
  }
  vofod_ctx* m_ctx;
  typename PointCloudIn::ConstPtr m_input;
  float m_leaf = 0.f;
  bool m_align = false;
  float m_center[3] = {0.f, 0.f, 0.f};
};

// PointIn needs x,y,z (ouster_ros::Point in the nodelet); PointOut needs x,y,z,range (vofod::PointXYZR)
template <class PointIn, class PointOut>
class VoxelGridWeighted : public VoxelGridBase<PointIn, PointOut>
{
  using B = VoxelGridBase<PointIn, PointOut>;

public:
  using B::B;
  void filter(typename B::PointCloudOut& output)
  {
    if (!this->m_input)
      return;
    const size_t n = this->m_input->points.size();
    std::vector<float> xyz(3 * n);
    for (size_t i = 0; i < n; i++)
    {
      xyz[3 * i] = this->m_input->points[i].x;
      xyz[3 * i + 1] = this->m_input->points[i].y;
      xyz[3 * i + 2] = this->m_input->points[i].z;
    }
    std::vector<vofod_vox> out(n ? n : 1);
    size_t m = 0;
    this->ck(vofod_voxel_grid_weighted(this->m_ctx, xyz.data(), n, this->m_leaf, this->m_align ? this->m_center : nullptr, out.data(), out.size(), &m));
    this->emit(out, m, output);
  }
};

// PointIn needs x,y,z,intensity (pcl::PointXYZI)
template <class PointIn, class PointOut>
class VoxelGridCounted : public VoxelGridBase<PointIn, PointOut>
{
  using B = VoxelGridBase<PointIn, PointOut>;

public:
  VoxelGridCounted(vofod_ctx* ctx, float threshold) : B(ctx), m_threshold(threshold) {}
  void filter(typename B::PointCloudOut& output)
  {
    if (!this->m_input)
      return;
    const size_t n = this->m_input->points.size();
    std::vector<vofod_xyzi> in(n);
    for (size_t i = 0; i < n; i++)
      in[i] = vofod_xyzi{this->m_input->points[i].x, this->m_input->points[i].y, this->m_input->points[i].z, this->m_input->points[i].intensity};
    std::vector<vofod_vox> out(n ? n : 1);
    size_t m = 0;
    this->ck(vofod_voxel_grid_counted(this->m_ctx, in.data(), n, this->m_leaf, m_threshold, this->m_align ? this->m_center : nullptr, out.data(), out.size(), &m));
    this->emit(out, m, output);
  }

private:
  float m_threshold;
};
}  // namespace vofod_b200
