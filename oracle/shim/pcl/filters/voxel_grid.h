// TEST INFRASTRUCTURE ONLY — the members of pcl::PCLBase / pcl::Filter / pcl::VoxelGrid (PCL 1.10) that the reference's
// VoxelGridWeighted / VoxelGridCounted subclasses read and write.
#pragma once
#include <pcl/common/common.h>

namespace pcl
{
template <class PointT>
class VoxelGrid
{
public:
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  VoxelGrid() : leaf_size_(0.f, 0.f, 0.f, 0.f), inverse_leaf_size_(0.f, 0.f, 0.f, 0.f), min_points_per_voxel_(0) {}
  virtual ~VoxelGrid() {}
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  // voxel_grid.h: leaf_size_[3] is forced to 1 "to avoid division by zero"; inverse = Ones()/leaf_size_.array()
  void setLeafSize(float lx, float ly, float lz)
  {
    leaf_size_ = Eigen::Vector4f(lx, ly, lz, 1.f);
    inverse_leaf_size_ = Eigen::Vector4f(1.0f / lx, 1.0f / ly, 1.0f / lz, 1.0f / 1.f);
  }

  // pcl::VoxelGrid<PointT>::applyFilter (filters/impl/voxel_grid.hpp, PCL 1.10) as the nodelet uses the stock class
  // (vofod_nodelet.cpp:331-335): no field filter, min_points_per_voxel 0, downsample_all_data -> CentroidPoint:
  // per leaf the fp32 sum of the member points IN THE ORDER std::sort LEFT THEM (the sort compares the leaf index only,
  // so the order inside a leaf is whatever libstdc++'s introsort produces) divided by the count.
  void filter(PointCloud<PointT>& output)
  {
    if (!initCompute())
      return;
    output.points.clear();
    output.height = 1;
    output.is_dense = true;
    Eigen::Vector4f min_p, max_p;
    getMinMax3D<PointT>(*input_, *indices_, min_p, max_p);
    const std::int64_t dx = static_cast<std::int64_t>((max_p[0] - min_p[0]) * inverse_leaf_size_[0]) + 1;
    const std::int64_t dy = static_cast<std::int64_t>((max_p[1] - min_p[1]) * inverse_leaf_size_[1]) + 1;
    const std::int64_t dz = static_cast<std::int64_t>((max_p[2] - min_p[2]) * inverse_leaf_size_[2]) + 1;
    if ((dx * dy * dz) > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max()))
    {
      output = *input_;  // "Leaf size is too small": PCL warns and passes the cloud through
      return;
    }
    int mnb[3], dv[3];
    for (int a = 0; a < 3; a++)
    {
      mnb[a] = static_cast<int>(std::floor(min_p[a] * inverse_leaf_size_[a]));
      dv[a] = static_cast<int>(std::floor(max_p[a] * inverse_leaf_size_[a])) - mnb[a] + 1;
    }
    const int mul[3] = {1, dv[0], dv[0] * dv[1]};
    struct cpi { unsigned idx; unsigned src; bool operator<(const cpi& o) const { return idx < o.idx; } };
    std::vector<cpi> iv;
    iv.reserve(indices_->size());
    for (const int i : *indices_)
    {
      const PointT& p = input_->points[i];
      if (!input_->is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)))
        continue;
      const int ijk0 = static_cast<int>(std::floor(p.x * inverse_leaf_size_[0]) - static_cast<float>(mnb[0]));
      const int ijk1 = static_cast<int>(std::floor(p.y * inverse_leaf_size_[1]) - static_cast<float>(mnb[1]));
      const int ijk2 = static_cast<int>(std::floor(p.z * inverse_leaf_size_[2]) - static_cast<float>(mnb[2]));
      iv.push_back(cpi{static_cast<unsigned>(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), static_cast<unsigned>(i)});
    }
    std::sort(iv.begin(), iv.end(), std::less<cpi>());
    std::size_t first = 0;
    while (first < iv.size())
    {
      std::size_t last = first + 1;
      while (last < iv.size() && iv[last].idx == iv[first].idx)
        ++last;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (std::size_t k = first; k < last; k++)
      {
        const PointT& p = input_->points[iv[k].src];
        sx += p.x; sy += p.y; sz += p.z;
      }
      const float n = static_cast<float>(last - first);
      PointT o;
      o.x = sx / n; o.y = sy / n; o.z = sz / n;
      output.points.push_back(o);
      first = last;
    }
    output.width = static_cast<std::uint32_t>(output.points.size());
  }

protected:
  // PCLBase::initCompute: without setIndices() a fake index list 0..N-1 is generated
  bool initCompute()
  {
    if (!input_)
      return false;
    if (!indices_)
      indices_ = boost::make_shared<std::vector<int>>();
    indices_->resize(input_->points.size());
    for (std::size_t i = 0; i < indices_->size(); i++)
      (*indices_)[i] = int(i);
    return true;
  }
  bool deinitCompute() { return true; }
  const std::string& getClassName() const { return filter_name_; }

  PointCloudConstPtr input_;
  boost::shared_ptr<std::vector<int>> indices_;
  std::string filter_name_;
  Eigen::Vector4f leaf_size_, inverse_leaf_size_;
  Eigen::Vector4i min_b_, max_b_, div_b_, divb_mul_;
  unsigned int min_points_per_voxel_;
};
}  // namespace pcl
