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
