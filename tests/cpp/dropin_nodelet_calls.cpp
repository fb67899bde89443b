// Drop-in check: this translation unit sees ONLY include/vofod_dropin + include (never the reference's headers) and
// uses the GPU-backed classes under the names vofod_nodelet.cpp uses — vofod::VoxelMap, vofod::VoxelGridWeighted, vofod::VoxelGridCounted,
// load_cloud — through every member function the nodelet calls on them (grep -o 'm_voxel_[a-z]*\.[A-Za-z]*' src/vofod_nodelet.cpp: 28
// names), with the argument types of the nodelet's call sites (cited per block).  It compiling is the point; it also runs on the GPU box.
#include <vofod/pc_loader.h>
#include <vofod/voxel_grid_counted.h>
#include <vofod/voxel_grid_weighted.h>
#include <vofod/voxel_map.h>
#include <vofod_b200/ouster_pack.hpp>

#include <algorithm>
#include <cstdio>
#include <fstream>

namespace vofod
{
// include/vofod/types.h:7-17
using pt_t = ouster_ros::Point;
using pc_t = pcl::PointCloud<pt_t>;
using pt_XYZ_t = pcl::PointXYZ;
using pc_XYZ_t = pcl::PointCloud<pt_XYZ_t>;
using pt_XYZR_t = vofod::PointXYZR;
using pc_XYZR_t = pcl::PointCloud<pt_XYZR_t>;
using vec3_t = Eigen::Vector3f;
using vec3i_t = Eigen::Vector3i;
// vofod_nodelet.cpp:110-119: a VoxelMap by value inside a copyable struct
struct cluster_t
{
  VoxelMap submap;
  float obb_size = 0.f;
};
}  // namespace vofod

#define CHECK(cond)                                                 \
  do                                                                \
  {                                                                 \
    if (!(cond))                                                    \
    {                                                               \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                     \
    }                                                               \
  } while (0)

namespace vofod
{
// like the nodelet's member functions: inside namespace vofod, where vofod::pc_t hides the loader's global pc_t
int run()
{
  VoxelMap m_voxel_map, m_voxel_flags, m_voxel_raycast, local_vmap;  // :2333-2339, :1283
  // reset() :1616-1628
  m_voxel_map.resize(0.0f, 0.0f, 13.75f, 40.0f, 40.0f, 30.0f, 0.5f);
  m_voxel_map.setTo(-740.0f);
  const auto [sx, sy, sz] = m_voxel_map.sizesIdx();
  CHECK(sx == 81 && sy == 81 && sz == 61);
  m_voxel_flags.resizeAs(m_voxel_map);
  std_msgs::ColorRGBA col;
  col.a = 1.f;
  m_voxel_flags.addVisualizationThreshold(2.0f - 0.1f, col);
  m_voxel_flags.clear();
  m_voxel_raycast.resizeAs(m_voxel_map);
  local_vmap.resizeAs(m_voxel_map);  // :1287
  // rangefinder seed :599-610
  CHECK(m_voxel_map.inLimits(1.0f, 2.0f, 0.1f));
  {
    auto& mapval = m_voxel_map.at(1.0f, 2.0f, 0.1f);
    mapval = (mapval + 0.0) / 2.0;
  }
  // filterAndTransform :661-668
  pc_t::Ptr cloud = boost::make_shared<pc_t>();
  for (int i = 0; i < 2000; i++)
  {
    pt_t p;
    p.x = 0.01f * float(i % 700) - 3.0f;
    p.y = 0.02f * float(i % 300) - 2.0f;
    p.z = 0.1f;
    p.intensity = 100.f;
    p.range = 5000;
    cloud->points.push_back(p);
  }
  pcl::PointCloud<pt_XYZR_t>::Ptr cloud_weighted = boost::make_shared<pcl::PointCloud<pt_XYZR_t>>();
  {
    VoxelGridWeighted vgw;
    vgw.setInputCloud(cloud);
    vgw.setLeafSize(0.5f, 0.5f, 0.5f);
    const auto [align_x, align_y, align_z] = m_voxel_map.idxToCoord(0, 0, 0);
    vgw.setVoxelAlign({align_x, align_y, align_z, 0.0f});
    vgw.filter(*cloud_weighted);
  }
  CHECK(!cloud_weighted->points.empty());
  std::vector<vofod_pt> packed;
  vofod_b200::pack_scan(*cloud, packed);
  CHECK(packed.size() == 2000 && packed[7].range_mm == 5000 && packed[7].x == cloud->points[7].x);
  // findCloseFarClusters :712-735, updateVoxel :779-796
  const uint64_t n_bg_pts = m_voxel_map.nVoxelsOver(-300.0f);
  CHECK(n_bg_pts == 0);
  for (const auto& pt : cloud_weighted->points)
  {
    (void)m_voxel_map.hasCloseTo(pt.x, pt.y, pt.z, 1.5f, -300.0f);
    const auto [xc, yc, zc] = m_voxel_map.coordToIdx(pt.x, pt.y, pt.z);
    auto& mapval = m_voxel_map.atIdx(xc, yc, zc);
    const float w = 1.0f / static_cast<float>(1lu << std::clamp(pt.range, 0u, 63u));
    mapval = w * mapval + (1.0f - w) * 0.0f;
    m_voxel_flags.atIdx(xc, yc, zc) = 2.0f;
  }
  CHECK(m_voxel_map.nVoxelsOver(-300.0f) > 0);
  // raycast_cloud :1430-1602
  m_voxel_raycast.clear();
  const vec3_t origin_pt(0.2f, 0.1f, 5.0f), dir(0.6f, 0.0f, -0.8f);
  CHECK(m_voxel_raycast.inLimits(origin_pt.x(), origin_pt.y(), origin_pt.z()));
  m_voxel_raycast.forEachRay(origin_pt, dir, 4.0f, [&](const VoxelMap::coord_t val, const VoxelMap::idx_t x_idx, const VoxelMap::idx_t y_idx, const VoxelMap::idx_t z_idx) {
    auto& raycastval = m_voxel_raycast.atIdx(x_idx, y_idx, z_idx);
    raycastval += val;
  });
  const float max_val = *std::max_element(std::begin(m_voxel_raycast), std::end(m_voxel_raycast));
  CHECK(max_val > 0.0f);
  size_t touched = 0;
  m_voxel_flags.forEachIdx([&](VoxelMap::data_t& flag, const VoxelMap::idx_t xc, const VoxelMap::idx_t yc, const VoxelMap::idx_t zc) {
    float raycastval;
    if (flag == 0.0f && (raycastval = m_voxel_raycast.atIdx(xc, yc, zc)) > 0.0f)
    {
      auto& mapval = m_voxel_map.atIdx(xc, yc, zc);
      const float w1 = std::pow(2, -1.0f * 0.0035f * raycastval);
      mapval = w1 * mapval + (1.0f - w1) * -1000.0f;
      touched++;
    }
  });
  CHECK(touched >= 8);
  m_voxel_flags.clear();
  std_msgs::Header header;
  (void)m_voxel_raycast.visualization(header);
  (void)m_voxel_map.borderVisualization(header);  // :2079, :2091
  m_voxel_map.clearVisualizationThresholds();      // :1018-1023
  m_voxel_map.addVisualizationThreshold(-300.0f, col);
  CHECK(!m_voxel_map.visualization(header).points.empty());
  // classify_cluster :1699-1715, extractDetections :850-861
  {
    const auto [is_connected, explored_idxs] = m_voxel_map.exploreToGround(0.3f, 0.3f, 3.0f, -750.0f, -300.0f, 8);
    if (!is_connected)
      for (const auto& idx : explored_idxs)
        m_voxel_map.at(idx) = -750.0f;
    cluster_t cl;
    cl.submap = m_voxel_map.getSubmapCopy(vec3_t(-1.0f, -1.0f, 2.0f), vec3_t(1.0f, 1.0f, 3.0f), 2);
    std::vector<cluster_t> clusters;
    clusters.push_back(cl);  // copies the VoxelMap
    clusters.back().submap.at(0.1f, 0.1f, 2.5f) = -1000.0f;
    double uncertainty_score = 0.0;
    for (auto& val : clusters.back().submap)
      uncertainty_score += 1.0 - val / -1000.0;
    CHECK(uncertainty_score > 0.0 && cl.submap.size() == clusters.back().submap.size());
  }
  // updateSeparatedBGClusters :1148-1167, :1252-1267
  local_vmap.copyDataIdx(m_voxel_map);
  VoxelMap::pc_t::Ptr vmap_pc_raw = local_vmap.voxelsAsVoxelPC(-300.0f);
  CHECK(!vmap_pc_raw->empty());
  auto vmap_pc_ds = boost::make_shared<vofod::VoxelGridCounted::PointCloudOut>();
  vofod::VoxelGridCounted vgc(-0.1f);
  vgc.setInputCloud(vmap_pc_raw);
  vgc.setLeafSize(1.0f, 1.0f, 1.0f);
  vgc.filter(*vmap_pc_ds);
  CHECK(!vmap_pc_ds->empty());
  const vec3i_t pos = vmap_pc_ds->points[0].getVector3fMap().cast<int>();
  const vec3i_t pt = pos + vec3i_t(1, 0, -1);
  if (m_voxel_map.inLimitsIdx(pt))
  {
    auto& mapval = m_voxel_map.atIdx(pt.x(), pt.y(), pt.z());
    mapval = 0.5f * mapval + 0.5f * -1000.0f;
    const vec3_t coords = local_vmap.idxToCoord(pt);
    (void)coords;
  }
  // debug topics :1001, :1009
  (void)m_voxel_map.voxelsAsPC(-750.0f, false);
  (void)m_voxel_map.voxelsAsPC(-300.0f, true);
  // initialize_apriori_map :319-341
  {
    const char* path = "/tmp/vofod_dropin_cloud.xyz";
    std::ofstream(path) << "1.0 2.0 0.2\n1.1 2.1 0.2 extra\n\nbad line\n";
    const pc_XYZ_t::Ptr loaded_cloud = load_cloud(path);
    CHECK(loaded_cloud != nullptr && loaded_cloud->size() == 2);
    CHECK(load_cloud("/nonexistent/cloud.xyz") == nullptr);
    for (const auto& p : *loaded_cloud)
      if (m_voxel_map.inLimits(p.x, p.y, p.z))
        m_voxel_map.at(p.x, p.y, p.z) = std::numeric_limits<float>::infinity();
    CHECK(m_voxel_map.nVoxelsOver(1e30f) == 1);
  }
  std::printf("drop-in names cover the nodelet's calls: OK\n");
  return 0;
}
}  // namespace vofod

int main() { return vofod::run(); }
