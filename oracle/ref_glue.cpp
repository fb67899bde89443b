// TEST INFRASTRUCTURE ONLY — C entry points over the REFERENCE's own vofod::VoxelMap / VoxelGridWeighted /
// VoxelGridCounted (compiled from /root/reference/src where they lie, against oracle/shim) so that tests can run the
// reference implementation call by call next to the oracle's restatement.  Built only where /root/reference exists.
#include <vofod/voxel_grid_counted.h>
#include <vofod/voxel_grid_weighted.h>
#include <vofod/voxel_map.h>

#include <cstring>

#include "../include/vofod_cuda.h"

using vofod::VoxelMap;

extern "C" {
VoxelMap* vr_create() { return new VoxelMap(); }
void vr_destroy(VoxelMap* m) { delete m; }
void vr_resize(VoxelMap* m, const float c[3], const float d[3], float vs) { m->resize(VoxelMap::vec3_t(c[0], c[1], c[2]), VoxelMap::vec3_t(d[0], d[1], d[2]), vs); }
void vr_resize_idx(VoxelMap* m, const float o[3], const int32_t s[3], float vs) { m->resize(VoxelMap::vec3_t(o[0], o[1], o[2]), VoxelMap::vec3i_t(s[0], s[1], s[2]), vs); }
void vr_info(VoxelMap* m, vofod_map_info* out)
{
  const auto o = m->origin();
  const auto s = m->sizes();
  for (int a = 0; a < 3; a++) { out->offset[a] = o[a]; out->sizes[a] = s[a]; }
  out->n_cells = m->size();
  out->voxel_size = m->dimensions()[0] / float(s[0]);
  out->slab_axis = 0; out->slab_lo = 0; out->slab_hi = s[0];
  for (int a = 0; a < 3; a++) { out->storage_lo[a] = 0; out->storage_size[a] = s[a]; }
  out->_pad = 0;
}
float* vr_data(VoxelMap* m) { return &*m->begin(); }
void vr_set_to(VoxelMap* m, float v) { m->setTo(v); }
uint64_t vr_count_over(VoxelMap* m, float thr) { return m->nVoxelsOver(thr); }
size_t vr_compact_over(VoxelMap* m, float thr, int greater, int metric, vofod_xyzi* out, size_t cap)
{
  const auto pc = metric ? m->voxelsAsPC(thr, greater != 0) : m->voxelsAsVoxelPC(thr, greater != 0);
  const size_t n = pc->size();
  for (size_t i = 0; i < n && i < cap; i++)
    out[i] = vofod_xyzi{pc->points[i].x, pc->points[i].y, pc->points[i].z, pc->points[i].intensity};
  return n;
}
void vr_has_close_to(VoxelMap* m, const float* xyz, size_t n, float max_dist, float thr, uint8_t* out)
{
  for (size_t i = 0; i < n; i++)
    out[i] = m->hasCloseTo(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], max_dist, thr);
}
size_t vr_explore_to_ground(VoxelMap* m, const float pt[3], float unk, float gnd, float maxd, int* connected, int32_t* idx3, size_t cap)
{
  const auto [c, v] = m->exploreToGround(pt[0], pt[1], pt[2], unk, gnd, maxd);
  *connected = c;
  for (size_t i = 0; i < v.size() && i < cap; i++)
  {
    idx3[3 * i] = std::get<0>(v[i]); idx3[3 * i + 1] = std::get<1>(v[i]); idx3[3 * i + 2] = std::get<2>(v[i]);
  }
  return v.size();
}
void vr_is_floating(VoxelMap* m, const float* xyz, size_t n, float thr, uint8_t* out)
{
  for (size_t i = 0; i < n; i++)
    out[i] = m->isFloating(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], thr);
}
size_t vr_submap_copy(VoxelMap* m, const float mn[3], const float mx[3], int inflate, float* out, size_t cap, int32_t sizes[3], float offset[3])
{
  VoxelMap s = m->getSubmapCopy(VoxelMap::vec3_t(mn[0], mn[1], mn[2]), VoxelMap::vec3_t(mx[0], mx[1], mx[2]), inflate);
  const auto sz = s.sizes();
  const auto of = s.origin();
  for (int a = 0; a < 3; a++) { sizes[a] = sz[a]; offset[a] = of[a]; }
  const size_t n = s.size();
  if (out && n <= cap)
    std::memcpy(out, &*s.begin(), n * sizeof(float));
  return n;
}
size_t vr_trace_ray(VoxelMap* m, const float start[3], const float dir[3], float length, float* ddist, int32_t* idx3, size_t cap)
{
  size_t n = 0;
  m->forEachRay(VoxelMap::vec3_t(start[0], start[1], start[2]), VoxelMap::vec3_t(dir[0], dir[1], dir[2]), length, [&](float d, int x, int y, int z) {
    if (n < cap) { ddist[n] = d; idx3[3 * n] = x; idx3[3 * n + 1] = y; idx3[3 * n + 2] = z; }
    n++;
  });
  return n;
}
void vr_coord_to_idx(VoxelMap* m, const float* xyz, size_t n, int32_t* idx3)
{
  for (size_t i = 0; i < n; i++)
  {
    const auto [x, y, z] = m->coordToIdx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    idx3[3 * i] = x; idx3[3 * i + 1] = y; idx3[3 * i + 2] = z;
  }
}
void vr_idx_to_coord(VoxelMap* m, const int32_t* idx3, size_t n, float* xyz)
{
  for (size_t i = 0; i < n; i++)
  {
    const auto [x, y, z] = m->idxToCoord(idx3[3 * i], idx3[3 * i + 1], idx3[3 * i + 2]);
    xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
  }
}
// raycast accumulate exactly as vofod_nodelet.cpp:1484-1489 does with the reference's forEachRay: raycast[v] += ddist
uint64_t vr_accumulate_rays(VoxelMap* m, const float* starts, const float* dirs, const float* lens, size_t n)
{
  uint64_t trav = 0;
  for (size_t i = 0; i < n; i++)
    m->forEachRay(VoxelMap::vec3_t(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]), VoxelMap::vec3_t(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), lens[i],
                  [&](float d, int x, int y, int z) { m->atIdx(x, y, z) += d; trav++; });
  return trav;
}

// same traversal, counting callbacks per voxel (float counts are exact below 2^24)
uint64_t vr_count_rays(VoxelMap* m, const float* starts, const float* dirs, const float* lens, size_t n)
{
  uint64_t trav = 0;
  for (size_t i = 0; i < n; i++)
    m->forEachRay(VoxelMap::vec3_t(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]), VoxelMap::vec3_t(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), lens[i],
                  [&](float, int x, int y, int z) { m->atIdx(x, y, z) += 1.0f; trav++; });
  return trav;
}

int vr_voxel_grid_weighted(const float* xyz, size_t n, float leaf, const float* align, int dense, vofod_vox* out, size_t cap, size_t* m)
{
  auto in = boost::make_shared<pcl::PointCloud<ouster_ros::Point>>();
  in->points.resize(n);
  for (size_t i = 0; i < n; i++) { in->points[i].x = xyz[3 * i]; in->points[i].y = xyz[3 * i + 1]; in->points[i].z = xyz[3 * i + 2]; }
  in->is_dense = dense != 0;
  vofod::VoxelGridWeighted vg;
  vg.setInputCloud(in);
  vg.setLeafSize(leaf, leaf, leaf);
  // align_voxels_ is uninitialised in the reference's class (voxel_grid_weighted.h:18); the nodelet always sets it
  vg.setVoxelAlign(align ? Eigen::Vector4f(align[0], align[1], align[2], 0.f) : Eigen::Vector4f(leaf / 2, leaf / 2, leaf / 2, 0.f));
  pcl::PointCloud<vofod::PointXYZR> res;
  vg.filter(res);
  *m = res.points.size();
  for (size_t i = 0; i < res.points.size() && i < cap; i++)
    out[i] = vofod_vox{res.points[i].x, res.points[i].y, res.points[i].z, res.points[i].range};
  return 0;
}
int vr_voxel_grid_counted(const vofod_xyzi* pts, size_t n, float leaf, float thr, const float* align, int dense, vofod_vox* out, size_t cap, size_t* m)
{
  auto in = boost::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
  in->points.resize(n);
  for (size_t i = 0; i < n; i++) { in->points[i].x = pts[i].x; in->points[i].y = pts[i].y; in->points[i].z = pts[i].z; in->points[i].intensity = pts[i].intensity; }
  in->is_dense = dense != 0;
  vofod::VoxelGridCounted vg(thr);
  vg.setInputCloud(in);
  vg.setLeafSize(leaf, leaf, leaf);
  if (align)
    vg.setVoxelAlign(Eigen::Vector4f(align[0], align[1], align[2], 0.f));
  pcl::PointCloud<vofod::PointXYZR> res;
  vg.filter(res);
  *m = res.points.size();
  for (size_t i = 0; i < res.points.size() && i < cap; i++)
    out[i] = vofod_vox{res.points[i].x, res.points[i].y, res.points[i].z, res.points[i].range};
  return 0;
}
}
