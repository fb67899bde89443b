// =====================================================================================================
// vofod_oracle — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-scan volumetric hot
// path (ctu-mrs/vofod).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (libvofod_cuda) never links or calls it.
//
// PARITY PINS.  The reference ships no tests, golden vectors or fixtures and cannot be built whole here
// (its nodelet needs ROS / PCL / Eigen / FLANN, none installed).  What pins this restatement:
//   (1) PINNED against the reference's own code — VoxelMap (C1), VoxelGridWeighted (C2), VoxelGridCounted (C3) and the
//       raycast accumulate loop around forEachRay: oracle/_ref = the reference's voxel_map.cpp / voxel_grid_weighted.cpp /
//       voxel_grid_counted.cpp compiled where they lie against the stand-in Eigen/PCL/ROS headers of oracle/shim
//       (oracle/Makefile target `ref`); tests/golden/make_golden.py turns its outputs into tests/golden/ref_vectors.npz and
//       tests/test_golden.py requires this file to reproduce them bit for bit (20 arrays incl. the sequential-fp32 grid).
//   (2) PARITY UNPINNED for what lives in the nodelet (vofod_nodelet.cpp needs ROS to compile) and in third-party code:
//       the per-scan orchestration, point / ray update rules, classification, detections, sepclusters, and PCL's CropBox /
//       transformPointCloud / EuclideanClusterExtraction / MomentOfInertiaEstimation (restated from their published
//       behaviour: PCL 1.10, FLANN 1.9, Eigen 3.3 — versions unpinned in the reference, package.xml:164-165).  These are
//       held by the hand-derived known-answer tests of SURVEY.md Appendix B (tests/test_oracle_kat.py) only.
//
// Build: g++ -std=c++17 -O3 -DNDEBUG -ffp-contract=off  (= reference CMakeLists.txt:14-15, no -march, so
// every fp32 op below is separately rounded exactly as in the reference's x86-64 build).
// All file:line citations are relative to /root/reference.
// =====================================================================================================
#include "../include/vofod_cuda.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace vo
{
using idx3_t = std::tuple<int, int, int>;

struct vec3f { float x, y, z; };
struct vec3i { int x, y, z; };

// ---------------------------------------------------------------------------------------------------
// VoxelMap — include/vofod/voxel_map.h:12-142, src/voxel_map.cpp
// ---------------------------------------------------------------------------------------------------
class VoxelMap
{
public:
  float m_offset_x = 0, m_offset_y = 0, m_offset_z = 0;
  float m_voxel_size = 0, m_voxel_size_inv = 0, m_voxel_half = 0;
  int m_size_x = 0, m_size_y = 0, m_size_z = 0;
  std::vector<float> m_data;

  // voxel_map.cpp:11-19
  void resize(const vec3f& center, const vec3f& dims, const float voxel_size)
  {
    const float inv = 1.0f / voxel_size;
    const vec3f offset{center.x - dims.x / 2.0f, center.y - dims.y / 2.0f, center.z - dims.z / 2.0f};
    const vec3i sizes{int(std::ceil(inv * dims.x)) + 1, int(std::ceil(inv * dims.y)) + 1, int(std::ceil(inv * dims.z)) + 1};
    resize(offset, sizes, voxel_size);
  }
  // voxel_map.cpp:21-48
  void resize(const vec3f& offset, const vec3i& sizes, const float voxel_size)
  {
    m_voxel_half = voxel_size / 2.0f;
    m_voxel_size = voxel_size;
    m_voxel_size_inv = 1.0f / m_voxel_size;
    m_offset_x = offset.x; m_offset_y = offset.y; m_offset_z = offset.z;
    m_size_x = sizes.x; m_size_y = sizes.y; m_size_z = sizes.z;
    const size_t tot = size_t(m_size_x) * size_t(m_size_y) * size_t(m_size_z);
    m_data.resize(tot);
  }
  void resizeAs(const VoxelMap& o) { resize(vec3f{o.m_offset_x, o.m_offset_y, o.m_offset_z}, vec3i{o.m_size_x, o.m_size_y, o.m_size_z}, o.m_voxel_size); }
  size_t size() const { return m_data.size(); }

  // voxel_map.cpp:592-599
  idx3_t coordToIdx(const float x, const float y, const float z) const
  {
    const int ix = int(std::floor((x - m_offset_x) * m_voxel_size_inv));
    const int iy = int(std::floor((y - m_offset_y) * m_voxel_size_inv));
    const int iz = int(std::floor((z - m_offset_z) * m_voxel_size_inv));
    return {ix, iy, iz};
  }
  // voxel_map.cpp:607-613
  std::tuple<float, float, float> idxToCoord(const int ix, const int iy, const int iz) const
  {
    const float x = (ix + 0.5f) * m_voxel_size + m_offset_x;
    const float y = (iy + 0.5f) * m_voxel_size + m_offset_y;
    const float z = (iz + 0.5f) * m_voxel_size + m_offset_z;
    return {x, y, z};
  }
  // voxel_map.cpp:289-300
  bool inLimitsIdx(const int ix, const int iy, const int iz) const
  {
    return ix >= 0 && ix < m_size_x && iy >= 0 && iy < m_size_y && iz >= 0 && iz < m_size_z;
  }
  bool inLimits(const float x, const float y, const float z) const
  {
    const auto [ix, iy, iz] = coordToIdx(x, y, z);
    return inLimitsIdx(ix, iy, iz);
  }
  // voxel_map.cpp:105-133 (bounds-checked like std::vector::at)
  float& atIdx(const int ix, const int iy, const int iz) { return m_data.at(size_t(ix) + size_t(iy) * m_size_x + size_t(iz) * m_size_x * m_size_y); }
  float atIdx(const int ix, const int iy, const int iz) const { return m_data.at(size_t(ix) + size_t(iy) * m_size_x + size_t(iz) * m_size_x * m_size_y); }
  float& at(const float x, const float y, const float z)
  {
    const auto [ix, iy, iz] = coordToIdx(x, y, z);
    return atIdx(ix, iy, iz);
  }
  float& at(const idx3_t& i) { return atIdx(std::get<0>(i), std::get<1>(i), std::get<2>(i)); }
  float at(const idx3_t& i) const { return atIdx(std::get<0>(i), std::get<1>(i), std::get<2>(i)); }

  void setTo(const float v) { std::fill(m_data.begin(), m_data.end(), v); }  // :275-278
  void clear() { setTo(0); }                                                 // :268-271
  void copyDataIdx(const VoxelMap& from) { m_data.assign(from.m_data.begin(), from.m_data.end()); }  // :282-285

  // voxel_map.cpp:216-222
  uint64_t nVoxelsOver(const float threshold) const
  {
    uint64_t ret = 0;
    for (const auto val : m_data)
      ret += val > threshold;
    return ret;
  }

  // voxel_map.cpp:229-263 — Amanatides & Woo; every arithmetic op is a separate fp32 rounding
  void forEachRay(const vec3f& start, const vec3f& dir, const float length, const std::function<void(float, int, int, int)>& f) const
  {
    const float absdir[3] = {std::fabs(dir.x), std::fabs(dir.y), std::fabs(dir.z)};
    const float dirv[3] = {dir.x, dir.y, dir.z};
    int step[3];
    float tdelta[3];
    for (int a = 0; a < 3; a++)
    {
      step[a] = (dirv[a] > 0.0f) - (dirv[a] < 0.0f);      // cwiseSign().cast<int>()
      tdelta[a] = (1.0f / absdir[a]) * m_voxel_size;      // cwiseInverse()*voxel_size
    }
    const auto [cx, cy, cz] = coordToIdx(start.x, start.y, start.z);
    int cur[3] = {cx, cy, cz};
    const auto [ccx, ccy, ccz] = idxToCoord(cx, cy, cz);
    const float ctr[3] = {ccx - start.x, ccy - start.y, ccz - start.z};
    float tmax[3];
    for (int a = 0; a < 3; a++)
      tmax[a] = (m_voxel_half + float(step[a]) * ctr[a]) / absdir[a];
    const int sizes[3] = {m_size_x, m_size_y, m_size_z};
    float last[3];  // stored as float in the reference (vec3_t last_voxel), compared == with int
    for (int a = 0; a < 3; a++)
      last[a] = step[a] > 0 ? float(sizes[a] - 1) : 0.0f;

    float prev_dist = 0.0f;
    while (prev_dist < length)
    {
      int i = 0;  // minCoeff(&i): first minimum, strict '<'
      float dist = tmax[0];
      if (tmax[1] < dist) { dist = tmax[1]; i = 1; }
      if (tmax[2] < dist) { dist = tmax[2]; i = 2; }
      const float ddist = std::min(dist, length) - prev_dist;
      f(ddist, cur[0], cur[1], cur[2]);
      prev_dist = dist;
      if (float(cur[i]) == last[i])
        break;
      cur[i] += step[i];
      tmax[i] += tdelta[i];
    }
  }

  // voxel_map.cpp:518-534 — x-outer / z-inner visiting order, std::function per cell, vector::at
  void forEachIdx(const std::function<void(float&, int, int, int)>& f, const int offset = 0)
  {
    const int max_x = m_size_x - offset, max_y = m_size_y - offset, max_z = m_size_z - offset;
    for (int x = offset; x < max_x; x++)
      for (int y = offset; y < max_y; y++)
        for (int z = offset; z < max_z; z++)
        {
          float& mapval = m_data.at(size_t(x) + size_t(y) * m_size_x + size_t(z) * m_size_x * m_size_y);
          f(mapval, x, y, z);
        }
  }

  // voxel_map.cpp:376-400
  bool hasCloseTo(const float x, const float y, const float z, const float max_dist, const float threshold) const
  {
    const auto [ox, oy, oz] = coordToIdx(x, y, z);
    const float max_dist_idx = max_dist * m_voxel_size_inv;
    const int mv = int(std::ceil(max_dist_idx));
    const int bx = std::max(ox - mv, 0), by = std::max(oy - mv, 0), bz = std::max(oz - mv, 0);
    const int ex = std::min(ox + mv, m_size_x), ey = std::min(oy + mv, m_size_y), ez = std::min(oz + mv, m_size_z);
    for (int xi = bx; xi < ex; xi++)
      for (int yi = by; yi < ey; yi++)
        for (int zi = bz; zi < ez; zi++)
        {
          // Eigen int-vector .norm(): int(sqrt(squaredNorm)) — truncation
          const int dx = xi - ox, dy = yi - oy, dz = zi - oz;
          const int nrm = int(std::sqrt(double(dx * dx + dy * dy + dz * dz)));
          if (atIdx(xi, yi, zi) > threshold && float(nrm) <= max_dist_idx)
            return true;
        }
    return false;
  }

  static int manhattan(const idx3_t& a, const idx3_t& b)
  {
    return std::abs(std::get<0>(a) - std::get<0>(b)) + std::abs(std::get<1>(a) - std::get<1>(b)) + std::abs(std::get<2>(a) - std::get<2>(b));
  }

  // voxel_map.cpp:402-488
  std::tuple<bool, std::vector<idx3_t>> exploreToGround(const float x, const float y, const float z, const float unknown_threshold,
                                                        const float ground_threshold, const float max_voxel_dist) const
  {
    const std::tuple<bool, std::vector<idx3_t>> connected_ret = {true, {}};
    const idx3_t orig = coordToIdx(x, y, z);
    const auto [xi, yi, zi] = orig;
    if (xi <= 0 || yi <= 0 || zi <= 0)
      return connected_ret;
    if (xi >= m_size_x - 1 || yi >= m_size_y - 1 || zi >= m_size_z - 1)
      return connected_ret;

    auto pack = [](const idx3_t& i) { return (uint64_t(uint32_t(std::get<0>(i))) << 42) ^ (uint64_t(uint32_t(std::get<1>(i))) << 21) ^ uint64_t(uint32_t(std::get<2>(i))); };
    std::unordered_set<uint64_t> explored;
    std::vector<idx3_t> explored_unknown;
    std::vector<idx3_t> to_explore;
    to_explore.push_back(orig);
    while (!to_explore.empty())
    {
      const idx3_t cur = to_explore.back();
      to_explore.pop_back();  // DFS
      const float cur_val = at(cur);
      if (cur_val > ground_threshold)
        return connected_ret;
      if (cur_val > unknown_threshold)
      {
        explored_unknown.push_back(cur);
        if (float(manhattan(orig, cur)) == max_voxel_dist - 1)
          return connected_ret;
        const int c[3] = {std::get<0>(cur), std::get<1>(cur), std::get<2>(cur)};
        const int sizes[3] = {m_size_x, m_size_y, m_size_z};
        for (int a = 0; a < 3; a++)  // +x, +y, +z
          if (c[a] < sizes[a] - 1)
          {
            int n[3] = {c[0], c[1], c[2]};
            n[a] += 1;
            const idx3_t to_add{n[0], n[1], n[2]};
            if (explored.count(pack(to_add)) == 0 && float(manhattan(orig, to_add)) <= max_voxel_dist)
              to_explore.push_back(to_add);
          }
        for (int a = 0; a < 3; a++)  // -x, -y, -z
          if (c[a] > 0)
          {
            int n[3] = {c[0], c[1], c[2]};
            n[a] -= 1;
            const idx3_t to_add{n[0], n[1], n[2]};
            if (explored.count(pack(to_add)) == 0 && float(manhattan(orig, to_add)) <= max_voxel_dist)
              to_explore.push_back(to_add);
          }
      }
      explored.insert(pack(cur));
    }
    return {false, explored_unknown};
  }

  // voxel_map.cpp:497-516
  bool isFloatingIdx(const int xi, const int yi, const int zi, const float threshold) const
  {
    if (xi <= 0 || yi <= 0 || zi <= 0)
      return false;
    if (xi >= m_size_x - 1 || yi >= m_size_y - 1 || zi >= m_size_z - 1)
      return false;
    for (int x = xi - 1; x <= xi + 1; x++)
      for (int y = yi - 1; y <= yi + 1; y++)
        for (int z = zi - 1; z <= zi + 1; z++)
          if (atIdx(x, y, z) > threshold)
            return false;
    return true;
  }

  // voxel_map.cpp:547-584
  VoxelMap getSubmapCopy(const vec3f& min_pt, const vec3f& max_pt, const int inflate) const
  {
    auto [nx, ny, nz] = coordToIdx(min_pt.x, min_pt.y, min_pt.z);
    auto [xx, xy, xz] = coordToIdx(max_pt.x, max_pt.y, max_pt.z);
    nx = std::clamp(nx - inflate, 0, m_size_x - 1);
    ny = std::clamp(ny - inflate, 0, m_size_y - 1);
    nz = std::clamp(nz - inflate, 0, m_size_z - 1);
    xx = std::clamp(xx + inflate, 0, m_size_x - 1);
    xy = std::clamp(xy + inflate, 0, m_size_y - 1);
    xz = std::clamp(xz + inflate, 0, m_size_z - 1);
    const auto [cx, cy, cz] = idxToCoord(nx, ny, nz);
    const vec3f sub_off{cx - m_voxel_size / 2.0f, cy - m_voxel_size / 2.0f, cz - m_voxel_size / 2.0f};
    const vec3i sub_size{xx - nx + 1, xy - ny + 1, xz - nz + 1};
    VoxelMap ret;
    ret.resize(sub_off, sub_size, m_voxel_size);
    for (int x = 0; x < sub_size.x; x++)
      for (int y = 0; y < sub_size.y; y++)
        for (int z = 0; z < sub_size.z; z++)
          ret.atIdx(x, y, z) = atIdx(x + nx, y + ny, z + nz);
    return ret;
  }

  // voxel_map.cpp:157-212 — emission order x-outer, y, z-inner
  std::vector<vofod_xyzi> voxelsAsPC(const float threshold, const bool greater_than, const bool metric) const
  {
    std::vector<vofod_xyzi> cloud;
    cloud.reserve(size_t(m_size_x) * m_size_y * m_size_z / 10);
    for (int x = 0; x < m_size_x; x++)
      for (int y = 0; y < m_size_y; y++)
        for (int z = 0; z < m_size_z; z++)
        {
          const float mapval = m_data.at(size_t(x) + size_t(y) * m_size_x + size_t(z) * m_size_x * m_size_y);
          if ((mapval > threshold) == greater_than)
          {
            vofod_xyzi pt;
            if (metric)
            {
              const auto [cx, cy, cz] = idxToCoord(x, y, z);
              pt.x = cx; pt.y = cy; pt.z = cz;
            } else
            {
              pt.x = float(x); pt.y = float(y); pt.z = float(z);
            }
            pt.intensity = mapval;
            cloud.push_back(pt);
          }
        }
    return cloud;
  }
};

// ---------------------------------------------------------------------------------------------------
// Voxel grids — src/voxel_grid_weighted.cpp:41-190, src/voxel_grid_counted.cpp:49-196
// (pcl::VoxelGrid base: leaf_size_, inverse_leaf_size_ = 1/leaf, min_points_per_voxel_ = 0; getMinMax3D)
// ---------------------------------------------------------------------------------------------------
struct cpidx
{
  unsigned idx;
  unsigned cloud_point_index;
  int ijk0, ijk1, ijk2;
  bool operator<(const cpidx& p) const { return idx < p.idx; }
};

struct vg_layout
{
  float offset[3];
  float inv_leaf;
  float leaf;
  int min_b[3], max_b[3], div_b[3];
};

// returns <0 on overflow (voxel_grid_weighted.cpp:61-69)
template <class PT>
static int vg_common(const std::vector<PT>& in, const float leaf, const bool align_voxels, const float align_center[3], vg_layout& L,
                     std::vector<cpidx>& index_vector, std::vector<std::pair<unsigned, unsigned>>& runs)
{
  index_vector.clear();
  runs.clear();
  if (in.empty())
    return 0;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Ones()/leaf_size_
  L.leaf = leaf;
  L.inv_leaf = inv;
  float min_p[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float max_p[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  for (const auto& p : in)  // pcl::getMinMax3D
  {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))
      continue;
    min_p[0] = std::min(min_p[0], p.x); min_p[1] = std::min(min_p[1], p.y); min_p[2] = std::min(min_p[2], p.z);
    max_p[0] = std::max(max_p[0], p.x); max_p[1] = std::max(max_p[1], p.y); max_p[2] = std::max(max_p[2], p.z);
  }
  const int64_t dx = int64_t((max_p[0] - min_p[0]) * inv) + 2;
  const int64_t dy = int64_t((max_p[1] - min_p[1]) * inv) + 2;
  const int64_t dz = int64_t((max_p[2] - min_p[2]) * inv) + 2;
  if (dx * dy * dz > int64_t(std::numeric_limits<int32_t>::max()))
    return -1;
  for (int a = 0; a < 3; a++)
  {
    L.min_b[a] = int(std::floor(min_p[a] * inv));
    L.max_b[a] = int(std::floor(max_p[a] * inv));
    L.offset[a] = float(L.min_b[a]) * leaf;
  }
  if (align_voxels)
  {
    for (int a = 0; a < 3; a++)
    {
      float aco = std::fmod(align_center[a] - leaf / 2, leaf);
      if (aco < 0)
        aco += leaf;
      L.offset[a] -= aco;
      L.min_b[a] = int(std::floor(L.offset[a] * inv));
    }
  }
  for (int a = 0; a < 3; a++)
    L.div_b[a] = L.max_b[a] - L.min_b[a] + 1;
  const int mul0 = 1, mul1 = L.div_b[0], mul2 = L.div_b[0] * L.div_b[1];

  index_vector.reserve(in.size());
  for (unsigned it = 0; it < in.size(); it++)
  {
    const auto& p = in[it];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))
      continue;
    const int ijk0 = int(std::floor((p.x - L.offset[0]) * inv));
    const int ijk1 = int(std::floor((p.y - L.offset[1]) * inv));
    const int ijk2 = int(std::floor((p.z - L.offset[2]) * inv));
    const int idx = ijk0 * mul0 + ijk1 * mul1 + ijk2 * mul2;
    index_vector.push_back({unsigned(idx), it, ijk0, ijk1, ijk2});
  }
  std::sort(index_vector.begin(), index_vector.end());
  unsigned index = 0;
  runs.reserve(index_vector.size());
  while (index < index_vector.size())
  {
    unsigned i = index + 1;
    while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx)
      ++i;
    runs.emplace_back(index, i);  // min_points_per_voxel_ == 0
    index = i;
  }
  return 0;
}

template <class PT>
static int voxel_grid_weighted(const std::vector<PT>& in, const float leaf, const bool align, const float align_center[3], std::vector<vofod_vox>& out)
{
  vg_layout L;
  std::vector<cpidx> iv;
  std::vector<std::pair<unsigned, unsigned>> runs;
  out.clear();
  if (vg_common(in, leaf, align, align_center, L, iv, runs) < 0)
    return -1;
  out.resize(runs.size());
  for (size_t cp = 0; cp < runs.size(); cp++)
  {
    const auto& c = iv[runs[cp].first];
    out[cp].x = (float(c.ijk0) + 0.5f) * L.leaf + L.offset[0];
    out[cp].y = (float(c.ijk1) + 0.5f) * L.leaf + L.offset[1];
    out[cp].z = (float(c.ijk2) + 0.5f) * L.leaf + L.offset[2];
    out[cp].count = uint32_t(runs[cp].second - runs[cp].first);
  }
  return 0;
}

static int voxel_grid_counted(const std::vector<vofod_xyzi>& in, const float leaf, const float threshold, const bool align, const float align_center[3],
                              std::vector<vofod_vox>& out)
{
  vg_layout L;
  std::vector<cpidx> iv;
  std::vector<std::pair<unsigned, unsigned>> runs;
  out.clear();
  if (vg_common(in, leaf, align, align_center, L, iv, runs) < 0)
    return -1;
  out.resize(runs.size());
  for (size_t cp = 0; cp < runs.size(); cp++)
  {
    const auto& c = iv[runs[cp].first];
    out[cp].x = (float(c.ijk0) + 0.5f) * L.leaf + L.offset[0];
    out[cp].y = (float(c.ijk1) + 0.5f) * L.leaf + L.offset[1];
    out[cp].z = (float(c.ijk2) + 0.5f) * L.leaf + L.offset[2];
    // voxel_grid_counted.cpp:185-187 — positions in the SORTED run list applied to the UNSORTED input
    unsigned count = 0;
    for (unsigned k = runs[cp].first; k < runs[cp].second; k++)
      count += in[k].intensity > threshold;
    out[cp].count = count;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// pcl::EuclideanClusterExtraction restated (vofod_nodelet.cpp:689-698): connected components of
// d^2 < r^2 (strict; fp32 d^2 summed x,y,z as FLANN L2_Simple; r^2 = float(double(tol)^2)), seeds in
// ascending index, per-cluster indices ascending, clusters by size descending (ties: min index first —
// the reference's std::sort leaves ties unspecified).  Neighbour search by a uniform cell list instead of
// a kd-tree: same result set.
// ---------------------------------------------------------------------------------------------------
struct clusters_t
{
  std::vector<int32_t> labels;                 // min index of the component
  std::vector<std::vector<int>> clusters;      // sorted as described
};

static inline uint64_t cell_key(int64_t cx, int64_t cy, int64_t cz)
{
  return (uint64_t(cx + (1 << 20)) & 0x1FFFFF) | ((uint64_t(cy + (1 << 20)) & 0x1FFFFF) << 21) | ((uint64_t(cz + (1 << 20)) & 0x1FFFFF) << 42);
}

// std_sort_ties: order the clusters with PCL's own call, std::sort through reverse iterators by size — not stable: clusters of equal size end up
// in whatever order libstdc++'s introsort leaves (what the reference build produces).  Default: ties by the smallest point index, which is what
// the GPU implements; the switch exists to show that the tie order is the ONLY thing that separates the two (tests/test_ref_nodelet.py).
static clusters_t euclidean_clusters(const float* xyz, const size_t stride, const size_t m, const float tol, const bool std_sort_ties = false)
{
  clusters_t ret;
  ret.labels.assign(m, -1);
  const float r2 = float(double(tol) * double(tol));
  const double cell = double(tol) * (1.0 + 1e-6);
  const double inv_cell = tol > 0 ? 1.0 / cell : 0.0;
  std::unordered_map<uint64_t, std::vector<int>> cells;
  cells.reserve(m);
  auto P = [&](size_t i, int a) { return xyz[i * stride + a]; };
  auto cc = [&](size_t i, int a) { return int64_t(std::floor(double(P(i, a)) * inv_cell)); };
  if (tol > 0)
    for (size_t i = 0; i < m; i++)
      cells[cell_key(cc(i, 0), cc(i, 1), cc(i, 2))].push_back(int(i));
  std::vector<char> processed(m, 0);
  std::vector<int> queue;
  for (size_t i = 0; i < m; i++)
  {
    if (processed[i])
      continue;
    queue.clear();
    queue.push_back(int(i));
    processed[i] = 1;
    for (size_t q = 0; q < queue.size(); q++)
    {
      const int a = queue[q];
      if (!(tol > 0))
        continue;
      const int64_t cx = cc(a, 0), cy = cc(a, 1), cz = cc(a, 2);
      for (int64_t dz = -1; dz <= 1; dz++)
        for (int64_t dy = -1; dy <= 1; dy++)
          for (int64_t dx = -1; dx <= 1; dx++)
          {
            const auto it = cells.find(cell_key(cx + dx, cy + dy, cz + dz));
            if (it == cells.end())
              continue;
            for (const int b : it->second)
            {
              if (processed[b])
                continue;
              float d2 = 0.0f;
              float diff = P(a, 0) - P(b, 0); d2 += diff * diff;
              diff = P(a, 1) - P(b, 1); d2 += diff * diff;
              diff = P(a, 2) - P(b, 2); d2 += diff * diff;
              if (d2 < r2)
              {
                processed[b] = 1;
                queue.push_back(b);
              }
            }
          }
    }
    std::vector<int> idcs(queue.begin(), queue.end());
    std::sort(idcs.begin(), idcs.end());
    for (const int k : idcs)
      ret.labels[k] = idcs.front();
    ret.clusters.push_back(std::move(idcs));
  }
  if (std_sort_ties)
    std::sort(ret.clusters.rbegin(), ret.clusters.rend(), [](const auto& a, const auto& b) { return a.size() < b.size(); });  // extract_clusters.hpp, 1.10
  else
    std::stable_sort(ret.clusters.begin(), ret.clusters.end(), [](const auto& a, const auto& b) {
      if (a.size() != b.size())
        return a.size() > b.size();
      return a.front() < b.front();
    });
  return ret;
}

// ---------------------------------------------------------------------------------------------------
// pcl::MomentOfInertiaEstimation restated (vofod_nodelet.cpp:1655-1672): fp32 mean / covariance in point
// order, eigenvectors, OBB.  The eigen-solve is a fixed-sweep cyclic Jacobi in fp64 (Eigen::EigenSolver's
// Hessenberg+QR iteration is not restated: source absent) — the OBB is ill-defined anyway when eigenvalues
// tie; `eig_gap` lets the tests skip obb comparisons in that case.
// ---------------------------------------------------------------------------------------------------
static void jacobi3(const double Ain[3][3], double eval[3], double V[3][3])
{
  double A[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      A[i][j] = Ain[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 16; sweep++)
  {
    const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
    if (off == 0.0)
      break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++)
      {
        if (A[p][q] == 0.0)
          continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0);
        const double s = t * c;
        const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
        A[p][p] = app - t * apq;
        A[q][q] = aqq + t * apq;
        A[p][q] = A[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; k++)
        {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  eval[0] = A[0][0]; eval[1] = A[1][1]; eval[2] = A[2][2];
}

static void moment_of_inertia(const std::vector<vofod_vox>& cloud, const std::vector<int>& idcs, vofod_cluster_info& ci)
{
  const size_t n = idcs.size();
  float mean[3] = {0, 0, 0};
  float amin[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float amax[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  for (const int i : idcs)  // computeMeanValue
  {
    const float p[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    for (int a = 0; a < 3; a++)
    {
      mean[a] += p[a];
      if (p[a] <= amin[a]) amin[a] = p[a];
      if (p[a] >= amax[a]) amax[a] = p[a];
    }
  }
  const unsigned np = n == 0 ? 1 : unsigned(n);
  for (int a = 0; a < 3; a++)
    mean[a] /= float(np);
  float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (const int i : idcs)  // computeCovarianceMatrix
  {
    const float d[3] = {cloud[i].x - mean[0], cloud[i].y - mean[1], cloud[i].z - mean[2]};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        cov[r][c] += d[r] * d[c];
  }
  const float factor = 1.0f / float((long(n) - 1 > 0) ? (n - 1) : 1);
  double A[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
    {
      cov[r][c] *= factor;
      A[r][c] = double(cov[r][c]);
    }
  double evald[3], Vd[3][3];
  jacobi3(A, evald, Vd);
  const float ev[3] = {float(evald[0]), float(evald[1]), float(evald[2])};
  // computeEigenVectors: index shuffling exactly as PCL
  unsigned major = 0, middle = 1, minor = 2;
  if (ev[major] < ev[middle]) std::swap(major, middle);
  if (ev[major] < ev[minor]) std::swap(major, minor);
  if (ev[middle] < ev[minor]) std::swap(minor, middle);
  float ax[3][3];  // ax[k] = axis k (major, middle, minor)
  const unsigned order[3] = {major, middle, minor};
  for (int k = 0; k < 3; k++)
  {
    float v[3] = {float(Vd[0][order[k]]), float(Vd[1][order[k]]), float(Vd[2][order[k]])};
    const float nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int a = 0; a < 3; a++)
      ax[k][a] = v[a] / nrm;
  }
  // det = major . (middle x minor)
  const float cx = ax[1][1] * ax[2][2] - ax[1][2] * ax[2][1];
  const float cy = ax[1][2] * ax[2][0] - ax[1][0] * ax[2][2];
  const float cz = ax[1][0] * ax[2][1] - ax[1][1] * ax[2][0];
  const float det = ax[0][0] * cx + ax[0][1] * cy + ax[0][2] * cz;
  if (det <= 0.0f)
    for (int a = 0; a < 3; a++)
      ax[0][a] = -ax[0][a];
  // computeOBB
  float omin[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float omax[3] = {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
  for (const int i : idcs)
  {
    const float d[3] = {cloud[i].x - mean[0], cloud[i].y - mean[1], cloud[i].z - mean[2]};
    for (int k = 0; k < 3; k++)
    {
      const float v = d[0] * ax[k][0] + d[1] * ax[k][1] + d[2] * ax[k][2];
      if (v <= omin[k]) omin[k] = v;
      if (v >= omax[k]) omax[k] = v;
    }
  }
  float shift[3];
  for (int k = 0; k < 3; k++)
  {
    shift[k] = (omax[k] + omin[k]) / 2.0f;
    omin[k] -= shift[k];
    omax[k] -= shift[k];
  }
  for (int a = 0; a < 3; a++)
  {
    ci.aabb_min[a] = amin[a];
    ci.aabb_max[a] = amax[a];
    ci.obb_min[a] = omin[a];
    ci.obb_max[a] = omax[a];
    // position = mean + Rot*shift, Rot columns = axes; row a: ax[0][a]*s0 + ax[1][a]*s1 + ax[2][a]*s2 (Eigen 3-term redux: e0 + (e1 + e2))
    ci.obb_center[a] = mean[a] + (ax[0][a] * shift[0] + (ax[1][a] * shift[1] + ax[2][a] * shift[2]));
    for (int k = 0; k < 3; k++)
      ci.obb_rot[a * 3 + k] = ax[k][a];
  }
  // parity aid: smallest relative gap between sorted eigenvalues
  double s[3] = {evald[0], evald[1], evald[2]};
  std::sort(s, s + 3);
  const double scale = std::max(std::fabs(s[2]), 1e-30);
  ci.eig_gap = float(std::min(s[1] - s[0], s[2] - s[1]) / scale);
}

// ---------------------------------------------------------------------------------------------------
// The nodelet state + per-scan functions (vofod_nodelet.cpp)
// ---------------------------------------------------------------------------------------------------
static constexpr float vflags_unmarked = 0.0f, vflags_point = 2.0f, vflags_unknown = 3.0f;  // :2335-2337

struct stage_clock
{
  std::chrono::steady_clock::time_point t0;
  void start() { t0 = std::chrono::steady_clock::now(); }
  double lap()
  {
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
    return ms;
  }
};

struct Oracle
{
  VoxelMap voxel_map, voxel_flags, voxel_raycast, bg_local;
  // deterministic twin of the raycast accumulator: per-cell callback count and exact fixed-point length
  std::vector<uint32_t> ray_count;
  std::vector<int64_t> ray_fixed;
  int frac_bits = 26;
  bool track_counts = true;       // maintain ray_count / ray_fixed (off for pure CPU-baseline timing)
  bool apply_from_fixed = false;  // apply uses float(fixed) instead of the sequential fp32 sum
  bool std_sort_ties = false;     // cluster order among equal sizes as libstdc++'s std::sort leaves it (see euclidean_clusters)

  int W = 0, H = 0;
  std::vector<float> lut_dirs, lut_offs;  // 3 x N column-major
  std::vector<uint8_t> mask;

  bool background_pts_sufficient = false, sure_background_sufficient = false;
  uint32_t last_detection_id = 0;
  int detection_its = 0;
  float voxel_size = 0;

  // outputs of the last scan
  std::vector<vofod_vox> cloud_weighted;
  clusters_t clusters;
  std::vector<uint8_t> point_close;
  std::vector<std::vector<int>> close_clusters, far_clusters;
  std::vector<vofod_cluster_info> cluster_infos;
  std::vector<vofod_detection> detections;
  uint64_t n_bg = 0, n_traversals = 0;
  uint32_t n_filtered = 0;
  double stage_ms[VOFOD_N_STAGES] = {0};

  // ---- reset(): vofod_nodelet.cpp:1610-1632
  void reset(const vofod_params& p, const float vs)
  {
    voxel_size = vs;
    const float oz = p.oparea_offset[2] + p.oparea_size[2] / 2.0f;  // :212
    voxel_map.resize(vec3f{p.oparea_offset[0], p.oparea_offset[1], oz}, vec3f{p.oparea_size[0], p.oparea_size[1], p.oparea_size[2]}, vs);
    voxel_map.setTo(p.score_init);
    voxel_flags.resizeAs(voxel_map);
    voxel_flags.clear();
    voxel_raycast.resizeAs(voxel_map);
    voxel_raycast.clear();
    bg_local.resizeAs(voxel_map);
    on_resize();
    detection_its = 0;
    raycast_pending = false;
    background_pts_sufficient = sure_background_sufficient = false;
    last_detection_id = 0;
  }
  void on_resize()
  {
    ray_count.assign(voxel_map.size(), 0);
    ray_fixed.assign(voxel_map.size(), 0);
  }

  // ---- rangefinder seed: vofod_nodelet.cpp:581-613
  void range_update(const float pt[3], const vofod_params& p)
  {
    if (!voxel_map.inLimits(pt[0], pt[1], pt[2]))
      return;
    float& mapval = voxel_map.at(pt[0], pt[1], pt[2]);
    mapval = float((double(mapval) + p.score_point) / 2.0);
  }

  // ---- filterAndTransform: vofod_nodelet.cpp:621-684
  struct pt3 { float x, y, z; };
  int filter_and_transform(const vofod_pt* scan, const size_t n, const vofod_pose& tf, const vofod_params& p)
  {
    std::vector<pt3> filtered;
    filtered.reserve(n);
    {
      const float ez = p.exclude_box_offset[2] + p.exclude_box_size[2] / 2.0f;  // :204
      const float mx[3] = {p.exclude_box_offset[0] + p.exclude_box_size[0] / 2, p.exclude_box_offset[1] + p.exclude_box_size[1] / 2, ez + p.exclude_box_size[2] / 2};
      const float mn[3] = {p.exclude_box_offset[0] - p.exclude_box_size[0] / 2, p.exclude_box_offset[1] - p.exclude_box_size[1] / 2, ez - p.exclude_box_size[2] / 2};
      for (size_t i = 0; i < n; i++)  // pcl::CropBox, negative
      {
        const auto& q = scan[i];
        if (!std::isfinite(q.x) || !std::isfinite(q.y) || !std::isfinite(q.z))
          continue;
        const bool outside = (q.x < mn[0] || q.y < mn[1] || q.z < mn[2]) || (q.x > mx[0] || q.y > mx[1] || q.z > mx[2]);
        if (outside)
          filtered.push_back({q.x, q.y, q.z});
      }
    }
    // pcl::transformPointCloud(Affine3f), PCL 1.10 SSE order: x*c0 + (y*c1 + (z*c2 + c3))
    for (auto& q : filtered)
    {
      const float x = q.x, y = q.y, z = q.z;
      q.x = x * tf.R[0] + (y * tf.R[1] + (z * tf.R[2] + tf.t[0]));
      q.y = x * tf.R[3] + (y * tf.R[4] + (z * tf.R[5] + tf.t[1]));
      q.z = x * tf.R[6] + (y * tf.R[7] + (z * tf.R[8] + tf.t[2]));
    }
    {
      const float oz = p.oparea_offset[2] + p.oparea_size[2] / 2.0f;  // :212
      const float mx[3] = {p.oparea_offset[0] + p.oparea_size[0] / 2, p.oparea_offset[1] + p.oparea_size[1] / 2, oz + p.oparea_size[2] / 2};
      const float mn[3] = {p.oparea_offset[0] - p.oparea_size[0] / 2, p.oparea_offset[1] - p.oparea_size[1] / 2, oz - p.oparea_size[2] / 2};
      std::vector<pt3> kept;
      kept.reserve(filtered.size());
      for (const auto& q : filtered)  // pcl::CropBox, positive
      {
        if (!std::isfinite(q.x) || !std::isfinite(q.y) || !std::isfinite(q.z))
          continue;
        const bool outside = (q.x < mn[0] || q.y < mn[1] || q.z < mn[2]) || (q.x > mx[0] || q.y > mx[1] || q.z > mx[2]);
        if (!outside)
          kept.push_back(q);
      }
      filtered.swap(kept);
    }
    n_filtered = uint32_t(filtered.size());
    const auto [ax, ay, az] = voxel_map.idxToCoord(0, 0, 0);  // :664
    const float align[3] = {ax, ay, az};
    return voxel_grid_weighted(filtered, voxel_size, true, align, cloud_weighted);
  }

  // ---- findCloseFarClusters: vofod_nodelet.cpp:703-750
  void find_close_far(const vofod_params& p)
  {
    const float max_dist = float(p.ground_points_max_distance);
    const float thr = float(p.thr_new_obstacles);
    n_bg = voxel_map.nVoxelsOver(thr);
    const float n_voxels_xy = p.oparea_size[0] / voxel_size * p.oparea_size[1] / voxel_size;  // :229
    const uint64_t min_sufficient = uint64_t(n_voxels_xy * p.background_sufficient_points_ratio);  // :230
    if (n_bg > min_sufficient)
      background_pts_sufficient = true;
    close_clusters.clear();
    far_clusters.clear();
    point_close.assign(cloud_weighted.size(), 0);
    for (const auto& cl : clusters.clusters)
    {
      bool is_close = false;
      for (const int idx : cl)
      {
        const auto& pt = cloud_weighted[idx];
        if (voxel_map.hasCloseTo(pt.x, pt.y, pt.z, max_dist, thr))
        {
          is_close = true;
          break;
        }
      }
      if (is_close)
      {
        close_clusters.push_back(cl);
        for (const int idx : cl)
          point_close[idx] = 1;
      } else
        far_clusters.push_back(cl);
    }
  }

  // ---- updateVoxel / updateVMaps: vofod_nodelet.cpp:777-809
  void update_voxel(const vofod_vox& pt, const float vmap_score, const float vflags)
  {
    const auto [xc, yc, zc] = voxel_map.coordToIdx(pt.x, pt.y, pt.z);
    if (!voxel_map.inLimitsIdx(xc, yc, zc))
      return;  // the reference would throw std::out_of_range / corrupt; never happens after the op-area crop
    float& mapval = voxel_map.atIdx(xc, yc, zc);
    const float w = 1.0f / float(1lu << std::clamp(pt.count, 0u, 63u));
    mapval = w * mapval + (1.0f - w) * vmap_score;
    voxel_flags.atIdx(xc, yc, zc) = vflags;
  }
  void update_vmaps(const std::vector<std::vector<int>>& cls, const float score, const float flag)
  {
    for (const auto& cl : cls)
      for (const int idx : cl)
        update_voxel(cloud_weighted[idx], score, flag);
  }

  // ---- raycast_cloud accumulate: vofod_nodelet.cpp:1425-1492
  int raycast_accumulate(const vofod_pt* scan, const size_t n, const vofod_pose& tf, const vofod_params& p)
  {
    n_traversals = 0;
    if (p.raycast_pause)
      return VOFOD_W_PAUSED;
    if (n != size_t(W) * size_t(H))
      return VOFOD_E_DIMS;
    voxel_raycast.clear();  // :1430
    if (track_counts)
    {
      std::fill(ray_count.begin(), ray_count.end(), 0u);
      std::fill(ray_fixed.begin(), ray_fixed.end(), int64_t(0));
    }
    if (!voxel_raycast.inLimits(tf.t[0], tf.t[1], tf.t[2]))
      return VOFOD_W_SENSOR_OOB;
    const float max_dist = float(p.raycast_max_distance);
    const float min_intensity = float(p.raycast_min_intensity);
    const float vs = voxel_size;
    const double scale = std::ldexp(1.0, frac_bits);
    const size_t sx = voxel_raycast.m_size_x, sxy = size_t(voxel_raycast.m_size_x) * voxel_raycast.m_size_y;
    uint64_t trav = 0;
    for (int row = 0; row < H; row++)
      for (int col = 0; col < W; col++)
      {
        const unsigned idx = unsigned(row) * unsigned(W) + unsigned(col);
        const auto& pt = scan[idx];
        if (pt.intensity < min_intensity || (!mask[idx] && pt.range_mm == 0))
          continue;
        const float* d1 = &lut_dirs[size_t(idx) * 3];
        const float* o1 = &lut_offs[size_t(idx) * 3];
        // Eigen 3x3 * 3 coefficient product: r_i = R_i0*v0 + (R_i1*v1 + R_i2*v2)
        const vec3f dir{tf.R[0] * d1[0] + (tf.R[1] * d1[1] + tf.R[2] * d1[2]), tf.R[3] * d1[0] + (tf.R[4] * d1[1] + tf.R[5] * d1[2]),
                        tf.R[6] * d1[0] + (tf.R[7] * d1[1] + tf.R[8] * d1[2])};
        const float ray_dist = 0.001f * float(pt.range_mm);
        const float dist = ray_dist == 0.0f ? max_dist : std::min(ray_dist - vs, max_dist);
        const vec3f start{(tf.R[0] * o1[0] + (tf.R[1] * o1[1] + tf.R[2] * o1[2])) + tf.t[0], (tf.R[3] * o1[0] + (tf.R[4] * o1[1] + tf.R[5] * o1[2])) + tf.t[1],
                          (tf.R[6] * o1[0] + (tf.R[7] * o1[1] + tf.R[8] * o1[2])) + tf.t[2]};
        if (voxel_raycast.inLimits(start.x, start.y, start.z))
        {
          voxel_raycast.forEachRay(start, dir, dist, [&](const float val, const int xi, const int yi, const int zi) {
            float& raycastval = voxel_raycast.atIdx(xi, yi, zi);
            raycastval += val;
            trav++;
            if (track_counts)
            {
              const size_t li = size_t(xi) + size_t(yi) * sx + size_t(zi) * sxy;
              ray_count[li] += 1;
              ray_fixed[li] += int64_t(std::llrint(double(val) * scale));
            }
          });
        }
      }
    n_traversals = trav;
    return VOFOD_OK;
  }

  // ---- raycast_cloud apply: vofod_nodelet.cpp:1539-1602
  int raycast_apply(const int its_diff, const vofod_params& p)
  {
    if (p.raycast_pause)
      return VOFOD_W_PAUSED;
    const float detection_its_diff = float(its_diff);
    if (apply_from_fixed)
    {
      const double inv_scale = std::ldexp(1.0, -frac_bits);
      for (size_t i = 0; i < voxel_raycast.m_data.size(); i++)
        voxel_raycast.m_data[i] = float(double(ray_fixed[i]) * inv_scale);
    }
    const float max_val = *std::max_element(voxel_raycast.m_data.begin(), voxel_raycast.m_data.end());
    if (max_val == 0.0f)
      return VOFOD_W_EMPTY_RAYCAST;
    const float ray_update_score = float(p.score_ray);
    const float ray_update_weight = float(p.raycast_weight_coefficient);
    if (p.raycast_new_update_rule)
    {
      const float voxel_diag = float(std::sqrt(3.0) * double(voxel_size));  // std::sqrt(3) is double
      const float weighting_factor = ray_update_weight / voxel_diag;
      voxel_flags.forEachIdx([&](float& flag, const int xc, const int yc, const int zc) {
        float raycastval;
        if (flag == vflags_unmarked && (raycastval = voxel_raycast.atIdx(xc, yc, zc)) > 0.0f)
        {
          float& mapval = voxel_map.atIdx(xc, yc, zc);
          const float n_int = weighting_factor * raycastval;
          const float w1 = float(std::pow(2.0, double(-detection_its_diff * n_int)));  // std::pow(int, float) -> double
          const float w2 = 1.0f - w1;
          mapval = w1 * mapval + w2 * ray_update_score;
        }
      });
    } else
    {
      voxel_flags.forEachIdx([&](float& flag, const int xc, const int yc, const int zc) {
        float raycastval;
        float& mapval = voxel_map.atIdx(xc, yc, zc);
        if (flag == vflags_unmarked && (raycastval = voxel_raycast.atIdx(xc, yc, zc)) > 0.0f)
        {
          const float norm_val = raycastval / max_val;
          const float w_update_single = ray_update_weight * std::sqrt(norm_val);
          const float w1 = std::clamp(std::pow(1.0f - w_update_single, detection_its_diff), 0.0f, 1.0f);
          const float w2 = 1.0f - w1;
          mapval = w1 * mapval + w2 * ray_update_score;
        }
      });
    }
    voxel_flags.clear();  // :1602
    return VOFOD_OK;
  }

  // ---- classify_cluster: vofod_nodelet.cpp:1648-1731
  vofod_cluster_info classify_cluster(const std::vector<int>& idcs, const vofod_pose& tf, const vofod_params& p)
  {
    vofod_cluster_info ret;
    std::memset(&ret, 0, sizeof(ret));
    ret.cclass = VOFOD_CLASS_INVALID;
    ret.label = idcs.front();
    ret.n_points = int32_t(idcs.size());
    ret.obb_size = std::numeric_limits<float>::quiet_NaN();
    moment_of_inertia(cloud_weighted, idcs, ret);
    if (int(idcs.size()) < p.cls_min_points)
      return ret;
    const float ddx = tf.t[0] - ret.obb_center[0], ddy = tf.t[1] - ret.obb_center[1], ddz = tf.t[2] - ret.obb_center[2];
    const double dist = double(std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz));
    if (dist > p.cls_max_distance)
      return ret;
    const float ex = ret.obb_max[0] - ret.obb_min[0], ey = ret.obb_max[1] - ret.obb_min[1], ez = ret.obb_max[2] - ret.obb_min[2];
    ret.obb_size = std::sqrt(ex * ex + ey * ey + ez * ez);
    if (double(ret.obb_size) > p.cls_max_size)
      return ret;
    bool is_floating = true;
    if (background_pts_sufficient && sure_background_sufficient)
    {
      const int max_explore_voxel_size = int((double(ret.obb_size) + p.cls_max_explore_distance) / double(voxel_size));
      for (const int idx : idcs)
      {
        const auto& pt = cloud_weighted[idx];
        const auto [is_connected, explored] =
            voxel_map.exploreToGround(pt.x, pt.y, pt.z, float(p.thr_frontiers), float(p.thr_new_obstacles), float(max_explore_voxel_size));
        if (is_connected)
        {
          is_floating = false;
          break;
        } else
          for (const auto& e : explored)
            voxel_map.at(e) = float(p.thr_frontiers);
      }
    } else
      is_floating = false;
    ret.cclass = is_floating ? VOFOD_CLASS_MAV : VOFOD_CLASS_UNKNOWN;
    return ret;
  }

  // ---- extractDetections: vofod_nodelet.cpp:834-879
  void extract_detections(const vofod_pose& tf, const vofod_params& p)
  {
    detections.clear();
    for (size_t c = 0; c < cluster_infos.size(); c++)
    {
      const auto& cl = cluster_infos[c];
      if (cl.cclass != VOFOD_CLASS_MAV)
        continue;
      const auto& idcs = far_clusters[c];
      const float ddx = tf.t[0] - cl.obb_center[0], ddy = tf.t[1] - cl.obb_center[1], ddz = tf.t[2] - cl.obb_center[2];
      const double det_dist = double(std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz));
      vofod_detection det;
      std::memset(&det, 0, sizeof(det));
      det.id = int32_t(last_detection_id++);
      det.label = cl.label;
      det.n_points = idcs.size();
      for (int a = 0; a < 3; a++)
      {
        det.aabb_min[a] = cl.aabb_min[a]; det.aabb_max[a] = cl.aabb_max[a];
        det.obb_min[a] = cl.obb_min[a]; det.obb_max[a] = cl.obb_max[a];
        det.position[a] = cl.obb_center[a];
      }
      std::memcpy(det.obb_rot, cl.obb_rot, sizeof(det.obb_rot));
      const float cv = float(std::sqrt(det_dist) * p.output_position_sigma);
      det.covariance[0] = det.covariance[4] = det.covariance[8] = cv;
      VoxelMap submap = voxel_map.getSubmapCopy(vec3f{cl.aabb_min[0], cl.aabb_min[1], cl.aabb_min[2]}, vec3f{cl.aabb_max[0], cl.aabb_max[1], cl.aabb_max[2]}, 2);
      for (const int idx : idcs)
      {
        const auto& pt = cloud_weighted[idx];
        submap.at(pt.x, pt.y, pt.z) = float(p.score_ray);
      }
      double uncertainty = 0.0;
      for (const float val : submap.m_data)
        uncertainty += 1.0 - double(val) / p.score_ray;
      uncertainty /= double(idcs.size());
      det.confidence = double(float(1.0 / std::exp(uncertainty)));
      const double vray_res = double(p.sensor_vfov) / double(H);
      const double hray_res = 2 * M_PI / double(W);
      const double pdet_vert = std::min(std::atan(1.0 / det_dist) / (vray_res * p.cls_min_points), 1.0);
      const double pdet_hori = std::min(std::atan(1.0 / det_dist) / hray_res, 1.0);
      det.detection_probability = pdet_vert * pdet_hori;
      detections.push_back(det);
    }
  }

  // ---- updateSeparatedBGClusters: vofod_nodelet.cpp:1126-1278
  std::vector<vofod_xyzi> sep_raw;
  std::vector<vofod_vox> sep_ds;
  clusters_t sep_clusters;
  int sepclusters(const int its_diff_in, const vofod_params& p)
  {
    if (p.sep_pause)
      return VOFOD_W_PAUSED;
    const double max_dist = p.sep_max_bg_distance;
    const float thr_new = float(p.thr_new_obstacles);
    const float thr_sure = float(p.thr_sure_obstacles);
    const unsigned n_pts_sure_cluster = unsigned(p.sep_min_sure_points);
    const float max_dist_idx = float(max_dist / double(voxel_size));
    const int max_voxel_dist = int(std::ceil(max_dist_idx));
    bg_local.copyDataIdx(voxel_map);  // :1148
    sep_raw = bg_local.voxelsAsPC(thr_new, true, false);  // voxelsAsVoxelPC
    if (sep_raw.empty())
      return VOFOD_W_EMPTY;
    const float lsz = float(std::max(max_voxel_dist - 1, 0));
    const float no_align[3] = {0, 0, 0};
    if (voxel_grid_counted(sep_raw, lsz, thr_sure, false, no_align, sep_ds) < 0)
      return VOFOD_E_OVERFLOW;
    sep_clusters = euclidean_clusters(&sep_ds[0].x, 4, sep_ds.size(), float(max_voxel_dist), std_sort_ties);
    std::vector<size_t> n_sure;
    n_sure.reserve(sep_clusters.clusters.size());
    for (const auto& cl : sep_clusters.clusters)
    {
      int acc = 0;  // std::accumulate with int init
      for (const int idx : cl)
        acc = int(unsigned(acc) + sep_ds[idx].count);
      n_sure.push_back(size_t(acc));
    }
    const size_t n_sure_clusters = std::count_if(n_sure.begin(), n_sure.end(), [&](const size_t a) { return a >= n_pts_sure_cluster; });
    if (n_sure_clusters == 0)
    {
      sure_background_sufficient = false;
      return VOFOD_OK;
    }
    sure_background_sufficient = true;
    const float detection_its_diff = float(std::max(its_diff_in, 1));
    std::vector<vec3i> offsets;
    for (int x = -max_voxel_dist; x <= max_voxel_dist; x++)
      for (int y = -max_voxel_dist; y <= max_voxel_dist; y++)
        for (int z = -max_voxel_dist; z <= max_voxel_dist; z++)
        {
          const int nrm = int(std::sqrt(double(x * x + y * y + z * z)));  // Eigen int .norm() truncation
          if (float(nrm) <= max_dist_idx)
            offsets.push_back({x, y, z});
        }
    const float update_val = float(p.score_ray);
    const float w1 = std::clamp(std::pow(1.0f - 0.5f, detection_its_diff), 0.0f, 1.0f);
    const float w2 = 1.0f - w1;
    for (size_t it = 0; it < sep_clusters.clusters.size(); it++)
    {
      if (unsigned(n_sure[it]) < n_pts_sure_cluster)
      {
        for (const int idx : sep_clusters.clusters[it])
        {
          const int px = int(sep_ds[idx].x), py = int(sep_ds[idx].y), pz = int(sep_ds[idx].z);  // cast<int>() truncation
          for (const auto& o : offsets)
          {
            const int x = px + o.x, y = py + o.y, z = pz + o.z;
            if (!voxel_map.inLimitsIdx(x, y, z))
              continue;
            float& mapval = voxel_map.atIdx(x, y, z);
            mapval = w1 * mapval + w2 * update_val;
          }
        }
      }
    }
    return VOFOD_OK;
  }

  // ---- processMsg in the deterministic schedule S1
  bool raycast_pending = false;  // an accumulate whose apply was deferred to a later scan
  int process_scan(const vofod_pt* scan, const size_t n, const vofod_pose& tf, const vofod_params& p, const vofod_schedule& s, vofod_scan_result& res)
  {
    std::memset(&res, 0, sizeof(res));
    std::fill(stage_ms, stage_ms + VOFOD_N_STAGES, 0.0);
    stage_clock clk;
    clk.start();
    for (int r = 0; r < s.n_range_seeds; r++)
      range_update(s.range_pt, p);
    stage_ms[0] = clk.lap();
    if (n != size_t(W) * size_t(H))
      return VOFOD_E_DIMS;
    const int rc = filter_and_transform(scan, n, tf, p);
    if (rc < 0)
      return VOFOD_E_OVERFLOW;
    stage_ms[1] = clk.lap();  // "filtering"
    clusters = euclidean_clusters(cloud_weighted.empty() ? nullptr : &cloud_weighted[0].x, 4, cloud_weighted.size(), float(p.ground_points_max_distance), std_sort_ties);
    stage_ms[2] = clk.lap();  // "clusterization"
    find_close_far(p);
    stage_ms[3] = clk.lap();  // "close X far"
    update_vmaps(close_clusters, float(p.score_point), vflags_point);
    update_vmaps(far_clusters, float(p.score_unknown), vflags_unknown);
    detection_its++;
    stage_ms[4] = clk.lap();  // "vmap update"
    res.raycast_status = VOFOD_W_PAUSED;
    n_traversals = 0;  // reported per scan
    // the raycast thread (:1397-1606): accumulate, wait for a detection to finish, apply.  With raycast_defer_apply the apply waits
    // for the NEXT scan's point update (raycast_apply_pending), as the reference's threads settle (:950-957, 1530-1539)
    if (s.raycast_apply_pending && raycast_pending)
    {
      raycast_pending = false;
      res.raycast_status = raycast_apply(std::max(s.raycast_its_diff, 1), p);
      stage_ms[6] = clk.lap();
    }
    if (s.do_raycast)
    {
      res.raycast_status = raycast_accumulate(scan, n, tf, p);
      stage_ms[5] = clk.lap();  // "raycasting"
      if (res.raycast_status >= 0 && res.raycast_status != VOFOD_W_PAUSED && res.raycast_status != VOFOD_W_SENSOR_OOB)
      {
        if (s.raycast_defer_apply)
          raycast_pending = true;
        else
          res.raycast_status = std::max(res.raycast_status, raycast_apply(std::max(s.raycast_its_diff, 1), p));
      }
      stage_ms[6] += clk.lap();  // raycast "vmap update"
    }
    cluster_infos.clear();
    detections.clear();
    if (s.do_classify)
    {
      for (const auto& cl : far_clusters)
        cluster_infos.push_back(classify_cluster(cl, tf, p));
      stage_ms[7] = clk.lap();  // "classification"
      extract_detections(tf, p);
      stage_ms[8] = clk.lap();
    }
    res.sep_status = VOFOD_W_PAUSED;
    if (s.do_sepclusters)
    {
      res.sep_status = sepclusters(s.sep_its_diff, p);
      stage_ms[9] = clk.lap();  // "sep bg clusters"
    }
    res.n_traversals = n_traversals;
    res.n_bg = n_bg;
    res.n_filtered = n_filtered;
    res.n_voxels = uint32_t(cloud_weighted.size());
    res.n_clusters = uint32_t(clusters.clusters.size());
    res.n_close_clusters = uint32_t(close_clusters.size());
    res.n_far_clusters = uint32_t(far_clusters.size());
    res.n_detections = uint32_t(detections.size());
    res.background_pts_sufficient = background_pts_sufficient;
    res.sure_background_sufficient = sure_background_sufficient;
    return VOFOD_OK;
  }
};
}  // namespace vo

// =====================================================================================================
// C API (ctypes).  Mirrors libvofod_cuda's entry points one to one under the vo_ prefix so that the
// parity tests read the same on both sides.
// =====================================================================================================
using vo::Oracle;
extern "C" {

Oracle* vo_create() { return new Oracle(); }
void vo_destroy(Oracle* o) { delete o; }

void vo_set_modes(Oracle* o, int track_counts, int apply_from_fixed, int frac_bits)
{
  o->track_counts = track_counts != 0;
  o->apply_from_fixed = apply_from_fixed != 0;
  o->frac_bits = frac_bits;
}

void vo_set_std_sort_ties(Oracle* o, int on) { o->std_sort_ties = on != 0; }

void vo_reset(Oracle* o, const vofod_params* p, float voxel_size) { o->reset(*p, voxel_size); }

void vo_map_resize(Oracle* o, const float center[3], const float dims[3], float vs)
{
  o->voxel_size = vs;
  o->voxel_map.resize(vo::vec3f{center[0], center[1], center[2]}, vo::vec3f{dims[0], dims[1], dims[2]}, vs);
  o->voxel_flags.resizeAs(o->voxel_map);
  o->voxel_raycast.resizeAs(o->voxel_map);
  o->bg_local.resizeAs(o->voxel_map);
  o->voxel_flags.clear();
  o->voxel_raycast.clear();
  o->on_resize();
}
void vo_map_resize_idx(Oracle* o, const float offset[3], const int32_t sizes[3], float vs)
{
  o->voxel_size = vs;
  o->voxel_map.resize(vo::vec3f{offset[0], offset[1], offset[2]}, vo::vec3i{sizes[0], sizes[1], sizes[2]}, vs);
  o->voxel_flags.resizeAs(o->voxel_map);
  o->voxel_raycast.resizeAs(o->voxel_map);
  o->bg_local.resizeAs(o->voxel_map);
  o->voxel_flags.clear();
  o->voxel_raycast.clear();
  o->on_resize();
}
void vo_map_info_get(const Oracle* o, vofod_map_info* out)
{
  const auto& m = o->voxel_map;
  out->offset[0] = m.m_offset_x; out->offset[1] = m.m_offset_y; out->offset[2] = m.m_offset_z;
  out->sizes[0] = m.m_size_x; out->sizes[1] = m.m_size_y; out->sizes[2] = m.m_size_z;
  out->voxel_size = m.m_voxel_size;
  out->n_cells = m.size();
  out->slab_axis = 0; out->slab_lo = 0; out->slab_hi = m.m_size_x;
  for (int a = 0; a < 3; a++) { out->storage_lo[a] = 0; out->storage_size[a] = out->sizes[a]; }
  out->_pad = 0;
}
static vo::VoxelMap& which_map(Oracle* o, int which) { return which == VOFOD_MAP_SCORE ? o->voxel_map : which == VOFOD_MAP_FLAGS ? o->voxel_flags : o->voxel_raycast; }
void vo_map_set_to(Oracle* o, int which, float v) { which_map(o, which).setTo(v); }
float* vo_map_data(Oracle* o, int which) { return which_map(o, which).m_data.data(); }
uint32_t* vo_ray_counts(Oracle* o) { return o->ray_count.data(); }
int64_t* vo_ray_fixed(Oracle* o) { return o->ray_fixed.data(); }
void vo_map_set_inf(Oracle* o, const float* xyz, size_t n)
{
  for (size_t i = 0; i < n; i++)
    if (o->voxel_map.inLimits(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]))
      o->voxel_map.at(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]) = std::numeric_limits<float>::infinity();
}
uint64_t vo_map_count_over(Oracle* o, float thr) { return o->voxel_map.nVoxelsOver(thr); }
size_t vo_map_compact_over(Oracle* o, float thr, int greater, int metric, vofod_xyzi* out, size_t cap)
{
  const auto v = o->voxel_map.voxelsAsPC(thr, greater != 0, metric != 0);
  if (out)
    std::memcpy(out, v.data(), std::min(cap, v.size()) * sizeof(vofod_xyzi));
  return v.size();
}
void vo_map_has_close_to(Oracle* o, const float* xyz, size_t n, float max_dist, float thr, uint8_t* out)
{
  for (size_t i = 0; i < n; i++)
    out[i] = o->voxel_map.hasCloseTo(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], max_dist, thr);
}
size_t vo_map_explore_to_ground(Oracle* o, const float pt[3], float unk, float gnd, float maxd, int* connected, int32_t* idx3, size_t cap)
{
  const auto [c, v] = o->voxel_map.exploreToGround(pt[0], pt[1], pt[2], unk, gnd, maxd);
  *connected = c;
  for (size_t i = 0; i < std::min(cap, v.size()); i++)
  {
    idx3[3 * i] = std::get<0>(v[i]); idx3[3 * i + 1] = std::get<1>(v[i]); idx3[3 * i + 2] = std::get<2>(v[i]);
  }
  return v.size();
}
void vo_map_is_floating(Oracle* o, const float* xyz, size_t n, float thr, uint8_t* out)
{
  for (size_t i = 0; i < n; i++)
  {
    const auto [x, y, z] = o->voxel_map.coordToIdx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[i] = o->voxel_map.isFloatingIdx(x, y, z, thr);
  }
}
size_t vo_map_submap_copy(Oracle* o, const float mn[3], const float mx[3], int inflate, float* out, size_t cap, int32_t sizes[3], float offset[3])
{
  const auto s = o->voxel_map.getSubmapCopy(vo::vec3f{mn[0], mn[1], mn[2]}, vo::vec3f{mx[0], mx[1], mx[2]}, inflate);
  sizes[0] = s.m_size_x; sizes[1] = s.m_size_y; sizes[2] = s.m_size_z;
  offset[0] = s.m_offset_x; offset[1] = s.m_offset_y; offset[2] = s.m_offset_z;
  if (out)
    std::memcpy(out, s.m_data.data(), std::min(cap, s.m_data.size()) * sizeof(float));
  return s.m_data.size();
}
size_t vo_map_trace_ray(Oracle* o, const float start[3], const float dir[3], float length, float* ddist, int32_t* idx3, size_t cap)
{
  size_t n = 0;
  o->voxel_map.forEachRay(vo::vec3f{start[0], start[1], start[2]}, vo::vec3f{dir[0], dir[1], dir[2]}, length, [&](float d, int x, int y, int z) {
    if (n < cap)
    {
      ddist[n] = d;
      idx3[3 * n] = x; idx3[3 * n + 1] = y; idx3[3 * n + 2] = z;
    }
    n++;
  });
  return n;
}
void vo_coord_to_idx(Oracle* o, const float* xyz, size_t n, int32_t* idx3)
{
  for (size_t i = 0; i < n; i++)
  {
    const auto [x, y, z] = o->voxel_map.coordToIdx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    idx3[3 * i] = x; idx3[3 * i + 1] = y; idx3[3 * i + 2] = z;
  }
}
void vo_idx_to_coord(Oracle* o, const int32_t* idx3, size_t n, float* xyz)
{
  for (size_t i = 0; i < n; i++)
  {
    const auto [x, y, z] = o->voxel_map.idxToCoord(idx3[3 * i], idx3[3 * i + 1], idx3[3 * i + 2]);
    xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
  }
}

int vo_set_sensor(Oracle* o, int W, int H, const float* dirs, const float* offs, const uint8_t* mask)
{
  const size_t n = size_t(W) * size_t(H);
  o->W = W; o->H = H;
  o->lut_dirs.assign(dirs, dirs + 3 * n);
  if (offs)
    o->lut_offs.assign(offs, offs + 3 * n);
  else
    o->lut_offs.assign(3 * n, 0.0f);
  if (mask)
    o->mask.assign(mask, mask + n);
  else
    o->mask.assign(n, 1);  // vofod_nodelet.cpp:558
  return 0;
}
// initialize_sensor_lut_simulation: vofod_nodelet.cpp:374-420 (double math, stored fp32, not renormalised)
void vo_sim_lut(int w, int h, double vfov_in, float* dirs3xN)
{
  const float m_sensor_vfov = float(vfov_in);  // a float member of the nodelet (:2311) — pinned by oracle/_ref (the sliced function itself)
  const double vfov = m_sensor_vfov;
  const double yAngle_step = (2.0 * M_PI - 0.0) / (w - 1);
  const double pAngle_step = (vfov / 2.0 - (-vfov / 2.0)) / (h - 1);
  for (int row = 0; row < h; row++)
    for (int col = 0; col < w; col++)
    {
      const double yAngle = col * yAngle_step + 0.0;
      const double pAngle = row * pAngle_step + (-vfov / 2.0);
      float* d = dirs3xN + 3 * (size_t(col) + size_t(row) * w);
      d[0] = float(cos(pAngle) * cos(yAngle));
      d[1] = float(cos(pAngle) * sin(yAngle));
      d[2] = float(sin(pAngle));
    }
}

int vo_filter_voxelize(Oracle* o, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, vofod_vox* out, size_t cap, size_t* m)
{
  if (o->filter_and_transform(scan, n, *tf, *p) < 0)
    return VOFOD_E_OVERFLOW;
  *m = o->cloud_weighted.size();
  if (*m > cap)
    return VOFOD_E_CAPACITY;
  std::memcpy(out, o->cloud_weighted.data(), *m * sizeof(vofod_vox));
  return 0;
}
int vo_voxel_grid_weighted(const float* xyz, size_t n, float leaf, const float* align, vofod_vox* out, size_t cap, size_t* m)
{
  std::vector<Oracle::pt3> in(n);
  for (size_t i = 0; i < n; i++)
    in[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  std::vector<vofod_vox> o;
  const float zero[3] = {0, 0, 0};
  if (vo::voxel_grid_weighted(in, leaf, align != nullptr, align ? align : zero, o) < 0)
    return VOFOD_E_OVERFLOW;
  *m = o.size();
  if (*m > cap)
    return VOFOD_E_CAPACITY;
  std::memcpy(out, o.data(), *m * sizeof(vofod_vox));
  return 0;
}
int vo_voxel_grid_counted(const vofod_xyzi* pts, size_t n, float leaf, float thr, const float* align, vofod_vox* out, size_t cap, size_t* m)
{
  std::vector<vofod_xyzi> in(pts, pts + n);
  std::vector<vofod_vox> o;
  const float zero[3] = {0, 0, 0};
  if (vo::voxel_grid_counted(in, leaf, thr, align != nullptr, align ? align : zero, o) < 0)
    return VOFOD_E_OVERFLOW;
  *m = o.size();
  if (*m > cap)
    return VOFOD_E_CAPACITY;
  std::memcpy(out, o.data(), *m * sizeof(vofod_vox));
  return 0;
}
int vo_cluster(const float* xyz, size_t m, float tol, int32_t* labels, size_t* n_clusters)
{
  const auto c = vo::euclidean_clusters(xyz, 3, m, tol);
  std::memcpy(labels, c.labels.data(), m * sizeof(int32_t));
  *n_clusters = c.clusters.size();
  return 0;
}
// use externally provided voxels + labels as the state of the "current scan" (staged parity tests)
static void adopt(Oracle* o, const vofod_vox* pts, const int32_t* labels, size_t m)
{
  o->cloud_weighted.assign(pts, pts + m);
  o->clusters.labels.assign(labels, labels + m);
  std::unordered_map<int32_t, std::vector<int>> by;
  for (size_t i = 0; i < m; i++)
    by[labels[i]].push_back(int(i));
  o->clusters.clusters.clear();
  for (auto& kv : by)
    o->clusters.clusters.push_back(std::move(kv.second));
  std::sort(o->clusters.clusters.begin(), o->clusters.clusters.end(), [](const auto& a, const auto& b) {
    if (a.size() != b.size())
      return a.size() > b.size();
    return a.front() < b.front();
  });
}
int vo_close_far(Oracle* o, const vofod_vox* pts, const int32_t* labels, size_t m, const vofod_params* p, uint8_t* point_close, uint64_t* n_bg)
{
  adopt(o, pts, labels, m);
  o->find_close_far(*p);
  std::memcpy(point_close, o->point_close.data(), m);
  *n_bg = o->n_bg;
  return 0;
}
int vo_range_update(Oracle* o, const float pt[3], const vofod_params* p)
{
  o->range_update(pt, *p);
  return 0;
}
int vo_update_points(Oracle* o, const vofod_vox* pts, const uint8_t* sel, int sel_value, size_t n, float score, float flag)
{
  for (size_t i = 0; i < n; i++)
    if (!sel || sel[i] == sel_value)
      o->update_voxel(pts[i], score, flag);
  return 0;
}
int vo_raycast_accumulate(Oracle* o, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, uint64_t* n_trav)
{
  const int rc = o->raycast_accumulate(scan, n, *tf, *p);
  *n_trav = o->n_traversals;
  return rc;
}
int vo_raycast_apply(Oracle* o, int its_diff, const vofod_params* p) { return o->raycast_apply(its_diff, *p); }
int vo_classify_detect(Oracle* o, const vofod_vox* pts, const int32_t* labels, const uint8_t* point_close, size_t m, const vofod_pose* tf,
                       const vofod_params* p, vofod_detection* dets, size_t det_cap, size_t* n_dets, vofod_cluster_info* cls, size_t cl_cap, size_t* n_far)
{
  adopt(o, pts, labels, m);
  o->far_clusters.clear();
  for (const auto& cl : o->clusters.clusters)
    if (!point_close[cl.front()])
      o->far_clusters.push_back(cl);
  o->cluster_infos.clear();
  for (const auto& cl : o->far_clusters)
    o->cluster_infos.push_back(o->classify_cluster(cl, *tf, *p));
  o->extract_detections(*tf, *p);
  *n_far = o->cluster_infos.size();
  *n_dets = o->detections.size();
  if (cls)
    std::memcpy(cls, o->cluster_infos.data(), std::min(cl_cap, *n_far) * sizeof(vofod_cluster_info));
  if (dets)
    std::memcpy(dets, o->detections.data(), std::min(det_cap, *n_dets) * sizeof(vofod_detection));
  return 0;
}
int vo_sepclusters(Oracle* o, int its_diff, const vofod_params* p, int* sure)
{
  const int rc = o->sepclusters(its_diff, *p);
  *sure = o->sure_background_sufficient;
  return rc;
}
void vo_state_get(const Oracle* o, int* bg, int* sure, uint32_t* id)
{
  *bg = o->background_pts_sufficient; *sure = o->sure_background_sufficient; *id = o->last_detection_id;
}
void vo_state_set(Oracle* o, int bg, int sure, uint32_t id)
{
  o->background_pts_sufficient = bg; o->sure_background_sufficient = sure; o->last_detection_id = id;
}
int vo_process_scan(Oracle* o, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                    vofod_detection* dets, size_t det_cap)
{
  const int rc = o->process_scan(scan, n, *tf, *p, *s, *res);
  if (rc == 0 && dets)
    std::memcpy(dets, o->detections.data(), std::min(det_cap, o->detections.size()) * sizeof(vofod_detection));
  return rc;
}
size_t vo_last_voxels(Oracle* o, vofod_vox* out, int32_t* labels, uint8_t* in_close, size_t cap)
{
  const size_t m = o->cloud_weighted.size();
  const size_t k = std::min(cap, m);
  if (out) std::memcpy(out, o->cloud_weighted.data(), k * sizeof(vofod_vox));
  if (labels) std::memcpy(labels, o->clusters.labels.data(), k * sizeof(int32_t));
  if (in_close && o->point_close.size() >= k) std::memcpy(in_close, o->point_close.data(), k);
  return m;
}
size_t vo_last_clusters(Oracle* o, vofod_cluster_info* out, size_t cap)
{
  const size_t n = o->cluster_infos.size();
  if (out) std::memcpy(out, o->cluster_infos.data(), std::min(cap, n) * sizeof(vofod_cluster_info));
  return n;
}
void vo_stage_times(const Oracle* o, double ms[VOFOD_N_STAGES]) { std::memcpy(ms, o->stage_ms, sizeof(o->stage_ms)); }
}
