// vofod_b200::VoxelMap — host-side adaptor with the public interface of the reference's vofod::VoxelMap
// (include/vofod/voxel_map.h:27-112), backed by libvofod_cuda: the grid lives on the GPU, a host mirror is synchronised
// lazily for the accessors that hand out references / iterators (at(), atIdx(), begin()/end(), forEachIdx, forEach), which
// is what the out-of-scope debug / ROS code of vofod_nodelet.cpp uses.  Row N1 of SURVEY.md §8f.
//
// Same types as the reference header: Eigen vectors, pcl::PointCloud<pcl::PointXYZI>, std::tuple index triples.  In a ROS
// workspace <pcl/common/common.h> is the real PCL; in this repository's tests it is the stand-in of oracle/shim.
// visualization() builds its CUBE_LIST from the GPU compaction (the reference's x-outer / z-inner loop order is the compaction's
// emission order); borderVisualization() / *VisualizationThreshold are host code.  Copies are deep (a new context on the same device), as
// cheap as the reference's for the empty maps that live in every cluster_t (vofod_nodelet.cpp:110-119).
// Differences a caller can observe: exploreToGround returns each explored cell once (the reference's DFS may list a cell
// several times) and in a different order; the class is as thread-unsafe as the reference's.
#pragma once
#include <pcl/common/common.h>
#include <visualization_msgs/Marker.h>
#include <vofod_cuda.h>

#include <algorithm>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace vofod_b200
{
class VoxelMap
{
public:
  using data_t = float;
  using data_container_t = std::vector<data_t>;
  using coord_t = float;
  using idx_t = int;
  using pt_t = pcl::PointXYZI;
  using pc_t = pcl::PointCloud<pt_t>;
  using vec3_t = Eigen::Matrix<coord_t, 3, 1>;
  using vec3i_t = Eigen::Matrix<idx_t, 3, 1>;
  using idx3_t = std::tuple<idx_t, idx_t, idx_t>;

  explicit VoxelMap(int device = 0) : m_device(device) {}  // the device context is created with the first resize
  ~VoxelMap()
  {
    if (m_ctx)
      vofod_destroy(m_ctx);
  }
  VoxelMap(const VoxelMap& o) : m_device(o.m_device) { *this = o; }
  VoxelMap& operator=(const VoxelMap& o)
  {
    if (this == &o)
      return *this;
    m_thresholds = o.m_thresholds;
    if (!o.m_ctx || o.size() == 0)
    {
      if (m_ctx)
        vofod_destroy(m_ctx);
      m_ctx = nullptr;
      m_info = vofod_map_info{};
      m_host.clear();
      m_host_valid = m_host_dirty = false;
      return *this;
    }
    resizeAs(o);
    copyDataIdx(const_cast<VoxelMap&>(o));
    to_device();
    return *this;
  }
  VoxelMap(VoxelMap&& o) noexcept { *this = std::move(o); }
  VoxelMap& operator=(VoxelMap&& o) noexcept
  {
    if (this != &o)
    {
      if (m_ctx)
        vofod_destroy(m_ctx);
      m_ctx = o.m_ctx;
      o.m_ctx = nullptr;
      m_device = o.m_device;
      m_info = o.m_info;
      o.m_info = vofod_map_info{};
      m_host = std::move(o.m_host);
      m_host_valid = o.m_host_valid;
      m_host_dirty = o.m_host_dirty;
      m_thresholds = std::move(o.m_thresholds);
    }
    return *this;
  }
  vofod_ctx* handle()
  {
    need_ctx();
    return m_ctx;
  }

  // ---- modifiers (voxel_map.cpp:11-63, 268-285) ----
  void resize(const vec3_t& offset, const vec3i_t& sizes, const coord_t voxel_size)
  {
    const float off[3] = {offset.x(), offset.y(), offset.z()};
    const int32_t sz[3] = {sizes.x(), sizes.y(), sizes.z()};
    need_ctx();
    ck(vofod_map_resize_idx(m_ctx, off, sz, voxel_size));
    after_resize();
  }
  void resize(const vec3_t& center, const vec3_t& dimensions, const coord_t voxel_size)
  {
    const float c[3] = {center.x(), center.y(), center.z()};
    const float d[3] = {dimensions.x(), dimensions.y(), dimensions.z()};
    need_ctx();
    ck(vofod_map_resize(m_ctx, c, d, voxel_size));
    after_resize();
  }
  void resize(const coord_t cx, const coord_t cy, const coord_t cz, const coord_t dx, const coord_t dy, const coord_t dz, const coord_t voxel_size)
  {
    resize(vec3_t(cx, cy, cz), vec3_t(dx, dy, dz), voxel_size);
  }
  void resizeAs(const VoxelMap& o) { resize(o.origin(), o.sizes(), o.m_info.voxel_size); }
  void setTo(const data_t value)
  {
    ck(vofod_map_set_to(m_ctx, VOFOD_MAP_SCORE, value));
    m_host_valid = m_host_dirty = false;
  }
  void clear() { setTo(data_t(0)); }
  void copyDataIdx(VoxelMap& from)
  {
    from.to_host();
    m_host = from.m_host;
    m_host_valid = m_host_dirty = true;
  }

  // ---- geometry (voxel_map.cpp:289-374, 592-619): host fp32, one rounding per operation as in the reference ----
  vec3_t dimensions() const { return vec3_t(m_info.voxel_size * m_info.sizes[0], m_info.voxel_size * m_info.sizes[1], m_info.voxel_size * m_info.sizes[2]); }
  vec3_t origin() const { return vec3_t(m_info.offset[0], m_info.offset[1], m_info.offset[2]); }
  idx3_t sizesIdx() const { return {m_info.sizes[0], m_info.sizes[1], m_info.sizes[2]}; }
  vec3i_t sizes() const { return vec3i_t(m_info.sizes[0], m_info.sizes[1], m_info.sizes[2]); }
  size_t size() const { return size_t(m_info.n_cells); }
  idx3_t coordToIdx(const coord_t x, const coord_t y, const coord_t z) const
  {
    const coord_t inv = coord_t(1) / m_info.voxel_size;
    const idx_t ix = idx_t(std::floor((x - m_info.offset[0]) * inv));
    const idx_t iy = idx_t(std::floor((y - m_info.offset[1]) * inv));
    const idx_t iz = idx_t(std::floor((z - m_info.offset[2]) * inv));
    return {ix, iy, iz};
  }
  vec3i_t coordToIdx(const vec3_t& c) const
  {
    const auto [x, y, z] = coordToIdx(c.x(), c.y(), c.z());
    return vec3i_t(x, y, z);
  }
  std::tuple<coord_t, coord_t, coord_t> idxToCoord(const idx_t ix, const idx_t iy, const idx_t iz) const
  {
    const coord_t x = (ix + coord_t(0.5)) * m_info.voxel_size + m_info.offset[0];
    const coord_t y = (iy + coord_t(0.5)) * m_info.voxel_size + m_info.offset[1];
    const coord_t z = (iz + coord_t(0.5)) * m_info.voxel_size + m_info.offset[2];
    return {x, y, z};
  }
  vec3_t idxToCoord(const vec3i_t& i) const
  {
    const auto [x, y, z] = idxToCoord(i.x(), i.y(), i.z());
    return vec3_t(x, y, z);
  }
  bool inLimitsIdx(const int ix, const int iy, const int iz) const
  {
    return ix >= 0 && ix < m_info.sizes[0] && iy >= 0 && iy < m_info.sizes[1] && iz >= 0 && iz < m_info.sizes[2];
  }
  bool inLimitsIdx(const vec3i_t& i) const { return inLimitsIdx(i.x(), i.y(), i.z()); }
  bool inLimits(const coord_t x, const coord_t y, const coord_t z) const
  {
    const auto [ix, iy, iz] = coordToIdx(x, y, z);
    return inLimitsIdx(ix, iy, iz);
  }
  idx_t manhattanDist(const vec3i_t& a, const vec3i_t& b) const { return (a - b).cwiseAbs().sum(); }
  idx_t manhattanDist(const idx3_t& a, const idx3_t& b) const
  {
    return std::abs(std::get<0>(a) - std::get<0>(b)) + std::abs(std::get<1>(a) - std::get<1>(b)) + std::abs(std::get<2>(a) - std::get<2>(b));
  }

  // ---- element access through the host mirror (voxel_map.cpp:67-154): std::vector::at semantics ----
  data_t& atIdx(const int ix, const int iy, const int iz)
  {
    to_host();
    m_host_dirty = true;
    return m_host.at(lin(ix, iy, iz));
  }
  data_t atIdx(const int ix, const int iy, const int iz) const
  {
    const_cast<VoxelMap*>(this)->to_host();
    return m_host.at(lin(ix, iy, iz));
  }
  data_t& at(const coord_t x, const coord_t y, const coord_t z)
  {
    const auto [ix, iy, iz] = coordToIdx(x, y, z);
    return atIdx(ix, iy, iz);
  }
  data_t at(const coord_t x, const coord_t y, const coord_t z) const
  {
    const auto [ix, iy, iz] = coordToIdx(x, y, z);
    return atIdx(ix, iy, iz);
  }
  data_t& at(const vec3i_t& i) { return atIdx(i.x(), i.y(), i.z()); }
  data_t at(const vec3i_t& i) const { return atIdx(i.x(), i.y(), i.z()); }
  data_t& at(const idx3_t& i) { return atIdx(std::get<0>(i), std::get<1>(i), std::get<2>(i)); }
  data_t at(const idx3_t& i) const { return atIdx(std::get<0>(i), std::get<1>(i), std::get<2>(i)); }
  data_container_t::iterator begin()
  {
    to_host();
    m_host_dirty = true;
    return m_host.begin();
  }
  data_container_t::iterator end()
  {
    to_host();
    return m_host.end();
  }

  // ---- whole-grid queries on the GPU ----
  uint64_t nVoxelsOver(const data_t threshold)  // voxel_map.cpp:216-222
  {
    to_device();
    uint64_t n = 0;
    ck(vofod_map_count_over(m_ctx, threshold, &n));
    return n;
  }
  pc_t::Ptr voxelsAsPC(const data_t threshold = std::numeric_limits<data_t>::lowest(), const bool greater_than = true, const pcl::PCLHeader& header = {})
  {
    return compact(threshold, greater_than, 1, header);  // :157-184
  }
  pc_t::Ptr voxelsAsVoxelPC(const data_t threshold = std::numeric_limits<data_t>::lowest(), const bool greater_than = true, const pcl::PCLHeader& header = {})
  {
    return compact(threshold, greater_than, 0, header);  // :187-212
  }
  bool hasCloseTo(const coord_t x, const coord_t y, const coord_t z, const coord_t max_dist, const data_t threshold)  // :376-400
  {
    to_device();
    const float p[3] = {x, y, z};
    uint8_t r = 0;
    ck(vofod_map_has_close_to(m_ctx, p, 1, max_dist, threshold, &r));
    return r != 0;
  }
  std::tuple<bool, std::vector<idx3_t>> exploreToGround(const coord_t x, const coord_t y, const coord_t z, const data_t unknown_threshold,
                                                        const data_t ground_threshold, const coord_t max_voxel_dist)  // :402-488
  {
    to_device();
    const float p[3] = {x, y, z};
    int connected = 0;
    size_t n = 0;
    std::vector<int32_t> idx(3 * 4096);
    int rc = vofod_map_explore_to_ground(m_ctx, p, unknown_threshold, ground_threshold, max_voxel_dist, &connected, idx.data(), idx.size() / 3, &n);
    if (rc == VOFOD_E_CAPACITY)
    {
      idx.resize(3 * n);
      rc = vofod_map_explore_to_ground(m_ctx, p, unknown_threshold, ground_threshold, max_voxel_dist, &connected, idx.data(), n, &n);
    }
    ck(rc);
    std::vector<idx3_t> cells;
    if (!connected)
      for (size_t i = 0; i < n; i++)
        cells.emplace_back(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
    return {connected != 0, cells};
  }
  bool isFloating(const coord_t x, const coord_t y, const coord_t z, const data_t threshold = data_t(-100))  // :491-495
  {
    to_device();
    const float p[3] = {x, y, z};
    uint8_t r = 0;
    ck(vofod_map_is_floating(m_ctx, p, 1, threshold, &r));
    return r != 0;
  }
  bool isFloatingIdx(const idx_t ix, const idx_t iy, const idx_t iz, const data_t threshold = data_t(-100))
  {
    const auto [x, y, z] = idxToCoord(ix, iy, iz);
    return isFloating(x, y, z, threshold);
  }
  VoxelMap getSubmapCopy(const vec3_t& min_pt, const vec3_t& max_pt, const int inflate = 0)  // :547-584
  {
    to_device();
    const float mn[3] = {min_pt.x(), min_pt.y(), min_pt.z()}, mx[3] = {max_pt.x(), max_pt.y(), max_pt.z()};
    int32_t sz[3];
    float off[3];
    std::vector<float> buf(1 << 16);
    int rc = vofod_map_submap_copy(m_ctx, mn, mx, inflate, buf.data(), buf.size(), sz, off);
    if (rc == VOFOD_E_CAPACITY)
    {
      buf.resize(size_t(sz[0]) * sz[1] * sz[2]);
      rc = vofod_map_submap_copy(m_ctx, mn, mx, inflate, buf.data(), buf.size(), sz, off);
    }
    ck(rc);
    VoxelMap ret(m_device);
    ret.resize(vec3_t(off[0], off[1], off[2]), vec3i_t(sz[0], sz[1], sz[2]), m_info.voxel_size);
    ret.m_host.assign(buf.begin(), buf.begin() + size_t(sz[0]) * sz[1] * sz[2]);
    ret.m_host_valid = ret.m_host_dirty = true;
    return ret;
  }

  // ---- visitors with host callbacks ----
  void forEachIdx(const std::function<void(data_t&, const idx_t, const idx_t, const idx_t)> f, const idx_t offset = 0)  // :518-534, same x/y/z order
  {
    to_host();
    m_host_dirty = true;
    for (idx_t x = offset; x < m_info.sizes[0] - offset; x++)
      for (idx_t y = offset; y < m_info.sizes[1] - offset; y++)
        for (idx_t z = offset; z < m_info.sizes[2] - offset; z++)
          f(m_host.at(lin(x, y, z)), x, y, z);
  }
  void forEach(const std::function<void(data_t&, const coord_t, const coord_t, const coord_t)> f, const idx_t offset = 0)
  {
    forEachIdx(
        [&](data_t& v, const idx_t ix, const idx_t iy, const idx_t iz) {
          const auto [x, y, z] = idxToCoord(ix, iy, iz);
          f(v, x, y, z);
        },
        offset);
  }
  // the traversal itself (voxel_map.cpp:229-263) runs on the GPU (vofod_map_trace_ray); the callback sees the same sequence
  void forEachRay(const vec3_t& start_pt, const vec3_t& dir, const coord_t length, const std::function<void(coord_t, const idx_t, const idx_t, const idx_t)> f)
  {
    const float s[3] = {start_pt.x(), start_pt.y(), start_pt.z()}, d[3] = {dir.x(), dir.y(), dir.z()};
    std::vector<float> dd(1024);
    std::vector<int32_t> idx(3 * 1024);
    size_t n = 0;
    int rc = vofod_map_trace_ray(m_ctx, s, d, length, dd.data(), idx.data(), dd.size(), &n);
    if (rc == VOFOD_E_CAPACITY)
    {
      dd.resize(n);
      idx.resize(3 * n);
      rc = vofod_map_trace_ray(m_ctx, s, d, length, dd.data(), idx.data(), dd.size(), &n);
    }
    ck(rc);
    for (size_t i = 0; i < n; i++)
      f(dd[i], idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
  }

  // ---- RViz markers (voxel_map.cpp:622-786) ----
  void clearVisualizationThresholds() { m_thresholds.clear(); }
  void addVisualizationThreshold(const data_t th, const std_msgs::ColorRGBA& th_color)
  {
    m_thresholds.emplace_back(th, th_color);
    std::sort(std::begin(m_thresholds), std::end(m_thresholds), [](const auto& v1, const auto& v2) { return v1.first < v2.first; });
  }
  // CUBE_LIST of every voxel above the lowest threshold, coloured by the highest threshold it exceeds (:622-669).  The reference walks
  // x (outer), y, z (inner) — the emission order of vofod_map_compact_over, so the cells come straight from the GPU compaction.
  visualization_msgs::Marker visualization(const std_msgs::Header& header) const
  {
    visualization_msgs::Marker ret;
    ret.header = header;
    ret.pose.position.x = m_info.offset[0] + m_info.voxel_size / coord_t(2);
    ret.pose.position.y = m_info.offset[1] + m_info.voxel_size / coord_t(2);
    ret.pose.position.z = m_info.offset[2] + m_info.voxel_size / coord_t(2);
    ret.pose.orientation.w = 1.0;
    ret.scale.x = ret.scale.y = ret.scale.z = m_info.voxel_size;
    ret.color.a = 1.0;
    ret.type = visualization_msgs::Marker::CUBE_LIST;
    if (m_thresholds.empty() || !m_ctx || size() == 0)
      return ret;
    VoxelMap* self = const_cast<VoxelMap*>(this);
    self->to_device();
    const data_t lowest = m_thresholds.front().first;
    size_t n = 0;
    int rc = vofod_map_compact_over(m_ctx, lowest, 1, 0, nullptr, 0, &n);
    std::vector<vofod_xyzi> buf(n);
    if (n)
      rc = vofod_map_compact_over(m_ctx, lowest, 1, 0, buf.data(), buf.size(), &n);
    ck(n ? rc : VOFOD_OK);
    ret.points.reserve(n);
    ret.colors.reserve(n);
    for (const vofod_xyzi& b : buf)
    {
      std_msgs::ColorRGBA color;
      for (const auto& [th, clr] : m_thresholds)
        if (b.intensity > th)
          color = clr;
      geometry_msgs::Point pt;
      pt.x = idx_t(b.x) * m_info.voxel_size;
      pt.y = idx_t(b.y) * m_info.voxel_size;
      pt.z = idx_t(b.z) * m_info.voxel_size;
      ret.points.push_back(pt);
      ret.colors.push_back(color);
    }
    return ret;
  }
  // LINE_LIST of the 12 edges of the map's box (:672-786): bottom loop, one riser, top loop, the three other risers
  visualization_msgs::Marker borderVisualization(const std_msgs::Header& header) const
  {
    visualization_msgs::Marker ret;
    ret.header = header;
    ret.pose.position.x = m_info.offset[0];
    ret.pose.position.y = m_info.offset[1];
    ret.pose.position.z = m_info.offset[2];
    ret.pose.orientation.w = 1.0;
    ret.scale.x = 0.05;
    ret.color.r = ret.color.g = ret.color.b = ret.color.a = 1.0;
    ret.type = visualization_msgs::Marker::LINE_LIST;
    const coord_t dim[3] = {m_info.sizes[0] * m_info.voxel_size, m_info.sizes[1] * m_info.voxel_size, m_info.sizes[2] * m_info.voxel_size};
    // corners as bit masks (x = 1, y = 2, z = 4), edges in the reference's order
    static const unsigned char edges[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {0, 4}, {4, 5}, {5, 7}, {7, 6}, {6, 4}, {1, 5}, {3, 7}, {2, 6}};
    ret.points.reserve(24);
    for (const auto& e : edges)
      for (int k = 0; k < 2; k++)
      {
        geometry_msgs::Point pt;
        pt.x = (e[k] & 1) ? dim[0] : coord_t(0);
        pt.y = (e[k] & 2) ? dim[1] : coord_t(0);
        pt.z = (e[k] & 4) ? dim[2] : coord_t(0);
        ret.points.push_back(pt);
      }
    return ret;
  }

private:
  void need_ctx()
  {
    if (!m_ctx && vofod_create(m_device, &m_ctx) != VOFOD_OK)
      throw std::runtime_error(std::string("vofod_create: ") + vofod_last_error(nullptr));  // no CPU fallback
  }
  void ck(const int rc) const
  {
    if (rc < 0)
      throw std::runtime_error(std::string("libvofod_cuda: ") + vofod_last_error(m_ctx));
  }
  size_t lin(const int ix, const int iy, const int iz) const
  {
    // the reference indexes with int arithmetic and lets std::vector::at reject what falls outside the array (voxel_map.cpp:81-82)
    return size_t(ix + iy * m_info.sizes[0] + iz * m_info.sizes[0] * m_info.sizes[1]);
  }
  void after_resize()
  {
    ck(vofod_map_info_get(m_ctx, &m_info));
    m_host.clear();
    m_host_valid = m_host_dirty = false;
  }
  void to_host()
  {
    if (m_host_valid)
      return;
    if (!m_ctx)  // never sized: an empty map, as the reference's default-constructed one
    {
      m_host.clear();
      m_host_valid = true;
      return;
    }
    m_host.resize(size());
    ck(vofod_map_download(m_ctx, VOFOD_MAP_SCORE, m_host.data(), m_host.size()));
    m_host_valid = true;
    m_host_dirty = false;
  }
  void to_device()
  {
    if (m_host_valid && m_host_dirty)
    {
      ck(vofod_map_upload(m_ctx, VOFOD_MAP_SCORE, m_host.data(), m_host.size()));
      m_host_dirty = false;
    }
  }
  pc_t::Ptr compact(const data_t threshold, const bool greater_than, const int metric, const pcl::PCLHeader& header)
  {
    to_device();
    size_t n = 0;
    int rc = vofod_map_compact_over(m_ctx, threshold, greater_than, metric, nullptr, 0, &n);
    std::vector<vofod_xyzi> buf(n);
    if (n)
      rc = vofod_map_compact_over(m_ctx, threshold, greater_than, metric, buf.data(), buf.size(), &n);
    ck(n ? rc : VOFOD_OK);
    pc_t::Ptr cloud = boost::make_shared<pc_t>();
    cloud->reserve(n);
    for (const auto& b : buf)
    {
      pt_t pt;
      pt.x = b.x;
      pt.y = b.y;
      pt.z = b.z;
      pt.intensity = b.intensity;
      cloud->push_back(pt);
    }
    cloud->header = header;
    return cloud;
  }

  vofod_ctx* m_ctx = nullptr;
  int m_device = 0;
  vofod_map_info m_info{};
  data_container_t m_host;
  bool m_host_valid = false, m_host_dirty = false;
  std::vector<std::pair<data_t, std_msgs::ColorRGBA>> m_thresholds;
};
}  // namespace vofod_b200
