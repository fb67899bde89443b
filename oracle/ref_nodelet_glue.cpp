// TEST INFRASTRUCTURE ONLY — the reference's per-scan member functions of src/vofod_nodelet.cpp, cut out of the reference file
// at build time (oracle/slice_nodelet.py -> oracle/_ref/nodelet_{types,members}.inc, never committed) and compiled inside
// the stand-in class below against oracle/shim.  What runs when a test calls vn_process_scan is therefore the reference's
// OWN statements (see the list in oracle/shim/nodelet_shim.h); this file only supplies the member variables
// (vofod_nodelet.cpp:2262-2339), the few lines of onInit that derive them from the parameters (:196-230) and the order in
// which the functions are called — the deterministic schedule S1 of SURVEY.md §8d, which follows processMsg (:882-964).
#include <nodelet_shim.h>

#include <vofod/pc_loader.h>
#include <vofod/types.h>
#include <vofod/voxel_grid_counted.h>
#include <vofod/voxel_grid_weighted.h>
#include <vofod/voxel_map.h>

#include <functional>

#include "../include/vofod_cuda.h"

namespace vofod
{
// vofod/DetectionParamsConfig.h (generated from config/dynamic_reconfigure/DetectionParams.cfg:16-44)
struct DetectionParamsConfig
{
  double ground_points_max_distance, output__position_sigma;
  double voxel_map__scores__point, voxel_map__scores__unknown, voxel_map__scores__ray;
  double voxel_map__thresholds__apriori_map, voxel_map__thresholds__new_obstacles, voxel_map__thresholds__sure_obstacles, voxel_map__thresholds__frontiers;
  int classification__min_points;
  double classification__max_size, classification__max_distance, classification__max_explore_distance;
  bool raycast__pause, raycast__new_update_rule;
  double raycast__max_distance, raycast__min_intensity, raycast__weight_coefficient;
  bool sepclusters__pause;
  double sepclusters__max_bg_distance;
  int sepclusters__min_sure_points;
};
struct drmgr_t { DetectionParamsConfig config; };

// hooks of the stand-in condition variable / scope timer: the places where the reference's other threads would advance
// m_detection_its while raycast_cloud / updateSeparatedBGClusters run (vofod_nodelet.cpp:950,1530-1539,1212)
static thread_local std::function<void()> g_on_cv_wait;
static thread_local std::function<void(const std::string&)> g_on_checkpoint;
struct FakeCv
{
  template <class L, class D>
  std::cv_status wait_for(L&, const D&) { if (g_on_cv_wait) g_on_cv_wait(); return std::cv_status::no_timeout; }
  void notify_one() {}
};
struct HookedTimer : mrs_lib::ScopeTimer
{
  using mrs_lib::ScopeTimer::ScopeTimer;
  void checkpoint(const std::string& s) { if (g_on_checkpoint) g_on_checkpoint(s); }
};
}  // namespace vofod
// the sliced code names mrs_lib::ScopeTimer: route it to the hooked one
namespace mrs_lib_hooked { using ScopeTimer = vofod::HookedTimer; using AtomicScopeFlag = mrs_lib::AtomicScopeFlag; }
#define mrs_lib mrs_lib_hooked

namespace vofod
{
#include "_ref/nodelet_types.inc"

class VoFOD
{
public:
  // ---- stand-ins for members the sliced functions name -----------------------------------------------------------
  drmgr_t m_drmgr_storage;
  drmgr_t* m_drmgr_ptr = &m_drmgr_storage;
  std::string m_world_frame_id = "world", m_node_name = "VoFOD";
  ros::Duration m_throttle_period, m_bgclusters_period;
  ros::Publisher m_pub_filtered_input_pc, m_pub_weighted_input_pc, m_pub_sepclusters_cluster_pc, m_pub_sepclusters_pc, m_pub_lidar_raycast, m_pub_lidar_fov,
      m_pub_apriori_pc, m_pub_lidar_mask;
  Eigen::Affine3d m_tf_to_world = Eigen::Affine3d::Identity();
  bool get_transform_to_world(const std::string&, const ros::Time&, Eigen::Affine3d& out) { out = m_tf_to_world; return true; }
  void publish_profile_start(profile_routines_t) {}
  void publish_profile_end(profile_routines_t) {}
  int lidar_visualization(const std_msgs::Header&, const std::vector<float>&) { return 0; }

  // ---- vofod_nodelet.cpp:2262-2339 ---------------------------------------------------------------------------------
  bool m_sensor_simulation = true, m_sensor_mask_mangle = true;
  float m_vmap_voxel_size = 0.5f, m_vmap_init_score = -740.f;
  std_msgs::ColorRGBA m_vflags_color_background, m_vflags_color_unknown;
  float m_exclude_box_offset_x, m_exclude_box_offset_y, m_exclude_box_offset_z, m_exclude_box_size_x, m_exclude_box_size_y, m_exclude_box_size_z;
  float m_oparea_offset_x, m_oparea_offset_y, m_oparea_offset_z, m_oparea_size_x, m_oparea_size_y, m_oparea_size_z;
  xyz_lut_t m_sensor_xyz_lut;
  std::vector<uint8_t> m_sensor_mask;
  float m_sensor_vfov = 0.f;
  int m_sensor_vrays = 0, m_sensor_hrays = 0;
  bool m_apriori_map_initialized = false, m_sensor_initialized = false, m_sensor_params_checked = true, m_sensor_params_ok = true;
  uint32_t m_last_detection_id = 0;
  std::atomic<bool> m_background_pts_sufficient{false}, m_sure_background_sufficient{false};
  uint64_t m_background_min_sufficient_pts = 0;
  std::mutex m_voxels_mtx;
  FakeCv m_detection_cv;
  std::thread m_raycast_thread;
  std::atomic<bool> m_raycast_running{false};
  std::atomic<int> m_detection_its{0};
  VoxelMap m_voxel_map;
  static constexpr float m_vflags_unmarked = 0.0f;
  static constexpr float m_vflags_point = 2.0f;
  static constexpr float m_vflags_unknown = 3.0f;
  VoxelMap m_voxel_flags;
  VoxelMap m_voxel_raycast;

  // ---- the reference's member functions, verbatim ------------------------------------------------------------------
#include "_ref/nodelet_members.inc"

  // ---- onInit, the lines that derive members from parameters (vofod_nodelet.cpp:196-230) ----------------------------
  void load(const vofod_params& p, const float voxel_size)
  {
    DetectionParamsConfig& c = m_drmgr_ptr->config;
    c.ground_points_max_distance = p.ground_points_max_distance;
    c.output__position_sigma = p.output_position_sigma;
    c.voxel_map__scores__point = p.score_point;
    c.voxel_map__scores__unknown = p.score_unknown;
    c.voxel_map__scores__ray = p.score_ray;
    c.voxel_map__thresholds__apriori_map = p.thr_apriori_map;
    c.voxel_map__thresholds__new_obstacles = p.thr_new_obstacles;
    c.voxel_map__thresholds__sure_obstacles = p.thr_sure_obstacles;
    c.voxel_map__thresholds__frontiers = p.thr_frontiers;
    c.classification__min_points = p.cls_min_points;
    c.classification__max_size = p.cls_max_size;
    c.classification__max_distance = p.cls_max_distance;
    c.classification__max_explore_distance = p.cls_max_explore_distance;
    c.raycast__pause = p.raycast_pause != 0;
    c.raycast__new_update_rule = p.raycast_new_update_rule != 0;
    c.raycast__max_distance = p.raycast_max_distance;
    c.raycast__min_intensity = p.raycast_min_intensity;
    c.raycast__weight_coefficient = p.raycast_weight_coefficient;
    c.sepclusters__pause = p.sep_pause != 0;
    c.sepclusters__max_bg_distance = p.sep_max_bg_distance;
    c.sepclusters__min_sure_points = p.sep_min_sure_points;
    m_vmap_voxel_size = voxel_size;
    m_vmap_init_score = p.score_init;
    m_sensor_vfov = p.sensor_vfov;
    m_exclude_box_offset_x = p.exclude_box_offset[0]; m_exclude_box_offset_y = p.exclude_box_offset[1]; m_exclude_box_offset_z = p.exclude_box_offset[2];
    m_exclude_box_size_x = p.exclude_box_size[0]; m_exclude_box_size_y = p.exclude_box_size[1]; m_exclude_box_size_z = p.exclude_box_size[2];
    m_exclude_box_offset_z = m_exclude_box_offset_z + m_exclude_box_size_z / 2.0f;  // :204
    m_oparea_offset_x = p.oparea_offset[0]; m_oparea_offset_y = p.oparea_offset[1]; m_oparea_offset_z = p.oparea_offset[2];
    m_oparea_size_x = p.oparea_size[0]; m_oparea_size_y = p.oparea_size[1]; m_oparea_size_z = p.oparea_size[2];
    m_oparea_offset_z = m_oparea_offset_z + m_oparea_size_z / 2.0f;                 // :212
    const auto background_sufficient_points_ratio = p.background_sufficient_points_ratio;
    const auto n_voxels_xy = m_oparea_size_x / m_vmap_voxel_size * m_oparea_size_y / m_vmap_voxel_size;  // :229
    m_background_min_sufficient_pts = n_voxels_xy * background_sufficient_points_ratio;                  // :230
  }

  // ---- results of the last scan, kept for the tests ------------------------------------------------------------------
  pc_XYZR_t::Ptr last_cloud;
  std::vector<pcl::PointIndices> last_clusters;
  std::vector<uint8_t> last_in_close;
  std::vector<cluster_t> last_far;
  std::vector<float> last_gaps;
  std::vector<detection_t> last_dets;
  VoxelMap local_vmap;  // bgclusters_loop's copy (:1283-1288)
  bool local_sized = false;
};
}  // namespace vofod
#undef mrs_lib

using vofod::VoFOD;

static void fill_tf(Eigen::Affine3f& tf, const vofod_pose& p)
{
  for (int r = 0; r < 3; r++)
  {
    for (int c = 0; c < 3; c++)
      tf.lin.m[r][c] = p.R[3 * r + c];
    tf.t.v[r] = p.t[r];
  }
}

extern "C" {
VoFOD* vn_create() { return new VoFOD(); }
void vn_destroy(VoFOD* v) { delete v; }

// onInit's parameter block + VoFOD::reset() (:1610-1632)
void vn_reset(VoFOD* v, const vofod_params* p, float voxel_size)
{
  v->load(*p, voxel_size);
  v->reset();
  v->m_last_detection_id = 0;
  v->m_background_pts_sufficient = false;
  v->m_sure_background_sufficient = false;
  v->local_sized = false;
}

// dynamic_reconfigure: the tunables are re-read on every use (no reset)
void vn_set_params(VoFOD* v, const vofod_params* p) { v->load(*p, v->m_vmap_voxel_size); }

// initialize_sensor_lut_simulation (:374-420) + the default all-ones mask of load_mask (:558)
void vn_sensor_sim(VoFOD* v, int W, int H)
{
  v->m_sensor_hrays = W;
  v->m_sensor_vrays = H;
  v->initialize_sensor_lut_simulation(size_t(W), size_t(H));
  v->m_sensor_mask.assign(size_t(W) * H, 1);
}
void vn_sensor_set(VoFOD* v, const float* dirs, const float* offs, const uint8_t* mask)
{
  const size_t n = size_t(v->m_sensor_hrays) * v->m_sensor_vrays;
  for (size_t i = 0; i < n; i++)
  {
    if (dirs)
      v->m_sensor_xyz_lut.directions.col(long(i)) = Eigen::Vector3f(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    if (offs)
      v->m_sensor_xyz_lut.offsets.col(long(i)) = Eigen::Vector3f(offs[3 * i], offs[3 * i + 1], offs[3 * i + 2]);
  }
  if (mask)
    v->m_sensor_mask.assign(mask, mask + n);
}
void vn_sensor_get(VoFOD* v, float* dirs, float* offs)
{
  const size_t n = size_t(v->m_sensor_xyz_lut.directions.cols());
  std::memcpy(dirs, v->m_sensor_xyz_lut.directions.d.data(), n * 12);
  if (offs)
    std::memcpy(offs, v->m_sensor_xyz_lut.offsets.d.data(), n * 12);
}
// load_mask (:506-560) on an in-memory image (H rows x W cols, row-major u8) instead of a PNG on disk
void vn_load_mask(VoFOD* v, const uint8_t* img, int cols, int rows, int W, int H, int mangle, const int32_t* pixel_shift_by_row, uint8_t* out)
{
  cv::Mat m(cv::Size{cols, rows}, CV_8UC1, 0);
  std::memcpy(m.data, img, size_t(cols) * rows);
  cv::shim_images()["<memory>"] = m;
  v->m_sensor_hrays = W;
  v->m_sensor_vrays = H;
  v->m_sensor_mask_mangle = mangle != 0;
  const std::vector<int> shift(pixel_shift_by_row, pixel_shift_by_row + H);
  const std::vector<uint8_t> r = v->load_mask(img ? "<memory>" : "<missing>", size_t(W), size_t(H), shift);
  std::memcpy(out, r.data(), r.size());
}

// processMsg(Range) (:581-613) with the range point already in the world frame
void vn_range(VoFOD* v, const float world_pt[3])
{
  auto msg = std::make_shared<sensor_msgs::Range>();
  msg->range = 0.0f;
  msg->min_range = -1.0f;  // "range <= min && range >= max" (:585) must not trigger
  msg->max_range = 1.0f;
  v->m_tf_to_world = Eigen::Affine3d::Identity();
  for (int a = 0; a < 3; a++)
    v->m_tf_to_world.t.v[a] = double(world_pt[a]);
  v->processMsg(sensor_msgs::Range::ConstPtr(msg));
}

float* vn_map(VoFOD* v, int which, size_t* n)
{
  vofod::VoxelMap& m = which == VOFOD_MAP_SCORE ? v->m_voxel_map : which == VOFOD_MAP_FLAGS ? v->m_voxel_flags : v->m_voxel_raycast;
  *n = m.size();
  return &*m.begin();
}
void vn_map_info(VoFOD* v, vofod_map_info* out)
{
  const auto o = v->m_voxel_map.origin();
  const auto s = v->m_voxel_map.sizes();
  std::memset(out, 0, sizeof(*out));
  for (int a = 0; a < 3; a++) { out->offset[a] = o[a]; out->sizes[a] = s[a]; out->storage_size[a] = s[a]; }
  out->n_cells = v->m_voxel_map.size();
  out->voxel_size = v->m_vmap_voxel_size;
  out->slab_hi = s[0];
}
void vn_state_get(VoFOD* v, int* bg, int* sure, uint32_t* id)
{
  *bg = v->m_background_pts_sufficient;
  *sure = v->m_sure_background_sufficient;
  *id = v->m_last_detection_id;
}
void vn_state_set(VoFOD* v, int bg, int sure, uint32_t id)
{
  v->m_background_pts_sufficient = bg != 0;
  v->m_sure_background_sufficient = sure != 0;
  v->m_last_detection_id = id;
}

// One scan of schedule S1, every stage a call of the reference's own member function, in the order of
// processMsg (:926-964) with the raycast run in line (one in flight, applied with its_diff) and the background pass after
// the detections.
int vn_process_scan(VoFOD* v, const vofod_pt* scan, size_t n, const vofod_pose* pose, const vofod_schedule* s, vofod_scan_result* res)
{
  using namespace vofod;
  std::memset(res, 0, sizeof(*res));
  for (int k = 0; k < s->n_range_seeds; k++)
    vn_range(v, s->range_pt);
  vofod::pc_t::Ptr cloud = boost::make_shared<vofod::pc_t>();
  cloud->points.resize(n);
  for (size_t i = 0; i < n; i++)
  {
    vofod::pt_t& q = cloud->points[i];
    q.x = scan[i].x; q.y = scan[i].y; q.z = scan[i].z;
    q.intensity = scan[i].intensity;
    q.range = scan[i].range_mm;
  }
  cloud->width = std::uint32_t(v->m_sensor_hrays);
  cloud->height = std::uint32_t(v->m_sensor_vrays);
  Eigen::Affine3f tf;
  fill_tf(tf, *pose);

  const pc_XYZR_t::Ptr cloud_weighted = v->filterAndTransform(cloud, tf);                                                     // :928
  const std::vector<pcl::PointIndices> clusters_indices = v->clusterCloud(cloud_weighted, v->m_drmgr_ptr->config.ground_points_max_distance);  // :932
  res->n_bg = v->m_voxel_map.nVoxelsOver(v->m_drmgr_ptr->config.voxel_map__thresholds__new_obstacles);  // what :712 will see
  const auto [close_clusters_indices, far_clusters_indices] = v->findCloseFarClusters(cloud_weighted, clusters_indices);     // :936
  v->updateVMaps(cloud_weighted, close_clusters_indices, v->m_drmgr_ptr->config.voxel_map__scores__point, VoFOD::m_vflags_point);      // :946
  v->updateVMaps(cloud_weighted, far_clusters_indices, v->m_drmgr_ptr->config.voxel_map__scores__unknown, VoFOD::m_vflags_unknown);  // :948
  v->m_detection_its++;                                                                                                       // :949
  res->raycast_status = VOFOD_W_PAUSED;
  if (s->do_raycast)
  {
    const int diff = s->raycast_its_diff > 1 ? s->raycast_its_diff : 1;
    g_on_cv_wait = [v, diff]() { v->m_detection_its += diff; };  // the detections that finish while the rays are cast (:1530-1539)
    v->raycast_cloud(cloud, tf);
    g_on_cv_wait = nullptr;
    v->m_detection_its -= diff;
    res->raycast_status = v->m_drmgr_ptr->config.raycast__pause ? VOFOD_W_PAUSED : VOFOD_OK;
  }
  std::vector<cluster_t> clusters;
  std::vector<detection_t> detections;
  if (s->do_classify)
  {
    pcl::shim_moi_gaps().clear();
    clusters = v->classifyClusters(cloud_weighted, far_clusters_indices, tf);       // :960
    v->last_gaps = pcl::shim_moi_gaps();
    detections = v->extractDetections(clusters, tf.translation());                // :962
  }
  res->sep_status = VOFOD_W_PAUSED;
  if (s->do_sepclusters)
  {
    if (!v->local_sized)
    {
      v->local_vmap.resizeAs(v->m_voxel_map);  // :1287
      v->local_sized = true;
    }
    const int diff = s->sep_its_diff > 1 ? s->sep_its_diff : 1;
    g_on_checkpoint = [v, diff](const std::string& name) { if (name == "mutex lock1") v->m_detection_its += diff; };
    v->updateSeparatedBGClusters(v->local_vmap);
    g_on_checkpoint = nullptr;
    v->m_detection_its -= diff;
    res->sep_status = v->m_drmgr_ptr->config.sepclusters__pause ? VOFOD_W_PAUSED : VOFOD_OK;
  }
  // ---- book-keeping for the tests
  v->last_cloud = cloud_weighted;
  v->last_clusters = clusters_indices;
  v->last_in_close.assign(cloud_weighted->size(), 0);
  for (const auto& c : close_clusters_indices)
    for (const int i : c->indices)
      v->last_in_close[size_t(i)] = 1;
  v->last_far = clusters;
  v->last_dets = detections;
  res->n_voxels = std::uint32_t(cloud_weighted->size());
  res->n_clusters = std::uint32_t(clusters_indices.size());
  res->n_close_clusters = std::uint32_t(close_clusters_indices.size());
  res->n_far_clusters = std::uint32_t(far_clusters_indices.size());
  res->n_detections = std::uint32_t(detections.size());
  res->background_pts_sufficient = v->m_background_pts_sufficient;
  res->sure_background_sufficient = v->m_sure_background_sufficient;
  return 0;
}

size_t vn_last_voxels(VoFOD* v, vofod_vox* out, int32_t* labels, uint8_t* in_close, size_t cap)
{
  if (!v->last_cloud)
    return 0;
  const size_t m = v->last_cloud->size();
  if (m > cap)
    return m;
  for (size_t i = 0; i < m; i++)
  {
    const auto& p = v->last_cloud->points[i];
    out[i] = vofod_vox{p.x, p.y, p.z, p.range};
    in_close[i] = v->last_in_close[i];
  }
  for (const auto& c : v->last_clusters)
    for (const int i : c.indices)
      labels[i] = c.indices.front();
  return m;
}
// far clusters in the order the reference classified them
size_t vn_last_clusters(VoFOD* v, vofod_cluster_info* out, size_t cap)
{
  const size_t k = v->last_far.size();
  for (size_t i = 0; i < k && i < cap; i++)
  {
    const vofod::cluster_t& c = v->last_far[i];
    vofod_cluster_info ci;
    std::memset(&ci, 0, sizeof(ci));
    ci.label = c.pc_indices->indices.front();
    ci.n_points = int(c.pc_indices->indices.size());
    ci.cclass = c.cclass == vofod::cluster_class_t::mav ? VOFOD_CLASS_MAV : c.cclass == vofod::cluster_class_t::unknown ? VOFOD_CLASS_UNKNOWN : VOFOD_CLASS_INVALID;
    for (int a = 0; a < 3; a++)
    {
      ci.aabb_min[a] = c.aabb.min_pt[a]; ci.aabb_max[a] = c.aabb.max_pt[a];
      ci.obb_min[a] = c.obb.min_pt[a]; ci.obb_max[a] = c.obb.max_pt[a]; ci.obb_center[a] = c.obb.center_pt[a];
      for (int b = 0; b < 3; b++)
        ci.obb_rot[3 * a + b] = c.obb.orientation(a, b);
    }
    ci.obb_size = c.obb_size;
    ci.eig_gap = i < v->last_gaps.size() ? v->last_gaps[i] : 0.f;
    out[i] = ci;
  }
  return k;
}
size_t vn_last_detections(VoFOD* v, vofod_detection* out, size_t cap)
{
  const size_t k = v->last_dets.size();
  for (size_t i = 0; i < k && i < cap; i++)
  {
    const vofod::detection_t& d = v->last_dets[i];
    vofod_detection o;
    std::memset(&o, 0, sizeof(o));
    o.id = d.id;
    o.label = -1;
    o.n_points = d.n_points;
    for (int a = 0; a < 3; a++)
    {
      o.aabb_min[a] = d.aabb.min_pt[a]; o.aabb_max[a] = d.aabb.max_pt[a];
      o.obb_min[a] = d.obb.min_pt[a]; o.obb_max[a] = d.obb.max_pt[a]; o.position[a] = d.obb.center_pt[a];
      for (int b = 0; b < 3; b++)
      {
        o.obb_rot[3 * a + b] = d.obb.orientation(a, b);
        o.covariance[3 * a + b] = d.covariance(a, b);
      }
    }
    o.confidence = d.confidence;
    o.detection_probability = d.detection_probability;
    out[i] = o;
  }
  return k;
}

// load_cloud (src/pc_loader.cpp:17-90): returns the number of points, or -1 when the reference returns nullptr
long vn_load_cloud(const char* filename, float* xyz, size_t cap)
{
  const ::pc_t::Ptr c = load_cloud(filename);
  if (c == nullptr)
    return -1;
  for (size_t i = 0; i < c->size() && i < cap; i++)
  {
    xyz[3 * i] = c->points[i].x; xyz[3 * i + 1] = c->points[i].y; xyz[3 * i + 2] = c->points[i].z;
  }
  return long(c->size());
}
// initialize_apriori_map (:305-353): load, transform, pcl::VoxelGrid centroid down-sample, +inf stamping
void vn_apriori(VoFOD* v, const char* filename, const vofod_pose* pose)
{
  Eigen::Affine3f tf;
  fill_tf(tf, *pose);
  v->initialize_apriori_map(filename, tf);
}
}  // extern "C"
