/*
 * libvofod_cuda — C ABI of the B200-native VoFOD per-scan volumetric hot path.
 *
 * The reference (ctu-mrs/vofod) has no FFI around this path: the nodelet holds three
 * vofod::VoxelMap members by value (src/vofod_nodelet.cpp:2333-2339) and calls class methods and
 * PCL templates inline.  This header is the boundary a maintainer binds instead: every entry point
 * cites the reference code it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns int: 0 = VOFOD_OK, >0 = informational status (the reference logs and
 *     carries on), <0 = error; vofod_last_error() gives a human string for the last error.
 *   - the caller owns all host buffers and passes capacities; the library owns all device memory.
 *   - no exceptions, no abort, no stdout/stderr output cross the ABI.
 *   - one vofod_ctx = one CUDA device; calls on one ctx must be serialised by the caller exactly
 *     as m_voxels_mtx does in the reference (src/vofod_nodelet.cpp:608,712,943,1146,1210,1530,1612);
 *     distinct contexts are independent.
 *   - there is NO CPU fallback: if no CUDA device is usable vofod_create fails with VOFOD_E_CUDA.
 */
#ifndef VOFOD_CUDA_H
#define VOFOD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define VOFOD_OK              0
#define VOFOD_W_SENSOR_OOB    1  /* sensor origin outside the map: raycast skipped (vofod_nodelet.cpp:1432,1525) */
#define VOFOD_W_EMPTY_RAYCAST 2  /* max raycast value is zero: apply + flag clear skipped (:1544-1548)          */
#define VOFOD_W_PAUSED        3  /* raycast__pause / sepclusters__pause set (:1400,1128)                        */
#define VOFOD_W_EMPTY         4  /* nothing to do (e.g. empty voxel-map cloud, :1155-1159)                      */
#define VOFOD_E_INVALID      -1  /* bad argument                                                                */
#define VOFOD_E_CUDA         -2  /* CUDA runtime failure (string in vofod_last_error)                           */
#define VOFOD_E_CAPACITY     -3  /* output buffer too small; the needed count is written to the out-param       */
#define VOFOD_E_STATE        -4  /* map not sized / sensor not set                                              */
#define VOFOD_E_DIMS         -5  /* cloud size != sensor LUT size (:895-899, :1407-1411)                        */
#define VOFOD_E_OVERFLOW     -6  /* voxel-grid index overflow (voxel_grid_weighted.cpp:61-69)                   */
#define VOFOD_E_NOMEM        -7
#define VOFOD_E_INTERNAL     -8  /* device-side watchdog tripped (bounded spin exceeded)                        */
#define VOFOD_E_IO           -9  /* file could not be opened (load_cloud returns nullptr, pc_loader.cpp:22-27)  */

/* which grid (vofod_nodelet.cpp:2333-2339) */
#define VOFOD_MAP_SCORE   0      /* m_voxel_map     : fp32 score                                   */
#define VOFOD_MAP_FLAGS   1      /* m_voxel_flags   : 0 unmarked, 2 background point, 3 unknown    */
#define VOFOD_MAP_RAYCAST 2      /* m_voxel_raycast : accumulated path length [m] of the last scan */

/* cluster classes (vofod_nodelet.cpp:85-90) */
#define VOFOD_CLASS_MAV     0
#define VOFOD_CLASS_UNKNOWN 1
#define VOFOD_CLASS_INVALID 2

typedef struct vofod_ctx vofod_ctx;

/* packed scan point: the fields of ouster_ros::Point the path reads (x,y,z,intensity,range) */
typedef struct vofod_pt { float x, y, z, intensity; uint32_t range_mm; } vofod_pt;
/* payload of vofod::PointXYZR (include/vofod/point_types.h:51-56) */
typedef struct vofod_vox { float x, y, z; uint32_t count; } vofod_vox;
/* payload of pcl::PointXYZI as produced by VoxelMap::voxelsAs[Voxel]PC (voxel_map.cpp:157-212) */
typedef struct vofod_xyzi { float x, y, z, intensity; } vofod_xyzi;
/* rigid sensor->world transform: R row-major (used both as Affine linear part, vofod_nodelet.cpp:640,
 * and as tf.rotation(), :1428), t translation */
typedef struct vofod_pose { float R[9]; float t[3]; } vofod_pose;

/* Every tunable of config/detection_params.yaml + DetectionParams.cfg, passed PER CALL so that
 * dynamic_reconfigure semantics (values re-read on every use) are preserved.
 * double where the reference holds a dynamic_reconfigure double, float where it holds a float member. */
typedef struct vofod_params {
  double ground_points_max_distance;          /* 1.5   */
  double output_position_sigma;               /* 0.1   */
  double score_point;                         /* 0     voxel_map__scores__point   */
  double score_unknown;                       /* -740  voxel_map__scores__unknown */
  double score_ray;                           /* -1000 voxel_map__scores__ray     */
  double thr_apriori_map;                     /* 0     */
  double thr_sure_obstacles;                  /* -0.1  */
  double thr_new_obstacles;                   /* -300  */
  double thr_frontiers;                       /* -750  */
  double cls_max_size;                        /* 3.0   */
  double cls_max_distance;                    /* 50    */
  double cls_max_explore_distance;            /* 3.0   */
  double raycast_max_distance;                /* 20    */
  double raycast_min_intensity;               /* 0     */
  double raycast_weight_coefficient;          /* 0.003 */
  double sep_max_bg_distance;                 /* 0.8   */
  int32_t cls_min_points;                     /* 2     */
  int32_t raycast_pause;                      /* 0     */
  int32_t raycast_new_update_rule;            /* 1     */
  int32_t sep_pause;                          /* 0     */
  int32_t sep_min_sure_points;                /* 24    */
  int32_t _pad0;
  float score_init;                           /* -740  voxel_map/scores/init (static)     */
  float background_sufficient_points_ratio;   /* 0.15  (static)                            */
  float exclude_box_offset[3];                /* yaml values; z is the LOWEST face (:204)  */
  float exclude_box_size[3];
  float oparea_offset[3];                     /* yaml values; z is the LOWEST face (:212)  */
  float oparea_size[3];
  float sensor_vfov;                          /* radians (:869)                            */
  float _pad1;
} vofod_params;

typedef struct vofod_map_info {
  float offset[3];        /* corner of voxel (0,0,0)  (voxel_map.cpp:351-354)       */
  int32_t sizes[3];       /* cells per axis = ceil(dims/vs)+1 (voxel_map.cpp:16)    */
  float voxel_size;
  uint64_t n_cells;
  int32_t slab_axis, slab_lo, slab_hi;  /* owned index range along slab_axis (whole axis when unsharded) */
  int32_t storage_lo[3], storage_size[3]; /* box of global cells this context holds (= 0 / sizes when unsharded): the layout of
                                             vofod_map_download / _upload is x + y*storage_size[0] + z*storage_size[0]*storage_size[1] */
  int32_t _pad;
} vofod_map_info;

/* one entry per far cluster, in classification order (vofod_nodelet.cpp:110-119) */
typedef struct vofod_cluster_info {
  int32_t label;            /* canonical label = minimum point index of the cluster */
  int32_t n_points;
  int32_t cclass;           /* VOFOD_CLASS_*                                         */
  float aabb_min[3], aabb_max[3];
  float obb_min[3], obb_max[3], obb_center[3];
  float obb_rot[9];         /* row-major, columns = major/middle/minor axes          */
  float obb_size;           /* NaN when a gate returned before computing it (:115)   */
  float eig_gap;            /* min relative gap between covariance eigenvalues (parity-test aid) */
} vofod_cluster_info;

/* msgs/Detection.msg:1-12 + detection_t (vofod_nodelet.cpp:121-130) */
typedef struct vofod_detection {
  int32_t id;
  int32_t label;            /* canonical label of the source cluster */
  uint64_t n_points;
  float aabb_min[3], aabb_max[3];
  float position[3];        /* = obb.center (:979-981) */
  float obb_min[3], obb_max[3];
  float obb_rot[9];
  float covariance[9];
  double confidence;
  double detection_probability;
} vofod_detection;

/* which optional stages vofod_process_scan runs (deterministic schedule S1) */
typedef struct vofod_schedule {
  int32_t n_range_seeds;        /* rangefinder seeds before the scan (A23), world_pt = range_pt */
  float   range_pt[3];
  int32_t do_raycast;           /* accumulate + apply(its_diff) + flags.clear()                 */
  int32_t raycast_its_diff;     /* >= 1                                                         */
  int32_t do_classify;          /* classifyClusters + extractDetections                         */
  int32_t do_sepclusters;       /* updateSeparatedBGClusters(its_diff) after the detections     */
  int32_t sep_its_diff;
  /* The reference's steady state (vofod_nodelet.cpp:950-957, 1530-1539, 1602): the raycast thread accumulates scan k while the scan
   * thread carries on, and applies it only after scan k + 1 has done its point update — with the flags of BOTH scans still set — and
   * no new raycast starts while one is in flight.  Reproduce it with
   *   scan k    : do_raycast = 1, raycast_defer_apply = 1      (accumulate, leave the result pending)
   *   scan k + 1: do_raycast = 0, raycast_apply_pending = 1    (after this scan's point update: apply(raycast_its_diff) + flags.clear()) */
  int32_t raycast_defer_apply;
  int32_t raycast_apply_pending;
  /* Software pipelining of the background thread (bgclusters_loop, :1280-1294, runs beside the scan thread in the reference): with
   * sep_deferred = 1 the separated-background pass this scan asks for (do_sepclusters) is carried out at the START of the next
   * vofod_process_scan[_resident] call — before that scan touches the map, so the order of all map operations is still schedule S1's —
   * where it overlaps the next scan's map-independent front end (crop, voxel grid, clustering).  The sep_status / sure_background_sufficient
   * fields of a result then describe the pass that ran in that call (the previous scan's).  Any other entry point that reads or writes
   * the map, and vofod_flush, carry out a pending pass first. */
  int32_t sep_deferred;
} vofod_schedule;

typedef struct vofod_scan_result {
  uint64_t n_traversals;        /* forEachRay callbacks of this scan (voxel_map.cpp:253)  */
  uint64_t n_bg;                /* nVoxelsOver(thr_new_obstacles) seen by findCloseFarClusters */
  uint32_t n_filtered;          /* points surviving both crop boxes                        */
  uint32_t n_voxels;            /* M = cloud_weighted->size()                              */
  uint32_t n_clusters;
  uint32_t n_close_clusters;
  uint32_t n_far_clusters;
  uint32_t n_detections;
  int32_t  background_pts_sufficient;
  int32_t  sure_background_sufficient;
  int32_t  raycast_status;      /* VOFOD_OK / VOFOD_W_*                                     */
  int32_t  sep_status;
} vofod_scan_result;

/* ---- lifetime ------------------------------------------------------------------------------ */
int vofod_create(int device, vofod_ctx** out);
int vofod_destroy(vofod_ctx* ctx);
const char* vofod_last_error(const vofod_ctx* ctx);      /* never NULL; ctx may be NULL */
int vofod_synchronize(vofod_ctx* ctx);
void vofod_default_params(vofod_params* p);               /* config/detection_params.yaml verbatim */
/* VoFOD::reset() (vofod_nodelet.cpp:1610-1632): resize the 3 grids from the operation area, score=init,
 * flags=0, detection_its=0; also clears background flags and the detection id counter. */
int vofod_reset(vofod_ctx* ctx, const vofod_params* p, float voxel_size);

/* ---- C1: vofod::VoxelMap (include/vofod/voxel_map.h:27-112, src/voxel_map.cpp) ------------- */
int vofod_map_resize(vofod_ctx*, const float center[3], const float dims[3], float voxel_size);       /* voxel_map.cpp:11-19 */
int vofod_map_resize_idx(vofod_ctx*, const float offset[3], const int32_t sizes[3], float voxel_size); /* :21-48 */
int vofod_map_info_get(const vofod_ctx*, vofod_map_info* out);                                         /* :343-374 */
int vofod_map_set_to(vofod_ctx*, int which, float value);                                              /* :268-279 */
int vofod_map_set_inf(vofod_ctx*, const float* xyz, size_t n);            /* apriori voxels, vofod_nodelet.cpp:339-341 */
int vofod_map_download(vofod_ctx*, int which, float* host, size_t n_cells);    /* begin()/end() mirror (voxel_map.h:50-51) */
int vofod_map_upload(vofod_ctx*, int which, const float* host, size_t n_cells);
int vofod_map_get(vofod_ctx*, int which, int ix, int iy, int iz, float* value);  /* atIdx (voxel_map.cpp:105-133) */
int vofod_map_set(vofod_ctx*, int which, int ix, int iy, int iz, float value);
int vofod_map_count_over(vofod_ctx*, float threshold, uint64_t* out);            /* nVoxelsOver :216-222 */
/* voxelsAsPC (metric=1) / voxelsAsVoxelPC (metric=0): emission order x-outer, y, z-inner (:157-212) */
int vofod_map_compact_over(vofod_ctx*, float threshold, int greater_than, int metric,
                           vofod_xyzi* out, size_t cap, size_t* n);
int vofod_map_has_close_to(vofod_ctx*, const float* xyz, size_t n, float max_dist, float threshold,
                           uint8_t* out);                                         /* :376-400 */
int vofod_map_explore_to_ground(vofod_ctx*, const float pt[3], float unknown_threshold,
                                float ground_threshold, float max_voxel_dist, int* connected,
                                int32_t* explored_idx3, size_t cap, size_t* n_explored); /* :402-488 */
int vofod_map_is_floating(vofod_ctx*, const float* xyz, size_t n, float threshold, uint8_t* out); /* :491-516 */
int vofod_map_submap_copy(vofod_ctx*, const float min_pt[3], const float max_pt[3], int inflate,
                          float* out, size_t cap, int32_t sizes_out[3], float offset_out[3]); /* :547-584 */
/* forEachRay (:229-263): returns the callback sequence (ddist, ix,iy,iz) of ONE ray */
int vofod_map_trace_ray(vofod_ctx*, const float start[3], const float dir[3], float length,
                        float* ddist, int32_t* idx3, size_t cap, size_t* n);

/* ---- apriori map ingest (vofod_nodelet.cpp:305-353, src/pc_loader.cpp:17-90) ------------------- */
/* load_cloud: text cloud (one "x y z ..." per line; a ".pts" file has the point count on its first line) -> xyz triples.  Host only.
 * *n = points in the file (up to cap are written).  VOFOD_E_IO: cannot open (the reference returns nullptr); VOFOD_E_CAPACITY: call
 * again with cap >= *n. */
int vofod_load_cloud(const char* filename, float* xyz, size_t cap, size_t* n);
/* initialize_apriori_map from a loaded cloud: transformPointCloud(tf) -> pcl::VoxelGrid centroid down-sample at the map's voxel size
 * -> voxels of the in-map centroids := +inf -> both background latches set (:343-344).  centroids (optional, cap points) receives the
 * down-sampled cloud the reference publishes (:347-352), *n_voxels its size. */
int vofod_apriori_map(vofod_ctx*, const float* xyz, size_t n, const vofod_pose* tf, float* centroids, size_t cap, size_t* n_voxels);

/* ---- sensor model (vofod_nodelet.cpp:77-81, 374-420, 506-560) ------------------------------- */
/* dirs/offs: 3xN column-major as Eigen stores them (= N consecutive xyz triples), ray id = row*W+col;
 * mask: N bytes, nonzero = valid pixel (NULL = all ones, :558) */
int vofod_set_sensor(vofod_ctx*, int W, int H, const float* dirs3xN, const float* offs3xN,
                     const uint8_t* mask);

/* host-only helpers of the sensor model (no device needed) */
/* load_mask (vofod_nodelet.cpp:506-560) on a decoded rows x cols u8 image (NULL = file missing): the "mangle" re-ordering
 * out[((v + pixel_shift_by_row[u]) % W) * H + u] = img[u * W + v] (:537-548), the plain copy, or all ones on missing / wrong-size input */
int vofod_mask_mangle(const uint8_t* img, int cols, int rows, int W, int H, int mangle, const int32_t* pixel_shift_by_row, uint8_t* out);
/* initialize_sensor_lut (:358-371): ouster::make_xyz_lut, cast to float, directions normalised; transform4x4 row-major (NULL = identity) */
int vofod_make_xyz_lut(int w, int h, double range_unit, double lidar_origin_to_beam_origin_mm, const double* transform4x4,
                       const double* azimuth_angles_deg, const double* altitude_angles_deg, float* dirs3xN, float* offs3xN);
/* initialize_sensor_lut_simulation (:374-420); offs may be NULL */
int vofod_sim_xyz_lut(int w, int h, float vfov, float* dirs3xN, float* offs3xN);
/* 48-byte ouster_ros::Point records (include/vofod/types.h:7) -> packed vofod_pt; stride 0 = that layout */
int vofod_pack_ouster(const void* points, size_t n, size_t stride, size_t intensity_offset, size_t range_offset, vofod_pt* out);

/* ---- C2/C3: voxel grids ---------------------------------------------------------------------- */
/* filterAndTransform (vofod_nodelet.cpp:621-684): exclude-box crop, rigid transform, op-area crop,
 * VoxelGridWeighted aligned to the map.  out: ascending voxel key order (voxel_grid_weighted.cpp:143). */
int vofod_filter_voxelize(vofod_ctx*, const vofod_pt* scan, size_t n, const vofod_pose*,
                          const vofod_params*, vofod_vox* out, size_t cap, size_t* m);
/* VoxelGridWeighted::filter on an arbitrary cloud (voxel_grid_weighted.cpp:41-190); align may be NULL */
int vofod_voxel_grid_weighted(vofod_ctx*, const float* xyz, size_t n, float leaf, const float align[3],
                              vofod_vox* out, size_t cap, size_t* m);
/* VoxelGridCounted::filter (voxel_grid_counted.cpp:49-196), including its input-slice quirk (:185-187) */
int vofod_voxel_grid_counted(vofod_ctx*, const vofod_xyzi* pts, size_t n, float leaf, float threshold,
                             const float align[3], vofod_vox* out, size_t cap, size_t* m);

/* ---- A13: clusterCloud (vofod_nodelet.cpp:689-698; pcl::EuclideanClusterExtraction) ---------- */
/* labels[i] = minimum point index of i's connected component under strict d^2 < tol^2 */
int vofod_cluster(vofod_ctx*, const float* xyz, size_t m, float tol, int32_t* labels, size_t* n_clusters);
/* Host-only diagnostic (no device needed): the table the scan's own clustering works from.  Its points are centres of grid
 * leaves, so "within the tolerance" is a property of the index offset (dx, dy, dz): for every forward row (dy, dz) listed,
 * |dx| <= R is inside for sure, |dx| = R + 1 is decided by the fp32 distance when shell == 1 (exact squared distance == tol^2),
 * and shell == 2 (R = -1) says that dx = 0 itself is such a border case; rows not listed are outside.  Returns the number of
 * rows, or 0 when the neighbourhood needs more than 16 rows (the scan then clusters with the generic vofod_cluster path). */
int vofod_cluster_grid_rows(float tolerance, float leaf, int8_t* dy, int8_t* dz, int8_t* R, int8_t* shell, int cap);

/* ---- A14/A15: findCloseFarClusters (vofod_nodelet.cpp:703-750) -------------------------------- */
int vofod_close_far(vofod_ctx*, const vofod_vox* pts, const int32_t* labels, size_t m,
                    const vofod_params*, uint8_t* point_in_close_cluster, uint64_t* n_bg);

/* ---- A23: rangefinder ground seed (vofod_nodelet.cpp:581-613) --------------------------------- */
int vofod_range_update(vofod_ctx*, const float world_pt[3], const vofod_params*);

/* ---- A11: updateVMaps (vofod_nodelet.cpp:777-815); sel may be NULL (all) else byte mask --------- */
int vofod_update_points(vofod_ctx*, const vofod_vox* pts, const uint8_t* sel, int sel_value, size_t n,
                        float score, float flag);

/* ---- A3..A9: raycast_cloud (vofod_nodelet.cpp:1397-1606) -------------------------------------- */
int vofod_raycast_accumulate(vofod_ctx*, const vofod_pt* scan, size_t n, const vofod_pose*,
                             const vofod_params*, uint64_t* n_traversals);
/* parity/debug view of the accumulator before apply: per-cell callback count and path length.
 * either pointer may be NULL. */
int vofod_raycast_download(vofod_ctx*, uint32_t* counts, float* lengths, size_t n_cells);
int vofod_raycast_apply(vofod_ctx*, int its_diff, const vofod_params*);
/* fractional bits of the fixed-point path-length accumulator chosen for the current sensor / voxel size */
int vofod_raycast_frac_bits(const vofod_ctx*);

/* ---- A16..A19: classifyClusters + extractDetections (vofod_nodelet.cpp:819-879, 1648-1731) ---- */
int vofod_classify_detect(vofod_ctx*, const vofod_vox* pts, const int32_t* labels,
                          const uint8_t* point_in_close_cluster, size_t m, const vofod_pose*,
                          const vofod_params*, vofod_detection* dets, size_t det_cap, size_t* n_dets,
                          vofod_cluster_info* clusters, size_t cl_cap, size_t* n_far_clusters);

/* ---- C6: updateSeparatedBGClusters (vofod_nodelet.cpp:1126-1278) ------------------------------ */
int vofod_sepclusters(vofod_ctx*, int its_diff, const vofod_params*, int* sure_background_sufficient);

/* ---- nodelet state that lives beside the maps (vofod_nodelet.cpp:2323-2332) -------------------- */
int vofod_state_get(const vofod_ctx*, int* background_pts_sufficient, int* sure_background_sufficient,
                    uint32_t* last_detection_id);
int vofod_state_set(vofod_ctx*, int background_pts_sufficient, int sure_background_sufficient,
                    uint32_t last_detection_id);

/* ---- whole scan, device-resident, no host round trip between stages (schedule S1) ------------- */
int vofod_process_scan(vofod_ctx*, const vofod_pt* scan, size_t n, const vofod_pose*,
                       const vofod_params*, const vofod_schedule*, vofod_scan_result* res,
                       vofod_detection* dets, size_t det_cap);
/* announce the NEXT scan: its host->device copy runs on a copy stream next to the current scan's kernels; the following
 * vofod_process_scan (or vofod_slab_process_scan on rank 0) with the same `scan` pointer consumes it.  Keep the host buffer (ideally pinned)
 * untouched until then.  An announcement is matched by host pointer and is good for the next TWO scan calls only (announce k+1, process k,
 * process k+1): one that was not consumed by then is void, so a buffer that is refilled later is copied again. */
int vofod_prefetch_scan(vofod_ctx*, const vofod_pt* scan, size_t n);
/* carries out a pending deferred separated-background pass (vofod_schedule::sep_deferred); no-op when there is none */
int vofod_flush(vofod_ctx*);
/* a sequence of scans from host buffers back to back (replay, benchmarks): the copy of scan k + 1 overlaps the kernels of scan k and the
 * host spends microseconds between two scans.  poses / scheds / results: one per scan; dets: det_cap records per scan (may be NULL). */
int vofod_process_scan_batch(vofod_ctx*, const vofod_pt* const* scans, size_t n_scans, size_t n, const vofod_pose* poses,
                             const vofod_params*, const vofod_schedule* scheds, vofod_scan_result* results,
                             vofod_detection* dets, size_t det_cap, uint32_t* n_dets, size_t* n_done);
/* same, but the scan is already in device memory (bench "value" leg: inputs resident in HBM) */
int vofod_upload_scan(vofod_ctx*, int slot, const vofod_pt* scan, size_t n);
int vofod_process_scan_resident(vofod_ctx*, int slot, const vofod_pose*, const vofod_params*,
                                const vofod_schedule*, vofod_scan_result* res,
                                vofod_detection* dets, size_t det_cap);
/* outputs of the last process_scan kept on the device, fetched on demand (debug topics) */
int vofod_last_voxels(vofod_ctx*, vofod_vox* out, int32_t* labels, uint8_t* in_close, size_t cap, size_t* m);
int vofod_last_clusters(vofod_ctx*, vofod_cluster_info* out, size_t cap, size_t* n);

/* ---- multi-GPU slab mode (no counterpart in the reference; SURVEY.md §8e, BASELINE.json configs[4]) -------------------------------- */
/* The global grid is cut along a horizontal axis (0 = x or 1 = y), one slab per context / GPU: this context keeps only cells
 * lo-halo <= idx[axis] < hi+halo and OWNS lo <= idx[axis] < hi.  Call after vofod_map_resize / vofod_reset; the grid contents are
 * unspecified afterwards (vofod_map_set_to).  Map downloads / uploads then address the storage box (vofod_map_info_get reports the
 * global geometry and the own range).  halo >= vofod_slab_min_halo(params, voxel_size) keeps every result bit-identical to the
 * unsharded run (checked when a scan starts: VOFOD_E_INVALID otherwise). */
int vofod_set_slab(vofod_ctx*, int axis, int lo, int hi, int halo);
int vofod_slab_min_halo(const vofod_params*, float voxel_size);
/* A scan of schedule S1 in slab mode = 4 phases; after each, some device buffers are combined over all slabs (vofod_csrc/slab.cu says
 * which and why): rays are clipped to the slab, grids and grid passes are sharded, the small per-scan point pipeline is replicated. */
#define VOFOD_XCHG_SUM_U64    0   /* all-reduce, sum of uint64                                          */
#define VOFOD_XCHG_MAX_I32    1   /* all-reduce, max of int32                                           */
#define VOFOD_XCHG_SUM_U32    2   /* all-reduce, sum of uint32 (every element is non-zero on one slab)  */
#define VOFOD_XCHG_GATHER_U32 3   /* all-gather: `count` uint32 of every slab into gather_out, rank-major */
#define VOFOD_W_REDO          5   /* vofod_slab_phase(3): a gathered list overflowed, nothing was touched: repeat phases 2 and 3 */
typedef struct vofod_slab_exchange {
  int32_t kind;        /* VOFOD_XCHG_*                                   */
  int32_t _pad;
  void*   buf;         /* device memory; all-reduces are done in place   */
  size_t  count;       /* elements                                       */
  void*   gather_out;  /* VOFOD_XCHG_GATHER_U32 only: nranks * count elements */
} vofod_slab_exchange;
/* (a) caller-driven: phase 0 takes the scan (host pointer, or device pointer with scan_on_device != 0) and the arguments, phases 1..3
 * continue it and ignore them; after every phase combine the buffers vofod_slab_exchanges lists (at most 4) over all slabs — a host
 * loop when several slabs live on one device (tests), any collective library otherwise; phase 3 delivers the results.
 * vofod_slab_set_world tells a context how many slabs there are (it sizes the gather buffers). */
int vofod_slab_set_world(vofod_ctx*, int rank, int nranks);
int vofod_slab_phase(vofod_ctx*, int phase, const vofod_pt* scan, int scan_on_device, size_t n, const vofod_pose*, const vofod_params*,
                     const vofod_schedule*, vofod_scan_result* res, vofod_detection* dets, size_t det_cap);
int vofod_slab_exchanges(vofod_ctx*, int phase, vofod_slab_exchange out[4], int* n_out);
/* (b) over NCCL inside the library (libnccl.so.2 is resolved at run time): one communicator per context, created from an id that rank 0
 * makes with vofod_comm_unique_id (128 bytes) and hands to the other ranks.  vofod_slab_process_scan then runs the four phases with
 * ncclBroadcast of the packed scan (rank 0 passes it, in host memory; other ranks may pass NULL) and ncclAllReduce / ncclAllGather of
 * the exchange buffers on the context's stream; pose, parameters and schedule are given on every rank.  One host wait per scan. */
int vofod_comm_unique_id(void* out128);
int vofod_comm_init(vofod_ctx*, int rank, int nranks, const void* nccl_unique_id);
int vofod_slab_process_scan(vofod_ctx*, const vofod_pt* scan, size_t n, const vofod_pose*, const vofod_params*, const vofod_schedule*,
                            vofod_scan_result* res, vofod_detection* dets, size_t det_cap);
/* device time [ms] of the parts of the last vofod_slab_process_scan on this rank: [0] scan copy + broadcast, [1],[3],[5] the kernels of phases
 * 0..2, [2],[4],[6] the exchange after each (includes waiting for the slowest slab), [7] phase 3 with its read-back */
int vofod_slab_times(vofod_ctx*, float ms[8]);
/* voxels of the last scan within `margin` cells of a face of the owned range, with their cluster labels (the cluster fragments that
 * reach into the neighbouring slab) */
int vofod_slab_boundary(vofod_ctx*, int margin, int32_t* point_idx, int32_t* labels, size_t cap, size_t* n);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* device time [ms] of the stages of the last process_scan, names follow the reference's ScopeTimer
 * checkpoints (vofod_nodelet.cpp:924-964, 1527-1604, 1147-1274) */
#define VOFOD_N_STAGES 12
int vofod_stage_times(vofod_ctx*, float ms[VOFOD_N_STAGES]);
const char* vofod_stage_name(int i);
/* number of kernels the library launched since the context was created */
uint64_t vofod_kernel_launches(const vofod_ctx*);
/* options: VOFOD_OPT_GRAPH (default 1) replays vofod_process_scan[_resident] as a CUDA graph once its launch sequence has
 * been seen twice unchanged; with 0 every scan is enqueued kernel by kernel and vofod_stage_times is filled per stage */
#define VOFOD_OPT_GRAPH 1
#define VOFOD_OPT_RAYCAST_BLOCK 5  /* tuning: rays per thread block of the raycast accumulate kernel: 64 (default), 128 or 256 */
#define VOFOD_OPT_OVERLAP 4        /* default 1: independent stages run as parallel branches of the scan graph */
#define VOFOD_OPT_SEP_GENERAL 3    /* test switch (default 0): sepclusters never takes its leaf-size-1 fast path */
#define VOFOD_OPT_SEP_CAP 9        /* test switch (default 0 = automatic): fixed capacity of the background-voxel list of the capture-friendly sepclusters
                                      pass; a list that overflows leaves the map untouched and the pass is redone exactly */
#define VOFOD_OPT_CLUSTER_HASH 8   /* test switch (default 0): the scan path clusters with the generic spatial-hash clustering */
#define VOFOD_OPT_VG_SORT 7        /* test switch (default 0): the scan path voxelizes with the generic sort-based voxel grid */
#define VOFOD_OPT_PDL 6            /* default 1: consecutive kernels are chained by programmatic dependent launch */
#define VOFOD_OPT_RAYCAST_NO_AGG 2 /* tuning switch (default 0): one RED per traversal instead of warp-aggregated REDs */
#define VOFOD_OPT_SLAB_PATCH_WORDS 11 /* slab mode: capacity (uint32 words) of the buffer that carries the classification candidates' map boxes (0 = automatic) */
#define VOFOD_OPT_ACC_SPARSE 12   /* raycast accumulator "touched" marks (apply visits only groups of 32 cells the rays added to): 0 automatic (windows of 2^25 cells and
                                      more, i.e. long rays on a fine grid), 1 always, 2 never */
#define VOFOD_OPT_RAYCAST_EXP 13   /* MEASUREMENT ONLY (the raycast results are wrong while it is set): 1 = the accumulate kernel does everything but the RED itself,
                                      2 = the DDA alone (no match / redux / RED): what the instruction stream costs without the memory side */
#define VOFOD_OPT_RAYCAST_SPREAD 14 /* tuning (default 64): the accumulate kernel merges the updates of a warp's lanes up to this many voxel sizes along a ray; beyond it
                                     neighbouring rays stand in different voxels anyway and every lane adds its own value */
#define VOFOD_OPT_CLASSIFY_SEQ 15   /* test switch (default 0): exploreToGround / detection extraction in one thread block, cluster after cluster (round 1's kernel)
                                     instead of in parallel over the far clusters whose explore boxes do not meet */
#define VOFOD_OPT_RAYCAST_STATS 10 /* instrumentation switch (default 0): the accumulate kernel also fills per warp-step histograms, see vofod_raycast_stats */
int vofod_set_option(vofod_ctx*, int option, int value);
/* VOFOD_OPT_RAYCAST_STATS: out[0..32] = warp-steps with that many lanes (rays) in the loop, out[33..65] = warp-steps with that many distinct
 * voxels (= REDs issued), out[66] / out[67] = warp-steps of the fast / general loop, out[68] = DDA steps skipped by the slab fast-forward;
 * accumulated since the last call (the call resets them).  n <= 72. */
int vofod_raycast_stats(vofod_ctx*, uint64_t* out, size_t n);
/* scan-replay statistics: which = 0 graph replays, 1 captures, 2 failed captures, 3 kernel-by-kernel scans, 4 reason code of the last failed capture */
uint64_t vofod_get_stat(const vofod_ctx*, int which);
/* raw CUDA stream handle (cudaStream_t) the context enqueues on, for event timing by the caller */
void* vofod_stream(vofod_ctx*);
/* stream-ordered, synchronous copies between host memory and a device buffer the library handed out (vofod_slab_exchanges) */
int vofod_dev_read(vofod_ctx*, const void* dev, void* host, size_t bytes);
int vofod_dev_write(vofod_ctx*, void* dev, const void* host, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* VOFOD_CUDA_H */
