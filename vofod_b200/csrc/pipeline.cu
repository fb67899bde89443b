// Per-scan orchestration of libvofod_cuda: the L3 functions of the nodelet (vofod_nodelet.cpp:581-613, 703-815,
// 882-964) as device stages, and vofod_process_scan, which enqueues a whole scan of the deterministic schedule S1
// on the context's stream with every count kept in DEVICE memory (no host round trip between stages) and reads
// the results back once at the end.
#include <math.h>

#include "common.cuh"
#include "prims.cuh"

// ---- A23 rangefinder ground seed (vofod_nodelet.cpp:581-613) ----------------------------------------
__global__ void k_range_update(float* __restrict__ score, const Geom g, const float x, const float y, const float z, const double score_point, const int repeats)
{
  const int ix = coord_to_idx1(x, g.off[0], g.inv), iy = coord_to_idx1(y, g.off[1], g.inv), iz = coord_to_idx1(z, g.off[2], g.inv);
  if (!in_limits_idx(g, ix, iy, iz))  // :599
    return;
  const long long ci = cell_index(g, ix, iy, iz);
  if (ci < 0)
    return;
  float m = score[ci];
  for (int r = 0; r < repeats; r++)
    m = (float)(((double)m + score_point) / 2.0);  // :610
  score[ci] = m;
}

int vf_range_update_dev(vofod_ctx* ctx, const float pt[3], const vofod_params& p, int repeats)
{
  if (repeats <= 0)
    return 0;
  LAUNCH(k_range_update, 1, 1, 0, ctx->score.as<float>(), ctx->g, pt[0], pt[1], pt[2], p.score_point, repeats);
  return 0;
}

// ---- A11 updateVoxel / updateVMaps (vofod_nodelet.cpp:777-809) ---------------------------------------
__global__ void __launch_bounds__(256) k_update_points(float* __restrict__ score, uint8_t* __restrict__ flags, const Geom g, const vofod_vox* __restrict__ vox,
                                                       const uint8_t* __restrict__ sel, const int sel_value, const unsigned long long* __restrict__ d_m,
                                                       const size_t m_cap, const float vmap_score, const uint8_t vflag, uint32_t* __restrict__ flagged,
                                                       const size_t flagged_cap, unsigned long long* __restrict__ counters)
{
  const size_t m = prims::dev_count(d_m, m_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    if (sel && sel[i] != (uint8_t)sel_value)
      continue;
    const vofod_vox v = vox[i];
    const int xc = coord_to_idx1(v.x, g.off[0], g.inv), yc = coord_to_idx1(v.y, g.off[1], g.inv), zc = coord_to_idx1(v.z, g.off[2], g.inv);
    if (!in_limits_idx(g, xc, yc, zc))  // the reference's vector::at would throw; unreachable after the op-area crop
      continue;
    const long long ci = cell_index(g, xc, yc, zc);
    if (ci < 0 || !cell_owned(g, xc, yc, zc))
      continue;
    const unsigned c = v.count > 63u ? 63u : v.count;                   // std::clamp(pt.range, 0u, 63u)
    const float w = 1.0f / (float)(1ull << c);                          // :791
    score[ci] = w * score[ci] + (1.0f - w) * vmap_score;                // :794
    flags[ci] = vflag;                                                  // :796
    const unsigned long long k = atomicAdd(counters + CNT_FLAGGED, 1ull);
    if (k < flagged_cap)
      flagged[k] = (uint32_t)ci;
    else
      counters[CNT_FLAGGED_OVERFLOW] = 1ull;
  }
}

static int ensure_flagged(vofod_ctx* ctx, size_t want)
{
  if (ctx->flagged_cap >= want)
    return 0;
  // growing would lose the cells recorded so far: fall back to one full clear
  if (ctx->flagged_cap)
    ctx->flags_full_dirty = true;
  ENSURE(ctx->flagged, want * 4);
  ctx->flagged_cap = want;
  return 0;
}

int vf_update_points_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_sel, int sel_value, const unsigned long long* d_m, size_t m_cap, float score, float flag)
{
  if (m_cap == 0)
    return 0;
  RET(ensure_flagged(ctx, 4 * m_cap > (size_t(1) << 20) ? 4 * m_cap : (size_t(1) << 20)));
  LAUNCH(k_update_points, vf_blocks(ctx, m_cap, 256, 8), 256, 0, ctx->score.as<float>(), ctx->flags.as<uint8_t>(), ctx->g, d_vox, d_sel, sel_value, d_m, m_cap, score,
         (uint8_t)flag, ctx->flagged.as<uint32_t>(), ctx->flagged_cap, ctx->d_counters.as<unsigned long long>());
  return 0;
}

// ---- A14/A15 findCloseFarClusters (vofod_nodelet.cpp:703-750) ----------------------------------------
__global__ void k_bg_state(unsigned long long* __restrict__ counters, const unsigned long long min_sufficient)
{
  if (counters[CNT_NBG] > min_sufficient)  // :716-721
    counters[CNT_STATE_BG] = 1ull;
}

// one WARP per point: the lanes stride over the <= (2*mv)^3 window cells of VoxelMap::hasCloseTo (voxel_map.cpp:376-400)
__global__ void __launch_bounds__(256) k_close_points(const float* __restrict__ score, const Geom g, const vofod_vox* __restrict__ vox, const int* __restrict__ labels,
                                                      const unsigned long long* __restrict__ d_m, const size_t m_cap, const float max_dist, const float thr,
                                                      int* __restrict__ cl_close)
{
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const float md = max_dist * g.inv;
  const int mv = (int)ceilf(md);
  for (size_t i = warp0; i < m; i += n_warps)
  {
    const int label = labels[i];
    // a cluster is close iff ANY of its points is (:730-741): once one point has said so the others need not look
    if (((volatile int*)cl_close)[label])
      continue;
    const vofod_vox v = vox[i];
    const int ox = coord_to_idx1(v.x, g.off[0], g.inv), oy = coord_to_idx1(v.y, g.off[1], g.inv), oz = coord_to_idx1(v.z, g.off[2], g.inv);
    const int bx = max(ox - mv, 0), by = max(oy - mv, 0), bz = max(oz - mv, 0);
    const int ex = min(ox + mv, g.size[0]), ey = min(oy + mv, g.size[1]), ez = min(oz + mv, g.size[2]);
    const int nx = ex - bx, ny = ey - by, nz = ez - bz;
    bool hit = false;
    if (nx > 0 && ny > 0 && nz > 0)
    {
      const int total = nx * ny * nz;
      for (int base = 0; base < total && !hit; base += 32)
      {
        const int t = base + (int)lane;
        bool h = false;
        if (t < total)
        {
          const int xi = bx + t % nx, yi = by + (t / nx) % ny, zi = bz + t / (nx * ny);
          const long long ci = cell_index(g, xi, yi, zi);
          if (ci >= 0 && score[ci] > thr)
          {
            const int ddx = xi - ox, ddy = yi - oy, ddz = zi - oz;
            const int nrm = (int)sqrt((double)(ddx * ddx + ddy * ddy + ddz * ddz));  // Eigen int-vector norm(): truncation
            h = (float)nrm <= md;
          }
        }
        hit = __any_sync(VOFOD_FULL, h);
      }
    }
    if (hit && lane == 0)
      cl_close[label] = 1;
  }
}
__global__ void __launch_bounds__(256) k_close_finish(const int* __restrict__ labels, const unsigned long long* __restrict__ d_m, const size_t m_cap,
                                                      const int* __restrict__ cl_close, uint8_t* __restrict__ pt_close, unsigned long long* __restrict__ counters)
{
  const size_t m = prims::dev_count(d_m, m_cap);
  unsigned n_close = 0, n_far = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const int l = labels[i];
    const int c = cl_close[l];
    pt_close[i] = (uint8_t)c;
    if (l == (int)i)
    {
      n_close += c != 0;
      n_far += c == 0;
    }
  }
  n_close = prims::warp_sum(n_close);
  n_far = prims::warp_sum(n_far);
  if ((threadIdx.x & 31) == 0)
  {
    if (n_close)
      atomicAdd(counters + CNT_NCLOSE, (unsigned long long)n_close);
    if (n_far)
      atomicAdd(counters + CNT_NFAR, (unsigned long long)n_far);
  }
}

int vf_close_far_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const unsigned long long* d_m, size_t m_cap, const vofod_params& p)
{
  const float max_dist = (float)p.ground_points_max_distance;
  const float thr = (float)p.thr_new_obstacles;
  RET(vf_count_over_dev(ctx, thr, vf_cnt(ctx, CNT_NBG)));
  // vofod_nodelet.cpp:229-230 (fp32, left to right), converted to uint64_t
  const float vs = ctx->cfg_voxel_size > 0.f ? ctx->cfg_voxel_size : ctx->g.vs;
  volatile float a0 = p.oparea_size[0] / vs;
  volatile float a1 = a0 * p.oparea_size[1];
  volatile float a2 = a1 / vs;
  volatile float a3 = a2 * p.background_sufficient_points_ratio;
  const unsigned long long min_sufficient = (unsigned long long)a3;
  LAUNCH(k_bg_state, 1, 1, 0, ctx->d_counters.as<unsigned long long>(), min_sufficient);
  CK(cudaMemsetAsync(vf_cnt(ctx, CNT_NCLOSE), 0, 16, ctx->stream));  // NCLOSE, NFAR
  if (m_cap == 0)
    return 0;
  ENSURE(ctx->cl_close, m_cap * 4);
  ENSURE(ctx->pt_close, m_cap + 64);
  CK(cudaMemsetAsync(ctx->cl_close.p, 0, m_cap * 4, ctx->stream));
  LAUNCH(k_close_points, vf_blocks(ctx, m_cap * 32, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, d_vox, d_labels, d_m, m_cap, max_dist, thr, ctx->cl_close.as<int>());
  LAUNCH(k_close_finish, vf_blocks(ctx, m_cap, 256, 8), 256, 0, d_labels, d_m, m_cap, ctx->cl_close.as<int>(), ctx->pt_close.as<uint8_t>(),
         ctx->d_counters.as<unsigned long long>());
  return 0;
}

// ======================================================================================================
// host side
// ======================================================================================================
#define NEED_MAP()                                                                        \
  if (!ctx)                                                                               \
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");                             \
  CK(cudaSetDevice(ctx->device));                                                         \
  if (!ctx->map_ready)                                                                    \
  return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized (call vofod_map_resize / vofod_reset first)")

// refresh the host mirror of the device-resident nodelet state; the stream must be idle
static int pull_state(vofod_ctx* ctx)
{
  unsigned long long s[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(s, vf_cnt(ctx, CNT_STATE_BG), 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->background_pts_sufficient = s[0] != 0;
  ctx->sure_background_sufficient = s[1] != 0;
  ctx->last_detection_id = (uint32_t)s[2];
  return 0;
}
extern "C" {

int vofod_state_get(const vofod_ctx* ctx, int* bg, int* sure, uint32_t* id)
{
  if (!ctx)
    return VOFOD_E_INVALID;
  if (bg) *bg = ctx->background_pts_sufficient;
  if (sure) *sure = ctx->sure_background_sufficient;
  if (id) *id = ctx->last_detection_id;
  return VOFOD_OK;
}
int vofod_state_set(vofod_ctx* ctx, int bg, int sure, uint32_t id)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  const unsigned long long s[3] = {bg != 0 ? 1ull : 0ull, sure != 0 ? 1ull : 0ull, (unsigned long long)id};
  CK(cudaMemcpyAsync(vf_cnt(ctx, CNT_STATE_BG), s, 3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->background_pts_sufficient = bg != 0;
  ctx->sure_background_sufficient = sure != 0;
  ctx->last_detection_id = id;
  return VOFOD_OK;
}

int vofod_range_update(vofod_ctx* ctx, const float world_pt[3], const vofod_params* p)
{
  NEED_MAP();
  if (!world_pt || !p)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  RET(vf_range_update_dev(ctx, world_pt, *p, 1));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_update_points(vofod_ctx* ctx, const vofod_vox* pts, const uint8_t* sel, int sel_value, size_t n, float score, float flag)
{
  NEED_MAP();
  if (n == 0)
    return VOFOD_OK;
  if (!pts)
    return vf_fail(ctx, VOFOD_E_INVALID, "pts is NULL");
  ENSURE(ctx->vox, prims::padded(n) * sizeof(vofod_vox));
  CK(cudaMemcpyAsync(ctx->vox.p, pts, n * sizeof(vofod_vox), cudaMemcpyHostToDevice, ctx->stream));
  const uint8_t* d_sel = nullptr;
  if (sel)
  {
    ENSURE(ctx->scratch_b, n + 64);
    CK(cudaMemcpyAsync(ctx->scratch_b.p, sel, n, cudaMemcpyHostToDevice, ctx->stream));
    d_sel = ctx->scratch_b.as<uint8_t>();
  }
  RET(vf_update_points_dev(ctx, ctx->vox.as<vofod_vox>(), d_sel, sel_value, nullptr, n, score, flag));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_close_far(vofod_ctx* ctx, const vofod_vox* pts, const int32_t* labels, size_t m, const vofod_params* p, uint8_t* point_in_close_cluster, uint64_t* n_bg)
{
  NEED_MAP();
  if (!p || (m && (!pts || !labels || !point_in_close_cluster)))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  for (size_t i = 0; i < m; i++)
    if (labels[i] < 0 || (size_t)labels[i] >= m)
      return vf_fail(ctx, VOFOD_E_INVALID, "labels[%zu] = %d is not a point index", i, labels[i]);
  if (m)
  {
    ENSURE(ctx->vox, prims::padded(m) * sizeof(vofod_vox));
    ENSURE(ctx->labels, m * 4);
    CK(cudaMemcpyAsync(ctx->vox.p, pts, m * sizeof(vofod_vox), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->labels.p, labels, m * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  RET(vf_close_far_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), nullptr, m, *p));
  unsigned long long nbg = 0;
  CK(cudaMemcpyAsync(&nbg, vf_cnt(ctx, CNT_NBG), 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (m)
    CK(cudaMemcpyAsync(point_in_close_cluster, ctx->pt_close.p, m, cudaMemcpyDeviceToHost, ctx->stream));
  RET(pull_state(ctx));
  if (n_bg)
    *n_bg = nbg;
  ctx->last_m = m;
  return VOFOD_OK;
}

int vofod_upload_scan(vofod_ctx* ctx, int slot, const vofod_pt* scan, size_t n)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (slot < 0 || slot >= VOFOD_SCAN_SLOTS || !scan || n == 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_upload_scan: bad arguments (%d slots)", VOFOD_SCAN_SLOTS);
  ENSURE(ctx->scan_slot[slot], n * sizeof(vofod_pt) + 64);
  CK(cudaMemcpyAsync(ctx->scan_slot[slot].p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->scan_slot_n[slot] = n;
  return VOFOD_OK;
}
}  // extern "C"

// classify.cu / sepclusters.cu read-back helpers
int vf_classify_readback(vofod_ctx* ctx, vofod_detection* dets, size_t det_cap, size_t* n_dets, vofod_cluster_info* clusters, size_t cl_cap, size_t* n_far);

static int process_scan_dev(vofod_ctx* ctx, const vofod_pt* d_scan, size_t n, const vofod_pose& tf, const vofod_params& p, const vofod_schedule& s, vofod_scan_result* res,
                            vofod_detection* dets, size_t det_cap)
{
  if (res)
    memset(res, 0, sizeof(*res));
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  cudaStream_t st = ctx->stream;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  int e = 0;
  CK(cudaEventRecord(ctx->ev[e++], st));
  // rangefinder seeds (A23)
  RET(vf_range_update_dev(ctx, s.range_pt, p, s.n_range_seeds));
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 0 "range"
  // filterAndTransform (:928)
  RET(vf_filter_voxelize_dev(ctx, d_scan, n, tf, p));
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 1 "filtering"
  // clusterCloud (:932)
  ENSURE(ctx->labels, n * 4);
  RET(vf_cluster_dev(ctx, ctx->cl, reinterpret_cast<const float*>(ctx->vox.p), 4, cnt + CNT_VG_M, n, (float)p.ground_points_max_distance, ctx->labels.as<int>(),
                     cnt + CNT_NCLUSTERS));
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 2 "clusterization"
  // findCloseFarClusters (:936)
  RET(vf_close_far_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), cnt + CNT_VG_M, n, p));
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 3 "close X far"
  // updateVMaps (:946-949)
  RET(vf_update_points_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->pt_close.as<uint8_t>(), 1, cnt + CNT_VG_M, n, (float)p.score_point, 2.0f));
  RET(vf_update_points_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->pt_close.as<uint8_t>(), 0, cnt + CNT_VG_M, n, (float)p.score_unknown, 3.0f));
  ctx->detection_its++;
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 4 "vmap update"
  int raycast_status = VOFOD_W_PAUSED;
  bool applied = false;
  if (s.do_raycast)
  {
    raycast_status = vf_raycast_accumulate_dev(ctx, d_scan, n, tf, p);
    if (raycast_status < 0)
      return raycast_status;
  }
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 5 "raycasting"
  if (s.do_raycast && raycast_status == VOFOD_OK)
  {
    const int rc = vf_raycast_apply_dev(ctx, s.raycast_its_diff > 1 ? s.raycast_its_diff : 1, p);
    if (rc < 0)
      return rc;
    if (rc == VOFOD_OK)
      applied = true;
    else
      raycast_status = rc;
  }
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 6 raycast "vmap update"
  CK(cudaMemsetAsync(cnt + CNT_NDET, 0, 8, st));
  ctx->last_far = 0;
  if (s.do_classify)
    RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, tf, p));
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 7 "classification" (+ 8 detections, fused)
  CK(cudaEventRecord(ctx->ev[e++], st));
  int sep_status = VOFOD_W_PAUSED;
  if (s.do_sepclusters)
  {
    sep_status = vf_sepclusters_dev(ctx, s.sep_its_diff, p);
    if (sep_status < 0)
      return sep_status;
  }
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 9 "sep bg clusters"
  // one read-back of every count
  unsigned long long h[CNT_N_SLOTS];
  unsigned long long* hp = ctx->pinned ? (unsigned long long*)ctx->pinned : h;
  CK(cudaMemcpyAsync(hp, cnt, CNT_N_SLOTS * 8, cudaMemcpyDeviceToHost, st));
  size_t nd_cap = 0;
  vofod_detection* hdets = nullptr;
  if (s.do_classify && dets && det_cap)
  {
    // detections are few: copy a bounded prefix speculatively with the counters, the rest (rare) afterwards
    nd_cap = det_cap < 16 ? det_cap : 16;
    if (ctx->pinned && ctx->dets.p)
    {
      hdets = (vofod_detection*)((char*)ctx->pinned + 4096);
      CK(cudaMemcpyAsync(hdets, ctx->dets.p, nd_cap * sizeof(vofod_detection), cudaMemcpyDeviceToHost, st));
    }
  }
  CK(cudaEventRecord(ctx->ev[e++], st));                         // 10 "readback"
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < 11; i++)
    cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
  cudaEventElapsedTime(&ctx->stage_ms[11], ctx->ev[0], ctx->ev[11]);
  if (hp[CNT_WATCHDOG])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", hp[CNT_WATCHDOG]);
  if (hp[CNT_OOB])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "%llu traversals fell outside the accumulator window", hp[CNT_OOB]);
  if (hp[CNT_VG_OVERFLOW])
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "leaf size too small for the input: integer indices would overflow");
  if (applied)
  {
    if (hp[CNT_APPLY_ANY])
      ctx->flags_full_dirty = false;
    else
      raycast_status = VOFOD_W_EMPTY_RAYCAST;
  }
  ctx->background_pts_sufficient = hp[CNT_STATE_BG] != 0;
  ctx->sure_background_sufficient = hp[CNT_STATE_SURE] != 0;
  ctx->last_detection_id = (uint32_t)hp[CNT_DET_ID];
  ctx->last_m = (size_t)hp[CNT_VG_M];
  const size_t n_det = s.do_classify ? (size_t)hp[CNT_NDET] : 0;
  if (s.do_classify)
    ctx->last_far = (size_t)hp[CNT_NFARPTS];
  if (res)
  {
    res->n_traversals = s.do_raycast ? hp[CNT_TRAVERSALS] : 0;
    res->n_bg = hp[CNT_NBG];
    res->n_filtered = (uint32_t)hp[CNT_VG_NVALID];
    res->n_voxels = (uint32_t)hp[CNT_VG_M];
    res->n_clusters = (uint32_t)hp[CNT_NCLUSTERS];
    res->n_close_clusters = (uint32_t)hp[CNT_NCLOSE];
    res->n_far_clusters = (uint32_t)hp[CNT_NFAR];
    res->n_detections = (uint32_t)n_det;
    res->background_pts_sufficient = ctx->background_pts_sufficient;
    res->sure_background_sufficient = ctx->sure_background_sufficient;
    res->raycast_status = raycast_status;
    res->sep_status = sep_status;
  }
  if (n_det && dets)
  {
    const size_t k = n_det < det_cap ? n_det : det_cap;
    if (hdets && k <= nd_cap)
      memcpy(dets, hdets, k * sizeof(vofod_detection));
    else
    {
      CK(cudaMemcpyAsync(dets, ctx->dets.p, k * sizeof(vofod_detection), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  if (n_det > det_cap && dets)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "process_scan: %zu detections, capacity %zu", n_det, det_cap);
  return VOFOD_OK;
}

extern "C" {

int vofod_process_scan(vofod_ctx* ctx, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                       vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  if (!scan || !tf || !p || !s)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
  CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  return process_scan_dev(ctx, ctx->scan_staging.as<vofod_pt>(), n, *tf, *p, *s, res, dets, det_cap);
}

int vofod_process_scan_resident(vofod_ctx* ctx, int slot, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                                vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  if (!tf || !p || !s || slot < 0 || slot >= VOFOD_SCAN_SLOTS || !ctx->scan_slot_n[slot])
    return vf_fail(ctx, VOFOD_E_INVALID, "bad argument / empty scan slot");
  return process_scan_dev(ctx, ctx->scan_slot[slot].as<vofod_pt>(), ctx->scan_slot_n[slot], *tf, *p, *s, res, dets, det_cap);
}

int vofod_last_voxels(vofod_ctx* ctx, vofod_vox* out, int32_t* labels, uint8_t* in_close, size_t cap, size_t* m)
{
  NEED_MAP();
  if (!m)
    return vf_fail(ctx, VOFOD_E_INVALID, "m is NULL");
  *m = ctx->last_m;
  if (ctx->last_m > cap)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "last_voxels: need capacity %zu", ctx->last_m);
  const size_t k = ctx->last_m;
  if (k)
  {
    if (out)
      CK(cudaMemcpyAsync(out, ctx->vox.p, k * sizeof(vofod_vox), cudaMemcpyDeviceToHost, ctx->stream));
    if (labels)
      CK(cudaMemcpyAsync(labels, ctx->labels.p, k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (in_close)
      CK(cudaMemcpyAsync(in_close, ctx->pt_close.p, k, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_last_clusters(vofod_ctx* ctx, vofod_cluster_info* out, size_t cap, size_t* n)
{
  NEED_MAP();
  if (!n)
    return vf_fail(ctx, VOFOD_E_INVALID, "n is NULL");
  *n = ctx->last_far;
  if (ctx->last_far > cap)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "last_clusters: need capacity %zu", ctx->last_far);
  if (ctx->last_far && out)
    CK(cudaMemcpyAsync(out, ctx->cl_info.p, ctx->last_far * sizeof(vofod_cluster_info), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}
}  // extern "C"
