// Multi-GPU slab mode (no counterpart in the reference; SURVEY.md §8e, BASELINE.json configs[4]): the global grid is cut along a
// horizontal axis, one slab per context / GPU / process.
//
// What is sharded and what is replicated
//   * every grid and every pass over a grid is sharded: a context holds its own range + a halo (the storage box);
//   * the rays are clipped: a slab walks only the part of a ray that lies in its storage box — the DDA is fast-forwarded from the
//     GLOBAL start (same fp32 additions per axis as walking, raycast.cu: ray_skip_to_slab), so voxel sequence and path lengths are those
//     of the monolithic traversal, and the path-length sums are exact integers: halo copies agree with the owner's bit for bit;
//   * the per-scan POINT pipeline (crop, voxel grid, Euclidean clustering, moments of inertia) works on ~10^4..10^5 voxels and is run
//     replicated on the broadcast scan: every slab holds identical voxels, labels and far-cluster records;
//   * point / rangefinder / frontier / decay writes are applied to every HELD cell (own or halo): they are deterministic functions of
//     replicated data, so halos never need an exchange.
// What crosses slabs (a scan = 4 phases, after each the listed buffers are combined over all slabs):
//   phase 0  seeds, filter, clustering, hasCloseTo of the points a slab OWNS, its share of nVoxelsOver
//            -> SUM  u64  background count;  MAX  i32[n]  per-cluster close flags
//   phase 1  close/far split, point update, clipped raycast + apply, far clusters + gates, candidates' map boxes packed
//            -> SUM  u64  traversal count;   SUM  u32[budget]  the candidates' boxes (each cell has exactly one owner, classify.cu)
//   phase 2  exploreToGround / frontier write-back / detections on the summed boxes (replicated), owned background voxels packed
//            -> GATHER u32[cap + 1]  background-voxel lists (cluster fragments of the whole map, sepclusters.cu)
//   phase 3  VoxelGridCounted + clustering + sure counts on the gathered list (replicated), decay of the held cells, read-back
// The caller either drives the phases and combines the buffers itself (vofod_slab_phase / vofod_slab_exchanges — several slabs on one
// device in the tests) or hands the library an NCCL communicator (vofod_comm_init) and calls vofod_slab_process_scan, which puts
// ncclBroadcast / ncclAllReduce / ncclAllGather on the context's own stream between the phases: one host synchronisation per scan.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the library is resolved at run time (libnccl.so.2), libvofod_cuda has no link dependency on it

#include "common.cuh"
#include "prims.cuh"

int vf_map_alloc(vofod_ctx* ctx);  // ctx.cu

#define NEED_MAP()                                                                        \
  if (!ctx)                                                                               \
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");                             \
  CK(cudaSetDevice(ctx->device));                                                         \
  if (!ctx->map_ready)                                                                    \
  return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized (call vofod_map_resize / vofod_reset first)")

// ---- NCCL, resolved at run time -------------------------------------------------------------------------------------------------
namespace
{
struct NcclApi
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api()
{
  static NcclApi api;
  static bool tried = false;
  if (!tried)
  {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // the copy the process already has (torch's) when there is one
    if (!h)
      h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h)
    {
      api.handle = h;
      *(void**)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId");
      *(void**)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
      *(void**)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
      *(void**)&api.AllReduce = dlsym(h, "ncclAllReduce");
      *(void**)&api.Broadcast = dlsym(h, "ncclBroadcast");
      *(void**)&api.AllGather = dlsym(h, "ncclAllGather");
      *(void**)&api.GroupStart = dlsym(h, "ncclGroupStart");
      *(void**)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
      *(void**)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast || !api.AllGather || !api.GroupStart || !api.GroupEnd)
        api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}
}  // namespace
#define NCK(call)                                                                                                   \
  do                                                                                                                \
  {                                                                                                                 \
    ctx->pdl_chain = false;                                                                                         \
    const ncclResult_t r__ = (call);                                                                                \
    if (r__ != ncclSuccess)                                                                                         \
      return vf_fail(ctx, VOFOD_E_CUDA, "%s failed: %s", #call, nc->GetErrorString ? nc->GetErrorString(r__) : "NCCL error"); \
  } while (0)

// smallest halo that keeps every result bit-identical to the unsharded run: the slab that owns a point's voxel answers hasCloseTo for
// it and must hold the whole window [o - mv, o + mv) (voxel_map.cpp:384-391).  Everything else either works on exchanged data
// (classification boxes, background lists) or writes replicated values into whatever cells are held.
static int slab_min_halo(const vofod_params& p, const float vs) { return (int)ceilf((float)p.ground_points_max_distance * (1.0f / vs)) + 1; }

extern "C" {
int vofod_slab_min_halo(const vofod_params* p, float voxel_size)
{
  if (!p || !(voxel_size > 0.f))
    return VOFOD_E_INVALID;
  return slab_min_halo(*p, voxel_size);
}

int vofod_set_slab(vofod_ctx* ctx, int axis, int lo, int hi, int halo)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  Geom& g = ctx->g;
  if (axis < 0 || axis > 1 || lo < 0 || hi > g.size[axis] || lo >= hi || halo < 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_set_slab: bad range [%d,%d) halo %d on axis %d (horizontal axes only)", lo, hi, halo, axis);
  for (int a = 0; a < 3; a++)
  {
    g.st_lo[a] = 0;
    g.st_size[a] = g.size[a];
  }
  const int slo = lo - halo > 0 ? lo - halo : 0;
  const int shi = hi + halo < g.size[axis] ? hi + halo : g.size[axis];
  g.st_lo[axis] = slo;
  g.st_size[axis] = shi - slo;
  g.slab_axis = axis;
  g.own_lo = lo;
  g.own_hi = hi;
  ctx->slab_on = !(lo == 0 && hi == g.size[axis]);
  ctx->slab_halo = halo;
  ctx->slab_next_phase = 0;
  // the grids are re-allocated for the storage box; their contents are unspecified until the next vofod_map_set_to
  return vf_map_alloc(ctx);
}

int vofod_comm_unique_id(void* out128)
{
  NcclApi* nc = nccl_api();
  if (!nc || !out128)
    return VOFOD_E_STATE;
  ncclUniqueId id;
  if (nc->GetUniqueId(&id) != ncclSuccess)
    return VOFOD_E_CUDA;
  memcpy(out128, &id, sizeof(id));
  return VOFOD_OK;
}

int vofod_comm_init(vofod_ctx* ctx, int rank, int nranks, const void* nccl_unique_id)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_unique_id))
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_comm_init: bad rank %d / %d", rank, nranks);
  ctx->slab_rank = rank;
  ctx->slab_nranks = nranks;
  if (nranks == 1)
    return VOFOD_OK;
  NcclApi* nc = nccl_api();
  if (!nc)
    return vf_fail(ctx, VOFOD_E_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  NCK(nc->CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  return VOFOD_OK;
}

}  // extern "C"
// vofod_destroy
void vf_slab_destroy(vofod_ctx* ctx)
{
  NcclApi* nc = ctx->nccl_comm ? nccl_api() : nullptr;
  if (nc && nc->CommDestroy)
    nc->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}
extern "C" {
/* number of slabs the exchanged buffers are sized for, when the caller combines them itself (vofod_comm_init sets it otherwise) */
int vofod_slab_set_world(vofod_ctx* ctx, int rank, int nranks)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_slab_set_world: bad rank %d / %d", rank, nranks);
  ctx->slab_rank = rank;
  ctx->slab_nranks = nranks;
  return VOFOD_OK;
}
}  // extern "C"

// ---- boundary fragments --------------------------------------------------------------------------------------------------------
// voxels of the last scan within `margin` cells of a face of the OWNED range, with their cluster labels: the fragments of clusters
// that reach into the neighbouring slab.  With the replicated point pipeline the labels are global already, so this is a consistency
// probe (tests compare the fragments of neighbouring slabs) and the hook for a sharded point pipeline.
__global__ void __launch_bounds__(256) k_slab_boundary(const vofod_vox* __restrict__ vox, const int* __restrict__ labels, const size_t m, const Geom g, const int margin,
                                                       int32_t* __restrict__ out_idx, int32_t* __restrict__ out_label, const size_t cap, unsigned long long* __restrict__ n_out)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const vofod_vox v = vox[i];
    const int c = g.slab_axis == 0 ? coord_to_idx1(v.x, g.off[0], g.inv) : coord_to_idx1(v.y, g.off[1], g.inv);
    const bool near_lo = g.own_lo > 0 && c >= g.own_lo - margin && c < g.own_lo + margin;
    const bool near_hi = g.own_hi < g.size[g.slab_axis] && c >= g.own_hi - margin && c < g.own_hi + margin;
    if (near_lo || near_hi)
    {
      const unsigned long long o = atomicAdd(n_out, 1ull);
      if (o < cap)
      {
        out_idx[o] = (int32_t)i;
        out_label[o] = labels[i];
      }
    }
  }
}

extern "C" int vofod_slab_boundary(vofod_ctx* ctx, int margin, int32_t* point_idx, int32_t* labels, size_t cap, size_t* n)
{
  NEED_MAP();
  if (!n || margin < 0 || (cap && (!point_idx || !labels)))
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_slab_boundary: bad arguments");
  *n = 0;
  const size_t m = ctx->last_m;
  if (m == 0)
    return VOFOD_OK;
  ENSURE(ctx->scratch_a, (cap + 1) * 4);
  ENSURE(ctx->scratch_b, (cap + 1) * 4);
  unsigned long long* d_n = vf_cnt(ctx, CNT_SCRATCH0);
  CK(cudaMemsetAsync(d_n, 0, 8, ctx->stream));
  LAUNCH(k_slab_boundary, vf_blocks(ctx, m, 256, 8), 256, 0, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), m, ctx->g, margin, ctx->scratch_a.as<int32_t>(),
         ctx->scratch_b.as<int32_t>(), cap, d_n);
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, d_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = (size_t)h;
  const size_t k = h < cap ? (size_t)h : cap;
  if (k)
  {
    CK(cudaMemcpyAsync(point_idx, ctx->scratch_a.p, k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(labels, ctx->scratch_b.p, k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return h > cap ? vf_fail(ctx, VOFOD_E_CAPACITY, "vofod_slab_boundary: need capacity %llu", h) : VOFOD_OK;
}

// ---- the four phases of a scan ----------------------------------------------------------------------------------------------------
static size_t patch_budget_words(const vofod_ctx* ctx, const vofod_params& p)
{
  // room for a handful of worst-case candidates (max_size cluster + explore radius on every side), at least 256 K words
  const double vs = (double)ctx->g.vs;
  const double side = ceil(p.cls_max_size / vs) + 2.0 * ((p.cls_max_size + p.cls_max_explore_distance) / vs + 2.0) + 3.0;
  double w = 3.0 * side * side * side;
  if (w < 262144.0)
    w = 262144.0;
  if (w > 16777216.0)
    w = 16777216.0;
  return ((size_t)w + 1023) & ~(size_t)1023;
}

static int slab_phase(vofod_ctx* ctx, const int phase, vofod_scan_result* res, vofod_detection* dets, const size_t det_cap)
{
  const vofod_params& p = ctx->slab_p;
  const vofod_schedule& s = ctx->slab_s;
  const size_t n = ctx->slab_n;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  const bool classify = s.do_classify != 0;
  const bool sep = s.do_sepclusters != 0 && !p.sep_pause;
  switch (phase)
  {
    case 0:
    {
      RET(vf_begin_call(ctx));
      RET(vf_dyn_push(ctx));
      RET(vf_range_update_dev(ctx, p));
      // The clipped raycast reads only the scan, the LUT and the per-scan arguments and writes only the accumulator window: it runs on the
      // second stream beside phase 0, the exchange after it and the head of phase 1 (an issue-bound kernel next to chains of short,
      // latency-bound ones); phase 1 joins before the apply.  Only for windows that live in L2 (20 m rays: 1.28 -> 1.18 ms per scan on two
      // slabs): a GB-sized window's accumulate runs for milliseconds with every warp slot taken, and the short kernels of phase 0 then
      // wait for slots launch after launch (200 m rays: 4.34 -> 4.52 ms, measured).
      ctx->slab_acc_forked = false;
      if (s.do_raycast && ctx->slab_raycast_status == VOFOD_OK && ctx->stream2 != nullptr && ctx->overlap_enabled && !ctx->acc_sparse)
      {
        cudaStream_t st = ctx->stream;
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        ctx->stream = ctx->stream2;
        const int rrc = vf_raycast_accumulate_dev(ctx, n, vofod_pose(), p);
        ctx->stream = st;
        if (rrc < 0)
          return rrc;
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
        ctx->slab_acc_forked = true;
      }
      const int seeded = vf_filter_voxelize_dev(ctx, n, p, true);
      if (seeded < 0)
        return seeded;
      ENSURE(ctx->labels, n * 4);
      if (seeded)
      {
        // the voxel list is a set of grid cells: connected components on the occupancy words the filter left behind (as in vofod_process_scan)
        RunRows rows;
        vf_run_rows((float)p.ground_points_max_distance, ctx->g.vs, rows);
        RET(vf_cluster_runs_dev(ctx, ctx->cl, ctx->cl_cellkey.as<uint32_t>(), ctx->cl_words.as<RunWord>(),
                                reinterpret_cast<const VgLayout*>(ctx->scratch_d.as<char>() + 64), rows, (float)p.ground_points_max_distance, cnt + CNT_VG_M, n,
                                ctx->labels.as<int>(), cnt + CNT_NCLUSTERS));
      } else
        RET(vf_cluster_dev(ctx, ctx->cl, reinterpret_cast<const float*>(ctx->vox.p), 4, cnt + CNT_VG_M, n, (float)p.ground_points_max_distance, ctx->labels.as<int>(),
                           cnt + CNT_NCLUSTERS));
      RET(vf_close_far_phase(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), cnt + CNT_VG_M, n, p, 1, false));
      return VOFOD_OK;
    }
    case 1:
    {
      RET(vf_close_far_phase(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), cnt + CNT_VG_M, n, p, 2, true));
      RET(vf_update_points_scan_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p));
      ctx->detection_its++;
      if (s.do_raycast && ctx->slab_raycast_status == VOFOD_OK)
      {
        if (ctx->slab_acc_forked)
          CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        else
          RET(vf_raycast_accumulate_dev(ctx, n, vofod_pose(), p));
        ctx->slab_acc_forked = false;
        const int rc = vf_raycast_apply_dev(ctx, 0, p);
        if (rc < 0)
          return rc;
        ctx->slab_applied = rc == VOFOD_OK;
      } else
      {
        CK(cudaMemsetAsync(cnt + CNT_TRAVERSALS, 0, 8, ctx->stream));
        CK(cudaMemsetAsync(cnt + CNT_OOB, 0, 8, ctx->stream));
        ctx->slab_applied = false;
      }
      CK(cudaMemsetAsync(cnt + CNT_NDET, 0, 8, ctx->stream));
      if (classify)
      {
        ENSURE(ctx->slab_patch, ctx->slab_patch_words * 4);
        ENSURE(ctx->slab_patch_desc, 64 * 64);
        ENSURE(ctx->slab_patch_meta, 64);
        RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p, 1));
        RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p, 3));
      }
      return VOFOD_OK;
    }
    case 2:
    {
      if (classify && !ctx->slab_redo_sep)
        RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p, 4));
      CK(cudaMemsetAsync(cnt + CNT_SEP_K, 0, 8, ctx->stream));
      if (sep)
      {
        ENSURE(ctx->slab_bg_send, (ctx->slab_bg_cap + 1) * 4);
        ENSURE(ctx->slab_bg_recv, (ctx->slab_bg_cap + 1) * 4 * (size_t)ctx->slab_nranks);
        if (ctx->slab_redo_sep)
          RET(vf_begin_call(ctx));
        RET(vf_sep_slab_pack(ctx, p, ctx->slab_bg_cap));
      }
      return VOFOD_OK;
    }
    default:
      break;
  }
  // ---- phase 3
  int sep_status = VOFOD_W_PAUSED;
  if (sep)
  {
    sep_status = vf_sep_slab_finish(ctx, s.sep_its_diff, p, ctx->slab_bg_cap, ctx->slab_nranks, ctx->slab_bg_consume);
    if (sep_status < 0)
      return sep_status;
  }
  CK(cudaMemcpyAsync(ctx->pinned, cnt, CNT_N_SLOTS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long* pm = (unsigned long long*)ctx->pinned + CNT_N_SLOTS;
  if (classify)
    CK(cudaMemcpyAsync(pm, ctx->slab_patch_meta.p, 24, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const unsigned long long* hp = (const unsigned long long*)ctx->pinned;
  if (hp[CNT_WATCHDOG])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", hp[CNT_WATCHDOG]);
  if (hp[CNT_OOB])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "%llu traversals fell outside the accumulator window", hp[CNT_OOB]);
  if (hp[CNT_VG_OVERFLOW])
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "leaf size too small for the input: integer indices would overflow");
  if (sep)
  {
    const unsigned long long K = hp[CNT_SEP_K];
    const size_t cap_all = ctx->slab_bg_cap * (size_t)ctx->slab_nranks;
    if (K > cap_all || K > ctx->slab_bg_consume)
    {
      // some slab's list (or their sum) did not fit: nothing was touched (k_sep_decay), every slab sees the same gathered counts and
      // takes the same decision — grow and let the caller repeat phases 2 and 3
      ctx->slab_bg_cap *= 4;
      ctx->slab_bg_consume *= 4;
      ctx->slab_redo_sep = true;
      ctx->slab_next_phase = 2;
      return VOFOD_W_REDO;
    }
    if (hp[CNT_SEP_NUNIQ])
      return vf_fail(ctx, VOFOD_E_OVERFLOW, "sepclusters: voxel-grid index overflow");
    if (K == 0)
      sep_status = VOFOD_W_EMPTY;
    // capacities follow the list (a single slab may come to hold most of it): both are functions of the GLOBAL count, which every slab
    // reads from the same gathered buffer, so all slabs keep identical sizes — the allgather needs that
    if (!ctx->sep_cap_forced && (K * 3 / 2 > ctx->slab_bg_cap || K * 8 < ctx->slab_bg_cap))
      ctx->slab_bg_cap = (size_t)K * 2 + 4096;
    if (!ctx->sep_cap_forced && (K * 3 / 2 > ctx->slab_bg_consume || K * 8 < ctx->slab_bg_consume))
      ctx->slab_bg_consume = (size_t)K * 2 + 4096;
    if (ctx->sep_cap_forced && K * 3 / 2 > ctx->slab_bg_cap)
    {
      ctx->slab_bg_cap = (size_t)K * 2 + 4096;
      ctx->slab_bg_consume = ctx->slab_bg_cap;
    }
    if ((size_t)K * 3 / 2 > ctx->sep_table_hint)
      ctx->sep_table_hint = (size_t)K * 4 + (size_t(1) << 18);
  }
  ctx->slab_redo_sep = false;
  if (classify && pm[2])
    return vf_fail(ctx, VOFOD_E_CAPACITY, "slab classification: the candidates' map boxes exceed the exchange buffer (%zu words, %d boxes): raise VOFOD_OPT_SLAB_PATCH_WORDS",
                   ctx->slab_patch_words, 64);
  int raycast_status = ctx->slab_raycast_status;
  if (ctx->slab_applied)
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
  const size_t n_det = classify ? (size_t)hp[CNT_NDET] : 0;
  ctx->last_far = classify ? (size_t)hp[CNT_NFARPTS] : 0;
  if (res)
  {
    memset(res, 0, sizeof(*res));
    res->n_traversals = hp[CNT_TRAVERSALS];
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
    CK(cudaMemcpyAsync(dets, ctx->dets.p, k * sizeof(vofod_detection), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->slab_n = 0;
  if (n_det > det_cap && dets)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "slab scan: %zu detections, capacity %zu", n_det, det_cap);
  return VOFOD_OK;
}

// host-side set-up of a scan (no launches): arguments, window, buffer sizes
static int slab_begin(vofod_ctx* ctx, const vofod_pt* d_scan, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s)
{
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  if (!p->raycast_new_update_rule && s->do_raycast && ctx->slab_on)
    return vf_fail(ctx, VOFOD_E_INVALID, "slab mode supports the new raycast update rule only (the old rule needs the global max path length)");
  if (ctx->slab_on && ctx->slab_halo < slab_min_halo(*p, ctx->g.vs))
    return vf_fail(ctx, VOFOD_E_INVALID, "slab halo of %d cells is too small: ground_points_max_distance needs %d (vofod_slab_min_halo)", ctx->slab_halo,
                   slab_min_halo(*p, ctx->g.vs));
  ctx->slab_p = *p;
  ctx->slab_s = *s;
  ctx->slab_raycast_status = VOFOD_W_PAUSED;
  if (s->do_raycast)
  {
    ctx->slab_raycast_status = vf_raycast_prepare(ctx, n, *tf, *p);
    if (ctx->slab_raycast_status < 0)
      return ctx->slab_raycast_status;
  }
  ScanDyn* hd = ctx->h_dyn;
  memcpy(hd->tf.R, tf->R, sizeof(tf->R));
  memcpy(hd->tf.t, tf->t, sizeof(tf->t));
  for (int a = 0; a < 3; a++)
    hd->range_pt[a] = s->range_pt[a];
  hd->n_seeds = s->n_range_seeds > 0 ? s->n_range_seeds : 0;
  hd->scan = d_scan;
  hd->its_raycast = s->raycast_its_diff > 1 ? s->raycast_its_diff : 1;
  ctx->slab_n = n;
  if (ctx->slab_patch_words_forced)
    ctx->slab_patch_words = ctx->slab_patch_words_forced;
  else if (ctx->slab_patch_words == 0)
    ctx->slab_patch_words = patch_budget_words(ctx, *p);
  if (ctx->slab_bg_cap == 0)
  {
    ctx->slab_bg_cap = ctx->sep_cap_forced ? ctx->sep_cap_forced : (size_t(1) << 16);
    ctx->slab_bg_consume = ctx->slab_bg_cap;
  }
  ctx->slab_redo_sep = false;
  return VOFOD_OK;
}

extern "C" {

/* One scan of schedule S1 in slab mode, phase by phase; the caller combines the buffers of vofod_slab_exchanges(phase) over all slabs
 * after every phase.  Phase 0 takes the scan (host pointer, or a device pointer when scan_on_device) and the arguments; phases 1..3
 * continue it (scan / tf / p / s are ignored); phase 3 delivers the results.  VOFOD_W_REDO from phase 3: repeat phases 2 and 3. */
int vofod_slab_phase(vofod_ctx* ctx, int phase, const vofod_pt* scan, int scan_on_device, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s,
                     vofod_scan_result* res, vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (phase < 0 || phase > 3)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_slab_phase: phase %d", phase);
  if (phase == 0)
  {
    if (!scan || !tf || !p || !s)
      return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
    const vofod_pt* d_scan = scan;
    if (!scan_on_device)
    {
      ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
      CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
      d_scan = ctx->scan_staging.as<vofod_pt>();
    }
    RET(slab_begin(ctx, d_scan, n, tf, p, s));
  } else if (!ctx->slab_n || phase != ctx->slab_next_phase)
    return vf_fail(ctx, VOFOD_E_STATE, "vofod_slab_phase(%d) out of order (expected %d)", phase, ctx->slab_n ? ctx->slab_next_phase : 0);
  ctx->slab_next_phase = phase + 1;
  return slab_phase(ctx, phase, res, dets, det_cap);
}

/* the buffers to combine after `phase` (at most 4); kind: VOFOD_XCHG_* */
int vofod_slab_exchanges(vofod_ctx* ctx, int phase, vofod_slab_exchange* out, int* n_out)
{
  NEED_MAP();
  if (!out || !n_out || !ctx->slab_n)
    return vf_fail(ctx, VOFOD_E_STATE, "vofod_slab_exchanges: no scan in flight / NULL argument");
  const vofod_schedule& s = ctx->slab_s;
  int k = 0;
  auto add = [&](int kind, void* buf, size_t count, void* gather_out) {
    out[k].kind = kind;
    out[k]._pad = 0;
    out[k].buf = buf;
    out[k].count = count;
    out[k].gather_out = gather_out;
    k++;
  };
  if (phase == 0)
  {
    add(VOFOD_XCHG_SUM_U64, vf_cnt(ctx, CNT_NBG), 1, nullptr);
    add(VOFOD_XCHG_MAX_I32, ctx->cl_close.p, ctx->slab_n, nullptr);
  } else if (phase == 1)
  {
    add(VOFOD_XCHG_SUM_U64, vf_cnt(ctx, CNT_TRAVERSALS), 1, nullptr);
    if (s.do_classify)
      add(VOFOD_XCHG_SUM_U32, ctx->slab_patch.p, ctx->slab_patch_words, nullptr);
  } else if (phase == 2)
  {
    if (s.do_sepclusters && !ctx->slab_p.sep_pause)
      add(VOFOD_XCHG_GATHER_U32, ctx->slab_bg_send.p, ctx->slab_bg_cap + 1, ctx->slab_bg_recv.p);
  }
  *n_out = k;
  return VOFOD_OK;
}

/* The same scan over NCCL (vofod_comm_init): rank 0 passes the scan (host memory), the others may pass NULL; pose, parameters and
 * schedule are given on every rank (a few hundred bytes of host metadata the caller distributes however it likes).  The scan is
 * broadcast and all exchanges run on the context's stream; the host waits once, at the end. */
int vofod_slab_process_scan(vofod_ctx* ctx, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                            vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (!tf || !p || !s || (ctx->slab_rank == 0 && !scan))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  NcclApi* nc = ctx->slab_nranks > 1 ? nccl_api() : nullptr;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (ctx->slab_nranks > 1 && (!nc || !comm))
    return vf_fail(ctx, VOFOD_E_STATE, "vofod_slab_process_scan on %d ranks needs vofod_comm_init", ctx->slab_nranks);
  ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
  // where the scan's time goes (vofod_slab_times): events on the stream before the copy, after the broadcast, after every phase's kernels and
  // after every exchange
  CK(cudaEventRecord(ctx->ev_slab[0], ctx->stream));
  const void* d_src = ctx->scan_staging.p;
  if (ctx->slab_rank == 0)
  {
    // a scan announced with vofod_prefetch_scan has been copied next to the previous scan's kernels
    int hit = -1;
    ctx->scan_calls++;
    for (int i = 0; i < 2; i++)
      if (ctx->prefetched_host[i] && ctx->scan_calls - ctx->prefetched_call[i] > 2)
        ctx->prefetched_host[i] = nullptr;  // (see vofod_process_scan)
    for (int i = 0; i < 2; i++)
      if (ctx->prefetched_host[i] == (const void*)scan && ctx->prefetched_n[i] == n && ctx->prefetch_buf[i].p)
        hit = i;
    if (hit >= 0)
    {
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_prefetch[hit], 0));
      ctx->prefetched_host[hit] = nullptr;
      ctx->stat_prefetch_hits++;
      d_src = ctx->prefetch_buf[hit].p;
    } else
      CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  }
  const vofod_pt* d_scan = ctx->scan_staging.as<vofod_pt>();
  if (nc)
    NCK(nc->Broadcast(d_src, ctx->scan_staging.p, n * sizeof(vofod_pt), ncclUint8, 0, comm, ctx->stream));
  else
    d_scan = (const vofod_pt*)d_src;
  CK(cudaEventRecord(ctx->ev_slab[1], ctx->stream));
  RET(slab_begin(ctx, d_scan, n, tf, p, s));
  int rc = VOFOD_OK;
  bool redone = false;
  for (int phase = 0; phase <= 3; phase++)
  {
    ctx->slab_next_phase = phase + 1;
    if (phase == 3)
      CK(cudaEventRecord(ctx->ev_slab[8], ctx->stream));  // (phase 3 ends with the host wait: its end is taken on the host side below)
    rc = slab_phase(ctx, phase, res, dets, det_cap);
    if (rc < 0)
      return rc;
    if (rc == VOFOD_W_REDO)
    {
      phase = 1;  // continue with phase 2
      redone = true;
      continue;
    }
    if (phase == 3)
    {
      // the stream is idle here (phase 3 waited for it)
      if (!redone)
      {
        for (int i = 0; i < 7; i++)
          cudaEventElapsedTime(&ctx->slab_ms[i], ctx->ev_slab[i], ctx->ev_slab[i + 1]);
        CK(cudaEventRecord(ctx->ev_slab[0], ctx->stream));
        CK(cudaEventSynchronize(ctx->ev_slab[0]));
        cudaEventElapsedTime(&ctx->slab_ms[7], ctx->ev_slab[8], ctx->ev_slab[0]);
      }
      continue;
    }
    CK(cudaEventRecord(ctx->ev_slab[2 + 2 * phase], ctx->stream));
    vofod_slab_exchange x[4];
    int nx = 0;
    RET(vofod_slab_exchanges(ctx, phase, x, &nx));
    if (!nc)
    {
      // a single slab: sums and maxima over one slab are the buffers themselves, the gathered list is the slab's own
      for (int i = 0; i < nx; i++)
        if (x[i].kind == VOFOD_XCHG_GATHER_U32)
          CK(cudaMemcpyAsync(x[i].gather_out, x[i].buf, x[i].count * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      CK(cudaEventRecord(ctx->ev_slab[3 + 2 * phase], ctx->stream));
      continue;
    }
    if (nx == 0)
    {
      CK(cudaEventRecord(ctx->ev_slab[3 + 2 * phase], ctx->stream));
      continue;
    }
    NCK(nc->GroupStart());
    for (int i = 0; i < nx; i++)
    {
      ncclResult_t r = ncclSuccess;
      switch (x[i].kind)
      {
        case VOFOD_XCHG_SUM_U64: r = nc->AllReduce(x[i].buf, x[i].buf, x[i].count, ncclUint64, ncclSum, comm, ctx->stream); break;
        case VOFOD_XCHG_MAX_I32: r = nc->AllReduce(x[i].buf, x[i].buf, x[i].count, ncclInt32, ncclMax, comm, ctx->stream); break;
        case VOFOD_XCHG_SUM_U32: r = nc->AllReduce(x[i].buf, x[i].buf, x[i].count, ncclUint32, ncclSum, comm, ctx->stream); break;
        default: r = nc->AllGather(x[i].buf, x[i].gather_out, x[i].count, ncclUint32, comm, ctx->stream); break;
      }
      if (r != ncclSuccess)
      {
        nc->GroupEnd();
        return vf_fail(ctx, VOFOD_E_CUDA, "NCCL exchange %d of phase %d failed: %s", i, phase, nc->GetErrorString ? nc->GetErrorString(r) : "?");
      }
    }
    NCK(nc->GroupEnd());
    CK(cudaEventRecord(ctx->ev_slab[3 + 2 * phase], ctx->stream));
  }
  return rc;
}

/* device time [ms] of the parts of the last vofod_slab_process_scan on this rank: scan copy + broadcast, then (phase k kernels, exchange
 * after phase k) for k = 0..2, then phase 3 (with its read-back).  An exchange's time includes waiting for the slowest rank. */
int vofod_slab_times(vofod_ctx* ctx, float ms[8])
{
  if (!ctx || !ms)
    return VOFOD_E_INVALID;
  memcpy(ms, ctx->slab_ms, sizeof(ctx->slab_ms));
  return VOFOD_OK;
}
}  // extern "C"
