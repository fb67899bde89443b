// Context lifetime + vofod::VoxelMap (C1) entry points of libvofod_cuda.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "prims.cuh"

static thread_local std::string g_last_error = "";

int vf_fail(vofod_ctx* c, int code, const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (c)
    c->err = buf;
  return code;
}

int vf_ensure(vofod_ctx* ctx, DevBuf& b, size_t bytes)
{
  if (bytes == 0)
    bytes = 256;
  if (b.cap >= bytes)
    return 0;
  if (ctx->capturing)
  {
    ctx->capture_broken = true;  // an allocation cannot happen inside a stream capture: the caller falls back to the eager path
    return vf_fail(ctx, VOFOD_E_STATE, "buffer growth during graph capture");
  }
  ctx->alloc_gen++;
  // grow geometrically for small buffers, exactly for large ones
  size_t want = bytes < (size_t(64) << 20) ? bytes + bytes / 2 : bytes;
  want = (want + 255) & ~size_t(255);
  if (b.p)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess)
  {
    b.p = nullptr;
    cudaGetLastError();
    return vf_fail(ctx, VOFOD_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  CK(cudaMemsetAsync(b.p, 0, want, ctx->stream));
  return 0;
}

extern "C" {

const char* vofod_last_error(const vofod_ctx* ctx)
{
  if (ctx)
    return ctx->err.c_str();
  return g_last_error.c_str();
}

int vofod_create(int device, vofod_ctx** out)
{
  if (!out)
    return vf_fail(nullptr, VOFOD_E_INVALID, "vofod_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return vf_fail(nullptr, VOFOD_E_CUDA, "vofod_create: no CUDA device (%s); libvofod_cuda has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev)
    return vf_fail(nullptr, VOFOD_E_INVALID, "vofod_create: device %d out of range [0,%d)", device, ndev);
  vofod_ctx* ctx = new vofod_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess)
  {
    delete ctx;
    return vf_fail(nullptr, VOFOD_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
  {
    delete ctx;
    return vf_fail(nullptr, VOFOD_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->stream4, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->ev_fork4, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_sepfill, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_front, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_done[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_done[1], cudaEventDisableTiming);
  for (auto& e : ctx->ev_slab)
    cudaEventCreate(&e);
  cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->ev_prefetch[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_prefetch[1], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fills, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_cls, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_cp, cudaEventDisableTiming);
  for (int i = 0; i <= VOFOD_N_STAGES; i++)
    cudaEventCreate(&ctx->ev[i]);
  ctx->ev_ok = true;
  memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms));
  memset(&ctx->g, 0, sizeof(ctx->g));
  memset(&ctx->win, 0, sizeof(ctx->win));
  int rc = vf_ensure(ctx, ctx->d_counters, CNT_N_SLOTS * sizeof(unsigned long long));
  if (rc < 0)
  {
    delete ctx;
    return rc;
  }
  if (getenv("VOFOD_NO_OVERLAP"))
    ctx->overlap_enabled = false;
  if (const char* e = getenv("VOFOD_RAYCAST_EXP"))  // A/B runs of the accumulate kernel's variants under the whole test suite (values >= 3 give correct results)
    ctx->raycast_exp = atoi(e);
  if (const char* e = getenv("VOFOD_CLASSIFY_SEQ"))
    ctx->cls_force_seq = atoi(e) != 0;
  if (const char* e = getenv("VOFOD_RAYCAST_SPREAD"))
    ctx->raycast_spread_voxels = atoi(e);
  ctx->pinned_bytes = 1 << 20;
  if (cudaHostAlloc(&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault) != cudaSuccess)
  {
    ctx->pinned = nullptr;
    cudaGetLastError();
  }
  rc = vf_ensure(ctx, ctx->dyn, sizeof(ScanDyn) + 64);
  if (rc < 0 || !ctx->pinned)
  {
    delete ctx;
    return rc < 0 ? rc : vf_fail(nullptr, VOFOD_E_NOMEM, "cudaHostAlloc failed");
  }
  ctx->h_dyn = reinterpret_cast<ScanDyn*>((char*)ctx->pinned + 65536);
  memset(ctx->h_dyn, 0, sizeof(ScanDyn));
  cudaStreamSynchronize(ctx->stream);
  *out = ctx;
  return VOFOD_OK;
}

static void free_buf(DevBuf& b)
{
  if (b.p)
    cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int vofod_destroy(vofod_ctx* ctx)
{
  if (!ctx)
    return VOFOD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  vf_slab_destroy(ctx);
  DevBuf* bufs[] = {&ctx->slab_patch, &ctx->slab_patch_desc, &ctx->slab_patch_meta, &ctx->slab_bg_send, &ctx->slab_bg_recv, &ctx->score, &ctx->flags, &ctx->col_dirty, &ctx->upd_owner, &ctx->upd_leftover, &ctx->flagged, &ctx->acc, &ctx->lut_dir, &ctx->lut_off, &ctx->mask, &ctx->dyn, &ctx->scan_staging, &ctx->prefetch_buf[0], &ctx->prefetch_buf[1], &ctx->vg_pts, &ctx->vg_keys_a, &ctx->vg_keys_b, &ctx->vgh_cnt, &ctx->vgh_bits, &ctx->vgh_list, &ctx->cl_cellkey, &ctx->cl_words, &ctx->vg_flags, &ctx->vg_scan, &ctx->vg_ustart,
                    &ctx->vg_ukey, &ctx->vg_pref, &ctx->vox, &ctx->d_counters, &ctx->tile_state, &ctx->tile_state2, &ctx->tile_state_b, &ctx->tile_state2_b, &ctx->sort_hist, &ctx->cl.pts, &ctx->cl.table_key,
                    &ctx->cl.table_head, &ctx->cl.next, &ctx->cl.parent, &ctx->cl.sizes, &ctx->cl.root, &ctx->cl.minidx, &ctx->cl.cellpts, &ctx->cl_bg.cellpts, &ctx->cl_bg.root, &ctx->cl_bg.minidx, &ctx->cl_bg.pts, &ctx->cl_bg.table_key, &ctx->cl_bg.table_head,
                    &ctx->cl_bg.next, &ctx->cl_bg.parent, &ctx->cl_bg.sizes, &ctx->labels, &ctx->pt_close, &ctx->cl_close, &ctx->far_list,
                    &ctx->cl_info, &ctx->dets, &ctx->explore_ws, &ctx->scratch_a, &ctx->scratch_b, &ctx->scratch_d,
                    &ctx->sep_colcnt, &ctx->sep_coloff, &ctx->sep_raw, &ctx->sep_ds, &ctx->sep_labels, &ctx->sep_nsure, &ctx->sep_offsets, &ctx->sep_segcnt, &ctx->sep_segoff, &ctx->sep_live, &ctx->sep_unsure,
                    &ctx->cls_sizes, &ctx->cls_maxidx, &ctx->cls_seg, &ctx->cls_okeys_a, &ctx->cls_okeys_b, &ctx->cls_queues, &ctx->cls_terms, &ctx->ray_stats};
  for (DevBuf* b : bufs)
    free_buf(*b);
  for (int i = 0; i < VOFOD_SCAN_SLOTS; i++)
    free_buf(ctx->scan_slot[i]);
  for (auto& gs : ctx->gslot)
    if (gs.exec)
      cudaGraphExecDestroy(gs.exec);
  if (ctx->pinned)
    cudaFreeHost(ctx->pinned);
  if (ctx->ev_ok)
    for (int i = 0; i <= VOFOD_N_STAGES; i++)
      cudaEventDestroy(ctx->ev[i]);
  if (ctx->ev_fork)
    cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join)
    cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_fills)
    cudaEventDestroy(ctx->ev_fills);
  if (ctx->ev_fork2)
    cudaEventDestroy(ctx->ev_fork2);
  if (ctx->ev_cls)
    cudaEventDestroy(ctx->ev_cls);
  if (ctx->ev_fork3)
    cudaEventDestroy(ctx->ev_fork3);
  if (ctx->ev_cp)
    cudaEventDestroy(ctx->ev_cp);
  if (ctx->stream2)
    cudaStreamDestroy(ctx->stream2);
  if (ctx->stream3)
    cudaStreamDestroy(ctx->stream3);
  if (ctx->stream4)
    cudaStreamDestroy(ctx->stream4);
  if (ctx->ev_fork4)
    cudaEventDestroy(ctx->ev_fork4);
  if (ctx->ev_sepfill)
    cudaEventDestroy(ctx->ev_sepfill);
  if (ctx->ev_front)
    cudaEventDestroy(ctx->ev_front);
  for (int i = 0; i < 2; i++)
    if (ctx->ev_done[i])
      cudaEventDestroy(ctx->ev_done[i]);
  for (int i = 0; i < 2; i++)
    if (ctx->ev_prefetch[i])
      cudaEventDestroy(ctx->ev_prefetch[i]);
  if (ctx->stream_copy)
  {
    cudaStreamSynchronize(ctx->stream_copy);
    cudaStreamDestroy(ctx->stream_copy);
  }
  cudaStreamDestroy(ctx->stream);
  for (auto& e : ctx->ev_slab)
    if (e)
      cudaEventDestroy(e);
  delete ctx;
  return VOFOD_OK;
}

int vofod_synchronize(vofod_ctx* ctx)
{
  if (!ctx)
    return VOFOD_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_set_option(vofod_ctx* ctx, int option, int value)
{
  if (!ctx)
    return VOFOD_E_INVALID;
  if (option == VOFOD_OPT_GRAPH)
  {
    ctx->graph_enabled = value != 0;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_RAYCAST_BLOCK)
  {
    if (value != 64 && value != 128 && value != 256)
      return vf_fail(ctx, VOFOD_E_INVALID, "raycast block must be 64, 128 or 256");
    ctx->raycast_block = value;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_OVERLAP)
  {
    ctx->overlap_enabled = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_PDL)
  {
    ctx->pdl_enabled = value != 0;
    ctx->pdl_in_graph = value == 2;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_SEP_CAP)
  {
    ctx->sep_cap_forced = value > 0 ? (size_t)value : 0;
    ctx->sep_cap = ctx->sep_cap_forced;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_CLUSTER_HASH)
  {
    ctx->cl_force_hash = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_VG_SORT)
  {
    ctx->vg_force_sort = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_SEP_GENERAL)
  {
    ctx->sep_force_general = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_RAYCAST_STATS)
  {
    ctx->raycast_stats = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_ACC_SPARSE)
  {
    ctx->acc_sparse_mode = value;
    ctx->acc_cells_max = 0;  // forces vf_raycast_prepare to lay the accumulator out again (and clear it)
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_SLAB_PATCH_WORDS)
  {
    ctx->slab_patch_words_forced = value > 0 ? (size_t)value : 0;
    ctx->slab_patch_words = 0;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_CLASSIFY_SEQ)
  {
    ctx->cls_force_seq = value != 0;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_RAYCAST_SPREAD)
  {
    if (value < 0)
      return vf_fail(ctx, VOFOD_E_INVALID, "raycast spread distance must be >= 0 voxels");
    ctx->raycast_spread_voxels = value;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_RAYCAST_EXP)
  {
    ctx->raycast_exp = value;
    ctx->alloc_gen++;
    return VOFOD_OK;
  }
  if (option == VOFOD_OPT_RAYCAST_NO_AGG)
  {
    ctx->raycast_no_agg = value != 0;
    ctx->alloc_gen++;  // invalidates a captured graph
    return VOFOD_OK;
  }
  return vf_fail(ctx, VOFOD_E_INVALID, "unknown option %d", option);
}

void* vofod_stream(vofod_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
/* stream-ordered copies between host memory and a device buffer the library handed out (vofod_slab_exchanges): lets a host loop
 * combine the exchange buffers of several contexts that share one device */
int vofod_dev_read(vofod_ctx* ctx, const void* dev, void* host, size_t bytes)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (bytes && (!dev || !host))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (bytes)
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}
int vofod_dev_write(vofod_ctx* ctx, void* dev, const void* host, size_t bytes)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (bytes && (!dev || !host))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (bytes)
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}
uint64_t vofod_kernel_launches(const vofod_ctx* ctx) { return ctx ? ctx->n_launches : 0; }

void vofod_default_params(vofod_params* p)
{
  memset(p, 0, sizeof(*p));
  p->ground_points_max_distance = 1.5;
  p->output_position_sigma = 0.1;
  p->score_point = 0.0;
  p->score_unknown = -740.0;
  p->score_ray = -1000.0;
  p->thr_apriori_map = 0.0;
  p->thr_sure_obstacles = -0.1;
  p->thr_new_obstacles = -300.0;
  p->thr_frontiers = -750.0;
  p->cls_max_size = 3.0;
  p->cls_max_distance = 50.0;
  p->cls_max_explore_distance = 3.0;
  p->raycast_max_distance = 20.0;
  p->raycast_min_intensity = 0.0;
  p->raycast_weight_coefficient = 0.003;
  p->sep_max_bg_distance = 0.8;
  p->cls_min_points = 2;
  p->raycast_new_update_rule = 1;
  p->sep_min_sure_points = 24;
  p->score_init = -740.0f;
  p->background_sufficient_points_ratio = 0.15f;
  const float eo[3] = {0.09f, 0.0f, -0.75f}, es[3] = {2.5f, 2.5f, 1.6f}, oo[3] = {40.0f, 20.0f, -1.25f}, os[3] = {120.0f, 100.0f, 25.0f};
  for (int a = 0; a < 3; a++)
  {
    p->exclude_box_offset[a] = eo[a];
    p->exclude_box_size[a] = es[a];
    p->oparea_offset[a] = oo[a];
    p->oparea_size[a] = os[a];
  }
  p->sensor_vfov = 1.5707963267948966f;
}
}  // extern "C"

// ======================================================================================================
// kernels
// ======================================================================================================
// `scan` != 0 (vofod_process_scan): also the rangefinder ground seed (A23, vofod_nodelet.cpp:581-613) and the reset of the
// filter's min/max cell — two single-thread kernels less on the scan's critical path
__global__ void k_begin_call(unsigned long long* __restrict__ counters, const int zero_scan_counters, const int scan, float* score, const Geom g,
                             const ScanDyn* dyn, const double score_point, uint8_t* col_dirty, MinMax* mm)
{
  pdl_enter();
  if (scan && threadIdx.x == 31)
  {
    minmax_init(mm);
    if (scan == 1)  // (2: a deferred separated-background pass comes first, the seeds follow it — k_range_update)
      range_update(score, g, dyn, score_point, col_dirty);
  }
  if (threadIdx.x == 0)
    counters[CNT_EPOCH_BASE] += EPOCH_STRIDE;
  if (zero_scan_counters)
  {
    // every counter a scan accumulates into, in one place instead of ~15 eight-byte memset nodes
    const int slots[] = {CNT_TRAVERSALS, CNT_OOB, CNT_APPLY_ANY, CNT_MAXVAL, CNT_NBG, CNT_NCLOSE, CNT_NFAR, CNT_NDET, CNT_NFARPTS, CNT_CLS_CURSOR,
                         CNT_CL_CURSOR, CNT_SEP_K, CNT_SEP_NUNIQ, CNT_SEP_ANY_SURE, CNT_NCLUSTERS, CNT_SEP_NCL, CNT_SEP_LIVE, CNT_VGH_LIST, CNT_SEP_NUNSURE,
                         CNT_VG_OVERFLOW, CNT_CLS_TICKET, CNT_CLS_FINISHED};
    if (threadIdx.x < (int)(sizeof(slots) / sizeof(int)))
      counters[slots[threadIdx.x]] = 0ull;
  }
}

static int begin_call(vofod_ctx* ctx, bool zero_scan_counters, const vofod_params* scan_params, bool seed_now);
int vf_begin_call(vofod_ctx* ctx, bool zero_scan_counters) { return begin_call(ctx, zero_scan_counters, nullptr, true); }
int vf_begin_scan(vofod_ctx* ctx, const vofod_params& p, bool seed_now) { return begin_call(ctx, true, &p, seed_now); }
static int begin_call(vofod_ctx* ctx, bool zero_scan_counters, const vofod_params* scan_params, bool seed_now)
{
  // work done ahead of time by a previous call's side branch does not carry over (that call may have failed half way)
  ctx->nbg_precounted = false;
  ctx->close_points_done = false;
  ctx->cls_prefilled = 0;
  ctx->sep_prefilled = 0;
  ctx->cl.prefilled_tsize = 0;
  ctx->cl_bg.prefilled_tsize = 0;
  ctx->epoch_local = 0;
  ctx->epoch_calls++;
  // the generation field of a look-back state has 30 bits: before it can repeat, forget every old state
  if (((ctx->epoch_calls * EPOCH_STRIDE) & 0x3fffffffull) < EPOCH_STRIDE && ctx->tile_state.p && !ctx->capturing)
  {
    CK(cudaMemsetAsync(ctx->tile_state.p, 0, ctx->tile_state.cap, ctx->stream));
    if (ctx->tile_state2.p)
      CK(cudaMemsetAsync(ctx->tile_state2.p, 0, ctx->tile_state2.cap, ctx->stream));
    if (ctx->tile_state_b.p)
      CK(cudaMemsetAsync(ctx->tile_state_b.p, 0, ctx->tile_state_b.cap, ctx->stream));
    if (ctx->tile_state2_b.p)
      CK(cudaMemsetAsync(ctx->tile_state2_b.p, 0, ctx->tile_state2_b.cap, ctx->stream));
  }
  ctx->pdl_chain = false;  // whatever precedes an API call on the stream (the caller's own work included) completes first
  if (scan_params)
    ENSURE(ctx->scratch_d, 1024);  // MinMax + VgLayout of the filter (voxelgrid.cu)
  LAUNCH(k_begin_call, 1, 32, 0, ctx->d_counters.as<unsigned long long>(), zero_scan_counters ? 1 : 0, scan_params ? (seed_now ? 1 : 2) : 0, ctx->score.as<float>(), ctx->g,
         ctx->dyn.as<ScanDyn>(), scan_params ? scan_params->score_point : 0.0, ctx->col_dirty.as<uint8_t>(), ctx->scratch_d.as<MinMax>());
  return 0;
}

struct FillJobs
{
  FillJob j[4];
};
__global__ void __launch_bounds__(256) k_fill_words(const FillJobs jobs)
{
  pdl_enter();
  const FillJob J = jobs.j[blockIdx.y];
  if (!J.p)
    return;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  // scalar head up to 16-byte alignment, uint4 body, scalar tail
  size_t head = ((16 - ((size_t)J.p & 15)) & 15) / 4;
  if (head > J.n_words)
    head = J.n_words;
  const size_t n4 = (J.n_words - head) / 4;
  uint4* p4 = reinterpret_cast<uint4*>(J.p + head);
  const uint4 v4 = make_uint4(J.value, J.value, J.value, J.value);
  for (size_t i = tid; i < n4; i += nth)
    p4[i] = v4;
  for (size_t i = tid; i < head; i += nth)
    J.p[i] = J.value;
  for (size_t i = head + n4 * 4 + tid; i < J.n_words; i += nth)
    J.p[i] = J.value;
}
int vf_fill(vofod_ctx* ctx, const FillJob* jobs, int n_jobs)
{
  FillJobs js = {};
  size_t most = 0;
  int k = 0;
  for (int i = 0; i < n_jobs && k < 4; i++)
    if (jobs[i].p && jobs[i].n_words)
    {
      js.j[k++] = jobs[i];
      if (jobs[i].n_words > most)
        most = jobs[i].n_words;
    }
  if (k == 0)
    return 0;
  LAUNCH(k_fill_words, dim3((unsigned)vf_blocks(ctx, most / 4 + 1, 256, 8), (unsigned)k), 256, 0, js);
  return 0;
}

int vf_dyn_push(vofod_ctx* ctx)
{
  CK(cudaMemcpyAsync(ctx->dyn.p, ctx->h_dyn, sizeof(ScanDyn), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

__global__ void k_fill_f32(float* __restrict__ p, const float v, const size_t n)
{
  pdl_enter();
  const size_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4 v4 = make_float4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    p4[i] = v4;
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void k_flags_to_f32(const uint8_t* __restrict__ f, float* __restrict__ out, const size_t n)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)f[i];
}
__global__ void k_f32_to_flags(const float* __restrict__ in, uint8_t* __restrict__ f, const size_t n)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    f[i] = (uint8_t)in[i];
}

// nVoxelsOver (voxel_map.cpp:216-222): streaming count of val > thr
__global__ void __launch_bounds__(256) k_count_over(const float* __restrict__ p, const size_t n, const float thr, unsigned long long* out)
{
  pdl_enter();
  unsigned cnt = 0;
  const size_t n4 = n / 4;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
  {
    const float4 v = __ldg(p4 + i);
    cnt += (v.x > thr) + (v.y > thr) + (v.z > thr) + (v.w > thr);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    cnt += p[i] > thr;
  cnt = prims::warp_sum(cnt);
  __shared__ unsigned ws[8];
  if ((threadIdx.x & 31) == 0)
    ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned t = 0;
    for (int i = 0; i < 8; i++)
      t += ws[i];
    if (t)
      atomicAdd(out, (unsigned long long)t);
  }
}

// nVoxelsOver restricted to the columns that can hold a value above the threshold (see vofod_ctx::col_dirty)
// blockIdx.y selects a chunk of 32 z-levels: a thread walking a whole column is a chain of ~20 dependent memory round trips,
// which (not bandwidth) set the pace of this kernel
__global__ void __launch_bounds__(128) k_count_over_cols(const float* __restrict__ p, const Geom g, const uint8_t* __restrict__ dirty, const float thr,
                                                         unsigned long long* out)
{
  pdl_enter();
  const int ncol = g.st_size[0] * g.st_size[1];
  const size_t sxy = (size_t)ncol;
  const int z_lo = blockIdx.y * DIRTY_ZC, z_hi = min(z_lo + DIRTY_ZC, g.st_size[2]);
  unsigned cnt = 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x)
  {
    if ((dirty && !dirty[(size_t)blockIdx.y * ncol + c]) || !column_owned(g, c % g.st_size[0], c / g.st_size[0]))
      continue;
    for (int z0 = z_lo; z0 < z_hi; z0 += 16)
    {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++)
        v[k] = (z0 + k < z_hi) ? p[(size_t)c + (size_t)(z0 + k) * sxy] : __int_as_float(0xff800000);
#pragma unroll
      for (int k = 0; k < 16; k++)
        cnt += v[k] > thr;
    }
  }
  cnt = prims::warp_sum(cnt);
  if ((threadIdx.x & 31) == 0 && cnt)
    atomicAdd(out, (unsigned long long)cnt);
}

__global__ void k_set_inf(float* __restrict__ score, const Geom g, const float* __restrict__ xyz, const size_t n, uint8_t* __restrict__ col_dirty)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = coord_to_idx1(xyz[3 * i], g.off[0], g.inv), y = coord_to_idx1(xyz[3 * i + 1], g.off[1], g.inv), z = coord_to_idx1(xyz[3 * i + 2], g.off[2], g.inv);
    if (!in_limits_idx(g, x, y, z))
      continue;
    const long long ci = cell_index(g, x, y, z);
    if (ci >= 0)
    {
      score[ci] = __int_as_float(0x7f800000);
      col_dirty[dirty_index(g, x - g.st_lo[0], y - g.st_lo[1], z - g.st_lo[2])] = 1;
    }
  }
}

__global__ void k_has_close_to(const float* __restrict__ score, const Geom g, const float* __restrict__ xyz, const int stride, const size_t n, const float max_dist,
                               const float thr, uint8_t* __restrict__ out)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = has_close_to(score, g, xyz[stride * i], xyz[stride * i + 1], xyz[stride * i + 2], max_dist, thr);
}

// isFloatingIdx (voxel_map.cpp:497-516)
__global__ void k_is_floating(const float* __restrict__ score, const Geom g, const float* __restrict__ xyz, const size_t n, const float thr, uint8_t* __restrict__ out)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = coord_to_idx1(xyz[3 * i], g.off[0], g.inv), y = coord_to_idx1(xyz[3 * i + 1], g.off[1], g.inv), z = coord_to_idx1(xyz[3 * i + 2], g.off[2], g.inv);
    bool ret = true;
    if (x <= 0 || y <= 0 || z <= 0 || x >= g.size[0] - 1 || y >= g.size[1] - 1 || z >= g.size[2] - 1)
      ret = false;
    else
      for (int zi = z - 1; zi <= z + 1 && ret; zi++)
        for (int yi = y - 1; yi <= y + 1 && ret; yi++)
          for (int xi = x - 1; xi <= x + 1; xi++)
          {
            const long long ci = cell_index(g, xi, yi, zi);
            if (ci >= 0 && score[ci] > thr)
            {
              ret = false;
              break;
            }
          }
    out[i] = ret;
  }
}

// forEachRay (voxel_map.cpp:229-263) for ONE ray, returning the callback sequence
__global__ void k_trace_ray(const Geom g, const float sx, const float sy, const float sz, const float dx, const float dy, const float dz, const float length,
                            float* __restrict__ ddist_out, int* __restrict__ idx3_out, const size_t cap, unsigned long long* n_out)
{
  pdl_enter();
  if (blockIdx.x || threadIdx.x)
    return;
  const float start[3] = {sx, sy, sz}, dir[3] = {dx, dy, dz};
  int step[3], cur[3];
  float tdelta[3], tmax[3], absdir[3];
  int last[3];
  for (int a = 0; a < 3; a++)
  {
    absdir[a] = fabsf(dir[a]);
    step[a] = (dir[a] > 0.0f) - (dir[a] < 0.0f);
    tdelta[a] = (1.0f / absdir[a]) * g.vs;
    cur[a] = coord_to_idx1(start[a], g.off[a], g.inv);
    const float ctr = idx_to_coord1(cur[a], g.off[a], g.vs) - start[a];
    tmax[a] = (g.half + (float)step[a] * ctr) / absdir[a];
    last[a] = step[a] > 0 ? g.size[a] - 1 : 0;
  }
  float prev = 0.0f;
  size_t n = 0;
  while (prev < length)
  {
    int i = 0;
    float dist = tmax[0];
    if (tmax[1] < dist) { dist = tmax[1]; i = 1; }
    if (tmax[2] < dist) { dist = tmax[2]; i = 2; }
    const float ddist = (length < dist ? length : dist) - prev;
    if (n < cap)
    {
      ddist_out[n] = ddist;
      idx3_out[3 * n] = cur[0]; idx3_out[3 * n + 1] = cur[1]; idx3_out[3 * n + 2] = cur[2];
    }
    n++;
    prev = dist;
    if (cur[i] == last[i])
      break;
    cur[i] += step[i];
    tmax[i] += tdelta[i];
    if (n > (size_t)1 << 24)
      break;  // watchdog
  }
  *n_out = n;
}

// getSubmapCopy (voxel_map.cpp:547-584)
__global__ void k_submap_copy(const float* __restrict__ score, const Geom g, const int nx, const int ny, const int nz, const int sx, const int sy, const int sz,
                              float* __restrict__ out)
{
  pdl_enter();
  const size_t n = (size_t)sx * sy * sz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = (int)(i % sx), y = (int)((i / sx) % sy), z = (int)(i / ((size_t)sx * sy));
    const long long ci = cell_index(g, x + nx, y + ny, z + nz);
    out[i] = ci >= 0 ? score[ci] : 0.0f;
  }
}

// voxelsAsPC / voxelsAsVoxelPC (voxel_map.cpp:157-212): ordered compaction, emission order x-outer, y, z-inner.
// pass 1: one thread per (x,y) column counts its matches (adjacent threads = adjacent x => coalesced reads)
__global__ void k_compact_count(const float* __restrict__ score, const Geom g, const float thr, const int greater, uint32_t* __restrict__ colcnt,
                                const uint8_t* __restrict__ dirty)
{
  pdl_enter();
  const int ncol = g.st_size[0] * g.st_size[1];
  const size_t sxy = (size_t)g.st_size[0] * g.st_size[1];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x)
  {
    const int x = c % g.st_size[0], y = c / g.st_size[0];
    uint32_t cnt = 0;
    if (column_owned(g, x, y))
      for (int z_lo = 0; z_lo < g.st_size[2]; z_lo += DIRTY_ZC)
      {
        if (dirty && !dirty[(size_t)(z_lo / DIRTY_ZC) * ncol + c])
          continue;
        const int z_hi = min(z_lo + DIRTY_ZC, g.st_size[2]);
        for (int z = z_lo; z < z_hi; z++)
          cnt += ((score[(size_t)c + (size_t)z * sxy] > thr) == (greater != 0));
      }
    colcnt[(size_t)x * g.st_size[1] + y] = cnt;  // x-major so that the scan runs in emission order
  }
}
// pass 2: emit.  Columns without a match (the vast majority) are skipped from their count alone; the others are read in
// batches of 16 independent loads before the (order-preserving, hence serial) emission.
__global__ void k_compact_emit(const float* __restrict__ score, const Geom g, const float thr, const int greater, const int metric, const uint32_t* __restrict__ colcnt,
                               const uint32_t* __restrict__ coloff, vofod_xyzi* __restrict__ out, const size_t cap)
{
  pdl_enter();
  const int ncol = g.st_size[0] * g.st_size[1];
  const size_t sxy = (size_t)g.st_size[0] * g.st_size[1];
  const int sz = g.st_size[2];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x)
  {
    const int x = c % g.st_size[0], y = c / g.st_size[0];
    const size_t t = (size_t)x * g.st_size[1] + y;
    if (colcnt[t] == 0)
      continue;
    size_t o = coloff[t];
    for (int z0 = 0; z0 < sz; z0 += 16)
    {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++)
        v[k] = (z0 + k < sz) ? score[(size_t)c + (size_t)(z0 + k) * sxy] : __int_as_float(0x7fc00000);
#pragma unroll
      for (int k = 0; k < 16; k++)
      {
        if (z0 + k < sz && ((v[k] > thr) == (greater != 0)))
        {
          if (o < cap)
          {
            vofod_xyzi p;
            const int gx = x + g.st_lo[0], gy = y + g.st_lo[1], gz = z0 + k + g.st_lo[2];
            if (metric)
            {
              p.x = idx_to_coord1(gx, g.off[0], g.vs);
              p.y = idx_to_coord1(gy, g.off[1], g.vs);
              p.z = idx_to_coord1(gz, g.off[2], g.vs);
            } else
            {
              p.x = (float)gx; p.y = (float)gy; p.z = (float)gz;
            }
            p.intensity = v[k];
            out[o] = p;
          }
          o++;
        }
      }
    }
  }
}

// shared with sepclusters.cu.  Two modes:
//   host_total != NULL : exact — one 8-byte read-back sizes the output (staged API, first scans)
//   host_total == NULL : `cap` rows are reserved up front, no host round trip (graph replay); *d_total still receives the
//                        true count, rows beyond cap are dropped and the caller must check d_total <= cap
int vf_compact_over_dev(vofod_ctx* ctx, float thr, int greater, int metric, DevBuf& out, unsigned long long* d_total, size_t* host_total, size_t cap,
                        const vofod_params* p)
{
  const Geom& g = ctx->g;
  const size_t ncol = (size_t)g.st_size[0] * g.st_size[1];
  ENSURE(ctx->sep_colcnt, prims::padded(ncol) * sizeof(uint32_t));
  ENSURE(ctx->sep_coloff, prims::padded(ncol) * sizeof(uint32_t));
  ctx->sep_prefilled = 0;  // sep_colcnt is about to be overwritten
  const uint8_t* dirty = greater ? vf_dirty_cols(ctx, thr, p) : nullptr;
  LAUNCH(k_compact_count, vf_blocks(ctx, ncol, 128, 16), 128, 0, ctx->score.as<float>(), g, thr, greater, ctx->sep_colcnt.as<uint32_t>(), dirty);
  RET(prims::scan_excl_u32(ctx, ctx->sep_colcnt.as<uint32_t>(), ctx->sep_coloff.as<uint32_t>(), nullptr, ncol, d_total));
  if (host_total)
  {
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *host_total = (size_t)total;
    cap = (size_t)total;
    if (total == 0)
      return 0;
  }
  ENSURE(out, prims::padded(cap) * sizeof(vofod_xyzi));
  LAUNCH(k_compact_emit, vf_blocks(ctx, ncol, 128, 16), 128, 0, ctx->score.as<float>(), g, thr, greater, metric, ctx->sep_colcnt.as<uint32_t>(),
         ctx->sep_coloff.as<uint32_t>(), out.as<vofod_xyzi>(), cap);
  return 0;
}

// ======================================================================================================
// C ABI — map
// ======================================================================================================
#define NEED_CTX()                   \
  if (!ctx)                          \
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL"); \
  CK(cudaSetDevice(ctx->device))
#define NEED_MAP()                   \
  NEED_CTX();                        \
  if (!ctx->map_ready)               \
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized (call vofod_map_resize / vofod_reset first)")

int vf_map_alloc(vofod_ctx* ctx)
{
  Geom& g = ctx->g;
  const long long n = geom_cells(g);
  if (n <= 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "map has no cells");
  ENSURE(ctx->score, (size_t)n * sizeof(float) + 64);
  ENSURE(ctx->flags, (size_t)n + 64);
  ENSURE(ctx->col_dirty, (size_t)g.st_size[0] * g.st_size[1] * dirty_chunks(g) + 64);
  CK(cudaMemsetAsync(ctx->col_dirty.p, 0, (size_t)g.st_size[0] * g.st_size[1] * dirty_chunks(g), ctx->stream));
  ctx->col_all_dirty = true;  // the grid contents are unspecified until the first setTo
  CK(cudaMemsetAsync(ctx->flags.p, 0, (size_t)n, ctx->stream));
  ctx->flags_full_dirty = false;
  // every counter except the persistent generations (exploreToGround stamps, look-back states keep old values)
  CK(cudaMemsetAsync(ctx->d_counters.p, 0, CNT_FIRST_PERSISTENT * sizeof(unsigned long long), ctx->stream));
  ctx->alloc_gen++;
  ctx->win_valid = false;
  ctx->acc_has_data = false;
  ctx->ray_pending = false;
  ctx->sep_pending = false;
  ctx->map_ready = true;
  return 0;
}

extern "C" {

int vofod_map_resize_idx(vofod_ctx* ctx, const float offset[3], const int32_t sizes[3], float voxel_size)
{
  NEED_CTX();
  if (!offset || !sizes || !(voxel_size > 0.0f) || sizes[0] <= 0 || sizes[1] <= 0 || sizes[2] <= 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_map_resize_idx: bad arguments");
  if ((long long)sizes[0] * sizes[1] * sizes[2] > (long long)INT32_MAX)
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "voxel map larger than INT32_MAX cells (reference indexes with int, voxel_map.cpp:46)");
  Geom& g = ctx->g;
  g.vs = voxel_size;
  g.half = voxel_size / 2.0f;
  g.inv = 1.0f / voxel_size;
  for (int a = 0; a < 3; a++)
  {
    g.off[a] = offset[a];
    g.size[a] = sizes[a];
    g.st_lo[a] = 0;
    g.st_size[a] = sizes[a];
  }
  g.slab_axis = 0;
  g.own_lo = 0;
  g.own_hi = sizes[0];
  ctx->slab_on = false;
  ctx->cfg_voxel_size = voxel_size;
  return vf_map_alloc(ctx);
}

int vofod_map_resize(vofod_ctx* ctx, const float center[3], const float dims[3], float voxel_size)
{
  NEED_CTX();
  if (!center || !dims || !(voxel_size > 0.0f))
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_map_resize: bad arguments");
  // voxel_map.cpp:11-19 (host fp32, separately rounded ops)
  const float inv = 1.0f / voxel_size;
  float offset[3];
  int32_t sizes[3];
  for (int a = 0; a < 3; a++)
  {
    volatile float half_dim = dims[a] / 2.0f;
    offset[a] = center[a] - half_dim;
    volatile float prod = inv * dims[a];
    sizes[a] = (int32_t)ceilf(prod) + 1;
  }
  return vofod_map_resize_idx(ctx, offset, sizes, voxel_size);
}

int vofod_reset(vofod_ctx* ctx, const vofod_params* p, float voxel_size)
{
  NEED_CTX();
  if (!p)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_reset: params is NULL");
  volatile float oz = p->oparea_offset[2] + p->oparea_size[2] / 2.0f;  // vofod_nodelet.cpp:212
  const float center[3] = {p->oparea_offset[0], p->oparea_offset[1], oz};
  RET(vofod_map_resize(ctx, center, p->oparea_size, voxel_size));
  RET(vofod_map_set_to(ctx, VOFOD_MAP_SCORE, p->score_init));
  ctx->detection_its = 0;
  ctx->background_pts_sufficient = false;
  ctx->sure_background_sufficient = false;
  ctx->last_detection_id = 0;
  return VOFOD_OK;
}

int vofod_map_info_get(const vofod_ctx* ctx, vofod_map_info* out)
{
  if (!ctx || !out)
    return VOFOD_E_INVALID;
  if (!ctx->map_ready)
    return VOFOD_E_STATE;
  const Geom& g = ctx->g;
  for (int a = 0; a < 3; a++)
  {
    out->offset[a] = g.off[a];
    out->sizes[a] = g.size[a];
  }
  out->voxel_size = g.vs;
  out->n_cells = (uint64_t)g.size[0] * g.size[1] * g.size[2];
  out->slab_axis = g.slab_axis;
  out->slab_lo = g.own_lo;
  out->slab_hi = g.own_hi;
  for (int a = 0; a < 3; a++)
  {
    out->storage_lo[a] = g.st_lo[a];
    out->storage_size[a] = g.st_size[a];
  }
  out->_pad = 0;
  return VOFOD_OK;
}

int vofod_map_set_to(vofod_ctx* ctx, int which, float value)
{
  NEED_MAP();
  FLUSH_PENDING();
  const size_t n = (size_t)geom_cells(ctx->g);
  if (which == VOFOD_MAP_SCORE)
  {
    LAUNCH(k_fill_f32, vf_blocks(ctx, n / 4 + 1, 256), 256, 0, ctx->score.as<float>(), value, n);
    CK(cudaMemsetAsync(ctx->col_dirty.p, 0, (size_t)ctx->g.st_size[0] * ctx->g.st_size[1] * dirty_chunks(ctx->g), ctx->stream));
    ctx->col_all_dirty = !(value == value) || value == __builtin_inff();  // NaN / +inf fills defeat the bound
    ctx->untouched_max = value;
  } else if (which == VOFOD_MAP_FLAGS)
  {
    CK(cudaMemsetAsync(ctx->flags.p, (int)(uint8_t)value, n, ctx->stream));
    CK(cudaMemsetAsync(vf_cnt(ctx, CNT_FLAGGED), 0, 2 * sizeof(unsigned long long), ctx->stream));
    ctx->flags_full_dirty = value != 0.0f;
  } else if (which == VOFOD_MAP_RAYCAST)
  {
    if (value != 0.0f)
      return vf_fail(ctx, VOFOD_E_INVALID, "the raycast accumulator can only be cleared");
    if (ctx->win_valid)
      CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_total_bytes, ctx->stream));
    ctx->acc_has_data = false;
    ctx->ray_pending = false;
  } else
    return vf_fail(ctx, VOFOD_E_INVALID, "bad map id %d", which);
  return VOFOD_OK;
}

int vofod_map_set_inf(vofod_ctx* ctx, const float* xyz, size_t n)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (n == 0)
    return VOFOD_OK;
  if (!xyz)
    return vf_fail(ctx, VOFOD_E_INVALID, "xyz is NULL");
  ENSURE(ctx->scratch_a, n * 12);
  CK(cudaMemcpyAsync(ctx->scratch_a.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_set_inf, vf_blocks(ctx, n, 256), 256, 0, ctx->score.as<float>(), ctx->g, ctx->scratch_a.as<float>(), n, ctx->col_dirty.as<uint8_t>());
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}


int vofod_map_download(vofod_ctx* ctx, int which, float* host, size_t n_cells)
{
  NEED_MAP();
  FLUSH_PENDING();
  const size_t n = (size_t)geom_cells(ctx->g);
  if (!host || n_cells != n)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_map_download: expected %zu cells", n);
  if (which == VOFOD_MAP_SCORE)
    CK(cudaMemcpyAsync(host, ctx->score.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  else if (which == VOFOD_MAP_FLAGS)
  {
    ENSURE(ctx->scratch_a, n * 4);
    LAUNCH(k_flags_to_f32, vf_blocks(ctx, n, 256), 256, 0, ctx->flags.as<uint8_t>(), ctx->scratch_a.as<float>(), n);
    CK(cudaMemcpyAsync(host, ctx->scratch_a.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  } else if (which == VOFOD_MAP_RAYCAST)
  {
    ENSURE(ctx->scratch_a, n * 4);
    RET(vf_raycast_expand(ctx, nullptr, ctx->scratch_a.as<float>()));
    CK(cudaMemcpyAsync(host, ctx->scratch_a.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  } else
    return vf_fail(ctx, VOFOD_E_INVALID, "bad map id %d", which);
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_upload(vofod_ctx* ctx, int which, const float* host, size_t n_cells)
{
  NEED_MAP();
  FLUSH_PENDING();
  const size_t n = (size_t)geom_cells(ctx->g);
  if (!host || n_cells != n)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_map_upload: expected %zu cells", n);
  if (which == VOFOD_MAP_SCORE)
  {
    CK(cudaMemcpyAsync(ctx->score.p, host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->col_all_dirty = true;
  }
  else if (which == VOFOD_MAP_FLAGS)
  {
    ENSURE(ctx->scratch_a, n * 4);
    CK(cudaMemcpyAsync(ctx->scratch_a.p, host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_f32_to_flags, vf_blocks(ctx, n, 256), 256, 0, ctx->scratch_a.as<float>(), ctx->flags.as<uint8_t>(), n);
    ctx->flags_full_dirty = true;
  } else
    return vf_fail(ctx, VOFOD_E_INVALID, "only the score and flags grids can be uploaded");
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_get(vofod_ctx* ctx, int which, int ix, int iy, int iz, float* value)
{
  NEED_MAP();
  FLUSH_PENDING();
  const Geom& g = ctx->g;
  // std::vector::at semantics (voxel_map.cpp:82,117): out of range is an error, not UB
  if (!value || ix < 0 || iy < 0 || iz < 0 || ix >= g.size[0] || iy >= g.size[1] || iz >= g.size[2])
    return vf_fail(ctx, VOFOD_E_INVALID, "index [%d,%d,%d] out of range", ix, iy, iz);
  const long long lx = ix - g.st_lo[0], ly = iy - g.st_lo[1], lz = iz - g.st_lo[2];
  if (lx < 0 || ly < 0 || lz < 0 || lx >= g.st_size[0] || ly >= g.st_size[1] || lz >= g.st_size[2])
    return vf_fail(ctx, VOFOD_E_INVALID, "index [%d,%d,%d] not held by this slab", ix, iy, iz);
  const size_t ci = (size_t)lx + (size_t)ly * g.st_size[0] + (size_t)lz * g.st_size[0] * g.st_size[1];
  if (which == VOFOD_MAP_SCORE)
    CK(cudaMemcpyAsync(value, ctx->score.as<float>() + ci, 4, cudaMemcpyDeviceToHost, ctx->stream));
  else if (which == VOFOD_MAP_FLAGS)
  {
    uint8_t f = 0;
    CK(cudaMemcpyAsync(&f, ctx->flags.as<uint8_t>() + ci, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *value = (float)f;
    return VOFOD_OK;
  } else
    return vf_fail(ctx, VOFOD_E_INVALID, "bad map id %d", which);
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_set(vofod_ctx* ctx, int which, int ix, int iy, int iz, float value)
{
  NEED_MAP();
  FLUSH_PENDING();
  const Geom& g = ctx->g;
  if (ix < 0 || iy < 0 || iz < 0 || ix >= g.size[0] || iy >= g.size[1] || iz >= g.size[2])
    return vf_fail(ctx, VOFOD_E_INVALID, "index [%d,%d,%d] out of range", ix, iy, iz);
  const long long lx = ix - g.st_lo[0], ly = iy - g.st_lo[1], lz = iz - g.st_lo[2];
  if (lx < 0 || ly < 0 || lz < 0 || lx >= g.st_size[0] || ly >= g.st_size[1] || lz >= g.st_size[2])
    return VOFOD_OK;  // another slab owns it
  const size_t ci = (size_t)lx + (size_t)ly * g.st_size[0] + (size_t)lz * g.st_size[0] * g.st_size[1];
  if (which == VOFOD_MAP_SCORE)
  {
    CK(cudaMemcpyAsync(ctx->score.as<float>() + ci, &value, 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->col_all_dirty = true;
  }
  else if (which == VOFOD_MAP_FLAGS)
  {
    const uint8_t f = (uint8_t)value;
    CK(cudaMemcpyAsync(ctx->flags.as<uint8_t>() + ci, &f, 1, cudaMemcpyHostToDevice, ctx->stream));
    ctx->flags_full_dirty = true;
  } else
    return vf_fail(ctx, VOFOD_E_INVALID, "bad map id %d", which);
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}
}  // extern "C"

// Columns worth reading for a "value > thr" pass, or NULL when every column must be read.  A column nobody raised holds
// only the fill value and what the decaying updates made of it: ray apply and the background-cluster decay mix towards
// score_ray, classification writes thr_frontiers.  If thr is strictly above all three, such a column cannot match.
const uint8_t* vf_dirty_cols(vofod_ctx* ctx, float thr, const vofod_params* p)
{
  if (ctx->col_all_dirty || !p || !ctx->col_dirty.p)
    return nullptr;
  float bound = ctx->untouched_max;
  if ((float)p->score_ray > bound) bound = (float)p->score_ray;
  if ((float)p->thr_frontiers > bound) bound = (float)p->thr_frontiers;
  return thr > bound ? ctx->col_dirty.as<uint8_t>() : nullptr;
}

int vf_count_over_dev(vofod_ctx* ctx, float thr, unsigned long long* d_out, const vofod_params* p)
{
  const size_t n = (size_t)geom_cells(ctx->g);
  if (!(ctx->scan_prezero && d_out == vf_cnt(ctx, CNT_NBG)))
    CK(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), ctx->stream));
  const uint8_t* dirty = vf_dirty_cols(ctx, thr, p);
  if (dirty || ctx->slab_on)  // a slab counts its own range only (halo columns belong to the neighbour)
    LAUNCH(k_count_over_cols, dim3((unsigned)vf_blocks(ctx, (size_t)ctx->g.st_size[0] * ctx->g.st_size[1], 128, 4), (unsigned)dirty_chunks(ctx->g)), 128, 0,
           ctx->score.as<float>(), ctx->g, dirty, thr, d_out);
  else
    LAUNCH(k_count_over, vf_blocks(ctx, n / 4 + 1, 256, 8), 256, 0, ctx->score.as<float>(), n, thr, d_out);
  return 0;
}

extern "C" {

int vofod_map_count_over(vofod_ctx* ctx, float threshold, uint64_t* out)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (!out)
    return vf_fail(ctx, VOFOD_E_INVALID, "out is NULL");
  RET(vf_count_over_dev(ctx, threshold, vf_cnt(ctx, CNT_NBG)));
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, vf_cnt(ctx, CNT_NBG), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = v;
  return VOFOD_OK;
}

int vofod_map_compact_over(vofod_ctx* ctx, float threshold, int greater_than, int metric, vofod_xyzi* out, size_t cap, size_t* n)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (!n)
    return vf_fail(ctx, VOFOD_E_INVALID, "n is NULL");
  size_t total = 0;
  RET(vf_begin_call(ctx));
  RET(vf_compact_over_dev(ctx, threshold, greater_than, metric, ctx->sep_raw, vf_cnt(ctx, CNT_SEP_K), &total, 0, nullptr));
  *n = total;
  if (total > cap || (!out && total))
    return vf_fail(ctx, VOFOD_E_CAPACITY, "compact_over: need capacity %zu", total);
  if (total)
    CK(cudaMemcpyAsync(out, ctx->sep_raw.p, total * sizeof(vofod_xyzi), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_has_close_to(vofod_ctx* ctx, const float* xyz, size_t n, float max_dist, float threshold, uint8_t* out)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (n == 0)
    return VOFOD_OK;
  if (!xyz || !out)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL buffer");
  ENSURE(ctx->scratch_a, n * 12);
  ENSURE(ctx->scratch_b, n);
  CK(cudaMemcpyAsync(ctx->scratch_a.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_has_close_to, vf_blocks(ctx, n, 128, 16), 128, 0, ctx->score.as<float>(), ctx->g, ctx->scratch_a.as<float>(), 3, n, max_dist, threshold,
         ctx->scratch_b.as<uint8_t>());
  CK(cudaMemcpyAsync(out, ctx->scratch_b.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_is_floating(vofod_ctx* ctx, const float* xyz, size_t n, float threshold, uint8_t* out)
{
  NEED_MAP();
  FLUSH_PENDING();
  if (n == 0)
    return VOFOD_OK;
  if (!xyz || !out)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL buffer");
  ENSURE(ctx->scratch_a, n * 12);
  ENSURE(ctx->scratch_b, n);
  CK(cudaMemcpyAsync(ctx->scratch_a.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_is_floating, vf_blocks(ctx, n, 128, 16), 128, 0, ctx->score.as<float>(), ctx->g, ctx->scratch_a.as<float>(), n, threshold, ctx->scratch_b.as<uint8_t>());
  CK(cudaMemcpyAsync(out, ctx->scratch_b.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_map_trace_ray(vofod_ctx* ctx, const float start[3], const float dir[3], float length, float* ddist, int32_t* idx3, size_t cap, size_t* n)
{
  NEED_MAP();
  if (!start || !dir || !n || (cap && (!ddist || !idx3)))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL buffer");
  ENSURE(ctx->scratch_a, cap * 4 + 16);
  ENSURE(ctx->scratch_b, cap * 12 + 16);
  LAUNCH(k_trace_ray, 1, 32, 0, ctx->g, start[0], start[1], start[2], dir[0], dir[1], dir[2], length, ctx->scratch_a.as<float>(), ctx->scratch_b.as<int>(), cap,
         vf_cnt(ctx, CNT_SCRATCH0));
  unsigned long long cnt = 0;
  CK(cudaMemcpyAsync(&cnt, vf_cnt(ctx, CNT_SCRATCH0), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = (size_t)cnt;
  const size_t k = cnt < cap ? (size_t)cnt : cap;
  if (k)
  {
    CK(cudaMemcpyAsync(ddist, ctx->scratch_a.p, k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(idx3, ctx->scratch_b.p, k * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return cnt > cap ? vf_fail(ctx, VOFOD_E_CAPACITY, "trace_ray: need capacity %llu", cnt) : VOFOD_OK;
}

int vofod_map_submap_copy(vofod_ctx* ctx, const float min_pt[3], const float max_pt[3], int inflate, float* out, size_t cap, int32_t sizes_out[3], float offset_out[3])
{
  NEED_MAP();
  FLUSH_PENDING();
  if (!min_pt || !max_pt || !sizes_out || !offset_out)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL buffer");
  const Geom& g = ctx->g;
  int mn[3], mx[3];
  for (int a = 0; a < 3; a++)
  {
    volatile float d0 = min_pt[a] - g.off[a];
    volatile float q0 = d0 * g.inv;
    volatile float d1 = max_pt[a] - g.off[a];
    volatile float q1 = d1 * g.inv;
    mn[a] = (int)floorf(q0) - inflate;
    mx[a] = (int)floorf(q1) + inflate;
    mn[a] = mn[a] < 0 ? 0 : (mn[a] > g.size[a] - 1 ? g.size[a] - 1 : mn[a]);
    mx[a] = mx[a] < 0 ? 0 : (mx[a] > g.size[a] - 1 ? g.size[a] - 1 : mx[a]);
    sizes_out[a] = mx[a] - mn[a] + 1;
    volatile float c0 = ((float)mn[a] + 0.5f);
    volatile float c1 = c0 * g.vs;
    volatile float c2 = c1 + g.off[a];
    volatile float h = g.vs / 2.0f;
    offset_out[a] = c2 - h;
  }
  const size_t n = (size_t)sizes_out[0] * sizes_out[1] * sizes_out[2];
  if (n > cap || !out)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "submap_copy: need capacity %zu", n);
  ENSURE(ctx->scratch_a, n * 4);
  LAUNCH(k_submap_copy, vf_blocks(ctx, n, 256), 256, 0, ctx->score.as<float>(), g, mn[0], mn[1], mn[2], sizes_out[0], sizes_out[1], sizes_out[2],
         ctx->scratch_a.as<float>());
  CK(cudaMemcpyAsync(out, ctx->scratch_a.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

// ---- sensor ----------------------------------------------------------------------------------------
int vofod_set_sensor(vofod_ctx* ctx, int W, int H, const float* dirs, const float* offs, const uint8_t* mask)
{
  NEED_CTX();
  if (W <= 0 || H <= 0 || !dirs)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_set_sensor: bad arguments");
  const size_t n = (size_t)W * H;
  if (n >= (size_t(1) << 20))
    return vf_fail(ctx, VOFOD_E_INVALID, "at most 2^20-1 rays per scan are supported by the packed accumulator (got %zu)", n);
  std::vector<float> d4(n * 4), o4(n * 4, 0.0f);
  std::vector<uint8_t> m(n, 1);  // missing mask => all ones (vofod_nodelet.cpp:558)
  float max_off = 0.f;
  bool has_off = false;
  for (size_t i = 0; i < n; i++)
  {
    d4[4 * i] = dirs[3 * i]; d4[4 * i + 1] = dirs[3 * i + 1]; d4[4 * i + 2] = dirs[3 * i + 2]; d4[4 * i + 3] = 0.f;
    if (offs)
    {
      o4[4 * i] = offs[3 * i]; o4[4 * i + 1] = offs[3 * i + 1]; o4[4 * i + 2] = offs[3 * i + 2];
      const float l = sqrtf(offs[3 * i] * offs[3 * i] + offs[3 * i + 1] * offs[3 * i + 1] + offs[3 * i + 2] * offs[3 * i + 2]);
      if (l > max_off) max_off = l;
      if (l != 0.f) has_off = true;
    }
    if (mask)
      m[i] = mask[i];
  }
  ENSURE(ctx->lut_dir, n * 16);
  ENSURE(ctx->lut_off, n * 16);
  ENSURE(ctx->mask, n);
  CK(cudaMemcpyAsync(ctx->lut_dir.p, d4.data(), n * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->lut_off.p, o4.data(), n * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->mask.p, m.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->W = W;
  ctx->H = H;
  ctx->lut_has_off = has_off;
  ctx->lut_max_off = max_off;
  ctx->win_valid = false;
  return VOFOD_OK;
}

static const char* k_stage_names[VOFOD_N_STAGES] = {"range", "filtering", "clusterization", "close X far", "vmap update", "raycasting",
                                                    "raycast vmap update", "classification", "detections", "sep bg clusters", "readback", "total"};
const char* vofod_stage_name(int i) { return (i >= 0 && i < VOFOD_N_STAGES) ? k_stage_names[i] : ""; }
int vofod_stage_times(vofod_ctx* ctx, float ms[VOFOD_N_STAGES])
{
  if (!ctx || !ms)
    return VOFOD_E_INVALID;
  memcpy(ms, ctx->stage_ms, sizeof(ctx->stage_ms));
  return VOFOD_OK;
}
}  // extern "C"
