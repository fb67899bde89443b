// updateSeparatedBGClusters (vofod_nodelet.cpp:1126-1278) on the GPU: the background thread's pass that finds
// background voxels which are NOT attached to a large "sure" background body and decays them towards the ray score.
//
//   compaction of voxels > new_obstacles in the reference's x-outer/z-inner order (K11, ctx.cu)
//   -> VoxelGridCounted (voxelgrid.cu, incl. its input-slice quirk) -> Euclidean clustering in index units (cluster.cu)
//   -> per-cluster sum of sure counts -> decay scatter (K12).
// The reference applies map = w1*map + w2*ray sequentially, so a cell hit by several (voxel, offset) pairs is
// updated several times; every application is the SAME affine function, hence the result only depends on how
// often a cell is hit: each hit is applied as one atomic CAS round and the outcome is bit-identical.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "prims.cuh"

int vf_compact_over_dev(vofod_ctx* ctx, float thr, int greater, int metric, DevBuf& out, unsigned long long* d_total, size_t* host_total, size_t cap,
                        const vofod_params* p);  // ctx.cu
int vf_voxel_grid_counted_dev(vofod_ctx* ctx, const vofod_xyzi* d_in, const unsigned long long* d_n, size_t cap, float leaf, float thr, DevBuf& out);  // voxelgrid.cu

// :1174-1183 — n_sure[cluster] = std::accumulate(range, int 0)
__global__ void __launch_bounds__(256) k_sep_nsure(const vofod_vox* __restrict__ ds, const int* __restrict__ labels, const unsigned long long* __restrict__ d_k,
                                                   const size_t cap, int* __restrict__ nsure)
{
  const size_t k = prims::dev_count(d_k, cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < k; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool valid = i < k;
    const int l = valid ? labels[i] : -1 - (int)lane;
    const unsigned c = valid ? ds[i].count : 0u;
    // one atomic per (warp, cluster): the big background body would otherwise serialise every point on one word
    const unsigned grp = __match_any_sync(VOFOD_FULL, l);
    const unsigned sum = __reduce_add_sync(grp, c);
    if (valid && lane == (unsigned)(__ffs(grp) - 1))
      atomicAdd(reinterpret_cast<unsigned*>(nsure) + l, sum);
  }
}
// :1186-1206
__global__ void __launch_bounds__(256) k_sep_any(const int* __restrict__ labels, const int* __restrict__ nsure, const unsigned long long* __restrict__ d_k, const size_t cap,
                                                 const unsigned min_sure, unsigned long long* __restrict__ counters)
{
  const size_t k = prims::dev_count(d_k, cap);
  bool any = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (size_t)gridDim.x * blockDim.x)
    if (labels[i] == (int)i && (unsigned long long)(long long)nsure[i] >= (unsigned long long)min_sure)  // size_t(int) >= unsigned
      any = true;
  if (__any_sync(VOFOD_FULL, any) && (threadIdx.x & 31) == 0)
    counters[CNT_SEP_ANY_SURE] = 1ull;
}
__global__ void k_sep_state(unsigned long long* __restrict__ counters, const unsigned long long k_cap)
{
  const unsigned long long k = counters[CNT_SEP_K];
  if (k == 0ull || k > k_cap)
    return;  // empty cloud: the reference returns before touching the flag (:1155-1159); overflow: the host redoes the pass
  counters[CNT_STATE_SURE] = counters[CNT_SEP_ANY_SURE] ? 1ull : 0ull;  // :1196 / :1205
}

// K12 — :1244-1272
__global__ void __launch_bounds__(256) k_sep_decay(float* score, const Geom g, const vofod_vox* __restrict__ ds, const int* __restrict__ labels,
                                                   const int* __restrict__ nsure, const unsigned long long* __restrict__ d_k, const size_t cap,
                                                   const int3* __restrict__ offsets, const int n_off, const unsigned min_sure, const float w1, const float w2,
                                                   const float update_val, const unsigned long long* __restrict__ counters, const unsigned long long k_cap)
{
  if (counters[CNT_SEP_ANY_SURE] == 0ull || counters[CNT_SEP_K] > k_cap)
    return;  // :1192-1199 (and: list overflow => nothing is touched, the host redoes the pass with a larger list)
  const size_t k = prims::dev_count(d_k, cap);
  const size_t total = k * (size_t)n_off;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = t / n_off;
    const int o = (int)(t - i * n_off);
    if (!((unsigned)nsure[labels[i]] < min_sure))  // only the unsure clusters (:1246)
      continue;
    const vofod_vox v = ds[i];
    const int3 of = offsets[o];
    const int x = (int)v.x + of.x, y = (int)v.y + of.y, z = (int)v.z + of.z;  // cast<int>() truncation (:1252)
    if (!in_limits_idx(g, x, y, z))
      continue;
    const long long ci = cell_index(g, x, y, z);
    if (ci < 0 || !cell_owned(g, x, y, z))
      continue;
    unsigned* addr = reinterpret_cast<unsigned*>(score + ci);
    unsigned old = *addr;
    while (true)
    {
      const float m = __uint_as_float(old);
      const float nv = w1 * m + w2 * update_val;  // :1259
      const unsigned prev = atomicCAS(addr, old, __float_as_uint(nv));
      if (prev == old)
        break;
      old = prev;
    }
  }
}

// k_cap == 0: exact mode (one host read-back sizes the background-voxel list).  k_cap > 0: the list is capped at k_cap rows and
// nothing returns to the host; when the true count exceeds k_cap the pass leaves the map untouched and the caller, who sees
// CNT_SEP_K > k_cap in its read-back, repeats it in exact mode.
int vf_sepclusters_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t k_cap)
{
  if (p.sep_pause)
    return VOFOD_W_PAUSED;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  const float vs = ctx->cfg_voxel_size > 0.f ? ctx->cfg_voxel_size : ctx->g.vs;
  const float thr_new = (float)p.thr_new_obstacles;
  const float thr_sure = (float)p.thr_sure_obstacles;
  const unsigned min_sure = (unsigned)p.sep_min_sure_points;
  const float max_dist_idx = (float)(p.sep_max_bg_distance / (double)vs);  // :1142
  const int mv = (int)ceilf(max_dist_idx);                                 // :1143
  // :1146-1153 copy + voxelsAsVoxelPC (the private copy is unnecessary here: calls on a context are serialised)
  size_t K = k_cap;
  if (k_cap == 0)
  {
    RET(vf_compact_over_dev(ctx, thr_new, 1, 0, ctx->sep_raw, cnt + CNT_SEP_K, &K, 0, &p));
    if (K == 0)
      return VOFOD_W_EMPTY;  // :1155-1159
  } else
    RET(vf_compact_over_dev(ctx, thr_new, 1, 0, ctx->sep_raw, cnt + CNT_SEP_K, nullptr, k_cap, &p));
  const unsigned long long cap_guard = k_cap ? (unsigned long long)k_cap : ~0ull;
  const float lsz = (float)(mv - 1 > 0 ? mv - 1 : 0);  // :1163
  if (!(lsz > 0.0f))
    return vf_fail(ctx, VOFOD_E_INVALID, "sepclusters: max_bg_distance/voxel_size <= 1 gives a zero leaf size (the reference divides by it)");
  RET(vf_voxel_grid_counted_dev(ctx, ctx->sep_raw.as<vofod_xyzi>(), cnt + CNT_SEP_K, K, lsz, thr_sure, ctx->sep_ds));
  ENSURE(ctx->sep_labels, K * 4);
  ENSURE(ctx->sep_nsure, K * 4);
  RET(vf_cluster_dev(ctx, ctx->cl_bg, reinterpret_cast<const float*>(ctx->sep_ds.p), 4, cnt + CNT_SEP_KDS, K, (float)mv, ctx->sep_labels.as<int>(), cnt + CNT_SEP_NCL));
  CK(cudaMemsetAsync(ctx->sep_nsure.p, 0, K * 4, ctx->stream));
  CK(cudaMemsetAsync(cnt + CNT_SEP_ANY_SURE, 0, 8, ctx->stream));
  const int nb = vf_blocks(ctx, K, 256, 8);
  LAUNCH(k_sep_nsure, nb, 256, 0, ctx->sep_ds.as<vofod_vox>(), ctx->sep_labels.as<int>(), cnt + CNT_SEP_KDS, K, ctx->sep_nsure.as<int>());
  LAUNCH(k_sep_any, nb, 256, 0, ctx->sep_labels.as<int>(), ctx->sep_nsure.as<int>(), cnt + CNT_SEP_KDS, K, min_sure, cnt);
  LAUNCH(k_sep_state, 1, 1, 0, cnt, cap_guard);
  // :1219-1237 ball of offsets with Eigen's truncated integer norm (uploaded once per parameter change)
  if (ctx->sep_off_n < 0 || ctx->sep_off_mv != mv || ctx->sep_off_md != max_dist_idx)
  {
    std::vector<int3> offs;
    for (int x = -mv; x <= mv; x++)
      for (int y = -mv; y <= mv; y++)
        for (int z = -mv; z <= mv; z++)
        {
          const int nrm = (int)sqrt((double)(x * x + y * y + z * z));
          if ((float)nrm <= max_dist_idx)
            offs.push_back(make_int3(x, y, z));
        }
    ENSURE(ctx->sep_offsets, (offs.size() + 1) * sizeof(int3));
    if (!offs.empty())
      CK(cudaMemcpyAsync(ctx->sep_offsets.p, offs.data(), offs.size() * sizeof(int3), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // `offs` is pageable and dies here
    ctx->sep_off_n = (int)offs.size();
    ctx->sep_off_mv = mv;
    ctx->sep_off_md = max_dist_idx;
  }
  const size_t n_off = (size_t)ctx->sep_off_n;
  if (n_off == 0)
    return VOFOD_OK;
  const float dits = (float)(its_diff > 1 ? its_diff : 1);           // :1210-1217
  float w1 = powf(1.0f - 0.5f, dits);                                 // :1239-1241
  w1 = w1 < 0.0f ? 0.0f : (1.0f < w1 ? 1.0f : w1);
  volatile float w2v = 1.0f - w1;
  const float w2 = w2v;
  LAUNCH(k_sep_decay, vf_blocks(ctx, K * n_off, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, ctx->sep_ds.as<vofod_vox>(), ctx->sep_labels.as<int>(),
         ctx->sep_nsure.as<int>(), cnt + CNT_SEP_KDS, K, ctx->sep_offsets.as<int3>(), (int)n_off, min_sure, w1, w2, (float)p.score_ray, cnt, cap_guard);
  return VOFOD_OK;
}

extern "C" int vofod_sepclusters(vofod_ctx* ctx, int its_diff, const vofod_params* p, int* sure_background_sufficient)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  if (!p)
    return vf_fail(ctx, VOFOD_E_INVALID, "params is NULL");
  RET(vf_begin_call(ctx));
  const int rc = vf_sepclusters_dev(ctx, its_diff, *p, 0);
  if (rc < 0)
    return rc;
  unsigned long long h[CNT_N_SLOTS];
  CK(cudaMemcpyAsync(h, ctx->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (h[CNT_WATCHDOG])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", h[CNT_WATCHDOG]);
  if (h[CNT_SEP_NUNIQ])
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "sepclusters: voxel-grid index overflow");
  ctx->sure_background_sufficient = h[CNT_STATE_SURE] != 0;
  if (sure_background_sufficient)
    *sure_background_sufficient = ctx->sure_background_sufficient;
  return rc;
}
