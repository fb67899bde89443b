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

// ---------------------------------------------------------------------------------------------------------------
// Fast path for the reference's default geometry: ceil(max_bg_distance / voxel_size) == 2, i.e. VoxelGridCounted runs with
// leaf size 1 on the integer voxel coordinates that voxelsAsVoxelPC emits.  Every input point then is its own leaf, and
// the filter degenerates to a permutation:
//     output r (r-th voxel in KEY order = storage order z,y,x) = centre (x+0.5, y+0.5, z+0.5) of that voxel,
//     its count = [intensity > sure threshold] of the r-th voxel in INPUT order (= x-outer/z-inner emission order; this is
//     the input-slice quirk of voxel_grid_counted.cpp:185-187 with runs of length one).
// Both orders come out of ONE counting pass and ONE emission pass over the (dirty columns of the) grid — no sort:
// a warp owns a 32-wide x segment of a row y and walks z; ballots give the rank inside the segment.
// ---------------------------------------------------------------------------------------------------------------
// A warp owns (row y, 32-wide x segment, chunk of SEP_ZC z-levels): walking a whole column is a chain of ~20 dependent
// memory round trips, which — not bandwidth — would set the pace.  Column counts are kept per (column, z-chunk) with the
// chunk as the fastest index, so that their scan still runs in the reference's x-outer / z-inner emission order.
#define SEP_ZC DIRTY_ZC
__global__ void __launch_bounds__(256) k_sep_fast_count(const float* __restrict__ score, const Geom g, const float thr, const uint8_t* __restrict__ dirty,
                                                        const int nseg, const int nzc, uint32_t* __restrict__ colcnt, uint32_t* __restrict__ segcnt)
{
  const unsigned lane = threadIdx.x & 31;
  const int sx = g.st_size[0], sy = g.st_size[1], sz = g.st_size[2];
  const size_t sxy = (size_t)sx * sy;
  const int n_items = sy * nseg * nzc;
  for (int item = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); item < n_items; item += (int)(((size_t)gridDim.x * blockDim.x) >> 5))
  {
    const int zc = item % nzc, seg = (item / nzc) % nseg, y = item / (nzc * nseg);
    const int x = seg * 32 + (int)lane;
    const bool in = x < sx;
    const size_t c = (size_t)y * sx + x;
    const bool live = in && (!dirty || dirty[(size_t)zc * sxy + c]) && column_owned(g, x, y);
    if (!__any_sync(VOFOD_FULL, live))
      continue;  // colcnt / segcnt were zero-filled
    const int z_lo = zc * SEP_ZC, z_hi = min(z_lo + SEP_ZC, sz);
    uint32_t cnt = 0;
    for (int z0 = z_lo; z0 < z_hi; z0 += 8)
    {
      // 8 independent loads in flight, then the (cheap) ballots
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = (live && z0 + k < z_hi) ? score[c + (size_t)(z0 + k) * sxy] : __int_as_float(0xff800000);
      unsigned mine = 0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        mine |= (v[k] > thr ? 1u : 0u) << k;
      cnt += __popc(mine);
      if (!__any_sync(VOFOD_FULL, mine != 0))
        continue;
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        const unsigned bal = __ballot_sync(VOFOD_FULL, (mine >> k) & 1u);
        if (bal && lane == 0)
          segcnt[((size_t)(z0 + k) * sy + y) * nseg + seg] = __popc(bal);
      }
    }
    if (cnt)
      colcnt[((size_t)x * sy + y) * nzc + zc] = cnt;
  }
}
__global__ void __launch_bounds__(256) k_sep_fast_emit(const float* __restrict__ score, const Geom g, const float thr, const float thr_sure, const int nseg,
                                                       const int nzc, const uint8_t* __restrict__ dirty, const uint32_t* __restrict__ colcnt,
                                                       const uint32_t* __restrict__ coloff,
                                                       const uint32_t* __restrict__ segoff, vofod_vox* __restrict__ ds, uint32_t* __restrict__ flag_in_order,
                                                       uint32_t* __restrict__ idgrid, const size_t cap)
{
  const unsigned lane = threadIdx.x & 31;
  const int sx = g.st_size[0], sy = g.st_size[1], sz = g.st_size[2];
  const size_t sxy = (size_t)sx * sy;
  const int n_items = sy * nseg * nzc;
  for (int item = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); item < n_items; item += (int)(((size_t)gridDim.x * blockDim.x) >> 5))
  {
    const int zc = item % nzc, seg = (item / nzc) % nseg, y = item / (nzc * nseg);
    const int x = seg * 32 + (int)lane;
    const bool in = x < sx;
    const size_t c = (size_t)y * sx + x;
    const size_t ci = ((size_t)(in ? x : 0) * sy + y) * nzc + zc;
    const bool live = in && (!dirty || dirty[(size_t)zc * sxy + c]) && colcnt[ci] != 0;  // the (coalesced) marks spare most of the strided count reads
    if (!__any_sync(VOFOD_FULL, live))
      continue;
    size_t o = live ? coloff[ci] : 0;
    const int z_lo = zc * SEP_ZC, z_hi = min(z_lo + SEP_ZC, sz);
    for (int z0 = z_lo; z0 < z_hi; z0 += 8)
    {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = (live && z0 + k < z_hi) ? score[c + (size_t)(z0 + k) * sxy] : __int_as_float(0xff800000);
      unsigned mine = 0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        mine |= (v[k] > thr ? 1u : 0u) << k;
      if (!__any_sync(VOFOD_FULL, mine != 0))
        continue;
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        const bool match = (mine >> k) & 1u;
        const unsigned bal = __ballot_sync(VOFOD_FULL, match);
        if (!bal)
          continue;
        const int z = z0 + k;
        const size_t base = segoff[((size_t)z * sy + y) * nseg + seg];
        if (match)
        {
          const size_t r = base + __popc(bal & prims::lanemask_lt());
          if (r < cap)
          {
            vofod_vox out;
            out.x = (float)(x + g.st_lo[0]) + 0.5f;  // (float(ijk) + 0.5f) * 1 + float(min_b) with ijk = idx - min_b: exact
            out.y = (float)(y + g.st_lo[1]) + 0.5f;
            out.z = (float)(z + g.st_lo[2]) + 0.5f;
            out.count = 0;
            ds[r] = out;
            idgrid[c + (size_t)z * sxy] = (uint32_t)r;  // cell -> point number, for the 26-connectivity clustering
          }
          if (o < cap)
            flag_in_order[o] = v[k] > thr_sure ? 1u : 0u;
          o++;
        }
      }
    }
  }
}
// + the initial state of the union-find that vf_cluster_grid26_dev works on
__global__ void __launch_bounds__(256) k_sep_fast_pair(vofod_vox* __restrict__ ds, const uint32_t* __restrict__ flag_in_order, const unsigned long long* __restrict__ d_k,
                                                       const size_t cap, int* __restrict__ parent, int* __restrict__ sizes, int* __restrict__ minidx)
{
  const size_t k = prims::dev_count(d_k, cap);
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < k; r += (size_t)gridDim.x * blockDim.x)
  {
    ds[r].count = flag_in_order[r];
    parent[r] = (int)r;
    sizes[r] = 0;
    minidx[r] = 0x7fffffff;
  }
}

// lists for the fast path; *host_total as in vf_compact_over_dev
static int sep_fast_lists(vofod_ctx* ctx, const float thr, const float thr_sure, const vofod_params& p, size_t* host_total, size_t cap)
{
  using namespace prims;
  const Geom& g = ctx->g;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  const int nseg = (g.st_size[0] + 31) / 32;
  const int nzc = (g.st_size[2] + SEP_ZC - 1) / SEP_ZC;
  const size_t ncc = (size_t)g.st_size[0] * g.st_size[1] * nzc;                // (column, z-chunk) counts
  const size_t nsegs = (size_t)g.st_size[1] * g.st_size[2] * nseg;             // (z, y, x-segment) counts
  if (nsegs >= (size_t(1) << 31) || ncc >= (size_t(1) << 31))
    return 1;  // not representable here: use the general path
  ENSURE(ctx->sep_colcnt, padded(ncc) * 4);
  ENSURE(ctx->sep_coloff, padded(ncc) * 4);
  ENSURE(ctx->sep_segcnt, padded(nsegs) * 4);
  ENSURE(ctx->sep_segoff, padded(nsegs) * 4);
  const uint8_t* dirty = vf_dirty_cols(ctx, thr, &p);
  CK(cudaMemsetAsync(ctx->sep_colcnt.p, 0, ncc * 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->sep_segcnt.p, 0, nsegs * 4, ctx->stream));
  const int nb = vf_blocks(ctx, (size_t)g.st_size[1] * nseg * nzc * 32, 256, 8);
  LAUNCH(k_sep_fast_count, nb, 256, 0, ctx->score.as<float>(), g, thr, dirty, nseg, nzc, ctx->sep_colcnt.as<uint32_t>(), ctx->sep_segcnt.as<uint32_t>());
  // the two scans are independent: the second one runs on the side stream (a parallel branch under graph replay)
  const bool fork = ctx->stream2 != nullptr && host_total == nullptr && ctx->overlap_enabled;
  cudaStream_t st = ctx->stream;
  if (fork)
  {
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    ctx->stream = ctx->stream2;
    const int rc2 = scan_excl_u32(ctx, ctx->sep_segcnt.as<uint32_t>(), ctx->sep_segoff.as<uint32_t>(), nullptr, nsegs, nullptr, &ctx->tile_state2);
    ctx->stream = st;
    if (rc2 < 0)
      return rc2;
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  }
  RET(scan_excl_u32(ctx, ctx->sep_colcnt.as<uint32_t>(), ctx->sep_coloff.as<uint32_t>(), nullptr, ncc, cnt + CNT_SEP_K));
  if (fork)
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  else
    RET(scan_excl_u32(ctx, ctx->sep_segcnt.as<uint32_t>(), ctx->sep_segoff.as<uint32_t>(), nullptr, nsegs, nullptr));
  if (host_total)
  {
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, cnt + CNT_SEP_K, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *host_total = (size_t)total;
    cap = (size_t)total;
    if (total == 0)
      return 0;
  }
  ENSURE(ctx->sep_ds, padded(cap) * sizeof(vofod_vox));
  ENSURE(ctx->vg_flags, padded(cap) * 4);
  ENSURE(ctx->sep_idgrid, (size_t)geom_cells(g) * 4);  // only the listed cells are ever written or read: no clearing
  ENSURE(ctx->cl_bg.parent, cap * 4);
  ENSURE(ctx->cl_bg.sizes, cap * 4);
  ENSURE(ctx->cl_bg.minidx, cap * 4);
  LAUNCH(k_sep_fast_emit, nb, 256, 0, ctx->score.as<float>(), g, thr, thr_sure, nseg, nzc, dirty, ctx->sep_colcnt.as<uint32_t>(), ctx->sep_coloff.as<uint32_t>(),
         ctx->sep_segoff.as<uint32_t>(), ctx->sep_ds.as<vofod_vox>(), ctx->vg_flags.as<uint32_t>(), ctx->sep_idgrid.as<uint32_t>(), cap);
  LAUNCH(k_sep_fast_pair, vf_blocks(ctx, cap, 256, 8), 256, 0, ctx->sep_ds.as<vofod_vox>(), ctx->vg_flags.as<uint32_t>(), cnt + CNT_SEP_K, cap,
         ctx->cl_bg.parent.as<int>(), ctx->cl_bg.sizes.as<int>(), ctx->cl_bg.minidx.as<int>());
  return 0;
}

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
    if (ci < 0)
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
  const float lsz = (float)(mv - 1 > 0 ? mv - 1 : 0);  // :1163
  if (!(lsz > 0.0f))
    return vf_fail(ctx, VOFOD_E_INVALID, "sepclusters: max_bg_distance/voxel_size <= 1 gives a zero leaf size (the reference divides by it)");
  size_t K = k_cap;
  const unsigned long long* d_kds = cnt + CNT_SEP_KDS;
  bool fast = lsz == 1.0f && !ctx->sep_force_general;
  if (fast)
  {
    const int frc = sep_fast_lists(ctx, thr_new, thr_sure, p, k_cap == 0 ? &K : nullptr, k_cap);
    if (frc < 0)
      return frc;
    fast = frc == 0;
    if (fast)
    {
      if (k_cap == 0 && K == 0)
        return VOFOD_W_EMPTY;  // :1155-1159
      d_kds = cnt + CNT_SEP_K;  // every voxel is its own leaf
      ZERO_CNT(CNT_SEP_NUNIQ, 1);
    }
  }
  if (!fast)
  {
    K = k_cap;
    if (k_cap == 0)
    {
      RET(vf_compact_over_dev(ctx, thr_new, 1, 0, ctx->sep_raw, cnt + CNT_SEP_K, &K, 0, &p));
      if (K == 0)
        return VOFOD_W_EMPTY;  // :1155-1159
    } else
      RET(vf_compact_over_dev(ctx, thr_new, 1, 0, ctx->sep_raw, cnt + CNT_SEP_K, nullptr, k_cap, &p));
    RET(vf_voxel_grid_counted_dev(ctx, ctx->sep_raw.as<vofod_xyzi>(), cnt + CNT_SEP_K, K, lsz, thr_sure, ctx->sep_ds));
  }
  const unsigned long long cap_guard = k_cap ? (unsigned long long)k_cap : ~0ull;
  ENSURE(ctx->sep_labels, K * 4);
  ENSURE(ctx->sep_nsure, K * 4);
  if (fast)  // leaf size 1 <=> tolerance 2 on distinct voxel centres: 26-connectivity, read off the grid
    RET(vf_cluster_grid26_dev(ctx, ctx->cl_bg, ctx->sep_ds.as<vofod_vox>(), ctx->sep_idgrid.as<uint32_t>(), thr_new, d_kds, K, ctx->sep_labels.as<int>(),
                              cnt + CNT_SEP_NCL));
  else
    RET(vf_cluster_dev(ctx, ctx->cl_bg, reinterpret_cast<const float*>(ctx->sep_ds.p), 4, d_kds, K, (float)mv, ctx->sep_labels.as<int>(), cnt + CNT_SEP_NCL,
                       k_cap ? ctx->sep_table_hint : 0));
  CK(cudaMemsetAsync(ctx->sep_nsure.p, 0, K * 4, ctx->stream));
  ZERO_CNT(CNT_SEP_ANY_SURE, 1);
  const int nb = vf_blocks(ctx, K, 256, 8);
  LAUNCH(k_sep_nsure, nb, 256, 0, ctx->sep_ds.as<vofod_vox>(), ctx->sep_labels.as<int>(), d_kds, K, ctx->sep_nsure.as<int>());
  LAUNCH(k_sep_any, nb, 256, 0, ctx->sep_labels.as<int>(), ctx->sep_nsure.as<int>(), d_kds, K, min_sure, cnt);
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
         ctx->sep_nsure.as<int>(), d_kds, K, ctx->sep_offsets.as<int3>(), (int)n_off, min_sure, w1, w2, (float)p.score_ray, cnt, cap_guard);
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
