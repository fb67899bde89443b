// updateSeparatedBGClusters (vofod_nodelet.cpp:1126-1278) on the GPU: the background thread's pass that finds
// background voxels which are NOT attached to a large "sure" background body and decays them towards the ray score.
//
// General path (any max_bg_distance / voxel_size):
//   compaction of voxels > new_obstacles in the reference's x-outer/z-inner order (K11, ctx.cu)
//   -> VoxelGridCounted (voxelgrid.cu, incl. its input-slice quirk) -> Euclidean clustering in index units (cluster.cu)
//   -> per-cluster sum of sure counts -> decay scatter (K12).
// Fast path (the default geometry, leaf size 1 — see below): occupancy masks + popcount scan instead of compaction and sort,
//   26-connectivity run against run instead of the spatial hash, sure counts per union-find root.
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
// Both orders come out of ONE counting pass and ONE emission pass — no sort.  The unit of work is
//     item = (row y, 32-wide x segment, chunk of SEP_ZC z-levels),   one warp per item, lane = x inside the segment.
// Only a few percent of the items can hold a voxel above the threshold (the "raised" marks of vofod_ctx::col_dirty say
// which), so a first kernel lists those and the two passes walk the list: with one listed item per warp and all its
// SEP_ZC loads in flight at once the passes cost a few memory round trips instead of a sweep over the grid.
//   count: per (z, y, segment) the BIT MASK of matching x (ranks in key order come from a popcount scan of the masks; the
//          masks also are the occupancy structure the clustering works on) and per (x, y, chunk) the number of matches
//          (chunk fastest, so that their scan runs in the reference's x-outer / z-inner emission order)
//   emit : centres in key order, sure flags in input order, and the union-find forest of the clustering seeded with
//          "every voxel hangs under the first voxel of its run" (run = consecutive set bits of one mask)
// ---------------------------------------------------------------------------------------------------------------
#define SEP_ZC DIRTY_ZC
static_assert(SEP_ZC == 32, "one lane per z-level of a chunk");
// loads the SEP_ZC levels of the item's lane column and returns the bit mask of levels above thr
__device__ __forceinline__ unsigned sep_load_levels(const float* __restrict__ score, const size_t c, const size_t sxy, const int z_lo, const int z_hi, const bool on,
                                                    const float thr, float (&v)[SEP_ZC])
{
#pragma unroll
  for (int k = 0; k < SEP_ZC; k++)
    v[k] = (on && z_lo + k < z_hi) ? score[c + (size_t)(z_lo + k) * sxy] : __int_as_float(0xff800000);
  unsigned mine = 0;
#pragma unroll
  for (int k = 0; k < SEP_ZC; k++)
    mine |= (v[k] > thr ? 1u : 0u) << k;
  return mine;
}
// counting pass over ALL items; the ones that can hold a match are appended to `live` for the emission pass
__global__ void __launch_bounds__(256) k_sep_fast_count(const float* __restrict__ score, const Geom g, const float thr, const uint8_t* __restrict__ dirty,
                                                        uint2* __restrict__ live, unsigned long long* __restrict__ n_live, const int nseg, const int nzc,
                                                        uint32_t* __restrict__ colcnt, uint32_t* __restrict__ segbits)
{
  pdl_enter();
  const unsigned lane = threadIdx.x & 31;
  const int sx = g.st_size[0], sy = g.st_size[1], sz = g.st_size[2];
  const size_t sxy = (size_t)sx * sy;
  const int n_items = sy * nseg * nzc;
  for (int item = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); item < n_items; item += (int)(((size_t)gridDim.x * blockDim.x) >> 5))
  {
    const int zc = item % nzc, seg = (item / nzc) % nseg, y = item / (nzc * nseg);
    const int x = seg * 32 + (int)lane;
    const bool on = x < sx && (!dirty || dirty[((size_t)zc * sy + y) * sx + x]);
    const unsigned mask = __ballot_sync(VOFOD_FULL, on);
    if (!mask)
      continue;
    const int z_lo = zc * SEP_ZC, z_hi = min(z_lo + SEP_ZC, sz);
    float v[SEP_ZC];
    const unsigned mine = sep_load_levels(score, (size_t)y * sx + x, sxy, z_lo, z_hi, on, thr, v);
    if (!__any_sync(VOFOD_FULL, mine != 0))
      continue;  // colcnt / segbits were zero-filled
    if (lane == 0)
      live[atomicAdd(n_live, 1ull)] = make_uint2((unsigned)item, mask);
    unsigned my_level_mask = 0;  // lane k keeps the mask of level z_lo + k
#pragma unroll
    for (int k = 0; k < SEP_ZC; k++)
    {
      const unsigned bal = __ballot_sync(VOFOD_FULL, (mine >> k) & 1u);
      if (lane == (unsigned)k)
        my_level_mask = bal;
    }
    if (my_level_mask)
      segbits[((size_t)(z_lo + (int)lane) * sy + y) * nseg + seg] = my_level_mask;
    if (mine)
      colcnt[((size_t)x * sy + y) * nzc + zc] = __popc(mine);
  }
}
__global__ void __launch_bounds__(256) k_sep_fast_emit(const float* __restrict__ score, const Geom g, const float thr, const float thr_sure, const uint2* __restrict__ live,
                                                       const unsigned long long* __restrict__ n_live, const int nseg, const int nzc,
                                                       const uint32_t* __restrict__ coloff, const uint32_t* __restrict__ segoff, vofod_vox* __restrict__ ds,
                                                       uint32_t* __restrict__ flag_in_order, int* __restrict__ parent, int* __restrict__ nsure, const size_t cap)
{
  pdl_enter();
  const unsigned lane = threadIdx.x & 31;
  const int sx = g.st_size[0], sy = g.st_size[1], sz = g.st_size[2];
  const size_t sxy = (size_t)sx * sy;
  const size_t n = (size_t)*after_wait(n_live);
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += ((size_t)gridDim.x * blockDim.x) >> 5)
  {
    const uint2 it = live[w];
    const int item = (int)it.x;
    const int zc = item % nzc, seg = (item / nzc) % nseg, y = item / (nzc * nseg);
    const int x = seg * 32 + (int)lane;
    const bool on = (it.y >> lane) & 1u;
    const int z_lo = zc * SEP_ZC, z_hi = min(z_lo + SEP_ZC, sz);
    // issued before the level loads so that everything is in flight together: the input-order offset of this lane's column
    // chunk, and (lane k) the key-order offset of level z_lo + k
    size_t o = on ? coloff[((size_t)x * sy + y) * nzc + zc] : 0;
    const uint32_t my_level_off = (z_lo + (int)lane < z_hi) ? segoff[((size_t)(z_lo + (int)lane) * sy + y) * nseg + seg] : 0u;
    float v[SEP_ZC];
    const unsigned mine = sep_load_levels(score, (size_t)y * sx + x, sxy, z_lo, z_hi, on, thr, v);
    if (!__any_sync(VOFOD_FULL, mine != 0))
      continue;
#pragma unroll
    for (int k = 0; k < SEP_ZC; k++)
    {
      const bool match = (mine >> k) & 1u;
      const unsigned bal = __ballot_sync(VOFOD_FULL, match);
      if (!bal)
        continue;
      const size_t base = __shfl_sync(VOFOD_FULL, my_level_off, k);
      if (match)
      {
        const size_t r = base + __popc(bal & prims::lanemask_lt());
        if (r < cap)
        {
          vofod_vox out;
          out.x = (float)(x + g.st_lo[0]) + 0.5f;  // (float(ijk) + 0.5f) * 1 + float(min_b) with ijk = idx - min_b: exact
          out.y = (float)(y + g.st_lo[1]) + 0.5f;
          out.z = (float)(z_lo + k + g.st_lo[2]) + 0.5f;
          out.count = 0;  // the counts live in flag_in_order (see k_sep_nsure)
          ds[r] = out;
          // first lane of this lane's run of consecutive set bits
          const unsigned zeros_below = ~bal & prims::lanemask_lt();
          const int head = zeros_below ? 32 - __clz(zeros_below) : 0;
          parent[r] = (int)(base + __popc(bal & ((1u << head) - 1u)));
          nsure[r] = 0;
        }
        if (o < cap)
          flag_in_order[o] = v[k] > thr_sure ? 1u : 0u;
        o++;
      }
    }
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
  const size_t nsegs = (size_t)g.st_size[1] * g.st_size[2] * nseg;             // (z, y, x-segment) masks
  const size_t n_items = (size_t)g.st_size[1] * nseg * nzc;
  if (nsegs >= (size_t(1) << 31) || ncc >= (size_t(1) << 31))
    return 1;  // not representable here: use the general path
  ENSURE(ctx->sep_colcnt, padded(ncc) * 4);
  ENSURE(ctx->sep_coloff, padded(ncc) * 4);
  ENSURE(ctx->sep_segcnt, padded(nsegs) * 4);
  ENSURE(ctx->sep_segoff, padded(nsegs) * 4);
  ENSURE(ctx->sep_live, (n_items + 1) * sizeof(uint2));
  const uint8_t* dirty = vf_dirty_cols(ctx, thr, &p);
  if (ctx->sep_prefilled != ncc + nsegs)
  {
    const FillJob fj[2] = {{ctx->sep_colcnt.as<uint32_t>(), ncc, 0u}, {ctx->sep_segcnt.as<uint32_t>(), nsegs, 0u}};
    RET(vf_fill(ctx, fj, 2));
  }
  ctx->sep_prefilled = 0;
  ZERO_CNT(CNT_SEP_LIVE, 1);
  const int nb = vf_blocks(ctx, n_items * 32, 256, 8);
  LAUNCH(k_sep_fast_count, nb, 256, 0, ctx->score.as<float>(), g, thr, dirty, ctx->sep_live.as<uint2>(), cnt + CNT_SEP_LIVE, nseg, nzc,
         ctx->sep_colcnt.as<uint32_t>(), ctx->sep_segcnt.as<uint32_t>());
  RET(scan_excl_u32_pair(ctx, ctx->sep_colcnt.as<uint32_t>(), ctx->sep_coloff.as<uint32_t>(), nullptr, ncc, cnt + CNT_SEP_K, false, ctx->sep_segcnt.as<uint32_t>(),
                         ctx->sep_segoff.as<uint32_t>(), nsegs, nullptr, true, nullptr, nullptr, true));  // offsets are read at non-empty entries only
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
  ENSURE(ctx->cl_bg.parent, cap * 4);
  ENSURE(ctx->sep_nsure, cap * 4);
  LAUNCH(k_sep_fast_emit, nb, 256, 0, ctx->score.as<float>(), g, thr, thr_sure, ctx->sep_live.as<uint2>(), cnt + CNT_SEP_LIVE, nseg, nzc,
         ctx->sep_coloff.as<uint32_t>(), ctx->sep_segoff.as<uint32_t>(), ctx->sep_ds.as<vofod_vox>(), ctx->vg_flags.as<uint32_t>(), ctx->cl_bg.parent.as<int>(),
         ctx->sep_nsure.as<int>(), cap);
  return 0;
}

// :1174-1183 — n_sure[cluster] = std::accumulate(range, int 0)
// `counts` != NULL: the per-voxel counts are there (fast path), not in ds[].count
__global__ void __launch_bounds__(256) k_sep_nsure(const vofod_vox* __restrict__ ds, const uint32_t* __restrict__ counts, const int* __restrict__ labels,
                                                   const unsigned long long* __restrict__ d_k, const size_t cap, int* __restrict__ nsure)
{
  pdl_enter();
  const size_t k = prims::dev_count(d_k, cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < k; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool valid = i < k;
    const int l = valid ? labels[i] : -1 - (int)lane;
    const unsigned c = valid ? (counts ? counts[i] : ds[i].count) : 0u;
    // one atomic per (warp, cluster): the big background body would otherwise serialise every point on one word
    const unsigned grp = __match_any_sync(VOFOD_FULL, l);
    const unsigned sum = __reduce_add_sync(grp, c);
    if (valid && lane == (unsigned)(__ffs(grp) - 1))
      atomicAdd(reinterpret_cast<unsigned*>(nsure) + l, sum);
  }
}
// fast path: root of every voxel + n_sure per ROOT in one pass over the forest.  Which member names a cluster is irrelevant to
// this stage (only sums per cluster and "is any cluster sure" matter), so the canonical min-index labels are not computed.
__global__ void __launch_bounds__(256) k_sep_roots_nsure(const int* __restrict__ parent, const uint32_t* __restrict__ counts, const unsigned long long* __restrict__ d_k,
                                                         const size_t cap, int* __restrict__ root, int* __restrict__ nsure)
{
  pdl_enter();
  const size_t k = prims::dev_count(d_k, cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < k; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool valid = i < k;
    int r = valid ? (int)i : -1 - (int)lane;
    if (valid)
      while (true)
      {
        const int pr = parent[r];
        if (pr == r)
          break;
        r = pr;
      }
    if (valid)
      root[i] = r;
    const unsigned c = valid ? counts[i] : 0u;
    // one atomic per (warp, cluster): the big background body would otherwise serialise every point on one word
    const unsigned grp = __match_any_sync(VOFOD_FULL, r);
    const unsigned sum = __reduce_add_sync(grp, c);
    if (valid && sum && lane == (unsigned)(__ffs(grp) - 1))
      atomicAdd(reinterpret_cast<unsigned*>(nsure) + r, sum);
  }
}
// :1186-1206 — is any cluster sure?  + the list of the voxels of UNSURE clusters (normally a handful) for the decay
__global__ void __launch_bounds__(256) k_sep_any(const int* __restrict__ labels, const int* __restrict__ nsure, const unsigned long long* __restrict__ d_k, const size_t cap,
                                                 const unsigned min_sure, unsigned long long* __restrict__ counters, uint32_t* __restrict__ unsure,
                                                 const int count_clusters)
{
  pdl_enter();
  const size_t k = prims::dev_count(d_k, cap);
  bool any = false;
  unsigned n_roots = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (size_t)gridDim.x * blockDim.x)
  {
    const int l = labels[i];
    const int ns = nsure[l];
    if (l == (int)i && (unsigned long long)(long long)ns >= (unsigned long long)min_sure)  // size_t(int) >= unsigned
      any = true;
    n_roots += l == (int)i;
    if ((unsigned)ns < min_sure)  // :1246
      unsure[atomicAdd(counters + CNT_SEP_NUNSURE, 1ull)] = (uint32_t)i;
  }
  if (__any_sync(VOFOD_FULL, any) && (threadIdx.x & 31) == 0)
    counters[CNT_SEP_ANY_SURE] = 1ull;
  n_roots = prims::warp_sum(n_roots);
  if (count_clusters && n_roots && (threadIdx.x & 31) == 0)
    atomicAdd(counters + CNT_SEP_NCL, (unsigned long long)n_roots);
}
__device__ __forceinline__ void sep_state(unsigned long long* __restrict__ counters, const unsigned long long k_cap)
{
  const unsigned long long k = counters[CNT_SEP_K];
  if (k == 0ull || k > k_cap)
    return;  // empty cloud: the reference returns before touching the flag (:1155-1159); overflow: the host redoes the pass
  counters[CNT_STATE_SURE] = counters[CNT_SEP_ANY_SURE] ? 1ull : 0ull;  // :1196 / :1205
}
__global__ void k_sep_state(unsigned long long* __restrict__ counters, const unsigned long long k_cap)
{
  pdl_enter();
  sep_state(counters, k_cap);
}

// K12 — :1244-1272: one thread per (listed voxel, offset)
__global__ void __launch_bounds__(256) k_sep_decay(float* score, const Geom g, const vofod_vox* __restrict__ ds, const uint32_t* __restrict__ unsure,
                                                   const size_t cap, const int3* __restrict__ offsets, const int n_off, const float w1, const float w2,
                                                   const float update_val, unsigned long long* counters, const unsigned long long k_cap)
{
  pdl_enter();
  if (blockIdx.x == 0 && threadIdx.x == 0)
    sep_state(counters, k_cap);
  if (counters[CNT_SEP_ANY_SURE] == 0ull || counters[CNT_SEP_K] > k_cap)
    return;  // :1192-1199 (and: list overflow => nothing is touched, the host redoes the pass with a larger list)
  const size_t n_list = prims::dev_count(counters + CNT_SEP_NUNSURE, cap);
  const size_t total = n_list * (size_t)n_off;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t e = t / n_off;
    const int o = (int)(t - e * n_off);
    const vofod_vox v = ds[unsure[e]];
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
      const float mval = __uint_as_float(old);
      const float nv = w1 * mval + w2 * update_val;  // :1259
      const unsigned prev = atomicCAS(addr, old, __float_as_uint(nv));
      if (prev == old)
        break;
      old = prev;
    }
  }
}

// the clears of the next fast-path pass, ahead of time (on the scan's side branch); a no-op when that pass would not take
// the fast path
int vf_sepclusters_prefill(vofod_ctx* ctx, const vofod_params& p)
{
  using namespace prims;
  const float vs = ctx->cfg_voxel_size > 0.f ? ctx->cfg_voxel_size : ctx->g.vs;
  const int mv = (int)ceilf((float)(p.sep_max_bg_distance / (double)vs));
  if (p.sep_pause || mv - 1 != 1 || ctx->sep_force_general || ctx->slab_on)
    return 0;
  const Geom& g = ctx->g;
  const int nseg = (g.st_size[0] + 31) / 32;
  const int nzc = (g.st_size[2] + SEP_ZC - 1) / SEP_ZC;
  const size_t ncc = (size_t)g.st_size[0] * g.st_size[1] * nzc;
  const size_t nsegs = (size_t)g.st_size[1] * g.st_size[2] * nseg;
  if (nsegs >= (size_t(1) << 31) || ncc >= (size_t(1) << 31))
    return 0;
  ENSURE(ctx->sep_colcnt, padded(ncc) * 4);
  ENSURE(ctx->sep_segcnt, padded(nsegs) * 4);
  const FillJob fj[2] = {{ctx->sep_colcnt.as<uint32_t>(), ncc, 0u}, {ctx->sep_segcnt.as<uint32_t>(), nsegs, 0u}};
  RET(vf_fill(ctx, fj, 2));
  ctx->sep_prefilled = ncc + nsegs;
  return 0;
}

// k_cap == 0: exact mode (one host read-back sizes the background-voxel list).  k_cap > 0: the list is capped at k_cap rows and
// nothing returns to the host; when the true count exceeds k_cap the pass leaves the map untouched and the caller, who sees
// CNT_SEP_K > k_cap in its read-back, repeats it in exact mode.
// ---- slab mode (slab.cu): the background voxels of all slabs, gathered ------------------------------------------------------------
// Every slab compacts the cells it OWNS (x-outer / z-inner, global index coordinates) and packs them as
//   word 0 = its true count, word 1 + i = global linear cell index | (value > sure threshold) << 31.
// After the allgather (fixed capacity per slab) every slab unpacks the segments in rank order — slabs cut the x axis in ascending
// order, so the concatenation IS the reference's emission order over the whole map — and runs the rest of the pass replicated on the
// same list; the decay touches the cells a slab holds.
__global__ void __launch_bounds__(256) k_sep_slab_pack(const vofod_xyzi* __restrict__ raw, const unsigned long long* __restrict__ d_k, const size_t cap, const Geom g,
                                                       const float thr_sure, uint32_t* __restrict__ send)
{
  pdl_enter();
  const unsigned long long k_true = *after_wait(d_k);
  const size_t k = k_true < cap ? (size_t)k_true : cap;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    send[0] = k_true > 0xffffffffull ? 0xffffffffu : (uint32_t)k_true;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (size_t)gridDim.x * blockDim.x)
  {
    const vofod_xyzi v = raw[i];
    const unsigned lin = (unsigned)((int)v.x + ((int)v.y + (int)v.z * g.size[1]) * g.size[0]);
    send[1 + i] = lin | (v.intensity > thr_sure ? 0x80000000u : 0u);
  }
}
__global__ void __launch_bounds__(256) k_sep_slab_unpack(const uint32_t* __restrict__ recv, const int nranks, const size_t cap, const size_t k_consume, const Geom g,
                                                         vofod_xyzi* __restrict__ raw, unsigned long long* __restrict__ counters)
{
  pdl_enter();
  __shared__ unsigned long long s_off[65];
  __shared__ int s_over;
  if (threadIdx.x == 0)
  {
    unsigned long long o = 0;
    int over = 0;
    for (int r = 0; r < nranks; r++)
    {
      s_off[r] = o;
      const unsigned long long c = after_wait(recv)[(size_t)r * (cap + 1)];
      over |= c > cap;
      o += c < cap ? c : cap;
    }
    s_off[nranks] = o;
    over |= o > k_consume;  // the rest of the pass is sized for k_consume rows (about twice the list of the previous scan), not for nranks * cap
    s_over = over;
    if (blockIdx.x == 0)
      // an overflowing slab: report a count above every capacity, the pass then leaves the map untouched (k_sep_decay) and the host redoes it
      counters[CNT_SEP_K] = over ? ~0ull >> 1 : o;
  }
  __syncthreads();
  if (s_over)
    return;
  const size_t total = (size_t)s_off[nranks];
  const size_t sxy = (size_t)g.size[0] * g.size[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    int r = 0;
    while (i >= s_off[r + 1])
      r++;
    const uint32_t w = recv[(size_t)r * (cap + 1) + 1 + (i - s_off[r])];
    const unsigned lin = w & 0x7fffffffu;
    vofod_xyzi v;
    v.x = (float)(lin % (unsigned)g.size[0]);
    v.y = (float)((lin / (unsigned)g.size[0]) % (unsigned)g.size[1]);
    v.z = (float)(lin / sxy);
    v.intensity = (w >> 31) ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);  // only "> sure threshold" is ever asked of it (voxel_grid_counted.cpp:185-187)
    raw[i] = v;
  }
}
// phase A: this slab's list into ctx->slab_bg_send ((cap + 1) words)
int vf_sep_slab_pack(vofod_ctx* ctx, const vofod_params& p, size_t cap)
{
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  if (geom_cells(ctx->g) <= 0 || (long long)ctx->g.size[0] * ctx->g.size[1] * ctx->g.size[2] >= (1ll << 31))
    return vf_fail(ctx, VOFOD_E_INVALID, "slab sepclusters: the global grid must have fewer than 2^31 cells");
  RET(vf_compact_over_dev(ctx, (float)p.thr_new_obstacles, 1, 0, ctx->sep_raw, cnt + CNT_SEP_K, nullptr, cap, &p));
  LAUNCH(k_sep_slab_pack, vf_blocks(ctx, cap, 256, 8), 256, 0, ctx->sep_raw.as<vofod_xyzi>(), cnt + CNT_SEP_K, cap, ctx->g, (float)p.thr_sure_obstacles,
         ctx->slab_bg_send.as<uint32_t>());
  return 0;
}
int vf_sepclusters_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t k_cap, int slab_ranks);
// phase B: the gathered lists (ctx->slab_bg_recv, nranks x (cap + 1) words) -> the rest of the pass
int vf_sep_slab_finish(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t cap, int nranks, size_t k_consume)
{
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  const size_t K = k_consume < cap * (size_t)nranks ? k_consume : cap * (size_t)nranks;
  ENSURE(ctx->sep_raw, prims::padded(K) * sizeof(vofod_xyzi));
  LAUNCH(k_sep_slab_unpack, vf_blocks(ctx, K, 256, 8), 256, 0, ctx->slab_bg_recv.as<uint32_t>(), nranks, cap, K, ctx->g, ctx->sep_raw.as<vofod_xyzi>(), cnt);
  return vf_sepclusters_dev(ctx, its_diff, p, K, nranks);
}

// slab_ranks > 0: the raw list (ctx->sep_raw, CNT_SEP_K) is there already (vf_sep_slab_finish)
int vf_sepclusters_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t k_cap, int slab_ranks)
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
  bool fast = lsz == 1.0f && !ctx->sep_force_general && !ctx->slab_on && slab_ranks == 0;
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
    if (slab_ranks > 0)
      ;
    else if (k_cap == 0)
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
  if (fast)  // leaf size 1 <=> tolerance 2 on distinct voxel centres: 26-connectivity, read off the occupancy masks
    RET(vf_cluster_runs26_dev(ctx, ctx->cl_bg, ctx->sep_ds.as<vofod_vox>(), ctx->sep_segcnt.as<uint32_t>(), ctx->sep_segoff.as<uint32_t>(), d_kds, K, nullptr,
                              cnt + CNT_SEP_NCL));
  else
    RET(vf_cluster_dev(ctx, ctx->cl_bg, reinterpret_cast<const float*>(ctx->sep_ds.p), 4, d_kds, K, (float)mv, ctx->sep_labels.as<int>(), cnt + CNT_SEP_NCL,
                       k_cap ? ctx->sep_table_hint : 0));
  if (!fast)  // (the fast path's emission pass cleared its entries)
  {
    const FillJob fj[1] = {{ctx->sep_nsure.as<uint32_t>(), K, 0u}};
    RET(vf_fill(ctx, fj, 1));
  }
  ZERO_CNT(CNT_SEP_ANY_SURE, 1);
  const int nb = vf_blocks(ctx, K, 256, 8);
  if (fast)  // sep_labels = root of every voxel (any member may name a cluster here)
    LAUNCH(k_sep_roots_nsure, nb, 256, 0, ctx->cl_bg.parent.as<int>(), ctx->vg_flags.as<uint32_t>(), d_kds, K, ctx->sep_labels.as<int>(), ctx->sep_nsure.as<int>());
  else
    LAUNCH(k_sep_nsure, nb, 256, 0, ctx->sep_ds.as<vofod_vox>(), (const uint32_t*)nullptr, ctx->sep_labels.as<int>(), d_kds, K, ctx->sep_nsure.as<int>());
  ENSURE(ctx->sep_unsure, K * 4);
  ZERO_CNT(CNT_SEP_NUNSURE, 1);
  LAUNCH(k_sep_any, nb, 256, 0, ctx->sep_labels.as<int>(), ctx->sep_nsure.as<int>(), d_kds, K, min_sure, cnt, ctx->sep_unsure.as<uint32_t>(), fast ? 1 : 0);
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
  {
    LAUNCH(k_sep_state, 1, 1, 0, cnt, cap_guard);
    return VOFOD_OK;
  }
  const float dits = (float)(its_diff > 1 ? its_diff : 1);           // :1210-1217
  float w1 = powf(1.0f - 0.5f, dits);                                 // :1239-1241
  w1 = w1 < 0.0f ? 0.0f : (1.0f < w1 ? 1.0f : w1);
  volatile float w2v = 1.0f - w1;
  const float w2 = w2v;
  // sized for a few thousand listed voxels per wave; the kernel strides over whatever the list holds
  LAUNCH(k_sep_decay, vf_blocks(ctx, (K < 8192 ? K : 8192) * n_off, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, ctx->sep_ds.as<vofod_vox>(),
         ctx->sep_unsure.as<uint32_t>(), K, ctx->sep_offsets.as<int3>(), (int)n_off, w1, w2, (float)p.score_ray, cnt, cap_guard);
  return VOFOD_OK;
}

extern "C" int vofod_sepclusters(vofod_ctx* ctx, int its_diff, const vofod_params* p, int* sure_background_sufficient)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  FLUSH_PENDING();
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
