// Per-scan orchestration of libvofod_cuda: the L3 functions of the nodelet (vofod_nodelet.cpp:581-613, 703-815,
// 882-964) as device stages, and vofod_process_scan, which enqueues a whole scan of the deterministic schedule S1
// on the context's stream with every count kept in DEVICE memory (no host round trip between stages) and reads
// the results back once at the end.  enqueue_scan is the launch sequence that gets captured into the scan's CUDA graph:
// a main chain and two side branches (streams 2 and 3) for the work that does not depend on it.
#include <math.h>

#include "common.cuh"
#include "prims.cuh"

// ---- A23 rangefinder ground seed (vofod_nodelet.cpp:581-613) ----------------------------------------
__global__ void k_range_update(float* __restrict__ score, const Geom g, const ScanDyn* __restrict__ dyn, const double score_point, uint8_t* __restrict__ col_dirty)
{
  pdl_enter();
  range_update(score, g, dyn, score_point, col_dirty);
}

int vf_range_update_dev(vofod_ctx* ctx, const vofod_params& p)
{
  LAUNCH(k_range_update, 1, 1, 0, ctx->score.as<float>(), ctx->g, ctx->dyn.as<ScanDyn>(), p.score_point, ctx->col_dirty.as<uint8_t>());
  return 0;
}

// ---- A11 updateVoxel / updateVMaps (vofod_nodelet.cpp:777-809) ---------------------------------------
// The reference walks the cloud twice (:946-948: points of close clusters with score_point / flag 2, then the others with
// score_unknown / flag 3) and updates the voxel of each point in turn.  Inside a scan the points are the centroids of a
// voxel grid that is aligned with the map (:664-665), one per cell; for any other cloud (the staged entry point, a
// differently aligned filter) two points can share a cell, and then the reference applies both updates, in cloud order.
// (Cloud order is exact for the whole-cloud overload, :811, which is what the staged entry point mirrors.  The per-cluster overload of
// the scan path, :946-948, walks PCL's size-sorted clusters: for a filter that is NOT aligned with the map two points of one cell would
// be applied in cluster order there, and the last bit of w*m + (1-w)*s could differ.  Unreachable with the reference's aligned filter.)
// To get exactly that from a parallel kernel every active point first CLAIMS its cell with
// atomicMin(owner[cell], key), key = (pass << 30) | index = its place in the reference's order.  The update kernel lets
// only the claim holder apply its update (and release the claim); every other point of an already claimed cell goes to a
// left-over list, which the last block to finish sorts by key and applies one by one (normally the list is empty).
// Both passes of a scan run in ONE launch; their claims ride along in k_close_finish.
#define UPD_EMPTY 0xFFFFFFFFu
struct UpdArgs
{
  float score[2];    // [pass]
  uint8_t flag[2];
  int both;          // 1: every point is active, pass = (sel[i] == sel_value ? 0 : 1); 0: only points with sel[i] == sel_value (or all, sel == NULL), pass 0
  int sel_value;
};
__device__ __forceinline__ bool upd_point(const Geom& g, const vofod_vox& v, long long& ci, int& xc, int& yc, int& zc)
{
  xc = coord_to_idx1(v.x, g.off[0], g.inv), yc = coord_to_idx1(v.y, g.off[1], g.inv), zc = coord_to_idx1(v.z, g.off[2], g.inv);
  if (!in_limits_idx(g, xc, yc, zc))  // the reference's vector::at would throw; unreachable after the op-area crop
    return false;
  ci = cell_index(g, xc, yc, zc);
  return ci >= 0;  // slab mode: every held cell (own range AND halo) takes the update, which keeps halos consistent without any exchange
}
__device__ __forceinline__ void upd_apply(float* __restrict__ score, uint8_t* __restrict__ flags, uint8_t* __restrict__ col_dirty, const Geom& g, const vofod_vox& v,
                                          const long long ci, const int xc, const int yc, const int zc, const float vmap_score, const uint8_t vflag)
{
  const unsigned c = v.count > 63u ? 63u : v.count;                   // std::clamp(pt.range, 0u, 63u)
  const float w = 1.0f / (float)(1ull << c);                          // :791
  score[ci] = w * score[ci] + (1.0f - w) * vmap_score;                // :794
  flags[ci] = vflag;                                                  // :796
  col_dirty[dirty_index(g, xc - g.st_lo[0], yc - g.st_lo[1], zc - g.st_lo[2])] = 1;
}
__device__ __forceinline__ void upd_claim(const Geom& g, const vofod_vox& v, const unsigned key, unsigned* __restrict__ owner)
{
  long long ci;
  int xc, yc, zc;
  if (upd_point(g, v, ci, xc, yc, zc))
    atomicMin(owner + ci, key);
}
// claims for the staged entry point (inside a scan k_close_finish places them)
__global__ void __launch_bounds__(256) k_update_claim(const Geom g, const vofod_vox* __restrict__ vox, const uint8_t* __restrict__ sel, const UpdArgs a,
                                                      const unsigned long long* __restrict__ d_m, const size_t m_cap, unsigned* __restrict__ owner)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const bool is_sel = !sel || sel[i] == (uint8_t)a.sel_value;
    if (!a.both && !is_sel)
      continue;
    upd_claim(g, vox[i], ((is_sel ? 0u : 1u) << 30) | (unsigned)i, owner);
  }
}
__global__ void __launch_bounds__(256) k_update_points(float* __restrict__ score, uint8_t* __restrict__ flags, const Geom g, const vofod_vox* __restrict__ vox,
                                                       const uint8_t* __restrict__ sel, const UpdArgs a, const unsigned long long* __restrict__ d_m, const size_t m_cap,
                                                       uint32_t* __restrict__ flagged, const size_t flagged_cap, unsigned long long* __restrict__ counters,
                                                       uint8_t* __restrict__ col_dirty, unsigned* __restrict__ owner, unsigned* __restrict__ leftover)
{
  pdl_enter();
  __shared__ bool s_last;
  const size_t m = prims::dev_count(d_m, m_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const bool is_sel = !sel || sel[i] == (uint8_t)a.sel_value;
    if (!a.both && !is_sel)
      continue;
    const int pass = is_sel ? 0 : 1;
    const unsigned key = ((unsigned)pass << 30) | (unsigned)i;
    const vofod_vox v = vox[i];
    long long ci;
    int xc, yc, zc;
    if (!upd_point(g, v, ci, xc, yc, zc))
      continue;
    if (owner[ci] != key)
    {
      leftover[atomicAdd(counters + CNT_UPD_LEFT, 1ull)] = key;  // < m entries
      continue;
    }
    upd_apply(score, flags, col_dirty, g, v, ci, xc, yc, zc, a.score[pass], a.flag[pass]);
    owner[ci] = UPD_EMPTY;
    const unsigned long long k = atomicAdd(counters + CNT_FLAGGED, 1ull);
    if (k < flagged_cap)
      flagged[k] = (uint32_t)ci;
    else
      counters[CNT_FLAGGED_OVERFLOW] = 1ull;
  }
  // the last block to get here applies the left-overs in the reference's order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(counters + CNT_UPD_TICKET, 1ull) == (unsigned long long)gridDim.x - 1ull;
  __syncthreads();
  if (!s_last || threadIdx.x != 0)
    return;
  __threadfence();
  const unsigned long long n_left = *(volatile unsigned long long*)(counters + CNT_UPD_LEFT);
  for (unsigned long long t = 1; t < n_left; t++)  // insertion sort: the list is empty or tiny
  {
    const unsigned key = *(volatile unsigned*)(leftover + t);
    unsigned long long u = t;
    for (; u > 0 && *(volatile unsigned*)(leftover + u - 1) > key; u--)
      leftover[u] = *(volatile unsigned*)(leftover + u - 1);
    leftover[u] = key;
  }
  for (unsigned long long t = 0; t < n_left; t++)
  {
    const unsigned key = *(volatile unsigned*)(leftover + t);
    const int pass = (int)(key >> 30);
    const vofod_vox v = vox[key & 0x3FFFFFFFu];
    long long ci;
    int xc, yc, zc;
    if (upd_point(g, v, ci, xc, yc, zc))
    {
      volatile float* sc = score + ci;  // written by another block a moment ago
      const unsigned c = v.count > 63u ? 63u : v.count;
      const float w = 1.0f / (float)(1ull << c);
      *sc = w * *sc + (1.0f - w) * a.score[pass];
      flags[ci] = a.flag[pass];
    }
  }
  counters[CNT_UPD_LEFT] = 0ull;
  counters[CNT_UPD_TICKET] = 0ull;
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

static int update_points_launch(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_sel, const UpdArgs& a, const unsigned long long* d_m, size_t m_cap, bool claim)
{
  if (m_cap == 0)
    return 0;
  if (m_cap >= (size_t(1) << 30))
    return vf_fail(ctx, VOFOD_E_INVALID, "too many points for one update");
  RET(ensure_flagged(ctx, 4 * m_cap > (size_t(1) << 20) ? 4 * m_cap : (size_t(1) << 20)));
  ENSURE(ctx->upd_leftover, m_cap * 4);
  RET(vf_update_owner(ctx));
  const int nb = vf_blocks(ctx, m_cap, 256, 8);
  if (claim)
    LAUNCH(k_update_claim, nb, 256, 0, ctx->g, d_vox, d_sel, a, d_m, m_cap, ctx->upd_owner.as<unsigned>());
  LAUNCH(k_update_points, nb, 256, 0, ctx->score.as<float>(), ctx->flags.as<uint8_t>(), ctx->g, d_vox, d_sel, a, d_m, m_cap, ctx->flagged.as<uint32_t>(), ctx->flagged_cap,
         ctx->d_counters.as<unsigned long long>(), ctx->col_dirty.as<uint8_t>(), ctx->upd_owner.as<unsigned>(), ctx->upd_leftover.as<unsigned>());
  return 0;
}
// the claim grid: one word per cell, UPD_EMPTY whenever no update is in flight (claims are released by their holders)
int vf_update_owner(vofod_ctx* ctx)
{
  const size_t n = (size_t)geom_cells(ctx->g);
  if (ctx->upd_owner.cap >= n * 4 && ctx->upd_owner_cells == n)
    return 0;
  ENSURE(ctx->upd_owner, n * 4);
  CK(cudaMemsetAsync(ctx->upd_owner.p, 0xFF, n * 4, ctx->stream));
  ctx->upd_owner_cells = n;
  return 0;
}

// one pass (staged entry point): the points with sel[i] == sel_value (all when sel is NULL)
int vf_update_points_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_sel, int sel_value, const unsigned long long* d_m, size_t m_cap, float score, float flag)
{
  UpdArgs a = {};
  a.score[0] = score;
  a.flag[0] = (uint8_t)flag;
  a.both = 0;
  a.sel_value = sel_value;
  return update_points_launch(ctx, d_vox, d_sel, a, d_m, m_cap, true);
}
// both passes of a scan (:946-948); the claims were placed by k_close_finish
int vf_update_points_scan_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_in_close, const unsigned long long* d_m, size_t m_cap, const vofod_params& p)
{
  UpdArgs a = {};
  a.score[0] = (float)p.score_point;
  a.flag[0] = 2;
  a.score[1] = (float)p.score_unknown;
  a.flag[1] = 3;
  a.both = 1;
  a.sel_value = 1;
  return update_points_launch(ctx, d_vox, d_in_close, a, d_m, m_cap, false);
}

// ---- A14/A15 findCloseFarClusters (vofod_nodelet.cpp:703-750) ----------------------------------------
__global__ void k_bg_state(unsigned long long* __restrict__ counters, const unsigned long long min_sufficient)
{
  pdl_enter();
  if (counters[CNT_NBG] > min_sufficient)  // :716-721
    counters[CNT_STATE_BG] = 1ull;
}

// one WARP per point: the lanes stride over the <= (2*mv)^3 window cells of VoxelMap::hasCloseTo (voxel_map.cpp:376-400),
// 8 independent loads per lane in flight (the default window, 6^3 cells, is one such batch: one memory round trip).
// The per-point result goes to pt_hit; clusters are marked afterwards (a shared per-cluster flag polled and written from
// here funnels every warp into one L2 sector once the ground cluster is background: measured 177 us instead of 10).
// Second part (n_bg_out != NULL): nVoxelsOver(thr) of :712 over the raised (chunk, column) cells — the same threshold, the
// same grid, no dependency between the two, so they share a launch.
__global__ void __launch_bounds__(256) k_close_points(const float* __restrict__ score, const Geom g, const vofod_vox* __restrict__ vox,
                                                      const unsigned long long* __restrict__ d_m, const size_t m_cap, const float max_dist, const float thr,
                                                      uint8_t* __restrict__ pt_hit, int* __restrict__ cl_close, const uint8_t* __restrict__ dirty,
                                                      unsigned long long* __restrict__ n_bg_out)
{
  pdl_enter();
  const size_t m = vox ? prims::dev_count(d_m, m_cap) : 0;
  const unsigned lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const float md = max_dist * g.inv;
  const int mv = (int)ceilf(md);
  for (size_t i = warp0; i < m; i += n_warps)
  {
    const vofod_vox v = vox[i];
    const int ox = coord_to_idx1(v.x, g.off[0], g.inv), oy = coord_to_idx1(v.y, g.off[1], g.inv), oz = coord_to_idx1(v.z, g.off[2], g.inv);
    if (!cell_owned(g, ox, oy, oz))
    {
      // slab mode: the slab owning the point's voxel holds the whole window (halo >= mv) and answers for it
      if (lane == 0)
      {
        pt_hit[i] = 0;
        cl_close[i] = 0;
      }
      continue;
    }
    const int bx = max(ox - mv, 0), by = max(oy - mv, 0), bz = max(oz - mv, 0);
    const int ex = min(ox + mv, g.size[0]), ey = min(oy + mv, g.size[1]), ez = min(oz + mv, g.size[2]);
    const int nx = ex - bx, ny = ey - by, nz = ez - bz;
    bool hit = false;
    if (nx > 0 && ny > 0 && nz > 0)
    {
      // The answer is "any cell of the window ...", so the order of the cells is free: the point's own z-plane goes first, in
      // a round of its own — for a point on mapped background its neighbours in that plane already decide, and the other
      // planes (5/6 of the window's memory sectors, which is what this kernel is made of) are never fetched.
      const int nxy = nx * ny;
      const int zown = (oz >= bz && oz < ez) ? oz - bz : 0;  // window plane visited first
      const int total = nxy * nz;
      for (int base = 0; base < total && !hit;)
      {
        const int span = base == 0 ? nxy : 32 * 8;  // first round: one plane (<= 2 loads per lane with the default window)
        const int end = min(base + span, total);
        for (int b0 = base; b0 < end && !hit; b0 += 32 * 8)
        {
          float val[8];
#pragma unroll
          for (int q = 0; q < 8; q++)
          {
            const int t = b0 + q * 32 + (int)lane;
            val[q] = __int_as_float(0xff800000);
            if (t < end)
            {
              const int zr = t / nxy, rem = t - zr * nxy;
              const int zi = zr == 0 ? zown : (zr <= zown ? zr - 1 : zr);
              const long long ci = cell_index(g, bx + rem % nx, by + rem / nx, bz + zi);
              if (ci >= 0)
                val[q] = score[ci];
            }
          }
          bool h = false;
#pragma unroll
          for (int q = 0; q < 8; q++)
            if (val[q] > thr)
            {
              const int t = b0 + q * 32 + (int)lane;
              const int zr = t / nxy, rem = t - zr * nxy;
              const int zi = zr == 0 ? zown : (zr <= zown ? zr - 1 : zr);
              const int ddx = bx + rem % nx - ox, ddy = by + rem / nx - oy, ddz = bz + zi - oz;
              // Eigen int-vector norm(): int(sqrt(d2)) <= md; the integer square root of d2 <= 3*mv^2 is exact in fp32 too
              const int nrm = (int)sqrtf((float)(ddx * ddx + ddy * ddy + ddz * ddz));
              h = h || (float)nrm <= md;
            }
          hit = __any_sync(VOFOD_FULL, h);
        }
        base = end;
      }
    }
    if (lane == 0)
    {
      pt_hit[i] = hit ? 1 : 0;
      cl_close[i] = 0;  // per-cluster flags (indexed by label < m): cleared here, set by k_close_mark
    }
  }
  if (!n_bg_out)
    return;
  // nVoxelsOver: one thread per (chunk of DIRTY_ZC levels, column), column fastest (coalesced marks and scores)
  const size_t ncol = (size_t)g.st_size[0] * g.st_size[1];
  const size_t n_items = ncol * (size_t)dirty_chunks(g);
  unsigned cnt = 0;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (size_t)gridDim.x * blockDim.x)
  {
    if (dirty && !dirty[it])
      continue;
    const size_t c = it % ncol;
    const int zc = (int)(it / ncol);
    if (!column_owned(g, (int)(c % g.st_size[0]), (int)(c / g.st_size[0])))
      continue;
    const int z_lo = zc * DIRTY_ZC, z_hi = min(z_lo + DIRTY_ZC, g.st_size[2]);
    for (int z0 = z_lo; z0 < z_hi; z0 += 16)
    {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++)
        v[k] = (z0 + k < z_hi) ? score[c + (size_t)(z0 + k) * ncol] : __int_as_float(0xff800000);
#pragma unroll
      for (int k = 0; k < 16; k++)
        cnt += v[k] > thr;
    }
  }
  cnt = prims::warp_sum(cnt);
  if (lane == 0 && cnt)
    atomicAdd(n_bg_out, (unsigned long long)cnt);
}
// a cluster is close iff ANY of its points is (:730-741)
// (+ the background-sufficient latch of :716-721 when `counters` is given: the count it reads was finished a kernel ago)
__global__ void __launch_bounds__(256) k_close_mark(const uint8_t* __restrict__ pt_hit, const int* __restrict__ labels, const unsigned long long* __restrict__ d_m,
                                                    const size_t m_cap, int* __restrict__ cl_close, unsigned long long* counters, const unsigned long long min_sufficient)
{
  pdl_enter();
  if (counters && blockIdx.x == 0 && threadIdx.x == 0 && counters[CNT_NBG] > min_sufficient)
    counters[CNT_STATE_BG] = 1ull;
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < m; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool hit = i < m && pt_hit[i] != 0;
    const int l = hit ? labels[i] : -1 - (int)lane;
    const unsigned grp = __match_any_sync(VOFOD_FULL, l);
    if (hit && lane == (unsigned)(__ffs(grp) - 1))
      cl_close[l] = 1;
  }
}
// + (owner != NULL) the cell claims of the point update that follows inside a scan (see k_update_points)
__global__ void __launch_bounds__(256) k_close_finish(const int* __restrict__ labels, const unsigned long long* __restrict__ d_m, const size_t m_cap,
                                                      const int* __restrict__ cl_close, uint8_t* __restrict__ pt_close, unsigned long long* __restrict__ counters,
                                                      const Geom g, const vofod_vox* __restrict__ vox, unsigned* __restrict__ owner)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  unsigned n_close = 0, n_far = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const int l = labels[i];
    const int c = cl_close[l];
    pt_close[i] = (uint8_t)c;
    if (owner)
      upd_claim(g, vox[i], ((c != 0 ? 0u : 1u) << 30) | (unsigned)i, owner);
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

// phase 1 = everything up to the cross-slab exchange point (local background count, per-cluster close flags of the points this
// slab answers for), phase 2 = the rest, phase 0 = both (unsharded)
// nVoxelsOver(thr_new_obstacles) into CNT_NBG (zeroed by the scan's first kernel): the counting part of k_close_points alone
int vf_count_bg_dev(vofod_ctx* ctx, const vofod_params& p)
{
  const float thr = (float)p.thr_new_obstacles;
  const size_t items = (size_t)ctx->g.st_size[0] * ctx->g.st_size[1] * dirty_chunks(ctx->g);
  LAUNCH(k_close_points, vf_blocks(ctx, items, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, (const vofod_vox*)nullptr, (const unsigned long long*)nullptr, (size_t)0,
         0.0f, thr, (uint8_t*)nullptr, (int*)nullptr, vf_dirty_cols(ctx, thr, &p), vf_cnt(ctx, CNT_NBG));
  return 0;
}

// hasCloseTo of every point (the part of findCloseFarClusters that does not look at the labels), ahead of time
int vf_close_points_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const unsigned long long* d_m, size_t m_cap, const vofod_params& p)
{
  if (m_cap == 0)
    return 0;
  ENSURE(ctx->cl_close, m_cap * 4);
  ENSURE(ctx->pt_close, m_cap + 64);
  LAUNCH(k_close_points, vf_blocks(ctx, m_cap * 32, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, d_vox, d_m, m_cap, (float)p.ground_points_max_distance,
         (float)p.thr_new_obstacles, ctx->pt_close.as<uint8_t>(), ctx->cl_close.as<int>(), (const uint8_t*)nullptr, (unsigned long long*)nullptr);
  ctx->close_points_done = true;
  return 0;
}

int vf_close_far_phase(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const unsigned long long* d_m, size_t m_cap, const vofod_params& p, int phase,
                       bool claim_for_update)
{
  const float max_dist = (float)p.ground_points_max_distance;
  const float thr = (float)p.thr_new_obstacles;
  // vofod_nodelet.cpp:229-230 (fp32, left to right), converted to uint64_t
  const float vs = ctx->cfg_voxel_size > 0.f ? ctx->cfg_voxel_size : ctx->g.vs;
  volatile float a0 = p.oparea_size[0] / vs;
  volatile float a1 = a0 * p.oparea_size[1];
  volatile float a2 = a1 / vs;
  volatile float a3 = a2 * p.background_sufficient_points_ratio;
  const unsigned long long min_sufficient = (unsigned long long)a3;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  bool state_done = false;
  if (phase != 2)
  {
    if (m_cap)
    {
      // :712 nVoxelsOver rides in the hasCloseTo kernel
      ENSURE(ctx->cl_close, m_cap * 4);
      ENSURE(ctx->pt_close, m_cap + 64);
      const bool precounted = ctx->nbg_precounted;  // (the scan's side branch has done it)
      ctx->nbg_precounted = false;
      if (!ctx->scan_prezero && !precounted)
        CK(cudaMemsetAsync(cnt + CNT_NBG, 0, sizeof(unsigned long long), ctx->stream));
      const size_t items = (size_t)ctx->g.st_size[0] * ctx->g.st_size[1] * dirty_chunks(ctx->g);
      const size_t work = m_cap * 32 > items || precounted ? m_cap * 32 : items;
      const bool points_done = ctx->close_points_done && precounted;  // (the scan's second side branch has done it)
      ctx->close_points_done = false;
      if (!points_done)
        LAUNCH(k_close_points, vf_blocks(ctx, work, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, d_vox, d_m, m_cap, max_dist, thr, ctx->pt_close.as<uint8_t>(),
               ctx->cl_close.as<int>(), vf_dirty_cols(ctx, thr, &p), precounted ? nullptr : cnt + CNT_NBG);
      state_done = phase == 0;  // unsharded: the count is final, the latch can ride in the next kernel
      LAUNCH(k_close_mark, vf_blocks(ctx, m_cap, 256, 8), 256, 0, ctx->pt_close.as<uint8_t>(), d_labels, d_m, m_cap, ctx->cl_close.as<int>(),
             state_done ? cnt : nullptr, min_sufficient);
    } else if (!ctx->nbg_precounted)
      RET(vf_count_over_dev(ctx, thr, cnt + CNT_NBG, &p));
    ctx->nbg_precounted = false;
  }
  if (phase != 1)
  {
    if (!state_done)
      LAUNCH(k_bg_state, 1, 1, 0, cnt, min_sufficient);
    ZERO_CNT(CNT_NCLOSE, 2);  // NCLOSE, NFAR
    if (m_cap)
    {
      if (claim_for_update)
        RET(vf_update_owner(ctx));
      LAUNCH(k_close_finish, vf_blocks(ctx, m_cap, 256, 8), 256, 0, d_labels, d_m, m_cap, ctx->cl_close.as<int>(), ctx->pt_close.as<uint8_t>(), cnt, ctx->g, d_vox,
             claim_for_update ? ctx->upd_owner.as<unsigned>() : nullptr);
    }
  }
  return 0;
}
int vf_close_far_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const unsigned long long* d_m, size_t m_cap, const vofod_params& p, bool claim_for_update)
{
  return vf_close_far_phase(ctx, d_vox, d_labels, d_m, m_cap, p, 0, claim_for_update);
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
  FLUSH_PENDING();
  FLUSH_PENDING();
  if (!world_pt || !p)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  for (int a = 0; a < 3; a++)
    ctx->h_dyn->range_pt[a] = world_pt[a];
  ctx->h_dyn->n_seeds = 1;
  RET(vf_dyn_push(ctx));
  RET(vf_range_update_dev(ctx, *p));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_update_points(vofod_ctx* ctx, const vofod_vox* pts, const uint8_t* sel, int sel_value, size_t n, float score, float flag)
{
  NEED_MAP();
  FLUSH_PENDING();
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
  FLUSH_PENDING();
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
  RET(vf_close_far_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), nullptr, m, *p, false));
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

// ======================================================================================================
// vofod_process_scan: one scan of schedule S1
// ======================================================================================================

// Enqueues every device operation of one scan on ctx->stream.  Called directly (eager mode) or under stream capture
// (graph mode): it must not synchronise, allocate or touch pageable host memory when plan.sep_cap != 0.
static int enqueue_scan(vofod_ctx* ctx, const ScanPlan& plan, int* sep_status_out, bool* applied_out)
{
  cudaStream_t st = ctx->stream;
  const vofod_params& p = plan.p;
  const vofod_schedule& s = plan.s;
  const size_t n = plan.n;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  int e = 0;
#define STAGE_EVENT()                         \
  do                                          \
  {                                           \
    if (plan.timed)                           \
      CK(cudaEventRecord(ctx->ev[e], st));    \
    e++;                                      \
  } while (0)
  struct Prezero
  {
    vofod_ctx* c;
    explicit Prezero(vofod_ctx* c_) : c(c_) { c->scan_prezero = true; }
    ~Prezero() { c->scan_prezero = false; }
  } prezero_guard(ctx);
  // (the ScanDyn upload and the read-back of the counters are enqueued by the caller, outside a captured graph: their host addresses
  //  belong to the pinned slot of the scan, and a graph would bake them in)
  RET(vf_begin_scan(ctx, p, !plan.sep_first));  // + rangefinder seeds (A23; after a deferred pass when there is one) + the filter's min/max reset
  // The raycast accumulate reads only the scan, the LUT and the per-scan arguments and writes only the accumulator window:
  // it is independent of the whole filter -> cluster -> close/far -> point-update chain.  In replay mode it runs as a
  // parallel branch of the graph (issue-bound kernel next to a chain of latency-bound ones); with per-stage timing on it
  // stays in line so that the stage table means what it says.
  // The stage-local clears (hash table of the clustering, work arrays of classification and sepclusters) depend on nothing
  // either: they open the side branch, so that the main chain finds them done.
  const bool side = !plan.timed && ctx->stream2 != nullptr && ctx->overlap_enabled;
  const bool overlap_raycast = plan.raycast_on && side && !plan.apply_first;
  if (side)
  {
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    ctx->stream = ctx->stream2;
    RunRows rows_probe;
    const bool hash_cluster = ctx->cl_force_hash || ctx->vg_force_sort || !vf_run_rows((float)p.ground_points_max_distance, ctx->g.vs, rows_probe);
    // a deferred pass is the first thing on the main chain: its clears come first here, with an event of their own for it to wait on
    // (without the wait the pass raced these fills — its counts could be wiped under it, and it then found an empty background cloud)
    int rrc = 0;
    if (plan.sep_first)
    {
      rrc = vf_sepclusters_prefill(ctx, plan.sep_first_p);
      if (rrc >= 0)
        CK(cudaEventRecord(ctx->ev_sepfill, ctx->stream2));
    }
    if (rrc >= 0 && hash_cluster)
      rrc = vf_cluster_prefill(ctx, ctx->cl, n, 0);
    if (rrc >= 0 && s.do_classify)
      rrc = vf_classify_prefill(ctx, n);
    if (rrc >= 0 && s.do_sepclusters && !plan.sep_first)
      rrc = vf_sepclusters_prefill(ctx, p);
    // nVoxelsOver of :712 only needs the map as the previous scan (and this scan's rangefinder seed) left it — with a deferred
    // separated-background pass still to come it has to wait for that (it then rides in the hasCloseTo kernel)
    if (rrc >= 0 && !plan.sep_first)
      rrc = vf_count_bg_dev(ctx, p);
    ctx->stream = st;
    if (rrc < 0)
      return rrc;
    ctx->nbg_precounted = !plan.sep_first;
    CK(cudaEventRecord(ctx->ev_fills, ctx->stream2));
    if (overlap_raycast)
    {
      ctx->stream = ctx->stream2;
      rrc = vf_raycast_accumulate_dev(ctx, n, vofod_pose(), p);
      ctx->stream = st;
      if (rrc < 0)
        return rrc;
    }
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  }
  STAGE_EVENT();
  // The previous scan's separated-background pass (vofod_schedule::sep_deferred): it must precede everything of this scan that touches
  // the map (first of all the rangefinder seeds), but this scan's front end — crop, voxel grid, clustering — never looks at the map:
  // in replay mode the front end runs on a branch of its own beside the pass.
  *sep_status_out = VOFOD_W_PAUSED;
  const bool front_side = plan.sep_first && side && ctx->stream4 != nullptr && !ctx->vg_force_sort && !ctx->cl_force_hash;
  if (front_side)
  {
    CK(cudaEventRecord(ctx->ev_fork4, st));
    CK(cudaStreamWaitEvent(ctx->stream4, ctx->ev_fork4, 0));
    CK(cudaStreamWaitEvent(ctx->stream4, ctx->ev_fills, 0));
  }
  if (plan.sep_first)
  {
    if (side)
      CK(cudaStreamWaitEvent(st, ctx->ev_sepfill, 0));
    std::swap(ctx->tile_state, ctx->tile_state_b);
    std::swap(ctx->tile_state2, ctx->tile_state2_b);
    const int src = vf_sepclusters_dev(ctx, plan.sep_first_its, plan.sep_first_p, plan.sep_cap);
    std::swap(ctx->tile_state, ctx->tile_state_b);
    std::swap(ctx->tile_state2, ctx->tile_state2_b);
    if (src < 0)
      return src;
    *sep_status_out = src;
    RET(vf_range_update_dev(ctx, p));  // this scan's rangefinder seeds (A23)
  }
  STAGE_EVENT();  // 0 "range": the rangefinder seeds (A23) ride in the scan's first kernel (vf_begin_scan)
  // filterAndTransform (:928)
  if (front_side)
    ctx->stream = ctx->stream4;
  const int seeded = vf_filter_voxelize_dev(ctx, n, p, true);
  if (seeded < 0)
  {
    ctx->stream = st;
    return seeded;
  }
  STAGE_EVENT();  // 1 "filtering"
  // hasCloseTo of the voxels does not look at their labels: second side branch, next to the clustering
  const bool cp_side = side && ctx->stream3 != nullptr && ctx->nbg_precounted;
  if (cp_side)
  {
    CK(cudaEventRecord(ctx->ev_fork3, st));
    CK(cudaStreamWaitEvent(ctx->stream3, ctx->ev_fork3, 0));
    ctx->stream = ctx->stream3;
    const int prc = vf_close_points_dev(ctx, ctx->vox.as<vofod_vox>(), cnt + CNT_VG_M, n, p);
    ctx->stream = st;
    if (prc < 0)
      return prc;
    CK(cudaEventRecord(ctx->ev_cp, ctx->stream3));
  }
  // clusterCloud (:932)
  if (side && !front_side)
    CK(cudaStreamWaitEvent(st, ctx->ev_fills, 0));
  ENSURE(ctx->labels, n * 4);
  if (seeded)
  {
    // the voxel list is a set of grid cells: connected components on the occupancy words the filter left behind
    RunRows rows;
    vf_run_rows((float)p.ground_points_max_distance, ctx->g.vs, rows);
    RET(vf_cluster_runs_dev(ctx, ctx->cl, ctx->cl_cellkey.as<uint32_t>(), ctx->cl_words.as<RunWord>(),
                            reinterpret_cast<const VgLayout*>(ctx->scratch_d.as<char>() + 64), rows, (float)p.ground_points_max_distance, cnt + CNT_VG_M, n,
                            ctx->labels.as<int>(), cnt + CNT_NCLUSTERS));
  } else
    RET(vf_cluster_dev(ctx, ctx->cl, reinterpret_cast<const float*>(ctx->vox.p), 4, cnt + CNT_VG_M, n, (float)p.ground_points_max_distance, ctx->labels.as<int>(),
                       cnt + CNT_NCLUSTERS));
  if (front_side)
  {
    ctx->stream = st;
    CK(cudaEventRecord(ctx->ev_front, ctx->stream4));
    CK(cudaStreamWaitEvent(st, ctx->ev_front, 0));
    CK(cudaStreamWaitEvent(st, ctx->ev_fills, 0));
  }
  STAGE_EVENT();  // 2 "clusterization"
  // findCloseFarClusters (:936)
  if (cp_side)
    CK(cudaStreamWaitEvent(st, ctx->ev_cp, 0));
  RET(vf_close_far_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), cnt + CNT_VG_M, n, p, true));
  STAGE_EVENT();  // 3 "close X far"
  // classifyClusters, the part that only looks at the voxel list: on the side branch, next to the point update and the ray apply
  const bool cls_side = side && s.do_classify;
  if (cls_side)
  {
    CK(cudaEventRecord(ctx->ev_fork2, st));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork2, 0));
    ctx->stream = ctx->stream2;
    const int crc = vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p, 1);
    ctx->stream = st;
    if (crc < 0)
      return crc;
    CK(cudaEventRecord(ctx->ev_cls, ctx->stream2));
  }
  // updateVMaps (:946-949)
  RET(vf_update_points_scan_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p));
  STAGE_EVENT();  // 4 "vmap update"
  *applied_out = false;
  if (side)
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));  // join: the apply needs the accumulator
  if (plan.apply_first)
  {
    // the raycast thread wakes up after this scan's point update: apply the pending accumulate, clear the flags (:1530-1602)
    const int rc = vf_raycast_apply_dev(ctx, 0, p);
    if (rc < 0)
      return rc;
    *applied_out = rc == VOFOD_OK;
    ZERO_CNT(CNT_TRAVERSALS, 1);
    ZERO_CNT(CNT_OOB, 1);
  } else if (overlap_raycast)
    ;
  else if (plan.raycast_on)
    RET(vf_raycast_accumulate_dev(ctx, n, vofod_pose(), p));
  else
  {
    ZERO_CNT(CNT_TRAVERSALS, 1);
    ZERO_CNT(CNT_OOB, 1);
    // the reference clears m_voxel_raycast before the sensor-in-map test (:1430-1432)
    if (s.do_raycast && plan.raycast_status == VOFOD_W_SENSOR_OOB && ctx->acc_has_data && ctx->acc.p)
    {
      CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_total_bytes, st));
      ctx->acc_has_data = false;
    }
  }
  STAGE_EVENT();  // 5 "raycasting"
  if (plan.apply_on && !plan.apply_first)
  {
    const int rc = vf_raycast_apply_dev(ctx, 0, p);
    if (rc < 0)
      return rc;
    *applied_out = rc == VOFOD_OK;
  }
  STAGE_EVENT();  // 6 raycast "vmap update"
  ZERO_CNT(CNT_NDET, 1);
  if (cls_side)
  {
    CK(cudaStreamWaitEvent(st, ctx->ev_cls, 0));
    RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p, 2));
  } else if (s.do_classify)
    RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), cnt + CNT_VG_M, n, p));
  STAGE_EVENT();  // 7 "classification" (+ 8 detections, fused)
  STAGE_EVENT();
  if (!plan.sep_first)
    ZERO_CNT(CNT_SEP_K, 1);
  if (s.do_sepclusters && !s.sep_deferred)
  {
    const int rc = vf_sepclusters_dev(ctx, s.sep_its_diff, p, plan.sep_cap);
    if (rc < 0)
      return rc;
    *sep_status_out = rc;
  }
  STAGE_EVENT();  // 9 "sep bg clusters"
  STAGE_EVENT();  // 10 "readback" (enqueued by the caller right behind this sequence)
#undef STAGE_EVENT
  return 0;
}

static uint64_t fnv1a(const void* data, size_t len, uint64_t h)
{
  const unsigned char* b = (const unsigned char*)data;
  for (size_t i = 0; i < len; i++)
  {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}

// pinned slot `slot` of the context: 64 result counters (+ spare), the ScanDyn source, the first PIPE_DETS detection records
#define PIPE_DETS 32
static inline unsigned long long* slot_counters(vofod_ctx* ctx, int slot) { return (unsigned long long*)((char*)ctx->pinned + slot * 2048); }
static inline ScanDyn* slot_dyn(vofod_ctx* ctx, int slot) { return reinterpret_cast<ScanDyn*>((char*)ctx->pinned + 65536 + slot * 1024); }
static inline vofod_detection* slot_dets(vofod_ctx* ctx, int slot) { return reinterpret_cast<vofod_detection*>((char*)ctx->pinned + 131072 + slot * 16384); }

// A scan = scan_launch (everything the host enqueues: no wait) + scan_finish (wait for it, read the results, host-side bookkeeping).
// vofod_process_scan runs them back to back; vofod_process_scan_batch keeps two scans in flight.
static int scan_launch(vofod_ctx* ctx, const vofod_pt* d_scan, size_t n, const vofod_pose& tf, const vofod_params& p, const vofod_schedule& s, const int slot,
                       const bool pipelined)
{
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  cudaStream_t st = ctx->stream;
  ctx->h_dyn = slot_dyn(ctx, slot);
  ScanInFlight& fl = ctx->fl[slot];
  fl = ScanInFlight();
  fl.pipelined = pipelined;
  // a pending deferred pass runs at the start of this call, beside the front end — unless this scan wants a pass of its own right away
  if (ctx->sep_pending && s.do_sepclusters && !s.sep_deferred)
    RET(vf_flush_pending(ctx));
  ScanPlan plan = {};
  plan.n = n;
  plan.p = p;
  plan.s = s;
  plan.sep_first = ctx->sep_pending;
  plan.sep_first_its = ctx->sep_pending_its;
  if (plan.sep_first)
    plan.sep_first_p = ctx->sep_pending_p;
  const bool sep_inline = s.do_sepclusters && !s.sep_deferred && !p.sep_pause;
  const bool sep_ran = plan.sep_first || sep_inline;
  plan.raycast_status = VOFOD_W_PAUSED;
  plan.raycast_on = false;
  plan.apply_first = s.raycast_apply_pending && ctx->ray_pending && !p.raycast_pause;
  if (plan.apply_first)
  {
    // a raycast is "in flight": like the reference (:952-957) no new one starts in this call, whatever do_raycast says
    ctx->h_dyn->win_apply = ctx->ray_pending_win;
    plan.raycast_status = VOFOD_OK;
    plan.apply_on = true;
  } else if (s.do_raycast)
  {
    plan.raycast_status = vf_raycast_prepare(ctx, n, tf, p);  // host only; fills h_dyn->win
    if (plan.raycast_status < 0)
      return plan.raycast_status;
    plan.raycast_on = plan.raycast_status == VOFOD_OK;
    plan.apply_on = plan.raycast_on && !s.raycast_defer_apply;
  }
  // per-scan dynamic arguments
  ScanDyn* hd = ctx->h_dyn;
  memcpy(hd->tf.R, tf.R, sizeof(tf.R));
  memcpy(hd->tf.t, tf.t, sizeof(tf.t));
  for (int a = 0; a < 3; a++)
    hd->range_pt[a] = s.range_pt[a];
  hd->n_seeds = s.n_range_seeds > 0 ? s.n_range_seeds : 0;
  hd->scan = d_scan;
  hd->its_raycast = s.raycast_its_diff > 1 ? s.raycast_its_diff : 1;

  // ---- signature of the launch sequence: everything by-value or host-decided that enqueue_scan depends on
  uint64_t sig = 1469598103934665603ull;
  sig = fnv1a(&p, sizeof(p), sig);
  const int flags[10] = {s.do_raycast, s.do_classify, s.do_sepclusters, s.sep_its_diff, plan.raycast_on ? 1 : 0, plan.raycast_status, ctx->flags_full_dirty ? 1 : 0,
                         ctx->acc_has_data ? 1 : 0, plan.apply_on ? 1 : 0, plan.apply_first ? 1 : 0};
  sig = fnv1a(flags, sizeof(flags), sig);
  sig = fnv1a(&n, sizeof(n), sig);
  sig = fnv1a(&ctx->g, sizeof(ctx->g), sig);
  sig = fnv1a(&ctx->acc_cells_max, sizeof(size_t), sig);
  sig = fnv1a(&ctx->frac_bits, sizeof(int), sig);
  sig = fnv1a(&ctx->sep_cap, sizeof(size_t), sig);
  sig = fnv1a(&ctx->sep_table_hint, sizeof(size_t), sig);
  sig = fnv1a(&ctx->alloc_gen, sizeof(uint64_t), sig);
  const int dirty_mode = ctx->col_all_dirty ? 1 : 0;
  sig = fnv1a(&dirty_mode, sizeof(int), sig);
  const int sepf[3] = {plan.sep_first ? 1 : 0, plan.sep_first ? plan.sep_first_its : 0, s.sep_deferred};
  sig = fnv1a(sepf, sizeof(sepf), sig);
  if (plan.sep_first)
    sig = fnv1a(&plan.sep_first_p, sizeof(vofod_params), sig);
  sig = fnv1a(&ctx->untouched_max, sizeof(float), sig);

  int sep_status = VOFOD_W_PAUSED;
  bool applied = false;
  bool used_graph = false;
  const bool epoch_wrap_soon = (((ctx->epoch_calls + 1) * EPOCH_STRIDE) & 0x3fffffffull) < EPOCH_STRIDE;
  const bool graph_ok = ctx->graph_enabled && !epoch_wrap_soon && (!sep_ran || ctx->sep_cap > 0);
  RET(vf_dyn_push(ctx));
  CK(cudaEventRecord(ctx->ev[0], st));
  // what the accumulator holds after this scan
  const bool acc_after = plan.apply_on ? !p.raycast_new_update_rule
                                       : (plan.raycast_on ? true : ((s.do_raycast && plan.raycast_status == VOFOD_W_SENSOR_OOB) ? false : ctx->acc_has_data));
  vofod_ctx::GraphSlot* hit = nullptr;
  vofod_ctx::GraphSlot* seen = nullptr;
  for (auto& gs : ctx->gslot)
  {
    if (gs.exec && gs.sig == sig)
      hit = &gs;
    if (gs.seen_sig == sig && gs.seen_gen == ctx->alloc_gen)
      seen = &gs;
  }
  if (graph_ok && hit)
  {
    // ---- replay
    plan.sep_cap = ctx->sep_cap;  // the capacity the captured pass runs with: the overflow check below needs it
    ctx->epoch_calls++;
    ctx->stat_replays++;
    hit->last_use = ++ctx->gslot_clock;
    CK(cudaGraphLaunch(hit->exec, st));
    ctx->n_launches += hit->kernels;
    sep_status = sep_ran ? VOFOD_OK : VOFOD_W_PAUSED;
    applied = plan.apply_on;
    ctx->acc_has_data = acc_after;
    used_graph = true;
  } else if (graph_ok && seen)
  {
    // ---- a kernel-by-kernel scan with this signature ran before without any allocation: capture, instantiate, launch
    plan.sep_cap = ctx->sep_cap;
    plan.timed = false;
    const bool acc_before = ctx->acc_has_data;
    const uint64_t calls_before = ctx->epoch_calls;
    const uint64_t launches_before = ctx->n_launches;
    cudaGraph_t graph = nullptr;
    ctx->capturing = true;
    ctx->capture_broken = false;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int erc = ce == cudaSuccess ? enqueue_scan(ctx, plan, &sep_status, &applied) : -1;
    if (ce == cudaSuccess)
      ce = cudaStreamEndCapture(st, &graph);
    ctx->capturing = false;
    cudaGraphExec_t exec = nullptr;
    ctx->stat_captures++;
    int why = 0;
    if (ce != cudaSuccess)
      why = graph ? 2 : 5;
    else if (erc != 0)
      why = 1;
    else if (ctx->capture_broken)
      why = 3;
    else if (!graph)
      why = 2;
    else if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
      why = 4;
    if (why)
    {
      ctx->stat_capture_failures++;
      ctx->stat_last_capture_error = why;
    }
    if (why == 0)
    {
      if (seen->exec)
        cudaGraphExecDestroy(seen->exec);
      seen->exec = exec;
      seen->sig = sig;
      seen->kernels = ctx->n_launches - launches_before;
      seen->last_use = ++ctx->gslot_clock;
      cudaGraphDestroy(graph);
      CK(cudaGraphLaunch(seen->exec, st));
      ctx->acc_has_data = acc_after;
      used_graph = true;
    } else
    {
      // capture failed: forget it and run this scan eagerly
      if (graph)
        cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->acc_has_data = acc_before;
      ctx->epoch_calls = calls_before;
      ctx->n_launches = launches_before;
      seen->seen_sig = 0;
      ctx->err.clear();
    }
  }
  if (!used_graph)
  {
    // ---- eager
    ctx->stat_eager++;
    plan.sep_cap = (graph_ok && ctx->sep_cap > 0) ? ctx->sep_cap : 0;
    plan.timed = true;
    const uint64_t gen_before = ctx->alloc_gen;
    RET(enqueue_scan(ctx, plan, &sep_status, &applied));
    ctx->acc_has_data = acc_after;
    if (gen_before == ctx->alloc_gen)
    {
      // remember the signature in the slot used longest ago (a slot that already remembers it, first)
      vofod_ctx::GraphSlot* slot = &ctx->gslot[0];
      for (auto& gs : ctx->gslot)
        if (gs.last_use < slot->last_use)
          slot = &gs;
      for (auto& gs : ctx->gslot)
        if (gs.sig == sig || gs.seen_sig == sig)
          slot = &gs;
      slot->seen_sig = sig;
      slot->seen_gen = ctx->alloc_gen;
      slot->last_use = ++ctx->gslot_clock;
    }
  }
  if (plan.raycast_on && !plan.apply_on)
  {
    ctx->ray_pending = true;  // its apply comes with a later scan (raycast_apply_pending)
    ctx->ray_pending_win = ctx->win;
  } else if (plan.apply_first)
    ctx->ray_pending = false;
  ctx->detection_its++;
  // one read-back of every count into the scan's pinned slot (detection records follow on demand: most scans have none — a pipelined
  // scan takes the first PIPE_DETS along, because the next scan's classification reuses the device array)
  CK(cudaMemcpyAsync(slot_counters(ctx, slot), ctx->d_counters.p, CNT_N_SLOTS * 8, cudaMemcpyDeviceToHost, st));
  if (pipelined && s.do_classify && ctx->dets.p)
    CK(cudaMemcpyAsync(slot_dets(ctx, slot), ctx->dets.p, PIPE_DETS * sizeof(vofod_detection), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[11], st));
  CK(cudaEventRecord(ctx->ev_done[slot], st));
  // the pass this scan asked for waits for the next call (or vofod_flush)
  ctx->sep_pending = s.do_sepclusters && s.sep_deferred && !p.sep_pause;
  if (ctx->sep_pending)
  {
    ctx->sep_pending_its = s.sep_its_diff;
    ctx->sep_pending_p = p;
  }
  fl.plan = plan;
  fl.sep_status = sep_status;
  fl.applied = applied;
  fl.used_graph = used_graph;
  fl.sep_ran = sep_ran;
  fl.active = true;
  return VOFOD_OK;
}

static int scan_finish(vofod_ctx* ctx, const int slot, vofod_scan_result* res, vofod_detection* dets, size_t det_cap)
{
  ScanInFlight& fl = ctx->fl[slot];
  if (res)
    memset(res, 0, sizeof(*res));
  if (!fl.active)
    return vf_fail(ctx, VOFOD_E_STATE, "no scan in flight in slot %d", slot);
  fl.active = false;
  cudaStream_t st = ctx->stream;
  const ScanPlan& plan = fl.plan;
  const vofod_params& p = plan.p;
  const vofod_schedule& s = plan.s;
  int sep_status = fl.sep_status;
  const bool applied = fl.applied, used_graph = fl.used_graph, sep_ran = fl.sep_ran;
  CK(cudaEventSynchronize(ctx->ev_done[slot]));
  if (!fl.pipelined)
  {
    memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms));
    if (!used_graph)
      for (int i = 0; i < 11; i++)
        cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    cudaEventElapsedTime(&ctx->stage_ms[11], ctx->ev[0], ctx->ev[11]);
  }

  const unsigned long long* hp = slot_counters(ctx, slot);
  if (hp[CNT_WATCHDOG])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", hp[CNT_WATCHDOG]);
  if (hp[CNT_OOB])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "%llu traversals fell outside the accumulator window", hp[CNT_OOB]);
  if (hp[CNT_VG_OVERFLOW])
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "leaf size too small for the input: integer indices would overflow");
  int raycast_status = plan.raycast_status;
  if (applied)
  {
    if (hp[CNT_APPLY_ANY])
      ctx->flags_full_dirty = false;
    else
      raycast_status = VOFOD_W_EMPTY_RAYCAST;
  }
  bool sure_flag = hp[CNT_STATE_SURE] != 0;
  if (sep_ran)
  {
    const size_t K = (size_t)hp[CNT_SEP_K];
    if (plan.sep_cap > 0)
    {
      if (K > plan.sep_cap && plan.sep_first)
      {
        // the deferred pass ran with a list that was too short and left the map untouched — but this scan has been applied on top of
        // that already, so the exact redo the in-line pass gets is no longer possible: grow the list and tell the caller
        ctx->sep_cap = K * 4 + (size_t(1) << 20);
        return vf_fail(ctx, VOFOD_E_CAPACITY, "the deferred separated-background pass found %zu background voxels, its list held %zu: the pass was skipped "
                                              "(the map now differs from schedule S1); the list has been grown", K, plan.sep_cap);
      }
      if (K > plan.sep_cap && fl.pipelined)
      {
        ctx->sep_cap = K * 4 + (size_t(1) << 20);
        return vf_fail(ctx, VOFOD_E_CAPACITY, "pipelined scan: the separated-background pass found %zu background voxels, its list held %zu, and the next scan is "
                                              "in flight already: the pass was skipped (the map now differs from schedule S1); the list has been grown", K, plan.sep_cap);
      }
      if (K > plan.sep_cap)
      {
        // the capped list overflowed: the pass did not touch the map; redo it exactly, and grow the list for the next scans
        RET(vf_begin_call(ctx));
        sep_status = vf_sepclusters_dev(ctx, s.sep_its_diff, p, 0);
        if (sep_status < 0)
          return sep_status;
        unsigned long long sure = 0;
        CK(cudaMemcpyAsync(&sure, vf_cnt(ctx, CNT_STATE_SURE), 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        sure_flag = sure != 0;
      } else if (K == 0)
        sep_status = VOFOD_W_EMPTY;
      if (hp[CNT_SEP_NUNIQ])
        return vf_fail(ctx, VOFOD_E_OVERFLOW, "sepclusters: voxel-grid index overflow");
    }
    // keep the list far larger than what the map currently holds: a change of capacity re-allocates and re-captures the graph
    // (two kernel-by-kernel scans + one capture), which must stay a rare event while the map is being explored
    if (ctx->sep_cap_forced)
      ctx->sep_cap = ctx->sep_cap_forced;  // test switch: keeps the overflow-and-redo path in use
    else if (K * 5 / 4 + 1024 > ctx->sep_cap)
      ctx->sep_cap = K * 4 + (size_t(1) << 20);
    // the clustering hash table is sized (and memset) for the points expected, not for the list capacity
    // (a growth re-allocates the table and re-captures the graph: a multi-millisecond hiccup, so start roomy and grow 4x)
    if (K * 3 / 2 > ctx->sep_table_hint)
      ctx->sep_table_hint = K * 4 + (size_t(1) << 18);
  }
  // an empty background cloud (:1155) is reported the same way whichever list the pass ran with (a pipelined scan launched before the
  // previous one's counts were read may still use the exact list while a scan-by-scan caller has moved on to the capped one)
  if (sep_ran && sep_status == VOFOD_OK && hp[CNT_SEP_K] == 0ull)
    sep_status = VOFOD_W_EMPTY;
  ctx->background_pts_sufficient = hp[CNT_STATE_BG] != 0;
  ctx->sure_background_sufficient = sure_flag;
  ctx->last_detection_id = (uint32_t)hp[CNT_DET_ID];
  ctx->last_m = (size_t)hp[CNT_VG_M];
  const size_t n_det = s.do_classify ? (size_t)hp[CNT_NDET] : 0;
  ctx->last_far = s.do_classify ? (size_t)hp[CNT_NFARPTS] : 0;
  if (res)
  {
    res->n_traversals = plan.raycast_on ? hp[CNT_TRAVERSALS] : 0;
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
  if (n_det && dets && fl.pipelined)
  {
    size_t k = n_det < det_cap ? n_det : det_cap;
    if (k > PIPE_DETS)
      k = PIPE_DETS;
    memcpy(dets, slot_dets(ctx, slot), k * sizeof(vofod_detection));
    if (n_det > PIPE_DETS && det_cap > PIPE_DETS)
      return vf_fail(ctx, VOFOD_E_CAPACITY, "pipelined scan: %zu detections, only the first %d are kept (use vofod_process_scan for more)", n_det, PIPE_DETS);
  } else if (n_det && dets)
  {
    const size_t k = n_det < det_cap ? n_det : det_cap;
    CK(cudaMemcpyAsync(dets, ctx->dets.p, k * sizeof(vofod_detection), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (n_det > det_cap && dets)
    return vf_fail(ctx, VOFOD_E_CAPACITY, "process_scan: %zu detections, capacity %zu", n_det, det_cap);
  return VOFOD_OK;
}

static int process_scan_dev(vofod_ctx* ctx, const vofod_pt* d_scan, size_t n, const vofod_pose& tf, const vofod_params& p, const vofod_schedule& s, vofod_scan_result* res,
                            vofod_detection* dets, size_t det_cap)
{
  if (res)
    memset(res, 0, sizeof(*res));
  RET(scan_launch(ctx, d_scan, n, tf, p, s, 0, false));
  return scan_finish(ctx, 0, res, dets, det_cap);
}

extern "C" {

/* instrumentation: 0 graph replays, 1 captures, 2 failed captures, 3 kernel-by-kernel scans, 4 reason of the last failed capture */
uint64_t vofod_get_stat(const vofod_ctx* ctx, int which)
{
  if (!ctx)
    return 0;
  switch (which)
  {
    case 0: return ctx->stat_replays;
    case 1: return ctx->stat_captures;
    case 2: return ctx->stat_capture_failures;
    case 3: return ctx->stat_eager;
    case 4: return (uint64_t)ctx->stat_last_capture_error;
    case 5: return ctx->stat_prefetch_hits;
    default: return 0;
  }
}

int vofod_process_scan(vofod_ctx* ctx, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                       vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  if (!scan || !tf || !p || !s)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  // an announcement that was not consumed by the two scan calls after it is void: the caller has moved on, and the host buffer may since
  // have been refilled with another scan
  ctx->scan_calls++;
  for (int i = 0; i < 2; i++)
    if (ctx->prefetched_host[i] && ctx->scan_calls - ctx->prefetched_call[i] > 2)
      ctx->prefetched_host[i] = nullptr;
  for (int i = 0; i < 2; i++)
    if (ctx->prefetched_host[i] == (const void*)scan && ctx->prefetched_n[i] == n && ctx->prefetch_buf[i].p)
    {
      // this scan was announced with vofod_prefetch_scan: its copy has been running next to the previous scan's kernels
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_prefetch[i], 0));
      ctx->prefetched_host[i] = nullptr;
      ctx->stat_prefetch_hits++;
      return process_scan_dev(ctx, ctx->prefetch_buf[i].as<vofod_pt>(), n, *tf, *p, *s, res, dets, det_cap);
    }
  ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
  CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  return process_scan_dev(ctx, ctx->scan_staging.as<vofod_pt>(), n, *tf, *p, *s, res, dets, det_cap);
}

/* Start the host->device copy of a COMING scan on a copy stream and return at once; the vofod_process_scan call that is later
 * given the same host pointer consumes it instead of copying.  Two scans can be announced ahead (typical use: announce scan
 * k+1, then process scan k).  The host buffer must stay untouched until it is consumed (pinned memory for a truly asynchronous
 * copy). */
int vofod_prefetch_scan(vofod_ctx* ctx, const vofod_pt* scan, size_t n)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!scan || n == 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_prefetch_scan: bad arguments");
  // the buffer that was announced longest ago; an unconsumed record in it is dropped (its scan will be copied the normal way)
  const int i = ctx->prefetch_next;
  ctx->prefetch_next ^= 1;
  if (ctx->prefetch_buf[i].cap < n * sizeof(vofod_pt) + 64)
  {
    ENSURE(ctx->prefetch_buf[i], n * sizeof(vofod_pt) + 64);
    CK(cudaStreamSynchronize(ctx->stream));  // the allocation's zero-fill ran on the main stream
  }
  CK(cudaMemcpyAsync(ctx->prefetch_buf[i].p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream_copy));
  CK(cudaEventRecord(ctx->ev_prefetch[i], ctx->stream_copy));
  ctx->prefetched_host[i] = scan;
  ctx->prefetched_n[i] = n;
  ctx->prefetched_call[i] = ctx->scan_calls;
  return VOFOD_OK;
}

int vofod_process_scan_resident(vofod_ctx* ctx, int slot, const vofod_pose* tf, const vofod_params* p, const vofod_schedule* s, vofod_scan_result* res,
                                vofod_detection* dets, size_t det_cap)
{
  NEED_MAP();
  if (!tf || !p || !s || slot < 0 || slot >= VOFOD_SCAN_SLOTS || !ctx->scan_slot_n[slot])
    return vf_fail(ctx, VOFOD_E_INVALID, "bad argument / empty scan slot");
  return process_scan_dev(ctx, ctx->scan_slot[slot].as<vofod_pt>(), ctx->scan_slot_n[slot], *tf, *p, *s, res, dets, det_cap);
}

}  // extern "C"
// a deferred separated-background pass, carried out in line (exact list, one host round trip inside)
int vf_flush_pending(vofod_ctx* ctx)
{
  if (!ctx->sep_pending)
    return 0;
  ctx->sep_pending = false;
  RET(vf_begin_call(ctx));
  const int rc = vf_sepclusters_dev(ctx, ctx->sep_pending_its, ctx->sep_pending_p, 0);
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
  return 0;
}
extern "C" {
int vofod_flush(vofod_ctx* ctx)
{
  NEED_MAP();
  RET(vf_flush_pending(ctx));
  return VOFOD_OK;
}

/* A sequence of scans from host buffers, back to back (rosbag replay, benchmarks): scan k + 1's host->device copy is announced before scan
 * k is processed, so it overlaps scan k's kernels, and the host spends a few microseconds between two scans instead of a round trip
 * through the caller's language binding.  results[k] / the detections of scan k (dets + k * det_cap, n_dets[k]) are those of
 * vofod_process_scan on the same inputs.  Stops at the first error and returns it; *n_done = scans completed. */
int vofod_process_scan_batch(vofod_ctx* ctx, const vofod_pt* const* scans, size_t n_scans, size_t n, const vofod_pose* poses, const vofod_params* p,
                             const vofod_schedule* scheds, vofod_scan_result* results, vofod_detection* dets, size_t det_cap, uint32_t* n_dets, size_t* n_done)
{
  NEED_MAP();
  if (!scans || !poses || !p || !scheds || !results)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (n_done)
    *n_done = 0;
  if (n_scans == 0)
    return VOFOD_OK;
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  // Two scans in flight: while scan k runs, the host launches scan k + 1 behind it (nothing it decides needs scan k's results), reads the
  // results of scan k - 1 and starts the copy of scan k + 1 — the GPU never waits for the host between two scans.  What a scan leaves
  // for the host to adapt (list capacities, the full-grid flag clear of the very first scans) takes effect one scan later.
  const size_t bytes = n * sizeof(vofod_pt);
  for (int i = 0; i < 2; i++)
  {
    if (ctx->prefetch_buf[i].cap < bytes + 64)
      ENSURE(ctx->prefetch_buf[i], bytes + 64);
    ctx->prefetched_host[i] = nullptr;  // records of vofod_prefetch_scan do not survive a batch
  }
  CK(cudaStreamSynchronize(ctx->stream));  // (fresh allocations are zero-filled on the main stream)
  auto copy_in = [&](size_t k) -> int {
    const int slot = (int)(k & 1);
    CK(cudaMemcpyAsync(ctx->prefetch_buf[slot].p, scans[k], bytes, cudaMemcpyHostToDevice, ctx->stream_copy));
    CK(cudaEventRecord(ctx->ev_prefetch[slot], ctx->stream_copy));
    return 0;
  };
  auto finish = [&](size_t k) -> int {
    const int rc = scan_finish(ctx, (int)(k & 1), results + k, dets ? dets + k * det_cap : nullptr, det_cap);
    if (rc < 0)
      return rc;
    if (n_dets)
      n_dets[k] = results[k].n_detections;
    if (n_done)
      *n_done = k + 1;
    return 0;
  };
  RET(copy_in(0));
  for (size_t k = 0; k < n_scans; k++)
  {
    const int slot = (int)(k & 1);
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_prefetch[slot], 0));
    const int lrc = scan_launch(ctx, ctx->prefetch_buf[slot].as<vofod_pt>(), n, poses[k], *p, scheds[k], slot, true);
    if (lrc < 0)
    {
      if (k >= 1)
        finish(k - 1);
      return lrc;
    }
    if (k >= 1)
    {
      const int frc = finish(k - 1);
      if (frc < 0)
      {
        scan_finish(ctx, slot, nullptr, nullptr, 0);  // drain the scan that is in flight
        return frc;
      }
    }
    if (k + 1 < n_scans)
      RET(copy_in(k + 1));  // into the buffer scan k - 1 has just finished with; overlaps scan k's kernels
  }
  RET(finish(n_scans - 1));
  return VOFOD_OK;
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

