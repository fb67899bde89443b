// clusterCloud (vofod_nodelet.cpp:689-698) = pcl::EuclideanClusterExtraction on the GPU.
//
// The reference builds a kd-tree and grows clusters by BFS; the result is the set of connected components of
// the graph "a ~ b  iff  fp32 ((ax-bx)^2 + (ay-by)^2) + (az-bz)^2 < float(tol^2)" (strict).  Here:
//   K4  spatial hash: cell pitch slightly above tol, open-addressing table keyed by the packed cell; the points of a
//       cell are stored contiguously (count -> reserve -> fill)
//   K5  one warp per point (one lane per neighbouring cell) scans the 27 cells and unites the point with every lower-indexed point
//       inside the radius (lock-free union-find with randomised linking)
//   K5b roots + atomicMin of the point index per root; K5c labels[i] = min index of i's component (the canonical label),
//       per-label sizes, number of components
// Linked-list order is scheduling dependent; the partition and the labels are not.
// That is the GENERIC path (arbitrary points: staged vofod_cluster, sepclusters outside its fast path).  The two clusterings
// a scan runs by default work on GRID CELLS and use run-based connected components on occupancy words instead (further
// down: k_cg_union for sepclusters' 26-connectivity, k_runs_union for the voxel list with the reference's tolerance).
#include <math.h>

#include "common.cuh"
#include "prims.cuh"

#define CL_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long cl_pack(const long long cx, const long long cy, const long long cz)
{
  return ((unsigned long long)(cx + (1 << 20)) & 0x1FFFFFull) | (((unsigned long long)(cy + (1 << 20)) & 0x1FFFFFull) << 21) |
         (((unsigned long long)(cz + (1 << 20)) & 0x1FFFFFull) << 42);
}
__device__ __forceinline__ unsigned cl_hash(unsigned long long k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}

// K4a: hash every point's cell; count the points per cell; remember the slot and the arrival rank of the point
__global__ void __launch_bounds__(256) k_cl_insert(const float* __restrict__ xyz, const int stride, const unsigned long long* __restrict__ d_m, const size_t m_cap,
                                                   const double inv_cell, float4* __restrict__ pts, unsigned long long* __restrict__ tkey, int* __restrict__ tcount,
                                                   const unsigned tmask, int* __restrict__ slot_of, int* __restrict__ rank_of, int* __restrict__ parent,
                                                   int* __restrict__ sizes, int* __restrict__ minidx, unsigned long long* __restrict__ watchdog)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const float x = xyz[i * stride], y = xyz[i * stride + 1], z = xyz[i * stride + 2];
    pts[i] = make_float4(x, y, z, 0.f);
    parent[i] = (int)i;
    sizes[i] = 0;
    minidx[i] = 0x7fffffff;
    slot_of[i] = -1;
    if (inv_cell == 0.0)
      continue;
    const unsigned long long key = cl_pack((long long)floor((double)x * inv_cell), (long long)floor((double)y * inv_cell), (long long)floor((double)z * inv_cell));
    unsigned slot = cl_hash(key) & tmask;
    unsigned probes = 0;
    while (true)
    {
      const unsigned long long prev = atomicCAS(tkey + slot, CL_EMPTY, key);
      if (prev == CL_EMPTY || prev == key)
      {
        slot_of[i] = (int)slot;
        rank_of[i] = atomicAdd(tcount + slot, 1);
        break;
      }
      slot = (slot + 1) & tmask;
      if (++probes > tmask)
      {
        atomicAdd(watchdog, 1ull);
        break;
      }
    }
  }
}
// K4b: the first point of every cell reserves the cell's contiguous range (any order will do)
__global__ void __launch_bounds__(256) k_cl_alloc(const unsigned long long* __restrict__ d_m, const size_t m_cap, const int* __restrict__ slot_of,
                                                  const int* __restrict__ rank_of, const int* __restrict__ tcount, int* __restrict__ tstart,
                                                  unsigned long long* __restrict__ cursor)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < m; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const int s = i < m ? slot_of[i] : -1;
    const bool first = s >= 0 && rank_of[i] == 0;
    const uint32_t need = first ? (uint32_t)tcount[s] : 0u;
    // one atomic per warp on the shared cursor
    const uint32_t incl = prims::warp_incl_scan(need);
    const uint32_t tot = __shfl_sync(VOFOD_FULL, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && tot)
      base = atomicAdd(cursor, (unsigned long long)tot);
    base = __shfl_sync(VOFOD_FULL, base, 31);
    if (first)
      tstart[s] = (int)(base + incl - need);
  }
}
// K4c: cell-contiguous copy of the points: (x, y, z, index)
__global__ void __launch_bounds__(256) k_cl_fill(const unsigned long long* __restrict__ d_m, const size_t m_cap, const float4* __restrict__ pts,
                                                 const int* __restrict__ slot_of, const int* __restrict__ rank_of, const int* __restrict__ tstart,
                                                 float4* __restrict__ cellpts)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
  {
    const int s = slot_of[i];
    if (s < 0)
      continue;
    float4 p = pts[i];
    p.w = __int_as_float((int)i);
    cellpts[tstart[s] + rank_of[i]] = p;
  }
}

// Union-find on plain (L1-cacheable) loads.  A load may return an OLD value of parent[x]; every old value is x itself or an
// ancestor of x, so walking it still climbs the tree, and the only decision that needs the truth — "a is a root, hang it
// under b" — is taken by the atomicCAS, whose return value (always fresh) is where the walk continues when it fails.
// Two same-address hot spots had to go (both measured): volatile loads all funnel into the L2 sector that holds the root
// of the ground cluster (180 us), and path compression inside find makes thousands of threads rewrite the same few
// near-root entries.  So find is READ-ONLY here; trees stay shallow because (a) links go by a pseudo-random priority
// (expected depth O(log n) — linking by index degenerates into one chain per grid row, the points arrive sorted by voxel
// key) and (b) every point compresses only ITS OWN entry once, after its unions.
__device__ __forceinline__ int uf_find(const int* __restrict__ parent, int a)
{
  while (true)
  {
    const int p = parent[a];
    if (p == a)
      return a;
    a = p;
  }
}
__device__ __forceinline__ unsigned uf_prio(const int a) { return (unsigned)a * 2654435761u; }  // odd multiplier: a bijection on u32
// unite the trees of roots-to-be a and b; returns the root both end up under (as far as this thread can tell)
__device__ __forceinline__ int uf_link(int* __restrict__ parent, int a, int b)
{
  while (true)
  {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b)
      return a;
    if (uf_prio(a) < uf_prio(b))
    {
      const int t = a;
      a = b;
      b = t;
    }
    // prio(a) > prio(b): hang a under b if a (still) is a root.  Parents always have the smaller priority => acyclic,
    // whether or not b is still a root.
    const int old = atomicCAS(parent + a, a, b);
    if (old == a)
      return b;
    a = old;  // a had been linked meanwhile: continue from its true parent
  }
}

// K5: one WARP per point, lane l < 27 owns neighbour cell l: one probe, then a walk over that cell's CONTIGUOUS point
// range (independent 16-byte loads; a per-cell linked list makes the same walk a chain of dependent L2 round trips,
// measured 68 us per call instead of ~20).
__global__ void __launch_bounds__(256) k_cl_union(const unsigned long long* __restrict__ d_m, const size_t m_cap, const double inv_cell, const float r2,
                                                  const float4* __restrict__ pts, const unsigned long long* __restrict__ tkey, const int* __restrict__ tcount,
                                                  const int* __restrict__ tstart, const unsigned tmask, const float4* __restrict__ cellpts, int* __restrict__ parent)
{
  pdl_enter();
  if (inv_cell == 0.0)
    return;
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  // every unordered pair of points is examined ONCE: lane 0 owns the point's own cell (pairs with a lower index only),
  // lanes 1..13 the 13 "forward" neighbour cells (dz,dy,dx) > (0,0,0) in lexicographic order (all their points)
  const int c = (int)lane + 13;  // 13 = the own cell in the 3x3x3 numbering
  const int dz = c / 9 - 1, dy = (c / 3) % 3 - 1, dx = c % 3 - 1;
  for (size_t i = warp0; i < m; i += n_warps)
  {
    const float4 a = pts[i];
    int first = 0, count = 0;
    if (lane < 14)
    {
      const long long cx = (long long)floor((double)a.x * inv_cell), cy = (long long)floor((double)a.y * inv_cell), cz = (long long)floor((double)a.z * inv_cell);
      const unsigned long long key = cl_pack(cx + dx, cy + dy, cz + dz);
      unsigned slot = cl_hash(key) & tmask;
      for (unsigned probes = 0; probes <= tmask; probes++)
      {
        const unsigned long long k = tkey[slot];
        if (k == key)
        {
          first = tstart[slot];
          count = tcount[slot];
          break;
        }
        if (k == CL_EMPTY)
          break;
        slot = (slot + 1) & tmask;
      }
    }
    // spread the candidates of all 14 cells evenly over the 32 lanes (the fullest cell would otherwise set the pace)
    const int incl = (int)prims::warp_incl_scan((uint32_t)count);
    const int excl = incl - count;
    const int total = __shfl_sync(VOFOD_FULL, incl, 31);
    int ri = (int)i;  // root of i as far as this lane knows
    for (int t0 = 0; t0 < total; t0 += 32)
    {
      const int t = t0 + (int)lane;
      // owner cell of candidate t: the largest lane l < 14 with excl_l <= t
      int l = 0;
#pragma unroll
      for (int step = 8; step >= 1; step >>= 1)
      {
        const int cand = l + step;
        const int e = __shfl_sync(VOFOD_FULL, excl, cand & 31);
        if (cand < 14 && e <= t)
          l = cand;
      }
      const int f = __shfl_sync(VOFOD_FULL, first, l);
      const int e0 = __shfl_sync(VOFOD_FULL, excl, l);
      if (t < total)
      {
        const float4 b = cellpts[f + (t - e0)];
        const int j = __float_as_int(b.w);
        if (l != 0 || j < (int)i)
        {
          // FLANN L2_Simple: result += diff*diff over x, y, z (fp32, separately rounded)
          float d2 = 0.0f;
          float diff = a.x - b.x;
          d2 += diff * diff;
          diff = a.y - b.y;
          d2 += diff * diff;
          diff = a.z - b.z;
          d2 += diff * diff;
          if (d2 < r2)
          {
            // the usual case: j already hangs directly under the root this lane knows for i (every finished point
            // compresses its own entry) — one load instead of two find chains
            const int pj = parent[j];
            if (pj != ri)
            {
              const int rj = uf_find(parent, pj);
              ri = uf_find(parent, ri);
              if (rj != ri)
                ri = uf_link(parent, ri, rj);
            }
          }
        }
      }
    }
    // compress this point's own entry (nobody else writes it once it is not a root)
    __syncwarp();
    if (lane == 0)
    {
      const int r = uf_find(parent, (int)i);
      if (r != (int)i && parent[i] != r)
        parent[i] = r;
    }
  }
}

// K5b: root of every point + minimum point index per root.  Neighbouring indices mostly share their root (the ground
// cluster): lanes with the same root are merged so that the hot root sees one atomic per warp, not 32.
__global__ void __launch_bounds__(256) k_cl_roots(const unsigned long long* __restrict__ d_m, const size_t m_cap, int* __restrict__ parent, int* __restrict__ root,
                                                  int* __restrict__ minidx)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < m; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool valid = i < m;
    int r = valid ? (int)i : -1 - (int)lane;
    if (valid)
      while (true)
      {
        const int p = parent[r];
        if (p == r)
          break;
        r = p;
      }
    if (valid)
      root[i] = r;
    const unsigned grp = __match_any_sync(VOFOD_FULL, r);
    const int mn = __reduce_min_sync(grp, valid ? (int)i : 0x7fffffff);
    if (valid && lane == (unsigned)(__ffs(grp) - 1))
      atomicMin(minidx + r, mn);
  }
}
// K5c: canonical label = minimum point index of the component; per-label sizes; number of components
__global__ void __launch_bounds__(256) k_cl_flatten(const unsigned long long* __restrict__ d_m, const size_t m_cap, const int* __restrict__ root,
                                                    const int* __restrict__ minidx, int* __restrict__ labels, int* __restrict__ sizes,
                                                    unsigned long long* __restrict__ d_ncl)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  unsigned roots = 0;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < m; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const bool valid = i < m;
    const int l = valid ? minidx[root[i]] : -1 - (int)lane;
    if (valid)
      labels[i] = l;
    const unsigned grp = __match_any_sync(VOFOD_FULL, l);
    if (valid && lane == (unsigned)(__ffs(grp) - 1))
      atomicAdd(sizes + l, __popc(grp));
    roots += (valid && l == (int)i);
  }
  roots = prims::warp_sum(roots);
  if (lane == 0 && roots)
    atomicAdd(d_ncl, (unsigned long long)roots);
}

static size_t cl_table_size(const size_t m_cap, const size_t table_points_hint)
{
  // open addressing degrades gracefully: a table sized for the points EXPECTED (hint) still works up to its slot count
  const size_t want = table_points_hint && table_points_hint < m_cap ? table_points_hint : m_cap;
  size_t tsize = 1024;
  while (tsize < 2 * want)
    tsize <<= 1;
  return tsize;
}
// the table clears of the NEXT vf_cluster_dev(ws, m_cap, hint), issued ahead of time (on the scan's side branch)
int vf_cluster_prefill(vofod_ctx* ctx, ClusterWs& ws, size_t m_cap, size_t table_points_hint)
{
  if (m_cap == 0)
    return 0;
  const size_t tsize = cl_table_size(m_cap, table_points_hint);
  ENSURE(ws.table_key, tsize * 8);
  ENSURE(ws.table_head, tsize * 4 * 2);
  const FillJob fj[2] = {{ws.table_key.as<uint32_t>(), tsize * 2, 0xFFFFFFFFu}, {ws.table_head.as<uint32_t>(), tsize, 0u}};  // keys = CL_EMPTY, counts = 0
  RET(vf_fill(ctx, fj, 2));
  ws.prefilled_tsize = tsize;
  return 0;
}

int vf_cluster_dev(vofod_ctx* ctx, ClusterWs& ws, const float* d_xyz, int stride_floats, const unsigned long long* d_m, size_t m_cap, float tol, int* d_labels,
                   unsigned long long* d_ncl, size_t table_points_hint)
{
  const bool second_use = &ws == &ctx->cl_bg;  // the allocator cursor was already used by the scan's own clustering
  if (!ctx->scan_prezero)
    CK(cudaMemsetAsync(d_ncl, 0, 8, ctx->stream));
  if (m_cap == 0)
    return 0;
  const size_t tsize = cl_table_size(m_cap, table_points_hint);
  ENSURE(ws.pts, m_cap * 16);
  ENSURE(ws.table_key, tsize * 8);
  ENSURE(ws.table_head, tsize * 4 * 2);  // per slot: point count, start of the cell's range
  ENSURE(ws.next, m_cap * 4 * 2);        // per point: slot, arrival rank inside the cell
  ENSURE(ws.cellpts, m_cap * 16);
  ENSURE(ws.parent, m_cap * 4);
  ENSURE(ws.sizes, m_cap * 4);
  ENSURE(ws.root, m_cap * 4);
  ENSURE(ws.minidx, m_cap * 4);
  if (ws.prefilled_tsize != tsize)
    RET(vf_cluster_prefill(ctx, ws, m_cap, table_points_hint));
  ws.prefilled_tsize = 0;
  if (!ctx->scan_prezero || second_use)
    CK(cudaMemsetAsync(vf_cnt(ctx, CNT_CL_CURSOR), 0, 8, ctx->stream));
  int* tcount = ws.table_head.as<int>();
  int* tstart = tcount + tsize;
  int* slot_of = ws.next.as<int>();
  int* rank_of = slot_of + m_cap;
  // pcl: r^2 = tolerance * tolerance evaluated in double, narrowed to the float the kd-tree compares with
  const float r2 = (float)((double)tol * (double)tol);
  // cell pitch a hair above tol: two points closer than tol in every axis always land in adjacent cells
  const double inv_cell = tol > 0.0f ? 1.0 / ((double)tol * (1.0 + 1e-6)) : 0.0;
  const int nb = vf_blocks(ctx, m_cap, 256, 8);
  LAUNCH(k_cl_insert, nb, 256, 0, d_xyz, stride_floats, d_m, m_cap, inv_cell, ws.pts.as<float4>(), ws.table_key.as<unsigned long long>(), tcount, (unsigned)(tsize - 1),
         slot_of, rank_of, ws.parent.as<int>(), ws.sizes.as<int>(), ws.minidx.as<int>(), vf_cnt(ctx, CNT_WATCHDOG));
  LAUNCH(k_cl_alloc, nb, 256, 0, d_m, m_cap, slot_of, rank_of, tcount, tstart, vf_cnt(ctx, CNT_CL_CURSOR));
  LAUNCH(k_cl_fill, nb, 256, 0, d_m, m_cap, ws.pts.as<float4>(), slot_of, rank_of, tstart, ws.cellpts.as<float4>());
  LAUNCH(k_cl_union, vf_blocks(ctx, m_cap * 32, 256, 8), 256, 0, d_m, m_cap, inv_cell, r2, ws.pts.as<float4>(), ws.table_key.as<unsigned long long>(), tcount, tstart,
         (unsigned)(tsize - 1), ws.cellpts.as<float4>(), ws.parent.as<int>());
  LAUNCH(k_cl_roots, vf_blocks(ctx, m_cap, 256, 8), 256, 0, d_m, m_cap, ws.parent.as<int>(), ws.root.as<int>(), ws.minidx.as<int>());
  LAUNCH(k_cl_flatten, vf_blocks(ctx, m_cap, 256, 8), 256, 0, d_m, m_cap, ws.root.as<int>(), ws.minidx.as<int>(), d_labels, ws.sizes.as<int>(), d_ncl);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Grid-aligned special case (updateSeparatedBGClusters with leaf size 1, sepclusters.cu): the points are the centres
// (i+0.5, j+0.5, k+0.5) of DISTINCT voxels and the tolerance is 2, so the fp32 squared distance is the exact integer
// dx^2+dy^2+dz^2 and "d2 < 4" is 26-connectivity.  No spatial hash and no distance tests: the occupancy is a bit mask per
// (z, y, 32-wide x segment) — `segbits` — and the number of a point is segoff[mask] + popcount(mask bits below it).
// The emission pass has already hung every voxel under the first voxel of its RUN (consecutive set bits of one mask), so
// only run heads work here: a head unites its run with
//   * the run that continues it in the next segment of the same row, and
//   * every run of the 4 "forward" rows (y+1,z), (y-1,z+1), (y,z+1), (y+1,z+1) that overlaps [x0-1, x1+1]
// — one union per pair of touching runs instead of one per pair of touching voxels (a ground plane has ~10x fewer).
// parent[] / sizes[] / minidx[] come initialised from the emission pass.
// ---------------------------------------------------------------------------------------------------------------
// FOUR threads per point, one per forward row: the work of a head is a chain of dependent loads (own mask -> row masks ->
// rank offsets -> parents), and four short chains side by side finish sooner than one long one.
__global__ void __launch_bounds__(256) k_cg_union(const unsigned long long* __restrict__ d_m, const size_t m_cap, const vofod_vox* __restrict__ ds, const Geom g,
                                                  const uint32_t* __restrict__ segbits, const uint32_t* __restrict__ segoff, int* __restrict__ parent)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const int sx = g.st_size[0], sy = g.st_size[1], sz = g.st_size[2];
  const int nseg = (sx + 31) / 32;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < 4 * m; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = t >> 2;
    const int q = (int)(t & 3);
    const vofod_vox v = ds[i];
    const int x = (int)v.x - g.st_lo[0], y = (int)v.y - g.st_lo[1], z = (int)v.z - g.st_lo[2];
    const int seg = x >> 5, b = x & 31;
    const size_t row = ((size_t)z * sy + y) * nseg;
    const uint32_t own = segbits[row + seg];
    if (b > 0 && ((own >> (b - 1)) & 1u))
      continue;  // not the head of its run
    const uint32_t up = ~(own >> b);              // first zero at or above b ends the run (bits shifted in from the top are zero)
    const int len = __ffs(up) ? __ffs(up) - 1 : 32;  // b == 0 and a full mask: 32
    const int x1 = b + len - 1;                    // last bit of the run
    int ri = (int)i;
    auto unite = [&](const uint32_t j) {
      if ((size_t)j >= m)
        return;  // list overflow: the pass is void and redone with a larger list; only memory safety matters here
      const int pj = parent[j];
      if (pj != ri)
      {
        const int rj = uf_find(parent, pj);
        ri = uf_find(parent, ri);
        if (rj != ri)
          ri = uf_link(parent, ri, rj);
      }
    };
    if (q == 0 && x1 == 31 && seg + 1 < nseg && (segbits[row + seg + 1] & 1u))
      unite(segoff[row + seg + 1]);
    const int dy = q == 1 ? -1 : (q == 2 ? 0 : 1), dz = q == 0 ? 0 : 1;  // forward rows (y+1,z), (y-1,z+1), (y,z+1), (y+1,z+1)
    const int ny = y + dy, nz = z + dz;
    if (ny < 0 || ny >= sy || nz >= sz)
      continue;
    const size_t nrow = ((size_t)nz * sy + ny) * nseg;
    const uint32_t wc = segbits[nrow + seg];
    const uint32_t wl = (b == 0 && seg > 0) ? segbits[nrow + seg - 1] : 0u;
    const uint32_t wr = (x1 == 31 && seg + 1 < nseg) ? segbits[nrow + seg + 1] : 0u;
    // 34-bit window, position p <-> x = 32*seg - 1 + p; of interest: x in [x0-1, x1+1] <-> p in [b, x1+2]
    unsigned long long w = (unsigned long long)(wl >> 31) | ((unsigned long long)wc << 1) | ((unsigned long long)(wr & 1u) << 33);
    w &= ((1ull << (len + 2)) - 1ull) << b;
    while (w)
    {
      const int p = __ffsll((long long)w) - 1;
      const unsigned long long inv = ~(w >> p);
      const int rl = __ffsll((long long)inv) - 1;  // >= 1; w has at most 34 bits, so inv always has a set bit
      w &= ~(((1ull << rl) - 1ull) << p);
      uint32_t j;
      if (p == 0)
        j = segoff[nrow + seg - 1] + (uint32_t)__popc(wl & 0x7fffffffu);
      else if (p == 33)
        j = segoff[nrow + seg + 1];
      else
        j = segoff[nrow + seg] + (uint32_t)__popc(wc & ((1u << (p - 1)) - 1u));
      unite(j);
    }
  }
}

// d_labels == NULL: only the unions — the caller reads the forest (ws.parent) itself
int vf_cluster_runs26_dev(vofod_ctx* ctx, ClusterWs& ws, const vofod_vox* d_ds, const uint32_t* d_segbits, const uint32_t* d_segoff, const unsigned long long* d_m,
                          size_t m_cap, int* d_labels, unsigned long long* d_ncl)
{
  if (!ctx->scan_prezero)
    CK(cudaMemsetAsync(d_ncl, 0, 8, ctx->stream));
  if (m_cap == 0)
    return 0;
  const int nb = vf_blocks(ctx, m_cap, 256, 8);
  LAUNCH(k_cg_union, vf_blocks(ctx, m_cap * 4, 256, 8), 256, 0, d_m, m_cap, d_ds, ctx->g, d_segbits, d_segoff, ws.parent.as<int>());
  if (!d_labels)
    return 0;
  ENSURE(ws.root, m_cap * 4);
  LAUNCH(k_cl_roots, nb, 256, 0, d_m, m_cap, ws.parent.as<int>(), ws.root.as<int>(), ws.minidx.as<int>());
  LAUNCH(k_cl_flatten, nb, 256, 0, d_m, m_cap, ws.root.as<int>(), ws.minidx.as<int>(), d_labels, ws.sizes.as<int>(), d_ncl);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The scan's own clustering (clusterCloud of the voxel-grid output, :932).  Its points are the CENTRES of the occupied
// leaves of a grid, (ijk + 0.5) * leaf + offset, so whether two of them are within the tolerance is a property of their
// index offset: n = dx^2+dy^2+dz^2 with n*leaf^2 clearly below tol^2 is always inside, clearly above never, and only when
// n*leaf^2 EQUALS tol^2 (n = 9 for 1.5 m / 0.5 m) the fp32 rounding of the two centres decides — for exactly those pairs
// the reference's fp32 distance is evaluated, on centres recomputed bit-identically from the indices.
// The points of a run of consecutive occupied x hang under the run's first point (seeded by the voxel-grid emission
// pass); one thread per (point, forward row) — only run heads work — unites the run with every run of that row inside the
// window the row allows.  Occupancy comes from RunWord entries tagged with the API call number (no clearing).
// ---------------------------------------------------------------------------------------------------------------
bool vf_run_rows(const float tol, const float leaf, RunRows& rr)
{
  rr.n = 0;
  if (!(tol > 0.0f) || !(leaf > 0.0f) || !(tol > leaf))
    return false;
  const double r2 = (double)(float)((double)tol * (double)tol);
  const double l2 = (double)leaf * (double)leaf;
  auto cls = [&](const long long n) {  // 0 inside, 1 border (decided in fp32), 2 outside
    const double v = (double)n * l2;
    if (fabs(v - r2) <= 1e-4 * r2)
      return 1;
    return v < r2 ? 0 : 2;
  };
  const int maxd = (int)ceil((double)tol / (double)leaf) + 1;
  for (int dz = 0; dz <= maxd; dz++)
    for (int dy = (dz == 0 ? 0 : -maxd); dy <= maxd; dy++)
    {
      const long long base = (long long)dy * dy + (long long)dz * dz;
      const bool same = dy == 0 && dz == 0;
      int R, shell;
      if (!same && cls(base) == 2)
        continue;
      if (!same && cls(base) == 1)
      {
        if (cls(base + 1) != 2)
          return false;
        R = -1;
        shell = 2;
      } else
      {
        R = 0;
        while (cls(base + (long long)(R + 1) * (R + 1)) == 0)
          R++;
        shell = cls(base + (long long)(R + 1) * (R + 1)) == 1 ? 1 : 0;
        if (cls(base + (long long)(R + 2) * (R + 2)) != 2)
          return false;
        if (same && R == 0 && shell == 0)
          return false;  // tol <= leaf: not a grid neighbourhood
      }
      if (R + 1 > 30 || rr.n >= RUN_ROWS_MAX)
        return false;
      rr.dy[rr.n] = (signed char)dy;
      rr.dz[rr.n] = (signed char)dz;
      rr.R[rr.n] = (signed char)R;
      rr.shell[rr.n] = (signed char)shell;
      rr.n++;
    }
  return rr.n > 0;
}

__global__ void __launch_bounds__(256) k_runs_union(const unsigned long long* __restrict__ d_m, const size_t m_cap, const uint32_t* __restrict__ cellkey,
                                                    const RunWord* __restrict__ words, const VgLayout* __restrict__ Lp, const RunRows rows, const float r2,
                                                    const unsigned long long* __restrict__ counters, int* __restrict__ parent)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const VgLayout L = *after_wait(Lp);
  const unsigned long long tag = *after_wait(counters + CNT_EPOCH_BASE);
  const int d0 = L.div[0], d1 = L.div[1], d2 = L.div[2];
  const uint32_t d01 = (uint32_t)d0 * (uint32_t)d1;
  const int nseg = (d0 + 31) / 32;
  const size_t total = m * (size_t)rows.n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = t / rows.n;
    const int q = (int)(t - i * rows.n);
    const uint32_t key = cellkey[i];
    const int k2 = (int)(key / d01), rem = (int)(key - (uint32_t)k2 * d01), k1 = rem / d0, k0 = rem - k1 * d0;
    const int seg = k0 >> 5, b = k0 & 31;
    const uint32_t own = words[((size_t)k2 * d1 + k1) * nseg + seg].bits;
    if (b > 0 && ((own >> (b - 1)) & 1u))
      continue;  // not the head of its run
    const uint32_t up = ~(own >> b);
    const int len = __ffs(up) ? __ffs(up) - 1 : 32;
    const int x0 = k0, x1 = k0 + len - 1;
    const int dy = rows.dy[q], dz = rows.dz[q], R = rows.R[q], shell = rows.shell[q];
    const int ny = k1 + dy, nz = k2 + dz;
    if (ny < 0 || ny >= d1 || nz >= d2)
      continue;
    const bool same = dy == 0 && dz == 0;
    // window of candidate cells in the row, and its part that is inside for sure
    int xa, xb, sa, sb;
    if (same)
    {
      xa = x1 + 1, xb = x1 + R + (shell == 1 ? 1 : 0);
      sa = x1 + 1, sb = x1 + R;
    } else if (shell == 2)
    {
      xa = x0, xb = x1;
      sa = 1, sb = 0;  // empty
    } else
    {
      xa = x0 - R - shell, xb = x1 + R + shell;
      sa = x0 - R, sb = x1 + R;
    }
    xa = max(xa, 0);
    xb = min(xb, d0 - 1);
    if (xa > xb)
      continue;
    int ri = (int)i;
    auto unite = [&](const uint32_t j) {
      if ((size_t)j >= m)
        return;
      const int pj = parent[j];
      if (pj != ri)
      {
        const int rj = uf_find(parent, pj);
        ri = uf_find(parent, ri);
        if (rj != ri)
          ri = uf_link(parent, ri, rj);
      }
    };
    const size_t nrow = ((size_t)nz * d1 + ny) * nseg;
    for (int ws = xa >> 5; ws <= (xb >> 5); ws++)
    {
      const RunWord rw = words[nrow + ws];
      if (rw.tag != tag)
        continue;  // nothing of this scan in that segment
      const int lo = max(xa - ws * 32, 0), hi = min(xb - ws * 32, 31);
      uint32_t wbits = rw.bits & (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
      while (wbits)
      {
        const int p = __ffs(wbits) - 1;
        const uint32_t inv = ~(wbits >> p);
        const int rl = __ffs(inv) ? __ffs(inv) - 1 : 32;
        wbits = rl >= 32 ? 0u : (wbits & ~(((1u << rl) - 1u) << p));
        const int p0 = ws * 32 + p, p1 = p0 + rl - 1;  // this run of the window, absolute x
        const uint32_t j = rw.rank + (uint32_t)__popc(rw.bits & ((1u << p) - 1u));
        if (p0 <= sb && p1 >= sa)
        {
          unite(j);  // some cell of it is inside for sure
          continue;
        }
        // a border cell: the reference's fp32 test (FLANN L2_Simple: diff*diff summed over x, y, z) on the two centres
        const int xh = shell == 2 ? p0 : (p0 < x0 ? x0 : x1);  // the cell of this run that faces it
        const float ax = ((float)xh + 0.5f) * L.leaf + L.offset[0], bx = ((float)p0 + 0.5f) * L.leaf + L.offset[0];
        const float ay = ((float)k1 + 0.5f) * L.leaf + L.offset[1], by = ((float)ny + 0.5f) * L.leaf + L.offset[1];
        const float az = ((float)k2 + 0.5f) * L.leaf + L.offset[2], bz = ((float)nz + 0.5f) * L.leaf + L.offset[2];
        float dd = 0.0f;
        float diff = ax - bx;
        dd += diff * diff;
        diff = ay - by;
        dd += diff * diff;
        diff = az - bz;
        dd += diff * diff;
        if (dd < r2)
          unite(j);
      }
    }
  }
}

int vf_cluster_runs_dev(vofod_ctx* ctx, ClusterWs& ws, const uint32_t* d_cellkey, const RunWord* d_words, const VgLayout* d_layout, const RunRows& rows, float tol,
                        const unsigned long long* d_m, size_t m_cap, int* d_labels, unsigned long long* d_ncl)
{
  if (!ctx->scan_prezero)
    CK(cudaMemsetAsync(d_ncl, 0, 8, ctx->stream));
  if (m_cap == 0)
    return 0;
  ENSURE(ws.root, m_cap * 4);
  const float r2 = (float)((double)tol * (double)tol);  // as vf_cluster_dev
  const int nb = vf_blocks(ctx, m_cap, 256, 8);
  LAUNCH(k_runs_union, vf_blocks(ctx, m_cap * (size_t)rows.n, 256, 8), 256, 0, d_m, m_cap, d_cellkey, d_words, d_layout, rows, r2,
         (const unsigned long long*)ctx->d_counters.as<unsigned long long>(), ws.parent.as<int>());
  LAUNCH(k_cl_roots, nb, 256, 0, d_m, m_cap, ws.parent.as<int>(), ws.root.as<int>(), ws.minidx.as<int>());
  LAUNCH(k_cl_flatten, nb, 256, 0, d_m, m_cap, ws.root.as<int>(), ws.minidx.as<int>(), d_labels, ws.sizes.as<int>(), d_ncl);
  return 0;
}

extern "C" int vofod_cluster_grid_rows(float tolerance, float leaf, int8_t* dy, int8_t* dz, int8_t* R, int8_t* shell, int cap)
{
  RunRows rr;
  if (!vf_run_rows(tolerance, leaf, rr) || rr.n > cap)
    return 0;
  for (int i = 0; i < rr.n; i++)
  {
    if (dy) dy[i] = rr.dy[i];
    if (dz) dz[i] = rr.dz[i];
    if (R) R[i] = rr.R[i];
    if (shell) shell[i] = rr.shell[i];
  }
  return rr.n;
}

extern "C" int vofod_cluster(vofod_ctx* ctx, const float* xyz, size_t m, float tol, int32_t* labels, size_t* n_clusters)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!n_clusters || (m && (!xyz || !labels)))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  *n_clusters = 0;
  if (m == 0)
    return VOFOD_OK;
  if (m > (size_t)INT32_MAX)
    return vf_fail(ctx, VOFOD_E_INVALID, "too many points");
  ENSURE(ctx->scratch_a, m * 12);
  ENSURE(ctx->labels, m * 4);
  CK(cudaMemcpyAsync(ctx->scratch_a.p, xyz, m * 12, cudaMemcpyHostToDevice, ctx->stream));
  RET(vf_cluster_dev(ctx, ctx->cl, ctx->scratch_a.as<float>(), 3, nullptr, m, tol, ctx->labels.as<int>(), vf_cnt(ctx, CNT_NCLUSTERS)));
  unsigned long long ncl = 0, wd = 0;
  CK(cudaMemcpyAsync(labels, ctx->labels.p, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&ncl, vf_cnt(ctx, CNT_NCLUSTERS), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&wd, vf_cnt(ctx, CNT_WATCHDOG), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (wd)
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", wd);
  *n_clusters = (size_t)ncl;
  return VOFOD_OK;
}
