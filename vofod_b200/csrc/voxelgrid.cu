// filterAndTransform (vofod_nodelet.cpp:621-684) + VoxelGridWeighted / VoxelGridCounted
// (src/voxel_grid_weighted.cpp:41-190, src/voxel_grid_counted.cpp:49-196) on the GPU:
//   K1a crop (exclude box, sensor frame) + rigid transform + crop (operation area) + min/max   [1 pass over the scan]
//   K1b layout: min_b / offset / div exactly as the reference computes them                     [1 thread]
//   K1c voxel key per point (invalid points get the sentinel 0xFFFFFFFF)
//   K2  hand-written LSD radix sort of the keys (prims.cuh) — keys only: the output needs neither the point
//       index nor ijk (ijk is recovered from the key), and VoxelGridCounted's counts are prefix sums over the
//       UNSORTED input (its slice quirk, voxel_grid_counted.cpp:185-187)
//   K3  run heads -> exclusive scan -> unique keys + run starts -> voxel centre + count
// The scan path (vf_filter_voxelize_dev) replaces K1c-K3 by a sort-free variant (k_vgh_count / popcount scan / k_vgh_emit,
// see there) whenever the operation area bounds the key range; K2/K3 serve arbitrary clouds (staged entry points,
// VoxelGridCounted of the sepclusters general path).
#include <math.h>

#include "common.cuh"
#include "prims.cuh"


struct CropArgs
{
  float ex_min[3], ex_max[3];
  float op_min[3], op_max[3];
  int n;
};

__global__ void k_minmax_init(MinMax* mm)
{
  pdl_enter();
  minmax_init(mm);
}

// block-wide min/max/count, then ONE set of 7 global atomics per block (a per-warp commit serialises ~8k warps on 7 words)
__device__ __forceinline__ void block_minmax_commit(const bool valid, const float x, const float y, const float z, MinMax* mm)
{
  __shared__ int s_mn[3][8], s_mx[3][8];
  __shared__ unsigned s_cnt[8];
  int mn[3] = {valid ? f2ord(x) : 0x7fffffff, valid ? f2ord(y) : 0x7fffffff, valid ? f2ord(z) : 0x7fffffff};
  int mx[3] = {valid ? f2ord(x) : (int)0x80000000, valid ? f2ord(y) : (int)0x80000000, valid ? f2ord(z) : (int)0x80000000};
  unsigned cnt = valid ? 1u : 0u;
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    mn[a] = __reduce_min_sync(VOFOD_FULL, mn[a]);
    mx[a] = __reduce_max_sync(VOFOD_FULL, mx[a]);
  }
  cnt = __reduce_add_sync(VOFOD_FULL, cnt);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
  {
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
      s_mn[a][w] = mn[a];
      s_mx[a][w] = mx[a];
    }
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 7)
  {
    const int nw = (blockDim.x + 31) >> 5;
    if (threadIdx.x < 3)
    {
      int v = s_mn[threadIdx.x][0];
      for (int i = 1; i < nw; i++)
        v = min(v, s_mn[threadIdx.x][i]);
      if (v != 0x7fffffff)
        atomicMin(&mm->mn[threadIdx.x], v);
    } else if (threadIdx.x < 6)
    {
      int v = s_mx[threadIdx.x - 3][0];
      for (int i = 1; i < nw; i++)
        v = max(v, s_mx[threadIdx.x - 3][i]);
      if (v != (int)0x80000000)
        atomicMax(&mm->mx[threadIdx.x - 3], v);
    } else
    {
      unsigned v = 0;
      for (int i = 0; i < nw; i++)
        v += s_cnt[i];
      if (v)
        atomicAdd(&mm->n_valid, v);
    }
  }
}

// K1a — pcl::CropBox(negative) -> pcl::transformPointCloud -> pcl::CropBox (vofod_nodelet.cpp:626-655)
__global__ void __launch_bounds__(256) k_crop_transform(const CropArgs a, const ScanDyn* __restrict__ dyn, float4* __restrict__ pts, MinMax* mm)
{
  pdl_enter();
  __shared__ __align__(16) uint32_t s_pts[256 * 5];
  const vofod_pt* __restrict__ scan = dyn->scan;
  const Pose33 tf = dyn->tf;
  const int blk_first = blockIdx.x * 256;
  {
    const int n_here = min(256, a.n - blk_first);
    const int n_words = n_here * 5;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(scan) + (size_t)blk_first * 5;
    const int n_vec = n_words / 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    for (int i = threadIdx.x; i < n_vec; i += 256)
      reinterpret_cast<uint4*>(s_pts)[i] = __ldg(src4 + i);
    for (int i = n_vec * 4 + threadIdx.x; i < n_words; i += 256)
      s_pts[i] = __ldg(src + i);
  }
  __syncthreads();
  const int idx = blk_first + threadIdx.x;
  bool valid = idx < a.n;
  float X = 0.f, Y = 0.f, Z = 0.f;
  if (valid)
  {
    const float x = __uint_as_float(s_pts[threadIdx.x * 5 + 0]), y = __uint_as_float(s_pts[threadIdx.x * 5 + 1]), z = __uint_as_float(s_pts[threadIdx.x * 5 + 2]);
    valid = isfinite(x) && isfinite(y) && isfinite(z);
    // keep points OUTSIDE the closed exclude box
    const bool outside1 = (x < a.ex_min[0] || y < a.ex_min[1] || z < a.ex_min[2]) || (x > a.ex_max[0] || y > a.ex_max[1] || z > a.ex_max[2]);
    valid = valid && outside1;
    // PCL 1.10 SSE transform order: x*c0 + (y*c1 + (z*c2 + c3))
    X = x * tf.R[0] + (y * tf.R[1] + (z * tf.R[2] + tf.t[0]));
    Y = x * tf.R[3] + (y * tf.R[4] + (z * tf.R[5] + tf.t[1]));
    Z = x * tf.R[6] + (y * tf.R[7] + (z * tf.R[8] + tf.t[2]));
    const bool outside2 = (X < a.op_min[0] || Y < a.op_min[1] || Z < a.op_min[2]) || (X > a.op_max[0] || Y > a.op_max[1] || Z > a.op_max[2]);
    valid = valid && isfinite(X) && isfinite(Y) && isfinite(Z) && !outside2;
  }
  // survivors are written COMPACTLY (one cursor atomic per block): the sort then handles ~45 % of the rays, and no sentinel
  // keys spoil its high digits.  Their order is scheduling dependent, which is harmless: only per-voxel COUNTS survive.
  {
    __shared__ unsigned s_wcnt[8], s_base;
    const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(VOFOD_FULL, valid);
    if (lane == 0)
      s_wcnt[wrp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0)
    {
      unsigned tot = 0;
      for (int i = 0; i < 8; i++)
      {
        const unsigned c = s_wcnt[i];
        s_wcnt[i] = tot;
        tot += c;
      }
      s_base = tot ? atomicAdd(&mm->n_append, tot) : 0u;
    }
    __syncthreads();
    if (valid)
      pts[s_base + s_wcnt[wrp] + __popc(bal & prims::lanemask_lt())] = make_float4(X, Y, Z, 1.0f);
  }
  block_minmax_commit(valid, X, Y, Z, mm);
}

// generic input: xyz triples (stride floats), w = 4th value (intensity) or 1; valid = finite
__global__ void __launch_bounds__(256) k_load_cloud(const float* __restrict__ in, const int stride, const unsigned long long* __restrict__ d_n, const int cap,
                                                   float4* __restrict__ pts, MinMax* mm)
{
  pdl_enter();
  const int n = (int)prims::dev_count(d_n, (size_t)cap);
  const int idx = blockIdx.x * 256 + threadIdx.x;
  bool valid = idx < n;
  float x = 0.f, y = 0.f, z = 0.f;
  if (valid)
  {
    x = in[(size_t)idx * stride];
    y = in[(size_t)idx * stride + 1];
    z = in[(size_t)idx * stride + 2];
    valid = isfinite(x) && isfinite(y) && isfinite(z);
    // w carries validity (weighted) — the counted variant reads intensity from the source array directly
    pts[idx] = make_float4(x, y, z, valid ? 1.0f : 0.0f);
  } else if (idx < cap)
    pts[idx] = make_float4(0.f, 0.f, 0.f, 0.0f);  // beyond the device-side count: invalid, sorts to the end
  block_minmax_commit(valid, x, y, z, mm);
}

// K1b — voxel_grid_weighted.cpp:56-111.  Returns the layout in L; n_valid == 0 gives the empty layout.
__device__ inline void vg_layout_compute(const MinMax* mm, const float leaf, const int align, const float ac0, const float ac1, const float ac2, VgLayout& L)
{
  const float inv = 1.0f / leaf;
  L.leaf = leaf;
  L.inv = inv;
  L.n_valid = mm->n_valid;
  L.overflow = 0;
  if (L.n_valid == 0)
  {
    for (int a = 0; a < 3; a++)
    {
      L.offset[a] = 0.f;
      L.min_b[a] = L.max_b[a] = 0;
      L.div[a] = 1;
    }
    return;
  }
  const float ac[3] = {ac0, ac1, ac2};
  float min_p[3], max_p[3];
  for (int a = 0; a < 3; a++)
  {
    min_p[a] = ord2f(mm->mn[a]);
    max_p[a] = ord2f(mm->mx[a]);
  }
  const long long dx = (long long)((max_p[0] - min_p[0]) * inv) + 2;
  const long long dy = (long long)((max_p[1] - min_p[1]) * inv) + 2;
  const long long dz = (long long)((max_p[2] - min_p[2]) * inv) + 2;
  if (dx * dy * dz > 2147483647ll)
    L.overflow = 1;
  for (int a = 0; a < 3; a++)
  {
    L.min_b[a] = (int)floorf(min_p[a] * inv);
    L.max_b[a] = (int)floorf(max_p[a] * inv);
    float off = (float)L.min_b[a] * leaf;
    if (align)
    {
      float aco = fmodf(ac[a] - leaf / 2, leaf);
      if (aco < 0)
        aco += leaf;
      off -= aco;
      L.min_b[a] = (int)floorf(off * inv);
    }
    L.offset[a] = off;
    L.div[a] = L.max_b[a] - L.min_b[a] + 1;
  }
}
__global__ void k_vg_layout(const MinMax* __restrict__ mm, const float leaf, const int align, const float ac0, const float ac1, const float ac2, VgLayout* __restrict__ L,
                            unsigned long long* __restrict__ counters, const int slot_nvalid, const int slot_overflow)
{
  pdl_enter();
  VgLayout l;
  vg_layout_compute(mm, leaf, align, ac0, ac1, ac2, l);
  *L = l;
  counters[slot_nvalid] = l.n_valid;
  counters[slot_overflow] = (unsigned long long)l.overflow;
}

// ---------------------------------------------------------------------------------------------------------------
// Sort-free VoxelGridWeighted for the scan path.  The filter's output is, per occupied leaf, its CENTRE and its point count,
// in ascending key order (key = i + j*div0 + k*div0*div1) — nothing depends on the order of the points inside a leaf.  The
// operation-area crop bounds the key range, so instead of sorting ~10^5 keys (histogram + 4 radix passes + run detection):
//   count: every point adds 1 to cnt[key] (dense over the key range, zero at rest) and sets bit (i & 31) of the occupancy
//          word of (k, j, i >> 5)                                  [one atomic pair per distinct key per warp]
//   scan : popcount scan of the occupancy words = rank of every occupied leaf in key order (prims.cuh)
//   emit : per set bit: centre + count -> out[rank]; the words and counts it read are put back to zero
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vgh_count(const float4* __restrict__ pts, const MinMax* mm, const float leaf, const float ac0, const float ac1,
                                                   const float ac2, VgLayout* Lout, unsigned long long* counters, const int slot_nvalid, const int slot_overflow,
                                                   const int budget_x, const int budget_y, const int budget_z, uint32_t* __restrict__ cnt,
                                                   uint32_t* __restrict__ bits)
{
  pdl_enter();
  __shared__ VgLayout sL;
  if (threadIdx.x == 0)
  {
    VgLayout l;
    vg_layout_compute(mm, leaf, 1, ac0, ac1, ac2, l);
    if (l.div[0] > budget_x || l.div[1] > budget_y || l.div[2] > budget_z)
      l.overflow = 1;  // cannot happen for points cropped to the operation area the budget was derived from
    sL = l;
    if (blockIdx.x == 0)
    {
      *Lout = l;
      counters[slot_nvalid] = l.n_valid;
      if (l.overflow)
        counters[slot_overflow] = 1ull;  // (zeroed by k_begin_call: every writer only ever raises it, in whatever order the blocks run)
      counters[CNT_VGH_WORDS] = l.overflow ? 0ull : (unsigned long long)((l.div[0] + 31) / 32) * (unsigned long long)l.div[1] * (unsigned long long)l.div[2];
    }
  }
  __syncthreads();
  const VgLayout L = sL;
  if (L.overflow)
    return;
  const unsigned lane = threadIdx.x & 31;
  const int rows = (int)L.n_valid;
  const uint32_t d0 = (uint32_t)L.div[0], d01 = (uint32_t)L.div[0] * (uint32_t)L.div[1], total = d01 * (uint32_t)L.div[2];
  const uint32_t nseg = (d0 + 31u) / 32u;
  for (int i0 = (blockIdx.x * 256 + threadIdx.x) & ~31; i0 < rows; i0 += gridDim.x * 256)
  {
    const int i = i0 + (int)lane;
    uint32_t key = 0xFFFFFF00u + lane;  // no point: a key of its own
    bool valid = i < rows;
    if (valid)
    {
      const float4 p = pts[i];
      const int ijk0 = (int)floorf((p.x - L.offset[0]) * L.inv);
      const int ijk1 = (int)floorf((p.y - L.offset[1]) * L.inv);
      const int ijk2 = (int)floorf((p.z - L.offset[2]) * L.inv);
      key = (uint32_t)(ijk0 + ijk1 * L.div[0] + ijk2 * L.div[0] * L.div[1]);
      // per axis, not just key < total: an index that leaves its axis aliases another leaf's key, and the emit kernel rebuilds ijk from the key
      if ((unsigned)ijk0 >= (unsigned)L.div[0] || (unsigned)ijk1 >= (unsigned)L.div[1] || (unsigned)ijk2 >= (unsigned)L.div[2] || key >= total)
      {
        // a point outside its own bounding box (not reachable with finite inputs): the sort path would emit it with a wrapped
        // key; here it is reported instead of written out of bounds
        counters[slot_overflow] = 1ull;
        valid = false;
        key = 0xFFFFFF00u + lane;
      }
    }
    const unsigned grp = __match_any_sync(VOFOD_FULL, key);
    if (valid && lane == (unsigned)(__ffs(grp) - 1))
    {
      const uint32_t k2 = key / d01, rem = key - k2 * d01, k1 = rem / d0, k0 = rem - k1 * d0;
      atomicAdd(cnt + key, (uint32_t)__popc(grp));
      atomicOr(bits + ((size_t)k2 * L.div[1] + k1) * nseg + (k0 >> 5), 1u << (k0 & 31u));
    }
  }
}
// one warp per NON-EMPTY occupancy word (listed by the scan together with its rank offset), lane = bit
// seeds != NULL: also the input of the grid clustering (cluster.cu, vf_cluster_runs_dev): key per point, tagged occupancy
// word, and the union-find forest with every point hung under the first point of its run
struct VghSeeds
{
  uint32_t* cellkey;
  RunWord* words;
  int *parent, *sizes, *minidx;
  const unsigned long long* counters;
};
__global__ void __launch_bounds__(256) k_vgh_emit(const VgLayout* __restrict__ Lp, const uint4* __restrict__ list, const unsigned long long* __restrict__ d_list_n,
                                                  const size_t list_cap, uint32_t* __restrict__ bits, uint32_t* __restrict__ cnt, vofod_vox* __restrict__ out,
                                                  const size_t out_cap, const VghSeeds seeds)
{
  pdl_enter();
  const VgLayout L = *after_wait(Lp);
  const size_t n_list = prims::dev_count(d_list_n, list_cap);
  const unsigned lane = threadIdx.x & 31;
  const uint32_t d0 = (uint32_t)L.div[0], d1 = (uint32_t)L.div[1], d01 = d0 * d1;
  const uint32_t nseg = (d0 + 31u) / 32u;
  for (size_t e = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5; e < n_list; e += ((size_t)gridDim.x * 256) >> 5)
  {
    const uint4 ent = list[e];  // (word index, occupancy bits, rank of its first set bit)
    const uint32_t w = ent.x, wb = ent.y;
    if (lane == 0)
    {
      bits[w] = 0u;
      if (seeds.words)
      {
        RunWord rw;
        rw.tag = *after_wait(seeds.counters + CNT_EPOCH_BASE);
        rw.bits = wb;
        rw.rank = ent.z;
        seeds.words[w] = rw;
      }
    }
    if (!((wb >> lane) & 1u))
      continue;
    const uint32_t seg = w % nseg, row = w / nseg;
    const uint32_t k0 = seg * 32u + lane, k1 = row % d1, k2 = row / d1;
    const uint32_t key = k0 + k1 * d0 + k2 * d01;
    const uint32_t c = cnt[key];
    cnt[key] = 0u;
    const size_t r = (size_t)ent.z + (size_t)__popc(wb & prims::lanemask_lt());
    if (r < out_cap)
    {
      vofod_vox o;
      o.x = ((float)(int)k0 + 0.5f) * L.leaf + L.offset[0];
      o.y = ((float)(int)k1 + 0.5f) * L.leaf + L.offset[1];
      o.z = ((float)(int)k2 + 0.5f) * L.leaf + L.offset[2];
      o.count = c;
      out[r] = o;
      if (seeds.words)
      {
        const unsigned zeros_below = ~wb & prims::lanemask_lt();
        const int head = zeros_below ? 32 - __clz(zeros_below) : 0;  // first lane of this lane's run of consecutive set bits
        seeds.cellkey[r] = key;
        seeds.parent[r] = (int)(ent.z + (uint32_t)__popc(wb & ((1u << head) - 1u)));
        seeds.sizes[r] = 0;
        seeds.minidx[r] = 0x7fffffff;
      }
    }
  }
}

// K1c — voxel_grid_weighted.cpp:122-139
// `compact`: the first L.n_valid rows of pts are the valid points (k_crop_transform) and only those get keys
__global__ void __launch_bounds__(256) k_vg_keys(const float4* __restrict__ pts, const int n, const VgLayout* __restrict__ Lp, uint32_t* __restrict__ keys, const int compact)
{
  pdl_enter();
  const VgLayout L = *Lp;
  const int rows = compact ? (int)L.n_valid : n;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < rows; i += gridDim.x * 256)
  {
    const float4 p = pts[i];
    uint32_t key = 0xFFFFFFFFu;
    if (p.w != 0.0f && !L.overflow)
    {
      const int ijk0 = (int)floorf((p.x - L.offset[0]) * L.inv);
      const int ijk1 = (int)floorf((p.y - L.offset[1]) * L.inv);
      const int ijk2 = (int)floorf((p.z - L.offset[2]) * L.inv);
      key = (uint32_t)(ijk0 + ijk1 * L.div[0] + ijk2 * L.div[0] * L.div[1]);
    }
    keys[i] = key;
  }
}

// K3a — run heads among the valid (non-sentinel) sorted keys
__global__ void __launch_bounds__(256) k_vg_heads(const uint32_t* __restrict__ keys, const int n, uint32_t* __restrict__ heads, const VgLayout* __restrict__ Lp,
                                                  const int compact)
{
  pdl_enter();
  const int rows = compact ? (int)Lp->n_valid : n;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
  {
    const uint32_t k = i < rows ? keys[i] : 0xFFFFFFFFu;
    heads[i] = (k != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
  }
}
// K3b — scatter run starts
__global__ void __launch_bounds__(256) k_vg_starts(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ heads_scan, const int n, uint32_t* __restrict__ ukey,
                                                   uint32_t* __restrict__ ustart, const VgLayout* __restrict__ Lp, const int compact)
{
  pdl_enter();
  const int rows = compact ? (int)Lp->n_valid : n;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < rows; i += gridDim.x * 256)
  {
    const uint32_t k = keys[i];
    if (k != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != k))
    {
      const uint32_t r = heads_scan[i];
      ukey[r] = k;
      ustart[r] = (uint32_t)i;
    }
  }
}
// K3c — voxel centre + weight (voxel_grid_weighted.cpp:169-188); counted variant: voxel_grid_counted.cpp:179-194
__global__ void __launch_bounds__(256) k_vg_emit(const uint32_t* __restrict__ ukey, const uint32_t* __restrict__ ustart, const VgLayout* __restrict__ Lp,
                                                 const unsigned long long* __restrict__ d_m, const uint32_t* __restrict__ over_prefix, vofod_vox* __restrict__ out,
                                                 const size_t out_cap)
{
  pdl_enter();
  const VgLayout L = *Lp;
  const unsigned m = (unsigned)*d_m;
  for (unsigned r = blockIdx.x * 256 + threadIdx.x; r < m && r < out_cap; r += gridDim.x * 256)
  {
    const uint32_t key = ukey[r];
    const uint32_t first = ustart[r];
    const uint32_t last = (r + 1 < m) ? ustart[r + 1] : L.n_valid;
    const int d0 = L.div[0], d01 = L.div[0] * L.div[1];
    const int ijk2 = (int)(key / (uint32_t)d01);
    const int rem = (int)(key - (uint32_t)ijk2 * (uint32_t)d01);
    const int ijk1 = rem / d0;
    const int ijk0 = rem - ijk1 * d0;
    vofod_vox v;
    v.x = ((float)ijk0 + 0.5f) * L.leaf + L.offset[0];
    v.y = ((float)ijk1 + 0.5f) * L.leaf + L.offset[1];
    v.z = ((float)ijk2 + 0.5f) * L.leaf + L.offset[2];
    v.count = over_prefix ? (over_prefix[last] - over_prefix[first]) : (last - first);
    out[r] = v;
  }
}
// counted variant: flag = intensity > threshold on the UNSORTED input
__global__ void __launch_bounds__(256) k_vg_over_flags(const vofod_xyzi* __restrict__ in, const unsigned long long* __restrict__ d_n, const int cap, const float thr,
                                                      uint32_t* __restrict__ flags)
{
  pdl_enter();
  const int n = (int)prims::dev_count(d_n, (size_t)cap);
  for (int i = blockIdx.x * 256 + threadIdx.x; i <= cap; i += gridDim.x * 256)
    flags[i] = (i < n && in[i].intensity > thr) ? 1u : 0u;
}

// shared tail: pts (float4, w = validity) + minmax -> ctx->vox / CNT_VG_M.  key_bits_hint = 0 => sort all 32 bits.
static int vg_run(vofod_ctx* ctx, const size_t n, const float leaf, const bool align, const float* align_center, const int key_bits_hint, const vofod_xyzi* d_counted_in,
                  const unsigned long long* d_counted_n, const float counted_thr, DevBuf& out_buf, const int slot_m = CNT_VG_M, const int slot_nvalid = CNT_VG_NVALID, const int slot_overflow = CNT_VG_OVERFLOW,
                  const bool compact = false)
{
  // `n` is the CAPACITY of the input (all kernels run over it); rows past the device-side count carry w == 0 (invalid)
  using namespace prims;
  const size_t np = padded(n);
  ENSURE(ctx->vg_keys_a, np * 4);
  ENSURE(ctx->vg_keys_b, np * 4);
  ENSURE(ctx->vg_flags, np * 4);
  ENSURE(ctx->vg_scan, np * 4);
  ENSURE(ctx->vg_ukey, np * 4);
  ENSURE(ctx->vg_ustart, np * 4);
  ENSURE(ctx->scratch_d, sizeof(MinMax) + sizeof(VgLayout) + 64);
  ENSURE(out_buf, np * sizeof(vofod_vox));
  MinMax* mm = ctx->scratch_d.as<MinMax>();
  VgLayout* L = reinterpret_cast<VgLayout*>(ctx->scratch_d.as<char>() + 64);
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  const float ac[3] = {align ? align_center[0] : 0.f, align ? align_center[1] : 0.f, align ? align_center[2] : 0.f};
  LAUNCH(k_vg_layout, 1, 1, 0, mm, leaf, align ? 1 : 0, ac[0], ac[1], ac[2], L, cnt, slot_nvalid, slot_overflow);
  const int nb = vf_blocks(ctx, n, 256, 8);
  LAUNCH(k_vg_keys, nb, 256, 0, ctx->vg_pts.as<float4>(), (int)n, L, ctx->vg_keys_a.as<uint32_t>(), compact ? 1 : 0);
  // compact input: only the first n_valid rows exist; that count sits in the counter slot k_vg_layout filled
  const unsigned long long* d_rows = compact ? cnt + slot_nvalid : nullptr;
  uint32_t* sorted = nullptr;
  const int bits = key_bits_hint > 0 && key_bits_hint < 32 ? key_bits_hint : 32;
  RET((radix_sort<uint32_t, false>(ctx, ctx->vg_keys_a.as<uint32_t>(), ctx->vg_keys_b.as<uint32_t>(), nullptr, nullptr, d_rows, n, 0, bits, &sorted, nullptr)));
  LAUNCH(k_vg_heads, nb, 256, 0, sorted, (int)n, ctx->vg_flags.as<uint32_t>(), L, compact ? 1 : 0);
  RET(scan_excl_u32(ctx, ctx->vg_flags.as<uint32_t>(), ctx->vg_scan.as<uint32_t>(), nullptr, n, cnt + slot_m));
  LAUNCH(k_vg_starts, nb, 256, 0, sorted, ctx->vg_scan.as<uint32_t>(), (int)n, ctx->vg_ukey.as<uint32_t>(), ctx->vg_ustart.as<uint32_t>(), L, compact ? 1 : 0);
  const uint32_t* over_prefix = nullptr;
  if (d_counted_in)
  {
    ENSURE(ctx->vg_pref, np * 4);
    // reuse vg_flags for the over-threshold flags (heads are no longer needed)
    LAUNCH(k_vg_over_flags, nb, 256, 0, d_counted_in, d_counted_n, (int)n, counted_thr, ctx->vg_flags.as<uint32_t>());
    RET(scan_excl_u32(ctx, ctx->vg_flags.as<uint32_t>(), ctx->vg_pref.as<uint32_t>(), nullptr, n + 1, nullptr));
    over_prefix = ctx->vg_pref.as<uint32_t>();
  }
  LAUNCH(k_vg_emit, nb, 256, 0, ctx->vg_ukey.as<uint32_t>(), ctx->vg_ustart.as<uint32_t>(), L, cnt + slot_m, over_prefix, out_buf.as<vofod_vox>(), np);
  return 0;
}

static int bits_for(unsigned long long v)
{
  int b = 0;
  while (v)
  {
    b++;
    v >>= 1;
  }
  return b;
}

// seed_cluster: also prepare the grid clustering of the output (returns 1 when that was done, 0 when the caller has to run
// the generic clustering)
int vf_filter_voxelize_dev(vofod_ctx* ctx, size_t n, const vofod_params& p, bool seed_cluster)
{
  const size_t np = prims::padded(n);
  ENSURE(ctx->vg_pts, np * 16);
  ENSURE(ctx->scratch_d, sizeof(MinMax) + sizeof(VgLayout) + 64);
  CropArgs a;
  {
    // vofod_nodelet.cpp:204, 626-629 (host fp32)
    volatile float ez = p.exclude_box_offset[2] + p.exclude_box_size[2] / 2.0f;
    for (int k = 0; k < 3; k++)
    {
      const float o = k == 2 ? (float)ez : p.exclude_box_offset[k];
      volatile float h = p.exclude_box_size[k] / 2;
      a.ex_max[k] = o + h;
      a.ex_min[k] = o - h;
    }
    // :212, 645-648
    volatile float oz = p.oparea_offset[2] + p.oparea_size[2] / 2.0f;
    for (int k = 0; k < 3; k++)
    {
      const float o = k == 2 ? (float)oz : p.oparea_offset[k];
      volatile float h = p.oparea_size[k] / 2;
      a.op_max[k] = o + h;
      a.op_min[k] = o - h;
    }
  }
  a.n = (int)n;
  MinMax* mm = ctx->scratch_d.as<MinMax>();
  if (!ctx->scan_prezero)  // inside vofod_process_scan the scan's first kernel has done it (vf_begin_scan)
    LAUNCH(k_minmax_init, 1, 1, 0, mm);
  LAUNCH(k_crop_transform, (int)((n + 255) / 256), 256, 0, a, ctx->dyn.as<ScanDyn>(), ctx->vg_pts.as<float4>(), mm);
  // align to the map: idxToCoord(0,0,0) (vofod_nodelet.cpp:664-665)
  const Geom& g = ctx->g;
  float ac[3];
  for (int k = 0; k < 3; k++)
  {
    volatile float c0 = (0 + 0.5f) * g.vs;
    ac[k] = c0 + g.off[k];
  }
  // the crop to the operation area bounds the key range: div <= size/leaf + 3 per axis
  unsigned long long cells = 1;
  int budget[3];
  for (int k = 0; k < 3; k++)
  {
    const double d = ceil((double)p.oparea_size[k] / (double)g.vs) + 3.0;
    budget[k] = d < 2147483647.0 ? (int)d : 2147483647;
    cells = (d < 2097152.0 && cells < (1ull << 42)) ? cells * (unsigned long long)d : ~0ull;  // (saturates instead of wrapping)
  }
  const int bits = cells < (1ull << 31) - 1 ? bits_for(cells + 1) : 32;
  if (!ctx->vg_force_sort && cells < (1ull << 31) && cells * 4 <= (size_t(8) << 30))
  {
    // sort-free path: dense counts + occupancy words over the key range the crop allows
    using namespace prims;
    size_t words_cap = 1;
    for (int k = 0; k < 3; k++)
    {
      const size_t d = (size_t)(ceil((double)p.oparea_size[k] / (double)g.vs) + 3.0);
      words_cap *= k == 0 ? (d + 31) / 32 : d;
    }
    ENSURE(ctx->vgh_cnt, (size_t)cells * 4);            // zero when (re)allocated, put back to zero by the emission pass
    ENSURE(ctx->vgh_bits, padded(words_cap) * 4);       // the same
    ENSURE(ctx->vox, np * sizeof(vofod_vox));
    unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
    VgLayout* L = reinterpret_cast<VgLayout*>(ctx->scratch_d.as<char>() + 64);
    ZERO_CNT(CNT_VG_OVERFLOW, 1);  // (inside a scan k_begin_call has done it: the kernel's blocks only ever raise the flag)
    LAUNCH(k_vgh_count, vf_blocks(ctx, n, 256, 8), 256, 0, ctx->vg_pts.as<float4>(), mm, g.vs, ac[0], ac[1], ac[2], L, cnt, (int)CNT_VG_NVALID, (int)CNT_VG_OVERFLOW,
           budget[0], budget[1], budget[2], ctx->vgh_cnt.as<uint32_t>(), ctx->vgh_bits.as<uint32_t>());
    // a word holds at least one of the n points
    const size_t list_cap = n < words_cap ? n : words_cap;
    ENSURE(ctx->vgh_list, (list_cap + 1) * sizeof(uint4));
    VghSeeds seeds = {};
    RunRows rows;
    const bool seed = seed_cluster && !ctx->cl_force_hash && vf_run_rows((float)p.ground_points_max_distance, g.vs, rows);
    if (seed)
    {
      ENSURE(ctx->cl_cellkey, np * 4);
      ENSURE(ctx->cl_words, (words_cap + 1) * sizeof(RunWord));
      ENSURE(ctx->cl.parent, n * 4);
      ENSURE(ctx->cl.sizes, n * 4);
      ENSURE(ctx->cl.minidx, n * 4);
      seeds.cellkey = ctx->cl_cellkey.as<uint32_t>();
      seeds.words = ctx->cl_words.as<RunWord>();
      seeds.parent = ctx->cl.parent.as<int>();
      seeds.sizes = ctx->cl.sizes.as<int>();
      seeds.minidx = ctx->cl.minidx.as<int>();
      seeds.counters = ctx->d_counters.as<unsigned long long>();
    }
    ZERO_CNT(CNT_VGH_LIST, 1);
    RET(scan_excl_u32_pair(ctx, ctx->vgh_bits.as<uint32_t>(), nullptr, cnt + CNT_VGH_WORDS, words_cap, cnt + CNT_VG_M, true, nullptr, nullptr, 0,
                           nullptr, false, ctx->vgh_list.as<uint4>(), cnt + CNT_VGH_LIST));
    LAUNCH(k_vgh_emit, vf_blocks(ctx, list_cap * 32, 256, 8), 256, 0, L, ctx->vgh_list.as<uint4>(), cnt + CNT_VGH_LIST, list_cap, ctx->vgh_bits.as<uint32_t>(),
           ctx->vgh_cnt.as<uint32_t>(), ctx->vox.as<vofod_vox>(), np, seeds);
    return seed ? 1 : 0;
  }
  return vg_run(ctx, n, g.vs, true, ac, bits, nullptr, nullptr, 0.f, ctx->vox, CNT_VG_M, CNT_VG_NVALID, CNT_VG_OVERFLOW, true);
}

static int read_m(vofod_ctx* ctx, size_t* m, int* overflow)
{
  unsigned long long v[2] = {0, 0};
  CK(cudaMemcpyAsync(&v[0], vf_cnt(ctx, CNT_VG_M), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&v[1], vf_cnt(ctx, CNT_VG_OVERFLOW), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  unsigned long long wd = 0;
  CK(cudaMemcpy(&wd, vf_cnt(ctx, CNT_WATCHDOG), 8, cudaMemcpyDeviceToHost));
  if (wd)
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", wd);
  *m = (size_t)v[0];
  *overflow = (int)v[1];
  return 0;
}

extern "C" {

int vofod_filter_voxelize(vofod_ctx* ctx, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, vofod_vox* out, size_t cap, size_t* m)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  if (!tf || !p || !m || (n && !scan))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  *m = 0;
  if (n == 0)
    return VOFOD_OK;
  ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
  CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  memcpy(ctx->h_dyn->tf.R, tf->R, sizeof(tf->R));
  memcpy(ctx->h_dyn->tf.t, tf->t, sizeof(tf->t));
  ctx->h_dyn->scan = ctx->scan_staging.as<vofod_pt>();
  RET(vf_begin_call(ctx));
  RET(vf_dyn_push(ctx));
  RET(vf_filter_voxelize_dev(ctx, n, *p));
  int overflow = 0;
  RET(read_m(ctx, m, &overflow));
  ctx->last_m = *m;
  if (overflow)
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "leaf size too small for the input: integer indices would overflow");
  if (*m > cap || (*m && !out))
    return vf_fail(ctx, VOFOD_E_CAPACITY, "filter_voxelize: need capacity %zu", *m);
  if (*m)
  {
    CK(cudaMemcpyAsync(out, ctx->vox.p, *m * sizeof(vofod_vox), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return VOFOD_OK;
}

static int vg_generic(vofod_ctx* ctx, const float* host_in, int stride, size_t n, float leaf, const float* align, bool counted, float thr, vofod_vox* out, size_t cap, size_t* m)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!m || (n && !host_in) || !(leaf > 0.0f))
    return vf_fail(ctx, VOFOD_E_INVALID, "bad argument (leaf size must be > 0)");
  *m = 0;
  if (n == 0)
    return VOFOD_OK;
  const size_t np = prims::padded(n);
  ENSURE(ctx->vg_pts, np * 16);
  ENSURE(ctx->scratch_a, n * stride * 4 + 64);
  ENSURE(ctx->scratch_d, sizeof(MinMax) + sizeof(VgLayout) + 64);
  CK(cudaMemcpyAsync(ctx->scratch_a.p, host_in, n * stride * 4, cudaMemcpyHostToDevice, ctx->stream));
  RET(vf_begin_call(ctx));
  MinMax* mm = ctx->scratch_d.as<MinMax>();
  LAUNCH(k_minmax_init, 1, 1, 0, mm);
  LAUNCH(k_load_cloud, (int)((n + 255) / 256), 256, 0, ctx->scratch_a.as<float>(), stride, nullptr, (int)n, ctx->vg_pts.as<float4>(), mm);
  RET(vg_run(ctx, n, leaf, align != nullptr, align, 0, counted ? ctx->scratch_a.as<vofod_xyzi>() : nullptr, nullptr, thr, ctx->sep_ds));
  int overflow = 0;
  RET(read_m(ctx, m, &overflow));
  if (overflow)
    return vf_fail(ctx, VOFOD_E_OVERFLOW, "leaf size too small for the input: integer indices would overflow");
  if (*m > cap || (*m && !out))
    return vf_fail(ctx, VOFOD_E_CAPACITY, "voxel_grid: need capacity %zu", *m);
  if (*m)
  {
    CK(cudaMemcpyAsync(out, ctx->sep_ds.p, *m * sizeof(vofod_vox), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return VOFOD_OK;
}

int vofod_voxel_grid_weighted(vofod_ctx* ctx, const float* xyz, size_t n, float leaf, const float align[3], vofod_vox* out, size_t cap, size_t* m)
{
  return vg_generic(ctx, xyz, 3, n, leaf, align, false, 0.f, out, cap, m);
}
int vofod_voxel_grid_counted(vofod_ctx* ctx, const vofod_xyzi* pts, size_t n, float leaf, float threshold, const float align[3], vofod_vox* out, size_t cap, size_t* m)
{
  return vg_generic(ctx, reinterpret_cast<const float*>(pts), 4, n, leaf, align, true, threshold, out, cap, m);
}
}

// used by sepclusters.cu: counted voxel grid over device-resident xyzi points; `cap` rows are processed, of which the
// first *d_n (device-side count, NULL = all) are real
int vf_voxel_grid_counted_dev(vofod_ctx* ctx, const vofod_xyzi* d_in, const unsigned long long* d_n, size_t cap, float leaf, float thr, DevBuf& out)
{
  const size_t np = prims::padded(cap);
  ENSURE(ctx->vg_pts, np * 16);
  ENSURE(ctx->scratch_d, sizeof(MinMax) + sizeof(VgLayout) + 64);
  MinMax* mm = ctx->scratch_d.as<MinMax>();
  LAUNCH(k_minmax_init, 1, 1, 0, mm);
  LAUNCH(k_load_cloud, (int)((cap + 255) / 256), 256, 0, reinterpret_cast<const float*>(d_in), 4, d_n, (int)cap, ctx->vg_pts.as<float4>(), mm);
  // separate counter slots: the per-scan voxel counts (CNT_VG_*) must survive the background-cluster pass
  return vg_run(ctx, cap, leaf, false, nullptr, 0, d_in, d_n, thr, out, CNT_SEP_KDS, CNT_SCRATCH1, CNT_SEP_NUNIQ);
}
