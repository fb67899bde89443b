// raycast_cloud (vofod_nodelet.cpp:1397-1606) on the GPU.
//
//  accumulate: one thread per LiDAR ray runs the reference's 3-D DDA (voxel_map.cpp:229-263) bit for bit.
//              Instead of `raycast[v] += ddist` in sequential fp32 (order dependent), every callback adds
//              (1 << 44) + round(ddist * 2^F) to a packed u64 cell of a dense accumulator WINDOW (the AABB of
//              sensor +- max_dist): callback COUNT in the top 20 bits (bit exact), path length as an exact,
//              order-independent fixed-point sum in the low 44 bits.  A warp (32 neighbouring columns of one
//              LiDAR row = 0.1 rad) mostly walks through the same voxels, so lanes that are in the same voxel at
//              the same step are merged with match.any + redux and issue ONE 64-bit RED to L2.
//  apply     : one pass over the window only (cells outside have raycast == 0 => untouched by the reference's
//              full-grid forEachIdx): decode, mix into the score grid with the reference's new/old rule,
//              and zero the accumulator cell for the next scan (fuses m_voxel_raycast.clear(), :1430).
//  flags     : m_voxel_flags.clear() (:1602) touches only the cells the point update flagged (list kept by
//              update_points), not the whole grid.
#include <math.h>

#include "common.cuh"
#include "prims.cuh"

struct RayArgs
{
  Geom g;
  float max_dist, min_intensity;
  float scale;  // 2^F
  int n;
  int has_off;
};

__device__ __forceinline__ void red_add_u64(unsigned long long* addr, const unsigned long long v)
{
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// RB = rays (threads) per block.  Ray lengths differ a lot between LiDAR rows (no-return rays walk max_dist, ground
// returns a few metres), so small blocks balance better: 128 threads = 4 warps = 128 neighbouring columns of one row.
template <int RB, bool AGG, bool SLAB>
__global__ void __launch_bounds__(RB) k_raycast_accumulate(const RayArgs a, const ScanDyn* __restrict__ dyn, const float4* __restrict__ lut_dir,
                                                           const float4* __restrict__ lut_off, const uint8_t* __restrict__ mask,
                                                           unsigned long long* __restrict__ acc, unsigned long long* __restrict__ counters)
{
  pdl_enter();
  // stage this block's packed points (20 B each) through shared memory with coalesced 16 B loads
  __shared__ __align__(16) uint32_t s_pts[RB * 5];
  const vofod_pt* __restrict__ scan = dyn->scan;
  const Pose33 tf = dyn->tf;
  const Window w = dyn->win;
  const int blk_first = blockIdx.x * RB;
  {
    const int n_here = min(RB, a.n - blk_first);
    const int n_words = n_here * 5;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(scan) + (size_t)blk_first * 5;
    const int n_vec = n_words / 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);  // blk_first*20 B is a multiple of 16 (RB % 4 == 0)
    for (int i = threadIdx.x; i < n_vec; i += RB)
      reinterpret_cast<uint4*>(s_pts)[i] = __ldg(src4 + i);
    for (int i = n_vec * 4 + threadIdx.x; i < n_words; i += RB)
      s_pts[i] = __ldg(src + i);
  }
  __syncthreads();

  const int idx = blk_first + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  bool alive = idx < a.n;
  float len = 0.f;
  float tmx = 0.f, tmy = 0.f, tmz = 0.f, tdx = 0.f, tdy = 0.f, tdz = 0.f;
  int remx = 0, remy = 0, remz = 0;  // steps left before the ray stands in the last voxel of the map along that axis
  int dwx = 0, dwy = 0, dwz = 0;     // window-index stride of one step along each axis
  int widx = 0;
  int spos = 0, sstep = 0;  // SLAB: window coordinate / step along the slab axis (the only axis a ray can leave the window on)
  const int ssize = w.size[a.g.slab_axis];
  const int wn = w.size[0] * w.size[1] * w.size[2];
  if (alive)
  {
    const float intensity = __uint_as_float(s_pts[threadIdx.x * 5 + 3]);
    const uint32_t range = s_pts[threadIdx.x * 5 + 4];
    // vofod_nodelet.cpp:1449
    if (intensity < a.min_intensity || (mask[idx] == 0 && range == 0))
      alive = false;
    const float4 d1 = __ldg(lut_dir + idx);
    // Eigen 3x3*3 coefficient product: r_i = R_i0*v0 + (R_i1*v1 + R_i2*v2)   (:1453)
    const float dx = tf.R[0] * d1.x + (tf.R[1] * d1.y + tf.R[2] * d1.z);
    const float dy = tf.R[3] * d1.x + (tf.R[4] * d1.y + tf.R[5] * d1.z);
    const float dz = tf.R[6] * d1.x + (tf.R[7] * d1.y + tf.R[8] * d1.z);
    const float ray_dist = 0.001f * (float)range;                                             // :1456
    const float dmv = ray_dist - a.g.vs;
    len = ray_dist == 0.0f ? a.max_dist : (a.max_dist < dmv ? a.max_dist : dmv);              // :1457 (std::min)
    float sx = tf.t[0], sy = tf.t[1], sz = tf.t[2];
    if (a.has_off)
    {
      const float4 o1 = __ldg(lut_off + idx);
      sx = (tf.R[0] * o1.x + (tf.R[1] * o1.y + tf.R[2] * o1.z)) + tf.t[0];                    // :1477
      sy = (tf.R[3] * o1.x + (tf.R[4] * o1.y + tf.R[5] * o1.z)) + tf.t[1];
      sz = (tf.R[6] * o1.x + (tf.R[7] * o1.y + tf.R[8] * o1.z)) + tf.t[2];
    }
    const int cx = coord_to_idx1(sx, a.g.off[0], a.g.inv);
    const int cy = coord_to_idx1(sy, a.g.off[1], a.g.inv);
    const int cz = coord_to_idx1(sz, a.g.off[2], a.g.inv);
    if (!in_limits_idx(a.g, cx, cy, cz))                                                      // :1482
      alive = false;
    // voxel_map.cpp:232-244
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const int stx = (dx > 0.0f) - (dx < 0.0f);
    const int sty = (dy > 0.0f) - (dy < 0.0f);
    const int stz = (dz > 0.0f) - (dz < 0.0f);
    tdx = (1.0f / ax) * a.g.vs;
    tdy = (1.0f / ay) * a.g.vs;
    tdz = (1.0f / az) * a.g.vs;
    const float ox = idx_to_coord1(cx, a.g.off[0], a.g.vs) - sx;
    const float oy = idx_to_coord1(cy, a.g.off[1], a.g.vs) - sy;
    const float oz = idx_to_coord1(cz, a.g.off[2], a.g.vs) - sz;
    tmx = (a.g.half + (float)stx * ox) / ax;
    tmy = (a.g.half + (float)sty * oy) / ay;
    tmz = (a.g.half + (float)stz * oz) / az;
    // `if (cur[i] == last[i]) break` with last = step > 0 ? size-1 : 0 (voxel_map.cpp:240-257), as a countdown
    remx = stx > 0 ? a.g.size[0] - 1 - cx : cx;
    remy = sty > 0 ? a.g.size[1] - 1 - cy : cy;
    remz = stz > 0 ? a.g.size[2] - 1 - cz : cz;
    dwx = stx;
    dwy = sty * w.size[0];
    dwz = stz * w.size[0] * w.size[1];
    widx = (cx - w.lo[0]) + (cy - w.lo[1]) * w.size[0] + (cz - w.lo[2]) * w.size[0] * w.size[1];
    if (SLAB)
    {
      spos = a.g.slab_axis == 0 ? cx - w.lo[0] : (a.g.slab_axis == 1 ? cy - w.lo[1] : cz - w.lo[2]);
      sstep = a.g.slab_axis == 0 ? stx : (a.g.slab_axis == 1 ? sty : stz);
    }
    if (!(0.0f < len))  // while (prev_dist < length) with prev_dist = 0
      alive = false;
  }
  float prev = 0.0f;
  unsigned steps = 0;
  unsigned oob = 0;
  int orq = 0;  // OR of all quantised path lengths: non-zero <=> some callback carried a positive length <=> max_element(raycast) > 0 (:1542-1548)
  // Every lane walks its own ray; the lanes that are in the loop together (__activemask) merge their updates.  No warp-wide
  // vote per step: lanes whose ray has ended simply leave the loop, and if the hardware ever splits the warp the only effect
  // is less merging — the additions commute.
  while (alive)
  {
    // tmax.minCoeff(&i): first minimum, strict '<'
    const bool use_y = tmy < tmx;
    const float d01 = use_y ? tmy : tmx;
    const bool use_z = tmz < d01;
    const float dist = use_z ? tmz : d01;
    const float ddist = (len < dist ? len : dist) - prev;                                     // voxel_map.cpp:252
    const int q = __float2int_rn(ddist * a.scale);
    int key;
    bool inside = true;
    if (SLAB)
    {
      // the window is cut at the slab's storage box: rays enter and leave it along the slab axis
      inside = (unsigned)spos < (unsigned)ssize;
      key = inside ? widx : -1 - (int)lane;
    } else
      // unsharded: the window holds every voxel within max_dist, so this clamp never bites; if it ever did, the update would
      // land in the spare cell behind the window (and the apply kernel reports it) instead of corrupting memory
      key = (int)min((unsigned)widx, (unsigned)wn);
    if (AGG)
    {
      // lanes standing in the same voxel issue ONE 64-bit RED: count in the top 20 bits, path length below
      const unsigned am = __activemask();
      const unsigned m = __match_any_sync(am, key);
      const int sum = __reduce_add_sync(m, q);
      if (inside && lane == (unsigned)(__ffs(m) - 1))
        red_add_u64(acc + key, ((unsigned long long)__popc(m) << ACC_LEN_BITS) + (unsigned long long)(long long)sum);
    } else if (inside)
      red_add_u64(acc + key, (1ull << ACC_LEN_BITS) + (unsigned long long)(long long)q);
    orq |= q;
    steps++;
    prev = dist;
    // voxel_map.cpp:257-261
    const int rem = use_z ? remz : (use_y ? remy : remx);
    alive = rem != 0 && dist < len;
    if (use_z)
    {
      remz--; tmz += tdz; widx += dwz;
      if (SLAB && a.g.slab_axis == 2) spos += sstep;
    } else if (use_y)
    {
      remy--; tmy += tdy; widx += dwy;
      if (SLAB && a.g.slab_axis == 1) spos += sstep;
    } else
    {
      remx--; tmx += tdx; widx += dwx;
      if (SLAB && a.g.slab_axis == 0) spos += sstep;
    }
  }
  __syncwarp();
  const bool anyq = orq != 0;
  if (__any_sync(VOFOD_FULL, anyq) && lane == 0)
    atomicOr(counters + CNT_APPLY_ANY, 1ull);
  // per-block totals
  unsigned tot = prims::warp_sum(steps);
  unsigned toob = prims::warp_sum(oob);
  __shared__ unsigned s_tot[RB / 32], s_oob[RB / 32];
  if (lane == 0)
  {
    s_tot[threadIdx.x >> 5] = tot;
    s_oob[threadIdx.x >> 5] = toob;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned t = 0, o = 0;
    for (int i = 0; i < RB / 32; i++)
    {
      t += s_tot[i];
      o += s_oob[i];
    }
    if (t)
      atomicAdd(counters + CNT_TRAVERSALS, (unsigned long long)t);
    if (o)
      atomicAdd(counters + CNT_OOB, (unsigned long long)o);
  }
}

struct ApplyArgs
{
  Geom g;
  float inv_scale_unused;
  double inv_scale;       // 2^-F
  float ray_score;        // :1553
  float weighting_factor; // :1556 (new rule)
  float ray_weight;       // :1554 (old rule)
  float max_val;          // old rule
  int new_rule;
};

// max_element of the accumulator (:1542) — only needed by the old update rule
__global__ void __launch_bounds__(256) k_raycast_max(const ScanDyn* __restrict__ dyn, const unsigned long long* __restrict__ acc, const double inv_scale, unsigned* __restrict__ out_bits)
{
  pdl_enter();
  const Window w = after_wait(dyn)->win;
  const long long n = (long long)w.size[0] * w.size[1] * w.size[2];
  float mx = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const unsigned long long p = acc[i];
    if (p)
    {
      unsigned c;
      long long lq;
      acc_decode(p, c, lq);
      const float rv = (float)((double)lq * inv_scale);
      mx = fmaxf(mx, rv);
    }
  }
  for (int o = 16; o > 0; o >>= 1)
    mx = fmaxf(mx, __shfl_xor_sync(VOFOD_FULL, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f)
    atomicMax(out_bits, __float_as_uint(mx));  // positive floats order like their bit patterns
}

__global__ void __launch_bounds__(256) k_raycast_apply(const ApplyArgs a, const ScanDyn* __restrict__ dyn, unsigned long long* __restrict__ acc, float* __restrict__ score,
                                                       const uint8_t* __restrict__ flags, const unsigned* __restrict__ max_bits,
                                                       unsigned long long* __restrict__ counters)
{
  pdl_enter();
  const Window w = dyn->win;
  const float its = (float)dyn->its_raycast;  // detection_its_diff as float (:1539)
  const long long n = (long long)w.size[0] * w.size[1] * w.size[2];
  // the spare cell behind the window catches updates the accumulate kernel had to clamp (must never happen)
  if (blockIdx.x == 0 && threadIdx.x == 0 && acc[n] != 0ull)
  {
    atomicAdd(counters + CNT_OOB, 1ull);
    acc[n] = 0ull;
  }
  const int wsx = w.size[0], wsy = w.size[1];
  float max_val = 0.f;
  if (!a.new_rule)
  {
    max_val = __uint_as_float(*max_bits);
    if (max_val == 0.0f)
      return;  // :1544-1548 (the host also skips the flag clear)
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const unsigned long long p = acc[i];
    if (!p)
      continue;
    acc[i] = 0ull;  // m_voxel_raycast.clear() for the next scan (:1430)
    unsigned c;
    long long lq;
    acc_decode(p, c, lq);
    const float rv = (float)((double)lq * a.inv_scale);
    if (!(rv > 0.0f))
      continue;
    const int wx = (int)(i % wsx), wy = (int)((i / wsx) % wsy), wz = (int)(i / ((long long)wsx * wsy));
    const long long ci = cell_index(a.g, wx + w.lo[0], wy + w.lo[1], wz + w.lo[2]);
    if (ci < 0 || flags[ci] != 0)  // flag == m_vflags_unmarked (:1561)
      continue;
    const float m = score[ci];
    float w1;
    if (a.new_rule)
    {
      const float n_int = a.weighting_factor * rv;                    // :1565
      w1 = (float)exp2((double)(-its * n_int));                     // :1567  std::pow(2, float) -> double pow
    } else
    {
      const float norm_val = rv / max_val;                            // :1587
      const float ws = a.ray_weight * sqrtf(norm_val);                // :1591
      // :1593 std::pow(float, float).  CUDA's powf is a few ulp off; the fp64 pow rounded to fp32 reproduces the host's
      // (correctly rounded) powf except in vanishingly rare double-rounding cases
      w1 = (float)pow((double)(1.0f - ws), (double)its);
      w1 = w1 < 0.0f ? 0.0f : (1.0f < w1 ? 1.0f : w1);                // std::clamp
    }
    const float w2 = 1.0f - w1;
    score[ci] = w1 * m + w2 * a.ray_score;                            // :1569 / :1597
  }
}

// m_voxel_flags.clear() (:1602) restricted to the cells flagged since the last clear; skipped when the apply was skipped
__global__ void k_clear_flags(uint8_t* __restrict__ flags, const uint32_t* __restrict__ flagged, const size_t flagged_cap, const long long n_cells,
                              unsigned long long* __restrict__ counters, const int force_full)
{
  pdl_enter();
  if (counters[CNT_APPLY_ANY] == 0ull)
    return;
  const bool full = force_full || counters[CNT_FLAGGED_OVERFLOW] != 0ull;
  if (full)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x)
      flags[i] = 0;
  } else
  {
    unsigned long long nf = counters[CNT_FLAGGED];
    if (nf > flagged_cap)
      nf = flagged_cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (unsigned long long)gridDim.x * blockDim.x)
      flags[flagged[i]] = 0;
  }
  // the last block to finish empties the list (every block has read its length by then)
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    if (atomicAdd(counters + CNT_CLEAR_TICKET, 1ull) == (unsigned long long)gridDim.x - 1ull)
    {
      counters[CNT_FLAGGED] = 0ull;
      counters[CNT_FLAGGED_OVERFLOW] = 0ull;
      counters[CNT_CLEAR_TICKET] = 0ull;
    }
  }
}
// parity/debug: expand the window accumulator into full-grid planes
__global__ void k_raycast_expand(const Geom g, const Window w, const unsigned long long* __restrict__ acc, const double inv_scale, uint32_t* __restrict__ counts,
                                 float* __restrict__ lengths)
{
  pdl_enter();
  const long long n = (long long)w.size[0] * w.size[1] * w.size[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const unsigned long long p = acc[i];
    if (!p)
      continue;
    unsigned c;
    long long lq;
    acc_decode(p, c, lq);
    const int wx = (int)(i % w.size[0]), wy = (int)((i / w.size[0]) % w.size[1]), wz = (int)(i / ((long long)w.size[0] * w.size[1]));
    const long long ci = cell_index(g, wx + w.lo[0], wy + w.lo[1], wz + w.lo[2]);
    if (ci < 0)
      continue;
    if (counts)
      counts[ci] = c;
    if (lengths)
      lengths[ci] = (float)((double)lq * inv_scale);
  }
}

static int choose_frac_bits(const size_t n_rays, const float vs)
{
  const double diag = (double)vs * 1.7320508075688772;
  const int b_sum = (int)ceil(log2((double)n_rays * diag + 1.0));   // bits of the largest possible per-cell sum [m]
  const int b_one = (int)ceil(log2(diag));                          // bits of one ddist [m]
  int f = 42 - b_sum;                                               // |sum| < 2^43 in the 44-bit signed field
  if (25 - b_one < f)
    f = 25 - b_one;                                                 // 32 lanes * q must fit int32 for the warp redux
  if (f < 4) f = 4;
  if (f > 30) f = 30;
  return f;
}

int vf_raycast_expand(vofod_ctx* ctx, uint32_t* d_counts, float* d_lengths)
{
  const size_t n = (size_t)geom_cells(ctx->g);
  if (d_counts)
    CK(cudaMemsetAsync(d_counts, 0, n * 4, ctx->stream));
  if (d_lengths)
    CK(cudaMemsetAsync(d_lengths, 0, n * 4, ctx->stream));
  if (!ctx->win_valid || !ctx->acc_has_data)
    return 0;
  const size_t wn = (size_t)ctx->win.size[0] * ctx->win.size[1] * ctx->win.size[2];
  LAUNCH(k_raycast_expand, vf_blocks(ctx, wn, 256), 256, 0, ctx->g, ctx->win, ctx->acc.as<unsigned long long>(), ldexp(1.0, -ctx->frac_bits), d_counts, d_lengths);
  return 0;
}

// Host-side preparation of a raycast (no launches): the sensor-in-map test of vofod_nodelet.cpp:1432 and the accumulator
// window = voxels that can be reached from the sensor within max_dist (+ LUT offsets), clamped to the held box.  Writes the
// window into ctx->h_dyn (the caller pushes h_dyn to the device before any kernel of the call runs).
// Returns VOFOD_OK, VOFOD_W_PAUSED or VOFOD_W_SENSOR_OOB.
int vf_raycast_prepare(vofod_ctx* ctx, size_t n, const vofod_pose& tf, const vofod_params& p)
{
  if (p.raycast_pause)
    return VOFOD_W_PAUSED;
  if (!ctx->W || n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  const Geom& g = ctx->g;
  const float max_dist = (float)p.raycast_max_distance;
  // sensor out of bounds => no raycast (:1432,1523-1526)
  for (int a = 0; a < 3; a++)
  {
    volatile float d = tf.t[a] - g.off[a];
    volatile float q = d * g.inv;
    const int c = (int)floorf(q);
    if (c < 0 || c >= g.size[a])
      return VOFOD_W_SENSOR_OOB;
  }
  Window w;
  const float reach = max_dist + ctx->lut_max_off;
  size_t wn_max = 1;
  for (int a = 0; a < 3; a++)
  {
    const int c = (int)floorf((tf.t[a] - g.off[a]) * g.inv);
    const int r = (int)ceilf(reach * g.inv) + 2;
    int lo = c - r, hi = c + r + 1;
    const int blo = g.st_lo[a], bhi = g.st_lo[a] + g.st_size[a];
    if (lo < blo) lo = blo;
    if (hi > bhi) hi = bhi;
    if (hi <= lo) { lo = blo; hi = blo + 1; }
    w.lo[a] = lo;
    w.size[a] = hi - lo;
    const int full = 2 * r + 1 < g.st_size[a] ? 2 * r + 1 : g.st_size[a];
    wn_max *= (size_t)full;
  }
  // sized for the largest window this max_dist can produce, so that the buffer (and the launch grids) never change
  if (ctx->acc.cap < (wn_max + 32) * 8)
  {
    ENSURE(ctx->acc, (wn_max + 32) * 8);  // fresh allocations are zero-filled
    ctx->acc_has_data = false;
  }
  ctx->acc_cells_max = wn_max;
  ctx->win = w;
  ctx->win_valid = true;
  ctx->h_dyn->win = w;
  ctx->frac_bits = choose_frac_bits(n, ctx->g.vs);
  return VOFOD_OK;
}

int vf_raycast_accumulate_dev(vofod_ctx* ctx, size_t n, const vofod_pose& tf, const vofod_params& p)
{
  (void)tf;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  ZERO_CNT(CNT_TRAVERSALS, 1);
  ZERO_CNT(CNT_OOB, 1);
  ZERO_CNT(CNT_APPLY_ANY, 1);
  // m_voxel_raycast.clear() (:1430): an accumulate that was never applied (or an old-rule apply that bailed out on
  // max_val == 0) must not leak into this one.  After a new-rule apply the accumulator is already all zero.
  if (ctx->acc_has_data)
  {
    CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_cells_max * 8, ctx->stream));
    ctx->acc_has_data = false;
  }
  RayArgs a;
  a.g = ctx->g;
  a.max_dist = (float)p.raycast_max_distance;
  a.min_intensity = (float)p.raycast_min_intensity;
  a.scale = ldexpf(1.0f, ctx->frac_bits);
  a.n = (int)n;
  a.has_off = ctx->lut_has_off ? 1 : 0;
  const bool slab = ctx->slab_on;
#define RAY_LAUNCH(RB_, AGG_, SLAB_)                                                                                                                         \
  LAUNCH((k_raycast_accumulate<RB_, AGG_, SLAB_>), (int)((n + RB_ - 1) / RB_), RB_, 0, a, ctx->dyn.as<ScanDyn>(), ctx->lut_dir.as<float4>(),               \
         ctx->lut_off.as<float4>(), ctx->mask.as<uint8_t>(), ctx->acc.as<unsigned long long>(), cnt)
  if (ctx->raycast_no_agg)
  {
    if (slab) RAY_LAUNCH(128, false, true); else RAY_LAUNCH(128, false, false);
  } else if (slab)
    RAY_LAUNCH(128, true, true);
  else if (ctx->raycast_block == 64)
    RAY_LAUNCH(64, true, false);
  else if (ctx->raycast_block == 256)
    RAY_LAUNCH(256, true, false);
  else
    RAY_LAUNCH(128, true, false);
#undef RAY_LAUNCH
  ctx->acc_has_data = true;
  return VOFOD_OK;
}

// its_diff comes from ctx->dyn (its_raycast)
int vf_raycast_apply_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p)
{
  (void)its_diff;
  if (p.raycast_pause)
    return VOFOD_W_PAUSED;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  ZERO_CNT(CNT_MAXVAL, 1);
  if (!ctx->win_valid || !ctx->acc_has_data)
    return VOFOD_W_EMPTY_RAYCAST;  // max_val == 0 (:1544-1548): nothing applied, flags NOT cleared
  ApplyArgs a;
  a.g = ctx->g;
  a.inv_scale_unused = 0.f;
  a.inv_scale = ldexp(1.0, -ctx->frac_bits);
  a.ray_score = (float)p.score_ray;
  a.ray_weight = (float)p.raycast_weight_coefficient;
  {
    const float voxel_diag = (float)(sqrt(3.0) * (double)ctx->g.vs);  // :1555 (std::sqrt(3) is double)
    volatile float wf = a.ray_weight / voxel_diag;                    // :1556
    a.weighting_factor = wf;
  }
  a.max_val = 0.f;
  a.new_rule = p.raycast_new_update_rule ? 1 : 0;
  const size_t wn = ctx->acc_cells_max;  // grid sized for the largest window: identical launches from scan to scan
  const ScanDyn* dyn = ctx->dyn.as<ScanDyn>();
  if (!a.new_rule)
    LAUNCH(k_raycast_max, vf_blocks(ctx, wn, 256), 256, 0, dyn, ctx->acc.as<unsigned long long>(), a.inv_scale, (unsigned*)(cnt + CNT_MAXVAL));
  LAUNCH(k_raycast_apply, vf_blocks(ctx, wn, 256), 256, 0, a, dyn, ctx->acc.as<unsigned long long>(), ctx->score.as<float>(), ctx->flags.as<uint8_t>(),
         (const unsigned*)(cnt + CNT_MAXVAL), cnt);
  // NOTE (old rule): when max_val == 0 the apply kernel returns before zeroing; the accumulator then only holds cells
  // whose length is <= 0, which the next accumulate clears because acc_has_data stays true.
  ctx->acc_has_data = !a.new_rule;
  const long long n_cells = geom_cells(ctx->g);
  const size_t work = ctx->flags_full_dirty ? (size_t)n_cells : ctx->flagged_cap;
  LAUNCH(k_clear_flags, vf_blocks(ctx, work ? work : 1, 256), 256, 0, ctx->flags.as<uint8_t>(), ctx->flagged.as<uint32_t>(), ctx->flagged_cap, n_cells, cnt,
         ctx->flags_full_dirty ? 1 : 0);
  // flags_full_dirty can only be dropped once we know (on the host) that the clear really ran; callers that
  // read CNT_APPLY_ANY back do that (see vofod_raycast_apply / process_scan).
  return VOFOD_OK;
}

extern "C" {

int vofod_raycast_frac_bits(const vofod_ctx* ctx) { return ctx ? ctx->frac_bits : 0; }

int vofod_raycast_accumulate(vofod_ctx* ctx, const vofod_pt* scan, size_t n, const vofod_pose* tf, const vofod_params* p, uint64_t* n_traversals)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  if (!ctx->W)
    return vf_fail(ctx, VOFOD_E_STATE, "sensor not set");
  if (!scan || !tf || !p)
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (n != (size_t)ctx->W * ctx->H)
    return vf_fail(ctx, VOFOD_E_DIMS, "cloud has %zu points, sensor LUT has %zu", n, (size_t)ctx->W * ctx->H);
  ENSURE(ctx->scan_staging, n * sizeof(vofod_pt) + 64);
  CK(cudaMemcpyAsync(ctx->scan_staging.p, scan, n * sizeof(vofod_pt), cudaMemcpyHostToDevice, ctx->stream));
  if (n_traversals)
    *n_traversals = 0;
  const int prc = vf_raycast_prepare(ctx, n, *tf, *p);
  if (prc != VOFOD_OK)
  {
    // the reference clears m_voxel_raycast before the sensor-in-map test (:1430-1432)
    if (prc == VOFOD_W_SENSOR_OOB && ctx->acc_has_data && ctx->acc.p)
    {
      CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_cells_max * 8, ctx->stream));
      ctx->acc_has_data = false;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return prc;
  }
  memcpy(ctx->h_dyn->tf.R, tf->R, sizeof(tf->R));
  memcpy(ctx->h_dyn->tf.t, tf->t, sizeof(tf->t));
  ctx->h_dyn->scan = ctx->scan_staging.as<vofod_pt>();
  RET(vf_dyn_push(ctx));
  const int rc = vf_raycast_accumulate_dev(ctx, n, *tf, *p);
  if (rc < 0)
    return rc;
  unsigned long long t[2] = {0, 0};
  CK(cudaMemcpyAsync(&t[0], vf_cnt(ctx, CNT_TRAVERSALS), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&t[1], vf_cnt(ctx, CNT_OOB), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_traversals)
    *n_traversals = t[0];
  if (t[1])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "%llu traversals fell outside the accumulator window", t[1]);
  return rc;
}

int vofod_raycast_download(vofod_ctx* ctx, uint32_t* counts, float* lengths, size_t n_cells)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  const size_t n = (size_t)geom_cells(ctx->g);
  if (n_cells != n)
    return vf_fail(ctx, VOFOD_E_INVALID, "expected %zu cells", n);
  ENSURE(ctx->scratch_a, n * 4);
  ENSURE(ctx->scratch_b, n * 4);
  RET(vf_raycast_expand(ctx, counts ? ctx->scratch_a.as<uint32_t>() : nullptr, lengths ? ctx->scratch_b.as<float>() : nullptr));
  if (counts)
    CK(cudaMemcpyAsync(counts, ctx->scratch_a.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (lengths)
    CK(cudaMemcpyAsync(lengths, ctx->scratch_b.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
}

int vofod_raycast_apply(vofod_ctx* ctx, int its_diff, const vofod_params* p)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  if (!p || its_diff < 1)
    return vf_fail(ctx, VOFOD_E_INVALID, "bad argument (its_diff must be >= 1)");
  ctx->h_dyn->its_raycast = its_diff;
  RET(vf_dyn_push(ctx));
  const int rc = vf_raycast_apply_dev(ctx, its_diff, *p);
  if (rc != VOFOD_OK)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
  }
  unsigned long long any = 0;
  CK(cudaMemcpyAsync(&any, vf_cnt(ctx, CNT_APPLY_ANY), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (!any)
    return VOFOD_W_EMPTY_RAYCAST;
  ctx->flags_full_dirty = false;
  return VOFOD_OK;
}
}
