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
  float spread_len;  // aggregation ends this far along a ray [m]
};

__device__ __forceinline__ void red_add_u64(unsigned long long* addr, const unsigned long long v)
{
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// ---- per-ray DDA state (voxel_map.cpp:229-263) --------------------------------------------------------------------
struct RayState
{
  float len, prev;
  float tmx, tmy, tmz, tdx, tdy, tdz;
  int remx, remy, remz;  // steps left before the ray stands in the last voxel of the map along that axis (`cur[i] == last[i]`)
  int dwx, dwy, dwz;     // window-index stride of one step along each axis
  int widx;
  int spos, sstep;       // SLAB: window coordinate / step along the slab axis (the only axis a ray can leave the window on)
};

// instrumentation (VOFOD_OPT_RAYCAST_STATS): per warp-step histograms of the lanes in the loop and of the distinct voxels among them
#define RAY_STATS_SLOTS 72  // [0..32] lanes alive, [33..65] groups, [66] warp-steps of the fast loop, [67] of the general loop, [68] skipped steps
template <bool STATS>
__device__ __forceinline__ void ray_stats(unsigned long long* stats, const unsigned am, const bool leader, const int loop_slot)
{
  if (STATS)
  {
    const unsigned lead = __ballot_sync(am, leader);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(am) - 1))
    {
      atomicAdd(stats + __popc(am), 1ull);
      atomicAdd(stats + 33 + __popc(lead), 1ull);
      atomicAdd(stats + loop_slot, 1ull);
    }
  }
}

// One lane = one ray; the lanes that are in the loop together (__activemask) merge their updates: lanes standing in the same
// voxel issue ONE 64-bit RED (count in the top 20 bits, path length below).  No warp-wide vote per step: lanes whose ray has
// ended simply leave the loop, and if the hardware ever splits the warp the only effect is less merging — additions commute.
//  FAST  = every ray of the warp provably ends by its length before it can reach the last voxel of the map along any axis
//          (and, unsharded, stays inside the accumulator window): no countdown, no clamp.  With SLAB the ray also stops at the
//          first step that leaves the slab's window along the slab axis (it never comes back).
//  !FAST = the reference's loop with the `cur[i] == last[i]` countdown; with SLAB, steps outside the window are walked but not
//          written.
// EXP (VOFOD_OPT_RAYCAST_EXP, measurement only — results are wrong): 1 = everything but the RED itself, 2 = the DDA alone (no match / redux / RED)
template <bool AGG, bool SLAB, bool FAST, bool STATS, int EXP = 0>
__device__ __forceinline__ unsigned ray_loop(RayState& r, bool alive, const float scale, const int slab_axis, const int ssize, const int own_lo, const int own_n,
                                             const int wn, unsigned long long* __restrict__ acc, const unsigned lane, const unsigned lanemask_lt,
                                             unsigned long long* stats, uint8_t* __restrict__ touched)
{
  unsigned steps = 0;
  while (alive)
  {
    // tmax.minCoeff(&i): first minimum, strict '<'
    const bool use_y = r.tmy < r.tmx;
    const float d01 = use_y ? r.tmy : r.tmx;
    const bool use_z = r.tmz < d01;
    const float dist = use_z ? r.tmz : d01;
    const float ddist = (r.len < dist ? r.len : dist) - r.prev;                               // voxel_map.cpp:252
    const int q = __float2int_rn(ddist * scale);
    int key;
    bool inside = true;
    if (FAST)
      key = r.widx;  // inside the window by construction (see the classification in the kernel)
    else if (SLAB)
    {
      // the window is cut at the slab's storage box: rays enter and leave it along the slab axis
      inside = (unsigned)r.spos < (unsigned)ssize;
      key = inside ? r.widx : -1 - (int)lane;
    } else
      // the window holds every voxel within max_dist, so this clamp never bites; if it ever did, the update would land in the
      // spare cell behind the window (and the apply kernel reports it) instead of corrupting memory
      key = (int)min((unsigned)r.widx, (unsigned)wn);
    if (EXP == 2)
    {
      if (q == 0x7fffffff && key == -12345)  // keeps q and key alive
        steps += 1000u;
    } else if (AGG)
    {
      const unsigned am = __activemask();
      const unsigned m = __match_any_sync(am, key);
      const int sum = __reduce_add_sync(m, q);
      const bool leader = (m & lanemask_lt) == 0;
      if (EXP == 1)
      {
        if (leader && sum == 0x7fffffff && __popc(m) == 33)
          steps += 1000u;
      } else if (inside && leader)
      {
        red_add_u64(acc + key, ((unsigned long long)__popc(m) << ACC_LEN_BITS) + (unsigned long long)(long long)sum);
        if (touched)
          touched[key >> 5] = 1;  // plain store: every writer stores the same value
      }
      ray_stats<STATS>(stats, am, leader, FAST ? 66 : 67);
    } else if (inside)
    {
      red_add_u64(acc + key, (1ull << ACC_LEN_BITS) + (unsigned long long)(long long)q);
      if (touched)
        touched[key >> 5] = 1;
    }
    // SLAB: windows of neighbouring slabs overlap in the halos; a traversal is counted by the slab that OWNS the voxel, so that the
    // counts of all slabs add up to the reference's
    steps += SLAB ? ((unsigned)(r.spos - own_lo) < (unsigned)own_n ? 1u : 0u) : 1u;
    r.prev = dist;
    // voxel_map.cpp:257-261
    if (FAST)
      alive = dist < r.len;
    else
    {
      const int rem = use_z ? r.remz : (use_y ? r.remy : r.remx);
      alive = rem != 0 && dist < r.len;
    }
    if (use_z)
    {
      r.tmz += r.tdz; r.widx += r.dwz;
      if (!FAST) r.remz--;
      if (SLAB && slab_axis == 2) r.spos += r.sstep;
    } else if (use_y)
    {
      r.tmy += r.tdy; r.widx += r.dwy;
      if (!FAST) r.remy--;
      if (SLAB && slab_axis == 1) r.spos += r.sstep;
    } else
    {
      r.tmx += r.tdx; r.widx += r.dwx;
      if (!FAST) r.remx--;
      if (SLAB && slab_axis == 0) r.spos += r.sstep;
    }
    // a ray that has left the window along the slab axis never comes back (the general loop is entered fast-forwarded as well)
    if (SLAB && (FAST || slab_axis < 2))
      alive = alive && (unsigned)r.spos < (unsigned)ssize;
  }
  return steps;
}

// ---- ray_loop3 (VER 5): the fewest instructions per step ---------------------------------------------------------------------------------
// The kernel is bound by instruction issue (~65 % of the slots filled, every attempt to trade instructions for fewer warp collectives made it
// slower — profiles/r02_raycast_variants.json): predicated axis updates, no bookkeeping that is not needed in the instantiation (TOUCH, SLAB,
// FAST are template switches), one match.any + one redux + one RED block per step.  Aggregation ends by DISTANCE, at no cost per step: the loop
// runs while dist < stop; the caller runs it once with AGG up to `spread_len` along the ray and once without for what is left (rays of a warp
// are 3 mrad apart: past ~50 voxels most lanes stand alone, and a redux per lane costs more than the REDs it saves).
template <bool AGG, bool SLAB, bool FAST, bool TOUCH>
__device__ __forceinline__ unsigned ray_loop3(RayState& r, bool& alive, const float stop, const float scale, const int slab_axis, const int ssize, const int own_lo,
                                              const int own_n, const int wn, unsigned long long* __restrict__ acc, const unsigned lane, const unsigned lanemask_lt,
                                              uint8_t* __restrict__ touched)
{
  unsigned steps = 0;
  bool go = alive && r.prev < stop;
  while (go)
  {
    const float dist = fminf(fminf(r.tmx, r.tmy), r.tmz);
    const bool step_x = r.tmx == dist;                 // tmax.minCoeff(&i): the first minimum
    const bool step_y = !step_x && r.tmy == dist;
    const bool step_z = !step_x && !step_y;
    const float ddist = fminf(r.len, dist) - r.prev;   // voxel_map.cpp:252 (no NaNs here: fminf == the reference's ?:)
    const int q = __float2int_rn(ddist * scale);
    int key;
    bool inside = true;
    if (FAST && !SLAB)
      key = r.widx;
    else if (SLAB)
    {
      inside = (unsigned)r.spos < (unsigned)ssize;
      key = inside ? r.widx : -1 - (int)lane;
    } else
      key = (int)min((unsigned)r.widx, (unsigned)wn);
    if (AGG)
    {
      const unsigned am = __activemask();
      const unsigned m = __match_any_sync(am, key);
      const int sum = __reduce_add_sync(m, q);
      if (inside && (m & lanemask_lt) == 0)
      {
        // (q can be negative: a start point that rounding puts a hair outside its voxel gives a first tmax below zero)
        red_add_u64(acc + key, ((unsigned long long)__popc(m) << ACC_LEN_BITS) + (unsigned long long)(long long)sum);
        if (TOUCH)
          touched[key >> 5] = 1;
      }
    } else if (inside)
    {
      red_add_u64(acc + key, (1ull << ACC_LEN_BITS) + (unsigned long long)(long long)q);
      if (TOUCH)
        touched[key >> 5] = 1;
    }
    steps += SLAB ? ((unsigned)(r.spos - own_lo) < (unsigned)own_n ? 1u : 0u) : 1u;
    r.prev = dist;
    // voxel_map.cpp:257-261
    alive = dist < r.len;
    if (!FAST)
    {
      const int rem = step_z ? r.remz : (step_y ? r.remy : r.remx);
      alive = alive && rem != 0;
    }
    if (step_x) r.tmx += r.tdx;
    if (step_y) r.tmy += r.tdy;
    if (step_z) r.tmz += r.tdz;
    r.widx += step_x ? r.dwx : (step_y ? r.dwy : r.dwz);
    if (!FAST)
    {
      if (step_x) r.remx--;
      if (step_y) r.remy--;
      if (step_z) r.remz--;
    }
    if (SLAB)
    {
      if (slab_axis == 0 ? step_x : (slab_axis == 1 ? step_y : step_z))
        r.spos += r.sstep;
      // a ray that has left the window along the slab axis never comes back (the general loop is entered fast-forwarded as well)
      if (FAST || slab_axis < 2)
        alive = alive && (unsigned)r.spos < (unsigned)ssize;
    }
    go = alive && dist < stop;
  }
  return steps;
}

// SLAB: bring the DDA of a ray that starts outside the slab's window to the state it has when it enters, WITHOUT walking the voxels
// in between.  The DDA is a 3-way merge of the increasing sequences tm_a + j*td_a (each built by the same chain of fp32 additions the
// loop performs) with ties going to x before y before z (first minimum).  The ka-th step along the slab axis consumes T = tm_a after
// ka-1 additions; a step along another axis b comes before it iff its value is < T, or == T when b has priority over the slab axis.
// The state after that step is therefore: tm_a = T + td_a, prev = T, every other axis advanced past T.  Bit-identical to walking (same
// additions in the same order per axis); the ray enters iff T < len — and, for a ray that can reach the edge of the MAP (!FAST), iff no
// axis runs out of voxels on the way: an axis that needs more steps than it has left (`cur[i] == last[i]`, voxel_map.cpp:257) ends the
// ray before it gets here.
// Returns false when the ray never enters the window.
template <bool FAST>
__device__ __forceinline__ bool ray_skip_to_slab(RayState& r, const int slab_axis, const int ssize, unsigned& skipped)
{
  if ((unsigned)r.spos < (unsigned)ssize)
    return true;
  const int ka = r.spos < 0 ? (r.sstep > 0 ? -r.spos : 0) : (r.sstep < 0 ? r.spos - (ssize - 1) : 0);
  if (ka <= 0)
    return false;  // parallel to the slab or heading away from it
  float& tma = slab_axis == 0 ? r.tmx : r.tmy;
  const float tda = slab_axis == 0 ? r.tdx : r.tdy;
  // the slab axis takes at most len/td_a + 1 steps before the ray ends
  if ((float)(ka - 2) * tda >= r.len)
    return false;
  if (!FAST)
  {
    int& rema = slab_axis == 0 ? r.remx : r.remy;
    if (rema < ka)
      return false;
    rema -= ka;
  }
  float T = tma;
  for (int j = 1; j < ka; j++)
    T += tda;
  if (!(T < r.len))
    return false;
  tma = T + tda;
  r.prev = T;
  r.spos += ka * r.sstep;
  r.widx += ka * (slab_axis == 0 ? r.dwx : r.dwy);
  skipped += (unsigned)ka;
  int nb = 0;
  if (slab_axis == 0)
  {
    while (r.tmy < T) { r.tmy += r.tdy; r.widx += r.dwy; nb++; }   // x has priority over y and z on ties
    if (!FAST) { if (r.remy < nb) return false; r.remy -= nb; }
  } else
  {
    while (r.tmx <= T) { r.tmx += r.tdx; r.widx += r.dwx; nb++; }  // x has priority over y
    if (!FAST) { if (r.remx < nb) return false; r.remx -= nb; }
  }
  skipped += (unsigned)nb;
  nb = 0;
  while (r.tmz < T) { r.tmz += r.tdz; r.widx += r.dwz; nb++; }
  if (!FAST) { if (r.remz < nb) return false; r.remz -= nb; }
  skipped += (unsigned)nb;
  return true;
}

// RB = rays (threads) per block.  Ray lengths differ a lot between LiDAR rows (no-return rays walk max_dist, ground
// returns a few metres), so small blocks balance better: 64 threads = 2 warps = 64 neighbouring columns of one row.
// VER 5 = ray_loop3 in 32 registers (2048 threads per SM: every ray of a 128 x 2048 scan is resident at once — with 56 registers only 36 of the
// 55 warps an SM receives were, and the rest ran as a second, half-empty wave); VER 0 = round 1's loop and register budget (statistics, A/B runs).
template <int RB, bool AGG, bool SLAB, bool STATS, int EXP = 0, int VER = 5, bool TOUCH = false>
__global__ void __launch_bounds__(RB, VER == 0 ? 0 : 2048 / RB) k_raycast_accumulate(const RayArgs a, const ScanDyn* __restrict__ dyn, const float4* __restrict__ lut_dir,
                                                           const float4* __restrict__ lut_off, const uint8_t* __restrict__ mask,
                                                           unsigned long long* __restrict__ acc, unsigned long long* __restrict__ counters,
                                                           unsigned long long* __restrict__ stats, uint8_t* __restrict__ touched)
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
  const unsigned lanemask_lt = (1u << lane) - 1u;
  bool alive = idx < a.n;
  bool safe = true;
  RayState r = {};
  const int slab_axis = a.g.slab_axis;
  const int ssize = w.size[slab_axis];
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
    r.len = ray_dist == 0.0f ? a.max_dist : (a.max_dist < dmv ? a.max_dist : dmv);            // :1457 (std::min)
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
    r.tdx = (1.0f / ax) * a.g.vs;
    r.tdy = (1.0f / ay) * a.g.vs;
    r.tdz = (1.0f / az) * a.g.vs;
    const float ox = idx_to_coord1(cx, a.g.off[0], a.g.vs) - sx;
    const float oy = idx_to_coord1(cy, a.g.off[1], a.g.vs) - sy;
    const float oz = idx_to_coord1(cz, a.g.off[2], a.g.vs) - sz;
    r.tmx = (a.g.half + (float)stx * ox) / ax;
    r.tmy = (a.g.half + (float)sty * oy) / ay;
    r.tmz = (a.g.half + (float)stz * oz) / az;
    // `if (cur[i] == last[i]) break` with last = step > 0 ? size-1 : 0 (voxel_map.cpp:240-257), as a countdown
    r.remx = stx > 0 ? a.g.size[0] - 1 - cx : cx;
    r.remy = sty > 0 ? a.g.size[1] - 1 - cy : cy;
    r.remz = stz > 0 ? a.g.size[2] - 1 - cz : cz;
    r.dwx = stx;
    r.dwy = sty * w.size[0];
    r.dwz = stz * w.size[0] * w.size[1];
    r.widx = (cx - w.lo[0]) + (cy - w.lo[1]) * w.size[0] + (cz - w.lo[2]) * w.size[0] * w.size[1];
    if (SLAB)
    {
      r.spos = slab_axis == 0 ? cx - w.lo[0] : (slab_axis == 1 ? cy - w.lo[1] : cz - w.lo[2]);
      r.sstep = slab_axis == 0 ? stx : (slab_axis == 1 ? sty : stz);
    }
    if (!(0.0f < r.len))  // while (prev_dist < length) with prev_dist = 0
      alive = false;
    // An axis steps when its tmax (>= j * td, up to rounding) is below len: at most len*|d|/vs + 1 times.  With more than that
    // (+1 for the rounding) to go before the last voxel of the map, the countdown of the reference's loop can never trigger.
    const float li = r.len * a.g.inv;
    safe = r.remx > (int)(li * ax) + 2 && r.remy > (int)(li * ay) + 2 && r.remz > (int)(li * az) + 2;
    // (such a ray also stays inside the accumulator window: vf_raycast_prepare sizes it ceil(reach/vs) + 2 cells around the sensor
    //  voxel, clamped to the held box — and a safe ray does not leave the map)
  }
  unsigned steps = 0, skipped = 0;
  const int own_lo = a.g.own_lo - w.lo[slab_axis], own_n = a.g.own_hi - a.g.own_lo;  // owned range in window coordinates along the slab axis
  // SLAB: max_element(raycast) > 0 (:1542-1548) is a property of the WHOLE map, and a slab only sees its part of the rays' voxels.  Every
  // slab runs this set-up for every ray, so "some ray is cast" is known to all of them without an exchange; it differs from the
  // reference's test only when every cast ray is shorter than one quantum of the fixed-point length (2^-F m).
  if (SLAB && __any_sync(VOFOD_FULL, alive) && lane == 0)
    atomicOr(counters + CNT_APPLY_ANY, 1ull);
  // one decision per warp, so that the lanes of a warp stay in the same loop and keep merging
  const bool fast = __all_sync(VOFOD_FULL, safe || !alive) && (!SLAB || slab_axis < 2);
  if (fast)
  {
    if (SLAB && alive)
      alive = ray_skip_to_slab<true>(r, slab_axis, ssize, skipped);
    if (VER == 0)
      steps = ray_loop<AGG, SLAB, true, STATS, EXP>(r, alive, a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, stats, touched);
    else
    {
      // merged up to spread_len along the ray, lane by lane beyond
      steps = ray_loop3<true, SLAB, true, TOUCH>(r, alive, a.spread_len, a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, touched);
      steps += ray_loop3<false, SLAB, true, TOUCH>(r, alive, __int_as_float(0x7f800000), a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, touched);
    }
  } else
  {
    if (SLAB && slab_axis < 2 && alive)
      alive = ray_skip_to_slab<false>(r, slab_axis, ssize, skipped);
    if (VER == 0)
      steps = ray_loop<AGG, SLAB, false, STATS, EXP>(r, alive, a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, stats, touched);
    else
    {
      // merged up to spread_len along the ray, lane by lane beyond
      steps = ray_loop3<true, SLAB, false, TOUCH>(r, alive, a.spread_len, a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, touched);
      steps += ray_loop3<false, SLAB, false, TOUCH>(r, alive, __int_as_float(0x7f800000), a.scale, slab_axis, ssize, own_lo, own_n, wn, acc, lane, lanemask_lt, touched);
    }
  }
  __syncwarp();
  if (STATS && skipped)
    atomicAdd(stats + 68, (unsigned long long)skipped);
  // per-block totals
  unsigned tot = prims::warp_sum(steps);
  __shared__ unsigned s_tot[RB / 32];
  if (lane == 0)
    s_tot[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned t = 0;
    for (int i = 0; i < RB / 32; i++)
      t += s_tot[i];
    if (t)
      atomicAdd(counters + CNT_TRAVERSALS, (unsigned long long)t);
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
  const Window w = after_wait(dyn)->win_apply;
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
                                                       unsigned long long* __restrict__ counters, uint8_t* __restrict__ touched)
{
  pdl_enter();
  const Window w = dyn->win_apply;
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
  bool any_pos = false;  // max_element(raycast) > 0 (:1542-1548): some cell holds a positive length
  // Every thread keeps APPLY_U cells in flight, stage by stage (accumulator words, then flags, then scores): with one cell per
  // iteration the pass is a chain of three dependent memory round trips per cell and runs at a third of the bandwidth.
  // `touched` != NULL (large windows): the accumulate kernel marked every group of 32 cells it added to; a warp fetches 32 marks at a
  // time and visits the marked groups only, APPLY_U groups per round.  Otherwise: every cell of the window.
  constexpr int APPLY_U = 4;
  const unsigned lane = threadIdx.x & 31;
  const long long n_groups = (n + 31) >> 5;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const unsigned wsxy = (unsigned)wsx * (unsigned)wsy;
  const bool idx32 = n < (1ll << 32);
  long long gbase = touched ? warp0 * 32 : 0;
  unsigned pending = 0;
  long long i_next = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  while (true)
  {
    long long idx[APPLY_U];
    if (touched)
    {
      bool have_first = false;  // warp-uniform: slot 0 got a group (slots are filled in order: none for slot 0 = nothing left)
#pragma unroll
      for (int k = 0; k < APPLY_U; k++)
      {
        // next marked group of this warp (all decisions are warp-uniform)
        while (pending == 0 && gbase < n_groups)
        {
          const long long g = gbase + lane;
          const bool mark = g < n_groups && touched[g] != 0;
          pending = __ballot_sync(VOFOD_FULL, mark);
          if (mark)
            touched[g] = 0;
          if (pending == 0)
            gbase += n_warps * 32;
        }
        idx[k] = -1;
        if (pending != 0)
        {
          const int bit = __ffs(pending) - 1;
          pending &= pending - 1;
          const long long c = ((gbase + bit) << 5) + lane;
          if (c < n)
            idx[k] = c;
          if (k == 0)
            have_first = true;
          if (pending == 0)
            gbase += n_warps * 32;
        }
      }
      if (!have_first)
        break;
    } else
    {
      if (i_next >= n)
        break;
#pragma unroll
      for (int k = 0; k < APPLY_U; k++)
      {
        idx[k] = i_next < n ? i_next : -1;
        i_next += stride;
      }
    }
    // stage 1: accumulator words
    unsigned long long pw[APPLY_U];
#pragma unroll
    for (int k = 0; k < APPLY_U; k++)
      pw[k] = idx[k] >= 0 ? acc[idx[k]] : 0ull;
    // stage 2: decode, window -> grid index
    float rv[APPLY_U];
    long long ci[APPLY_U];
#pragma unroll
    for (int k = 0; k < APPLY_U; k++)
    {
      rv[k] = 0.f;
      ci[k] = -1;
      if (pw[k])
      {
        acc[idx[k]] = 0ull;  // m_voxel_raycast.clear() for the next scan (:1430)
        unsigned c;
        long long lq;
        acc_decode(pw[k], c, lq);
        rv[k] = (float)((double)lq * a.inv_scale);
        if (rv[k] > 0.0f)
        {
          any_pos = true;
          int wx, wy, wz;
          if (idx32)
          {
            const unsigned u = (unsigned)idx[k];
            wz = (int)(u / wsxy);
            const unsigned r = u - (unsigned)wz * wsxy;
            wy = (int)(r / (unsigned)wsx);
            wx = (int)(r - (unsigned)wy * (unsigned)wsx);
          } else
          {
            wx = (int)(idx[k] % wsx);
            wy = (int)((idx[k] / wsx) % wsy);
            wz = (int)(idx[k] / ((long long)wsx * wsy));
          }
          ci[k] = cell_index(a.g, wx + w.lo[0], wy + w.lo[1], wz + w.lo[2]);
        }
      }
    }
    // stage 3: flags (flag == m_vflags_unmarked, :1561), stage 4: scores
    uint8_t fl[APPLY_U];
#pragma unroll
    for (int k = 0; k < APPLY_U; k++)
      fl[k] = ci[k] >= 0 ? flags[ci[k]] : (uint8_t)1;
    float mv[APPLY_U];
#pragma unroll
    for (int k = 0; k < APPLY_U; k++)
      mv[k] = fl[k] == 0 ? score[ci[k]] : 0.f;
#pragma unroll
    for (int k = 0; k < APPLY_U; k++)
    {
      if (fl[k] != 0)
        continue;
      float w1;
      if (a.new_rule)
      {
        const float n_int = a.weighting_factor * rv[k];                 // :1565
        w1 = (float)exp2((double)(-its * n_int));                     // :1567  std::pow(2, float) -> double pow
      } else
      {
        const float norm_val = rv[k] / max_val;                         // :1587
        const float ws = a.ray_weight * sqrtf(norm_val);                // :1591
        // :1593 std::pow(float, float).  CUDA's powf is a few ulp off; the fp64 pow rounded to fp32 reproduces the host's
        // (correctly rounded) powf except in vanishingly rare double-rounding cases
        w1 = (float)pow((double)(1.0f - ws), (double)its);
        w1 = w1 < 0.0f ? 0.0f : (1.0f < w1 ? 1.0f : w1);                // std::clamp
      }
      const float w2 = 1.0f - w1;
      score[ci[k]] = w1 * mv[k] + w2 * a.ray_score;                     // :1569 / :1597
    }
  }
  __syncwarp();
  if (__any_sync(VOFOD_FULL, any_pos) && (threadIdx.x & 31) == 0)
    atomicOr(counters + CNT_APPLY_ANY, 1ull);
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
  const size_t dirty_off = (wn_max + 32) * 8, total = dirty_off + wn_max / 32 + 64;
  if (ctx->acc.cap < total)
  {
    ENSURE(ctx->acc, total);  // fresh allocations are zero-filled
    ctx->acc_has_data = false;
  }
  if (ctx->acc_cells_max != wn_max && ctx->acc_total_bytes && (ctx->acc_has_data || ctx->acc_cells_max == 0))
  {
    // the layout changes with max_dist (dynamic_reconfigure): start from a clean buffer
    CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_total_bytes > total ? ctx->acc_total_bytes : total, ctx->stream));
    ctx->acc_has_data = false;
  }
  ctx->acc_dirty_off = dirty_off;
  ctx->acc_total_bytes = total;
  ctx->acc_sparse = ctx->acc_sparse_mode == 1 || (ctx->acc_sparse_mode == 0 && wn_max >= (size_t(1) << 25));
  ctx->acc_cells_max = wn_max;
  ctx->win = w;
  ctx->win_valid = true;
  ctx->h_dyn->win = w;
  ctx->h_dyn->win_apply = w;  // (a call that applies an earlier, deferred accumulate overrides it)
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
    CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_total_bytes, ctx->stream));
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
  uint8_t* touched = ctx->acc_sparse ? ctx->acc.as<uint8_t>() + ctx->acc_dirty_off : nullptr;
  unsigned long long* stats = nullptr;
  if (ctx->raycast_stats)
  {
    ENSURE(ctx->ray_stats, RAY_STATS_SLOTS * 8);
    stats = ctx->ray_stats.as<unsigned long long>();
  }
  // VOFOD_OPT_RAYCAST_NO_AGG: no merging at all = merging that ends at distance 0
  a.spread_len = ctx->raycast_no_agg ? 0.0f : (float)ctx->raycast_spread_voxels * ctx->g.vs;
#define RAY_ARGS a, ctx->dyn.as<ScanDyn>(), ctx->lut_dir.as<float4>(), ctx->lut_off.as<float4>(), ctx->mask.as<uint8_t>(), ctx->acc.as<unsigned long long>(), cnt, stats, touched
#define RAY_LAUNCH(RB_, SLAB_, TOUCH_) LAUNCH((k_raycast_accumulate<RB_, true, SLAB_, false, 0, 5, TOUCH_>), (int)((n + RB_ - 1) / RB_), RB_, 0, RAY_ARGS)
#define RAY_OLD(SLAB_, STATS_, EXP_) LAUNCH((k_raycast_accumulate<64, true, SLAB_, STATS_, EXP_, 0>), (int)((n + 63) / 64), 64, 0, RAY_ARGS)
  if (ctx->raycast_exp == 1 && !slab)
    RAY_OLD(false, false, 1);
  else if (ctx->raycast_exp == 2 && !slab)
    RAY_OLD(false, false, 2);
  else if (ctx->raycast_exp == 3)  // round 1's loop, for A/B runs (results are correct)
  {
    if (slab) RAY_OLD(true, false, 0); else RAY_OLD(false, false, 0);
  } else if (stats)
  {
    if (slab) RAY_OLD(true, true, 0); else RAY_OLD(false, true, 0);
  } else if (slab)
  {
    if (touched) RAY_LAUNCH(64, true, true); else RAY_LAUNCH(64, true, false);
  } else if (touched)
    RAY_LAUNCH(64, false, true);
  else if (ctx->raycast_block == 128)
    RAY_LAUNCH(128, false, false);
  else if (ctx->raycast_block == 256)
    RAY_LAUNCH(256, false, false);
  else
    RAY_LAUNCH(64, false, false);
#undef RAY_LAUNCH
#undef RAY_OLD
#undef RAY_ARGS
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
  // (old rule: when the apply bails out on max_val == 0 the marks stay set and the next accumulate's clear removes them with the cells)
  uint8_t* touched = ctx->acc_sparse ? ctx->acc.as<uint8_t>() + ctx->acc_dirty_off : nullptr;
  LAUNCH(k_raycast_apply, vf_blocks(ctx, touched ? wn / 8 : wn, 256), 256, 0, a, dyn, ctx->acc.as<unsigned long long>(), ctx->score.as<float>(), ctx->flags.as<uint8_t>(),
         (const unsigned*)(cnt + CNT_MAXVAL), cnt, touched);
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
      CK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc_total_bytes, ctx->stream));
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

int vofod_raycast_stats(vofod_ctx* ctx, uint64_t* out, size_t n)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!out || n > RAY_STATS_SLOTS)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_raycast_stats: bad arguments");
  memset(out, 0, n * 8);
  if (!ctx->ray_stats.p)
    return VOFOD_OK;
  CK(cudaMemcpyAsync(out, ctx->ray_stats.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemsetAsync(ctx->ray_stats.p, 0, RAY_STATS_SLOTS * 8, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VOFOD_OK;
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
  FLUSH_PENDING();
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
