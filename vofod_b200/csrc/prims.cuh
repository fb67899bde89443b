// Device-wide primitives written for this library (no CUB/Thrust on the product path):
//   * single-pass exclusive scan with decoupled look-back
//   * Onesweep-style LSD radix sort (one histogram sweep + one chained-scan scatter kernel per 8-bit digit),
//     keys u32/u64, optional u32 payload, stable
// All kernels are persistent (grid <= co-resident capacity), take their element count from DEVICE memory so
// a whole scan can be enqueued without a host round trip, and every spin is bounded by a watchdog.
#pragma once
#include "common.cuh"

namespace prims
{
constexpr int NT = 256;           // threads per block
constexpr int IPT = 8;            // items per thread
constexpr int TILE = NT * IPT;    // 2048
constexpr int SPIN_LIMIT = 1 << 22;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

template <class T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(VOFOD_FULL, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t t = __shfl_up_sync(VOFOD_FULL, v, o);
    if (lane_id() >= (unsigned)o)
      v += t;
  }
  return v;
}
// exclusive scan over the block (NT threads); `ws` = NT/32 words of shared memory; returns exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_excl_scan(const uint32_t v, uint32_t* ws, uint32_t& total)
{
  const uint32_t inc = warp_incl_scan(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if (lane_id() == 31)
    ws[w] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NT / 32; i++)
  {
    const uint32_t s = ws[i];
    if (i < w)
      base += s;
    tot += s;
  }
  total = tot;
  return base + inc - v;
}

// ---- decoupled look-back -------------------------------------------------------------------------------
// state word: [63:34] epoch | [33:32] status (1 = aggregate, 2 = inclusive prefix) | [31:0] value
__device__ __forceinline__ unsigned long long st_pack(const uint32_t epoch, const unsigned long long status, const uint32_t value)
{
  return ((unsigned long long)epoch << 34) | (status << 32) | (unsigned long long)value;
}
__device__ __forceinline__ void st_store(unsigned long long* p, const unsigned long long v) { *(volatile unsigned long long*)p = v; }
__device__ __forceinline__ unsigned long long st_load(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }

// one thread per chain; `chain` entries are `stride` apart.  Returns the exclusive prefix of `tile`.
__device__ inline uint32_t lookback(unsigned long long* chain, const int stride, const int tile, const uint32_t aggregate, const uint32_t epoch,
                                    unsigned long long* watchdog)
{
  if (tile == 0)
  {
    st_store(chain, st_pack(epoch, 2ull, aggregate));
    return 0u;
  }
  st_store(chain + (size_t)tile * stride, st_pack(epoch, 1ull, aggregate));
  uint32_t excl = 0;
  int t = tile - 1;
  int spins = 0;
  while (t >= 0)
  {
    const unsigned long long s = st_load(chain + (size_t)t * stride);
    const unsigned status = (unsigned)((s >> 32) & 3ull);
    if ((uint32_t)(s >> 34) == (epoch & 0x3fffffffu) && status != 0u)
    {
      excl += (uint32_t)s;
      if (status == 2u)
        break;
      t--;
    } else
    {
      if (++spins > SPIN_LIMIT)
      {
        atomicAdd(watchdog, 1ull);
        break;
      }
      __nanosleep(32);
    }
  }
  st_store(chain + (size_t)tile * stride, st_pack(epoch, 2ull, excl + aggregate));
  return excl;
}

// the same for ONE chain, walked by a whole warp: lane l inspects the l-th predecessor, so a chain of T tiles resolves in
// ~T/32 round trips instead of T (the serial walk made a 1M-element scan cost 16 us).  All 32 lanes must call.
__device__ inline uint32_t lookback_warp(unsigned long long* chain, const int tile, const uint32_t aggregate, const uint32_t epoch, unsigned long long* watchdog)
{
  const unsigned lane = lane_id();
  if (tile == 0)
  {
    if (lane == 0)
      st_store(chain, st_pack(epoch, 2ull, aggregate));
    return 0u;
  }
  if (lane == 0)
    st_store(chain + tile, st_pack(epoch, 1ull, aggregate));
  uint32_t excl = 0;
  int base = tile - 1;
  int spins = 0;
  while (true)
  {
    const int t = base - (int)lane;
    unsigned status = 2u;  // before tile 0: an inclusive prefix of 0
    uint32_t val = 0u;
    if (t >= 0)
    {
      const unsigned long long s = st_load(chain + t);
      status = (uint32_t)(s >> 34) == (epoch & 0x3fffffffu) ? (unsigned)((s >> 32) & 3ull) : 0u;
      val = (uint32_t)s;
    }
    const unsigned not_ready = __ballot_sync(VOFOD_FULL, status == 0u);
    const unsigned inclusive = __ballot_sync(VOFOD_FULL, status == 2u);
    const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
    const int first_in = inclusive ? __ffs(inclusive) - 1 : 32;
    const int take = first_in < first_nr ? first_in + 1 : first_nr;  // usable predecessors: lanes [0, take)
    excl += __reduce_add_sync(VOFOD_FULL, (int)lane < take ? val : 0u);
    if (first_in < first_nr)
      break;
    base -= take;
    if (take == 0)
    {
      if (++spins > SPIN_LIMIT)
      {
        if (lane == 0)
          atomicAdd(watchdog, 1ull);
        break;
      }
      __nanosleep(32);
    }
  }
  if (lane == 0)
    st_store(chain + tile, st_pack(epoch, 2ull, excl + aggregate));
  return excl;
}

__device__ __forceinline__ size_t dev_count(const unsigned long long* d_n, const size_t cap)
{
  if (!d_n)
    return cap;
  const unsigned long long n = *after_wait(d_n);
  return n < cap ? (size_t)n : cap;
}

// ---- exclusive scan of u32 ------------------------------------------------------------------------------
// out[i] = sum_{j<i} f(in[j]); *total (u64, may be null) = sum of all.  in/out may alias.  Buffers must be padded to TILE.
// f = identity, or popcount (the input then is a bit mask per element and the scan ranks the set bits).
// Up to two independent scans ride in one launch (blockIdx.y selects the job; each has its own look-back chain).
struct ScanJob
{
  const uint32_t* in;
  uint32_t* out;
  const unsigned long long* d_n;  // element count on the device (clamped to cap), or NULL = cap
  size_t cap;
  unsigned long long* state;
  unsigned long long* total;
  int popc;
  int sparse_out;  // 1: the caller only reads out[i] where in[i] != 0 — groups of four zero inputs are not stored
  // optional: every non-zero element is appended (in no particular order) as (index, raw value, exclusive prefix, 0)
  uint4* list;
  unsigned long long* list_n;
};
struct ScanJobs
{
  ScanJob j[2];
};
// SIPT = items per thread: 8 (2048-element tiles) for short inputs, 32 (8192) for long ones — the look-back chain of a long
// input is what its scan waits for, and it is 4x shorter with the larger tile.
constexpr int SCAN_BIG_IPT = 32;
constexpr int SCAN_BIG_TILE = NT * SCAN_BIG_IPT;  // buffers are padded to this (padded())
template <int SIPT>
static __global__ void __launch_bounds__(NT) k_scan_excl_u32(const ScanJobs jobs, const unsigned long long* __restrict__ epoch_base, const uint32_t epoch_local,
                                                             unsigned long long* watchdog)
{
  pdl_enter();
  constexpr int STILE = NT * SIPT;
  __shared__ uint32_t ws[NT / 32];
  __shared__ uint32_t s_base;
  const ScanJob& J = jobs.j[blockIdx.y];
  const uint32_t* __restrict__ in = J.in;
  uint32_t* __restrict__ out = J.out;
  unsigned long long* total = J.total;
  const uint32_t epoch = (uint32_t)(*epoch_base + epoch_local) & 0x3fffffffu;
  const size_t n = dev_count(J.d_n, J.cap);
  const int n_tiles = (int)((n + STILE - 1) / STILE);
  if (n_tiles == 0)
  {
    if (blockIdx.x == 0 && threadIdx.x == 0 && total)
      *total = 0ull;
    return;
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const size_t base_i = (size_t)tile * STILE + (size_t)threadIdx.x * SIPT;
    uint32_t v[SIPT];
#pragma unroll
    for (int q = 0; q < SIPT / 4; q++)
    {
      const uint4 a = *reinterpret_cast<const uint4*>(in + base_i + 4 * q);
      v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
    uint32_t raw_nz = 0;  // which of this thread's elements are non-zero
    uint32_t raw[SIPT];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SIPT; k++)
    {
      raw[k] = v[k];
      if (v[k] != 0u && base_i + k < n)
        raw_nz |= 1u << k;
      if (J.popc)
        v[k] = (uint32_t)__popc(v[k]);
      if (base_i + k >= n)
        v[k] = 0;
      tsum += v[k];
    }
    uint32_t agg;
    const uint32_t texcl = block_excl_scan(tsum, ws, agg);
    if (threadIdx.x < 32)
    {
      const uint32_t e = lookback_warp(J.state, tile, agg, epoch, watchdog);
      if (threadIdx.x == 0)
      {
        s_base = e;
        if (tile == n_tiles - 1 && total)
          *total = (unsigned long long)e + agg;
      }
    }
    __syncthreads();
    uint32_t run = s_base + texcl;
    if (J.list && raw_nz)
    {
      unsigned long long at = atomicAdd(J.list_n, (unsigned long long)__popc(raw_nz));
      uint32_t r2 = run;
#pragma unroll
      for (int k = 0; k < SIPT; k++)
      {
        if ((raw_nz >> k) & 1u)
          J.list[at++] = make_uint4((uint32_t)(base_i + k), raw[k], r2, 0u);
        r2 += v[k];
      }
    }
#pragma unroll
    for (int q = 0; q < SIPT / 4; q++)
    {
      uint4 o;
      o.x = run; run += v[4 * q];
      o.y = run; run += v[4 * q + 1];
      o.z = run; run += v[4 * q + 2];
      o.w = run; run += v[4 * q + 3];
      if (out && !(J.sparse_out && (v[4 * q] | v[4 * q + 1] | v[4 * q + 2] | v[4 * q + 3]) == 0u))  // (NULL: only list and total wanted)
        *reinterpret_cast<uint4*>(out + base_i + 4 * q) = o;
    }
    __syncthreads();
  }
}

// ---- radix sort --------------------------------------------------------------------------------------------
template <class KeyT>
__global__ void __launch_bounds__(NT) k_radix_hist(const KeyT* __restrict__ keys, const unsigned long long* d_n, const size_t cap, uint32_t* __restrict__ hist,
                                                   const int begin_bit, const int passes)
{
  pdl_enter();
  __shared__ uint32_t sh[8 * 256];
  for (int i = threadIdx.x; i < passes * 256; i += NT)
    sh[i] = 0;
  __syncthreads();
  const size_t n = dev_count(d_n, cap);
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT)
  {
    const KeyT k = keys[i];
    for (int p = 0; p < passes; p++)
      atomicAdd(&sh[p * 256 + (int)((k >> (begin_bit + 8 * p)) & 0xFF)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += NT)
    if (sh[i])
      atomicAdd(&hist[i], sh[i]);
}

template <class KeyT, bool HAS_VAL>
__global__ void __launch_bounds__(NT) k_radix_pass(const KeyT* __restrict__ kin, KeyT* __restrict__ kout, const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                   const unsigned long long* d_n, const size_t cap, const uint32_t* __restrict__ hist_pass,
                                                   unsigned long long* state, const unsigned long long* __restrict__ epoch_base, const uint32_t epoch_local, const int shift,
                                                   unsigned long long* watchdog)
{
  pdl_enter();
  const uint32_t epoch = (uint32_t)(*epoch_base + epoch_local) & 0x3fffffffu;
  __shared__ uint32_t wcnt[NT / 32][256];
  __shared__ uint32_t gbase[256];
  __shared__ uint32_t dbase[256];
  __shared__ uint32_t ws[NT / 32];
  const size_t n = dev_count(d_n, cap);
  const int n_tiles = (int)((n + TILE - 1) / TILE);
  if (n_tiles == 0)
    return;
  {
    uint32_t tot;
    gbase[threadIdx.x] = block_excl_scan(hist_pass[threadIdx.x], ws, tot);
  }
  // a digit that has the same value in every key (the high digit of keys that do not use their full bit budget) leaves the
  // order unchanged: plain copy, no ranking and no look-back chain
  {
    __shared__ int s_const;
    if (threadIdx.x == 0)
      s_const = 0;
    __syncthreads();
    if ((size_t)hist_pass[threadIdx.x] == n)
      s_const = 1;
    __syncthreads();
    if (s_const)
    {
      for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT)
      {
        kout[i] = kin[i];
        if (HAS_VAL)
          vout[i] = vin[i];
      }
      return;
    }
  }
  const int w = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
#pragma unroll
    for (int i = 0; i < NT / 32; i++)
      wcnt[i][threadIdx.x] = 0;
    __syncthreads();
    KeyT key[IPT];
    uint32_t val[IPT];
    uint32_t rank[IPT];
    const size_t wbase = (size_t)tile * TILE + (size_t)w * (32 * IPT);
#pragma unroll
    for (int k = 0; k < IPT; k++)
    {
      const size_t i = wbase + k * 32 + lane;
      const bool valid = i < n;
      key[k] = valid ? kin[i] : (KeyT)0;
      if (HAS_VAL)
        val[k] = valid ? vin[i] : 0u;
      const unsigned d = valid ? (unsigned)((key[k] >> shift) & 0xFF) : (256u + lane);
      const unsigned m = __match_any_sync(VOFOD_FULL, d);
      uint32_t r = 0;
      if (valid)
        r = wcnt[w][d] + __popc(m & lanemask_lt());
      __syncwarp();
      if (valid && lane == (unsigned)(__ffs(m) - 1))
        wcnt[w][d] += __popc(m);
      __syncwarp();
      rank[k] = r;
    }
    __syncthreads();
    {
      const int d = threadIdx.x;
      uint32_t running = 0;
#pragma unroll
      for (int i = 0; i < NT / 32; i++)
      {
        const uint32_t c = wcnt[i][d];
        wcnt[i][d] = running;
        running += c;
      }
      const uint32_t excl = lookback(state + d, 256, tile, running, epoch, watchdog);
      dbase[d] = gbase[d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < IPT; k++)
    {
      const size_t i = wbase + k * 32 + lane;
      if (i < n)
      {
        const unsigned d = (unsigned)((key[k] >> shift) & 0xFF);
        const uint32_t pos = dbase[d] + wcnt[w][d] + rank[k];
        kout[pos] = key[k];
        if (HAS_VAL)
          vout[pos] = val[k];
      }
    }
    __syncthreads();
  }
}

static inline int persistent_grid(const vofod_ctx* c, const size_t cap_items)
{
  size_t tiles = (cap_items + TILE - 1) / TILE;
  if (tiles < 1)
    tiles = 1;
  const size_t g = (size_t)c->num_sms * 2;  // 2 x 256-thread CTAs per SM are always co-resident for these kernels
  return (int)(tiles < g ? tiles : g);
}
static inline size_t padded(const size_t n) { return ((n + SCAN_BIG_TILE - 1) / SCAN_BIG_TILE + 1) * SCAN_BIG_TILE; }

// host drivers ------------------------------------------------------------------------------------------------
// one tile per CTA whenever the CTAs fit the machine at once: a CTA that loops over tiles pays the full load -> look-back ->
// store latency once per tile, back to back (measured: 18 us for a 1M-element scan with 2 CTAs per SM)
template <int SIPT>
static inline int scan_grid(const vofod_ctx* c, const size_t cap_items)
{
  static int occ = 0;
  if (occ == 0)
  {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_scan_excl_u32<SIPT>, NT, 0) != cudaSuccess || o < 1)
      o = 2;
    occ = o;
  }
  const size_t tile = (size_t)NT * SIPT;
  size_t tiles = (cap_items + tile - 1) / tile;
  if (tiles < 1)
    tiles = 1;
  const size_t g = (size_t)c->num_sms * occ;
  return (int)(tiles < g ? tiles : g);
}
// Two independent scans in one launch.  `b_*` may be all-null for a single scan.  Scan B uses ctx->tile_state2.
static inline int scan_excl_u32_pair(vofod_ctx* ctx, const uint32_t* a_in, uint32_t* a_out, const unsigned long long* a_dn, const size_t a_cap, unsigned long long* a_total,
                                     const bool a_popc, const uint32_t* b_in, uint32_t* b_out, const size_t b_cap, unsigned long long* b_total, const bool b_popc,
                                     uint4* a_list = nullptr, unsigned long long* a_list_n = nullptr, const bool sparse_out = false)
{
  const size_t a_tiles = (a_cap + TILE - 1) / TILE + 1;
  ENSURE(ctx->tile_state, a_tiles * 256 * sizeof(unsigned long long));
  ScanJobs jobs;
  jobs.j[0] = ScanJob{a_in, a_out, a_dn, a_cap, ctx->tile_state.as<unsigned long long>(), a_total, a_popc ? 1 : 0, sparse_out ? 1 : 0, a_list, a_list_n};
  jobs.j[1] = ScanJob{nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 0, nullptr, nullptr};
  const bool big = (a_cap > b_cap ? a_cap : b_cap) >= (size_t(1) << 18);
  int gx = big ? scan_grid<SCAN_BIG_IPT>(ctx, a_cap) : scan_grid<IPT>(ctx, a_cap), gy = 1;
  if (b_in)
  {
    const size_t b_tiles = (b_cap + TILE - 1) / TILE + 1;
    ENSURE(ctx->tile_state2, b_tiles * sizeof(unsigned long long));
    jobs.j[1] = ScanJob{b_in, b_out, nullptr, b_cap, ctx->tile_state2.as<unsigned long long>(), b_total, b_popc ? 1 : 0, sparse_out ? 1 : 0, nullptr, nullptr};
    const int gb = big ? scan_grid<SCAN_BIG_IPT>(ctx, b_cap) : scan_grid<IPT>(ctx, b_cap);
    gx = gx > gb ? gx : gb;
    gy = 2;
  }
  if (ctx->epoch_local >= EPOCH_STRIDE)
    return vf_fail(ctx, VOFOD_E_INTERNAL, "more than %d look-back launches in one call", EPOCH_STRIDE);
  if (big)
    LAUNCH(k_scan_excl_u32<SCAN_BIG_IPT>, dim3((unsigned)gx, (unsigned)gy), NT, 0, jobs, vf_cnt(ctx, CNT_EPOCH_BASE), (uint32_t)(ctx->epoch_local++),
           vf_cnt(ctx, CNT_WATCHDOG));
  else
    LAUNCH(k_scan_excl_u32<IPT>, dim3((unsigned)gx, (unsigned)gy), NT, 0, jobs, vf_cnt(ctx, CNT_EPOCH_BASE), (uint32_t)(ctx->epoch_local++), vf_cnt(ctx, CNT_WATCHDOG));
  return 0;
}
static inline int scan_excl_u32(vofod_ctx* ctx, const uint32_t* in, uint32_t* out, const unsigned long long* d_n, const size_t cap, unsigned long long* d_total)
{
  return scan_excl_u32_pair(ctx, in, out, d_n, cap, d_total, false, nullptr, nullptr, 0, nullptr, false);
}

// sorts bits [begin_bit, end_bit) ascending, stable.  a/b (and va/vb) are ping-pong buffers; *out_k / *out_v receive the result pointers.
template <class KeyT, bool HAS_VAL>
static inline int radix_sort(vofod_ctx* ctx, KeyT* a, KeyT* b, uint32_t* va, uint32_t* vb, const unsigned long long* d_n, const size_t cap, const int begin_bit,
                             const int end_bit, KeyT** out_k, uint32_t** out_v)
{
  const int passes = (end_bit - begin_bit + 7) / 8;
  if (passes <= 0 || passes > 8)
    return vf_fail(ctx, VOFOD_E_INVALID, "radix_sort: bad bit range [%d,%d)", begin_bit, end_bit);
  const size_t tiles = (cap + TILE - 1) / TILE + 1;
  ENSURE(ctx->tile_state, tiles * 256 * sizeof(unsigned long long));
  ENSURE(ctx->sort_hist, 8 * 256 * sizeof(uint32_t));
  {
    const FillJob fj[1] = {{ctx->sort_hist.as<uint32_t>(), 8 * 256, 0u}};
    RET(vf_fill(ctx, fj, 1));
  }
  uint32_t* hist = ctx->sort_hist.as<uint32_t>();
  LAUNCH((k_radix_hist<KeyT>), vf_blocks(ctx, cap, NT, 4), NT, 0, a, d_n, cap, hist, begin_bit, passes);
  KeyT* kin = a;
  KeyT* kout = b;
  uint32_t* vin = va;
  uint32_t* vout = vb;
  for (int p = 0; p < passes; p++)
  {
    if (ctx->epoch_local >= EPOCH_STRIDE)
      return vf_fail(ctx, VOFOD_E_INTERNAL, "more than %d look-back launches in one call", EPOCH_STRIDE);
    LAUNCH((k_radix_pass<KeyT, HAS_VAL>), persistent_grid(ctx, cap), NT, 0, kin, kout, vin, vout, d_n, cap, hist + p * 256,
           ctx->tile_state.as<unsigned long long>(), vf_cnt(ctx, CNT_EPOCH_BASE), (uint32_t)(ctx->epoch_local++), begin_bit + 8 * p, vf_cnt(ctx, CNT_WATCHDOG));
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *out_k = kin;
  if (out_v)
    *out_v = vin;
  return 0;
}
}  // namespace prims
