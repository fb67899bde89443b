// Apriori map ingest: VoFOD::initialize_apriori_map (vofod_nodelet.cpp:305-353) and load_cloud
// (src/pc_loader.cpp:17-90).
//
//   vofod_load_cloud   text cloud -> xyz (host only; the only on-disk format on the path)
//   vofod_apriori_map  rigid transform -> pcl::VoxelGrid centroid down-sample at the map's voxel size -> every down-sampled point that is
//                      inside the map stamps its voxel with +inf -> both "background sufficient" latches are set (:343-344)
// The down-sample runs on the GPU: leaf key per point exactly as PCL computes it (floor(p * inv_leaf) - min_b, dotted with
// (1, div0, div0*div1)), radix sort of (key, point index), per-leaf sequential fp32 sum, centroid = sum / float(count).
// One thing cannot be restated: PCL sorts the (key, index) pairs with std::sort comparing the key only, so the order of the points
// INSIDE a leaf — and with it the last bits of the fp32 sum — is whatever libstdc++'s introsort leaves behind.  Here the sort is
// stable (ascending point index inside a leaf).  The stamped voxel set is what the map sees; it can only differ when a centroid lies
// within an ulp of a voxel face (tests compare it with the reference's own function on random clouds: identical).
#include <errno.h>
#include <math.h>
#include <stdlib.h>

#include <fstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "prims.cuh"

struct ApLayout
{
  int min_b[3], div[3];
  int overflow;  // leaf too small: PCL warns and passes the cloud through un-down-sampled
  float inv;
};

__global__ void __launch_bounds__(256) k_ap_transform(const float* __restrict__ xyz, const size_t n, const Pose33 tf, float4* __restrict__ pts, MinMax* mm)
{
  pdl_enter();
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    // pcl::transformPointCloud(Affine3f), SSE path: x*c0 + (y*c1 + (z*c2 + c3))  (:328)
    float4 p;
    p.x = x * tf.R[0] + (y * tf.R[1] + (z * tf.R[2] + tf.t[0]));
    p.y = x * tf.R[3] + (y * tf.R[4] + (z * tf.R[5] + tf.t[1]));
    p.z = x * tf.R[6] + (y * tf.R[7] + (z * tf.R[8] + tf.t[2]));
    p.w = 1.0f;
    pts[i] = p;
    // getMinMax3D over a dense cloud: no finiteness test (load_cloud sets is_dense, pc_loader.cpp:88)
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int a = 0; a < 3; a++)
  {
    for (int o = 16; o > 0; o >>= 1)
    {
      mn[a] = fminf(mn[a], __shfl_xor_sync(VOFOD_FULL, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(VOFOD_FULL, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(&mm->mn[a], f2ord(mn[a]));
      atomicMax(&mm->mx[a], f2ord(mx[a]));
    }
  }
}
__global__ void k_ap_init(MinMax* mm)
{
  pdl_enter();
  minmax_init(mm);
}
// pcl::VoxelGrid::applyFilter, the set-up (voxel_grid.hpp, PCL 1.10)
__global__ void k_ap_layout(const MinMax* __restrict__ mm, const float leaf, ApLayout* L)
{
  pdl_enter();
  const float inv = 1.0f / leaf;
  long long d[3];
  ApLayout l;
  l.inv = inv;
  for (int a = 0; a < 3; a++)
  {
    const float mn = ord2f(mm->mn[a]), mx = ord2f(mm->mx[a]);
    d[a] = (long long)((mx - mn) * inv) + 1;
    l.min_b[a] = (int)floorf(mn * inv);
    l.div[a] = (int)floorf(mx * inv) - l.min_b[a] + 1;
  }
  l.overflow = d[0] * d[1] * d[2] > 2147483647ll ? 1 : 0;
  *L = l;
}
__global__ void __launch_bounds__(256) k_ap_keys(const float4* __restrict__ pts, const size_t n, const ApLayout* __restrict__ L, uint32_t* __restrict__ keys,
                                                 uint32_t* __restrict__ vals)
{
  pdl_enter();
  const ApLayout l = *after_wait(L);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const float4 p = pts[i];
    const int i0 = (int)(floorf(p.x * l.inv) - (float)l.min_b[0]);
    const int i1 = (int)(floorf(p.y * l.inv) - (float)l.min_b[1]);
    const int i2 = (int)(floorf(p.z * l.inv) - (float)l.min_b[2]);
    keys[i] = l.overflow ? (uint32_t)i : (uint32_t)(i0 + i1 * l.div[0] + i2 * l.div[0] * l.div[1]);  // pass-through: every point its own "leaf"
    vals[i] = (uint32_t)i;
  }
}
__global__ void __launch_bounds__(256) k_ap_heads(const uint32_t* __restrict__ keys, const size_t n, uint32_t* __restrict__ flags)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_ap_starts(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rank, const size_t n, uint32_t* __restrict__ start)
{
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (flags[i])
      start[rank[i]] = (uint32_t)i;
}
// one thread per leaf: CentroidPoint<PointXYZ> (sum of the members, divided by their number), then :339-341
__global__ void __launch_bounds__(256) k_ap_centroid_stamp(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ start,
                                                           const unsigned long long* __restrict__ d_m, const size_t n, float* __restrict__ score, const Geom g,
                                                           uint8_t* __restrict__ col_dirty, float* __restrict__ centroids, unsigned long long* __restrict__ n_stamped)
{
  pdl_enter();
  const size_t m = (size_t)*after_wait(d_m);
  unsigned stamped = 0;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < m; l += (size_t)gridDim.x * blockDim.x)
  {
    const size_t b = start[l], e = l + 1 < m ? start[l + 1] : n;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (size_t k = b; k < e; k++)
    {
      const float4 p = pts[vals[k]];
      sx += p.x; sy += p.y; sz += p.z;
    }
    const float cnt = (float)(e - b);
    const float cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
    if (centroids)
    {
      centroids[3 * l] = cx; centroids[3 * l + 1] = cy; centroids[3 * l + 2] = cz;
    }
    const int x = coord_to_idx1(cx, g.off[0], g.inv), y = coord_to_idx1(cy, g.off[1], g.inv), z = coord_to_idx1(cz, g.off[2], g.inv);
    if (!in_limits_idx(g, x, y, z))  // :340
      continue;
    stamped++;
    const long long ci = cell_index(g, x, y, z);
    if (ci >= 0)
    {
      score[ci] = __int_as_float(0x7f800000);  // :341
      col_dirty[dirty_index(g, x - g.st_lo[0], y - g.st_lo[1], z - g.st_lo[2])] = 1;
    }
  }
  stamped = prims::warp_sum(stamped);
  if ((threadIdx.x & 31) == 0 && stamped)
    atomicAdd(n_stamped, (unsigned long long)stamped);
}
__global__ void k_ap_state(unsigned long long* counters)
{
  pdl_enter();
  counters[CNT_STATE_SURE] = 1ull;  // :343
  counters[CNT_STATE_BG] = 1ull;    // :344
}

extern "C" {

/* load_cloud (src/pc_loader.cpp:17-90): whitespace-separated text, one point per line, at least three numbers per line (the first
 * three are x y z, the rest is ignored); lines with fewer tokens and empty lines are skipped; a ".pts" file carries the point count on
 * its first line.  Numbers are read with atof and narrowed to float.  *n receives the number of points in the file; up to `cap` are
 * written.  VOFOD_E_IO when the file cannot be opened (the reference returns nullptr and the node shuts down, vofod_nodelet.cpp:320-325),
 * VOFOD_E_CAPACITY when cap is too small (call again with *n). */
int vofod_load_cloud(const char* filename, float* xyz, size_t cap, size_t* n)
{
  if (!filename || !n || (cap && !xyz))
    return VOFOD_E_INVALID;
  *n = 0;
  std::ifstream fs;
  fs.open(filename, std::ios::binary);
  if (!fs.is_open() || fs.fail())
    return VOFOD_E_IO;
  const std::string fname(filename);
  const std::string ftype = fname.substr(fname.find_last_of(".") + 1);
  std::string line;
  if (ftype == "pts")
    std::getline(fs, line);  // the point count: only used to reserve memory (:37-42)
  size_t k = 0;
  const char* seps = "\t\r ";
  while (!fs.eof())
  {
    std::getline(fs, line);
    if (line.empty())
      continue;
    // boost::trim: isspace characters off both ends
    const char* ws = " \t\n\v\f\r";
    const size_t b = line.find_first_not_of(ws);
    if (b == std::string::npos)
      continue;  // one empty token: fewer than 3 (:68-72)
    const size_t e = line.find_last_not_of(ws);
    line = line.substr(b, e - b + 1);
    // boost::split(is_any_of("\t\r "), token_compress_on): only the first three tokens matter
    double v[3];
    size_t pos = 0;
    int got = 0;
    while (got < 3 && pos <= line.size())
    {
      const size_t q = line.find_first_of(seps, pos);
      const std::string tok = line.substr(pos, q == std::string::npos ? std::string::npos : q - pos);
      v[got++] = atof(tok.c_str());
      if (q == std::string::npos)
        break;
      pos = line.find_first_not_of(seps, q);
      if (pos == std::string::npos)
        pos = line.size() + 1;  // trailing separators cannot occur after the trim
    }
    if (got < 3)
      continue;
    if (k < cap)
    {
      xyz[3 * k] = (float)v[0];
      xyz[3 * k + 1] = (float)v[1];
      xyz[3 * k + 2] = (float)v[2];
    }
    k++;
  }
  *n = k;
  return k > cap ? VOFOD_E_CAPACITY : VOFOD_OK;
}

/* initialize_apriori_map (vofod_nodelet.cpp:305-353) from a loaded cloud: xyz = n points in the file's frame, tf = the apriori_map/tf
 * transform (:214-225).  centroids (optional, capacity cap points) receives the down-sampled cloud the reference publishes (:347-352);
 * *n_voxels = its size.  Also sets m_sure_background_sufficient and m_background_pts_sufficient (:343-344). */
int vofod_apriori_map(vofod_ctx* ctx, const float* xyz, size_t n, const vofod_pose* tf, float* centroids, size_t cap, size_t* n_voxels)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  FLUSH_PENDING();
  if (!tf || (n && !xyz) || (cap && !centroids))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (n >= (size_t(1) << 31))
    return vf_fail(ctx, VOFOD_E_INVALID, "apriori cloud of %zu points: at most 2^31 - 1", n);
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  if (n_voxels)
    *n_voxels = 0;
  RET(vf_begin_call(ctx));
  if (n)
  {
    using namespace prims;
    const size_t np = padded(n);
    ENSURE(ctx->scratch_a, n * 12);
    ENSURE(ctx->vg_pts, np * 16);
    ENSURE(ctx->vg_keys_a, np * 4);
    ENSURE(ctx->vg_keys_b, np * 4);
    ENSURE(ctx->vg_flags, np * 4);
    ENSURE(ctx->vg_scan, np * 4);
    ENSURE(ctx->vg_ukey, np * 4);
    ENSURE(ctx->vg_ustart, np * 4);
    ENSURE(ctx->scratch_d, sizeof(MinMax) + sizeof(ApLayout) + 128);
    ENSURE(ctx->scratch_b, (cap ? cap : 1) * 12);
    MinMax* mm = ctx->scratch_d.as<MinMax>();
    ApLayout* L = reinterpret_cast<ApLayout*>(ctx->scratch_d.as<char>() + 64);
    CK(cudaMemcpyAsync(ctx->scratch_a.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    Pose33 p33;
    memcpy(p33.R, tf->R, sizeof(p33.R));
    memcpy(p33.t, tf->t, sizeof(p33.t));
    const int nb = vf_blocks(ctx, n, 256, 8);
    LAUNCH(k_ap_init, 1, 1, 0, mm);
    LAUNCH(k_ap_transform, nb, 256, 0, ctx->scratch_a.as<float>(), n, p33, ctx->vg_pts.as<float4>(), mm);
    LAUNCH(k_ap_layout, 1, 1, 0, mm, ctx->g.vs, L);
    LAUNCH(k_ap_keys, nb, 256, 0, ctx->vg_pts.as<float4>(), n, L, ctx->vg_keys_a.as<uint32_t>(), ctx->vg_ukey.as<uint32_t>());
    uint32_t *sk = nullptr, *sv = nullptr;
    RET((radix_sort<uint32_t, true>(ctx, ctx->vg_keys_a.as<uint32_t>(), ctx->vg_keys_b.as<uint32_t>(), ctx->vg_ukey.as<uint32_t>(), ctx->vg_ustart.as<uint32_t>(), nullptr, n,
                                    0, 32, &sk, &sv)));
    LAUNCH(k_ap_heads, nb, 256, 0, sk, n, ctx->vg_flags.as<uint32_t>());
    CK(cudaMemsetAsync(cnt + CNT_SCRATCH0, 0, 16, ctx->stream));
    RET(scan_excl_u32(ctx, ctx->vg_flags.as<uint32_t>(), ctx->vg_scan.as<uint32_t>(), nullptr, n, cnt + CNT_SCRATCH0));
    // leaf starts go where the unsorted keys were (the sort left its result in the other buffer or in this one: use a buffer of its own)
    uint32_t* start = sk == ctx->vg_keys_a.as<uint32_t>() ? ctx->vg_keys_b.as<uint32_t>() : ctx->vg_keys_a.as<uint32_t>();
    LAUNCH(k_ap_starts, nb, 256, 0, ctx->vg_flags.as<uint32_t>(), ctx->vg_scan.as<uint32_t>(), n, start);
    unsigned long long h_m = 0;
    CK(cudaMemcpyAsync(&h_m, cnt + CNT_SCRATCH0, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const bool want_out = centroids && h_m <= cap;
    LAUNCH(k_ap_centroid_stamp, vf_blocks(ctx, (size_t)h_m, 256, 8), 256, 0, ctx->vg_pts.as<float4>(), sv, start, cnt + CNT_SCRATCH0, n, ctx->score.as<float>(), ctx->g,
           ctx->col_dirty.as<uint8_t>(), want_out ? ctx->scratch_b.as<float>() : nullptr, cnt + CNT_SCRATCH1);
    if (want_out && h_m)
      CK(cudaMemcpyAsync(centroids, ctx->scratch_b.p, (size_t)h_m * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_voxels)
      *n_voxels = (size_t)h_m;
    if (centroids && h_m > cap)
    {
      CK(cudaStreamSynchronize(ctx->stream));
      LAUNCH(k_ap_state, 1, 1, 0, cnt);
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->sure_background_sufficient = ctx->background_pts_sufficient = true;
      return vf_fail(ctx, VOFOD_E_CAPACITY, "vofod_apriori_map: %llu down-sampled points, capacity %zu (the map was stamped)", h_m, cap);
    }
  }
  LAUNCH(k_ap_state, 1, 1, 0, cnt);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->sure_background_sufficient = true;
  ctx->background_pts_sufficient = true;
  return VOFOD_OK;
}
}  // extern "C"
