// classifyClusters + extractDetections (vofod_nodelet.cpp:819-879, 1648-1731) on the GPU.
//
//   1. far clusters are put into the reference's processing order (clusters sorted by size, largest first —
//      pcl::EuclideanClusterExtraction; ties by smallest point index) by ranking their unique (size, label) keys, and
//      their points into ascending index order by ballot-compacting labels[label .. max member index] (no sort: n_far
//      is a handful in steady state)
//   2. K13: one thread per far cluster restates pcl::MomentOfInertiaEstimation (fp32 mean / covariance summed in
//      point order, eigenvectors, AABB, OBB) and evaluates the min_points / max_distance / max_size gates
//   3. K14/K15: ONE thread block walks the gated clusters in order, because VoxelMap::exploreToGround of one point
//      reads cells that the previous point's exploration wrote (frontiers write-back, :1712-1715).  Each
//      exploration is a block-parallel flood fill: the reference's DFS visits the same cell SET whatever the
//      order, and only the set (and whether ground / the search horizon was reached) is observable.  The same
//      block then emits the detections (submap uncertainty summed sequentially in double, as the reference does).
#include <math.h>

#include "common.cuh"
#include "prims.cuh"

#define CLS_CANDIDATE (-1)
#define MAX_DETS 16384
#define CLS_PAR_BLOCKS 64  // thread blocks (and stamp cubes) of the parallel classification

// far clusters: per-label size and largest member index, list of far roots (unordered) with their sort keys
__global__ void __launch_bounds__(256) k_cls_mark(const int* __restrict__ labels, const uint8_t* __restrict__ in_close, const unsigned long long* __restrict__ d_m,
                                                  const size_t m_cap, int* __restrict__ sizes, int* __restrict__ maxidx)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned lane = threadIdx.x & 31;
  for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < m; i0 += (size_t)gridDim.x * blockDim.x)
  {
    const size_t i = i0 + lane;
    const int l = i < m ? labels[i] : -1;
    const bool far = i < m && in_close[l] == 0;
    const int key = far ? l : -1 - (int)lane;
    // one atomic per (warp, cluster)
    const unsigned grp = __match_any_sync(VOFOD_FULL, key);
    if (far && lane == (unsigned)(31 - __clz(grp)))  // the highest lane holds the largest index of the group
    {
      atomicAdd(sizes + l, __popc(grp));
      atomicMax(maxidx + l, (int)i);
    }
  }
}
__global__ void __launch_bounds__(256) k_cls_roots(const int* __restrict__ labels, const uint8_t* __restrict__ in_close, const int* __restrict__ sizes,
                                                   const unsigned long long* __restrict__ d_m, const size_t m_cap, const int bits, unsigned long long* __restrict__ okeys,
                                                   unsigned long long* __restrict__ d_nfar)
{
  pdl_enter();
  const size_t m = prims::dev_count(d_m, m_cap);
  const unsigned long long maxv = (1ull << bits) - 1ull;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
    if (labels[i] == (int)i && in_close[i] == 0)
    {
      const unsigned long long pos = atomicAdd(d_nfar, 1ull);
      okeys[pos] = ((maxv - (unsigned long long)sizes[i]) << bits) | (unsigned long long)i;  // size descending, label ascending
    }
}
// The reference processes clusters largest first (ties: smallest point index).  The keys are unique, so the rank of a key
// IS its position: one warp per cluster counts the smaller keys.  n_far is a handful in steady state and a few thousand
// during bootstrap; either way this beats a multi-pass radix sort of n_far 64-bit keys.
__global__ void __launch_bounds__(256) k_cls_rank(const unsigned long long* __restrict__ okeys, const unsigned long long* __restrict__ d_nfar,
                                                  unsigned long long* __restrict__ sorted)
{
  pdl_enter();
  const unsigned long long n = *after_wait(d_nfar);
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long c = warp0; c < n; c += n_warps)
  {
    const unsigned long long key = okeys[c];
    unsigned rank = 0;
    for (unsigned long long j = lane; j < n; j += 32)
      rank += okeys[j] < key;
    rank = prims::warp_sum(rank);
    if (lane == 0)
      sorted[rank] = key;
  }
}
// member lists: the points of every far cluster in ascending index order (the order pcl::EuclideanClusterExtraction
// returns them in).  One warp per cluster walks labels[label .. maxidx] and compacts the matches with ballots.
__global__ void __launch_bounds__(256) k_cls_members(const int* __restrict__ labels, const int* __restrict__ sizes, const int* __restrict__ maxidx,
                                                     const unsigned long long* __restrict__ okeys, const unsigned long long* __restrict__ d_nfar, const int bits,
                                                     uint32_t* __restrict__ memb, int* __restrict__ seg_start, unsigned long long* __restrict__ cursor)
{
  pdl_enter();
  const unsigned long long n_far = *after_wait(d_nfar);
  const unsigned long long lmask = (1ull << bits) - 1ull;
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long c = warp0; c < n_far; c += n_warps)
  {
    const int label = (int)(okeys[c] & lmask);
    const int hi = maxidx[label];
    unsigned long long start = 0;
    if (lane == 0)
    {
      start = atomicAdd(cursor, (unsigned long long)sizes[label]);
      seg_start[label] = (int)start;
    }
    start = __shfl_sync(VOFOD_FULL, start, 0);
    unsigned cnt = 0;
    // 256 indices per round, the 8 loads of a lane in flight together (a far cluster can span most of the list, and one
    // dependent load per 32 indices made this kernel 40 us on such scans)
    for (int base = label; base <= hi; base += 256)
    {
      int lab[8];
#pragma unroll
      for (int q = 0; q < 8; q++)
      {
        const int idx = base + q * 32 + (int)lane;
        lab[q] = idx <= hi ? labels[idx] : -1;
      }
#pragma unroll
      for (int q = 0; q < 8; q++)
      {
        const bool match = lab[q] == label;
        const unsigned bal = __ballot_sync(VOFOD_FULL, match);
        if (match)
          memb[start + cnt + __popc(bal & prims::lanemask_lt())] = (uint32_t)(base + q * 32 + (int)lane);
        cnt += __popc(bal);
      }
    }
  }
}

// cyclic Jacobi, fixed sweep order, fp64 — identical operation sequence to the oracle's restatement
__device__ void jacobi3_dev(const double Ain[3][3], double eval[3], double V[3][3])
{
  double A[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      A[i][j] = Ain[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 16; sweep++)
  {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0)
      break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++)
      {
        if (A[p][q] == 0.0)
          continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
        A[p][p] = app - t * apq;
        A[q][q] = aqq + t * apq;
        A[p][q] = A[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; k++)
        {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  eval[0] = A[0][0];
  eval[1] = A[1][1];
  eval[2] = A[2][2];
}

struct ClsArgs
{
  Geom g;
  int min_points;
  double max_distance, max_size, max_explore_distance;
  float thr_frontiers, thr_new;
  double score_ray;
  double position_sigma;
  double vfov;
  int W, H;
  int bits;
  int side;        // stamp cube side = 2*Rmax+1
  int rmax;
  int terms_cap;
  int n_exact;     // clusters with more points cannot pass the max_size gate
};

// K13 — pcl::MomentOfInertiaEstimation restated (vofod_nodelet.cpp:1655-1672) + the gates (:1679-1690).
// One WARP per far cluster.  Two paths:
//   n <= n_exact (every cluster small enough to pass the max_size gate is): the lanes load 32 points at a time and then
//     ALL lanes run the same sequential fp32 sums over them via shuffles — the reference's summation order, bit for bit,
//     with the memory latency of the index indirection paid once per 32 points;
//   n >  n_exact (cannot fit into max_size, so the class is `invalid` whatever the last bits are): lane-strided fp64
//     partial sums + a fixed-shape warp reduction; the reported box agrees with the sequential fp32 one to ~1e-4
//     (measured 2e-5 on a 10^4-point cluster: the reference's own sequential fp32 sum is the less accurate of the two).
__global__ void __launch_bounds__(256) k_cluster_moi(const ClsArgs a, const ScanDyn* __restrict__ dyn, const vofod_vox* __restrict__ vox, const uint32_t* __restrict__ sidx, const int* __restrict__ seg_start,
                                                     const int* __restrict__ sizes, const unsigned long long* __restrict__ okeys, const unsigned long long* __restrict__ d_nfar,
                                                     vofod_cluster_info* __restrict__ out)
{
  pdl_enter();
  const unsigned long long n_far = *after_wait(d_nfar);
  const unsigned long long lmask = (1ull << a.bits) - 1ull;
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  const float FMAX = 3.402823466e+38f;
  for (unsigned long long c = warp0; c < n_far; c += n_warps)
  {
    const int label = (int)(okeys[c] & lmask);
    const int n = sizes[label];
    const uint32_t* idcs = sidx + seg_start[label];
    const bool exact = n <= a.n_exact;
    float mean[3] = {0.f, 0.f, 0.f};
    float amin[3] = {FMAX, FMAX, FMAX}, amax[3] = {-FMAX, -FMAX, -FMAX};
    float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const unsigned np = n == 0 ? 1u : (unsigned)n;
    if (exact)
    {
      for (int base = 0; base < n; base += 32)  // computeMeanValue
      {
        const int k = base + (int)lane;
        vofod_vox v = {0.f, 0.f, 0.f, 0u};
        if (k < n)
          v = vox[idcs[k]];
        const int cnt = min(32, n - base);
        for (int t = 0; t < cnt; t++)
        {
          const float p[3] = {__shfl_sync(VOFOD_FULL, v.x, t), __shfl_sync(VOFOD_FULL, v.y, t), __shfl_sync(VOFOD_FULL, v.z, t)};
#pragma unroll
          for (int q = 0; q < 3; q++)
          {
            mean[q] += p[q];
            if (p[q] <= amin[q]) amin[q] = p[q];
            if (p[q] >= amax[q]) amax[q] = p[q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 3; q++)
        mean[q] /= (float)np;
      for (int base = 0; base < n; base += 32)  // computeCovarianceMatrix
      {
        const int k = base + (int)lane;
        vofod_vox v = {0.f, 0.f, 0.f, 0u};
        if (k < n)
          v = vox[idcs[k]];
        const int cnt = min(32, n - base);
        for (int t = 0; t < cnt; t++)
        {
          const float d[3] = {__shfl_sync(VOFOD_FULL, v.x, t) - mean[0], __shfl_sync(VOFOD_FULL, v.y, t) - mean[1], __shfl_sync(VOFOD_FULL, v.z, t) - mean[2]};
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
              cov[r][q] += d[r] * d[q];
        }
      }
    } else
    {
      double sm[3] = {0.0, 0.0, 0.0};
      for (int k = (int)lane; k < n; k += 32)
      {
        const vofod_vox v = vox[idcs[k]];
        const float p[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int q = 0; q < 3; q++)
        {
          sm[q] += (double)p[q];
          amin[q] = fminf(amin[q], p[q]);
          amax[q] = fmaxf(amax[q], p[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 3; q++)
      {
        for (int o = 16; o > 0; o >>= 1)
        {
          sm[q] += __shfl_xor_sync(VOFOD_FULL, sm[q], o);
          amin[q] = fminf(amin[q], __shfl_xor_sync(VOFOD_FULL, amin[q], o));
          amax[q] = fmaxf(amax[q], __shfl_xor_sync(VOFOD_FULL, amax[q], o));
        }
        mean[q] = (float)(sm[q] / (double)np);
      }
      double sc[6] = {0, 0, 0, 0, 0, 0};
      for (int k = (int)lane; k < n; k += 32)
      {
        const vofod_vox v = vox[idcs[k]];
        const float d[3] = {v.x - mean[0], v.y - mean[1], v.z - mean[2]};
        sc[0] += (double)(d[0] * d[0]);
        sc[1] += (double)(d[0] * d[1]);
        sc[2] += (double)(d[0] * d[2]);
        sc[3] += (double)(d[1] * d[1]);
        sc[4] += (double)(d[1] * d[2]);
        sc[5] += (double)(d[2] * d[2]);
      }
#pragma unroll
      for (int q = 0; q < 6; q++)
        for (int o = 16; o > 0; o >>= 1)
          sc[q] += __shfl_xor_sync(VOFOD_FULL, sc[q], o);
      cov[0][0] = (float)sc[0]; cov[0][1] = cov[1][0] = (float)sc[1]; cov[0][2] = cov[2][0] = (float)sc[2];
      cov[1][1] = (float)sc[3]; cov[1][2] = cov[2][1] = (float)sc[4]; cov[2][2] = (float)sc[5];
    }
    const float factor = 1.0f / (float)((n - 1 > 0) ? (n - 1) : 1);
    double A[3][3];
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++)
      {
        cov[r][q] *= factor;
        A[r][q] = (double)cov[r][q];
      }
    double evald[3], Vd[3][3];
    jacobi3_dev(A, evald, Vd);
    const float ev[3] = {(float)evald[0], (float)evald[1], (float)evald[2]};
    unsigned major = 0, middle = 1, minor = 2, t;
    if (ev[major] < ev[middle]) { t = major; major = middle; middle = t; }
    if (ev[major] < ev[minor]) { t = major; major = minor; minor = t; }
    if (ev[middle] < ev[minor]) { t = minor; minor = middle; middle = t; }
    float ax[3][3];
    const unsigned order[3] = {major, middle, minor};
    for (int k = 0; k < 3; k++)
    {
      const float v[3] = {(float)Vd[0][order[k]], (float)Vd[1][order[k]], (float)Vd[2][order[k]]};
      const float nrm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int q = 0; q < 3; q++)
        ax[k][q] = v[q] / nrm;
    }
    const float cx = ax[1][1] * ax[2][2] - ax[1][2] * ax[2][1];
    const float cy = ax[1][2] * ax[2][0] - ax[1][0] * ax[2][2];
    const float cz = ax[1][0] * ax[2][1] - ax[1][1] * ax[2][0];
    const float det = ax[0][0] * cx + ax[0][1] * cy + ax[0][2] * cz;
    if (det <= 0.0f)
      for (int q = 0; q < 3; q++)
        ax[0][q] = -ax[0][q];
    // computeOBB: min / max of the projections — order independent, lane-strided + warp reduction
    float omin[3] = {FMAX, FMAX, FMAX}, omax[3] = {-FMAX, -FMAX, -FMAX};
    for (int k = (int)lane; k < n; k += 32)
    {
      const vofod_vox v = vox[idcs[k]];
      const float d[3] = {v.x - mean[0], v.y - mean[1], v.z - mean[2]};
#pragma unroll
      for (int q = 0; q < 3; q++)
      {
        const float pr = d[0] * ax[q][0] + d[1] * ax[q][1] + d[2] * ax[q][2];
        omin[q] = fminf(omin[q], pr);
        omax[q] = fmaxf(omax[q], pr);
      }
    }
#pragma unroll
    for (int q = 0; q < 3; q++)
      for (int o = 16; o > 0; o >>= 1)
      {
        omin[q] = fminf(omin[q], __shfl_xor_sync(VOFOD_FULL, omin[q], o));
        omax[q] = fmaxf(omax[q], __shfl_xor_sync(VOFOD_FULL, omax[q], o));
      }
    if (lane != 0)
      continue;
    vofod_cluster_info ci;
    ci.label = label;
    ci.n_points = n;
    ci.cclass = VOFOD_CLASS_INVALID;
    ci.obb_size = __int_as_float(0x7fc00000);
    float shift[3];
    for (int q = 0; q < 3; q++)
    {
      shift[q] = (omax[q] + omin[q]) / 2.0f;
      omin[q] -= shift[q];
      omax[q] -= shift[q];
    }
    for (int q = 0; q < 3; q++)
    {
      ci.aabb_min[q] = amin[q];
      ci.aabb_max[q] = amax[q];
      ci.obb_min[q] = omin[q];
      ci.obb_max[q] = omax[q];
      ci.obb_center[q] = mean[q] + (ax[0][q] * shift[0] + (ax[1][q] * shift[1] + ax[2][q] * shift[2]));
      for (int k = 0; k < 3; k++)
        ci.obb_rot[q * 3 + k] = ax[k][q];
    }
    {
      double s0 = evald[0], s1 = evald[1], s2 = evald[2], tt;
      if (s0 > s1) { tt = s0; s0 = s1; s1 = tt; }
      if (s1 > s2) { tt = s1; s1 = s2; s2 = tt; }
      if (s0 > s1) { tt = s0; s0 = s1; s1 = tt; }
      const double scale = fmax(fabs(s2), 1e-30);
      ci.eig_gap = (float)(fmin(s1 - s0, s2 - s1) / scale);
    }
    // gates (:1679-1690)
    if (n >= a.min_points)
    {
      const float ddx = dyn->tf.t[0] - ci.obb_center[0], ddy = dyn->tf.t[1] - ci.obb_center[1], ddz = dyn->tf.t[2] - ci.obb_center[2];
      const double dist = (double)sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
      if (!(dist > a.max_distance))
      {
        const float ex = omax[0] - omin[0], ey = omax[1] - omin[1], ez = omax[2] - omin[2];
        ci.obb_size = sqrtf(ex * ex + ey * ey + ez * ez);
        if (!((double)ci.obb_size > a.max_size))
          ci.cclass = CLS_CANDIDATE;
      }
    }
    out[c] = ci;
  }
}

// ---- slab mode: classification on exchanged PATCHES of the map (slab.cu) ------------------------------------------------------------
// A candidate's exploreToGround reads (and its frontier write-back writes) cells up to R voxels around its points, which may lie in
// other slabs.  Every slab therefore packs, for each candidate in classification order, the cells it OWNS of the box
// [aabb - R - 1, aabb + R + 1] as raw bit patterns (zero elsewhere); the sum over all slabs (each cell has one owner) gives every slab
// the same dense copy of every box, the sequential classification runs replicated on those copies, and writes go through to the held
// part of the grid and to every other box that contains the cell.
#define PATCH_MAX 64
struct PatchDesc
{
  int lo[3], size[3];
  unsigned off;  // first word of the box in the patch buffer
  int far_idx;   // index of the candidate among the far clusters
};
struct PatchSet
{
  PatchDesc* desc;            // PATCH_MAX entries
  uint32_t* words;            // bit patterns of the score cells
  unsigned long long* meta;   // [0] number of boxes, [1] words used, [2] overflow flag (a candidate did not fit: the host reports it)
  unsigned budget;            // capacity of `words`
};
__device__ __forceinline__ long long patch_cell(const PatchDesc& d, const int x, const int y, const int z)
{
  const int lx = x - d.lo[0], ly = y - d.lo[1], lz = z - d.lo[2];
  if (lx < 0 || ly < 0 || lz < 0 || lx >= d.size[0] || ly >= d.size[1] || lz >= d.size[2])
    return -1;
  return (long long)d.off + lx + (long long)ly * d.size[0] + (long long)lz * d.size[0] * d.size[1];
}
// value of a map cell as the sequential classification sees it
template <bool PATCH>
__device__ __forceinline__ float cls_load(const float* score, const Geom& g, const PatchDesc* pd, const uint32_t* words, const int x, const int y, const int z)
{
  if (PATCH)
  {
    const long long w = patch_cell(*pd, x, y, z);
    return w >= 0 ? __uint_as_float(words[w]) : 0.0f;
  }
  const long long ci = cell_index(g, x, y, z);
  return ci >= 0 ? score[ci] : 0.0f;
}
// the same read straight from L2: the parallel classification reads cells that a block on ANOTHER SM may have written earlier in the same
// launch, and an SM's L1 is not coherent with those writes
__device__ __forceinline__ float cls_load_cg(const float* score, const Geom& g, const int x, const int y, const int z)
{
  const long long ci = cell_index(g, x, y, z);
  return ci >= 0 ? __ldcg(score + ci) : 0.0f;
}
// explore radius of a candidate (:1698) and its box
__device__ __forceinline__ int cls_explore_radius(const float obb_size, const double max_explore_distance, const float vs)
{
  return (int)(((double)obb_size + max_explore_distance) / (double)vs);
}

// K14 — VoxelMap::exploreToGround (voxel_map.cpp:402-488) as a block-parallel flood fill.  All threads of the block
// call it with identical arguments.  Returns `connected`; when not connected, explored[0..*n_explored) holds the
// cube-relative ids of the visited "unknown" cells (decode with explore_decode).  Shared scratch: sh[0..1] queue
// sizes, sh[2] explored count, sh[3] connected flag.
struct ExploreWs
{
  unsigned* stamps;
  int *q0, *q1, *explored;
  int side, rm;
};
__device__ __forceinline__ void explore_decode(const ExploreWs& w, const int rel, int& dx, int& dy, int& dz)
{
  const int side2 = w.side * w.side;
  dz = rel / side2 - w.rm;
  dy = (rel / w.side) % w.side - w.rm;
  dx = rel % w.side - w.rm;
}
template <bool PATCH, bool CG = false>
__device__ bool explore_to_ground_block(const float* score, const Geom& g, const int ox, const int oy, const int oz, const float unknown_thr, const float ground_thr,
                                        const float maxd, const unsigned epoch, const ExploreWs& w, int* sh, const PatchDesc* pd = nullptr,
                                        const uint32_t* words = nullptr)
{
  const int tid = threadIdx.x;
  const int side = w.side, side2 = w.side * w.side, rm = w.rm;
  // voxel_map.cpp:408-416: border cells count as connected
  if (ox <= 0 || oy <= 0 || oz <= 0 || ox >= g.size[0] - 1 || oy >= g.size[1] - 1 || oz >= g.size[2] - 1)
  {
    __syncthreads();
    if (tid == 0)
      sh[2] = 0;
    __syncthreads();
    return true;
  }
  int* q[2] = {w.q0, w.q1};
  __syncthreads();
  if (tid == 0)
  {
    sh[0] = 1;
    sh[1] = 0;
    sh[2] = 0;
    sh[3] = 0;
    const int rel0 = rm + rm * side + rm * side2;
    w.q0[0] = rel0;
    w.stamps[rel0] = epoch;
  }
  __syncthreads();
  // The loop decision is latched into registers between the two trailing barriers, where nobody writes sh[3] or the
  // next queue's size: every thread of the block takes the same branch, so the barriers always pair up.
  int cur = 0, qn = 1;
  bool stop = false;
  while (qn != 0 && !stop)
  {
    for (int t = tid; t < qn; t += blockDim.x)
    {
      const int rel = q[cur][t];
      int dx, dy, dz;
      explore_decode(w, rel, dx, dy, dz);
      const int x = ox + dx, y = oy + dy, z = oz + dz;
      const float val = CG ? cls_load_cg(score, g, x, y, z) : cls_load<PATCH>(score, g, pd, words, x, y, z);
      if (val > ground_thr)  // :423 -> connected
        sh[3] = 1;
      else if (val > unknown_thr)  // :427
      {
        w.explored[atomicAdd(&sh[2], 1)] = rel;
        const int man = abs(dx) + abs(dy) + abs(dz);
        if ((float)man == maxd - 1.0f)  // :433
          sh[3] = 1;
        else
        {
          const int nb[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
#pragma unroll
          for (int e = 0; e < 6; e++)
          {
            const int nx = x + nb[e][0], ny = y + nb[e][1], nz = z + nb[e][2];
            // :438-478: stay inside the grid and inside the Manhattan horizon
            if (nx < 0 || ny < 0 || nz < 0 || nx > g.size[0] - 1 || ny > g.size[1] - 1 || nz > g.size[2] - 1)
              continue;
            const int ddx = dx + nb[e][0], ddy = dy + nb[e][1], ddz = dz + nb[e][2];
            const int man2 = abs(ddx) + abs(ddy) + abs(ddz);
            if (!((float)man2 <= maxd) || man2 > rm)
              continue;
            const int rel2 = (ddx + rm) + (ddy + rm) * side + (ddz + rm) * side2;
            if (atomicExch(w.stamps + rel2, epoch) != epoch)
              q[cur ^ 1][atomicAdd(&sh[cur ^ 1], 1)] = rel2;
          }
        }
      }
    }
    __syncthreads();
    stop = sh[3] != 0;
    qn = sh[cur ^ 1];
    if (tid == 0)
      sh[cur] = 0;
    cur ^= 1;
    __syncthreads();
  }
  return stop;
}

// staged entry point (vofod_map_explore_to_ground): one exploration, NO write-back (the caller of the reference's
// VoxelMap::exploreToGround decides what to do with the cells)
__global__ void __launch_bounds__(256) k_explore_single(const float* score, const Geom g, const float x, const float y, const float z, const float unknown_thr,
                                                        const float ground_thr, const float maxd, const ExploreWs w, int* __restrict__ out_idx3, const size_t cap,
                                                        unsigned long long* __restrict__ counters)
{
  pdl_enter();
  __shared__ int sh[4];
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0)
  {
    unsigned e = (unsigned)counters[CNT_EXPLORE_EPOCH] + 1u;
    if (e == 0u)
      e = 1u;
    s_epoch = e;
    counters[CNT_EXPLORE_EPOCH] = e;
  }
  __syncthreads();
  const int ox = coord_to_idx1(x, g.off[0], g.inv), oy = coord_to_idx1(y, g.off[1], g.inv), oz = coord_to_idx1(z, g.off[2], g.inv);
  const bool connected = explore_to_ground_block<false>(score, g, ox, oy, oz, unknown_thr, ground_thr, maxd, s_epoch, w, sh);
  const int ne = connected ? 0 : sh[2];
  for (int t = threadIdx.x; t < ne && (size_t)t < cap; t += blockDim.x)
  {
    int dx, dy, dz;
    explore_decode(w, w.explored[t], dx, dy, dz);
    out_idx3[3 * t] = ox + dx;
    out_idx3[3 * t + 1] = oy + dy;
    out_idx3[3 * t + 2] = oz + dz;
  }
  if (threadIdx.x == 0)
  {
    counters[CNT_EXPLORE_N] = (unsigned long long)ne;
    counters[CNT_SCRATCH0] = connected ? 1ull : 0ull;
  }
}

// K14 + K15 — one block, sequential over clusters (see file header).  PATCH: slab mode, the map is read through the exchanged boxes.
template <bool PATCH>
__global__ void __launch_bounds__(256) k_classify_seq(const ClsArgs a, const ScanDyn* __restrict__ dyn, float* score, const vofod_vox* __restrict__ vox, const uint32_t* __restrict__ sidx,
                                                      const int* __restrict__ seg_start, vofod_cluster_info* __restrict__ infos, const ExploreWs w,
                                                      double* __restrict__ terms, vofod_detection* __restrict__ dets, unsigned long long* __restrict__ counters,
                                                      const unsigned long long* __restrict__ d_nfar, const PatchSet ps)
{
  pdl_enter();
  const int n_patch = PATCH ? (int)ps.meta[0] : 0;
  int next_patch = 0;
  __shared__ int sh[4];
  __shared__ unsigned s_epoch;
  __shared__ unsigned long long s_det_id, s_ndet;
  const int tid = threadIdx.x;
  const unsigned long long n_far = *after_wait(d_nfar);
  const bool active = counters[CNT_STATE_BG] != 0ull && counters[CNT_STATE_SURE] != 0ull;  // :1695
  if (tid == 0)
  {
    s_epoch = (unsigned)counters[CNT_EXPLORE_EPOCH];
    s_det_id = counters[CNT_DET_ID];
    s_ndet = 0ull;
  }
  __syncthreads();
  const Geom& g = a.g;
  for (unsigned long long c = 0; c < n_far; c++)
  {
    if (infos[c].cclass != CLS_CANDIDATE)
      continue;  // uniform: every thread reads the same global word
    const PatchDesc* pd = nullptr;
    if (PATCH)
    {
      // the boxes were laid out in this order; a candidate without a box (buffer overflow, reported to the host) stays unclassified
      if (next_patch < n_patch && ps.desc[next_patch].far_idx == (int)c)
        pd = ps.desc + next_patch++;
      else
        continue;
    }
    bool is_floating = true;
    if (active)
    {
      const int label = infos[c].label, n = infos[c].n_points;
      const uint32_t* idcs = sidx + seg_start[label];
      const int R = cls_explore_radius(infos[c].obb_size, a.max_explore_distance, g.vs);  // :1698
      for (int k = 0; k < n && is_floating; k++)
      {
        const vofod_vox v = vox[idcs[k]];
        const int ox = coord_to_idx1(v.x, g.off[0], g.inv), oy = coord_to_idx1(v.y, g.off[1], g.inv), oz = coord_to_idx1(v.z, g.off[2], g.inv);
        __syncthreads();
        if (tid == 0)
        {
          s_epoch++;
          if (s_epoch == 0u)
            s_epoch = 1u;
        }
        __syncthreads();
        const bool connected = explore_to_ground_block<PATCH>(score, g, ox, oy, oz, a.thr_frontiers, a.thr_new, (float)R, s_epoch, w, sh, pd, ps.words);
        if (connected)
          is_floating = false;
        else
        {
          // :1712-1715 — mark the explored unknown cells as frontiers
          const int ne = sh[2];
          for (int t = tid; t < ne; t += blockDim.x)
          {
            int dx, dy, dz;
            explore_decode(w, w.explored[t], dx, dy, dz);
            const long long ci = cell_index(g, ox + dx, oy + dy, oz + dz);
            if (ci >= 0)
              score[ci] = a.thr_frontiers;
            if (PATCH)  // every box that holds the cell (later explorations and the detections' submaps read them)
              for (int q = 0; q < n_patch; q++)
              {
                const long long pw = patch_cell(ps.desc[q], ox + dx, oy + dy, oz + dz);
                if (pw >= 0)
                  ps.words[pw] = __float_as_uint(a.thr_frontiers);
              }
          }
        }
        __syncthreads();
      }
    } else
      is_floating = false;
    __syncthreads();
    if (tid == 0)
      infos[c].cclass = is_floating ? VOFOD_CLASS_MAV : VOFOD_CLASS_UNKNOWN;
    __syncthreads();
  }

  // ---- extractDetections (:834-879) ----
  next_patch = 0;
  for (unsigned long long c = 0; c < n_far; c++)
  {
    const PatchDesc* pd = nullptr;
    if (PATCH)
    {
      while (next_patch < n_patch && ps.desc[next_patch].far_idx < (int)c)
        next_patch++;
      if (next_patch < n_patch && ps.desc[next_patch].far_idx == (int)c)
        pd = ps.desc + next_patch;
    }
    if (infos[c].cclass != VOFOD_CLASS_MAV)
      continue;
    const vofod_cluster_info ci = infos[c];
    const uint32_t* idcs = sidx + seg_start[ci.label];
    // getSubmapCopy(aabb, inflate 2) (voxel_map.cpp:547-584)
    int lo[3], ssz[3];
    float sub_off[3];
#pragma unroll
    for (int q2 = 0; q2 < 3; q2++)
    {
      int mn = coord_to_idx1(ci.aabb_min[q2], g.off[q2], g.inv) - 2, mx = coord_to_idx1(ci.aabb_max[q2], g.off[q2], g.inv) + 2;
      mn = mn < 0 ? 0 : (mn > g.size[q2] - 1 ? g.size[q2] - 1 : mn);
      mx = mx < 0 ? 0 : (mx > g.size[q2] - 1 ? g.size[q2] - 1 : mx);
      lo[q2] = mn;
      ssz[q2] = mx - mn + 1;
      sub_off[q2] = idx_to_coord1(mn, g.off[q2], g.vs) - g.vs / 2.0f;
    }
    const long long ncell = (long long)ssz[0] * ssz[1] * ssz[2];
    const bool fits = ncell <= (long long)a.terms_cap;
    if (fits)
    {
      for (int t = tid; t < (int)ncell; t += blockDim.x)
      {
        const int x = t % ssz[0], y = (t / ssz[0]) % ssz[1], z = t / (ssz[0] * ssz[1]);
        const float val = cls_load<PATCH>(score, g, pd, ps.words, x + lo[0], y + lo[1], z + lo[2]);
        terms[t] = 1.0 - (double)val / a.score_ray;  // :862
      }
      __syncthreads();
      const float ray_f = (float)a.score_ray;
      const double self_term = 1.0 - (double)ray_f / a.score_ray;
      for (int k = tid; k < ci.n_points; k += blockDim.x)  // :855-859 cluster voxels count as certain
      {
        const vofod_vox v = vox[idcs[k]];
        const int x = coord_to_idx1(v.x, sub_off[0], g.inv), y = coord_to_idx1(v.y, sub_off[1], g.inv), z = coord_to_idx1(v.z, sub_off[2], g.inv);
        if (x >= 0 && y >= 0 && z >= 0 && x < ssz[0] && y < ssz[1] && z < ssz[2])
          terms[x + y * ssz[0] + z * ssz[0] * ssz[1]] = self_term;
      }
      __syncthreads();
    }
    if (tid == 0)
    {
      const float ddx = dyn->tf.t[0] - ci.obb_center[0], ddy = dyn->tf.t[1] - ci.obb_center[1], ddz = dyn->tf.t[2] - ci.obb_center[2];
      const double det_dist = (double)sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
      vofod_detection d;
      memset(&d, 0, sizeof(d));
      d.id = (int32_t)(uint32_t)s_det_id++;
      d.label = ci.label;
      d.n_points = (uint64_t)ci.n_points;
      for (int q2 = 0; q2 < 3; q2++)
      {
        d.aabb_min[q2] = ci.aabb_min[q2];
        d.aabb_max[q2] = ci.aabb_max[q2];
        d.obb_min[q2] = ci.obb_min[q2];
        d.obb_max[q2] = ci.obb_max[q2];
        d.position[q2] = ci.obb_center[q2];
      }
      for (int q2 = 0; q2 < 9; q2++)
        d.obb_rot[q2] = ci.obb_rot[q2];
      const float cv = (float)(sqrt(det_dist) * a.position_sigma);
      d.covariance[0] = d.covariance[4] = d.covariance[8] = cv;
      double u = 0.0;
      if (fits)
        for (long long t = 0; t < ncell; t++)
          u += terms[t];
      else
        u = __longlong_as_double(0x7ff8000000000000ll);
      u /= (double)ci.n_points;
      d.confidence = (double)(float)(1.0 / exp(u));
      const double vray_res = a.vfov / (double)a.H;
      const double hray_res = 2 * 3.14159265358979323846 / (double)a.W;
      const double pv = fmin(atan(1.0 / det_dist) / (vray_res * a.min_points), 1.0);
      const double ph = fmin(atan(1.0 / det_dist) / hray_res, 1.0);
      d.detection_probability = pv * ph;
      if (s_ndet < MAX_DETS)
        dets[s_ndet] = d;
      s_ndet++;
    }
    __syncthreads();
  }
  if (tid == 0)
  {
    counters[CNT_EXPLORE_EPOCH] = s_epoch;
    counters[CNT_DET_ID] = s_det_id;
    counters[CNT_NDET] = s_ndet;
  }
}

// K14 in parallel.  The order of the explorations matters only through the map: an exploration that does not reach the ground writes its
// visited cells back as frontiers (:1712-1715), and a later exploration that meets such a cell stops there.  Explorations of one cluster
// stay inside the cluster's index box grown by its explore radius, so two clusters whose grown boxes do not meet cannot see each other.
// NB blocks take the far clusters in the reference's order from a ticket counter; a block waits for every EARLIER cluster whose grown box
// meets its own to be finished (the blocks holding them are running: the lowest unfinished ticket never waits, so this cannot deadlock),
// then runs the cluster's explorations exactly as the sequential kernel does, on a stamp cube / queues of its own, reading the map from L2.
// The last block to finish counts the MAVs, hands k_extract_detections the first detection id and advances m_last_detection_id.
struct ClsParWs
{
  unsigned* stamps;   // [NB][cube]
  int* queues;        // [NB][3][cube]
  unsigned* done;     // [far cluster] == tag: finished in this call
  size_t cube;
};
__device__ __forceinline__ void cls_grown_box(const ClsArgs& a, const vofod_cluster_info& ci, int lo[3], int hi[3])
{
  const int R = cls_explore_radius(ci.obb_size, a.max_explore_distance, a.g.vs) + 1;
#pragma unroll
  for (int q = 0; q < 3; q++)
  {
    lo[q] = coord_to_idx1(ci.aabb_min[q], a.g.off[q], a.g.inv) - R;
    hi[q] = coord_to_idx1(ci.aabb_max[q], a.g.off[q], a.g.inv) + R;
  }
}
__global__ void __launch_bounds__(256) k_classify_par(const ClsArgs a, float* score, const vofod_vox* __restrict__ vox, const uint32_t* __restrict__ sidx,
                                                      const int* __restrict__ seg_start, vofod_cluster_info* infos, const ClsParWs ws,
                                                      unsigned long long* counters, const unsigned long long* __restrict__ d_nfar)
{
  pdl_enter();
  __shared__ int sh[4];
  __shared__ unsigned s_epoch;
  __shared__ unsigned long long s_ticket;
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const unsigned long long n_far = *after_wait(d_nfar);
  const bool active = counters[CNT_STATE_BG] != 0ull && counters[CNT_STATE_SURE] != 0ull;  // :1695
  const unsigned tag = (unsigned)(counters[CNT_EPOCH_BASE] / EPOCH_STRIDE) | 0x80000000u;   // this API call (k_begin_call advanced it)
  ExploreWs w;
  w.stamps = ws.stamps + (size_t)blockIdx.x * ws.cube;
  w.q0 = ws.queues + (size_t)blockIdx.x * 3 * ws.cube;
  w.q1 = w.q0 + ws.cube;
  w.explored = w.q1 + ws.cube;
  w.side = a.side;
  w.rm = a.rmax;
  const Geom& g = a.g;
  while (true)
  {
    __syncthreads();
    if (tid == 0)
    {
      s_ticket = atomicAdd(counters + CNT_CLS_TICKET, 1ull);
      s_last = 0;
    }
    __syncthreads();
    const unsigned long long c = s_ticket;
    if (c >= n_far)
      break;
    const vofod_cluster_info me = infos[c];
    if (me.cclass == CLS_CANDIDATE)
    {
      bool is_floating = true;
      if (active)
      {
        int lo[3], hi[3];
        cls_grown_box(a, me, lo, hi);
        // earlier clusters whose grown box meets this one: finished first (the geometry of every cluster is final since k_cluster_moi;
        // clusters that are no candidates finish at once)
        for (unsigned long long j = tid; j < c; j += blockDim.x)
        {
          int lj[3], hj[3];
          cls_grown_box(a, infos[j], lj, hj);
          if (lj[0] > hi[0] || hj[0] < lo[0] || lj[1] > hi[1] || hj[1] < lo[1] || lj[2] > hi[2] || hj[2] < lo[2])
            continue;
          const volatile unsigned* dj = ws.done + j;
          unsigned spins = 0;
          while (*dj != tag)
          {
            __nanosleep(64);
            if (++spins > (1u << 24))
            {
              atomicAdd(counters + CNT_WATCHDOG, 1ull);
              break;
            }
          }
        }
        __threadfence();
        __syncthreads();
        const int n = me.n_points;
        const uint32_t* idcs = sidx + seg_start[me.label];
        const int R = cls_explore_radius(me.obb_size, a.max_explore_distance, g.vs);  // :1698
        for (int k = 0; k < n && is_floating; k++)
        {
          const vofod_vox v = vox[idcs[k]];
          const int ox = coord_to_idx1(v.x, g.off[0], g.inv), oy = coord_to_idx1(v.y, g.off[1], g.inv), oz = coord_to_idx1(v.z, g.off[2], g.inv);
          __syncthreads();
          if (tid == 0)
          {
            // stamp generations come from ONE counter shared by all blocks (and by the sequential kernels): a cube's layout follows the
            // parameters, so after a change a block can find another block's old stamps in its cube — they must never equal its own
            unsigned e = (unsigned)atomicAdd(counters + CNT_EXPLORE_EPOCH, 1ull) + 1u;
            if (e == 0u)
              e = (unsigned)atomicAdd(counters + CNT_EXPLORE_EPOCH, 1ull) + 1u;
            s_epoch = e;
          }
          __syncthreads();
          const bool connected = explore_to_ground_block<false, true>(score, g, ox, oy, oz, a.thr_frontiers, a.thr_new, (float)R, s_epoch, w, sh);
          if (connected)
            is_floating = false;
          else
          {
            // :1712-1715 — mark the explored unknown cells as frontiers
            const int ne = sh[2];
            for (int t = tid; t < ne; t += blockDim.x)
            {
              int dx, dy, dz;
              explore_decode(w, w.explored[t], dx, dy, dz);
              const long long ci = cell_index(g, ox + dx, oy + dy, oz + dz);
              if (ci >= 0)
                __stcg(score + ci, a.thr_frontiers);
            }
          }
          __syncthreads();
        }
      } else
        is_floating = false;
      if (tid == 0)
        infos[c].cclass = is_floating ? VOFOD_CLASS_MAV : VOFOD_CLASS_UNKNOWN;
    }
    // publish: everything this block wrote for cluster c (frontier cells, the class) before the flag
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
      *(volatile unsigned*)(ws.done + c) = tag;
      __threadfence();
      s_last = atomicAdd(counters + CNT_CLS_FINISHED, 1ull) == n_far - 1ull ? 1 : 0;
    }
    __syncthreads();
    if (s_last)
    {
      // every cluster is finished: count the MAVs (= detections, :834-879), fix the ids
      __threadfence();
      if (tid == 0)
        sh[0] = 0;
      __syncthreads();
      int mine = 0;
      for (unsigned long long j = tid; j < n_far; j += blockDim.x)
        mine += __ldcg(&infos[j].cclass) == VOFOD_CLASS_MAV ? 1 : 0;
      if (mine)
        atomicAdd(&sh[0], mine);
      __syncthreads();
      if (tid == 0)
      {
        const unsigned long long base = counters[CNT_DET_ID];
        counters[CNT_DET_BASE] = base;
        counters[CNT_DET_ID] = base + (unsigned long long)sh[0];
        counters[CNT_NDET] = (unsigned long long)sh[0];
        // (the ticket and finished counters are zeroed before the next launch, not here: other blocks are still asking for tickets)
      }
    }
  }
}

// K15 in parallel — extractDetections (:834-879): one block per MAV, detection number = its rank among the MAVs in cluster order
__global__ void __launch_bounds__(256) k_extract_detections(const ClsArgs a, const ScanDyn* __restrict__ dyn, const float* __restrict__ score, const vofod_vox* __restrict__ vox,
                                                            const uint32_t* __restrict__ sidx, const int* __restrict__ seg_start,
                                                            const vofod_cluster_info* __restrict__ infos, double* __restrict__ terms_all,
                                                            vofod_detection* __restrict__ dets, const unsigned long long* counters,
                                                            const unsigned long long* __restrict__ d_nfar)
{
  pdl_enter();
  __shared__ int s_rank;
  const int tid = threadIdx.x;
  const unsigned long long n_far = *after_wait(d_nfar);
  const unsigned long long det_base = after_wait(counters)[CNT_DET_BASE];
  double* terms = terms_all + (size_t)blockIdx.x * (size_t)a.terms_cap;
  const Geom& g = a.g;
  for (unsigned long long c = blockIdx.x; c < n_far; c += gridDim.x)
  {
    if (infos[c].cclass != VOFOD_CLASS_MAV)
      continue;  // uniform
    __syncthreads();
    if (tid == 0)
      s_rank = 0;
    __syncthreads();
    int mine = 0;
    for (unsigned long long j = tid; j < c; j += blockDim.x)
      mine += infos[j].cclass == VOFOD_CLASS_MAV ? 1 : 0;
    if (mine)
      atomicAdd(&s_rank, mine);
    __syncthreads();
    const int rank = s_rank;
    const vofod_cluster_info ci = infos[c];
    const uint32_t* idcs = sidx + seg_start[ci.label];
    // getSubmapCopy(aabb, inflate 2) (voxel_map.cpp:547-584)
    int lo[3], ssz[3];
    float sub_off[3];
#pragma unroll
    for (int q2 = 0; q2 < 3; q2++)
    {
      int mn = coord_to_idx1(ci.aabb_min[q2], g.off[q2], g.inv) - 2, mx = coord_to_idx1(ci.aabb_max[q2], g.off[q2], g.inv) + 2;
      mn = mn < 0 ? 0 : (mn > g.size[q2] - 1 ? g.size[q2] - 1 : mn);
      mx = mx < 0 ? 0 : (mx > g.size[q2] - 1 ? g.size[q2] - 1 : mx);
      lo[q2] = mn;
      ssz[q2] = mx - mn + 1;
      sub_off[q2] = idx_to_coord1(mn, g.off[q2], g.vs) - g.vs / 2.0f;
    }
    const long long ncell = (long long)ssz[0] * ssz[1] * ssz[2];
    const bool fits = ncell <= (long long)a.terms_cap;
    if (fits)
    {
      for (int t = tid; t < (int)ncell; t += blockDim.x)
      {
        const int x = t % ssz[0], y = (t / ssz[0]) % ssz[1], z = t / (ssz[0] * ssz[1]);
        const float val = cls_load<false>(score, g, nullptr, nullptr, x + lo[0], y + lo[1], z + lo[2]);
        terms[t] = 1.0 - (double)val / a.score_ray;  // :862
      }
      __syncthreads();
      const float ray_f = (float)a.score_ray;
      const double self_term = 1.0 - (double)ray_f / a.score_ray;
      for (int k = tid; k < ci.n_points; k += blockDim.x)  // :855-859 cluster voxels count as certain
      {
        const vofod_vox v = vox[idcs[k]];
        const int x = coord_to_idx1(v.x, sub_off[0], g.inv), y = coord_to_idx1(v.y, sub_off[1], g.inv), z = coord_to_idx1(v.z, sub_off[2], g.inv);
        if (x >= 0 && y >= 0 && z >= 0 && x < ssz[0] && y < ssz[1] && z < ssz[2])
          terms[x + y * ssz[0] + z * ssz[0] * ssz[1]] = self_term;
      }
      __syncthreads();
    }
    if (tid == 0)
    {
      const float ddx = dyn->tf.t[0] - ci.obb_center[0], ddy = dyn->tf.t[1] - ci.obb_center[1], ddz = dyn->tf.t[2] - ci.obb_center[2];
      const double det_dist = (double)sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
      vofod_detection d;
      memset(&d, 0, sizeof(d));
      d.id = (int32_t)(uint32_t)(det_base + (unsigned long long)rank);
      d.label = ci.label;
      d.n_points = (uint64_t)ci.n_points;
      for (int q2 = 0; q2 < 3; q2++)
      {
        d.aabb_min[q2] = ci.aabb_min[q2];
        d.aabb_max[q2] = ci.aabb_max[q2];
        d.obb_min[q2] = ci.obb_min[q2];
        d.obb_max[q2] = ci.obb_max[q2];
        d.position[q2] = ci.obb_center[q2];
      }
      for (int q2 = 0; q2 < 9; q2++)
        d.obb_rot[q2] = ci.obb_rot[q2];
      const float cv = (float)(sqrt(det_dist) * a.position_sigma);
      d.covariance[0] = d.covariance[4] = d.covariance[8] = cv;
      double u = 0.0;
      if (fits)
        for (long long t = 0; t < ncell; t++)
          u += terms[t];  // the reference's order (one running double sum)
      else
        u = __longlong_as_double(0x7ff8000000000000ll);
      u /= (double)ci.n_points;
      d.confidence = (double)(float)(1.0 / exp(u));
      const double vray_res = a.vfov / (double)a.H;
      const double hray_res = 2 * 3.14159265358979323846 / (double)a.W;
      const double pv = fmin(atan(1.0 / det_dist) / (vray_res * a.min_points), 1.0);
      const double ph = fmin(atan(1.0 / det_dist) / hray_res, 1.0);
      d.detection_probability = pv * ph;
      if (rank < MAX_DETS)
        dets[rank] = d;
    }
    __syncthreads();
  }
}

// ---- slab mode: lay out and pack the candidates' boxes --------------------------------------------------------------------------
// one thread: boxes in classification order (few candidates: clusters that passed the size / distance / point-count gates)
__global__ void k_patch_layout(const ClsArgs a, const vofod_cluster_info* __restrict__ infos, const unsigned long long* __restrict__ d_nfar, const PatchSet ps)
{
  pdl_enter();
  const unsigned long long n_far = *after_wait(d_nfar);
  const Geom& g = a.g;
  unsigned used = 0;
  int n = 0;
  unsigned long long overflow = 0ull;
  for (unsigned long long c = 0; c < n_far; c++)
  {
    if (infos[c].cclass != CLS_CANDIDATE)
      continue;
    int R = cls_explore_radius(infos[c].obb_size, a.max_explore_distance, g.vs) + 1;
    if (R < 3)
      R = 3;  // the detection's submap reaches 2 voxels beyond the AABB (:850)
    PatchDesc d;
    unsigned long long vol = 1;
    for (int q = 0; q < 3; q++)
    {
      int mn = coord_to_idx1(infos[c].aabb_min[q], g.off[q], g.inv) - R, mx = coord_to_idx1(infos[c].aabb_max[q], g.off[q], g.inv) + R;
      mn = mn < 0 ? 0 : mn;
      mx = mx > g.size[q] - 1 ? g.size[q] - 1 : mx;
      d.lo[q] = mn;
      d.size[q] = mx >= mn ? mx - mn + 1 : 0;
      vol *= (unsigned long long)d.size[q];
    }
    if (n >= PATCH_MAX || used + vol > (unsigned long long)ps.budget)
    {
      overflow = 1ull;
      break;
    }
    d.off = used;
    d.far_idx = (int)c;
    ps.desc[n++] = d;
    used += (unsigned)vol;
  }
  ps.meta[0] = (unsigned long long)n;
  ps.meta[1] = (unsigned long long)used;
  ps.meta[2] = overflow;
}
// every word of the buffer: the bit pattern of the cell when this slab owns it, zero otherwise (and zero behind the last box)
__global__ void __launch_bounds__(256) k_patch_fill(const float* __restrict__ score, const Geom g, const PatchSet ps)
{
  pdl_enter();
  __shared__ PatchDesc sd[PATCH_MAX];
  const int n = (int)*after_wait(ps.meta);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    sd[i] = ps.desc[i];
  __syncthreads();
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < (size_t)ps.budget; w += (size_t)gridDim.x * blockDim.x)
  {
    int q = n - 1;
    while (q >= 0 && (size_t)sd[q].off > w)
      q--;
    uint32_t v = 0u;
    if (q >= 0)
    {
      const size_t r = w - sd[q].off;
      const size_t sxy = (size_t)sd[q].size[0] * sd[q].size[1];
      if (r < sxy * (size_t)sd[q].size[2])
      {
        const int z = (int)(r / sxy) + sd[q].lo[2], y = (int)((r % sxy) / sd[q].size[0]) + sd[q].lo[1], x = (int)(r % sd[q].size[0]) + sd[q].lo[0];
        const long long ci = cell_index(g, x, y, z);
        if (ci >= 0 && cell_owned(g, x, y, z))
          v = __float_as_uint(score[ci]);
      }
    }
    ps.words[w] = v;
  }
}

static int bits_for_u(unsigned long long v)
{
  int b = 0;
  while (v)
  {
    b++;
    v >>= 1;
  }
  return b < 1 ? 1 : b;
}

// the clears of the next vf_classify_detect_dev over m_cap points (ahead of time, on the scan's side branch)
int vf_classify_prefill(vofod_ctx* ctx, size_t m_cap)
{
  ENSURE(ctx->cls_sizes, m_cap * 4);
  ENSURE(ctx->cls_maxidx, m_cap * 4);
  const FillJob fj[2] = {{ctx->cls_sizes.as<uint32_t>(), m_cap, 0u}, {ctx->cls_maxidx.as<uint32_t>(), m_cap, 0u}};
  RET(vf_fill(ctx, fj, 2));
  ctx->cls_prefilled = m_cap;
  return 0;
}

// phase 1: everything that only looks at the voxel list (far clusters, member lists, moments of inertia + gates);
// phase 2: the sequential part that reads and writes the map (exploreToGround, detections); phase 0: both;
// slab mode: 1, then 3 (lay out + pack the candidates' boxes for the cross-slab sum), then 4 (phase 2 on the summed boxes).
// Inside a scan phase 1 runs on the side branch next to the point update and the ray apply.
int vf_classify_detect_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const uint8_t* d_in_close, const unsigned long long* d_m, size_t m_cap,
                           const vofod_params& p, int phase)
{
  using namespace prims;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  if (phase != 2 && phase != 3 && phase != 4)
  {
    ZERO_CNT(CNT_NDET, 1);
    ZERO_CNT(CNT_NFARPTS, 1);
  }
  if (m_cap == 0)
    return 0;
  const size_t np = padded(m_cap);
  const int bits = bits_for_u(m_cap);  // 2^bits > m_cap: labels and sizes both fit, the 0xFFFFFFFF sentinel sorts last
  // geometry-derived workspace bounds
  const double vs = (double)ctx->g.vs;
  const double rmax_d = (p.cls_max_size + p.cls_max_explore_distance) / vs;
  if (!(rmax_d >= 0.0) || rmax_d > 200.0)
    return vf_fail(ctx, VOFOD_E_INVALID, "classification max_size + max_explore_distance spans %.0f voxels (limit 200)", rmax_d);
  const int rmax = (int)rmax_d + 1;
  const int side = 2 * rmax + 1;
  const size_t cube = (size_t)side * side * side;
  const int sub_side = (int)ceil(p.cls_max_size / vs) + 8;
  const size_t terms_cap = (size_t)sub_side * sub_side * sub_side;

  const unsigned long long* okeys = ctx->cls_okeys_b.as<unsigned long long>();
  uint32_t* sidx = ctx->far_list.as<uint32_t>();
  ClsArgs a;
  a.g = ctx->g;
  a.min_points = p.cls_min_points;
  a.max_distance = p.cls_max_distance;
  a.max_size = p.cls_max_size;
  a.max_explore_distance = p.cls_max_explore_distance;
  a.thr_frontiers = (float)p.thr_frontiers;
  a.thr_new = (float)p.thr_new_obstacles;
  a.score_ray = p.score_ray;
  a.position_sigma = p.output_position_sigma;
  a.vfov = (double)p.sensor_vfov;
  a.W = ctx->W ? ctx->W : 1;
  a.H = ctx->H ? ctx->H : 1;
  a.bits = bits;
  a.side = side;
  a.rmax = rmax;
  a.terms_cap = (int)terms_cap;
  {
    // a cluster of distinct voxel centres whose OBB diagonal is <= max_size lies inside a ball of that diameter: it cannot
    // have more points than a cube of (max_size/vs + 2) voxels per side (generous bound)
    const double side_v = ceil(p.cls_max_size / vs) + 2.0;
    const double cap = side_v * side_v * side_v;
    a.n_exact = cap < 1e6 ? (int)cap : 1000000;
    if (a.n_exact < 64)
      a.n_exact = 64;
  }
  if (phase != 2 && phase != 3 && phase != 4)
  {
  ENSURE(ctx->far_list, np * 4);            // member lists
  ENSURE(ctx->cls_sizes, m_cap * 4);
  ENSURE(ctx->cls_maxidx, m_cap * 4);
  ENSURE(ctx->cls_seg, m_cap * 4);
  ENSURE(ctx->cls_okeys_a, np * 8);
  ENSURE(ctx->cls_okeys_b, np * 8);
  ENSURE(ctx->cl_info, m_cap * sizeof(vofod_cluster_info));
  ENSURE(ctx->dets, (size_t)MAX_DETS * sizeof(vofod_detection));
  ENSURE(ctx->explore_ws, cube * 4);          // stamps: zero-filled when (re)allocated
  ENSURE(ctx->cls_queues, cube * 4 * 3);      // q0, q1, explored
  ENSURE(ctx->cls_terms, terms_cap * 8);
  if (ctx->cls_prefilled != m_cap)
    RET(vf_classify_prefill(ctx, m_cap));
  ctx->cls_prefilled = 0;
  ZERO_CNT(CNT_CLS_CURSOR, 1);

  const int nb = vf_blocks(ctx, m_cap, 256, 8);
  LAUNCH(k_cls_mark, nb, 256, 0, d_labels, d_in_close, d_m, m_cap, ctx->cls_sizes.as<int>(), ctx->cls_maxidx.as<int>());
  LAUNCH(k_cls_roots, nb, 256, 0, d_labels, d_in_close, ctx->cls_sizes.as<int>(), d_m, m_cap, bits, ctx->cls_okeys_a.as<unsigned long long>(), cnt + CNT_NFARPTS);
  const int nbw = vf_blocks(ctx, m_cap * 32, 256, 4);
  LAUNCH(k_cls_rank, nbw, 256, 0, ctx->cls_okeys_a.as<unsigned long long>(), cnt + CNT_NFARPTS, ctx->cls_okeys_b.as<unsigned long long>());
  okeys = ctx->cls_okeys_b.as<unsigned long long>();
  sidx = ctx->far_list.as<uint32_t>();
  LAUNCH(k_cls_members, nbw, 256, 0, d_labels, ctx->cls_sizes.as<int>(), ctx->cls_maxidx.as<int>(), okeys, cnt + CNT_NFARPTS, bits, sidx, ctx->cls_seg.as<int>(),
         cnt + CNT_CLS_CURSOR);
  LAUNCH(k_cluster_moi, vf_blocks(ctx, m_cap * 32, 256, 4), 256, 0, a, ctx->dyn.as<ScanDyn>(), d_vox, sidx, ctx->cls_seg.as<int>(), ctx->cls_sizes.as<int>(), okeys, cnt + CNT_NFARPTS,
         ctx->cl_info.as<vofod_cluster_info>());
  }
  if (phase == 1)
    return 0;
  PatchSet ps = {};
  if (phase == 3 || phase == 4)
  {
    ps.desc = ctx->slab_patch_desc.as<PatchDesc>();
    ps.words = ctx->slab_patch.as<uint32_t>();
    ps.meta = ctx->slab_patch_meta.as<unsigned long long>();
    ps.budget = (unsigned)ctx->slab_patch_words;
  }
  if (phase == 3)
  {
    LAUNCH(k_patch_layout, 1, 1, 0, a, ctx->cl_info.as<vofod_cluster_info>(), cnt + CNT_NFARPTS, ps);
    LAUNCH(k_patch_fill, vf_blocks(ctx, ctx->slab_patch_words, 256, 8), 256, 0, ctx->score.as<float>(), ctx->g, ps);
    return 0;
  }
  int* qbase = ctx->cls_queues.as<int>();
  ExploreWs w;
  w.stamps = ctx->explore_ws.as<unsigned>();
  w.q0 = qbase;
  w.q1 = qbase + cube;
  w.explored = qbase + 2 * cube;
  w.side = side;
  w.rm = rmax;
  if (phase != 4 && !ctx->cls_force_seq)
  {
    // parallel over the far clusters (see k_classify_par); the sequential kernel stays for slab mode (the map is read through exchanged
    // boxes there) and as VOFOD_OPT_CLASSIFY_SEQ
    const int NB = CLS_PAR_BLOCKS;
    ENSURE(ctx->cls_par_stamps, (size_t)NB * cube * 4);   // zero-filled when (re)allocated: generation 0 is never used
    ENSURE(ctx->cls_par_queues, (size_t)NB * cube * 4 * 3);
    ENSURE(ctx->cls_par_done, m_cap * 4);
    ENSURE(ctx->cls_par_terms, (size_t)NB * terms_cap * 8);
    ZERO_CNT(CNT_CLS_TICKET, 2);  // + CNT_CLS_FINISHED
    ClsParWs pw;
    pw.stamps = ctx->cls_par_stamps.as<unsigned>();
    pw.queues = ctx->cls_par_queues.as<int>();
    pw.done = ctx->cls_par_done.as<unsigned>();
    pw.cube = cube;
    LAUNCH(k_classify_par, NB, 256, 0, a, ctx->score.as<float>(), d_vox, sidx, ctx->cls_seg.as<int>(), ctx->cl_info.as<vofod_cluster_info>(), pw, cnt, cnt + CNT_NFARPTS);
    LAUNCH(k_extract_detections, NB, 256, 0, a, ctx->dyn.as<ScanDyn>(), ctx->score.as<float>(), d_vox, sidx, ctx->cls_seg.as<int>(), ctx->cl_info.as<vofod_cluster_info>(),
           ctx->cls_par_terms.as<double>(), ctx->dets.as<vofod_detection>(), cnt, cnt + CNT_NFARPTS);
    return 0;
  }
  if (phase == 4)
    LAUNCH(k_classify_seq<true>, 1, 256, 0, a, ctx->dyn.as<ScanDyn>(), ctx->score.as<float>(), d_vox, sidx, ctx->cls_seg.as<int>(), ctx->cl_info.as<vofod_cluster_info>(), w,
           ctx->cls_terms.as<double>(), ctx->dets.as<vofod_detection>(), cnt, cnt + CNT_NFARPTS, ps);
  else
    LAUNCH(k_classify_seq<false>, 1, 256, 0, a, ctx->dyn.as<ScanDyn>(), ctx->score.as<float>(), d_vox, sidx, ctx->cls_seg.as<int>(), ctx->cl_info.as<vofod_cluster_info>(), w,
           ctx->cls_terms.as<double>(), ctx->dets.as<vofod_detection>(), cnt, cnt + CNT_NFARPTS, ps);
  return 0;
}

extern "C" int vofod_map_explore_to_ground(vofod_ctx* ctx, const float pt[3], float unknown_threshold, float ground_threshold, float max_voxel_dist, int* connected,
                                           int32_t* explored_idx3, size_t cap, size_t* n_explored)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  FLUSH_PENDING();
  if (!pt || !connected || !n_explored || (cap && !explored_idx3))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  if (!(max_voxel_dist >= 0.0f) || max_voxel_dist > 200.0f)
    return vf_fail(ctx, VOFOD_E_INVALID, "max_voxel_dist %.1f out of range [0,200]", max_voxel_dist);
  const int rmax = (int)ceilf(max_voxel_dist) + 1;
  const int side = 2 * rmax + 1;
  const size_t cube = (size_t)side * side * side;
  ENSURE(ctx->explore_ws, cube * 4);
  ENSURE(ctx->cls_queues, cube * 4 * 3);
  ENSURE(ctx->scratch_b, cap * 12 + 16);
  int* qbase = ctx->cls_queues.as<int>();
  ExploreWs w;
  w.stamps = ctx->explore_ws.as<unsigned>();
  w.q0 = qbase;
  w.q1 = qbase + cube;
  w.explored = qbase + 2 * cube;
  w.side = side;
  w.rm = rmax;
  unsigned long long* cnt = ctx->d_counters.as<unsigned long long>();
  LAUNCH(k_explore_single, 1, 256, 0, ctx->score.as<float>(), ctx->g, pt[0], pt[1], pt[2], unknown_threshold, ground_threshold, max_voxel_dist, w, ctx->scratch_b.as<int>(), cap,
         cnt);
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(&h[0], cnt + CNT_EXPLORE_N, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&h[1], cnt + CNT_SCRATCH0, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *connected = h[1] != 0;
  *n_explored = (size_t)h[0];
  const size_t k = h[0] < cap ? (size_t)h[0] : cap;
  if (k)
  {
    CK(cudaMemcpyAsync(explored_idx3, ctx->scratch_b.p, k * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return h[0] > cap ? vf_fail(ctx, VOFOD_E_CAPACITY, "explore_to_ground: need capacity %llu", h[0]) : VOFOD_OK;
}

extern "C" int vofod_classify_detect(vofod_ctx* ctx, const vofod_vox* pts, const int32_t* labels, const uint8_t* point_in_close_cluster, size_t m, const vofod_pose* tf,
                                     const vofod_params* p, vofod_detection* dets, size_t det_cap, size_t* n_dets, vofod_cluster_info* clusters, size_t cl_cap,
                                     size_t* n_far_clusters)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  FLUSH_PENDING();
  if (!tf || !p || !n_dets || !n_far_clusters || (m && (!pts || !labels || !point_in_close_cluster)))
    return vf_fail(ctx, VOFOD_E_INVALID, "NULL argument");
  *n_dets = 0;
  *n_far_clusters = 0;
  if (m == 0)
    return VOFOD_OK;
  for (size_t i = 0; i < m; i++)
    if (labels[i] < 0 || (size_t)labels[i] >= m)
      return vf_fail(ctx, VOFOD_E_INVALID, "labels[%zu] = %d is not a point index", i, labels[i]);
  ENSURE(ctx->vox, prims::padded(m) * sizeof(vofod_vox));
  ENSURE(ctx->labels, m * 4);
  ENSURE(ctx->pt_close, m + 64);
  CK(cudaMemcpyAsync(ctx->vox.p, pts, m * sizeof(vofod_vox), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->labels.p, labels, m * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->pt_close.p, point_in_close_cluster, m, cudaMemcpyHostToDevice, ctx->stream));
  memcpy(ctx->h_dyn->tf.R, tf->R, sizeof(tf->R));
  memcpy(ctx->h_dyn->tf.t, tf->t, sizeof(tf->t));
  RET(vf_begin_call(ctx));
  RET(vf_dyn_push(ctx));
  RET(vf_classify_detect_dev(ctx, ctx->vox.as<vofod_vox>(), ctx->labels.as<int>(), ctx->pt_close.as<uint8_t>(), nullptr, m, *p));
  unsigned long long h[CNT_N_SLOTS];
  CK(cudaMemcpyAsync(h, ctx->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (h[CNT_WATCHDOG])
    return vf_fail(ctx, VOFOD_E_INTERNAL, "device watchdog tripped (%llu)", h[CNT_WATCHDOG]);
  ctx->last_detection_id = (uint32_t)h[CNT_DET_ID];
  ctx->last_m = m;
  ctx->last_far = (size_t)h[CNT_NFARPTS];
  *n_dets = (size_t)h[CNT_NDET];
  *n_far_clusters = ctx->last_far;
  if (clusters)
  {
    const size_t k = ctx->last_far < cl_cap ? ctx->last_far : cl_cap;
    if (k)
      CK(cudaMemcpyAsync(clusters, ctx->cl_info.p, k * sizeof(vofod_cluster_info), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (dets)
  {
    size_t k = *n_dets < det_cap ? *n_dets : det_cap;
    if (k > MAX_DETS)
      k = MAX_DETS;
    if (k)
      CK(cudaMemcpyAsync(dets, ctx->dets.p, k * sizeof(vofod_detection), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if ((dets && *n_dets > det_cap) || (clusters && ctx->last_far > cl_cap))
    return vf_fail(ctx, VOFOD_E_CAPACITY, "classify_detect: %zu detections / %zu far clusters exceed the given capacities", *n_dets, ctx->last_far);
  return VOFOD_OK;
}
