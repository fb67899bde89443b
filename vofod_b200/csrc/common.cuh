// libvofod_cuda internals: context, geometry, device helpers.  sm_100a only.
// The whole library is compiled with -fmad=false: every fp32 op in the parity-critical device code is a
// separately rounded IEEE operation, exactly like the reference's x86-64 build without -march
// (CMakeLists.txt:10-15).  Division and sqrt are IEEE (--prec-div/--prec-sqrt default true).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/vofod_cuda.h"

#define VOFOD_FULL 0xffffffffu

// ---- geometry of the dense grid (voxel_map.cpp:21-48), passed by value to kernels -------------------
struct Geom
{
  float off[3];
  float vs, inv, half;
  int size[3];      // global cells per axis
  int st_lo[3];     // storage box (global index coords) held by this context: == 0 / size when unsharded
  int st_size[3];
  int own_lo, own_hi, slab_axis;  // owned range along slab_axis (cells outside are halo)
};

__host__ __device__ inline long long geom_cells(const Geom& g) { return (long long)g.st_size[0] * g.st_size[1] * g.st_size[2]; }

// voxel_map.cpp:592-599 — sub, mul, floor, each rounded separately
__device__ __forceinline__ int coord_to_idx1(const float x, const float off, const float inv) { return (int)floorf((x - off) * inv); }
// voxel_map.cpp:607-613
__device__ __forceinline__ float idx_to_coord1(const int i, const float off, const float vs) { return ((float)i + 0.5f) * vs + off; }
// voxel_map.cpp:289-300 (global limits)
__device__ __forceinline__ bool in_limits_idx(const Geom& g, const int x, const int y, const int z)
{
  return x >= 0 && x < g.size[0] && y >= 0 && y < g.size[1] && z >= 0 && z < g.size[2];
}
// storage index of a global cell, or -1 when this context does not hold it
// first statement of every kernel (see LAUNCH): let the next kernel of the stream get launched, then wait until the previous
// one has completed and its writes are visible.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The "memory" clobber does not hold back loads through `const T* __restrict__` kernel parameters: they compile to
// non-coherent loads (LDG.CONSTANT), the data counts as immutable for the kernel's lifetime, and ptxas scheduled such loads
// of per-scan counters ABOVE the wait — reading the previous scan's value.  A pointer laundered through a volatile asm
// cannot be dereferenced before that asm, and volatile asms keep their order.  Use it for whatever a kernel reads
// unconditionally at entry; tools/check_pdl_sass.py (run by the tests) rejects a build with any global access above the wait.
template <class T>
__device__ __forceinline__ T* after_wait(T* p)
{
  asm volatile("" : "+l"(p));
  return p;
}

__device__ __forceinline__ long long cell_index(const Geom& g, const int x, const int y, const int z)
{
  const int lx = x - g.st_lo[0], ly = y - g.st_lo[1], lz = z - g.st_lo[2];
  if (lx < 0 || lx >= g.st_size[0] || ly < 0 || ly >= g.st_size[1] || lz < 0 || lz >= g.st_size[2])
    return -1;
  return (long long)lx + (long long)ly * g.st_size[0] + (long long)lz * g.st_size[0] * g.st_size[1];
}
__device__ __forceinline__ bool cell_owned(const Geom& g, const int x, const int y, const int z)
{
  const int v = g.slab_axis == 0 ? x : (g.slab_axis == 1 ? y : z);
  return v >= g.own_lo && v < g.own_hi;
}

// hasCloseTo (voxel_map.cpp:376-400): window [o-mv, o+mv) clamped, truncated integer norm
__device__ inline bool has_close_to(const float* __restrict__ score, const Geom& g, const float x, const float y, const float z, const float max_dist, const float thr)
{
  const int ox = coord_to_idx1(x, g.off[0], g.inv), oy = coord_to_idx1(y, g.off[1], g.inv), oz = coord_to_idx1(z, g.off[2], g.inv);
  const float md = max_dist * g.inv;
  const int mv = (int)ceilf(md);
  const int bx = max(ox - mv, 0), by = max(oy - mv, 0), bz = max(oz - mv, 0);
  const int ex = min(ox + mv, g.size[0]), ey = min(oy + mv, g.size[1]), ez = min(oz + mv, g.size[2]);
  for (int zi = bz; zi < ez; zi++)
    for (int yi = by; yi < ey; yi++)
      for (int xi = bx; xi < ex; xi++)
      {
        const long long ci = cell_index(g, xi, yi, zi);
        if (ci < 0)
          continue;
        if (score[ci] > thr)
        {
          const int dx = xi - ox, dy = yi - oy, dz = zi - oz;
          const int nrm = (int)sqrt((double)(dx * dx + dy * dy + dz * dz));
          if ((float)nrm <= md)
            return true;
        }
      }
  return false;
}

// slab axis is horizontal (x or y): ownership is a property of the (x,y) column.  lx, ly = storage-local column coordinates.
__device__ __forceinline__ bool column_owned(const Geom& g, const int lx, const int ly)
{
  const int v = g.slab_axis == 0 ? lx + g.st_lo[0] : ly + g.st_lo[1];
  return v >= g.own_lo && v < g.own_hi;
}

// "raised" marks (vofod_ctx::col_dirty): one byte per (chunk of DIRTY_ZC z-levels, column) of the storage box, x fastest
#define DIRTY_ZC 32
__host__ __device__ __forceinline__ int dirty_chunks(const Geom& g) { return (g.st_size[2] + DIRTY_ZC - 1) / DIRTY_ZC; }
__device__ __forceinline__ size_t dirty_index(const Geom& g, const int lx, const int ly, const int lz)
{
  return ((size_t)(lz / DIRTY_ZC) * g.st_size[1] + ly) * g.st_size[0] + lx;
}

// monotone float <-> int mapping for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(const float f)
{
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(const int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

struct MinMax
{
  int mn[3], mx[3];
  unsigned n_valid;
  unsigned n_append;  // cursor of the compacting writers
};

__device__ __forceinline__ void minmax_init(MinMax* mm)
{
  for (int a = 0; a < 3; a++)
  {
    mm->mn[a] = f2ord(3.402823466e+38f);
    mm->mx[a] = f2ord(-3.402823466e+38f);
  }
  mm->n_valid = 0;
  mm->n_append = 0;
}

// layout of a voxel grid over a cloud (voxelgrid.cu): key = i + j*div[0] + k*div[0]*div[1], centre = (ijk + 0.5)*leaf + offset
struct VgLayout
{
  float offset[3];
  float leaf, inv;
  int min_b[3], max_b[3], div[3];
  int overflow;
  unsigned n_valid;
};

// ---- run-based connected components on an occupancy grid (cluster.cu: vf_cluster_runs_dev) -------------------------
// forward rows of the neighbourhood "squared index distance * leaf^2 < tol^2": for row (dy, dz) every cell with |dx| <= R is
// surely inside, the cell at |dx| = R + 1 needs the fp32 distance test when shell == 1 (its exact squared distance EQUALS
// tol^2, the rounding of the centres decides), and shell == 2 means R = -1 and dx = 0 itself is such a border case
#define RUN_ROWS_MAX 16
struct RunRows
{
  int n;
  signed char dy[RUN_ROWS_MAX], dz[RUN_ROWS_MAX], R[RUN_ROWS_MAX], shell[RUN_ROWS_MAX];
};
// occupancy word of the grid clustering: valid iff tag == the current API call number (no clearing between scans)
struct RunWord
{
  unsigned long long tag;
  uint32_t bits;   // occupied x of this 32-cell segment
  uint32_t rank;   // point number of its first occupied cell
};

// ---- raycast accumulator: one u64 per window cell = count (top 20 bits) | signed Q-length (low 44) ---
#define ACC_LEN_BITS 44
__device__ __forceinline__ void acc_decode(const unsigned long long p, unsigned& count, long long& len_q)
{
  len_q = ((long long)(p << (64 - ACC_LEN_BITS))) >> (64 - ACC_LEN_BITS);
  count = (unsigned)((p - (unsigned long long)len_q) >> ACC_LEN_BITS);
}
struct Window
{
  int lo[3];
  int size[3];
};

// ---- grow-only device buffer ------------------------------------------------------------------------
struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  template <class T>
  T* as() const { return (T*)p; }
};

struct Pose33
{
  float R[9];
  float t[3];
};

// Everything that changes from scan to scan and that kernels need.  It lives in device memory (ctx->dyn) and is refreshed
// from a pinned host copy by ONE memcpy per call, so that the kernel arguments themselves are scan-invariant and a
// whole scan can be replayed as a CUDA graph.
struct ScanDyn
{
  Pose33 tf;                 // sensor -> world
  Window win;                // raycast accumulator window of this scan
  float range_pt[3];         // rangefinder seed point (A23)
  int n_seeds;
  const vofod_pt* scan;      // packed scan of this call (staging buffer or resident slot)
  int its_raycast;           // detection_its_diff of the raycast apply
  int pad;
  Window win_apply;          // window of the accumulate that this call applies (= win, or the window of an earlier scan whose apply was deferred)
};
#define EPOCH_STRIDE 64      // look-back launches per API call are numbered 0..63

// A23 rangefinder ground seed (vofod_nodelet.cpp:581-613), one thread
__device__ inline void range_update(float* score, const Geom& g, const ScanDyn* dyn, const double score_point, uint8_t* col_dirty)
{
  const float x = dyn->range_pt[0], y = dyn->range_pt[1], z = dyn->range_pt[2];
  const int repeats = dyn->n_seeds;
  const int ix = coord_to_idx1(x, g.off[0], g.inv), iy = coord_to_idx1(y, g.off[1], g.inv), iz = coord_to_idx1(z, g.off[2], g.inv);
  if (repeats <= 0 || !in_limits_idx(g, ix, iy, iz))  // :599
    return;
  const long long ci = cell_index(g, ix, iy, iz);
  if (ci < 0)
    return;
  float m = score[ci];
  for (int r = 0; r < repeats; r++)
    m = (float)(((double)m + score_point) / 2.0);  // :610
  score[ci] = m;
  col_dirty[dirty_index(g, ix - g.st_lo[0], iy - g.st_lo[1], iz - g.st_lo[2])] = 1;
}

// Euclidean-cluster workspace handles (cluster.cu)
struct ClusterWs
{
  size_t prefilled_tsize = 0;  // != 0: the hash table was cleared ahead of time for a table of this many slots (vf_cluster_prefill)
  DevBuf pts;        // float4 per point
  DevBuf table_key;  // u64 per slot
  DevBuf table_head; // i32 per slot
  DevBuf cellpts;    // float4 per point, grouped by cell (x, y, z, index)
  DevBuf next;       // i32 x2 per point: hash slot, rank inside the cell
  DevBuf parent;     // i32 per point
  DevBuf sizes;      // i32 per point (size of the cluster labelled by this index)
  DevBuf root;       // i32 per point
  DevBuf minidx;     // i32 per point (min member index, valid at roots)
};

// ---- one scan of vofod_process_scan, as planned on the host (pipeline.cu) -------------------------------------------------------------
struct ScanPlan
{
  size_t n;
  vofod_params p;
  vofod_schedule s;
  bool raycast_on;     // do_raycast && not paused && sensor inside the map: this scan's rays are accumulated
  bool apply_on;       // an accumulate is applied in this call: this scan's (unless deferred) or a pending one
  bool apply_first;    // the pending one: before this scan's own accumulate may touch the accumulator
  int raycast_status;  // VOFOD_OK / W_PAUSED / W_SENSOR_OOB as known on the host before launching
  bool sep_first;      // a separated-background pass deferred by the previous call runs at the start of this one
  int sep_first_its;
  vofod_params sep_first_p;
  size_t sep_cap;      // 0 = exact sepclusters (host round trip inside), else capped list
  bool timed;          // record the per-stage events
};
// a scan whose device work has been enqueued and whose results have not been read yet
struct ScanInFlight
{
  ScanPlan plan;
  int sep_status = 0;
  bool applied = false, used_graph = false, sep_ran = false, pipelined = false, active = false;
};

#define MAX_TILE_STATES (1 << 15)
#define VOFOD_SCAN_SLOTS 256

struct vofod_ctx
{
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream3 = nullptr;   // second side branch: hasCloseTo of the voxel list next to its clustering
  cudaStream_t stream2 = nullptr;   // side branch for work that is independent of the main chain (raycast accumulate, second scan)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fills = nullptr, ev_fork2 = nullptr, ev_cls = nullptr, ev_fork3 = nullptr, ev_cp = nullptr;
  bool close_points_done = false;  // k_close_points of this scan ran on the second side branch
  bool nbg_precounted = false;  // CNT_NBG of this scan was counted on the side branch
  size_t cls_prefilled = 0;     // classification work arrays cleared ahead of time for this many points
  size_t sep_prefilled = 0;     // sepclusters fast-path count arrays cleared ahead of time (= their total length)
  std::string err;
  uint64_t n_launches = 0;

  // map
  bool map_ready = false;
  Geom g;
  DevBuf score;   // float
  DevBuf flags;   // uint8
  bool flags_full_dirty = false;
  DevBuf flagged; // u32 cell indices (storage index) written since the last clear
  // Columns (x,y) in which some cell was ever raised above the fill value by a point / rangefinder / apriori update.  Every
  // other column only holds values <= max(fill value, ray score, frontiers threshold), so threshold passes over the grid
  // (nVoxelsOver, voxelsAs*PC) can skip it without reading it.
  DevBuf upd_owner;      // u32 per cell: claim of the point update in flight (UPD_EMPTY otherwise), see k_update_points
  DevBuf upd_leftover;   // u32 keys of the points that found their cell claimed
  size_t upd_owner_cells = 0;
  DevBuf col_dirty;      // u8 per (z-chunk, column) of the storage box: someone RAISED a cell there (dirty_index)
  bool col_all_dirty = true;   // unknown contents (after an upload / single-cell set): every column must be read
  float untouched_max = 0.f;   // the fill value of the last setTo
  size_t flagged_cap = 0;

  // raycast accumulator
  DevBuf acc;     // u64 per window cell
  Window win;
  bool win_valid = false;   // window allocated + zeroed and matches `win`
  int frac_bits = 24;
  bool acc_has_data = false;
  // large windows (long rays on a fine grid: GBs): one "touched" byte per 32 accumulator cells behind the cells, set by the accumulate
  // kernel, so that the apply pass reads 1/256 of the window + the touched 256-byte groups instead of all of it
  bool acc_sparse = false;
  int acc_sparse_mode = 0;     // VOFOD_OPT_ACC_SPARSE: 0 automatic (windows of 2^25 cells and more), 1 always, 2 never
  size_t acc_dirty_off = 0;    // byte offset of the touched marks inside ctx->acc
  size_t acc_total_bytes = 0;  // cells + spare + marks: what a clear of the accumulator has to zero

  // sensor
  int W = 0, H = 0;
  DevBuf lut_dir, lut_off;  // float4 per ray
  DevBuf mask;              // u8 per ray
  bool lut_has_off = false;
  float lut_max_off = 0.f;

  // scan slots (device-resident scans)
  DevBuf scan_staging;  // device copy of the host scan of the current call
  // vofod_prefetch_scan: two device buffers used alternately; a record stays valid until a process_scan consumes it
  DevBuf prefetch_buf[2];
  cudaStream_t stream_copy = nullptr;
  cudaEvent_t ev_prefetch[2] = {nullptr, nullptr};
  const void* prefetched_host[2] = {nullptr, nullptr};
  size_t prefetched_n[2] = {0, 0};
  uint64_t prefetched_call[2] = {0, 0};  // scan_calls when the record was made: a record is good for the next two scan calls only
  uint64_t scan_calls = 0;
  int prefetch_next = 0;
  uint64_t stat_prefetch_hits = 0;
  DevBuf scan_slot[VOFOD_SCAN_SLOTS];
  size_t scan_slot_n[VOFOD_SCAN_SLOTS] = {0};
  void* pinned = nullptr;   // small pinned host block for result read-back
  size_t pinned_bytes = 0;

  // voxel-grid workspace
  DevBuf vg_pts;    // float4 per input point (x,y,z,valid/intensity)
  DevBuf vg_keys_a, vg_keys_b;
  DevBuf vgh_cnt, vgh_bits, vgh_list;
  DevBuf cl_cellkey, cl_words;        // grid clustering of the scan's voxel list: key per point, RunWord per occupancy word  // sort-free scan-path voxel grid: dense per-leaf counts, occupancy words, their popcount scan
  bool cl_force_hash = false;         // test switch: the scan clusters its voxel list with the generic spatial-hash clustering
  bool vg_force_sort = false;         // test switch: the scan path uses the generic sort-based voxel grid
  DevBuf vg_flags, vg_scan, vg_ustart, vg_ukey, vg_pref;
  DevBuf vox;       // vofod_vox per output voxel (cloud_weighted of the last scan)
  DevBuf d_counters;  // u32/u64 scratch counters (see enum below)
  DevBuf tile_state;  // u64 decoupled look-back states
  DevBuf tile_state2; // same, for a scan running on the side branch
  DevBuf sort_hist;   // u32 [passes][256]
  bool scan_prezero = false;  // inside vofod_process_scan: k_begin_call zeroed every per-scan counter, the stages skip their own 8-byte memsets
  size_t sep_table_hint = 0;  // hash-table sizing of the sepclusters clustering (points expected, not the list capacity)
  int epoch_local = 0;        // look-back launch number inside the current API call
  uint64_t epoch_calls = 0;   // host mirror of CNT_EPOCH_BASE / EPOCH_STRIDE

  // per-scan dynamic arguments
  DevBuf dyn;                 // ScanDyn
  ScanDyn* h_dyn = nullptr;   // pinned
  size_t acc_cells_max = 0;

  // CUDA-graph replay of vofod_process_scan
  bool slab_on = false;         // this context holds a slab (own range + halo) of the global grid
  int slab_halo = 0;
  size_t slab_n = 0;            // rays of the scan between vofod_slab_scan_begin and _end (0 = none in flight)
  int slab_raycast_status = 0;
  // slab mode v2 (slab.cu): exchanged buffers and the scan in flight between two phases
  DevBuf slab_patch, slab_patch_desc, slab_patch_meta;  // candidates' boxes for the classification (classify.cu: PatchSet)
  size_t slab_patch_words = 0;
  DevBuf slab_bg_send, slab_bg_recv;                    // packed background-voxel lists of the sepclusters pass: own / every slab's
  size_t slab_bg_cap = 0;                               // entries per slab (identical on every slab: it sizes the allgather)
  size_t slab_bg_consume = 0;                           // rows the replicated rest of the pass is sized for (identical on every slab)
  int slab_next_phase = 0;
  int slab_nranks = 1, slab_rank = 0;
  bool slab_redo_sep = false;
  bool slab_applied = false;
  size_t slab_patch_words_forced = 0;                   // VOFOD_OPT_SLAB_PATCH_WORDS
  vofod_params slab_p;
  vofod_schedule slab_s;
  void* nccl_comm = nullptr;                            // ncclComm_t (vofod_comm_init)
  int raycast_block = 64;       // tuning: rays per block of the accumulate kernel (64 / 128 / 256)
  int raycast_exp = 0;          // VOFOD_OPT_RAYCAST_EXP (measurement only)
  int raycast_spread_voxels = 64;  // VOFOD_OPT_RAYCAST_SPREAD: warp aggregation ends this many voxel sizes along a ray
  bool raycast_no_agg = false;  // experiment switch: one RED per traversal instead of warp-aggregated REDs
  bool raycast_stats = false;   // instrumentation switch: the accumulate kernel fills ray_stats (see RAY_STATS_SLOTS in raycast.cu)
  DevBuf ray_stats;
  bool pdl_enabled = true;      // programmatic dependent launch between consecutive kernels (see LAUNCH)
  bool pdl_in_graph = false;    // experiment switch (VOFOD_OPT_PDL = 2): keep the attribute under stream capture too
  bool pdl_chain = false;       // the last operation put on ctx->stream was a kernel launch of this library
  bool overlap_enabled = true;  // raycast accumulate / second sep scan on the side stream
  bool graph_enabled = true;
  bool capturing = false;
  bool capture_broken = false;
  uint64_t alloc_gen = 0;
  // a few captured scans, keyed by the signature of their launch sequence (the reference's steady-state schedule alternates between two)
  struct GraphSlot
  {
    uint64_t sig = 0;            // signature of `exec`
    cudaGraphExec_t exec = nullptr;
    uint64_t kernels = 0;
    uint64_t seen_sig = 0;       // signature of a kernel-by-kernel scan that ran without allocating (a second one like it gets captured)
    uint64_t seen_gen = ~0ull;
    uint64_t last_use = 0;
  };
  static constexpr int N_GRAPH_SLOTS = 4;
  GraphSlot gslot[N_GRAPH_SLOTS];
  uint64_t gslot_clock = 0;
  bool sep_pending = false;     // a separated-background pass deferred to the start of the next scan (vofod_schedule::sep_deferred)
  int sep_pending_its = 1;
  vofod_params sep_pending_p;
  DevBuf tile_state_b, tile_state2_b;  // look-back states of the deferred pass, which runs next to the front end's own scans
  cudaStream_t stream4 = nullptr;      // the next scan's front end beside the deferred pass
  cudaEvent_t ev_fork4 = nullptr, ev_front = nullptr, ev_sepfill = nullptr;
  ScanInFlight fl[2];           // vofod_process_scan_batch keeps two scans in flight; a single call uses slot 0
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_slab[9] = {};  // vofod_slab_process_scan: start, broadcast, (phase, exchange) x 3, phase 3
  float slab_ms[8] = {};        // vofod_slab_times
  bool slab_acc_forked = false; // the scan's raycast accumulate runs on stream2 (phase 0 .. phase 1)
  bool ray_pending = false;     // an accumulate whose apply was deferred (vofod_schedule::raycast_defer_apply)
  Window ray_pending_win;
  uint64_t stat_replays = 0, stat_captures = 0, stat_capture_failures = 0, stat_eager = 0;
  int stat_last_capture_error = 0;  // 1 enqueue failed, 2 EndCapture failed, 3 buffer growth during capture, 4 instantiate failed, 5 BeginCapture failed
  size_t sep_cap_forced = 0;  // VOFOD_OPT_SEP_CAP
  size_t sep_cap = 0;         // capacity of the background-voxel list when sepclusters runs without a host round trip

  // clustering / per-scan products
  ClusterWs cl, cl_bg;
  DevBuf labels;      // i32 per voxel
  DevBuf pt_close;    // u8 per voxel: hasCloseTo result, then "in close cluster"
  DevBuf cl_close;    // i32 per voxel: per-root close flag
  DevBuf far_list;    // u32: member lists of the far clusters, ascending point index inside a cluster
  DevBuf cl_info;     // vofod_cluster_info per far cluster
  DevBuf dets;        // vofod_detection
  DevBuf explore_ws;
  DevBuf cls_par_stamps, cls_par_queues, cls_par_done, cls_par_terms;  // k_classify_par / k_extract_detections
  bool cls_force_seq = false;  // VOFOD_OPT_CLASSIFY_SEQ
  DevBuf cls_sizes, cls_maxidx, cls_seg, cls_okeys_a, cls_okeys_b, cls_queues, cls_terms;
  DevBuf scratch_a, scratch_b, scratch_d;
  size_t last_m = 0, last_far = 0;

  // sepclusters workspace
  DevBuf sep_colcnt, sep_coloff, sep_raw, sep_ds, sep_labels, sep_nsure, sep_offsets, sep_segcnt, sep_segoff, sep_live, sep_unsure;
  bool sep_force_general = false;  // test switch: never take the leaf-size-1 fast path
  int sep_off_n = -1, sep_off_mv = 0;
  float sep_off_md = 0.f;

  // nodelet state (vofod_nodelet.cpp:2323-2332)
  bool background_pts_sufficient = false;
  bool sure_background_sufficient = false;
  uint32_t last_detection_id = 0;
  int detection_its = 0;
  float cfg_voxel_size = 0.f;

  // instrumentation
  cudaEvent_t ev[VOFOD_N_STAGES + 1];
  bool ev_ok = false;
  float stage_ms[VOFOD_N_STAGES];
};

// device counter slots inside ctx->d_counters (u64 each)
enum
{
  CNT_TRAVERSALS = 0,
  CNT_APPLY_ANY,
  CNT_FLAGGED,
  CNT_FLAGGED_OVERFLOW,
  CNT_NBG,
  CNT_VG_MIN,     // 3 ints packed in 2 slots (see voxelgrid.cu) -> uses slots CNT_VG_MIN .. +1
  CNT_VG_MIN2,
  CNT_VG_MAX,
  CNT_VG_MAX2,
  CNT_VG_NVALID,
  CNT_VG_M,
  CNT_VG_OVERFLOW,
  CNT_NCLUSTERS,
  CNT_NCLOSE,
  CNT_NFAR,
  CNT_NDET,
  CNT_WATCHDOG,
  CNT_MAXVAL,
  CNT_SEP_K,
  CNT_SEP_KDS,
  CNT_SEP_NCL,
  CNT_SEP_ANY_SURE,
  CNT_OOB,
  CNT_EXPLORE_N,
  CNT_SCRATCH0,
  CNT_SCRATCH1,
  CNT_STATE_BG,       // m_background_pts_sufficient (vofod_nodelet.cpp:2323) — device-resident so classification needs no host round trip
  CNT_STATE_SURE,     // m_sure_background_sufficient
  CNT_DET_ID,         // m_last_detection_id
  CNT_NFARPTS,
  CNT_SEP_NENT,
  CNT_SEP_NUNIQ,
  CNT_CL_CURSOR,      // range allocator of the clustering cell arrays
  CNT_CLS_CURSOR,     // range allocator of the far-cluster member lists
  CNT_SEP_LIVE,       // length of the sepclusters work list
  CNT_UPD_LEFT,       // k_update_points: left-over list length / block ticket (both return to 0 at the end of the kernel)
  CNT_UPD_TICKET,
  CNT_CLEAR_TICKET,   // k_clear_flags: blocks done (returns to 0)
  CNT_VGH_LIST,       // non-empty occupancy words listed by the scan
  CNT_SEP_NUNSURE,    // voxels of unsure clusters listed for the decay
  CNT_VGH_WORDS,      // occupancy words of the scan-path voxel grid (depends on the cloud's bounding box)
  CNT_CLS_TICKET,     // k_classify_par: next far cluster to take / clusters finished (adjacent: zeroed together)
  CNT_CLS_FINISHED,
  CNT_DET_BASE,       // id of the scan's first detection (k_classify_par -> k_extract_detections)
  // ---- persistent slots (never zeroed by a map resize) ----
  CNT_EXPLORE_EPOCH,  // stamp generation of the exploreToGround visited cube
  CNT_EPOCH_BASE,     // generation base of the decoupled look-back states, advanced on the DEVICE once per API call (graph replay safe)
  CNT_N_SLOTS = 64,
  CNT_FIRST_PERSISTENT = CNT_EXPLORE_EPOCH
};

// ---- host helpers ----------------------------------------------------------------------------------
int vf_fail(vofod_ctx* c, int code, const char* fmt, ...);
int vf_ensure(vofod_ctx* c, DevBuf& b, size_t bytes);
int vf_begin_call(vofod_ctx* ctx, bool zero_scan_counters = false);
int vf_flush_pending(vofod_ctx* ctx);  // pipeline.cu: carries out a deferred separated-background pass, if any
#define FLUSH_PENDING()          \
  do                             \
  {                              \
    if (ctx->sep_pending)        \
      RET(vf_flush_pending(ctx)); \
  } while (0)
int vf_begin_scan(vofod_ctx* ctx, const vofod_params& p, bool seed_now = true);  // vofod_process_scan: vf_begin_call + rangefinder seeds (A23) + min/max reset of the filter, one kernel  // advances the look-back generation; first thing of every entry point that sorts / scans
int vf_dyn_push(vofod_ctx* ctx);    // h_dyn -> device (stream ordered)
// up to 4 word fills in ONE kernel launch (a stage's clears; unlike memset nodes they chain by programmatic dependent launch)
struct FillJob
{
  uint32_t* p;
  size_t n_words;
  uint32_t value;
};
int vf_fill(vofod_ctx* ctx, const FillJob* jobs, int n_jobs);
// (any runtime call other than a kernel launch ends a chain of programmatically dependent launches: see LAUNCH)
#define CK(call)                                                                                         \
  do                                                                                                     \
  {                                                                                                      \
    ctx->pdl_chain = false;                                                                              \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return vf_fail(ctx, VOFOD_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define RET(call)          \
  do                       \
  {                        \
    int r__ = (call);      \
    if (r__ < 0)           \
      return r__;          \
  } while (0)
#define ENSURE(buf, bytes) RET(vf_ensure(ctx, buf, bytes))
// zero `n` consecutive u64 counter slots, unless k_begin_call already did it for the whole scan
#define ZERO_CNT(slot, n)                                                                        \
  do                                                                                             \
  {                                                                                              \
    if (!ctx->scan_prezero)                                                                      \
      CK(cudaMemsetAsync(vf_cnt(ctx, slot), 0, (size_t)(n) * 8, ctx->stream));                   \
  } while (0)
// kernel launch with accounting.  Every kernel of the library starts with pdl_enter() and is launched with the
// programmatic-stream-serialization attribute (programmatic dependent launch): the launch latency and block scheduling of
// kernel N+1 overlap the execution of kernel N, and N+1 only starts to touch memory after griddepcontrol.wait, i.e. after N
// has completed and flushed.  The scan is a chain of ~50 short kernels, so this latency is a large part of its duration.
// The attribute is only set when the previous operation this context put on the stream was one of its own kernel
// launches (pdl_chain): griddepcontrol.wait orders a kernel after its prerequisite GRIDS — measured: a kernel launched with
// the attribute right behind a cudaMemsetAsync ran concurrently with that memset.  For the same reason the stage-local
// clears on the scan path are kernels (vf_fill), not memset nodes.
// Under stream capture the attribute stays off: measured, a replayed graph with programmatic edges is ~2 % SLOWER than
// one with plain edges (the graph already launches its nodes back to back; blocks parked in griddepcontrol.wait only take
// SM slots from the side branch), while the kernel-by-kernel path gains 20 %.
#define LAUNCH(kern, grid, block, smem, ...)                                                             \
  do                                                                                                     \
  {                                                                                                      \
    cudaLaunchConfig_t cfg__ = {};                                                                       \
    cfg__.gridDim = dim3(grid);                                                                          \
    cfg__.blockDim = dim3(block);                                                                        \
    cfg__.dynamicSmemBytes = (smem);                                                                     \
    cfg__.stream = ctx->stream;                                                                          \
    cudaLaunchAttribute at__[1];                                                                         \
    at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                     \
    at__[0].val.programmaticStreamSerializationAllowed = (ctx->pdl_enabled && ctx->pdl_chain && (!ctx->capturing || ctx->pdl_in_graph)) ? 1 : 0; \
    cfg__.attrs = at__;                                                                                  \
    cfg__.numAttrs = 1;                                                                                  \
    cudaError_t e__ = cudaLaunchKernelEx(&cfg__, kern, __VA_ARGS__);                                     \
    ctx->n_launches++;                                                                                   \
    ctx->pdl_chain = true;                                                                               \
    if (e__ == cudaSuccess)                                                                              \
      e__ = cudaGetLastError();                                                                          \
    if (e__ != cudaSuccess)                                                                              \
      return vf_fail(ctx, VOFOD_E_CUDA, "launch %s failed: %s (%s:%d)", #kern, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static inline int vf_blocks(const vofod_ctx* c, size_t n, int block, int per_sm = 8)
{
  size_t b = (n + block - 1) / block;
  const size_t cap = (size_t)c->num_sms * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static inline unsigned long long* vf_cnt(vofod_ctx* c, int slot) { return c->d_counters.as<unsigned long long>() + slot; }

// ---- stage entry points shared between the staged C ABI and vofod_process_scan (device pointers) ----
// voxelgrid.cu
int vf_filter_voxelize_dev(vofod_ctx* ctx, size_t n, const vofod_params& p, bool seed_cluster = false);  // scan + pose come from ctx->dyn
// cluster.cu: clusters `m_cap`-bounded points whose count lives in d_m (u64 slot); labels = min index
int vf_cluster_dev(vofod_ctx* ctx, ClusterWs& ws, const float* d_xyz, int stride_floats, const unsigned long long* d_m, size_t m_cap, float tol,
                   int* d_labels, unsigned long long* d_ncl, size_t table_points_hint = 0);
int vf_cluster_prefill(vofod_ctx* ctx, ClusterWs& ws, size_t m_cap, size_t table_points_hint);
int vf_classify_prefill(vofod_ctx* ctx, size_t m_cap);
int vf_sepclusters_prefill(vofod_ctx* ctx, const vofod_params& p);
bool vf_run_rows(float tol, float leaf, RunRows& rr);  // false: neighbourhood too large for the grid clustering
int vf_cluster_runs_dev(vofod_ctx* ctx, ClusterWs& ws, const uint32_t* d_cellkey, const RunWord* d_words, const VgLayout* d_layout, const RunRows& rows, float tol,
                        const unsigned long long* d_m, size_t m_cap, int* d_labels, unsigned long long* d_ncl);
int vf_cluster_runs26_dev(vofod_ctx* ctx, ClusterWs& ws, const vofod_vox* d_ds, const uint32_t* d_segbits, const uint32_t* d_segoff, const unsigned long long* d_m,
                          size_t m_cap, int* d_labels, unsigned long long* d_ncl);
// raycast.cu
int vf_raycast_accumulate_dev(vofod_ctx* ctx, size_t n, const vofod_pose& tf, const vofod_params& p);  // scan comes from ctx->dyn
int vf_raycast_prepare(vofod_ctx* ctx, size_t n, const vofod_pose& tf, const vofod_params& p);  // host only: OOB test + window -> h_dyn
int vf_raycast_apply_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p);
int vf_raycast_expand(vofod_ctx* ctx, uint32_t* d_counts, float* d_lengths);
// pipeline.cu
int vf_range_update_dev(vofod_ctx* ctx, const vofod_params& p);  // point + repeat count come from ctx->dyn
int vf_close_far_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const unsigned long long* d_m, size_t m_cap, const vofod_params& p, bool claim_for_update);
int vf_close_far_phase(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const unsigned long long* d_m, size_t m_cap, const vofod_params& p, int phase,
                       bool claim_for_update);
int vf_update_points_scan_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_in_close, const unsigned long long* d_m, size_t m_cap, const vofod_params& p);
int vf_update_owner(vofod_ctx* ctx);
int vf_close_points_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const unsigned long long* d_m, size_t m_cap, const vofod_params& p);  // hasCloseTo per point, ahead of time
int vf_update_points_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const uint8_t* d_sel, int sel_value, const unsigned long long* d_m, size_t m_cap, float score,
                         float flag);
int vf_count_over_dev(vofod_ctx* ctx, float thr, unsigned long long* d_out, const vofod_params* p = nullptr);
const uint8_t* vf_dirty_cols(vofod_ctx* ctx, float thr, const vofod_params* p);  // NULL = every column must be read
// classify.cu
int vf_classify_detect_dev(vofod_ctx* ctx, const vofod_vox* d_vox, const int* d_labels, const uint8_t* d_in_close, const unsigned long long* d_m, size_t m_cap,
                           const vofod_params& p, int phase = 0);  // sensor position comes from ctx->dyn
// slab.cu
void vf_slab_destroy(vofod_ctx* ctx);
// sepclusters.cu
int vf_sepclusters_dev(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t k_cap = 0, int slab_ranks = 0);
int vf_sep_slab_pack(vofod_ctx* ctx, const vofod_params& p, size_t cap);
int vf_sep_slab_finish(vofod_ctx* ctx, int its_diff, const vofod_params& p, size_t cap, int nranks, size_t k_consume);
