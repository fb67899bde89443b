// Multi-GPU slab mode (no counterpart in the reference): the global grid is cut along a horizontal axis,
// one slab per context / GPU.  This file: the storage box of a slab and the two-phase scan whose middle is the only
// cross-slab exchange of the mapping path.
//
// Design: the per-scan POINT pipeline (crop, transform, voxel grid, Euclidean clustering) is tiny next to the grids, so every
// slab runs it on the whole (broadcast) scan and gets bit-identical voxels and labels — no cluster-fragment merge is needed
// for the mapping stages.  Everything that touches the grid works on the slab's storage box = own range + halo:
//   * point / rangefinder updates are applied to every HELD cell, own or halo.  They are deterministic functions of the
//     scan, so a halo cell always carries the same value as the neighbour's own copy — without any halo exchange;
//   * the raycast walks all rays but accumulates only inside the storage box.  The path-length sums are exact integers
//     (order independent), so the halo copies again agree bit for bit with the owner's;
//   * nVoxelsOver counts own columns only; hasCloseTo is answered by the slab that owns the query voxel (halo >= window).
// Exchange between the two phases (caller: NCCL all-reduce on the context's stream, or a host loop when several slabs are
// emulated on one device): SUM of the 8-byte background count, MAX of the per-cluster close flags.
// Not in slab mode yet (round 2): classification / detections (exploreToGround needs a wider, exchanged halo) and the
// separated-background-cluster pass (global components need the boundary-fragment merge).
#include "common.cuh"
#include "prims.cuh"

int vf_map_alloc(vofod_ctx* ctx);  // ctx.cu
void vf_bg_state_launch(vofod_ctx* ctx, const vofod_params& p);  // pipeline.cu

extern "C" {
int vofod_set_slab(vofod_ctx* ctx, int axis, int lo, int hi, int halo)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  Geom& g = ctx->g;
  if (axis < 0 || axis > 1 || lo < 0 || hi > g.size[axis] || lo >= hi || halo < 0)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_set_slab: bad range [%d,%d) halo %d on axis %d (horizontal axes only)", lo, hi, halo, axis);
  for (int a = 0; a < 3; a++)
  {
    g.st_lo[a] = 0;
    g.st_size[a] = g.size[a];
  }
  const int slo = lo - halo > 0 ? lo - halo : 0;
  const int shi = hi + halo < g.size[axis] ? hi + halo : g.size[axis];
  g.st_lo[axis] = slo;
  g.st_size[axis] = shi - slo;
  g.slab_axis = axis;
  g.own_lo = lo;
  g.own_hi = hi;
  ctx->slab_on = !(lo == 0 && hi == g.size[axis]);
  ctx->slab_halo = halo;
  // the grids are re-allocated for the storage box; their contents are unspecified until the next vofod_map_set_to
  return vf_map_alloc(ctx);
}

int vofod_slab_boundary(vofod_ctx* ctx, int32_t* point_idx, int32_t* labels, size_t cap, size_t* n)
{
  (void)point_idx;
  (void)labels;
  (void)cap;
  if (n)
    *n = 0;
  return vf_fail(ctx, VOFOD_E_STATE, "vofod_slab_boundary: cluster fragments are not exported yet (slab mode runs the mapping stages only)");
}
}  // extern "C"
