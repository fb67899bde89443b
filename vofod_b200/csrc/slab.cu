// Multi-GPU slab mode (no counterpart in the reference).
#include "common.cuh"

extern "C" {
int vofod_set_slab(vofod_ctx* ctx, int axis, int lo, int hi)
{
  if (!ctx)
    return vf_fail(nullptr, VOFOD_E_INVALID, "ctx is NULL");
  if (!ctx->map_ready)
    return vf_fail(ctx, VOFOD_E_STATE, "voxel map not sized");
  if (axis < 0 || axis > 2 || lo < 0 || hi > ctx->g.size[axis] || lo >= hi)
    return vf_fail(ctx, VOFOD_E_INVALID, "vofod_set_slab: bad range [%d,%d) on axis %d", lo, hi, axis);
  ctx->g.slab_axis = axis;
  ctx->g.own_lo = lo;
  ctx->g.own_hi = hi;
  return VOFOD_OK;
}
int vofod_slab_boundary(vofod_ctx* ctx, int32_t* point_idx, int32_t* labels, size_t cap, size_t* n)
{
  (void)point_idx;
  (void)labels;
  (void)cap;
  if (n)
    *n = 0;
  return vf_fail(ctx, VOFOD_E_STATE, "vofod_slab_boundary: cluster fragments are not exported in this build");
}
}
