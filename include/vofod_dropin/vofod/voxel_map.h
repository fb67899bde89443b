// Drop-in for the reference's include/vofod/voxel_map.h: put  -I <libvofod_cuda>/include/vofod_dropin -I <libvofod_cuda>/include
// in front of the reference's own include directory and vofod_nodelet.cpp compiles against the GPU-backed classes under the names it
// already uses (vofod::VoxelMap, vofod::VoxelGridWeighted, vofod::VoxelGridCounted, load_cloud).  SURVEY.md §8f N1, INTEGRATION.md.
#pragma once
#include <vofod_b200/voxel_map.hpp>

namespace vofod
{
using VoxelMap = vofod_b200::VoxelMap;
}
