// cellplan.cuh — per-call plan that groups the N points of each cloud by voxel cell, laid out for coalesced
// "one thread per cell" traversal (sliced-ELL), used by avg_voxelize forward and trilinear_devoxelize backward.
//
//   light cells (<= kLightMax points): sorted by point count (descending) and cut into slices of 32 cells; the i-th
//       record of the 32 cells of a slice is contiguous (lane l reads rec[slot0 + 32*i]), so a warp walking 32 cells
//       reads one coalesced line per step and pads only to the longest cell of its slice.
//   heavy cells (> kLightMax points): records contiguous per cell, reduced by a whole warp.
//
// Records are 4 bytes (point id) for mean plans and 16 bytes (point id + the three trilinear fractions) for
// trilinear plans.  Light cells keep their records in ascending point order (deterministic sums).
#pragma once
#include "common.cuh"

namespace pcfm {

constexpr int kLightMax = 16;

struct CellPlan {
  int* cells = nullptr;    // [B, lcap]  light cells, count-descending: (count << 24) | cell
  int* slot0 = nullptr;    // [B, lcap]  first record slot of that cell (its records are 32 slots apart)
  int* nlight = nullptr;   // [B]
  int* heavy = nullptr;    // [B, hcap, 3]  {cell, first slot, count}
  int* nheavy = nullptr;   // [B]
  void* rec = nullptr;     // [B, rcap] records
  int lcap = 0, hcap = 0, rcap = 0;
};

struct TriRec {  // 16-byte record of a trilinear plan
  int p;
  float fx, fy, fz;
};

size_t cellplan_workspace_bytes(int b, long long n, long long t, bool trilinear);

// keys: exactly one of
//   vox_coords int[B,3,N]   -> key = x*r^2+y*r+z, also written to ind_out[B,N]              (avg_voxelize)
//   float_coords f32[B,3,N] -> key = floor cell of the clamped coordinate, record carries the fractions (devoxelize)
//   cnt int[B,T] is overwritten with the per-cell counts.
int cellplan_build(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                   int* cnt, CellPlan* plan, Workspace* ws, cudaStream_t st);

}  // namespace pcfm
