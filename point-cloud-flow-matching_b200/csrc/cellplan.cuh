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

constexpr int kLightMax = 32;  // longer cells are reduced by a whole warp
constexpr int kSortMax = 16;   // light cells up to this size keep their records in ascending point order

// Parity classes (trilinear plans): cells are additionally grouped by the parity of (x,y,z).  Two cells of the same
// class are at least two voxels apart along every axis, so their 2x2x2 corner blocks are disjoint: all cells of one
// class can add their 8 corner sums into the plane concurrently with plain read-modify-writes, and a voxel receives
// its (up to) 8 contributions in class order 0..7 — no atomics, no per-corner barriers, fixed summation order.
constexpr int kMetaInts = 32;  // per batch element: [0..7] lstart[q], [8..15] lend[q], [16..24] hstart[q] (9 entries)

struct CellPlan {
  int* cells = nullptr;    // [B, lcap]  light cells, class-major then count-descending: (count << 24) | cell; 0 = hole
  int* slot0 = nullptr;    // [B, lcap]  (corner mask << 24) | first record slot (records are 32 slots apart)
  int* meta = nullptr;     // [B, kMetaInts]
  int* heavy = nullptr;    // [B, hcap, 3]  {cell, first slot, count}, class-major
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
//   parity: group the light / heavy lists by the parity class of the cell (see above); 0 = one class
int cellplan_build(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                   int* cnt, CellPlan* plan, Workspace* ws, cudaStream_t st, int parity = 0);

}  // namespace pcfm
