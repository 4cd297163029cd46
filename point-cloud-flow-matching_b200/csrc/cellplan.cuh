// cellplan.cuh — per-call plan that groups the N points of each cloud by voxel cell, laid out for coalesced
// "one thread per cell" traversal (sliced-ELL), used by avg_voxelize forward and trilinear_devoxelize backward.
//
//   light cells (<= light_max points): sorted by point count (descending) and cut into slices of 32 cells; the i-th
//       record of the 32 cells of a slice is contiguous (lane l reads rec[slot0 + 32*i]), so a warp walking 32 cells
//       reads one coalesced line per step and pads only to the longest cell of its slice.
//   heavy cells (> light_max points): records contiguous per cell, reduced by a whole warp per piece of the cell.
//
// Records are 4 bytes (point id) for mean plans and 16 bytes (point id + the three trilinear fractions) for
// trilinear plans.  In mean plans, cells of up to kSortMax points keep their records in ascending point order
// (deterministic sums, the order the oracle uses).
#pragma once
#include "common.cuh"

namespace pcfm {

// Cells longer than the plan's `light_max` are reduced by a whole warp.  Thread-per-cell costs about one instruction
// slot per point-lane but serialises a cell; warp-per-cell adds ~100 warp instructions per cell.  Measured on B200
// (N=20000, r=8..32, warps taking work items dynamically): 64 beats 16, 32 and 128 for both plan kinds.
constexpr int kLightCap = 64;      // upper bound of light_max (<= 255: 8-bit count field of CellPlan::cells)
constexpr int kLightMaxMean = 64;
constexpr int kLightMaxTri = 64;
constexpr int kHeavyChunk = 512;   // trilinear plans: a heavy cell is listed as pieces of at most this many records, so no
                                  // single warp owns a 2000-point cell (the pieces are combined with atomics anyway)
#ifndef PCFM_HEAVY_CHUNK_MEAN
#define PCFM_HEAVY_CHUNK_MEAN 512
#endif
constexpr int kHeavyChunkMean = PCFM_HEAVY_CHUNK_MEAN;  // mean plans: same, so that the 1500-point centre cells of a dense
                                  // cloud on a small grid (r = 8) are not the serial tail of one warp; entry .w = the
                                  // cell's total count (the divisor), pieces of a split cell are added with RED.ADD
constexpr int kSortMax = 16;   // light cells up to this size keep their records in ascending point order
constexpr int kMetaInts = 4;   // per batch element: [0] end of the light list (multiple of 32), [1] heavy cells

struct CellPlan {
  int* cells = nullptr;    // [B, lcap]  light cells, count-descending: (count << 24) | cell; 0 = hole
  int* slot0 = nullptr;    // [B, lcap]  (corner mask << 24) | first record slot (records are 32 slots apart)
  int* meta = nullptr;     // [B, kMetaInts]
  int* heavy = nullptr;    // [B, hcap, 4]  per piece of <= chunk records of a heavy cell: {cell, first slot, records in the
                           //               piece, corner mask (trilinear plans) | records in the whole cell (mean plans)}
  void* rec = nullptr;     // [B, rcap] records
  int lcap = 0, hcap = 0, rcap = 0, light_max = 0, chunk = 0;
  bool pdl = false;        // built by the fused kernel, which releases dependent launches early (launch_dependent)
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

// Mean plan (vox_coords -> ind_out, cnt, *pm in *wsm) and trilinear plan (float_coords -> *pt in *wst) of the same
// clouds by ONE kernel when the fused build applies, else by two builds.
int cellplan_build_dual(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                        int* cnt, CellPlan* pm, Workspace* wsm, CellPlan* pt, Workspace* wst, cudaStream_t st);

// Re-attaches to the plan that an earlier cellplan_build() with the same (kind, b, n, t) left in `ws` (same buffer,
// untouched since): no kernel is launched.  The plan is valid for exactly the coordinates it was built from.
int cellplan_attach(bool trilinear, int b, int n, int t, CellPlan* plan, Workspace* ws);

}  // namespace pcfm
