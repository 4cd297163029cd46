// cellplan.cuh — per-call plan that groups the N points of each cloud by voxel cell, laid out for coalesced
// "one thread per cell" traversal (sliced-ELL), used by avg_voxelize forward and trilinear_devoxelize backward.
//
//   light cells (<= light_max points): sorted by point count (descending) and cut into slices of 32 cells; the i-th
//       record of the 32 cells of a slice is contiguous (lane l reads rec[slot0 + 32*i]), so a warp walking 32 cells
//       reads one coalesced line per step and pads only to the longest cell of its slice.
//   heavy cells (> light_max points): records contiguous per cell, reduced by a whole warp per piece of the cell.
//
// Mean plans: a record is a 4-byte point id; cells of up to kSortMax points keep their records in ascending point
// order (deterministic sums, the order the oracle uses); a heavy cell's records start on a 16-byte boundary and are
// padded to a multiple of 4 with the id N (the consumer keeps an exact zero at row[N]), so a lane reads 4 ids at once.
//
// Trilinear plans: the unit is a PAIR of points, stored as two float4 (structure of arrays over pair slots):
//     recA[slot] = {off0, off1, fx0, fx1}     recB[slot] = {fy0, fy1, fz0, fz1}
// off = 4 * point id (the byte offset into a channel row), f* = the trilinear fractions.  One 16-byte load per array
// yields two points whose fractions sit in aligned register pairs, so the whole weight / accumulate arithmetic of the
// consumers issues as packed fp32 (FADD2 / FMUL2 / FFMA2: two points per instruction).  A cell with an odd number of
// points ends in a dummy half {off = 4*N, fractions 0}: the consumer keeps an exact zero (or a scratch slot) at row[N].
// Light cells are binned by their number of pairs; the i-th pair of the 32 cells of a slice is contiguous.
#pragma once
#include "common.cuh"

namespace pcfm {

// Cells longer than the plan's `light_max` are reduced by a whole warp.  Thread-per-cell costs about one instruction
// slot per point-lane but serialises a cell; warp-per-cell adds ~100 warp instructions per cell.  Measured on B200
// (N=20000, r=8..32, warps taking work items dynamically): 64 beats 16, 32 and 128 for both plan kinds.
constexpr int kLightCap = 64;      // upper bound of light_max (<= 255: 8-bit count field of CellPlan::cells)
#ifndef PCFM_LIGHT_MAX_MEAN
#define PCFM_LIGHT_MAX_MEAN 64
#endif
#ifndef PCFM_LIGHT_MAX_TRI
#define PCFM_LIGHT_MAX_TRI 64
#endif
#ifndef PCFM_HEAVY_CHUNK_TRI
#define PCFM_HEAVY_CHUNK_TRI 512
#endif
constexpr int kLightMaxMean = PCFM_LIGHT_MAX_MEAN;   // even, <= kLightCap
constexpr int kLightMaxTri = PCFM_LIGHT_MAX_TRI;     // even, <= kLightCap
constexpr int kHeavyChunk = PCFM_HEAVY_CHUNK_TRI;   // trilinear plans: a heavy cell is listed as pieces of at most this many records, so no
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
  void* rec = nullptr;     // [B, rcap] records (mean: int point ids; trilinear: float4 recA, see above)
  void* rec2 = nullptr;    // [B, rcap] trilinear plans: float4 recB
  int lcap = 0, hcap = 0, rcap = 0, light_max = 0, chunk = 0;
  bool pdl = false;        // built by the fused kernel, which releases dependent launches early (launch_dependent)
};

// A work item of the cell kernels and its header, fetched ONE ITEM AHEAD so that the header's L2 latency overlaps the
// previous item's work: light slice -> x = cells entry, y = slot0 entry of this lane's cell; heavy piece -> the int4.
struct CellItem {
  int item;
  int4 h;
};
__device__ __forceinline__ CellItem next_cell_item(int* counter, int nslice, int nitem, const int* __restrict__ cl,
                                                   const int* __restrict__ sl, const int4* __restrict__ hv) {
  const int lane = threadIdx.x & 31;
  CellItem c;
  c.item = 0;
  // one shared-memory atomic per warp (written in PTX: the compiler wraps a C++ atomicAdd under `if (lane == 0)` in
  // its warp-aggregation sequence, ~15 instructions per item)
  if (lane == 0)
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(c.item) : "r"(smem_u32(counter)) : "memory");
  c.item = __shfl_sync(0xffffffffu, c.item, 0);
  c.h = make_int4(0, 0, 0, 0);
  if (c.item < nslice) {
    c.h.x = __ldg(cl + ((c.item << 5) | lane));
    c.h.y = __ldg(sl + ((c.item << 5) | lane));
  } else if (c.item < nitem) {
    c.h = __ldg(hv + (c.item - nslice));
  }
  return c;
}
// Makes a prefetched item the current one.  The empty asm pins the header registers to THIS point of the program:
// without it the compiler moves the warp-uniform fields of a heavy header into uniform registers (R2UR) right behind
// the load, i.e. it waits for the L2 round trip one item early and the prefetch hides nothing (ncu: 500 stall samples
// on that R2UR).
__device__ __forceinline__ CellItem take_cell_item(CellItem n) {
  asm volatile("" : "+r"(n.item), "+r"(n.h.x), "+r"(n.h.y), "+r"(n.h.z), "+r"(n.h.w));
  return n;
}

// packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2, every lane rounded like the scalar instruction)
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// false when the plan's packed fields cannot hold this problem (24-bit record slots and cell ids)
bool cellplan_fits(long long n, long long t, bool trilinear);

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
