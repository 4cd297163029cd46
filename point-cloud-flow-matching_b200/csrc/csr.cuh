// csr.cuh — per-call "sorted segment" plan: edges (target, source, weight) grouped by target.
//
// Every scatter-add on the hot path (avg_voxelize fwd, trilinear_devoxelize bwd, grouping / gather / 3-NN bwd)
// is turned into a gather over segments: out[t] = sum_{e in seg(t)} w_e * row[src_e].  No fp32 atomics, every
// output element written exactly once (no memset), summation order = ascending edge id => deterministic.
#pragma once
#include "common.cuh"

namespace pcfm {

struct CsrPlan {
  int* start = nullptr;  // [B, T+1] exclusive prefix of segment sizes
  int* ssrc = nullptr;   // [B, E]   source index of the j-th sorted edge
  float* sw = nullptr;   // [B, E]   weight of the j-th sorted edge (nullptr when weights are implicit)
  int* heavy = nullptr;  // [B, hcap] targets whose segment is longer than kHeavySeg (reduced by a whole warp)
  int* nheavy = nullptr; // [B]
  int hcap = 0;
};

constexpr int kHeavySeg = 16;  // segments up to this length are summed by one thread, longer ones by one warp

// bytes of workspace csr_build needs (excluding `cnt`, which the caller provides: int[B,T])
size_t csr_workspace_bytes(int b, long long e, long long t, bool weights);

// Build the plan.
//   tgt   : int[B,E] target of edge e (edges with a target outside [0,T) are dropped), or nullptr when
//           coords != nullptr (voxel mode: tgt = x*r^2+y*r+z from int coords[B,3,E], also written to ind_out)
//   w     : float[B,E] or nullptr
//   src_mod: source index of edge e is (e % src_mod)
//   cnt   : int[B,T], overwritten with the segment sizes
int csr_build(const int* tgt, const int* coords, int r, int* ind_out, const float* w, int b, int e, int t,
              int src_mod, int* cnt, CsrPlan* plan, Workspace* ws, cudaStream_t st);

}  // namespace pcfm
