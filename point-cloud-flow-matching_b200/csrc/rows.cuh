// rows.cuh — the two channel-row engines every PVCNN feature op maps onto.
//
// All PVCNN feature tensors are channel-first ([B,C,N] / [B,C,r^3]): the gather / scatter index runs along the
// CONTIGUOUS axis, so one CTA owns one (batch, channel) row, stages the randomly-read side in shared memory
// (TMA bulk copy) and streams the other side coalesced:
//
//   row_gather :  out[b,c,j] = sum_{k<K} w[b,k,j] * plane[b,c, idx[b,k,j]]          (gather-type ops)
//   row_segsum :  out[b,c,t] = sum_{e in seg_b(t)} w_e * row[b,c, src_e]            (scatter-type ops, via CsrPlan)
#pragma once
#include "common.cuh"
#include "csr.cuh"

namespace pcfm {

constexpr int kMaxStageFloats = 55296;  // 216 KB of the 227 KB shared memory a CTA may use

enum SegMode { SEG_WEIGHTED = 0, SEG_UNIT = 1, SEG_MEAN = 2 };

// out[b,c,t] over a CsrPlan; src rows have length S, out rows length T, E edges per batch element.
int launch_row_segsum(SegMode mode, const float* src, const CsrPlan& plan, float* out, int b, int c, int s,
                      int t, int e, cudaStream_t st);

// K in {1,3}; w may be nullptr (weights 1, K must be 1).  planes [B,C,P], idx [B,K,J], out [B,C,J].
int launch_row_gather(int k, const float* planes, const int* idx, const float* w, float* out, int b, int c,
                      int p, int j, cudaStream_t st, bool dependent = false);

// channels per CTA for plane-staging kernels (0 = plane does not fit in shared memory)
int pick_group(int b, int c, int p);

}  // namespace pcfm
