// cellops.cu — avg_voxelize forward on the sliced-ELL mean plan:
//   out[b,c,v] = mean of the features of the points in voxel v                                  (vox.cu:48-72)
// One CTA per (batch, channel) row.  The randomly-read row (N floats) is staged in shared memory by TMA; a thread
// owns one light voxel cell and walks that cell's records with coalesced loads (cellplan.cuh), a warp owns a piece
// of a heavy cell and reads four point ids per lane and load.  (The trilinear walks live in tri.cu.)
#include "cellplan.cuh"
#include "rows.cuh"

#ifndef PCFM_VOX_THREADS
#define PCFM_VOX_THREADS 512
#endif

namespace pcfm {

constexpr int kVoxThreads = PCFM_VOX_THREADS;  // two CTAs per SM: 1024 threads leave 32 registers, 768 -> 40, 512 -> 64

// -------------------------------------------------------------------------------------------------
// avg_voxelize forward
// -------------------------------------------------------------------------------------------------
template <bool ROW_SMEM>
__global__ void __launch_bounds__(kVoxThreads, 2)
    vox_fwd_cell_kernel(const float* __restrict__ feat, const int* __restrict__ cells,
                        const int* __restrict__ slot0, const int* __restrict__ meta,
                        const int* __restrict__ heavy,
                        const int* __restrict__ rec, float* __restrict__ out, int C, int N, int T, int lcap,
                        int hcap, int rcap, int use_tma, int wait_first) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ int next_item;
  // plan reused from an earlier call: the predecessor on the stream is an arbitrary kernel (it may produce `feat` or
  // still read the memory `out` now occupies), so nothing global is touched before it has completed
  if (wait_first) griddep_wait();
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = feat + ((long long)b * C + c) * N;
  float* o = out + ((long long)b * C + c) * T;
  if (threadIdx.x == 0) {
    next_item = 0;
    if (ROW_SMEM) {
      smem[N] = 0.f;  // the pad ids of the heavy cells point here
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
  }
  __syncthreads();
  // issue the row copy, then zero the output row while it is in flight
  const bool aligned = ROW_SMEM && use_tma && ((((uintptr_t)row_g) & 15) == 0);
  const int bulk = aligned ? (N & ~3) : 0;
  if (bulk > 0 && threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)bulk * 4u);
    for (uint32_t off = 0; off < (uint32_t)bulk * 4u; off += kBulkChunk)
      bulk_g2s((char*)smem + off, (const char*)row_g + off, min(kBulkChunk, (uint32_t)bulk * 4u - off), &bar);
  }
  if ((((uintptr_t)o) & 15) == 0) {
    float4* o4 = (float4*)o;
    for (int i = threadIdx.x; i < (T >> 2); i += blockDim.x) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = (T & ~3) + threadIdx.x; i < T; i += blockDim.x) o[i] = 0.f;
  } else {
    for (int i = threadIdx.x; i < T; i += blockDim.x) o[i] = 0.f;
  }
  if (ROW_SMEM) {
    for (int i = bulk + threadIdx.x; i < N; i += blockDim.x) smem[i] = __ldg(row_g + i);
    if (bulk > 0) mbar_wait(&bar, 0);
  }
  griddep_wait();   // everything above is independent of the plan (PDL: overlaps the plan kernel)
  griddep_launch_dependents();  // chained PDL: the next kernel on the stream may be scheduled behind this one
  cells = after_wait(cells); slot0 = after_wait(slot0); meta = after_wait(meta); heavy = after_wait(heavy);
  rec = after_wait(rec);
  __syncthreads();  // row visible; zero stores ordered before the sums below (same CTA)
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const int* rc = rec + (long long)b * rcap;
  const int nslice = meta[b * kMetaInts + 0] >> 5, nitem = nslice + meta[b * kMetaInts + 1];
  const int lane = threadIdx.x & 31;
  // work items taken by whole warps from a shared counter: slices of 32 light cells (longest first), then heavy cells
  const int4* hv = (const int4*)heavy + (long long)b * hcap;
  CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
  while (cur.item < nitem) {
    const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);  // header in flight during this item
    if (cur.item < nslice) {
      const int d = cur.h.x;
      const int n = (int)((unsigned)d >> 24), t = d & 0xffffff;
      if (n != 0) {  // 0 = alignment hole
        const int s0 = cur.h.y & 0xffffff;
        const float rcp = __frcp_rn((float)n);  // == (float)(1.0/(float)cnt), vox.cu:65
        float acc = 0.f;
        // records walked 4 / 2 / 1 at a time in ascending order (independent loads in flight, immediate offsets, no
        // padded work); every product is rounded, then added: the reference's atomicAdd(f * div)
        auto feat = [&](int p) { return ROW_SMEM ? smem[p] : __ldg(row_g + p); };
        const int* rp = rc + s0;
        int i0 = 0;
        for (; i0 + 4 <= n; i0 += 4, rp += 128) {
          const int p0 = rp[0], p1 = rp[32], p2 = rp[64], p3 = rp[96];
          const float f0 = feat(p0), f1 = feat(p1), f2 = feat(p2), f3 = feat(p3);
          acc = __fadd_rn(acc, __fmul_rn(f0, rcp));
          acc = __fadd_rn(acc, __fmul_rn(f1, rcp));
          acc = __fadd_rn(acc, __fmul_rn(f2, rcp));
          acc = __fadd_rn(acc, __fmul_rn(f3, rcp));
        }
        if (i0 + 2 <= n) {
          const int p0 = rp[0], p1 = rp[32];
          const float f0 = feat(p0), f1 = feat(p1);
          acc = __fadd_rn(acc, __fmul_rn(f0, rcp));
          acc = __fadd_rn(acc, __fmul_rn(f1, rcp));
          i0 += 2;
          rp += 64;
        }
        if (i0 < n) acc = __fadd_rn(acc, __fmul_rn(feat(rp[0]), rcp));
        o[t] = acc;
      }
    } else {
      // Heavy cell piece: the whole warp; a lane reads FOUR point ids per load (the plan starts a heavy cell on a
      // 16-byte boundary and pads it with the id N, whose feature slot holds an exact zero), two loads in flight.
      const int t = cur.h.x, n4 = (cur.h.z + 3) >> 2;   // cur.h.z records in this piece, cur.h.w in the whole cell
      const int4* r4 = (const int4*)(rc + cur.h.y);
      const float rcp = __frcp_rn((float)cur.h.w);
      auto feat = [&](int p) { return ROW_SMEM ? smem[p] : (p < N ? __ldg(row_g + p) : 0.f); };
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int i = lane;
      for (; i + 32 < n4; i += 64) {
        const int4 p = __ldg(r4 + i), q = __ldg(r4 + i + 32);
        const float f0 = feat(p.x), f1 = feat(p.y), f2 = feat(p.z), f3 = feat(p.w);
        const float g0 = feat(q.x), g1 = feat(q.y), g2 = feat(q.z), g3 = feat(q.w);
        a0 += f0 + g0;
        a1 += f1 + g1;
        a2 += f2 + g2;
        a3 += f3 + g3;
      }
      if (i < n4) {
        const int4 p = __ldg(r4 + i);
        a0 += feat(p.x);
        a1 += feat(p.y);
        a2 += feat(p.z);
        a3 += feat(p.w);
      }
      const float acc = warp_sum((a0 + a1) + (a2 + a3)) * rcp;
      if (lane == 0) {
        if (cur.h.z == cur.h.w) o[t] = acc;
        else atomicAdd(o + t, acc);   // a piece of a split cell: the row was zeroed above (same CTA, before the barrier)
      }
    }
    cur = take_cell_item(nxt);
  }
}

// defined in rows.cu (A/B variants on the generic sorted-segment engine / global atomics)
int avg_voxelize_forward_generic(const float* features, const int* coords, int b, int c, int n, int r, float* out,
                                 int* ind, int* cnt, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                 int variant);

static int launch_vox_fwd(const float* features, const CellPlan& plan, int b, int c, int n, int t, float* out,
                          cudaStream_t st, bool reused) {
  dim3 grid(c, b);
  // fresh plan: programmatic dependent of the plan kernel; reused plan: chained PDL behind whatever ran before
  const bool dep = reused ? g_opt_pdl >= 2 : plan.pdl;
  PCFM_CUDA(allow_dyn_smem((const void*)vox_fwd_cell_kernel<true>, (int)(kMaxStageFloats * sizeof(float))));
  if (n < kMaxStageFloats)
    PCFM_CUDA(launch_dependent(vox_fwd_cell_kernel<true>, grid, dim3(kVoxThreads), ((size_t)n + 4) * sizeof(float), st, dep,
                               features, plan.cells, plan.slot0, plan.meta, plan.heavy, (const int*)plan.rec, out, c,
                               n, t, plan.lcap, plan.hcap, plan.rcap, (int)(g_opt_stage == 0), (int)reused));
  else
    PCFM_CUDA(launch_dependent(vox_fwd_cell_kernel<false>, grid, dim3(kVoxThreads), 0, st, dep, features, plan.cells,
                               plan.slot0, plan.meta, plan.heavy, (const int*)plan.rec, out, c, n, t, plan.lcap,
                               plan.hcap, plan.rcap, 0, (int)reused));
  PCFM_LAUNCHED();
  return 0;
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

size_t pcfm_avg_voxelize_forward_workspace(int b, int c, int n, int r) {
  (void)c;
  if (b <= 0 || n < 0 || r <= 0) return 0;
  const long long t = (long long)r * r * r;
  const size_t a = cellplan_workspace_bytes(b, n, t, false);
  const size_t s = csr_workspace_bytes(b, n, t, false) + 256;
  return a > s ? a : s;
}

int pcfm_avg_voxelize_forward(const float* features, const int* coords, int b, int c, int n, int r, float* out,
                              int* ind, int* cnt, void* workspace, size_t workspace_bytes,
                              pcfm_stream_t stream) {
  // 24-bit cell ids and record slots in the plan: r <= 256 and N below ~16.7 M points
  if (b < 0 || c < 0 || n < 0 || r <= 0 || !cellplan_fits(n, (long long)r * r * r, false)) return PCFM_E_BADARG;
  if (b == 0) return 0;
  if (!cnt || (c > 0 && !out) || (n > 0 && (!coords || !ind || (c > 0 && !features)))) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_opt_vox_fwd != 0)
    return avg_voxelize_forward_generic(features, coords, b, c, n, r, out, ind, cnt, workspace,
                                             workspace_bytes, st, g_opt_vox_fwd);
  const int t = r * r * r;
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = cellplan_build(coords, nullptr, r, ind, b, n, t, cnt, &plan, &ws, st);
  if (rc || c == 0) return rc;
  return launch_vox_fwd(features, plan, b, c, n, t, out, st, false);
}

int pcfm_avg_voxelize_forward_dual(const float* features, const int* coords, const float* float_coords, int b, int c,
                                   int n, int r, float* out, int* ind, int* cnt, void* workspace,
                                   size_t workspace_bytes, void* tri_workspace, size_t tri_workspace_bytes,
                                   pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0 || !cellplan_fits(n, (long long)r * r * r, true) ||
      !cellplan_fits(n, (long long)r * r * r, false))
    return PCFM_E_BADARG;
  if (b == 0) return 0;
  if (!cnt || (c > 0 && !out) || (n > 0 && (!coords || !float_coords || !ind || (c > 0 && !features))))
    return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int t = r * r * r;
  Workspace ws(workspace, workspace_bytes), wst(tri_workspace, tri_workspace_bytes);
  CellPlan plan, tri;
  int rc = cellplan_build_dual(coords, float_coords, r, ind, b, n, t, cnt, &plan, &ws, &tri, &wst, st);
  if (rc || c == 0) return rc;
  return launch_vox_fwd(features, plan, b, c, n, t, out, st, false);
}

int pcfm_avg_voxelize_forward_planned(const float* features, int b, int c, int n, int r, float* out,
                                      const void* plan_workspace, size_t plan_workspace_bytes,
                                      pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0 || !cellplan_fits(n, (long long)r * r * r, false)) return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!out || !plan_workspace || (n > 0 && !features)) return PCFM_E_BADARG;
  const int t = r * r * r;
  Workspace ws(const_cast<void*>(plan_workspace), plan_workspace_bytes);
  CellPlan plan;
  int rc = cellplan_attach(false, b, n, t, &plan, &ws);
  if (rc) return rc;
  return launch_vox_fwd(features, plan, b, c, n, t, out, (cudaStream_t)stream, true);
}

}  // extern "C"
