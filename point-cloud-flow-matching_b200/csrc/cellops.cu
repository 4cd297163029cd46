// cellops.cu — the two scatter-type PVConv primitives on the sliced-ELL cell plan:
//   avg_voxelize forward            out[b,c,v]  = mean of the features of the points in voxel v
//   trilinear_devoxelize backward   grad[b,c,v] = sum over the 8 cells around v of sum_p w_corner(p) * grad_y[b,c,p]
// One CTA per (batch, channel) row.  The randomly-read row (N floats) is staged in shared memory by TMA; a thread
// owns one voxel cell and walks that cell's records with coalesced loads (cellplan.cuh), so a point is visited
// ONCE per channel (the atomic formulation of the reference visits it 8 times for devoxelize backward).
#include "cellplan.cuh"
#include "rows.cuh"

#ifndef PCFM_HEAVY_U
#define PCFM_HEAVY_U 1
#endif
#ifndef PCFM_HEAVY_V
#define PCFM_HEAVY_V 4
#endif

namespace pcfm {

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
  if (lane == 0) c.item = atomicAdd(counter, 1);
  c.item = __shfl_sync(0xffffffffu, c.item, 0);
  c.h = make_int4(0, 0, 0, 0);
  if (c.item < nslice) {
    c.h.x = cl[(c.item << 5) | lane];
    c.h.y = sl[(c.item << 5) | lane];
  } else if (c.item < nitem) {
    c.h = hv[c.item - nslice];
  }
  return c;
}

// -------------------------------------------------------------------------------------------------
// avg_voxelize forward
// -------------------------------------------------------------------------------------------------
template <bool ROW_SMEM>
__global__ void __launch_bounds__(1024, 2)
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
      const int t = cur.h.x, s0 = cur.h.y, n = cur.h.z;   // n records in this piece, cur.h.w in the whole cell
      const float rcp = __frcp_rn((float)cur.h.w);
      float acc = 0.f;
      for (int i = lane; i < n; i += 32 * PCFM_HEAVY_U) {  // PCFM_HEAVY_U independent record loads in flight per lane
        int p[PCFM_HEAVY_U];
#pragma unroll
        for (int k = 0; k < PCFM_HEAVY_U; ++k) p[k] = rc[s0 + min(i + 32 * k, n - 1)];
#pragma unroll
        for (int k = 0; k < PCFM_HEAVY_U; ++k) {
          float f = ROW_SMEM ? smem[p[k]] : __ldg(row_g + p[k]);
          f = (i + 32 * k < n) ? f : 0.f;  // a padded slot contributes an exact +0
          acc = __fadd_rn(acc, __fmul_rn(f, rcp));
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        if (n == cur.h.w) o[t] = acc;
        else atomicAdd(o + t, acc);   // a piece of a split cell: the row was zeroed above (same CTA, before the barrier)
      }
    }
    cur = nxt;
  }
}

// -------------------------------------------------------------------------------------------------
// trilinear_devoxelize backward
// -------------------------------------------------------------------------------------------------
// 8 corner sums of one point: s[dx,dy,dz] += (x_dx*y_dy) * (z_dz*g).  (The reference rounds w = (x*y)*z, then w*g,
// then adds; this factorisation differs from it by ~1 ulp per term, far inside the 1e-5 bar, and halves the work.)
__device__ __forceinline__ void tri_accumulate(float fx, float fy, float fz, float g, float* s) {
  const float x0 = 1.0f - fx, y0 = 1.0f - fy, z0 = 1.0f - fz;
  const float gz0 = g * z0, gz1 = g * fz;
  const float a00 = x0 * y0, a01 = x0 * fy, a10 = fx * y0, a11 = fx * fy;
  s[0] = fmaf(a00, gz0, s[0]);
  s[1] = fmaf(a00, gz1, s[1]);
  s[2] = fmaf(a01, gz0, s[2]);
  s[3] = fmaf(a01, gz1, s[3]);
  s[4] = fmaf(a10, gz0, s[4]);
  s[5] = fmaf(a10, gz1, s[5]);
  s[6] = fmaf(a11, gz0, s[6]);
  s[7] = fmaf(a11, gz1, s[7]);
}

// PLANE_SMEM = false (grids whose plane does not fit in shared memory, r >= 38): the cell sums are added straight into
// the caller-zeroed output with global atomics (RED.ADD.F32) — 8 per CELL, where the reference issues 8 per point.
template <bool ROW_SMEM, bool PLANE_SMEM>
__global__ void __launch_bounds__(1024, 1)
    devox_bwd_cell_kernel(const float* __restrict__ grad_y, const int* __restrict__ cells,
                          const int* __restrict__ slot0, const int* __restrict__ meta,
                          const int* __restrict__ heavy, const float4* __restrict__ rec,
                          float* __restrict__ grad_x, int C, int N, int r, int lcap, int hcap, int rcap,
                          int use_tma, int wait_first) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ int next_item;
  if (wait_first) griddep_wait();  // reused plan: see vox_fwd_cell_kernel
  const int r2 = r * r, P = r2 * r;
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = grad_y + ((long long)b * C + c) * N;
  float* o = grad_x + ((long long)b * C + c) * P;
  float* plane = PLANE_SMEM ? smem : o;                   // [P]
  float* row = smem + (PLANE_SMEM ? ((P + 3) & ~3) : 0);  // [N] (ROW_SMEM)
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    next_item = 0;
  }
  __syncthreads();
  const bool aligned = ROW_SMEM && use_tma && ((((uintptr_t)row_g) & 15) == 0);
  const int bulk = aligned ? (N & ~3) : 0;
  if (bulk > 0 && threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)bulk * 4u);
    for (uint32_t off = 0; off < (uint32_t)bulk * 4u; off += kBulkChunk)
      bulk_g2s((char*)row + off, (const char*)row_g + off, min(kBulkChunk, (uint32_t)bulk * 4u - off), &bar);
  }
  for (int i = threadIdx.x; PLANE_SMEM && i < ((P + 3) >> 2); i += blockDim.x)  // overlaps the row copy
    ((float4*)plane)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ROW_SMEM) {
    for (int i = bulk + threadIdx.x; i < N; i += blockDim.x) row[i] = __ldg(row_g + i);
    if (bulk > 0) mbar_wait(&bar, 0);
  }
  griddep_wait();  // everything above is independent of the plan (PDL: overlaps the plan kernel)
  griddep_launch_dependents();  // chained PDL: the next kernel on the stream may be scheduled behind this one
  __syncthreads();
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const float4* rc = rec + (long long)b * rcap;
  const int l1 = meta[b * kMetaInts + 0], h1 = meta[b * kMetaInts + 1];
  const int lane = threadIdx.x & 31;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1);
  // Work items, taken from a shared counter by whole warps (dynamic: the static round-robin left most warps waiting
  // at the final barrier for the one that drew the longest cells — 3.5-5.6 stalled warps per issue slot in ncu):
  //   [0, nslice)           slices of 32 light cells, longest cells first (count-sorted plan)
  //   [nslice, nslice + h1) pieces (<= kHeavyChunk records) of heavy cells
  const int nslice = l1 >> 5, nitem = nslice + h1;
  const int4* hv = (const int4*)heavy + (long long)b * hcap;
  CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
  while (cur.item < nitem) {
    const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);  // header in flight during this item
    if (cur.item < nslice) {
      // Light cells (sliced-ELL): one thread per cell sums the 8 corner contributions of its points in registers,
      // then adds them into the plane.  Neighbouring cells share voxels, so the add is a shared-memory atomicAdd
      // (a CAS loop on sm_100a; collisions need two adjacent cells to finish in the same few cycles).
      const int d = cur.h.x;
      const int n = (int)((unsigned)d >> 24);  // 0 = alignment hole
      const int cell = d & 0xffffff;
      const unsigned se = (unsigned)cur.h.y;
      const int s0 = se & 0xffffff;
      const unsigned mask = n ? se >> 24 : 0u;
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      // records are walked 4 / 2 / 1 at a time: independent loads in flight, no padded work (slices are count-sorted,
      // so the lanes of a warp take the same path except at slice boundaries)
      int i0 = 0;
      for (; i0 + 4 <= n; i0 += 4) {
        float4 qr[4];
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) qr[u] = rc[s0 + 32 * (i0 + u)];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = __float_as_int(qr[u].x);
          g[u] = ROW_SMEM ? row[pp] : __ldg(row_g + pp);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) tri_accumulate(qr[u].y, qr[u].z, qr[u].w, g[u], s);
      }
      if (i0 + 2 <= n) {
        const float4 qa = rc[s0 + 32 * i0], qb = rc[s0 + 32 * (i0 + 1)];
        const int pa = __float_as_int(qa.x), pb = __float_as_int(qb.x);
        const float ga = ROW_SMEM ? row[pa] : __ldg(row_g + pa);
        const float gb = ROW_SMEM ? row[pb] : __ldg(row_g + pb);
        tri_accumulate(qa.y, qa.z, qa.w, ga, s);
        tri_accumulate(qb.y, qb.z, qb.w, gb, s);
        i0 += 2;
      }
      if (i0 < n) {
        const float4 qa = rc[s0 + 32 * i0];
        const int pa = __float_as_int(qa.x);
        const float ga = ROW_SMEM ? row[pa] : __ldg(row_g + pa);
        tri_accumulate(qa.y, qa.z, qa.w, ga, s);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((mask >> k) & 1u) atomicAdd(plane + cell + off[k], s[k]);
    } else {
      // Heavy cell piece: the whole warp; the 8 partial sums are reduced by recursive halving (9 shuffles instead
      // of 40): after the exchange lane l holds the total of corner (l >> 2) & 7.
      const int4 hd = cur.h;
      const int cell = hd.x, s0 = hd.y, n = hd.z;
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < n; i += 32 * PCFM_HEAVY_V) {  // PCFM_HEAVY_V independent record loads in flight per lane
        float4 q[PCFM_HEAVY_V];
#pragma unroll
        for (int k = 0; k < PCFM_HEAVY_V; ++k) q[k] = rc[s0 + min(i + 32 * k, n - 1)];
#pragma unroll
        for (int k = 0; k < PCFM_HEAVY_V; ++k) {
          const int pk = __float_as_int(q[k].x);
          float gk = ROW_SMEM ? row[pk] : __ldg(row_g + pk);
          gk = (i + 32 * k < n) ? gk : 0.f;  // a padded slot adds exact zeros
          tri_accumulate(q[k].y, q[k].z, q[k].w, gk, s);
        }
      }
      float a4[4], a2[2], a1;
      const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // lanes 0-15 keep corners 0-3, lanes 16-31 keep corners 4-7
        const float send = up16 ? s[k] : s[k + 4];
        const float keep = up16 ? s[k + 4] : s[k];
        a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = up8 ? a4[k] : a4[k + 2];
        const float keep = up8 ? a4[k + 2] : a4[k];
        a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const float send = up4 ? a2[0] : a2[1];
        const float keep = up4 ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
      if ((lane & 3) == 0) {
        const int k = lane >> 2;  // bit 4 of the lane -> dx, bit 3 -> dy, bit 2 -> dz
        if ((hd.w >> k) & 1) atomicAdd(plane + cell + (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1), a1);
      }
    }
    cur = nxt;
  }
  if (!PLANE_SMEM) return;
  __syncthreads();
  // ---- write the plane: one bulk (TMA) store when aligned, else a coalesced copy ----
  if ((((uintptr_t)o) & 15) == 0 && (P & 3) == 0 && use_tma) {
    fence_proxy_async();  // generic-proxy writes to `plane` -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t off2 = 0; off2 < (uint32_t)P * 4u; off2 += kBulkChunk)
        bulk_s2g((char*)o + off2, (const char*)plane + off2, min(kBulkChunk, (uint32_t)P * 4u - off2));
      bulk_commit();
      bulk_wait_read();  // the CTA may exit once the source has been read out of shared memory
    }
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) o[i] = plane[i];
  }
}

static size_t devox_bwd_smem(int n, int r, bool row_smem, bool plane_smem = true) {
  const size_t p = plane_smem ? (size_t)r * r * r : 0;
  return (((p + 3) & ~(size_t)3) + (row_smem ? n : 0)) * sizeof(float);
}
constexpr size_t kSmemLimit = 227 * 1024 - 512;  // opt-in maximum (232448 B) minus the static shared memory

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
  if (n <= kMaxStageFloats)
    PCFM_CUDA(launch_dependent(vox_fwd_cell_kernel<true>, grid, dim3(1024), (size_t)n * sizeof(float), st, dep,
                               features, plan.cells, plan.slot0, plan.meta, plan.heavy, (const int*)plan.rec, out, c,
                               n, t, plan.lcap, plan.hcap, plan.rcap, (int)(g_opt_stage == 0), (int)reused));
  else
    PCFM_CUDA(launch_dependent(vox_fwd_cell_kernel<false>, grid, dim3(1024), 0, st, dep, features, plan.cells,
                               plan.slot0, plan.meta, plan.heavy, (const int*)plan.rec, out, c, n, t, plan.lcap,
                               plan.hcap, plan.rcap, 0, (int)reused));
  PCFM_LAUNCHED();
  return 0;
}

// coords != nullptr: build the trilinear plan into `workspace`, else attach to the one an earlier call left there
static int devox_bwd_cells(const float* grad_y, const float* coords, int b, int c, int n, int r, float* grad_x,
                           void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int t = r * r * r;
  const bool reused = coords == nullptr;
  const bool plane_smem = devox_bwd_smem(0, r, false) <= kSmemLimit;  // r <= 37
  if (n == 0 || !plane_smem) {
    PCFM_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)b * c * t * sizeof(float), st));
    if (n == 0) return 0;
  }
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = reused ? cellplan_attach(true, b, n, t, &plan, &ws)
                  : cellplan_build(nullptr, coords, r, nullptr, b, n, t, nullptr, &plan, &ws, st);
  if (rc) return rc;
  PCFM_CUDA(cudaFuncSetAttribute(devox_bwd_cell_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemLimit));
  PCFM_CUDA(cudaFuncSetAttribute(devox_bwd_cell_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemLimit));
  PCFM_CUDA(cudaFuncSetAttribute(devox_bwd_cell_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemLimit));
  dim3 grid(c, b);
  const bool row_smem = devox_bwd_smem(n, r, true, plane_smem) <= kSmemLimit;
  const size_t smem = devox_bwd_smem(n, r, row_smem, plane_smem);
  // small footprints (r <= 16, or no plane): two CTAs of 512 threads per SM overlap one CTA's staging with the
  // other's compute
  const int threads = (row_smem && smem <= 110 * 1024) ? 512 : 1024;
  const int tma = (int)(g_opt_stage == 0);
  // a reused plan behind a memset must not start early (the kernel adds to the zeroed grid with global atomics)
  const bool dep = reused ? (g_opt_pdl >= 2 && plane_smem) : plan.pdl;
#define PCFM_DBWD(RS, PS)                                                                                         \
  PCFM_CUDA(launch_dependent(devox_bwd_cell_kernel<RS, PS>, grid, dim3(threads), smem, st, dep, grad_y,            \
                             plan.cells, plan.slot0, plan.meta, plan.heavy, (const float4*)plan.rec, grad_x, c, n, \
                             r, plan.lcap, plan.hcap, plan.rcap, tma, (int)reused))
  if (plane_smem && row_smem) PCFM_DBWD(true, true);
  else if (plane_smem) PCFM_DBWD(false, true);
  else if (row_smem) PCFM_DBWD(true, false);
  else PCFM_DBWD(false, false);
#undef PCFM_DBWD
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
  if (b < 0 || c < 0 || n < 0 || r <= 0 || (long long)r * r * r > (1 << 24)) return PCFM_E_BADARG;
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
  if (b < 0 || c < 0 || n < 0 || r <= 0 || (long long)r * r * r > (1 << 24)) return PCFM_E_BADARG;
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
  if (b < 0 || c < 0 || n < 0 || r <= 0 || (long long)r * r * r > (1 << 24)) return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!out || !plan_workspace || (n > 0 && !features)) return PCFM_E_BADARG;
  const int t = r * r * r;
  Workspace ws(const_cast<void*>(plan_workspace), plan_workspace_bytes);
  CellPlan plan;
  int rc = cellplan_attach(false, b, n, t, &plan, &ws);
  if (rc) return rc;
  return launch_vox_fwd(features, plan, b, c, n, t, out, (cudaStream_t)stream, true);
}

size_t pcfm_trilinear_devoxelize_backward_coords_workspace(int b, int c, int n, int r) {
  (void)c;
  if (b <= 0 || n < 0 || r <= 0) return 0;
  return cellplan_workspace_bytes(b, n, (long long)r * r * r, true);
}

int pcfm_trilinear_devoxelize_backward_coords(const float* grad_y, const float* coords, int b, int c, int n, int r,
                                              float* grad_x, void* workspace, size_t workspace_bytes,
                                              pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!grad_x || (n > 0 && (!grad_y || !coords))) return PCFM_E_BADARG;
  if ((long long)r * r * r > (1 << 24)) return PCFM_E_BADARG;  // 24-bit cell ids in the plan (r <= 256)
  return devox_bwd_cells(grad_y, coords, b, c, n, r, grad_x, workspace, workspace_bytes, (cudaStream_t)stream);
}

int pcfm_trilinear_devoxelize_backward_planned(const float* grad_y, int b, int c, int n, int r, float* grad_x,
                                               const void* plan_workspace, size_t plan_workspace_bytes,
                                               pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!grad_x || !plan_workspace || (n > 0 && !grad_y)) return PCFM_E_BADARG;
  if ((long long)r * r * r > (1 << 24)) return PCFM_E_BADARG;
  return devox_bwd_cells(grad_y, nullptr, b, c, n, r, grad_x, const_cast<void*>(plan_workspace),
                         plan_workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
