// tri.cu — the two trilinear PVConv primitives as walks over the trilinear cell plan (cellplan.cuh):
//   trilinear_devoxelize forward    out[b,c,p]  = sum_k w_k(p) * vox[b,c, cell(p) + corner_k]        (trilinear_devox.cu:38-103)
//   trilinear_devoxelize backward   grad[b,c,v] = sum over the 8 cells around v of sum_p w_corner(p) * grad_y[b,c,p]
//                                                                                                  (trilinear_devox.cu:119-162)
// One CTA per (batch, channel) row.  The side that is indexed by POINT (the output row of the forward, the gradient
// row of the backward: N floats, random access) lives in shared memory; the points are visited cell by cell — a thread
// owns a light cell of a count-sorted slice, a warp owns a piece of a heavy cell — so the 8 corner values of a cell are
// read (forward) or added (backward) ONCE per cell and stay in registers while its points stream by.  The plan stores
// PAIRS of points (two float4 per pair): both walks run on packed fp32, two points per FFMA2.
#include "cellplan.cuh"
#include "rows.cuh"

namespace pcfm {

constexpr size_t kTriSmemLimit = 227 * 1024 - 512;  // opt-in maximum (232448 B) minus the static shared memory

// ---- the two points of a pair slot -------------------------------------------------------------------------------
struct TriPair {
  int off0, off1;          // byte offsets of the two points in a channel row (4*N = the dummy half of an odd cell)
  f32x2_t fx, fy, fz;      // fractions of both points
};
__device__ __forceinline__ TriPair tri_pair(const float4 a, const float4 b) {
  TriPair t;
  t.off0 = __float_as_int(a.x);
  t.off1 = __float_as_int(a.y);
  t.fx = pk2(a.z, a.w);
  t.fy = pk2(b.x, b.y);
  t.fz = pk2(b.z, b.w);
  return t;
}

// backward: s[dx,dy,dz] += (x_dx * y_dy) * (z_dz * g) for both points.  (The reference rounds w = (x*y)*z, then w*g, then
// adds; this factorisation differs by ~1 ulp per term, far inside the 1e-5 bar.)
__device__ __forceinline__ void tri_bwd_accumulate(const TriPair& t, float g0, float g1, f32x2_t* s) {
  const f32x2_t one = pk2(1.0f, 1.0f);
  const f32x2_t g = pk2(g0, g1);
  const f32x2_t x0 = sub2(one, t.fx), y0 = sub2(one, t.fy), z0 = sub2(one, t.fz);
  const f32x2_t gz0 = mul2(g, z0), gz1 = mul2(g, t.fz);
  const f32x2_t a00 = mul2(x0, y0), a01 = mul2(x0, t.fy), a10 = mul2(t.fx, y0), a11 = mul2(t.fx, t.fy);
  s[0] = fma2(a00, gz0, s[0]);
  s[1] = fma2(a00, gz1, s[1]);
  s[2] = fma2(a01, gz0, s[2]);
  s[3] = fma2(a01, gz1, s[3]);
  s[4] = fma2(a10, gz0, s[4]);
  s[5] = fma2(a10, gz1, s[5]);
  s[6] = fma2(a11, gz0, s[6]);
  s[7] = fma2(a11, gz1, s[7]);
}

// forward: both points' interpolated values from the cell's 8 corner values v[k] (each duplicated into a pair),
// reduced along z, then y, then x
__device__ __forceinline__ f32x2_t tri_fwd_interpolate(const TriPair& t, const f32x2_t* v) {
  const f32x2_t one = pk2(1.0f, 1.0f);
  const f32x2_t x0 = sub2(one, t.fx), y0 = sub2(one, t.fy), z0 = sub2(one, t.fz);
  const f32x2_t t00 = fma2(t.fz, v[1], mul2(z0, v[0]));
  const f32x2_t t01 = fma2(t.fz, v[3], mul2(z0, v[2]));
  const f32x2_t t10 = fma2(t.fz, v[5], mul2(z0, v[4]));
  const f32x2_t t11 = fma2(t.fz, v[7], mul2(z0, v[6]));
  const f32x2_t u0 = fma2(t.fy, t01, mul2(y0, t00));
  const f32x2_t u1 = fma2(t.fy, t11, mul2(y0, t10));
  return fma2(t.fx, u1, mul2(x0, u0));
}

// Walks pair slots first, first + step, ... < end of one cell (or of one lane's share of a heavy piece), two slots per
// round (j and j + second), and hands every pair to `body`.  Slot j sits at pa[j * mul] / pb[j * mul] (light cell of a
// sliced-ELL slice: mul = 32, second = 1, step = 2; lane of a heavy piece: mul = 1, second = 32, step = 64, first = lane).
// PIPE (off: measured 5-10 % slower, the register rotation costs ~20 moves per round and the kernels are issue-bound):
// the loads of round i+1 are issued before round i is processed.  A missing second slot is replaced by a dummy pair (offsets 4*N, zero fractions): it gathers the exact zero behind the row / stores to
// the scratch slot behind it.
#ifndef PCFM_TRI_PIPE
#define PCFM_TRI_PIPE 0
#endif
template <typename Body>
__device__ __forceinline__ void tri_walk(const float4* __restrict__ pa, const float4* __restrict__ pb, int first, int end,
                                         int step, int second, int mul, int dummy, Body&& body) {
  const float4 da = make_float4(__int_as_float(dummy), __int_as_float(dummy), 0.f, 0.f);
  const float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = first;
#if PCFM_TRI_PIPE
  float4 a0 = da, b0 = db, a1 = da, b1 = db;
  if (j < end) {
    a0 = __ldg(pa + j * mul);
    b0 = __ldg(pb + j * mul);
    if (j + second < end) {
      a1 = __ldg(pa + (j + second) * mul);
      b1 = __ldg(pb + (j + second) * mul);
    }
  }
  while (j < end) {
    const int jn = j + step;
    float4 na0 = da, nb0 = db, na1 = da, nb1 = db;
    if (jn < end) {
      na0 = __ldg(pa + jn * mul);
      nb0 = __ldg(pb + jn * mul);
      if (jn + second < end) {
        na1 = __ldg(pa + (jn + second) * mul);
        nb1 = __ldg(pb + (jn + second) * mul);
      }
    }
    body(tri_pair(a0, b0));
    body(tri_pair(a1, b1));
    a0 = na0; b0 = nb0; a1 = na1; b1 = nb1;
    j = jn;
  }
#else
  for (; j < end; j += step) {
    const bool two = j + second < end;
    const float4 a0 = __ldg(pa + j * mul), b0 = __ldg(pb + j * mul);
    const float4 a1 = two ? __ldg(pa + (j + second) * mul) : da, b1 = two ? __ldg(pb + (j + second) * mul) : db;
    body(tri_pair(a0, b0));
    body(tri_pair(a1, b1));
  }
#endif
}

// =================================================================================================
// backward
// =================================================================================================
// PLANE_SMEM = false (grids whose plane does not fit in shared memory, r >= 38): the cell sums are added straight into
// the caller-zeroed output with global atomics (RED.ADD.F32) — 8 per CELL, where the reference issues 8 per point.
// Shared memory: [plane: P floats, rounded up to 4][row: N floats][one zero float: the dummy half of an odd cell].
template <bool ROW_SMEM, bool PLANE_SMEM>
__global__ void __launch_bounds__(1024, 1)
    tri_bwd_kernel(const float* __restrict__ grad_y, const int* __restrict__ cells, const int* __restrict__ slot0,
                   const int* __restrict__ meta, const int* __restrict__ heavy, const float4* __restrict__ recA,
                   const float4* __restrict__ recB, float* __restrict__ grad_x, int C, int N, int r, int lcap,
                   int hcap, int rcap, int use_tma, int wait_first, int phases) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ int next_item;
  // plan reused from an earlier call: the predecessor on the stream is an arbitrary kernel (it may produce `grad_y` or
  // still read the memory `grad_x` now occupies), so nothing global is touched before it has completed
  if (wait_first) griddep_wait();
  const int r2 = r * r, P = r2 * r;
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = grad_y + ((long long)b * C + c) * N;
  float* o = grad_x + ((long long)b * C + c) * P;
  float* plane = PLANE_SMEM ? smem : o;                   // [P]
  float* row = smem + (PLANE_SMEM ? ((P + 3) & ~3) : 0);  // [N + 1] (ROW_SMEM)
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    next_item = 0;
    if (ROW_SMEM) row[N] = 0.f;
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
  cells = after_wait(cells); slot0 = after_wait(slot0); meta = after_wait(meta); heavy = after_wait(heavy);
  recA = after_wait(recA); recB = after_wait(recB);
  __syncthreads();
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const float4* ra = recA + (long long)b * rcap;
  const float4* rb = recB + (long long)b * rcap;
  const int l1 = meta[b * kMetaInts + 0], h1 = meta[b * kMetaInts + 1];
  const int lane = threadIdx.x & 31;
  const int dummy = 4 * N;
  const char* rowb = (const char*)row;
  const char* rowg = (const char*)row_g;
  auto gather = [&](int off) -> float {
    if (ROW_SMEM) return *(const float*)(rowb + off);
    return off != dummy ? __ldg((const float*)(rowg + off)) : 0.f;
  };
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1);
  const int nslice = l1 >> 5, nitem = nslice + h1;
  const int4* hv = (const int4*)heavy + (long long)b * hcap;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5;
  // ---- walk: whole warps take work items from a shared counter: slices of 32 light cells (one thread per cell, the 8
  // corner sums of its points built in registers), then pieces of heavy cells (one warp per piece).  Neighbouring cells
  // share voxels, so a cell's sums must be ADDED to the plane with care.  A shared-memory float atomicAdd is a CAS loop
  // whose 32 lanes are served one by one (ncu: ~30 wavefronts per warp instruction, two thirds of the kernel's
  // shared-memory traffic).  Two forms, see `phases` below. ----
  // phases == 0: light cells add their sums with shared-memory atomics (heavy pieces always do: 8 per piece).
  // phases == 2 (pcfm_set_option("bwd_phases", 2)): a cell's 8 sums are parked in this CTA's own OUTPUT row (P floats of global memory that nothing
  // reads until the plane is stored at the very end: room for P/8 cells) instead of being added with atomics; when all
  // walks are done, every thread fetches the sums of up to kDeferCells cells and the CTA adds corner 0 of all cells,
  // then corner 1, ...: cell -> cell + offset(k) is injective, so inside a phase no two threads touch the same voxel and
  // the add is a plain load / add / store (8 barriers per row, perfectly balanced, and a voxel's light contributions
  // arrive in a fixed order).  Slices beyond the room use atomics.  Measured at B=8, N=20000: no faster than the atomics
  // (r=32: 0.141 vs 0.140 ms alone, 0.148 vs 0.137 inside the bench step; r=16: 0.154 vs 0.138) — the kernel is bound by
  // the latency chain of its longest slice, not by the atomics' wavefronts; kept as an option for its fixed summation
  // order.
  // (Phases after every ROUND of statically dealt slices were measured too: the rounds walk in lock step, 0-15 % slower.)
  constexpr int kDeferCells = 4;
  const bool defer_ok = PLANE_SMEM && phases == 2 && ((((uintptr_t)o) & 15) == 0);
  const int ndefer = defer_ok ? min(nslice, min(P >> 8, (int)(blockDim.x * kDeferCells) >> 5)) : 0;
  CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
  while (cur.item < nitem) {
    const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);  // header in flight during this item
    f32x2_t s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0ull;
    if (cur.item < nslice) {
      const int d = cur.h.x;
      const int n = (int)((unsigned)d >> 24);  // 0 = alignment hole
      const int cell = d & 0xffffff;
      const unsigned se = (unsigned)cur.h.y;
      const unsigned mask = n ? se >> 24 : 0u;
      const int np = (n + 1) >> 1;
      tri_walk(ra + (se & 0xffffff), rb + (se & 0xffffff), 0, np, 2, 1, 32, dummy,
               [&](const TriPair t) { tri_bwd_accumulate(t, gather(t.off0), gather(t.off1), s); });
      float sk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float lo, hi;
        upk2(s[k], lo, hi);
        sk[k] = lo + hi;
      }
      if (cur.item < ndefer) {   // deferred: the 8 sums of this cell wait in the (not yet written) output row
        float4* sc = (float4*)o + 2 * ((cur.item << 5) | lane);
        sc[0] = make_float4(sk[0], sk[1], sk[2], sk[3]);
        sc[1] = make_float4(sk[4], sk[5], sk[6], sk[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if ((mask >> k) & 1u) atomicAdd(plane + cell + off[k], sk[k]);
      }
    } else {
      // Heavy cell piece: the whole warp, lanes stride the pair slots; the 8 partial sums are reduced by recursive
      // halving (9 shuffles instead of 40): after the exchange lane l holds the total of corner (l >> 2) & 7.
      const int4 hd = cur.h;
      const int cell = hd.x, np = hd.z;
      tri_walk(ra + hd.y, rb + hd.y, lane, np, 64, 32, 1, dummy,
               [&](const TriPair t) { tri_bwd_accumulate(t, gather(t.off0), gather(t.off1), s); });
      float sk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float lo, hi;
        upk2(s[k], lo, hi);
        sk[k] = lo + hi;
      }
      float a4[4], a2[2], a1;
      const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // lanes 0-15 keep corners 0-3, lanes 16-31 keep corners 4-7
        const float send = up16 ? sk[k] : sk[k + 4];
        const float keep = up16 ? sk[k + 4] : sk[k];
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
    cur = take_cell_item(nxt);
  }
  if (!PLANE_SMEM) return;
  __syncthreads();
  if (ndefer > 0) {   // uniform over the CTA
    float dsum[kDeferCells][8];
    int dcell[kDeferCells];
    unsigned dmask[kDeferCells];
#pragma unroll
    for (int c = 0; c < kDeferCells; ++c) {
      const int i = threadIdx.x + c * blockDim.x;
      dmask[c] = 0u;
      dcell[c] = 0;
      if (i < (ndefer << 5)) {
        const int hx = cl[i];
        const int n = (int)((unsigned)hx >> 24);
        dcell[c] = hx & 0xffffff;
        dmask[c] = n ? (unsigned)sl[i] >> 24 : 0u;
        const float4 u = __ldcg((const float4*)o + 2 * i), w = __ldcg((const float4*)o + 2 * i + 1);
        dsum[c][0] = u.x; dsum[c][1] = u.y; dsum[c][2] = u.z; dsum[c][3] = u.w;
        dsum[c][4] = w.x; dsum[c][5] = w.y; dsum[c][6] = w.z; dsum[c][7] = w.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int c = 0; c < kDeferCells; ++c)
        if ((dmask[c] >> k) & 1u) plane[dcell[c] + off[k]] += dsum[c][k];
      __syncthreads();
    }
  }
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

static size_t tri_bwd_smem(int n, int r, bool row_smem, bool plane_smem = true) {
  const size_t p = plane_smem ? (size_t)r * r * r : 0;
  return (((p + 3) & ~(size_t)3) + (row_smem ? n + 1 : 0)) * sizeof(float) + 16;
}

// coords != nullptr: build the trilinear plan into `workspace`, else attach to the one an earlier call left there
static int devox_bwd_cells(const float* grad_y, const float* coords, int b, int c, int n, int r, float* grad_x,
                           void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int t = r * r * r;
  const bool reused = coords == nullptr;
  const bool plane_smem = tri_bwd_smem(0, r, false) <= kTriSmemLimit;  // r <= 37
  if (n == 0 || !plane_smem) {
    PCFM_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)b * c * t * sizeof(float), st));
    if (n == 0) return 0;
  }
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = reused ? cellplan_attach(true, b, n, t, &plan, &ws)
                  : cellplan_build(nullptr, coords, r, nullptr, b, n, t, nullptr, &plan, &ws, st);
  if (rc) return rc;
  dim3 grid(c, b);
  const bool row_smem = tri_bwd_smem(n, r, true, plane_smem) <= kTriSmemLimit;
  const size_t smem = tri_bwd_smem(n, r, row_smem, plane_smem);
  // small footprints (r <= 16, or no plane): two CTAs of 512 threads per SM overlap one CTA's staging with the
  // other's compute
  const int threads = (row_smem && smem <= 110 * 1024) ? 512 : 1024;
  const int tma = (int)(g_opt_stage == 0);
  // a reused plan behind a memset must not start early (the kernel adds to the zeroed grid with global atomics)
  const bool dep = reused ? (g_opt_pdl >= 2 && plane_smem) : plan.pdl;
  // light-cell adds: option bwd_phases = 0 -> shared-memory atomics (default: never slower), 2 = deferred conflict-free
  // phases (fixed summation order of the light cells)
  const int phases = g_opt_bwd_phases == 2 ? 2 : 0;
#define PCFM_DBWD(RS, PS)                                                                                        \
  do {                                                                                                           \
    PCFM_CUDA(allow_dyn_smem((const void*)tri_bwd_kernel<RS, PS>, (int)kTriSmemLimit));                          \
    PCFM_CUDA(launch_dependent(tri_bwd_kernel<RS, PS>, grid, dim3(threads), smem, st, dep, grad_y, plan.cells,   \
                               plan.slot0, plan.meta, plan.heavy, (const float4*)plan.rec,                       \
                               (const float4*)plan.rec2, grad_x, c, n, r, plan.lcap, plan.hcap, plan.rcap, tma,  \
                               (int)reused, phases));                                                            \
  } while (0)
  if (plane_smem && row_smem) PCFM_DBWD(true, true);
  else if (plane_smem) PCFM_DBWD(false, true);
  else if (row_smem) PCFM_DBWD(true, false);
  else PCFM_DBWD(false, false);
#undef PCFM_DBWD
  PCFM_LAUNCHED();
  return 0;
}

// =================================================================================================
// forward
// =================================================================================================
// Shared memory: [voxel plane: P floats (PLANE_SMEM)][output row: N floats][scratch: the dummy half of an odd cell].
// The lanes of a light slice own 32 unrelated cells (the plan sorts cells by size): read from global memory their 8
// corner loads touch 32 different lines each, and ncu showed those loads filling the L1 data pipe (63 % at r = 16).
// PLANE_SMEM: small planes (r <= 20: <= 32 KB) are staged by one TMA bulk copy and the corners come from shared
// memory (~3 wavefronts per gather instead of 32); two CTAs of 512 threads still share an SM at N = 20000.  Larger
// planes are read straight from global memory (L1-cached).
// RT > 0: the grid resolution at compile time (8 / 16 / 32: the 8 corner offsets become immediates of the loads).
template <int RT, bool PLANE_SMEM>
__global__ void __launch_bounds__(512, 2)
    tri_fwd_kernel(const float* __restrict__ vox, const int* __restrict__ cells, const int* __restrict__ slot0,
                   const int* __restrict__ meta, const int* __restrict__ heavy, const float4* __restrict__ recA,
                   const float4* __restrict__ recB, float* __restrict__ outs, int C, int N, int r, int lcap, int hcap,
                   int rcap, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ int next_item;
  __shared__ uint64_t bar;
  griddep_wait();  // the plan (and, chained, whatever produced `vox`) is complete
  griddep_launch_dependents();
  vox = after_wait(vox); cells = after_wait(cells); slot0 = after_wait(slot0); meta = after_wait(meta);
  heavy = after_wait(heavy); recA = after_wait(recA); recB = after_wait(recB);
  if (RT > 0) r = RT;
  const int r2 = r * r, P = r2 * r;
  const int c = blockIdx.x, b = blockIdx.y;
  const float* plane_g = vox + ((long long)b * C + c) * P;
  float* o = outs + ((long long)b * C + c) * N;
  float* row = smem + (PLANE_SMEM ? ((P + 3) & ~3) : 0);
  char* rowb = (char*)row;
  if (threadIdx.x == 0) {
    next_item = 0;
    if (PLANE_SMEM) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
  }
  __syncthreads();
  if (PLANE_SMEM) stage_floats(smem, plane_g, P, &bar, 0, use_tma != 0);
  const float* plane = PLANE_SMEM ? smem : plane_g;
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const float4* ra = recA + (long long)b * rcap;
  const float4* rb = recB + (long long)b * rcap;
  const int l1 = meta[b * kMetaInts + 0], h1 = meta[b * kMetaInts + 1];
  const int lane = threadIdx.x & 31;
  const int dummy = 4 * N;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1);
  const int nslice = l1 >> 5, nitem = nslice + h1;
  const int4* hv = (const int4*)heavy + (long long)b * hcap;
  CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
  while (cur.item < nitem) {
    const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);  // header in flight during this item
    const bool light = cur.item < nslice;
    const int n = light ? (int)((unsigned)cur.h.x >> 24) : 1;
    const int cell = light ? (cur.h.x & 0xffffff) : cur.h.x;
    const unsigned mask = light ? (n ? (unsigned)cur.h.y >> 24 : 0u) : (unsigned)cur.h.w;
    // the cell's 8 corner values: a corner beyond the grid carries weight exactly 0 for every point of the cell
    // (trilinear_devox.cu:64-75 skips it by not offsetting the index) and is not read.  Interior cells (all of a warp's
    // cells, nearly always) take 8 plain loads at constant offsets from one base pointer.
    f32x2_t v[8];
    const float* pc = plane + cell;
    if (__all_sync(0xffffffffu, mask == 0xffu || n == 0)) {   // holes (n == 0) sit on cell 0: harmless reads
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x = PLANE_SMEM ? pc[off[k]] : __ldg(pc + off[k]);
        v[k] = pk2(x, x);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x = ((mask >> k) & 1u) ? (PLANE_SMEM ? pc[off[k]] : __ldg(pc + off[k])) : 0.f;
        v[k] = pk2(x, x);
      }
    }
    auto emit = [&](const TriPair t) {
      float lo, hi;
      upk2(tri_fwd_interpolate(t, v), lo, hi);
      *(float*)(rowb + t.off0) = lo;
      *(float*)(rowb + t.off1) = hi;
    };
    if (light) tri_walk(ra + (cur.h.y & 0xffffff), rb + (cur.h.y & 0xffffff), 0, (n + 1) >> 1, 2, 1, 32, dummy, emit);
    else tri_walk(ra + cur.h.y, rb + cur.h.y, lane, cur.h.z, 64, 32, 1, dummy, emit);
    cur = take_cell_item(nxt);
  }
  __syncthreads();
  // ---- write the row: one bulk (TMA) store for the 16-byte aligned part, plain stores for the rest ----
  const bool aligned = use_tma && ((((uintptr_t)o) & 15) == 0);
  const int bulk = aligned ? (N & ~3) : 0;
  if (bulk > 0) {
    fence_proxy_async();  // generic-proxy writes to the row -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t off2 = 0; off2 < (uint32_t)bulk * 4u; off2 += kBulkChunk)
        bulk_s2g((char*)o + off2, (const char*)row + off2, min(kBulkChunk, (uint32_t)bulk * 4u - off2));
      bulk_commit();
    }
  }
  for (int i = bulk + threadIdx.x; i < N; i += blockDim.x) o[i] = row[i];
  if (bulk > 0 && threadIdx.x == 0) bulk_wait_read();  // the CTA may exit once the source has been read
}

constexpr int kTriFwdPlaneMax = 8192;  // floats: planes up to r = 20 are staged in shared memory
static size_t tri_fwd_smem(int n, int r = 0) {
  const size_t p = (size_t)r * r * r;
  return ((size_t)n + 4 + (p <= kTriFwdPlaneMax ? ((p + 3) & ~(size_t)3) : 0)) * sizeof(float);
}

// inds / wgts of the training-mode forward (trilinear_devox.cu:64-94), one thread per point
__global__ void __launch_bounds__(256) tri_index_kernel(const float* __restrict__ coords, int* __restrict__ inds,
                                                        float* __restrict__ wgts, int n, int r) {
  griddep_wait();
  griddep_launch_dependents();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r2 = r * r;
  const float* cb = coords + (long long)b * 3 * n;
  const float hi = (float)(r - 1);
  const float x = fminf(fmaxf(cb[i], 0.f), hi);
  const float y = fminf(fmaxf(cb[(long long)i + n], 0.f), hi);
  const float z = fminf(fmaxf(cb[(long long)i + 2LL * n], 0.f), hi);
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  const float x1 = x - xl, y1 = y - yl, z1 = z - zl;
  const float x0 = 1.0f - x1, y0 = 1.0f - y1, z0 = 1.0f - z1;
  const int ox = (x1 > 0.f) ? r2 : 0, oy = (y1 > 0.f) ? r : 0, oz = (z1 > 0.f) ? 1 : 0;
  const int i0 = (int)xl * r2 + (int)yl * r + (int)zl;
  int* ib = inds + (long long)b * 8 * n;
  float* wb = wgts + (long long)b * 8 * n;
  const float xs[2] = {x0, x1}, ys[2] = {y0, y1}, zs[2] = {z0, z1};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;
    ib[(long long)k * n + i] = i0 + (dx ? ox : 0) + (dy ? oy : 0) + (dz ? oz : 0);
    wb[(long long)k * n + i] = __fmul_rn(__fmul_rn(xs[dx], ys[dy]), zs[dz]);
  }
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

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
  if (!cellplan_fits(n, (long long)r * r * r, true)) return PCFM_E_BADARG;  // 24-bit cell ids / record slots
  return devox_bwd_cells(grad_y, coords, b, c, n, r, grad_x, workspace, workspace_bytes, (cudaStream_t)stream);
}

int pcfm_trilinear_devoxelize_backward_planned(const float* grad_y, int b, int c, int n, int r, float* grad_x,
                                               const void* plan_workspace, size_t plan_workspace_bytes,
                                               pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!grad_x || !plan_workspace || (n > 0 && !grad_y)) return PCFM_E_BADARG;
  if (!cellplan_fits(n, (long long)r * r * r, true)) return PCFM_E_BADARG;
  return devox_bwd_cells(grad_y, nullptr, b, c, n, r, grad_x, const_cast<void*>(plan_workspace),
                         plan_workspace_bytes, (cudaStream_t)stream);
}

int pcfm_trilinear_devoxelize_forward_cells_ok(int b, int c, int n, int r) {
  if (b <= 0 || c <= 0 || n <= 0 || r <= 0) return 0;
  return cellplan_fits(n, (long long)r * r * r, true) && tri_fwd_smem(n, r) <= kTriSmemLimit ? 1 : 0;
}

int pcfm_trilinear_devoxelize_forward_cells(const float* coords, const float* features, int b, int c, int n, int r,
                                            int is_training, float* outs, int* inds, float* wgts, void* workspace,
                                            size_t workspace_bytes, int planned, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0) return 0;
  if (!workspace || (c > 0 && (!features || !outs))) return PCFM_E_BADARG;
  if ((!planned || is_training) && !coords) return PCFM_E_BADARG;
  if (is_training && (!inds || !wgts)) return PCFM_E_BADARG;
  if (!pcfm_trilinear_devoxelize_forward_cells_ok(b, c > 0 ? c : 1, n, r)) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int t = r * r * r;
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = planned ? cellplan_attach(true, b, n, t, &plan, &ws)
                   : cellplan_build(nullptr, coords, r, nullptr, b, n, t, nullptr, &plan, &ws, st);
  if (rc) return rc;
  if (c > 0) {
    const size_t smem = tri_fwd_smem(n, r);
    const bool ps = (long long)r * r * r <= kTriFwdPlaneMax;
#define PCFM_TFWD(RT, PS)                                                                                           \
  do {                                                                                                              \
    PCFM_CUDA(allow_dyn_smem((const void*)tri_fwd_kernel<RT, PS>, (int)kTriSmemLimit));                             \
    PCFM_CUDA(launch_dependent(tri_fwd_kernel<RT, PS>, dim3(c, b), dim3(512), smem, st, g_opt_pdl >= 2, features,   \
                               plan.cells, plan.slot0, plan.meta, plan.heavy, (const float4*)plan.rec,              \
                               (const float4*)plan.rec2, outs, c, n, r, plan.lcap, plan.hcap, plan.rcap,            \
                               (int)(g_opt_stage == 0)));                                                           \
  } while (0)
    if (r == 32) PCFM_TFWD(32, false);
    else if (r == 16) PCFM_TFWD(16, true);
    else if (r == 8) PCFM_TFWD(8, true);
    else if (ps) PCFM_TFWD(0, true);
    else PCFM_TFWD(0, false);
#undef PCFM_TFWD
    PCFM_LAUNCHED();
  }
  if (is_training) {
    PCFM_CUDA(launch_dependent(tri_index_kernel, dim3(cdiv(n, 256), b), dim3(256), 0, st, g_opt_pdl >= 2, coords, inds,
                               wgts, n, r));
    PCFM_LAUNCHED();
  }
  return 0;
}

}  // extern "C"
