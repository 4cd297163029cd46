// tri2.cu — the trilinear walks of tri.cu for TWO channels per CTA.
//
// What ncu showed for the one-channel walks (profiles/r02_*): they are bound by what happens PER POINT and is
// independent of the channel — the 16-byte plan record that comes through L1 (0.20 cycles of the L1 data pipe per
// point), the random 4-byte access to the point-indexed row in shared memory (~3.5 wavefronts per warp, 0.17-0.27
// cycles per point) and the L2 round trip every round of a walk waits for — while the arithmetic pipes idle.  With two
// channels per CTA every record, every weight and every round trip is used twice, and the two values of a point travel
// as ONE 8-byte shared-memory access: the two rows (and the two voxel planes) are INTERLEAVED in shared memory,
// [index][2 channels], and all arithmetic is packed fp32 with the two channels in the two lanes of FMUL2 / FFMA2.
//
// Two interleaved rows of N = 20000 floats take 160 KB, so one CTA owns an SM: the kernels are PERSISTENT (grid = SMs,
// 1024 threads) and loop over (batch, channel pair); row / plane transposes go through registers (coalesced 16-byte
// global accesses on one side, 16-byte shared accesses on the other), which ends when the stores are issued — unlike a
// bulk store out of shared memory, nothing has to drain before the next pair starts.
#include "cellplan.cuh"
#include "rows.cuh"
#include <stdio.h>

namespace pcfm {

constexpr size_t kTri2SmemLimit = 227 * 1024 - 512;
constexpr int kTri2Threads = 1024;

// both channels of point `off` (byte offset 4*p into a plain row => 8*p into an interleaved one)
__device__ __forceinline__ f32x2_t* pair_at(char* base, int off) { return (f32x2_t*)(base + 2 * (long long)off); }

// transposes: out2[i] = (a[i], b[i]) and back; count elements, all threads of the CTA
__device__ __forceinline__ void interleave_rows(float* __restrict__ dst2, const float* __restrict__ a,
                                                const float* __restrict__ b, int count) {
  const bool vec = ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
  const int n4 = vec ? count >> 2 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 x = __ldg((const float4*)a + i), y = __ldg((const float4*)b + i);
    float4* d = (float4*)dst2 + 2 * i;
    d[0] = make_float4(x.x, y.x, x.y, y.y);
    d[1] = make_float4(x.z, y.z, x.w, y.w);
  }
  for (int i = 4 * n4 + threadIdx.x; i < count; i += blockDim.x)
    ((float2*)dst2)[i] = make_float2(__ldg(a + i), __ldg(b + i));
}
__device__ __forceinline__ void deinterleave_rows(const float* __restrict__ src2, float* __restrict__ a,
                                                  float* __restrict__ b, int count) {
  const bool vec = ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
  const int n4 = vec ? count >> 2 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4* s = (const float4*)src2 + 2 * i;
    const float4 u = s[0], w = s[1];
    ((float4*)a)[i] = make_float4(u.x, u.z, w.x, w.z);
    ((float4*)b)[i] = make_float4(u.y, u.w, w.y, w.w);
  }
  for (int i = 4 * n4 + threadIdx.x; i < count; i += blockDim.x) {
    const float2 v = ((const float2*)src2)[i];
    a[i] = v.x;
    b[i] = v.y;
  }
}

// one point of a pair record: fractions as scalars
struct TriPoint {
  int off;
  float fx, fy, fz;
};

// forward: both channels of one point from the cell's 8 corner pairs (reduced along z, y, x)
__device__ __forceinline__ f32x2_t tri2_interpolate(const TriPoint& t, const f32x2_t* v) {
  const float x0 = 1.0f - t.fx, y0 = 1.0f - t.fy, z0 = 1.0f - t.fz;
  const f32x2_t zl = pk2(z0, z0), zh = pk2(t.fz, t.fz), yl = pk2(y0, y0), yh = pk2(t.fy, t.fy);
  const f32x2_t t00 = fma2(zh, v[1], mul2(zl, v[0]));
  const f32x2_t t01 = fma2(zh, v[3], mul2(zl, v[2]));
  const f32x2_t t10 = fma2(zh, v[5], mul2(zl, v[4]));
  const f32x2_t t11 = fma2(zh, v[7], mul2(zl, v[6]));
  const f32x2_t u0 = fma2(yh, t01, mul2(yl, t00));
  const f32x2_t u1 = fma2(yh, t11, mul2(yl, t10));
  return fma2(pk2(t.fx, t.fx), u1, mul2(pk2(x0, x0), u0));
}

// backward: s[dx,dy,dz] (both channels) += (x_dx * y_dy) * (z_dz * g)
__device__ __forceinline__ void tri2_accumulate(const TriPoint& t, f32x2_t g, f32x2_t* s) {
  const float x0 = 1.0f - t.fx, y0 = 1.0f - t.fy, z0 = 1.0f - t.fz;
  const f32x2_t gz0 = mul2(g, pk2(z0, z0)), gz1 = mul2(g, pk2(t.fz, t.fz));
  const float a00 = x0 * y0, a01 = x0 * t.fy, a10 = t.fx * y0, a11 = t.fx * t.fy;
  const f32x2_t p00 = pk2(a00, a00), p01 = pk2(a01, a01), p10 = pk2(a10, a10), p11 = pk2(a11, a11);
  s[0] = fma2(p00, gz0, s[0]);
  s[1] = fma2(p00, gz1, s[1]);
  s[2] = fma2(p01, gz0, s[2]);
  s[3] = fma2(p01, gz1, s[3]);
  s[4] = fma2(p10, gz0, s[4]);
  s[5] = fma2(p10, gz1, s[5]);
  s[6] = fma2(p11, gz0, s[6]);
  s[7] = fma2(p11, gz1, s[7]);
}

// Pair slots first, first + step, ... < end; slot j at pa[j * mul] / pb[j * mul]; two slots (four points) per round.
// A missing second slot is replaced by a dummy pair (offsets 4*N, zero fractions): it reads the exact zero / writes the
// scratch element behind the row.
template <typename Body>
__device__ __forceinline__ void tri2_walk(const float4* __restrict__ pa, const float4* __restrict__ pb, int first, int end,
                                          int step, int second, int mul, int dummy, Body&& body) {
  const float4 da = make_float4(__int_as_float(dummy), __int_as_float(dummy), 0.f, 0.f);
  const float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = first; j < end; j += step) {
    const bool two = j + second < end;
    const float4 a0 = __ldg(pa + j * mul), b0 = __ldg(pb + j * mul);
    const float4 a1 = two ? __ldg(pa + (j + second) * mul) : da, b1 = two ? __ldg(pb + (j + second) * mul) : db;
    body(TriPoint{__float_as_int(a0.x), a0.z, b0.x, b0.z});
    body(TriPoint{__float_as_int(a0.y), a0.w, b0.y, b0.w});
    body(TriPoint{__float_as_int(a1.x), a1.z, b1.x, b1.z});
    body(TriPoint{__float_as_int(a1.y), a1.w, b1.y, b1.w});
  }
}

// =================================================================================================
// forward: shared memory = [interleaved planes: 2*P floats (PLANE_SMEM)][interleaved rows: 2*(N+1) floats]
// =================================================================================================
template <int RT, bool PLANE_SMEM>
__global__ void __launch_bounds__(kTri2Threads, 1)
    tri2_fwd_kernel(const float* __restrict__ vox, const int* __restrict__ cells, const int* __restrict__ slot0,
                    const int* __restrict__ meta, const int* __restrict__ heavy, const float4* __restrict__ recA,
                    const float4* __restrict__ recB, float* __restrict__ outs, int B, int C, int N, int r, int lcap,
                    int hcap, int rcap, int dbg) {
  extern __shared__ __align__(128) float smem[];
  __shared__ int next_item;
  griddep_wait();  // the plan (and, chained, whatever produced `vox`) is complete
  griddep_launch_dependents();
  vox = after_wait(vox); cells = after_wait(cells); slot0 = after_wait(slot0); meta = after_wait(meta);
  heavy = after_wait(heavy); recA = after_wait(recA); recB = after_wait(recB);
  if (RT > 0) r = RT;
  const int r2 = r * r, P = r2 * r;
  float* plane2 = smem;                                   // [P][2] (PLANE_SMEM)
  float* row2 = smem + (PLANE_SMEM ? 2 * ((P + 3) & ~3) : 0);   // [N + 1][2]
  char* rowb = (char*)row2;
  const int lane = threadIdx.x & 31;
  const int dummy = 4 * N;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1);
  const int half_c = C >> 1, npairs = B * half_c;
  for (int job = blockIdx.x; job < npairs; job += gridDim.x) {
    const int b = job / half_c, c0 = 2 * (job - b * half_c);
    const float* pg0 = vox + ((long long)b * C + c0) * P;
    const float* pg1 = pg0 + P;
    if (threadIdx.x == 0) next_item = 0;
    if (PLANE_SMEM) interleave_rows(plane2, pg0, pg1, P);
    __syncthreads();   // planes staged; the previous pair's rows have been read out (barrier at the end of the loop)
    const int* cl = cells + (long long)b * lcap;
    const int* sl = slot0 + (long long)b * lcap;
    const float4* ra = recA + (long long)b * rcap;
    const float4* rb = recB + (long long)b * rcap;
    const int l1 = meta[b * kMetaInts + 0], h1 = meta[b * kMetaInts + 1];
    const int nslice = l1 >> 5, nitem = nslice + h1;
    const int4* hv = (const int4*)heavy + (long long)b * hcap;
    CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
    while (cur.item < nitem) {
      const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);  // header in flight during this item
      const bool light = cur.item < nslice;
      const int n = light ? (int)((unsigned)cur.h.x >> 24) : 1;
      const int cell = light ? (cur.h.x & 0xffffff) : cur.h.x;
      const unsigned mask = light ? (n ? (unsigned)cur.h.y >> 24 : 0u) : (unsigned)cur.h.w;
      // the 8 corner values of both channels; a corner beyond the grid carries weight 0 and is not read
      f32x2_t v[8];
      const bool full = __all_sync(0xffffffffu, mask == 0xffu || n == 0);   // holes sit on cell 0: harmless reads
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = full || ((mask >> k) & 1u);
        if (PLANE_SMEM) {
          v[k] = ok ? *(const f32x2_t*)(plane2 + 2 * (cell + off[k])) : 0ull;
        } else {
          const float x0 = ok ? __ldg(pg0 + cell + off[k]) : 0.f, x1 = ok ? __ldg(pg1 + cell + off[k]) : 0.f;
          v[k] = pk2(x0, x1);
        }
      }
      auto emit = [&](const TriPoint t) {
        const f32x2_t o = tri2_interpolate(t, v);
        if (!(dbg & 4) || o == 0x1234567812345678ull) *pair_at(rowb, t.off) = o;
      };
      if ((light && (dbg & 1)) || (!light && (dbg & 2))) {
      } else if (light) tri2_walk(ra + (cur.h.y & 0xffffff), rb + (cur.h.y & 0xffffff), 0, (n + 1) >> 1, 2, 1, 32, dummy, emit);
      else tri2_walk(ra + cur.h.y, rb + cur.h.y, lane, cur.h.z, 64, 32, 1, dummy, emit);
      cur = take_cell_item(nxt);
    }
    __syncthreads();
    float* o0 = outs + ((long long)b * C + c0) * N;
    deinterleave_rows(row2, o0, o0 + N, N);
    __syncthreads();   // rows read out before the next pair writes them
  }
}

// =================================================================================================
// backward (planes that fit in shared memory twice): shared memory = [interleaved planes: 2*P][interleaved rows: 2*(N+1)]
// =================================================================================================
template <int RT>
__global__ void __launch_bounds__(kTri2Threads, 1)
    tri2_bwd_kernel(const float* __restrict__ grad_y, const int* __restrict__ cells, const int* __restrict__ slot0,
                    const int* __restrict__ meta, const int* __restrict__ heavy, const float4* __restrict__ recA,
                    const float4* __restrict__ recB, float* __restrict__ grad_x, int B, int C, int N, int r, int lcap,
                    int hcap, int rcap, int wait_first) {
  extern __shared__ __align__(128) float smem[];
  __shared__ int next_item;
  if (RT > 0) r = RT;
  const int r2 = r * r, P = r2 * r;
  float* plane2 = smem;                          // [P][2]
  float* row2 = smem + 2 * ((P + 3) & ~3);       // [N + 1][2]; element N stays zero (dummy half of an odd cell)
  const char* rowb = (const char*)row2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int dummy = 4 * N;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = (k >> 2) * r2 + ((k >> 1) & 1) * r + (k & 1);
  const int half_c = C >> 1, npairs = B * half_c;
  // reused plan: the predecessor is an arbitrary kernel (may produce grad_y, may still read grad_x) -> wait first.
  // fresh plan: the predecessor is the plan kernel; the first pair's rows are staged while it still runs (PDL).
  if (wait_first) griddep_wait();
  bool staged = false;
  if (!wait_first && (int)blockIdx.x < npairs) {
    const int b = blockIdx.x / half_c, c0 = 2 * (blockIdx.x - b * half_c);
    const float* g0 = grad_y + ((long long)b * C + c0) * N;
    interleave_rows(row2, g0, g0 + N, N);
    staged = true;
    griddep_wait();
  }
  griddep_launch_dependents();
  cells = after_wait(cells); slot0 = after_wait(slot0); meta = after_wait(meta); heavy = after_wait(heavy);
  recA = after_wait(recA); recB = after_wait(recB);
  for (int job = blockIdx.x; job < npairs; job += gridDim.x) {
    const int b = job / half_c, c0 = 2 * (job - b * half_c);
    const float* g0 = grad_y + ((long long)b * C + c0) * N;
    if (!staged) interleave_rows(row2, g0, g0 + N, N);
    staged = false;
    if (threadIdx.x == 0) {
      row2[2 * N] = 0.f;
      row2[2 * N + 1] = 0.f;
    }
    for (int i = threadIdx.x; i < ((2 * P + 3) >> 2); i += blockDim.x) ((float4*)plane2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int* cl = cells + (long long)b * lcap;
    const int* sl = slot0 + (long long)b * lcap;
    const float4* ra = recA + (long long)b * rcap;
    const float4* rb = recB + (long long)b * rcap;
    const int l1 = meta[b * kMetaInts + 0], h1 = meta[b * kMetaInts + 1];
    const int nslice = l1 >> 5, nitem = nslice + h1;
    const int4* hv = (const int4*)heavy + (long long)b * hcap;
    // ---- light cells: slices dealt to the warps in rounds; after every round the 8 corners are added in 8 phases
    // (inside a phase no two cells touch the same voxel: plain 8-byte read-modify-write, see tri.cu) ----
    {
      int hx = 0, hy = 0;
      if (warp < nslice) {
        hx = __ldg(cl + ((warp << 5) | lane));
        hy = __ldg(sl + ((warp << 5) | lane));
      }
      for (int rd = 0; rd * nw < nslice; ++rd) {
        const int n = (int)((unsigned)hx >> 24);
        const int cell = hx & 0xffffff;
        const unsigned se = (unsigned)hy;
        const unsigned mask = n ? se >> 24 : 0u;
        const int nx_item = (rd + 1) * nw + warp;
        hx = 0;
        if (nx_item < nslice) {
          hx = __ldg(cl + ((nx_item << 5) | lane));
          hy = __ldg(sl + ((nx_item << 5) | lane));
        }
        f32x2_t s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = 0ull;
        tri2_walk(ra + (se & 0xffffff), rb + (se & 0xffffff), 0, (n + 1) >> 1, 2, 1, 32, dummy,
                  [&](const TriPoint t) {
#ifdef PCFM_TRI2_CHECK
                    if ((unsigned)t.off > (unsigned)dummy || (t.off & 3)) {
                      printf("light bad off %d: job %d b %d rd %d warp %d lane %d n %d cell %d s0 %d nslice %d h1 %d\n", t.off, job, b, rd,
                             warp, lane, n, cell, (int)(se & 0xffffff), nslice, h1);
                      return;
                    }
#endif
                    tri2_accumulate(t, *(const f32x2_t*)(rowb + 2 * (long long)t.off), s); });
        f32x2_t* pv = (f32x2_t*)plane2 + cell;
        const f32x2_t one = pk2(1.0f, 1.0f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if ((mask >> k) & 1u) pv[off[k]] = fma2(s[k], one, pv[off[k]]);
          __syncthreads();
        }
        asm volatile("" : "+r"(hx), "+r"(hy));
      }
    }
    // ---- heavy pieces: whole warps, dynamic; sums reduced over the warp, added with atomics ----
    if (threadIdx.x == 0) next_item = nslice;
    __syncthreads();
    CellItem cur = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
    while (cur.item < nitem) {
      const CellItem nxt = next_cell_item(&next_item, nslice, nitem, cl, sl, hv);
      const int4 hd = cur.h;
      f32x2_t s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = 0ull;
      tri2_walk(ra + hd.y, rb + hd.y, lane, hd.z, 64, 32, 1, dummy,
                [&](const TriPoint t) {
#ifdef PCFM_TRI2_CHECK
                  if ((unsigned)t.off > (unsigned)dummy || (t.off & 3)) {
                    printf("heavy bad off %d: job %d b %d item %d lane %d hd %d %d %d %d nslice %d h1 %d\n", t.off, job, b, cur.item,
                           lane, hd.x, hd.y, hd.z, hd.w, nslice, h1);
                    return;
                  }
#endif
                  tri2_accumulate(t, *(const f32x2_t*)(rowb + 2 * (long long)t.off), s); });
      // 16 partial sums per lane (8 corners x 2 channels), recursive halving: lane l ends with value index (l >> 1) & 15
      float v16[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) upk2(s[k], v16[2 * k], v16[2 * k + 1]);
      float a8[8], a4[4], a2[2], a1;
      const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4, up2 = lane & 2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float send = up16 ? v16[k] : v16[k + 8];
        const float keep = up16 ? v16[k + 8] : v16[k];
        a8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = up8 ? a8[k] : a8[k + 4];
        const float keep = up8 ? a8[k + 4] : a8[k];
        a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = up4 ? a4[k] : a4[k + 2];
        const float keep = up4 ? a4[k + 2] : a4[k];
        a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      {
        const float send = up2 ? a2[0] : a2[1];
        const float keep = up2 ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
      if ((lane & 1) == 0) {
        const int idx = lane >> 1;           // bit 4 -> +8, bit 3 -> +4, bit 2 -> +2, bit 1 -> +1: value index 0..15
        const int k = idx >> 1, ch = idx & 1;  // v16[2k + ch]
        if ((hd.w >> k) & 1) atomicAdd(plane2 + 2 * (hd.x + off[k]) + ch, a1);
      }
      cur = take_cell_item(nxt);
    }
    __syncthreads();
    float* o0 = grad_x + ((long long)b * C + c0) * P;
    deinterleave_rows(plane2, o0, o0 + P, P);
    __syncthreads();
  }
}

size_t tri2_fwd_smem(int n, int r) {
  const size_t p = (size_t)r * r * r;
  const size_t plane = 2 * p * sizeof(float) <= 64 * 1024 ? 2 * ((p + 3) & ~(size_t)3) : 0;
  return (plane + 2 * ((size_t)n + 2)) * sizeof(float);
}
size_t tri2_bwd_smem(int n, int r) {
  const size_t p = (size_t)r * r * r;
  return (2 * ((p + 3) & ~(size_t)3) + 2 * ((size_t)n + 2)) * sizeof(float);
}
bool tri2_fwd_ok(int c, int n, int r) { return (c & 1) == 0 && c >= 2 && tri2_fwd_smem(n, r) <= kTri2SmemLimit; }
bool tri2_bwd_ok(int c, int n, int r) { return (c & 1) == 0 && c >= 2 && tri2_bwd_smem(n, r) <= kTri2SmemLimit; }

int launch_tri2_fwd(const float* vox, const CellPlan& plan, float* outs, int b, int c, int n, int r, cudaStream_t st) {
  const size_t smem = tri2_fwd_smem(n, r);
  const bool ps = 2 * (size_t)r * r * r * sizeof(float) <= 64 * 1024;
  const int grid = min(b * (c / 2), num_sms());
#define PCFM_T2F(RT, PS)                                                                                              \
  do {                                                                                                                \
    PCFM_CUDA(allow_dyn_smem((const void*)tri2_fwd_kernel<RT, PS>, (int)kTri2SmemLimit));                             \
    PCFM_CUDA(launch_dependent(tri2_fwd_kernel<RT, PS>, dim3(grid), dim3(kTri2Threads), smem, st, g_opt_pdl >= 2, vox, \
                               plan.cells, plan.slot0, plan.meta, plan.heavy, (const float4*)plan.rec,                \
                               (const float4*)plan.rec2, outs, b, c, n, r, plan.lcap, plan.hcap, plan.rcap,           \
                               g_opt_dbg));                                                                           \
  } while (0)
  if (r == 32) PCFM_T2F(32, false);
  else if (r == 16) PCFM_T2F(16, true);
  else if (r == 8) PCFM_T2F(8, true);
  else if (ps) PCFM_T2F(0, true);
  else PCFM_T2F(0, false);
#undef PCFM_T2F
  PCFM_LAUNCHED();
  return 0;
}

int launch_tri2_bwd(const float* grad_y, const CellPlan& plan, float* grad_x, int b, int c, int n, int r, bool reused,
                    cudaStream_t st) {
  const size_t smem = tri2_bwd_smem(n, r);
  const int grid = min(b * (c / 2), num_sms());
  const bool dep = reused ? g_opt_pdl >= 2 : plan.pdl;
#define PCFM_T2B(RT)                                                                                                  \
  do {                                                                                                                \
    PCFM_CUDA(allow_dyn_smem((const void*)tri2_bwd_kernel<RT>, (int)kTri2SmemLimit));                                 \
    PCFM_CUDA(launch_dependent(tri2_bwd_kernel<RT>, dim3(grid), dim3(kTri2Threads), smem, st, dep, grad_y, plan.cells, \
                               plan.slot0, plan.meta, plan.heavy, (const float4*)plan.rec, (const float4*)plan.rec2,  \
                               grad_x, b, c, n, r, plan.lcap, plan.hcap, plan.rcap, (int)reused));                    \
  } while (0)
  if (r == 16) PCFM_T2B(16);
  else if (r == 8) PCFM_T2B(8);
  else PCFM_T2B(0);
#undef PCFM_T2B
  PCFM_LAUNCHED();
  return 0;
}

}  // namespace pcfm
