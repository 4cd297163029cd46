// devox.cu — trilinear devoxelisation forward (trilinear_devox.cu:21-105 semantics).
//
// One CTA owns G consecutive channel planes of one batch element: the planes (G * r^3 floats) are staged in
// shared memory with 1-D TMA bulk copies, then every point gathers its 8 corners from shared memory and the
// result row is written coalesced.  Indices / weights are computed once per point and reused for G channels.
#include "rows.cuh"

namespace pcfm {

// Shared-memory plane layout: x-layer `x` of plane `g` starts at ((g*(r+1) + x) * sx) floats, sx = r*r + 4.
//   * every layer is one 16-byte-aligned TMA bulk copy (r*r floats),
//   * the +4 skew moves consecutive x-layers by four banks (the unpadded layout maps a voxel to bank z only, and a
//     centred cloud then fights over a handful of banks: 6-way conflicts measured on Gaussian clouds),
//   * layer r of every plane is a zeroed guard, so the 8 corners can always be read at the CONSTANT offsets
//     {0,1} + {0,r} + {0,sx}: a corner beyond the grid (or beyond a lattice plane) carries weight exactly 0
//     (trilinear_devox.cu:64-75 skips it by not offsetting the index), and 0 * finite = 0.
//   * IL (G % 4 == 0): the planes of 4 consecutive channels are interleaved, voxel-major ([voxel][4] floats), and a
//     corner is ONE 16-byte shared load for 4 channels: 4x fewer LSU instructions (the G = 4 kernels sat in
//     mio_throttle, 7-13 stalled warps per issue) and fewer wavefronts per gathered float (8 lanes per phase
//     instead of 32 fight for the banks).  Staged with plain loads (TMA cannot interleave).
template <int G, int RT, bool PLANE_SMEM, bool IL>
__global__ void __launch_bounds__(1024, 1)
    devox_fwd_kernel(const float* __restrict__ coords, const float* __restrict__ feat,
                     float* __restrict__ outs, int* __restrict__ inds, float* __restrict__ wgts, int C, int n,
                     int r_rt, int training, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  griddep_wait();  // chained PDL: this kernel may have been scheduled before its predecessor on the stream finished
  griddep_launch_dependents();
  const int r = RT > 0 ? RT : r_rt;
  const int c0 = blockIdx.x * G, b = blockIdx.y;
  const int r2 = r * r, P = r2 * r;
  const int sx = PLANE_SMEM ? r2 + 4 : r2;   // layer stride
  // plane stride: r layers + guard; the farthest corner read is r*sx + r*r + r floats from the plane base
  const int sp = PLANE_SMEM ? ((r + 1) * sx + r + 3) & ~3 : P;
  const float* pg = feat + ((long long)b * C + c0) * P;
  bool bulk_copy = false;
  if (PLANE_SMEM) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    const bool bulk = !IL && use_tma && ((r2 & 3) == 0) && ((((uintptr_t)pg) & 15) == 0);
    bulk_copy = bulk;
    if (IL) {
      for (int i = threadIdx.x; i < G * sp; i += blockDim.x) smem[i] = 0.f;  // skew + guard voxels
      __syncthreads();
      for (int i = threadIdx.x; i < G * P; i += blockDim.x) {  // coalesced read, 4-way strided shared store
        const int g = i / P, rem = i - g * P, x = rem / r2;
        smem[(((size_t)(g >> 2) * sp + (size_t)x * sx + (rem - x * r2)) << 2) | (g & 3)] = __ldg(pg + i);
      }
    } else if (bulk) {
      if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, (uint32_t)(G * P) * 4u);
      __syncthreads();  // expect_tx is armed before any copy can complete
      for (int l = threadIdx.x; l < G * r; l += blockDim.x) {   // one bulk copy per x-layer, issued by G*r threads
        const int g = l / r, x = l - g * r;
        bulk_g2s(smem + (size_t)g * sp + (size_t)x * sx, pg + (size_t)g * P + (size_t)x * r2, (uint32_t)r2 * 4u, &bar);
      }
    } else {
      for (int i = threadIdx.x; i < G * P; i += blockDim.x) {
        const int g = i / P, rem = i - g * P, x = rem / r2;
        smem[(size_t)g * sp + (size_t)x * sx + (rem - x * r2)] = __ldg(pg + i);
      }
    }
    // zero the 4-float skew of every layer and the guard layer of every plane (never written by the copies)
    for (int i = threadIdx.x; !IL && i < G * r * 4; i += blockDim.x) {
      const int l = i >> 2, g = l / r, x = l - g * r;
      smem[(size_t)g * sp + (size_t)x * sx + r2 + (i & 3)] = 0.f;
    }
    const int guard = sp - r * sx;
    for (int i = threadIdx.x; !IL && i < G * guard; i += blockDim.x) {
      const int g = i / guard;
      smem[(size_t)g * sp + (size_t)r * sx + (i - g * guard)] = 0.f;
    }
  }
  const float* cb = coords + (long long)b * 3 * n;
  float* ob = outs + ((long long)b * C + c0) * n;
  const bool save = training && blockIdx.x == 0;
  int* ib = inds ? inds + (long long)b * 8 * n : nullptr;
  float* wb = wgts ? wgts + (long long)b * 8 * n : nullptr;
  const float hi = (float)(r - 1);
  // software pipeline: the coordinates of the next point are requested before the current one is gathered, and those
  // of the first point while the planes are still landing
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if ((int)threadIdx.x < n) {
    nx = cb[threadIdx.x];
    ny = cb[(long long)threadIdx.x + n];
    nz = cb[(long long)threadIdx.x + 2LL * n];
  }
  if (PLANE_SMEM) {
    if (bulk_copy) mbar_wait(&bar, 0);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // coords are contractually in [0, r-1]; the clamp is a no-op there and keeps every index in range otherwise
    const float x = fminf(fmaxf(nx, 0.f), hi);
    const float y = fminf(fmaxf(ny, 0.f), hi);
    const float z = fminf(fmaxf(nz, 0.f), hi);
    {
      const int inext = i + blockDim.x;
      if (inext < n) {
        nx = cb[inext];
        ny = cb[(long long)inext + n];
        nz = cb[(long long)inext + 2LL * n];
      }
    }
    const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
    const float x1 = x - xl, y1 = y - yl, z1 = z - zl;
    const float x0 = 1.0f - x1, y0 = 1.0f - y1, z0 = 1.0f - z1;
    float w[8];
    w[0] = __fmul_rn(__fmul_rn(x0, y0), z0);
    w[1] = __fmul_rn(__fmul_rn(x0, y0), z1);
    w[2] = __fmul_rn(__fmul_rn(x0, y1), z0);
    w[3] = __fmul_rn(__fmul_rn(x0, y1), z1);
    w[4] = __fmul_rn(__fmul_rn(x1, y0), z0);
    w[5] = __fmul_rn(__fmul_rn(x1, y0), z1);
    w[6] = __fmul_rn(__fmul_rn(x1, y1), z0);
    w[7] = __fmul_rn(__fmul_rn(x1, y1), z1);
    const int xi = (int)xl, yi = (int)yl, zi = (int)zl;
    if (save) {  // the reference's index rule, written once per batch element (trilinear_devox.cu:64-94)
      const int ox = (x1 > 0.f) ? r2 : 0, oy = (y1 > 0.f) ? r : 0, oz = (z1 > 0.f) ? 1 : 0;
      const int i0 = xi * r2 + yi * r + zi;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ib[(long long)k * n + i] = i0 + ((k & 4) ? ox : 0) + ((k & 2) ? oy : 0) + ((k & 1) ? oz : 0);
        wb[(long long)k * n + i] = w[k];
      }
    }
    if (PLANE_SMEM && IL) {
      const float4* p0 = (const float4*)smem + (xi * sx + yi * r + zi);
#pragma unroll
      for (int q = 0; q < G / 4; ++q) {
        const float4* pl = p0 + q * sp;
        const int co[8] = {0, 1, r, r + 1, sx, sx + 1, sx + r, sx + r + 1};
        float4 v = pl[0];
        float a0 = w[0] * v.x, a1 = w[0] * v.y, a2 = w[0] * v.z, a3 = w[0] * v.w;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
          v = pl[co[k]];
          a0 = fmaf(w[k], v.x, a0);
          a1 = fmaf(w[k], v.y, a1);
          a2 = fmaf(w[k], v.z, a2);
          a3 = fmaf(w[k], v.w, a3);
        }
        float* oq = ob + (long long)(4 * q) * n + i;
        oq[0] = a0;
        oq[n] = a1;
        oq[2LL * n] = a2;
        oq[3LL * n] = a3;
      }
    } else if (PLANE_SMEM) {
      const float* p0 = smem + (xi * sx + yi * r + zi);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float* pl = p0 + g * sp;
        float acc = w[0] * pl[0];
        acc = fmaf(w[1], pl[1], acc);
        acc = fmaf(w[2], pl[r], acc);
        acc = fmaf(w[3], pl[r + 1], acc);
        acc = fmaf(w[4], pl[sx], acc);
        acc = fmaf(w[5], pl[sx + 1], acc);
        acc = fmaf(w[6], pl[sx + r], acc);
        acc = fmaf(w[7], pl[sx + r + 1], acc);
        ob[(long long)g * n + i] = acc;
      }
    } else {
      // plane does not fit in shared memory: read through L2 with the reference's in-range indices
      const int ox = (x1 > 0.f) ? r2 : 0, oy = (y1 > 0.f) ? r : 0, oz = (z1 > 0.f) ? 1 : 0;
      const int i0 = xi * r2 + yi * r + zi;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float* pl = pg + (long long)g * P + i0;
        float acc = w[0] * __ldg(pl);
        acc = fmaf(w[1], __ldg(pl + oz), acc);
        acc = fmaf(w[2], __ldg(pl + oy), acc);
        acc = fmaf(w[3], __ldg(pl + oy + oz), acc);
        acc = fmaf(w[4], __ldg(pl + ox), acc);
        acc = fmaf(w[5], __ldg(pl + ox + oz), acc);
        acc = fmaf(w[6], __ldg(pl + ox + oy), acc);
        acc = fmaf(w[7], __ldg(pl + ox + oy + oz), acc);
        ob[(long long)g * n + i] = acc;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Channel-lane variant for small grids (r <= 10, C % 32 == 0) — pcfm_set_option("devox_cl", 1 | 2), OFF by default.
//
// The kernel above maps lanes to POINTS: the 32 lanes of a gather hit 32 unrelated voxels, and a dense cloud on a
// small grid piles them onto a few banks (ncu, r = 8: 3.1 wavefronts per LDS, 68 % of the LSU pipe, the bound of the
// kernel).  Here lanes map to CHANNELS: the block of 32 channels is staged voxel-major, [voxel][32 ch] with a row of
// 36 floats, and a quarter-warp (8 lanes x 4 channels) reads the 128 contiguous bytes of ONE voxel with one LDS.128:
// conflict-free by construction, 4 wavefronts per warp load = the minimum for 512 bytes.  The cube is padded to
// (r+1)^3 with zero guard voxels so the 8 corners sit at constant offsets, as above.  Grid = (C/32, B, point slices).
// Measured (B=8, C=256, N=20000, r=8): shared-memory wavefronts fall from 68 % to 38 % of the pipe as designed, but
// the kernel is not faster: 0.116 ms vs 0.108 ms for the point-lane kernel.  Every lane now repeats the weight
// arithmetic of its point (amortised over 4 channels instead of 8), the output leaves as 8 rows x 64 bytes per store
// instruction instead of one 128-byte line (L1TEX 92 % busy with scalar stores, still the top unit with float4 ones),
// and the plane block is re-staged per point slice.  Kept as a tested option; the point-lane kernels stay the default.
constexpr int kClRow = 36;  // floats per voxel row: 32 channels + 4 (16-byte aligned, consecutive voxels 4 banks apart)
constexpr int kClThreads = 384;
// A warp step covers 16 consecutive points: quarter-warp q takes points 4q .. 4q+3 one after the other, so that a lane
// ends up with 4 channels x 4 CONSECUTIVE points and writes each channel with one 16-byte store (the four quarters of
// a warp fill 64 contiguous bytes of 8 channel rows per store instruction; one point per lane and step meant 8 rows x
// 16 bytes per scalar store and the L1TEX pipe at 92 %).  The 4 points' coordinates arrive as one float4 per axis.
__global__ void __launch_bounds__(kClThreads, 2)
    devox_fwd_cl_kernel(const float* __restrict__ coords, const float* __restrict__ feat, float* __restrict__ outs,
                        int* __restrict__ inds, float* __restrict__ wgts, int C, int n, int r, int training,
                        int pts_per_cta) {
  extern __shared__ __align__(128) float smem[];
  griddep_wait();
  griddep_launch_dependents();
  const int c0 = blockIdx.x * 32, b = blockIdx.y;
  const int r1 = r + 1, r2 = r * r, P = r2 * r, P1 = r1 * r1 * r1;
  const float* pg = feat + ((long long)b * C + c0) * P;
  {
    float4* z4 = (float4*)smem;
    for (int i = threadIdx.x; i < P1 * (kClRow / 4); i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  for (int v = threadIdx.x; v < P; v += blockDim.x) {  // a thread copies the 32 channels of its voxel(s): loads coalesced
    const int x = v / r2, y = (v - x * r2) / r, z = v - x * r2 - y * r;  // along v, 8 in flight; 4-way conflict per store
    float* dst = smem + ((x * r1 + y) * r1 + z) * kClRow;
    const float* src = pg + v;
#pragma unroll
    for (int c8 = 0; c8 < 32; c8 += 8) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = __ldg(src + (size_t)(c8 + k) * P);
      *(float4*)(dst + c8) = make_float4(t[0], t[1], t[2], t[3]);
      *(float4*)(dst + c8 + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, q = lane >> 3, j = lane & 7;  // points 4q..4q+3 of the warp step, channel quad j
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* cb = coords + (long long)b * 3 * n;
  float* ob = outs + ((long long)b * C + c0 + 4 * j) * n;
  const bool save = training && blockIdx.x == 0;
  int* ib = inds ? inds + (long long)b * 8 * n : nullptr;
  float* wb = wgts ? wgts + (long long)b * 8 * n : nullptr;
  const float hi = (float)(r - 1);
  const int p_lo = blockIdx.z * pts_per_cta, p_hi = min(n, p_lo + pts_per_cta);  // p_lo is a multiple of 16
  const bool vec = ((n & 3) == 0) && ((((uintptr_t)cb) & 15) == 0) && ((((uintptr_t)ob) & 15) == 0);
  const int co[8] = {0, 1, r1, r1 + 1, r1 * r1, r1 * r1 + 1, r1 * r1 + r1, r1 * r1 + r1 + 1};
  auto load4 = [&](int p, float (&cx)[4], float (&cy)[4], float (&cz)[4]) {  // coordinates of points p .. p+3 (< p_hi)
    if (vec && p + 3 < p_hi) {
      const float4 a = *(const float4*)(cb + p), bb = *(const float4*)(cb + (long long)p + n),
                   c = *(const float4*)(cb + (long long)p + 2LL * n);
      cx[0] = a.x; cx[1] = a.y; cx[2] = a.z; cx[3] = a.w;
      cy[0] = bb.x; cy[1] = bb.y; cy[2] = bb.z; cy[3] = bb.w;
      cz[0] = c.x; cz[1] = c.y; cz[2] = c.z; cz[3] = c.w;
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const bool in = p + s < p_hi;
        cx[s] = in ? cb[p + s] : 0.f;
        cy[s] = in ? cb[(long long)p + s + n] : 0.f;
        cz[s] = in ? cb[(long long)p + s + 2LL * n] : 0.f;
      }
    }
  };
  float nx[4], ny[4], nz[4];
  int p = p_lo + warp * 16 + 4 * q;
  if (p < p_hi) load4(p, nx, ny, nz);
  for (; p < p_hi; p += nw * 16) {
    float cx[4], cy[4], cz[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { cx[s] = nx[s]; cy[s] = ny[s]; cz[s] = nz[s]; }
    if (p + nw * 16 < p_hi) load4(p + nw * 16, nx, ny, nz);  // next step's coordinates in flight during this one
    float acc[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float x = fminf(fmaxf(cx[s], 0.f), hi);
      const float y = fminf(fmaxf(cy[s], 0.f), hi);
      const float z = fminf(fmaxf(cz[s], 0.f), hi);
      const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
      const float x1 = x - xl, y1 = y - yl, z1 = z - zl;
      const float x0 = 1.0f - x1, y0 = 1.0f - y1, z0 = 1.0f - z1;
      float w[8];
      w[0] = __fmul_rn(__fmul_rn(x0, y0), z0);
      w[1] = __fmul_rn(__fmul_rn(x0, y0), z1);
      w[2] = __fmul_rn(__fmul_rn(x0, y1), z0);
      w[3] = __fmul_rn(__fmul_rn(x0, y1), z1);
      w[4] = __fmul_rn(__fmul_rn(x1, y0), z0);
      w[5] = __fmul_rn(__fmul_rn(x1, y0), z1);
      w[6] = __fmul_rn(__fmul_rn(x1, y1), z0);
      w[7] = __fmul_rn(__fmul_rn(x1, y1), z1);
      const int xi = (int)xl, yi = (int)yl, zi = (int)zl;
      if (save && p + s < p_hi) {  // lane j of the quarter-warp writes corner j (trilinear_devox.cu:64-94)
        const int ox = (x1 > 0.f) ? r2 : 0, oy = (y1 > 0.f) ? r : 0, oz = (z1 > 0.f) ? 1 : 0;
        const int i0 = xi * r2 + yi * r + zi;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j == k) {
            ib[(long long)k * n + p + s] = i0 + ((k & 4) ? ox : 0) + ((k & 2) ? oy : 0) + ((k & 1) ? oz : 0);
            wb[(long long)k * n + p + s] = w[k];
          }
      }
      const float4* pv = (const float4*)(smem + ((xi * r1 + yi) * r1 + zi) * kClRow) + j;
      float4 v = pv[0];
      float a0 = w[0] * v.x, a1 = w[0] * v.y, a2 = w[0] * v.z, a3 = w[0] * v.w;
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        v = pv[co[k] * (kClRow / 4)];
        a0 = fmaf(w[k], v.x, a0);
        a1 = fmaf(w[k], v.y, a1);
        a2 = fmaf(w[k], v.z, a2);
        a3 = fmaf(w[k], v.w, a3);
      }
      acc[s][0] = a0; acc[s][1] = a1; acc[s][2] = a2; acc[s][3] = a3;
    }
    float* o = ob + p;
    if (vec && p + 3 < p_hi) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *(float4*)(o + (long long)i * n) = make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s)
        if (p + s < p_hi) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[(long long)i * n + s] = acc[s][i];
        }
    }
  }
}

static inline bool devox_cl_ok(int n) { return n >= 4; }
static size_t devox_cl_smem(int r) { return (size_t)(r + 1) * (r + 1) * (r + 1) * kClRow * sizeof(float); }

static int launch_devox_cl(const float* coords, const float* feat, float* outs, int* inds, float* wgts, int b, int c,
                           int n, int r, int training, cudaStream_t st) {
  const size_t smem = devox_cl_smem(r);
  PCFM_CUDA(allow_dyn_smem((const void*)devox_fwd_cl_kernel, (int)(227 * 1024 - 512)));
  // point slices: ~4 CTAs per SM over the whole grid, at least 1024 points each, a multiple of 16 points
  int slices = cdiv(4LL * num_sms(), (long long)(c / 32) * b);
  slices = max(1, min(slices, cdiv(n, 1024)));
  const int per = ((cdiv(n, slices) + 15) / 16) * 16;
  dim3 grid(c / 32, b, cdiv(n, per));
  PCFM_CUDA(launch_dependent(devox_fwd_cl_kernel, grid, dim3(kClThreads), smem, st, g_opt_pdl >= 2, coords, feat, outs, inds,
                             wgts, c, n, r, training, per));
  PCFM_LAUNCHED();
  return 0;
}

static size_t devox_fwd_smem(int g, int r) {
  const size_t sx = (size_t)r * r + 4;
  return (size_t)g * (((r + 1) * sx + r + 3) & ~(size_t)3) * sizeof(float);
}
constexpr size_t kDevoxSmemLimit = 227 * 1024 - 512;

template <int G, int RT, bool PLANE_SMEM, bool IL = false>
static int launch_devox_t(const float* coords, const float* feat, float* outs, int* inds, float* wgts, int b,
                          int c, int n, int r, int training, cudaStream_t st) {
  auto kern = devox_fwd_kernel<G, RT, PLANE_SMEM, IL>;
  const size_t smem = PLANE_SMEM ? devox_fwd_smem(G, r) : 0;
  PCFM_CUDA(allow_dyn_smem((const void*)kern, (int)kDevoxSmemLimit));
  dim3 grid(c / G, b);
  // two CTAs per SM when the planes are small (56+ registers: 512 threads each), else one CTA of 1024 threads
  const int threads = (PLANE_SMEM && smem <= 110 * 1024) ? 512 : 1024;
  PCFM_CUDA(launch_dependent(kern, grid, dim3(threads), smem, st, g_opt_pdl >= 2, coords, feat, outs, inds, wgts, c, n,
                             r, training, (int)(g_opt_stage == 0)));
  PCFM_LAUNCHED();
  return 0;
}

// channels per CTA: as many as fit in ~100 KB (two CTAs per SM) while keeping the grid at >= 2 CTAs per SM
static int devox_group(int b, int c, int r) {
  if (devox_fwd_smem(1, r) > kDevoxSmemLimit) return 0;
  int g = 8;
  while (g > 1 && (c % g != 0 || devox_fwd_smem(g, r) > 100 * 1024 || (long long)(c / g) * b * 2 < 3LL * num_sms()))
    g >>= 1;
  return g;
}

// inds/wgts only (C == 0 or nothing else to do)
__global__ void __launch_bounds__(256) devox_index_kernel(const float* __restrict__ coords,
                                                          int* __restrict__ inds, float* __restrict__ wgts,
                                                          int n, int r) {
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

extern "C" int pcfm_trilinear_devoxelize_forward(const float* coords, const float* features, int b, int c,
                                                 int n, int r, int is_training, float* outs, int* inds,
                                                 float* wgts, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0 || (long long)r * r * r > 0x7fffffffLL) return PCFM_E_BADARG;
  if (b == 0 || n == 0) return 0;
  if (!coords || (c > 0 && (!features || !outs))) return PCFM_E_BADARG;
  if (is_training && (!inds || !wgts)) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (c == 0) {
    if (is_training) {
      dim3 grid(cdiv(n, 256), b);
      devox_index_kernel<<<grid, 256, 0, st>>>(coords, inds, wgts, n, r);
      PCFM_LAUNCHED();
    }
    return 0;
  }
#define PCFM_DEVOX(G, RT) return launch_devox_t<G, RT, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st)
  // small grids: lanes along channels (conflict-free voxel rows); two CTAs per SM up to r = 8
  if (g_opt_devox_cl != 0 && (c % 32) == 0 && r <= (g_opt_devox_cl > 1 ? 10 : 8) && devox_cl_ok(n))
    return launch_devox_cl(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
  const int g = devox_group(b, c, r);
  if (g == 0) return launch_devox_t<1, 0, false>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
  if (r == 32 && g == 1) PCFM_DEVOX(1, 32);
#define PCFM_DEVOX_IL(G, RT) \
  return launch_devox_t<G, RT, true, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st)
  // measured (B=8, N=20000, C=256): r=16 0.158 -> 0.138 ms, r=8 0.116 -> 0.128 ms (a dense cloud on a tiny grid
  // mostly broadcasts, there the plain layout with TMA staging wins) => interleave from r = 12 up
  if (g_opt_devox_il != 0 && r >= 12) {
    if (r == 16 && g == 4) PCFM_DEVOX_IL(4, 16);
    if (g == 8) PCFM_DEVOX_IL(8, 0);
    if (g == 4) PCFM_DEVOX_IL(4, 0);
  }
#undef PCFM_DEVOX_IL
  if (r == 16 && g == 4) PCFM_DEVOX(4, 16);
  if (r == 16 && g == 2) PCFM_DEVOX(2, 16);
  if (r == 8 && g == 8) PCFM_DEVOX(8, 8);
  if (r == 8 && g == 4) PCFM_DEVOX(4, 8);
  switch (g) {
    case 8: PCFM_DEVOX(8, 0);
    case 4: PCFM_DEVOX(4, 0);
    case 2: PCFM_DEVOX(2, 0);
    default: PCFM_DEVOX(1, 0);
  }
#undef PCFM_DEVOX
}
