// devox.cu — trilinear devoxelisation forward (trilinear_devox.cu:21-105 semantics).
//
// One CTA owns G consecutive channel planes of one batch element: the planes (G * r^3 floats) are staged in
// shared memory with 1-D TMA bulk copies, then every point gathers its 8 corners from shared memory and the
// result row is written coalesced.  Indices / weights are computed once per point and reused for G channels.
#include "rows.cuh"

namespace pcfm {

template <int G, bool PLANE_SMEM>
__global__ void __launch_bounds__(1024, 1)
    devox_fwd_kernel(const float* __restrict__ coords, const float* __restrict__ feat,
                     float* __restrict__ outs, int* __restrict__ inds, float* __restrict__ wgts, int C, int n,
                     int r, int training, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int c0 = blockIdx.x * G, b = blockIdx.y;
  const int r2 = r * r, P = r2 * r;
  const float* pg = feat + ((long long)b * C + c0) * P;
  if (PLANE_SMEM) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    stage_floats(smem, pg, G * P, &bar, 0, use_tma != 0);
  }
  const float* cb = coords + (long long)b * 3 * n;
  float* ob = outs + ((long long)b * C + c0) * n;
  const bool save = training && blockIdx.x == 0;
  int* ib = inds ? inds + (long long)b * 8 * n : nullptr;
  float* wb = wgts ? wgts + (long long)b * 8 * n : nullptr;
  const float hi = (float)(r - 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // coords are contractually in [0, r-1]; the clamp is a no-op there and keeps every index in range otherwise
    const float x = fminf(fmaxf(cb[i], 0.f), hi);
    const float y = fminf(fmaxf(cb[(long long)i + n], 0.f), hi);
    const float z = fminf(fmaxf(cb[(long long)i + 2LL * n], 0.f), hi);
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
    const int ox = (x1 > 0.f) ? r2 : 0, oy = (y1 > 0.f) ? r : 0, oz = (z1 > 0.f) ? 1 : 0;
    int id[8];
    id[0] = (int)xl * r2 + (int)yl * r + (int)zl;
    id[1] = id[0] + oz;
    id[2] = id[0] + oy;
    id[3] = id[2] + oz;
    id[4] = id[0] + ox;
    id[5] = id[4] + oz;
    id[6] = id[4] + oy;
    id[7] = id[6] + oz;
    if (save) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ib[(long long)k * n + i] = id[k];
        wb[(long long)k * n + i] = w[k];
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float acc;
      if (PLANE_SMEM) {
        const float* pl = smem + g * P;
        acc = w[0] * pl[id[0]];
#pragma unroll
        for (int k = 1; k < 8; ++k) acc = fmaf(w[k], pl[id[k]], acc);
      } else {
        const float* pl = pg + (long long)g * P;
        acc = w[0] * __ldg(pl + id[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) acc = fmaf(w[k], __ldg(pl + id[k]), acc);
      }
      ob[(long long)g * n + i] = acc;
    }
  }
}

template <int G, bool PLANE_SMEM>
static int launch_devox_t(const float* coords, const float* feat, float* outs, int* inds, float* wgts, int b,
                          int c, int n, int r, int training, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = devox_fwd_kernel<G, PLANE_SMEM>;
  const size_t smem = PLANE_SMEM ? (size_t)G * r * r * r * sizeof(float) : 0;
  if (!attr_set) {
    PCFM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kMaxStageFloats * sizeof(float))));
    attr_set = true;
  }
  dim3 grid(c / G, b);
  kern<<<grid, 1024, smem, st>>>(coords, feat, outs, inds, wgts, c, n, r, training, g_opt_stage == 0);
  PCFM_LAUNCHED();
  return 0;
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
  const int p = r * r * r;
  switch (pick_group(b, c, p)) {
    case 0: return launch_devox_t<1, false>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
    case 8: return launch_devox_t<8, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
    case 4: return launch_devox_t<4, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
    case 2: return launch_devox_t<2, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
    default: return launch_devox_t<1, true>(coords, features, outs, inds, wgts, b, c, n, r, is_training, st);
  }
}
