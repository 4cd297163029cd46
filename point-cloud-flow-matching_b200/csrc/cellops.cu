// cellops.cu — the two scatter-type PVConv primitives on the sliced-ELL cell plan:
//   avg_voxelize forward            out[b,c,v]  = mean of the features of the points in voxel v
//   trilinear_devoxelize backward   grad[b,c,v] = sum over the 8 cells around v of sum_p w_corner(p) * grad_y[b,c,p]
// One CTA per (batch, channel) row.  The randomly-read row (N floats) is staged in shared memory by TMA; a thread
// owns one voxel cell and walks that cell's records with coalesced loads (cellplan.cuh), so a point is visited
// ONCE per channel (the atomic formulation of the reference visits it 8 times for devoxelize backward).
#include "cellplan.cuh"
#include "rows.cuh"

namespace pcfm {

// -------------------------------------------------------------------------------------------------
// avg_voxelize forward
// -------------------------------------------------------------------------------------------------
template <bool ROW_SMEM>
__global__ void __launch_bounds__(1024, 2)
    vox_fwd_cell_kernel(const float* __restrict__ feat, const int* __restrict__ cells,
                        const int* __restrict__ slot0, const int* __restrict__ nlight,
                        const int* __restrict__ heavy, const int* __restrict__ nheavy,
                        const int* __restrict__ rec, float* __restrict__ out, int C, int N, int T, int lcap,
                        int hcap, int rcap, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = feat + ((long long)b * C + c) * N;
  float* o = out + ((long long)b * C + c) * T;
  if (ROW_SMEM && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
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
  __syncthreads();  // row visible; zero stores ordered before the sums below (same CTA)
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const int* rc = rec + (long long)b * rcap;
  const int nl = nlight[b];
  for (int ci = threadIdx.x; ci < nl; ci += blockDim.x) {
    const int d = cl[ci];
    const int n = d >> 24, t = d & 0xffffff;
    const int s0 = sl[ci];
    const float rcp = __frcp_rn((float)n);  // == (float)(1.0/(float)cnt), vox.cu:65
    float acc = 0.f;
    for (int i0 = 0; i0 < n; i0 += 4) {  // four independent record loads in flight per thread
      int p[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) p[k] = (i0 + k < n) ? rc[s0 + 32 * (i0 + k)] : -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (p[k] >= 0) {
          const float f = ROW_SMEM ? smem[p[k]] : __ldg(row_g + p[k]);
          acc = __fadd_rn(acc, __fmul_rn(f, rcp));  // product rounded, then added: the reference's atomicAdd(f*div)
        }
      }
    }
    o[t] = acc;
  }
  const int nh = nheavy[b];
  const int lane = threadIdx.x & 31;
  for (int h = threadIdx.x >> 5; h < nh; h += blockDim.x >> 5) {
    const int* hd = heavy + ((long long)b * hcap + h) * 3;
    const int t = hd[0], s0 = hd[1], n = hd[2];
    const float rcp = __frcp_rn((float)n);
    float acc = 0.f;
    for (int i = lane; i < n; i += 32) {
      const int p = rc[s0 + i];
      const float f = ROW_SMEM ? smem[p] : __ldg(row_g + p);
      acc = __fadd_rn(acc, __fmul_rn(f, rcp));
    }
    acc = warp_sum(acc);
    if (lane == 0) o[t] = acc;
  }
}

// -------------------------------------------------------------------------------------------------
// trilinear_devoxelize backward
// -------------------------------------------------------------------------------------------------
constexpr int kHeavyBatch = 256;

__device__ __forceinline__ void tri_weights(float fx, float fy, float fz, float* w) {
  // trilinear_devox.cu:45-59, same association as devox_fwd_kernel: the weights equal the saved wgts bit for bit
  const float x0 = 1.0f - fx, y0 = 1.0f - fy, z0 = 1.0f - fz;
  w[0] = __fmul_rn(__fmul_rn(x0, y0), z0);
  w[1] = __fmul_rn(__fmul_rn(x0, y0), fz);
  w[2] = __fmul_rn(__fmul_rn(x0, fy), z0);
  w[3] = __fmul_rn(__fmul_rn(x0, fy), fz);
  w[4] = __fmul_rn(__fmul_rn(fx, y0), z0);
  w[5] = __fmul_rn(__fmul_rn(fx, y0), fz);
  w[6] = __fmul_rn(__fmul_rn(fx, fy), z0);
  w[7] = __fmul_rn(__fmul_rn(fx, fy), fz);
}

// add the 8 corner sums of one cell into the shared-memory plane, corner k in phase k: inside a phase every thread
// targets cell + off_k with a distinct cell, so plain read-modify-writes never collide (no atomics, fixed order)
#define PCFM_CORNER_PHASES(active, cell, S)                                                     \
  {                                                                                             \
    int cx = 0, cy = 0, cz = 0;                                                                 \
    if (active) {                                                                               \
      cx = (cell) / r2;                                                                         \
      cy = ((cell) / r) % r;                                                                    \
      cz = (cell) % r;                                                                          \
    }                                                                                           \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) {                                             \
      const int dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;                                     \
      /* a corner outside the grid only ever carries weight 0 (coords <= r-1): skip it */       \
      if ((active) && cx + dx < r && cy + dy < r && cz + dz < r) {                              \
        const int v = (cell) + dx * r2 + dy * r + dz;                                           \
        plane[v] = __fadd_rn(plane[v], (S)[k]);                                                 \
      }                                                                                         \
      __syncthreads();                                                                          \
    }                                                                                           \
  }

template <bool ROW_SMEM>
__global__ void __launch_bounds__(1024, 1)
    devox_bwd_cell_kernel(const float* __restrict__ grad_y, const int* __restrict__ cells,
                          const int* __restrict__ slot0, const int* __restrict__ nlight,
                          const int* __restrict__ heavy, const int* __restrict__ nheavy,
                          const float4* __restrict__ rec, float* __restrict__ grad_x, int C, int N, int r,
                          int lcap, int hcap, int rcap, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int r2 = r * r, P = r2 * r;
  float* plane = smem;                                        // [P]
  float* hsum = smem + ((P + 3) & ~3);                        // [kHeavyBatch][8]
  float* row = hsum + kHeavyBatch * 8;                        // [N] (ROW_SMEM)
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = grad_y + ((long long)b * C + c) * N;
  float* o = grad_x + ((long long)b * C + c) * P;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const bool aligned = ROW_SMEM && use_tma && ((((uintptr_t)row_g) & 15) == 0);
  const int bulk = aligned ? (N & ~3) : 0;
  if (bulk > 0 && threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)bulk * 4u);
    for (uint32_t off = 0; off < (uint32_t)bulk * 4u; off += kBulkChunk)
      bulk_g2s((char*)row + off, (const char*)row_g + off, min(kBulkChunk, (uint32_t)bulk * 4u - off), &bar);
  }
  for (int i = threadIdx.x; i < ((P + 3) >> 2); i += blockDim.x)  // overlaps the row copy
    ((float4*)plane)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ROW_SMEM) {
    for (int i = bulk + threadIdx.x; i < N; i += blockDim.x) row[i] = __ldg(row_g + i);
    if (bulk > 0) mbar_wait(&bar, 0);
  }
  __syncthreads();
  const int* cl = cells + (long long)b * lcap;
  const int* sl = slot0 + (long long)b * lcap;
  const float4* rc = rec + (long long)b * rcap;
  const int nl = nlight[b];
  // ---- light cells: one thread per cell, rounds of blockDim cells ----
  for (int base = 0; base < nl; base += blockDim.x) {
    const int ci = base + threadIdx.x;
    const bool active = ci < nl;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cell = 0;
    if (active) {
      const int d = cl[ci];
      const int n = d >> 24;
      cell = d & 0xffffff;
      const int s0 = sl[ci];
      for (int i0 = 0; i0 < n; i0 += 2) {  // two independent record loads in flight per thread
        float4 q[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
          q[k] = (i0 + k < n) ? rc[s0 + 32 * (i0 + k)] : make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = __float_as_int(q[u].x);
          if (p < 0) continue;
          const float g = ROW_SMEM ? row[p] : __ldg(row_g + p);
          float w[8];
          tri_weights(q[u].y, q[u].z, q[u].w, w);
#pragma unroll
          for (int k = 0; k < 8; ++k) s[k] = __fadd_rn(s[k], __fmul_rn(w[k], g));  // trilinear_devox.cu:152-159
        }
      }
    }
    PCFM_CORNER_PHASES(active, cell, s)
  }
  // ---- heavy cells: one warp per cell, batches of kHeavyBatch cells ----
  const int nh = nheavy[b];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int hb = 0; hb < nh; hb += kHeavyBatch) {
    const int hn = min(kHeavyBatch, nh - hb);
    for (int h = wid; h < hn; h += nw) {
      const int* hd = heavy + ((long long)b * hcap + hb + h) * 3;
      const int s0 = hd[1], n = hd[2];
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < n; i += 32) {
        const float4 q = rc[s0 + i];
        const int p = __float_as_int(q.x);
        const float g = ROW_SMEM ? row[p] : __ldg(row_g + p);
        float w[8];
        tri_weights(q.y, q.z, q.w, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = __fadd_rn(s[k], __fmul_rn(w[k], g));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float tot = warp_sum(s[k]);
        if (lane == k) hsum[h * 8 + k] = tot;
      }
    }
    __syncthreads();
    {
      const bool active = (int)threadIdx.x < hn;
      int cell = 0;
      float s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = active ? hsum[threadIdx.x * 8 + k] : 0.f;
      if (active) cell = heavy[((long long)b * hcap + hb + threadIdx.x) * 3];
      PCFM_CORNER_PHASES(active, cell, s)
    }
  }
  // ---- write the plane: one bulk (TMA) store when aligned, else a coalesced copy ----
  if ((((uintptr_t)o) & 15) == 0 && (P & 3) == 0 && use_tma) {
    fence_proxy_async();  // generic-proxy writes to `plane` -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t off = 0; off < (uint32_t)P * 4u; off += kBulkChunk)
        bulk_s2g((char*)o + off, (const char*)plane + off, min(kBulkChunk, (uint32_t)P * 4u - off));
      bulk_commit();
      bulk_wait_read();  // the CTA may exit once the source has been read out of shared memory
    }
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) o[i] = plane[i];
  }
}

static size_t devox_bwd_smem(int n, int r, bool row_smem) {
  const size_t p = (size_t)r * r * r;
  return (((p + 3) & ~(size_t)3) + kHeavyBatch * 8 + (row_smem ? n : 0)) * sizeof(float);
}
constexpr size_t kSmemLimit = 227 * 1024 - 512;  // opt-in maximum (232448 B) minus the static shared memory

// defined in rows.cu (A/B variants on the generic sorted-segment engine / global atomics)
int avg_voxelize_forward_generic(const float* features, const int* coords, int b, int c, int n, int r, float* out,
                                 int* ind, int* cnt, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                 int variant);

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
  if (!coords || !ind || !cnt || (c > 0 && (!features || !out))) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_opt_vox_fwd != 0)
    return avg_voxelize_forward_generic(features, coords, b, c, n, r, out, ind, cnt, workspace,
                                             workspace_bytes, st, g_opt_vox_fwd);
  const int t = r * r * r;
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = cellplan_build(coords, nullptr, r, ind, b, n, t, cnt, &plan, &ws, st);
  if (rc || c == 0) return rc;
  dim3 grid(c, b);
  static bool attr_set = false;
  if (!attr_set) {
    PCFM_CUDA(cudaFuncSetAttribute(vox_fwd_cell_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kMaxStageFloats * sizeof(float))));
    attr_set = true;
  }
  if (n <= kMaxStageFloats)
    vox_fwd_cell_kernel<true><<<grid, 1024, (size_t)n * sizeof(float), st>>>(
        features, plan.cells, plan.slot0, plan.nlight, plan.heavy, plan.nheavy, (const int*)plan.rec, out, c, n, t,
        plan.lcap, plan.hcap, plan.rcap, g_opt_stage == 0);
  else
    vox_fwd_cell_kernel<false><<<grid, 1024, 0, st>>>(features, plan.cells, plan.slot0, plan.nlight, plan.heavy,
                                                      plan.nheavy, (const int*)plan.rec, out, c, n, t, plan.lcap,
                                                      plan.hcap, plan.rcap, 0);
  PCFM_LAUNCHED();
  return 0;
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
  cudaStream_t st = (cudaStream_t)stream;
  const long long tl = (long long)r * r * r;
  if (devox_bwd_smem(0, r, false) > kSmemLimit || tl > (1 << 24)) return PCFM_E_BADARG;  // plane must fit (r <= 37)
  const int t = (int)tl;
  if (n == 0) {
    PCFM_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)b * c * t * sizeof(float), st));
    return 0;
  }
  Workspace ws(workspace, workspace_bytes);
  CellPlan plan;
  int rc = cellplan_build(nullptr, coords, r, nullptr, b, n, t, nullptr, &plan, &ws, st);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PCFM_CUDA(cudaFuncSetAttribute(devox_bwd_cell_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kSmemLimit));
    PCFM_CUDA(cudaFuncSetAttribute(devox_bwd_cell_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kSmemLimit));
    attr_set = true;
  }
  dim3 grid(c, b);
  const bool row_smem = devox_bwd_smem(n, r, true) <= kSmemLimit;
  // small planes (r <= 16): two CTAs of 512 threads per SM overlap one CTA's staging with the other's compute
  const int threads = (row_smem && devox_bwd_smem(n, r, true) <= 110 * 1024) ? 512 : 1024;
  if (row_smem)
    devox_bwd_cell_kernel<true><<<grid, threads, devox_bwd_smem(n, r, true), st>>>(
        grad_y, plan.cells, plan.slot0, plan.nlight, plan.heavy, plan.nheavy, (const float4*)plan.rec, grad_x, c, n,
        r, plan.lcap, plan.hcap, plan.rcap, g_opt_stage == 0);
  else
    devox_bwd_cell_kernel<false><<<grid, threads, devox_bwd_smem(n, r, false), st>>>(
        grad_y, plan.cells, plan.slot0, plan.nlight, plan.heavy, plan.nheavy, (const float4*)plan.rec, grad_x, c, n,
        r, plan.lcap, plan.hcap, plan.rcap, g_opt_stage == 0);
  PCFM_LAUNCHED();
  return 0;
}

}  // extern "C"
