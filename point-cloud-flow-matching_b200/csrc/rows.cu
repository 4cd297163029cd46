// rows.cu — row engines + the entry points built directly on them:
//   avg_voxelize fwd/bwd, grouping fwd/bwd, gather fwd/bwd, 3-NN interpolate bwd, trilinear devoxelize bwd.
#include "rows.cuh"

namespace pcfm {

// =================================================================================================
// row_segsum: grid (C, B), 1024 threads, one thread per output element, segment walked in plan order.
// =================================================================================================
template <int MODE, bool ROW_SMEM>
__global__ void __launch_bounds__(1024, 1)
    row_segsum_kernel(const float* __restrict__ src, const int* __restrict__ start,
                      const int* __restrict__ ssrc, const float* __restrict__ sw,
                      const int* __restrict__ heavy, const int* __restrict__ nheavy, int hcap,
                      float* __restrict__ out, int C, int S, int T, int E, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int c = blockIdx.x, b = blockIdx.y;
  const float* row_g = src + ((long long)b * C + c) * S;
  if (ROW_SMEM) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    stage_floats(smem, row_g, S, &bar, 0, use_tma != 0);
  }
  const int* st = start + (long long)b * (T + 1);
  const int* ss = ssrc + (long long)b * E;
  const float* wv = (MODE == SEG_WEIGHTED) ? sw + (long long)b * E : nullptr;
  float* o = out + ((long long)b * C + c) * T;
  // phase A: one thread per target, segments of at most kHeavySeg edges (ascending edge order)
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int s0 = st[t], s1 = st[t + 1];
    if (s1 - s0 > kHeavySeg) continue;  // phase B
    float acc = 0.f;
    if (s1 > s0) {
      float rcp = 1.f;
      if (MODE == SEG_MEAN) rcp = __frcp_rn((float)(s1 - s0));  // == (float)(1.0/(float)cnt), vox.cu:65
      for (int j = s0; j < s1; ++j) {
        const int si = ss[j];
        const float f = ROW_SMEM ? smem[si] : __ldg(row_g + si);
        // the reference rounds the product, then adds it with an atomic (vox.cu:67, trilinear_devox.cu:152):
        // keep product and sum as two roundings so a one-edge segment is bit-identical.
        if (MODE == SEG_MEAN) acc = __fadd_rn(acc, __fmul_rn(f, rcp));
        else if (MODE == SEG_WEIGHTED) acc = __fadd_rn(acc, __fmul_rn(wv[j], f));
        else acc = __fadd_rn(acc, f);
      }
    }
    o[t] = acc;
  }
  // phase B: one warp per long segment: lanes stride the segment (coalesced plan reads), fixed-shape butterfly
  const int nh = min(nheavy[b], hcap);
  const int* hv = heavy + (long long)b * hcap;
  const int lane = threadIdx.x & 31;
  for (int h = threadIdx.x >> 5; h < nh; h += blockDim.x >> 5) {
    const int t = hv[h];
    const int s0 = st[t], s1 = st[t + 1];
    float rcp = 1.f;
    if (MODE == SEG_MEAN) rcp = __frcp_rn((float)(s1 - s0));
    float acc = 0.f;
    for (int j = s0 + lane; j < s1; j += 32) {
      const int si = ss[j];
      const float f = ROW_SMEM ? smem[si] : __ldg(row_g + si);
      if (MODE == SEG_MEAN) acc = __fadd_rn(acc, __fmul_rn(f, rcp));
      else if (MODE == SEG_WEIGHTED) acc = __fadd_rn(acc, __fmul_rn(wv[j], f));
      else acc = __fadd_rn(acc, f);
    }
    acc = warp_sum(acc);
    if (lane == 0) o[t] = acc;
  }
}

template <int MODE, bool ROW_SMEM>
static int launch_segsum_t(const float* src, const CsrPlan& plan, float* out, int b, int c, int s, int t,
                           int e, cudaStream_t st) {
  auto kern = row_segsum_kernel<MODE, ROW_SMEM>;
  const size_t smem = ROW_SMEM ? (size_t)s * sizeof(float) : 0;
  PCFM_CUDA(allow_dyn_smem((const void*)kern, (int)(kMaxStageFloats * sizeof(float))));
  dim3 grid(c, b);
  kern<<<grid, 1024, smem, st>>>(src, plan.start, plan.ssrc, plan.sw, plan.heavy, plan.nheavy, plan.hcap, out, c,
                                 s, t, e, g_opt_stage == 0);
  PCFM_LAUNCHED();
  return 0;
}

int launch_row_segsum(SegMode mode, const float* src, const CsrPlan& plan, float* out, int b, int c, int s,
                      int t, int e, cudaStream_t st) {
  if (b <= 0 || c <= 0 || t <= 0) return 0;
  const bool in_smem = s <= kMaxStageFloats;
  switch (mode) {
    case SEG_WEIGHTED:
      return in_smem ? launch_segsum_t<SEG_WEIGHTED, true>(src, plan, out, b, c, s, t, e, st)
                     : launch_segsum_t<SEG_WEIGHTED, false>(src, plan, out, b, c, s, t, e, st);
    case SEG_UNIT:
      return in_smem ? launch_segsum_t<SEG_UNIT, true>(src, plan, out, b, c, s, t, e, st)
                     : launch_segsum_t<SEG_UNIT, false>(src, plan, out, b, c, s, t, e, st);
    default:
      return in_smem ? launch_segsum_t<SEG_MEAN, true>(src, plan, out, b, c, s, t, e, st)
                     : launch_segsum_t<SEG_MEAN, false>(src, plan, out, b, c, s, t, e, st);
  }
}

// =================================================================================================
// row_gather: grid (C/G, B), 1024 threads; G consecutive channel planes staged per CTA so the index /
// weight stream is read once per G channels.
// =================================================================================================
template <int K, int G, bool PLANE_SMEM, bool HAS_W>
__global__ void __launch_bounds__(1024, 1)
    row_gather_kernel(const float* __restrict__ planes, const int* __restrict__ idx,
                      const float* __restrict__ w, float* __restrict__ out, int C, int P, int J, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int c0 = blockIdx.x * G, b = blockIdx.y;
  const float* pg = planes + ((long long)b * C + c0) * P;
  // ---- issue the plane copy (TMA bulk when aligned), do not wait yet ----
  const int total = G * P;
  const bool aligned = PLANE_SMEM && use_tma && ((((uintptr_t)pg) & 15) == 0);
  const int bulk = aligned ? (total & ~3) : 0;
  if (PLANE_SMEM) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
      if (bulk > 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)bulk * 4u);
        for (uint32_t off = 0; off < (uint32_t)bulk * 4u; off += kBulkChunk)
          bulk_g2s((char*)smem + off, (const char*)pg + off, min(kBulkChunk, (uint32_t)bulk * 4u - off), &bar);
      }
    }
    for (int i = bulk + threadIdx.x; i < total; i += blockDim.x) smem[i] = __ldg(pg + i);
  }
  griddep_wait();  // the planes are an input of the operator; idx / w may come from the previous kernel (PDL)
  griddep_launch_dependents();
  const int* ib = idx + (long long)b * K * J;
  const float* wb = HAS_W ? w + (long long)b * K * J : nullptr;
  float* ob = out + ((long long)b * C + c0) * J;
  const unsigned pmax = (unsigned)(P - 1);
  // U outputs per thread and round, software-pipelined: the index / weight loads of round i+1 are in flight while round
  // i gathers, and those of round 0 while the plane copy lands (the loop was bound by the L2 latency of these loads)
  constexpr int U = (K == 1) ? 4 : 2;
  unsigned id[U][K], idn[U][K];
  float wk[U][K], wkn[U][K];
  auto load = [&](int j0, unsigned (&di)[U][K], float (&dw)[U][K]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * blockDim.x;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        // min(): memory safety only; valid indices are unchanged
        di[u][k] = (j < J) ? min((unsigned)ib[(long long)k * J + j], pmax) : 0u;
        dw[u][k] = (HAS_W && j < J) ? wb[(long long)k * J + j] : 1.f;
      }
    }
  };
  // PIPE costs ~12 registers: only where one CTA owns the SM anyway (G == 1: big planes); the small-plane variants stay
  // at 32 registers so that two CTAs of 1024 threads fit
  constexpr bool PIPE = (G == 1);
  if (PIPE) load(threadIdx.x, id, wk);
  if (PLANE_SMEM) {
    __syncthreads();  // mbarrier initialised (thread 0) and the tail loads done before anybody waits / reads
    if (bulk > 0) mbar_wait(&bar, 0);
  }
  for (int j0 = threadIdx.x; j0 < J; j0 += blockDim.x * U) {
    if (PIPE) load(j0 + blockDim.x * U, idn, wkn);  // out-of-range rounds load nothing (predicated)
    else load(j0, id, wk);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * blockDim.x;
      if (j >= J) break;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float f[K];
#pragma unroll
        for (int k = 0; k < K; ++k)
          f[k] = PLANE_SMEM ? smem[g * P + id[u][k]] : __ldg(pg + (long long)g * P + id[u][k]);
        float acc;
        if (K == 1) {
          acc = HAS_W ? __fmul_rn(f[0], wk[u][0]) : f[0];
        } else {
          acc = f[0] * wk[u][0];
#pragma unroll
          for (int k = 1; k < K; ++k) acc = fmaf(f[k], wk[u][k], acc);
        }
        ob[(long long)g * J + j] = acc;
      }
    }
    if (PIPE) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          id[u][k] = idn[u][k];
          wk[u][k] = wkn[u][k];
        }
    }
  }
}

template <int K, int G, bool PLANE_SMEM, bool HAS_W>
static int launch_gather_t(const float* planes, const int* idx, const float* w, float* out, int b, int c,
                           int p, int j, cudaStream_t st, bool dependent) {
  auto kern = row_gather_kernel<K, G, PLANE_SMEM, HAS_W>;
  const size_t smem = PLANE_SMEM ? (size_t)G * p * sizeof(float) : 0;
  PCFM_CUDA(allow_dyn_smem((const void*)kern, (int)(kMaxStageFloats * sizeof(float))));
  dim3 grid(c / G, b);
  PCFM_CUDA(launch_dependent(kern, grid, dim3(1024), smem, st, dependent && g_opt_pdl != 0, planes, idx, w, out, c, p,
                             j, (int)(g_opt_stage == 0)));
  PCFM_LAUNCHED();
  return 0;
}

// channels per CTA: as many as fit in ~96 KB while keeping >= 2 CTAs per SM in flight
int pick_group(int b, int c, int p) {
  if (p > kMaxStageFloats) return 0;  // plane does not fit: read through L2
  int g = 8;
  while (g > 1 && (c % g != 0 || (long long)g * p * 4 > 98304 || (long long)(c / g) * b < 2LL * num_sms()))
    g >>= 1;
  return g;
}

template <int K, bool HAS_W>
static int launch_gather_k(const float* planes, const int* idx, const float* w, float* out, int b, int c,
                           int p, int j, cudaStream_t st, bool dep) {
  switch (pick_group(b, c, p)) {
    case 0: return launch_gather_t<K, 1, false, HAS_W>(planes, idx, w, out, b, c, p, j, st, dep);
    case 8: return launch_gather_t<K, 8, true, HAS_W>(planes, idx, w, out, b, c, p, j, st, dep);
    case 4: return launch_gather_t<K, 4, true, HAS_W>(planes, idx, w, out, b, c, p, j, st, dep);
    case 2: return launch_gather_t<K, 2, true, HAS_W>(planes, idx, w, out, b, c, p, j, st, dep);
    default: return launch_gather_t<K, 1, true, HAS_W>(planes, idx, w, out, b, c, p, j, st, dep);
  }
}

// dependent: the previous kernel on the stream (which produced idx / w) has called griddep_launch_dependents(), so
// this one may stage its planes while that kernel is still running
int launch_row_gather(int k, const float* planes, const int* idx, const float* w, float* out, int b, int c,
                      int p, int j, cudaStream_t st, bool dependent) {
  if (b <= 0 || c <= 0 || j <= 0) return 0;
  if (p <= 0) return PCFM_E_BADARG;
  if (k == 1) return w ? launch_gather_k<1, true>(planes, idx, w, out, b, c, p, j, st, dependent)
                       : launch_gather_k<1, false>(planes, idx, nullptr, out, b, c, p, j, st, dependent);
  if (k == 3 && w) return launch_gather_k<3, true>(planes, idx, w, out, b, c, p, j, st, dependent);
  return PCFM_E_BADARG;
}

// =================================================================================================
// small helper kernels
// =================================================================================================
// rcp[b,i] = 1/cnt[b,ind[b,i]]  (0 when the count is not positive: the reference skips such points, vox.cu:103)
__global__ void __launch_bounds__(256) point_rcp_kernel(const int* __restrict__ ind,
                                                        const int* __restrict__ cnt, int n, int s,
                                                        float* __restrict__ rcp) {
  griddep_wait();               // chained PDL: may have started before its predecessor finished
  griddep_launch_dependents();  // the gather kernel behind this one may start staging its planes
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = ind[(long long)b * n + i];
  int k = 0;
  if (v >= 0 && v < s) k = cnt[(long long)b * s + v];
  rcp[(long long)b * n + i] = k > 0 ? __frcp_rn((float)k) : 0.f;
}

// A/B variant: the reference's formulation (global fp32 RED per point-channel) on a full grid.
__global__ void __launch_bounds__(256) vox_scatter_atomic_kernel(const float* __restrict__ feat,
                                                                 const int* __restrict__ ind,
                                                                 const int* __restrict__ cnt, int C, int n,
                                                                 int s, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = ind[(long long)b * n + i];
  if (v < 0 || v >= s) return;
  const int k = cnt[(long long)b * s + v];
  if (k <= 0) return;
  const float f = feat[((long long)b * C + c) * n + i];
  atomicAdd(out + ((long long)b * C + c) * s + v, __fmul_rn(f, __frcp_rn((float)k)));
}

__global__ void __launch_bounds__(256) devox_scatter_atomic_kernel(const float* __restrict__ grad_y,
                                                                   const int* __restrict__ inds,
                                                                   const float* __restrict__ wgts, int C,
                                                                   int n, int s, float* __restrict__ grad_x) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad_y[((long long)b * C + c) * n + i];
  float* gx = grad_x + ((long long)b * C + c) * s;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int v = inds[((long long)b * 8 + k) * n + i];
    const float w = wgts[((long long)b * 8 + k) * n + i];
    if (v >= 0 && v < s) atomicAdd(gx + v, __fmul_rn(w, g));
  }
}

// Scatter-add of one row into a shared-memory row: the CTA zeroes grad_x[b, c, :] in shared memory, streams
// grad_y[b, c, :] (E floats, coalesced) beside indices[b, :] (L2-resident: shared by the C rows of a cloud) and adds
// every element to its target with a shared-memory atomic; the row leaves with one bulk (TMA) store.  No plan, no
// memset, no global atomics (grouping.cu:36-58 issues one global atomicAdd per element).  Ball queries pad a
// neighbourhood with copies of its first neighbour: runs of equal targets in consecutive lanes are summed in the warp
// first (segmented suffix sum, 5 shuffles) so that a run costs one atomic instead of a CAS loop over its lanes.
// Rows longer than one CTA's shared memory are cut into `tiles` slices of `slice` floats, one CTA each: grid
// (C * tiles, B); every CTA of a row streams all E elements and keeps the ones that land in its slice.  (Dealing the
// elements to the CTAs of a cluster instead, with the adds going to the owner through distributed shared memory, was
// measured 7x slower: 13.8 ms against 1.9 ms at N = 100000 — remote float atomics are generic-address CAS loops.)
constexpr int kScatterUnroll = 16;
// WEIGHTED (3-NN interpolate backward, interpolate.cu:158-181): element e carries weight wgt[b, e] and the source value
// grad_y[b, c, e % S] (E = 3 * S: the three neighbours of every point); product and sum are rounded separately, as the
// reference's atomicAdd(grad_x + idx, grad_y * weight) does.
template <int THREADS, bool TILED, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
    row_scatter_kernel(const float* __restrict__ grad_y, const int* __restrict__ indices, const float* __restrict__ wgt,
                       float* __restrict__ grad_x, int C, int S, int T, int E, int tiles, int slice, int use_tma) {
  extern __shared__ __align__(128) float smem[];
  const int c = TILED ? blockIdx.x / tiles : blockIdx.x, tile = TILED ? blockIdx.x - c * tiles : 0, b = blockIdx.y;
  const float* src = grad_y + ((long long)b * C + c) * S;
  const int* idx = indices + (long long)b * E;
  const float* wv = WEIGHTED ? wgt + (long long)b * E : nullptr;
  const int t0 = tile * slice, count = max(0, min(slice, T - t0));  // this CTA's part of the row
  float* o = grad_x + ((long long)b * C + c) * T + t0;
  for (int i = threadIdx.x; i < (slice >> 2); i += THREADS) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kStep = 32 * kScatterUnroll, kWarps = THREADS >> 5;
  const bool few_wraps = (long long)E <= 3LL * S;
  for (int base = warp * kStep; base < E; base += kWarps * kStep) {
    int t[kScatterUnroll];
    float v[kScatterUnroll];
#pragma unroll
    for (int k = 0; k < kScatterUnroll; ++k) {  // all loads of the step in flight before the first is used
      const int e = base + k * 32 + lane;
      t[k] = e < E ? __ldg(idx + e) - t0 : -1;
      if (WEIGHTED) {
        int si = e;  // e % S without a division for the 3-NN shape (E = 3 * S)
        if (few_wraps) {
          if (si >= S) si -= S;
          if (si >= S) si -= S;
        } else {
          si = e % S;
        }
        v[k] = e < E ? __fmul_rn(__ldg(wv + e), __ldg(src + si)) : 0.f;
      }
      else v[k] = e < E ? (TILED ? __ldg(src + e) : __ldcs(src + e)) : 0.f;  // one pass: streaming, evict first
    }
#pragma unroll
    for (int k = 0; k < kScatterUnroll; ++k) {
      const int prev = __shfl_up_sync(0xffffffffu, t[k], 1);
      const bool head = lane == 0 || t[k] != prev;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      float x = v[k];
      if (heads != 0xffffffffu) {  // warp-uniform: some run is longer than one lane
        const unsigned above = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));
        const int end = above ? __ffs(above) - 1 : 32;  // this lane's run ends before lane `end`
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float y = __shfl_down_sync(0xffffffffu, x, d);
          if (lane + d < end) x += y;
        }
      }
      if (head && (unsigned)t[k] < (unsigned)count) atomicAdd(smem + t[k], x);
    }
  }
  __syncthreads();
  if (use_tma && (((uintptr_t)o) & 15) == 0 && (count & 3) == 0 && count > 0) {
    fence_proxy_async();  // generic-proxy writes to the row -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t off = 0; off < (uint32_t)count * 4u; off += kBulkChunk)
        bulk_s2g((char*)o + off, (const char*)smem + off, min(kBulkChunk, (uint32_t)count * 4u - off));
      bulk_commit();
      bulk_wait_read();
    }
  } else {
    for (int i = threadIdx.x; i < count; i += THREADS) o[i] = smem[i];
  }
}

constexpr int kScatterTwoCtaFloats = 28672;  // 112 KB: two CTAs per SM
constexpr int kScatterMaxTiles = 8;          // beyond 8 slices of 28672 points the sorted-segment path takes over
static bool scatter_rows_fits(long long t) { return t <= (long long)kScatterMaxTiles * kScatterTwoCtaFloats; }
// s = length of a source row (e for the unit form, where element e reads grad_y[b, c, e])
static int scatter_rows_smem(const float* grad_y, const int* indices, const float* wgt, int b, int c, int s, int t, int e,
                             float* grad_x, cudaStream_t st) {
  const int tiles = t <= kMaxStageFloats ? 1 : cdiv(t, kScatterTwoCtaFloats);
  const int slice = (cdiv(t, tiles) + 3) & ~3;
  const size_t smem = (size_t)slice * sizeof(float);
  const int tma = (int)(g_opt_stage == 0);
  dim3 grid(c * tiles, b);
#define PCFM_SCAT(TH, TI, WE)                                                                                        \
  do {                                                                                                               \
    PCFM_CUDA(allow_dyn_smem((const void*)row_scatter_kernel<TH, TI, WE>, (int)(kMaxStageFloats * sizeof(float))));  \
    row_scatter_kernel<TH, TI, WE><<<grid, TH, smem, st>>>(grad_y, indices, wgt, grad_x, c, s, t, e, tiles, slice,   \
                                                           tma);                                                     \
  } while (0)
#define PCFM_SCAT_W(TH, TI) \
  do {                      \
    if (wgt) PCFM_SCAT(TH, TI, true); else PCFM_SCAT(TH, TI, false); \
  } while (0)
  if (tiles > 1) PCFM_SCAT_W(512, true);
  else if (slice <= kScatterTwoCtaFloats) PCFM_SCAT_W(512, false);
  else PCFM_SCAT_W(1024, false);
#undef PCFM_SCAT_W
#undef PCFM_SCAT
  PCFM_LAUNCHED();
  return 0;
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

// ---------------------------------------------------------------------------------------------
// avg_voxelize
// ---------------------------------------------------------------------------------------------
}  // extern "C"

namespace pcfm {
// A/B variants of avg_voxelize forward kept for measurement (pcfm_set_option("vox_fwd", v)):
//   1 = the reference's formulation (global fp32 atomics) on a full grid, 2 = generic sorted-segment engine
int avg_voxelize_forward_generic(const float* features, const int* coords, int b, int c, int n, int r, float* out,
                                 int* ind, int* cnt, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                 int variant) {
  const int s = r * r * r;
  Workspace ws(workspace, workspace_bytes);
  CsrPlan plan;
  int rc = csr_build(nullptr, coords, r, ind, nullptr, b, n, s, n > 0 ? n : 1, cnt, &plan, &ws, st);
  if (rc) return rc;
  if (c == 0) return 0;
  if (variant == 1) {
    PCFM_CUDA(cudaMemsetAsync(out, 0, (size_t)b * c * s * sizeof(float), st));
    if (n > 0) {
      dim3 grid(cdiv(n, 256), c, b);
      vox_scatter_atomic_kernel<<<grid, 256, 0, st>>>(features, ind, cnt, c, n, s, out);
      PCFM_LAUNCHED();
    }
    return 0;
  }
  return launch_row_segsum(SEG_MEAN, features, plan, out, b, c, n, s, n, st);
}
}  // namespace pcfm

extern "C" {

size_t pcfm_avg_voxelize_backward_workspace(int b, int c, int n, int s) {
  (void)c; (void)s;
  if (b <= 0 || n <= 0) return 256;
  return align256((size_t)b * n * sizeof(float)) + 256;
}

int pcfm_avg_voxelize_backward(const float* grad_y, const int* ind, const int* cnt, int b, int c, int n, int s,
                               float* grad_x, void* workspace, size_t workspace_bytes, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || s <= 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!grad_y || !ind || !cnt || !grad_x) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  float* rcp = ws.take<float>((size_t)b * n);
  if (!ws.ok()) return PCFM_E_WORKSPACE;
  dim3 grid(cdiv(n, 256), b);
  PCFM_CUDA(launch_dependent(point_rcp_kernel, grid, dim3(256), 0, st, g_opt_pdl >= 2, ind, cnt, n, s, rcp));
  PCFM_LAUNCHED();
  return launch_row_gather(1, grad_y, ind, rcp, grad_x, b, c, s, n, st, true);
}

// ---------------------------------------------------------------------------------------------
// trilinear_devoxelize backward (generic: arbitrary inds / wgts)
// ---------------------------------------------------------------------------------------------
size_t pcfm_trilinear_devoxelize_backward_workspace(int b, int c, int n, int r) {
  (void)c;
  if (b <= 0 || n < 0 || r <= 0) return 0;
  const long long s = (long long)r * r * r;
  return csr_workspace_bytes(b, 8LL * n, s, true) + align256((size_t)b * s * sizeof(int)) + 256;
}

int pcfm_trilinear_devoxelize_backward(const float* grad_y, const int* inds, const float* wgts, int b, int c,
                                       int n, int r, float* grad_x, void* workspace, size_t workspace_bytes,
                                       pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || r <= 0 || (long long)r * r * r > 0x7fffffffLL || 8LL * n > 0x7fffffffLL)
    return PCFM_E_BADARG;
  if (b == 0 || c == 0) return 0;
  if (!grad_x || (n > 0 && (!grad_y || !inds || !wgts))) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int s = r * r * r;
  if (g_opt_devox_bwd == 1 || n == 0) {
    PCFM_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)b * c * s * sizeof(float), st));
    if (n > 0) {
      dim3 grid(cdiv(n, 256), c, b);
      devox_scatter_atomic_kernel<<<grid, 256, 0, st>>>(grad_y, inds, wgts, c, n, s, grad_x);
      PCFM_LAUNCHED();
    }
    return 0;
  }
  Workspace ws(workspace, workspace_bytes);
  int* cnt = ws.take<int>((size_t)b * s);
  if (!ws.ok()) return PCFM_E_WORKSPACE;
  CsrPlan plan;
  int rc = csr_build(inds, nullptr, 0, nullptr, wgts, b, 8 * n, s, n, cnt, &plan, &ws, st);
  if (rc) return rc;
  return launch_row_segsum(SEG_WEIGHTED, grad_y, plan, grad_x, b, c, n, s, 8 * n, st);
}

// ---------------------------------------------------------------------------------------------
// grouping / gather
// ---------------------------------------------------------------------------------------------
int pcfm_grouping_forward(const float* features, const int* indices, int b, int c, int n, int m, int u,
                          float* out, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0 || u < 0 || (long long)m * u > 0x7fffffffLL) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || m == 0 || u == 0) return 0;
  if (!features || !indices || !out || n == 0) return PCFM_E_BADARG;
  return launch_row_gather(1, features, indices, nullptr, out, b, c, n, m * u, (cudaStream_t)stream);
}

static size_t scatter_ws(int b, long long e, long long t) {
  return csr_workspace_bytes(b, e, t, false) + align256((size_t)b * t * sizeof(int)) + 256;
}

__global__ void __launch_bounds__(256) gather_bwd_atomic_kernel(const float* __restrict__ grad_y,
                                                                const int* __restrict__ indices, int c, int n, int m,
                                                                float* __restrict__ grad_x) {
  const int b = blockIdx.z, ch = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int t = indices[(long long)b * m + j];
  if (t < 0 || t >= n) return;
  atomicAdd(grad_x + ((long long)b * c + ch) * n + t, grad_y[((long long)b * c + ch) * m + j]);
}

static int scatter_unit(const float* grad_y, const int* indices, int b, int c, int t, int e, float* grad_x,
                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  Workspace ws(workspace, workspace_bytes);
  int* cnt = ws.take<int>((size_t)b * t);
  if (!ws.ok()) return PCFM_E_WORKSPACE;
  CsrPlan plan;
  int rc = csr_build(indices, nullptr, 0, nullptr, nullptr, b, e, t, e > 0 ? e : 1, cnt, &plan, &ws, st);
  if (rc) return rc;
  return launch_row_segsum(SEG_UNIT, grad_y, plan, grad_x, b, c, e, t, e, st);
}

size_t pcfm_grouping_backward_workspace(int b, int c, int n, int m, int u) {
  (void)c;
  if (b <= 0 || n <= 0) return 256;
  return scatter_ws(b, (long long)m * u, n);
}

int pcfm_grouping_backward(const float* grad_y, const int* indices, int b, int c, int n, int m, int u,
                           float* grad_x, void* workspace, size_t workspace_bytes, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0 || u < 0 || (long long)m * u > 0x7fffffffLL) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!grad_x || ((long long)m * u > 0 && (!grad_y || !indices))) return PCFM_E_BADARG;
  // rows of up to 229376 points: scatter into shared-memory rows (option "grouping_bwd" = 1: sorted-segment sums, whose
  // result does not depend on the scheduling order — and which any longer row takes)
  if (g_opt_grouping_bwd == 0 && scatter_rows_fits(n) && b <= 65535 && (long long)c * kScatterMaxTiles <= 0x7fffffffLL)
    return scatter_rows_smem(grad_y, indices, nullptr, b, c, m * u, n, m * u, grad_x, (cudaStream_t)stream);
  return scatter_unit(grad_y, indices, b, c, n, m * u, grad_x, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

int pcfm_gather_features_forward(const float* features, const int* indices, int b, int c, int n, int m,
                                 float* out, pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || m == 0) return 0;
  if (!features || !indices || !out || n == 0) return PCFM_E_BADARG;
  return launch_row_gather(1, features, indices, nullptr, out, b, c, n, m, (cudaStream_t)stream);
}

size_t pcfm_gather_features_backward_workspace(int b, int c, int n, int m) {
  (void)c;
  if (b <= 0 || n <= 0) return 256;
  return scatter_ws(b, m, n);
}

int pcfm_gather_features_backward(const float* grad_y, const int* indices, int b, int c, int n, int m,
                                  float* grad_x, void* workspace, size_t workspace_bytes,
                                  pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!grad_x || (m > 0 && (!grad_y || !indices))) return PCFM_E_BADARG;
  if (g_opt_gather_bwd == 0) {
    // sampling.cu:62-84: zero + one atomicAdd per (channel, sample).  Sampled indices are normally distinct (FPS), so
    // a target receives one addend and the result does not depend on the order; a plan build (the sorted-segment path
    // below) costs 3-5x the whole operator at these sizes (B*C*M ~ 10^6).
    cudaStream_t st = (cudaStream_t)stream;
    PCFM_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)b * c * n * sizeof(float), st));
    if (m == 0) return 0;
    dim3 grid(cdiv(m, 256), c, b);
    gather_bwd_atomic_kernel<<<grid, 256, 0, st>>>(grad_y, indices, c, n, m, grad_x);
    PCFM_LAUNCHED();
    return 0;
  }
  return scatter_unit(grad_y, indices, b, c, n, m, grad_x, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// 3-NN interpolate backward
// ---------------------------------------------------------------------------------------------
size_t pcfm_three_nn_interpolate_backward_workspace(int b, int c, int n, int m) {
  (void)c;
  if (b <= 0 || m <= 0) return 256;
  return csr_workspace_bytes(b, 3LL * n, m, true) + align256((size_t)b * m * sizeof(int)) + 256;
}

int pcfm_three_nn_interpolate_backward(const float* grad_y, const int* idx, const float* wgt, int b, int c,
                                       int n, int m, float* grad_x, void* workspace, size_t workspace_bytes,
                                       pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0 || 3LL * n > 0x7fffffffLL) return PCFM_E_BADARG;
  if (b == 0 || c == 0 || m == 0) return 0;
  if (!grad_x || (n > 0 && (!grad_y || !idx || !wgt))) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  // the m centres of a row fit in shared memory: scatter (option "three_nn_bwd" = 1: sorted-segment sums)
  if (g_opt_three_nn_bwd == 0 && n > 0 && scatter_rows_fits(m) && b <= 65535 && (long long)c * kScatterMaxTiles <= 0x7fffffffLL)
    return scatter_rows_smem(grad_y, idx, wgt, b, c, n, m, 3 * n, grad_x, st);
  Workspace ws(workspace, workspace_bytes);
  int* cnt = ws.take<int>((size_t)b * m);
  if (!ws.ok()) return PCFM_E_WORKSPACE;
  CsrPlan plan;
  int rc = csr_build(idx, nullptr, 0, nullptr, wgt, b, 3 * n, m, n > 0 ? n : 1, cnt, &plan, &ws, st);
  if (rc) return rc;
  return launch_row_segsum(SEG_WEIGHTED, grad_y, plan, grad_x, b, c, n, m, 3 * n, st);
}

}  // extern "C"
