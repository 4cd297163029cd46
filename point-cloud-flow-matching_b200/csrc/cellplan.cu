// cellplan.cu — builds the sliced-ELL cell plan (count -> classify -> layout -> fill -> order).
#include "cellplan.cuh"

namespace pcfm {

constexpr int kEllPad = 4608;  // >= 32 * (1+2+...+16): every count class can waste at most one slice of padding

__device__ __forceinline__ int floor_cell(const float* cb, int i, int n, int r, float* fx, float* fy, float* fz) {
  // identical arithmetic to devox_fwd_kernel (devox.cu): clamp (no-op for contractual inputs), floor, fraction
  const float hi = (float)(r - 1);
  const float x = fminf(fmaxf(cb[i], 0.f), hi);
  const float y = fminf(fmaxf(cb[(long long)i + n], 0.f), hi);
  const float z = fminf(fmaxf(cb[(long long)i + 2LL * n], 0.f), hi);
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  *fx = x - xl;
  *fy = y - yl;
  *fz = z - zl;
  return (int)xl * r * r + (int)yl * r + (int)zl;
}

// one thread per point: cell key, per-cell count (warp-aggregated atomics), arbitrary unique rank inside the cell
template <bool TRI>
__global__ void __launch_bounds__(256) cp_count_kernel(const int* __restrict__ vox_coords,
                                                       const float* __restrict__ fcoords, int r,
                                                       int* __restrict__ ind_out, int n, int t,
                                                       int* __restrict__ cnt, int* __restrict__ rank) {
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inb = i < n;
  int key = -1;
  if (inb) {
    if (TRI) {
      float fx, fy, fz;
      key = floor_cell(fcoords + base * 3, i, n, r, &fx, &fy, &fz);
    } else {
      const int* cb = vox_coords + base * 3;
      key = cb[i] * (r * r) + cb[(long long)i + n] * r + cb[(long long)i + 2LL * n];  // vox.cu:31
      ind_out[base + i] = key;
    }
  }
  const bool valid = inb && key >= 0 && key < t;
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  int myrank = -1;
  if (valid) {
    const unsigned peers = __match_any_sync(act, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    int old = 0;
    if (lane == leader) old = atomicAdd(cnt + (long long)b * t + key, __popc(peers));
    old = __shfl_sync(peers, old, leader);
    myrank = old + __popc(peers & ((1u << lane) - 1u));
  }
  if (inb) rank[base + i] = myrank;
}

// counters per batch element: hist[0..16] (hist[c] = #cells with c points, c in 1..16), [17] = #heavy cells,
// [18] = #points in heavy cells
constexpr int kCtr = 20;

// one thread per cell: position inside its count class (light) or heavy slot + record range (heavy)
__global__ void __launch_bounds__(256) cp_classify_kernel(const int* __restrict__ cnt, int t,
                                                          int* __restrict__ ctr, int* __restrict__ aux,
                                                          int* __restrict__ recbase, int* __restrict__ heavy,
                                                          int hcap) {
  const int b = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0) return;
  int* c = ctr + b * kCtr;
  if (n <= kLightMax) {
    aux[(long long)b * t + v] = atomicAdd(c + n, 1);
  } else {
    const int slot = atomicAdd(c + 17, 1);
    const int hs = atomicAdd(c + 18, n);
    if (slot < hcap) {
      int* h = heavy + ((long long)b * hcap + slot) * 3;
      h[0] = v; h[1] = hs; h[2] = n;
    }
    recbase[(long long)b * t + v] = hs;  // stride 1
  }
}

// one thread per cell: sorted index of a light cell (count-descending), its slice, and the slice's record base
__global__ void __launch_bounds__(256) cp_layout_kernel(const int* __restrict__ cnt, int t,
                                                        const int* __restrict__ ctr, const int* __restrict__ aux,
                                                        int* __restrict__ recbase, int* __restrict__ cells,
                                                        int* __restrict__ slot0, int* __restrict__ nlight,
                                                        int* __restrict__ nheavy, int lcap, int hcap) {
  __shared__ int P[kLightMax + 2];      // P[c] = #light cells with more than c points (sorted index where class c starts)
  __shared__ int F[kLightMax + 2];      // F[c] = first slice that STARTS in class c = ceil(P[c]/32)
  __shared__ int Bs[kLightMax + 2];     // Bs[c] = record offset of slice F[c]
  const int b = blockIdx.y;
  const int* c = ctr + b * kCtr;
  if (threadIdx.x == 0) {
    int p = 0;
    for (int k = kLightMax; k >= 1; --k) { P[k] = p; p += c[k]; }
    P[0] = p;  // total light cells
    int base = 0;
    for (int k = kLightMax; k >= 1; --k) {
      F[k] = (P[k] + 31) >> 5;
      Bs[k] = base;
      const int fnext = (P[k - 1] + 31) >> 5;
      base += 32 * k * (fnext - F[k]);
    }
    F[0] = (P[0] + 31) >> 5;
    if (blockIdx.x == 0) {
      nlight[b] = min(P[0], lcap);
      nheavy[b] = min(c[17], hcap);
    }
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0 || n > kLightMax) return;
  const int si = P[n] + aux[(long long)b * t + v];
  const int s = si >> 5, lane = si & 31;
  int k = kLightMax;  // class in which slice s starts: F[k] <= s < F[k-1]
  while (k > 1 && s >= F[k - 1]) --k;
  const int s0 = c[18] + Bs[k] + 32 * k * (s - F[k]) + lane;  // ELL area follows the heavy records
  recbase[(long long)b * t + v] = s0 | (int)0x80000000;       // stride 32
  if (si < lcap) {
    cells[(long long)b * lcap + si] = (n << 24) | v;
    slot0[(long long)b * lcap + si] = s0;
  }
}

// one thread per point: write its record into the slot (record base of its cell) + rank * stride
template <bool TRI>
__global__ void __launch_bounds__(256) cp_fill_kernel(const int* __restrict__ keys,
                                                      const float* __restrict__ fcoords, int r, int n, int t,
                                                      const int* __restrict__ rank,
                                                      const int* __restrict__ recbase, void* __restrict__ rec,
                                                      int rcap) {
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rk = rank[base + i];
  if (rk < 0) return;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  int key;
  if (TRI) key = floor_cell(fcoords + base * 3, i, n, r, &fx, &fy, &fz);
  else key = keys[base + i];
  const int rb = recbase[(long long)b * t + key];
  const int slot = (rb & 0x7fffffff) + rk * (rb < 0 ? 32 : 1);
  if (slot >= rcap) return;
  if (TRI) {
    float4 v = make_float4(__int_as_float(i), fx, fy, fz);
    ((float4*)rec)[(long long)b * rcap + slot] = v;
  } else {
    ((int*)rec)[(long long)b * rcap + slot] = i;
  }
}

// one thread per light cell: ascending point order inside the cell (deterministic, and the order the oracle uses)
template <bool TRI>
__global__ void __launch_bounds__(256) cp_order_kernel(const int* __restrict__ cells,
                                                       const int* __restrict__ slot0,
                                                       const int* __restrict__ nlight, void* __restrict__ rec,
                                                       int lcap, int rcap) {
  const int b = blockIdx.y;
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= nlight[b]) return;
  const int n = cells[(long long)b * lcap + ci] >> 24;
  if (n < 2) return;
  const long long s0 = (long long)b * rcap + slot0[(long long)b * lcap + ci];
  if (TRI) {
    float4* rr = (float4*)rec;
    float4 loc[kLightMax];
#pragma unroll
    for (int a = 0; a < kLightMax; ++a)
      if (a < n) loc[a] = rr[s0 + 32 * a];
    for (int a = 1; a < n; ++a) {
      const float4 x = loc[a];
      int j = a - 1;
      while (j >= 0 && __float_as_int(loc[j].x) > __float_as_int(x.x)) { loc[j + 1] = loc[j]; --j; }
      loc[j + 1] = x;
    }
    for (int a = 0; a < n; ++a) rr[s0 + 32 * a] = loc[a];
  } else {
    int* rr = (int*)rec;
    int loc[kLightMax];
#pragma unroll
    for (int a = 0; a < kLightMax; ++a)
      if (a < n) loc[a] = rr[s0 + 32 * a];
    for (int a = 1; a < n; ++a) {
      const int x = loc[a];
      int j = a - 1;
      while (j >= 0 && loc[j] > x) { loc[j + 1] = loc[j]; --j; }
      loc[j + 1] = x;
    }
    for (int a = 0; a < n; ++a) rr[s0 + 32 * a] = loc[a];
  }
}

static long long lcap_of(long long n, long long t) { return (n < t ? n : t) + 1; }

size_t cellplan_workspace_bytes(int b, long long n, long long t, bool trilinear) {
  size_t s = 0;
  const long long lcap = lcap_of(n, t), hcap = n / (kLightMax + 1) + 1, rcap = n + kEllPad;
  s += align256((size_t)b * lcap * 4) * 2;                     // cells, slot0
  s += align256((size_t)b * 4) * 2;                            // nlight, nheavy
  s += align256((size_t)b * hcap * 3 * 4);                     // heavy
  s += align256((size_t)b * rcap * (trilinear ? 16 : 4));      // rec
  s += align256((size_t)b * n * 4);                            // rank
  s += align256((size_t)b * t * 4) * 2;                        // aux, recbase
  s += align256((size_t)b * kCtr * 4);                         // counters
  if (trilinear) s += align256((size_t)b * t * 4);             // cnt (caller has none)
  return s + 256;
}

int cellplan_build(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                   int* cnt, CellPlan* plan, Workspace* ws, cudaStream_t st) {
  const bool tri = float_coords != nullptr;
  plan->lcap = (int)lcap_of(n, t);
  plan->hcap = n / (kLightMax + 1) + 1;
  plan->rcap = n + kEllPad;
  plan->cells = ws->take<int>((size_t)b * plan->lcap);
  plan->slot0 = ws->take<int>((size_t)b * plan->lcap);
  plan->nlight = ws->take<int>(b);
  plan->nheavy = ws->take<int>(b);
  plan->heavy = ws->take<int>((size_t)b * plan->hcap * 3);
  plan->rec = tri ? (void*)ws->take<float4>((size_t)b * plan->rcap) : (void*)ws->take<int>((size_t)b * plan->rcap);
  int* rank = ws->take<int>((size_t)b * n);
  int* aux = ws->take<int>((size_t)b * t);
  int* recbase = ws->take<int>((size_t)b * t);
  int* ctr = ws->take<int>((size_t)b * kCtr);
  if (!cnt) cnt = ws->take<int>((size_t)b * t);
  if (!ws->ok()) return PCFM_E_WORKSPACE;
  PCFM_CUDA(cudaMemsetAsync(cnt, 0, (size_t)b * t * sizeof(int), st));
  PCFM_CUDA(cudaMemsetAsync(ctr, 0, (size_t)b * kCtr * sizeof(int), st));
  dim3 gp(cdiv(n > 0 ? n : 1, 256), b), gt(cdiv(t, 256), b);
  if (n > 0) {
    if (tri) cp_count_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, nullptr, n, t, cnt, rank);
    else cp_count_kernel<false><<<gp, 256, 0, st>>>(vox_coords, nullptr, r, ind_out, n, t, cnt, rank);
    PCFM_LAUNCHED();
  }
  cp_classify_kernel<<<gt, 256, 0, st>>>(cnt, t, ctr, aux, recbase, plan->heavy, plan->hcap);
  PCFM_LAUNCHED();
  cp_layout_kernel<<<gt, 256, 0, st>>>(cnt, t, ctr, aux, recbase, plan->cells, plan->slot0, plan->nlight,
                                       plan->nheavy, plan->lcap, plan->hcap);
  PCFM_LAUNCHED();
  if (n > 0) {
    dim3 gl(cdiv(plan->lcap, 256), b);
    if (tri) {
      cp_fill_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, n, t, rank, recbase, plan->rec, plan->rcap);
      PCFM_LAUNCHED();
      cp_order_kernel<true><<<gl, 256, 0, st>>>(plan->cells, plan->slot0, plan->nlight, plan->rec, plan->lcap,
                                                plan->rcap);
    } else {
      cp_fill_kernel<false><<<gp, 256, 0, st>>>(ind_out, nullptr, r, n, t, rank, recbase, plan->rec, plan->rcap);
      PCFM_LAUNCHED();
      cp_order_kernel<false><<<gl, 256, 0, st>>>(plan->cells, plan->slot0, plan->nlight, plan->rec, plan->lcap,
                                                 plan->rcap);
    }
    PCFM_LAUNCHED();
  }
  return 0;
}

}  // namespace pcfm
