// cellplan.cu — builds the sliced-ELL cell plan (count -> classify -> layout -> fill -> order).
#include "cellplan.cuh"

namespace pcfm {

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

// counters per batch element: ctr[c] = #light cells with c points (c in 1..light_max), ctr[kLightCap+1] = #heavy
// cells, ctr[kLightCap+2] = #points in heavy cells (allocator of their contiguous record ranges)
constexpr int kCtr = kLightCap + 4;
constexpr int kEllPad = 32 * (kLightCap * (kLightCap + 1) / 2) + 1024;  // every count bin wastes < 1 slice

// which of the 8 corners (dx,dy,dz) of cell v stay inside the grid (trilinear plans; mean plans do not use it)
__device__ __forceinline__ int corner_mask(int v, int r, bool tri) {
  if (!tri) return 0xff;
  const int r2 = r * r;
  const int x = v / r2, y = (v / r) % r, z = v % r;
  int m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (x + (k >> 2) < r && y + ((k >> 1) & 1) < r && z + (k & 1) < r) m |= 1 << k;
  return m;
}

// one thread per cell: position inside its count bin (light) or inside the heavy list + a record range
__global__ void __launch_bounds__(256) cp_classify_kernel(const int* __restrict__ cnt, int t, int lm,
                                                          int* __restrict__ ctr, int* __restrict__ aux,
                                                          int* __restrict__ recbase) {
  const int b = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0) return;
  int* c = ctr + b * kCtr;
  if (n <= lm) {
    aux[(long long)b * t + v] = atomicAdd(c + n, 1);
  } else {
    aux[(long long)b * t + v] = atomicAdd(c + kLightCap + 1, 1);
    recbase[(long long)b * t + v] = atomicAdd(c + kLightCap + 2, n);  // stride-1 record range
  }
}

// one thread per cell: sorted index of a light cell, its slice and the slice's record base; heavy list entry
__global__ void __launch_bounds__(256) cp_layout_kernel(const int* __restrict__ cnt, int t, int r, int tri, int lm,
                                                        const int* __restrict__ ctr, const int* __restrict__ aux,
                                                        int* __restrict__ recbase, int* __restrict__ cells,
                                                        int* __restrict__ slot0, int* __restrict__ meta,
                                                        int* __restrict__ heavy, int lcap, int hcap) {
  // bins in processing order: bin j holds the cells with lm - j points (count descending)
  __shared__ int P[kLightCap + 1];   // sorted index where bin j starts
  __shared__ int F[kLightCap + 1];   // first slice that STARTS in bin j
  __shared__ int Bs[kLightCap + 1];  // record offset of slice F[j]
  __shared__ int c[kCtr];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kCtr; i += blockDim.x) c[i] = ctr[b * kCtr + i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int p = 0, base = 0;
    for (int k = lm; k >= 1; --k) {
      const int j = lm - k;
      P[j] = p;
      F[j] = (p + 31) >> 5;
      Bs[j] = base;
      p += c[k];
      const int pn = (k == 1) ? ((p + 31) & ~31) : p;  // where the next bin starts
      base += 32 * k * (((pn + 31) >> 5) - F[j]);      // slices that start in bin j are k records wide
    }
    P[lm] = (p + 31) & ~31;
    F[lm] = P[lm] >> 5;
    if (blockIdx.x == 0) {
      meta[b * kMetaInts + 0] = min(P[lm], lcap);
      meta[b * kMetaInts + 1] = min(c[kLightCap + 1], hcap);
    }
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0) return;
  const int mask = corner_mask(v, r, tri != 0);
  if (n > lm) {
    const int slot = aux[(long long)b * t + v];
    if (slot < hcap)
      ((int4*)heavy)[(long long)b * hcap + slot] = make_int4(v, recbase[(long long)b * t + v], n, mask);
    return;
  }
  int j = lm - n;
  const int si = P[j] + aux[(long long)b * t + v];
  const int s = si >> 5, lane = si & 31;
  while (s < F[j]) --j;  // the bin in which slice s starts
  const int width = lm - j;
  const int s0 = c[kLightCap + 2] + Bs[j] + 32 * width * (s - F[j]) + lane;  // the ELL area follows the heavy records
  recbase[(long long)b * t + v] = s0 | (int)0x80000000;                     // stride 32
  if (si < lcap) {
    cells[(long long)b * lcap + si] = (n << 24) | v;
    slot0[(long long)b * lcap + si] = (mask << 24) | s0;
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
                                                       const int* __restrict__ meta, void* __restrict__ rec,
                                                       int lcap, int rcap) {
  const int b = blockIdx.y;
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= meta[b * kMetaInts + 0]) return;
  const int n = (int)((unsigned)cells[(long long)b * lcap + ci] >> 24);
  if (n < 2 || n > kSortMax) return;
  const long long s0 = (long long)b * rcap + (slot0[(long long)b * lcap + ci] & 0xffffff);
  if (TRI) {
    float4* rr = (float4*)rec;
    float4 loc[kSortMax];
#pragma unroll
    for (int a = 0; a < kSortMax; ++a)
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
    int loc[kSortMax];
#pragma unroll
    for (int a = 0; a < kSortMax; ++a)
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

static long long lcap_of(long long n, long long t) { return (((n < t ? n : t) + 31) & ~31LL) + 32; }

size_t cellplan_workspace_bytes(int b, long long n, long long t, bool trilinear) {
  size_t s = 0;
  const long long lcap = lcap_of(n, t), hcap = n / (kLightMaxMean + 1) + 1, rcap = n + kEllPad;
  s += align256((size_t)b * lcap * 4) * 2 + align256((size_t)b * kCtr * 4);  // cells (+counters), slot0
  s += align256((size_t)b * kMetaInts * 4);                    // meta
  s += align256((size_t)b * hcap * 4 * 4);                     // heavy
  s += align256((size_t)b * rcap * (trilinear ? 16 : 4));      // rec
  s += align256((size_t)b * n * 4);                            // rank
  s += align256((size_t)b * t * 4) * 2;                        // aux, recbase
  if (trilinear) s += align256((size_t)b * t * 4);             // cnt (caller has none)
  return s + 256;
}

int cellplan_build(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                   int* cnt, CellPlan* plan, Workspace* ws, cudaStream_t st) {
  const bool tri = float_coords != nullptr;
  plan->lcap = (int)lcap_of(n, t);
  plan->light_max = tri ? kLightMaxTri : kLightMaxMean;
  plan->hcap = n / (plan->light_max + 1) + 1;
  plan->rcap = n + kEllPad;
  // cells and the counters are zeroed by ONE memset: they are carved from one block
  const size_t zero_ints = (size_t)b * plan->lcap + (size_t)b * kCtr;
  int* zero_block = ws->take<int>(zero_ints);
  plan->cells = zero_block;
  int* ctr = zero_block ? zero_block + (size_t)b * plan->lcap : nullptr;
  plan->slot0 = ws->take<int>((size_t)b * plan->lcap);
  plan->meta = ws->take<int>((size_t)b * kMetaInts);
  plan->heavy = (int*)ws->take<int4>((size_t)b * plan->hcap);
  plan->rec = tri ? (void*)ws->take<float4>((size_t)b * plan->rcap) : (void*)ws->take<int>((size_t)b * plan->rcap);
  int* rank = ws->take<int>((size_t)b * n);
  int* aux = ws->take<int>((size_t)b * t);
  int* recbase = ws->take<int>((size_t)b * t);
  if (!cnt) cnt = ws->take<int>((size_t)b * t);
  if (!ws->ok()) return PCFM_E_WORKSPACE;
  PCFM_CUDA(cudaMemsetAsync(cnt, 0, (size_t)b * t * sizeof(int), st));
  PCFM_CUDA(cudaMemsetAsync(zero_block, 0, zero_ints * sizeof(int), st));  // counters; holes in `cells` = empty
  dim3 gp(cdiv(n > 0 ? n : 1, 256), b), gt(cdiv(t, 256), b);
  if (n > 0) {
    if (tri) cp_count_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, nullptr, n, t, cnt, rank);
    else cp_count_kernel<false><<<gp, 256, 0, st>>>(vox_coords, nullptr, r, ind_out, n, t, cnt, rank);
    PCFM_LAUNCHED();
  }
  cp_classify_kernel<<<gt, 256, 0, st>>>(cnt, t, plan->light_max, ctr, aux, recbase);
  PCFM_LAUNCHED();
  cp_layout_kernel<<<gt, 256, 0, st>>>(cnt, t, r, tri ? 1 : 0, plan->light_max, ctr, aux, recbase, plan->cells, plan->slot0, plan->meta,
                                       plan->heavy, plan->lcap, plan->hcap);
  PCFM_LAUNCHED();
  if (n > 0) {
    if (tri) {
      // no ordering pass for trilinear plans: their consumer adds cell sums with shared-memory atomics, whose order
      // is not fixed anyway
      cp_fill_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, n, t, rank, recbase, plan->rec, plan->rcap);
      PCFM_LAUNCHED();
    } else {
      dim3 gl(cdiv(plan->lcap, 256), b);
      cp_fill_kernel<false><<<gp, 256, 0, st>>>(ind_out, nullptr, r, n, t, rank, recbase, plan->rec, plan->rcap);
      PCFM_LAUNCHED();
      cp_order_kernel<false><<<gl, 256, 0, st>>>(plan->cells, plan->slot0, plan->meta, plan->rec, plan->lcap,
                                                 plan->rcap);
      PCFM_LAUNCHED();
    }
  }
  return 0;
}

}  // namespace pcfm
