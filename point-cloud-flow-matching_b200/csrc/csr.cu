// csr.cu — builds the sorted-segment plan (count -> scan -> fill -> finalize).
#include "csr.cuh"

namespace pcfm {

static_assert(kHeavySeg >= 1, "");
constexpr int kSmallSeg = 16;      // sorted by one thread in local memory
constexpr int kMaxSortedSeg = 64;  // ranked by a warp (O(n^2/32)); longer segments keep their arrival order

// One thread per edge: key, per-target count (warp-aggregated atomics: equal keys inside a warp issue ONE
// atomicAdd), and a unique arbitrary rank inside the target's segment.
template <bool VOX>
__global__ void __launch_bounds__(256) csr_count_kernel(const int* __restrict__ tgt,
                                                        const int* __restrict__ coords, int r,
                                                        int* __restrict__ ind_out, int e, int t,
                                                        int* __restrict__ cnt, int* __restrict__ rank) {
  const int b = blockIdx.y;
  const long long base = (long long)b * e;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inb = i < e;
  int key = -1;
  if (inb) {
    if (VOX) {
      const int* cb = coords + base * 3;
      // vox.cu:31  ind = x*r2 + y*r + z
      key = cb[i] * (r * r) + cb[(long long)i + e] * r + cb[(long long)i + 2LL * e];
      ind_out[base + i] = key;
    } else {
      key = tgt[base + i];
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

// One CTA (1024 threads) per batch element: exclusive scan of cnt[b,0..t) into start[b,0..t].
__global__ void __launch_bounds__(1024) csr_scan_kernel(const int* __restrict__ cnt, int t,
                                                        int* __restrict__ start) {
  __shared__ int warp_tot[32];
  __shared__ int tile_total;
  const int b = blockIdx.x;
  cnt += (long long)b * t;
  start += (long long)b * (t + 1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < t; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < t) ? cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      const int w = warp_tot[lane];
      int xs = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, xs, o);
        if (lane >= o) xs += y;
      }
      warp_tot[lane] = xs - w;
      if (lane == 31) tile_total = xs;
    }
    __syncthreads();
    if (i < t) start[i] = carry + warp_tot[wid] + x - v;
    carry += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[t] = carry;
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int* __restrict__ keys,
                                                       const int* __restrict__ rank,
                                                       const int* __restrict__ start, int e, int t,
                                                       int* __restrict__ tmp) {
  const int b = blockIdx.y;
  const long long base = (long long)b * e;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  const int rk = rank[base + i];
  if (rk < 0) return;
  const int key = keys[base + i];
  tmp[base + start[(long long)b * (t + 1) + key] + rk] = i;
}

// One thread per target: put the segment's edge ids in ascending order and emit (source, weight) in that
// order.  Long segments are ranked cooperatively by the warp.
__global__ void __launch_bounds__(256) csr_finalize_kernel(const int* __restrict__ start,
                                                           const int* __restrict__ tmp,
                                                           const float* __restrict__ w, int e, int t,
                                                           int src_mod, int* __restrict__ ssrc,
                                                           float* __restrict__ sw, int* __restrict__ heavy,
                                                           int* __restrict__ nheavy, int hcap) {
  const int b = blockIdx.y;
  const long long base = (long long)b * e;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int* st = start + (long long)b * (t + 1);
  const int* tb = tmp + base;
  int s = 0, n = 0;
  if (v < t) {
    s = st[v];
    n = st[v + 1] - s;
  }
  if (n > kHeavySeg) {  // order of this list is irrelevant: every entry is reduced independently
    const int slot = atomicAdd(nheavy + b, 1);
    if (slot < hcap) heavy[(long long)b * hcap + slot] = v;
  }
  if (n >= 1 && n <= kSmallSeg) {
    int loc[kSmallSeg];
#pragma unroll
    for (int a = 0; a < kSmallSeg; ++a) loc[a] = (a < n) ? tb[s + a] : 0x7fffffff;
    for (int a = 1; a < n; ++a) {
      const int x = loc[a];
      int j = a - 1;
      while (j >= 0 && loc[j] > x) {
        loc[j + 1] = loc[j];
        --j;
      }
      loc[j + 1] = x;
    }
    for (int a = 0; a < n; ++a) {
      const int ee = loc[a];
      ssrc[base + s + a] = ee % src_mod;
      if (sw) sw[base + s + a] = w[base + ee];
    }
  }
  unsigned big = __ballot_sync(0xffffffffu, n > kSmallSeg);
  while (big) {
    const int src_lane = __ffs(big) - 1;
    big &= big - 1;
    const int S = __shfl_sync(0xffffffffu, s, src_lane);
    const int N = __shfl_sync(0xffffffffu, n, src_lane);
    const bool sorted = N <= kMaxSortedSeg;
    for (int a = lane; a < N; a += 32) {
      const int ea = tb[S + a];
      int rk = a;
      if (sorted) {
        rk = 0;
        for (int f = 0; f < N; ++f) rk += (tb[S + f] < ea) ? 1 : 0;
      }
      ssrc[base + S + rk] = ea % src_mod;
      if (sw) sw[base + S + rk] = w[base + ea];
    }
  }
}

size_t csr_workspace_bytes(int b, long long e, long long t, bool weights) {
  size_t n = 0;
  n += align256((size_t)b * (t + 1) * sizeof(int));  // start
  n += align256((size_t)b * e * sizeof(int));        // rank
  n += align256((size_t)b * e * sizeof(int));        // tmp
  n += align256((size_t)b * e * sizeof(int));        // ssrc
  if (weights) n += align256((size_t)b * e * sizeof(float));
  const long long hcap = e / (kHeavySeg + 1) + 1;
  n += align256((size_t)b * hcap * sizeof(int)) + align256((size_t)b * sizeof(int));  // heavy list + counts
  return n;
}

int csr_build(const int* tgt, const int* coords, int r, int* ind_out, const float* w, int b, int e, int t,
              int src_mod, int* cnt, CsrPlan* plan, Workspace* ws, cudaStream_t st) {
  plan->start = ws->take<int>((size_t)b * (t + 1));
  int* rank = ws->take<int>((size_t)b * e);
  int* tmp = ws->take<int>((size_t)b * e);
  plan->ssrc = ws->take<int>((size_t)b * e);
  plan->sw = w ? ws->take<float>((size_t)b * e) : nullptr;
  plan->hcap = e / (kHeavySeg + 1) + 1;  // at most this many segments can be longer than kHeavySeg
  plan->heavy = ws->take<int>((size_t)b * plan->hcap);
  plan->nheavy = ws->take<int>((size_t)b);
  if (!ws->ok()) return PCFM_E_WORKSPACE;
  PCFM_CUDA(cudaMemsetAsync(plan->nheavy, 0, (size_t)b * sizeof(int), st));
  PCFM_CUDA(cudaMemsetAsync(cnt, 0, (size_t)b * t * sizeof(int), st));
  if (e > 0) {
    dim3 ge(cdiv(e, 256), b);
    if (coords)
      csr_count_kernel<true><<<ge, 256, 0, st>>>(nullptr, coords, r, ind_out, e, t, cnt, rank);
    else
      csr_count_kernel<false><<<ge, 256, 0, st>>>(tgt, nullptr, 0, nullptr, e, t, cnt, rank);
    PCFM_LAUNCHED();
  }
  csr_scan_kernel<<<b, 1024, 0, st>>>(cnt, t, plan->start);
  PCFM_LAUNCHED();
  if (e > 0) {
    dim3 ge(cdiv(e, 256), b);
    csr_fill_kernel<<<ge, 256, 0, st>>>(coords ? ind_out : tgt, rank, plan->start, e, t, tmp);
    PCFM_LAUNCHED();
    dim3 gt(cdiv(t, 256), b);
    csr_finalize_kernel<<<gt, 256, 0, st>>>(plan->start, tmp, w, e, t, src_mod, plan->ssrc, plan->sw, plan->heavy,
                                            plan->nheavy, plan->hcap);
    PCFM_LAUNCHED();
  }
  return 0;
}

}  // namespace pcfm
