// cellplan.cu — builds the sliced-ELL cell plan: count -> classify -> layout -> fill -> order, as one fused kernel
// (one thread-block cluster per cloud, per-cell table in distributed shared memory) or, for very big grids /
// clouds, as separate kernels.
#include "cellplan.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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

// record slots a cell of c points occupies: trilinear plans store PAIRS of points, mean plans pad heavy cells to
// whole groups of 4 ids (cellplan.cuh)
template <bool TRI>
__device__ __forceinline__ int light_units(int c) { return TRI ? (c + 1) >> 1 : c; }
template <bool TRI>
__device__ __forceinline__ int heavy_units(int c) { return TRI ? (c + 1) >> 1 : (c + 3) & ~3; }

// the dummy half that completes the last pair of a cell with an odd number of points (trilinear plans): byte offset
// 4*n (the consumer's zero / scratch slot behind the row) and zero fractions
__device__ __forceinline__ void tri_dummy_half(float4* __restrict__ ra, float4* __restrict__ rb, long long slot, int n) {
  float* a = (float*)(ra + slot);
  float* b = (float*)(rb + slot);
  a[1] = __int_as_float(4 * n);
  a[3] = 0.f;
  b[1] = 0.f;
  b[3] = 0.f;
}
// one point's half of its pair slot
__device__ __forceinline__ void tri_fill_half(float4* __restrict__ ra, float4* __restrict__ rb, long long slot, int h,
                                              int i, float fx, float fy, float fz) {
  float* a = (float*)(ra + slot);
  float* b = (float*)(rb + slot);
  a[h] = __int_as_float(4 * i);
  a[2 + h] = fx;
  b[h] = fy;
  b[2 + h] = fz;
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
constexpr int kEllPadTri = 32 * ((kLightCap / 2) * (kLightCap / 2 + 1) / 2) + 1024;  // bins of 1 .. kLightCap/2 pairs

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
template <bool TRI>
__global__ void __launch_bounds__(256) cp_classify_kernel(const int* __restrict__ cnt, int t, int lm, int chunk,
                                                          int* __restrict__ ctr, int* __restrict__ aux,
                                                          int* __restrict__ recbase) {
  const int b = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0) return;
  int* c = ctr + b * kCtr;
  if (n <= lm) {
    aux[(long long)b * t + v] = atomicAdd(c + light_units<TRI>(n), 1);
  } else {
    aux[(long long)b * t + v] = atomicAdd(c + kLightCap + 1, ((n - 1) / chunk + 1));  // heavy list entries
    recbase[(long long)b * t + v] = atomicAdd(c + kLightCap + 2, heavy_units<TRI>(n));  // stride-1 slot range
  }
}

// one thread per cell: sorted index of a light cell, its slice and the slice's record base; heavy list entry
template <bool TRI>
__global__ void __launch_bounds__(256) cp_layout_kernel(const int* __restrict__ cnt, int t, int r, int lm, int chunk,
                                                        const int* __restrict__ ctr, const int* __restrict__ aux,
                                                        int* __restrict__ recbase, int* __restrict__ cells,
                                                        int* __restrict__ slot0, int* __restrict__ meta,
                                                        int* __restrict__ heavy, void* __restrict__ rec,
                                                        void* __restrict__ rec2, int npts, int lcap, int hcap,
                                                        int rcap) {
  // bins in processing order: bin j holds the cells with lmu - j record units (count descending)
  __shared__ int P[kLightCap + 1];   // sorted index where bin j starts
  __shared__ int F[kLightCap + 1];   // first slice that STARTS in bin j
  __shared__ int Bs[kLightCap + 1];  // record offset of slice F[j]
  __shared__ int c[kCtr];
  const int b = blockIdx.y;
  const int lmu = light_units<TRI>(lm);
  for (int i = threadIdx.x; i < kCtr; i += blockDim.x) c[i] = ctr[b * kCtr + i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int p = 0, base = 0;
    for (int k = lmu; k >= 1; --k) {
      const int j = lmu - k;
      P[j] = p;
      F[j] = (p + 31) >> 5;
      Bs[j] = base;
      p += c[k];
      const int pn = (k == 1) ? ((p + 31) & ~31) : p;  // where the next bin starts
      base += 32 * k * (((pn + 31) >> 5) - F[j]);      // slices that start in bin j are k units wide
    }
    P[lmu] = (p + 31) & ~31;
    F[lmu] = P[lmu] >> 5;
    if (blockIdx.x == 0) {
      meta[b * kMetaInts + 0] = min(P[lmu], lcap);
      meta[b * kMetaInts + 1] = min(c[kLightCap + 1], hcap);
    }
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t) return;
  const int n = cnt[(long long)b * t + v];
  if (n <= 0) return;
  const int mask = corner_mask(v, r, TRI);
  if (n > lm) {
    const int slot = aux[(long long)b * t + v];
    const int rb = recbase[(long long)b * t + v];
    for (int k = 0, o = 0; o < n; ++k, o += chunk)
      if (slot + k < hcap) {
        const int cnt_k = min(chunk, n - o);
        ((int4*)heavy)[(long long)b * hcap + slot + k] =
            TRI ? make_int4(v, rb + (o >> 1), (cnt_k + 1) >> 1, mask) : make_int4(v, rb + o, cnt_k, n);
      }
    if (TRI) {
      if ((n & 1) && rb + (n >> 1) < rcap)
        tri_dummy_half((float4*)rec, (float4*)rec2, (long long)b * rcap + rb + (n >> 1), npts);
    } else {
      for (int k = n; k < ((n + 3) & ~3); ++k)
        if (rb + k < rcap) ((int*)rec)[(long long)b * rcap + rb + k] = npts;
    }
    return;
  }
  const int u = light_units<TRI>(n);
  int j = lmu - u;
  const int si = P[j] + aux[(long long)b * t + v];
  const int s = si >> 5, lane = si & 31;
  while (s < F[j]) --j;  // the bin in which slice s starts
  const int width = lmu - j;
  const int s0 = c[kLightCap + 2] + Bs[j] + 32 * width * (s - F[j]) + lane;  // the ELL area follows the heavy records
  recbase[(long long)b * t + v] = s0 | (int)0x80000000;                     // stride 32
  if (TRI && (n & 1) && s0 + 32 * (n >> 1) < rcap)
    tri_dummy_half((float4*)rec, (float4*)rec2, (long long)b * rcap + s0 + 32 * (n >> 1), npts);
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
                                                      void* __restrict__ rec2, int rcap) {
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
  const int slot = (rb & 0x7fffffff) + (TRI ? rk >> 1 : rk) * (rb < 0 ? 32 : 1);
  if (slot >= rcap) return;
  if (TRI) tri_fill_half((float4*)rec, (float4*)rec2, (long long)b * rcap + slot, rk & 1, i, fx, fy, fz);
  else ((int*)rec)[(long long)b * rcap + slot] = i;
}

// ascending point order inside one light cell of a mean plan (deterministic, and the order the oracle uses)
__device__ __forceinline__ void order_cell(int* __restrict__ rr, long long s0, int n) {
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

// one thread per light cell (mean plans)
__global__ void __launch_bounds__(256) cp_order_kernel(const int* __restrict__ cells,
                                                       const int* __restrict__ slot0,
                                                       const int* __restrict__ meta, int* __restrict__ rec, int lcap,
                                                       int rcap) {
  const int b = blockIdx.y;
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= meta[b * kMetaInts + 0]) return;
  const int n = (int)((unsigned)cells[(long long)b * lcap + ci] >> 24);
  if (n < 2 || n > kSortMax) return;
  order_cell(rec, (long long)b * rcap + (slot0[(long long)b * lcap + ci] & 0xffffff), n);
}

// -------------------------------------------------------------------------------------------------
// Fused build: ONE kernel, one thread-block CLUSTER of kCl CTAs per cloud.  The per-cell table (point counts, later
// record bases) lives in distributed shared memory: cell v belongs to CTA (v >> 5) % kCl, so the hot cells of a
// centred cloud are spread over all CTAs; counting is a DSMEM atomicAdd, the phases are separated by cluster
// barriers instead of kernel boundaries.  Replaces 2 memsets + 4-5 dependent tiny launches (~35 us).
// (A single CTA per cloud was measured slower than the separate kernels: one SM's load bandwidth is not enough.)
// -------------------------------------------------------------------------------------------------
constexpr int kCl = 8;              // CTAs per cluster: the portable maximum, and the measured optimum (2 / 4 / 16 CTAs per
                                    // cloud: +30 / +9 / +7 us per plan at N = 20000)
constexpr int kPlanThreads = 1024;
constexpr int kPlanUnroll = 4;
constexpr int kNB = kLightCap + 4;  // [1..lm] cells with that many points, [kLightCap+1] heavy cells, [kLightCap+2] their points
constexpr int kH = kLightCap + 1, kHP = kLightCap + 2;

__device__ __forceinline__ int cell_owner(int v) { return (v >> 5) % kCl; }
__device__ __forceinline__ int cell_local(int v) { return (((v >> 5) / kCl) << 5) | (v & 31); }

// `b` = cloud of this cluster, `tab` = [tq] this CTA's share of the per-cell table (dynamic shared memory)
template <bool TRI>
__device__ __forceinline__ void cp_cluster_body(const int* __restrict__ vox_coords, const float* __restrict__ fcoords,
                                                int r, int* __restrict__ ind_out, int n, int t, int tq, int lm,
                                                int chunk, int* __restrict__ cnt_out, int* __restrict__ rank,
                                                int* __restrict__ cells, int* __restrict__ slot0,
                                                int* __restrict__ meta, int* __restrict__ heavy,
                                                void* __restrict__ rec, void* __restrict__ rec2, int lcap, int hcap,
                                                int rcap, const int b, int* tab) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int bin[kNB];      // this CTA's histogram of cell sizes (read by the other CTAs in phase 3)
  __shared__ int pos[kNB];      // running positions inside this CTA's part of every bin
  __shared__ int off[kNB];      // where this CTA's part of every bin starts (sum over lower-ranked CTAs)
  __shared__ int tot[kNB];      // cluster-wide histogram
  __shared__ int P[kLightCap + 1], F[kLightCap + 1], Bs[kLightCap + 1];
  __shared__ int p_end;
  const int q = (int)cluster.block_rank();
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lmu = light_units<TRI>(lm);  // light bins are indexed by record units (pairs for trilinear plans)
  const long long base = (long long)b * n;
  const float* fb = TRI ? fcoords + base * 3 : nullptr;
  const int* vb = TRI ? nullptr : vox_coords + base * 3;
  const int r2 = r * r;
  const int ppc = (n + kCl - 1) / kCl;
  const int i_lo = min(n, q * ppc), i_hi = min(n, i_lo + ppc);  // this CTA's points
  for (int i = tid; i < tq; i += nt) tab[i] = 0;
  if (tid < kNB) { bin[tid] = 0; pos[tid] = 0; }
  cluster.sync();
  // ---- 1. count: cell key of every point, its rank inside the cell ----
  // `single`: every thread sees each of its points once (at most kPlanUnroll of them), so key / rank / fractions stay
  // in registers until phase 5 instead of a round trip through global memory
  const bool single = ppc <= nt * kPlanUnroll;
  int key[kPlanUnroll], rk[kPlanUnroll];
  float fx[kPlanUnroll], fy[kPlanUnroll], fz[kPlanUnroll];
  for (int i0 = i_lo; i0 < i_hi; i0 += nt * kPlanUnroll) {
#pragma unroll
    for (int u = 0; u < kPlanUnroll; ++u) {
      const int i = i0 + u * nt + tid;
      key[u] = -1;
      if (i < i_hi) {
        if (TRI) key[u] = floor_cell(fb, i, n, r, &fx[u], &fy[u], &fz[u]);
        else key[u] = vb[i] * r2 + vb[(long long)i + n] * r + vb[(long long)i + 2LL * n];  // vox.cu:31
      }
    }
#pragma unroll
    for (int u = 0; u < kPlanUnroll; ++u) {
      const int i = i0 + u * nt + tid;
      const bool valid = i < i_hi && key[u] >= 0 && key[u] < t;
      rk[u] = -1;
      // lanes of a warp that hit the same cell share one DSMEM atomic (dense clouds: most of a warp at r = 8)
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const unsigned peers = __match_any_sync(act, key[u]);
        const int leader = __ffs(peers) - 1, lane = tid & 31;
        int old = 0;
        if (lane == leader)
          old = atomicAdd(cluster.map_shared_rank(tab, cell_owner(key[u])) + cell_local(key[u]), __popc(peers));
        rk[u] = __shfl_sync(peers, old, leader) + __popc(peers & ((1u << lane) - 1u));
      }
      if (i < i_hi) {
        if (!single) rank[base + i] = rk[u];
        if (!TRI) ind_out[base + i] = key[u];
      }
    }
  }
  cluster.sync();
  // ---- 2. histogram of the sizes of this CTA's cells (+ the per-cell counts, an output of avg_voxelize) ----
  for (int l = tid; l < tq; l += nt) {
    const int v = (((l >> 5) * kCl + q) << 5) | (l & 31);
    if (v >= t) continue;
    const int c = tab[l];
    if (cnt_out) cnt_out[(long long)b * t + v] = c;
    if (c > 0) atomicAdd(bin + (c <= lm ? light_units<TRI>(c) : kH), c <= lm ? 1 : ((c - 1) / chunk + 1));
    if (c > lm) atomicAdd(bin + kHP, heavy_units<TRI>(c));
  }
  cluster.sync();
  // ---- 3. cluster-wide bins, where this CTA's part of each bin starts, where every bin starts ----
  if (tid < kNB) {
    int sum = 0, mine = 0;
#pragma unroll
    for (int k = 0; k < kCl; ++k) {
      const int x = cluster.map_shared_rank(bin, k)[tid];
      if (k == q) mine = sum;
      sum += x;
    }
    tot[tid] = sum;
    off[tid] = mine;
  }
  __syncthreads();
  if (tid == 0) {  // bins in processing order: bin j holds the cells with lmu - j record units
    int p = 0, rb = 0;
    for (int k = lmu; k >= 1; --k) {
      const int j = lmu - k;
      P[j] = p;
      F[j] = (p + 31) >> 5;
      Bs[j] = rb;
      p += tot[k];
      const int pn = (k == 1) ? ((p + 31) & ~31) : p;
      rb += 32 * k * (((pn + 31) >> 5) - F[j]);
    }
    p_end = p;
    P[lmu] = (p + 31) & ~31;
    F[lmu] = P[lmu] >> 5;
    if (q == 0) {
      meta[b * kMetaInts + 0] = min(P[lmu], lcap);
      meta[b * kMetaInts + 1] = min(tot[kH], hcap);
    }
  }
  __syncthreads();
  // ---- 4. place this CTA's cells: light -> (sorted index, slice, record base), heavy -> list entry + records ----
  const int hpts = tot[kHP];
  for (int l = tid; l < tq; l += nt) {
    const int v = (((l >> 5) * kCl + q) << 5) | (l & 31);
    if (v >= t) continue;
    const int c = tab[l];
    if (c > 0 && c <= lm) {
      const int u = light_units<TRI>(c);
      int j = lmu - u;
      const int si = P[j] + off[u] + atomicAdd(pos + u, 1);
      const int sidx = si >> 5;
      while (sidx < F[j]) --j;  // the bin in which this slice starts
      const int s0 = hpts + Bs[j] + 32 * (lmu - j) * (sidx - F[j]) + (si & 31);  // the ELL area follows the heavy records
      tab[l] = s0 | (int)0x80000000;
      if (TRI && (c & 1) && s0 + 32 * (c >> 1) < rcap)
        tri_dummy_half((float4*)rec, (float4*)rec2, (long long)b * rcap + s0 + 32 * (c >> 1), n);
      if (si < lcap) {
        cells[(long long)b * lcap + si] = (c << 24) | v;
        slot0[(long long)b * lcap + si] = (corner_mask(v, r, TRI) << 24) | s0;
      }
    } else if (c > lm) {
      const int slot = off[kH] + atomicAdd(pos + kH, ((c - 1) / chunk + 1));
      const int rb = off[kHP] + atomicAdd(pos + kHP, heavy_units<TRI>(c));
      tab[l] = rb;
      const int mask = corner_mask(v, r, TRI);
      for (int k = 0, o = 0; o < c; ++k, o += chunk)
        if (slot + k < hcap) {
          const int cnt_k = min(chunk, c - o);
          ((int4*)heavy)[(long long)b * hcap + slot + k] =
              TRI ? make_int4(v, rb + (o >> 1), (cnt_k + 1) >> 1, mask) : make_int4(v, rb + o, cnt_k, c);
        }
      if (TRI) {
        if ((c & 1) && rb + (c >> 1) < rcap)
          tri_dummy_half((float4*)rec, (float4*)rec2, (long long)b * rcap + rb + (c >> 1), n);
      } else {
        for (int k = c; k < ((c + 3) & ~3); ++k)
          if (rb + k < rcap) ((int*)rec)[(long long)b * rcap + rb + k] = n;
      }
    }
  }
  if (q == 0)
    for (int si = p_end + tid; si < P[lmu] && si < lcap; si += nt) cells[(long long)b * lcap + si] = 0;  // holes
  cluster.sync();
  // ---- 5. fill the records of this CTA's points ----
  for (int i0 = i_lo; i0 < i_hi; i0 += nt * kPlanUnroll) {
    if (!single) {
#pragma unroll
      for (int u = 0; u < kPlanUnroll; ++u) {
        const int i = i0 + u * nt + tid;
        rk[u] = -1;
        if (i < i_hi) {
          rk[u] = rank[base + i];
          if (TRI) key[u] = floor_cell(fb, i, n, r, &fx[u], &fy[u], &fz[u]);
          else key[u] = ind_out[base + i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPlanUnroll; ++u) {
      const int i = i0 + u * nt + tid;
      if (rk[u] < 0) continue;
      const int rb = cluster.map_shared_rank(tab, cell_owner(key[u]))[cell_local(key[u])];
      const int slot = (rb & 0x7fffffff) + (TRI ? rk[u] >> 1 : rk[u]) * (rb < 0 ? 32 : 1);
      if (slot >= rcap) continue;
      if (TRI) tri_fill_half((float4*)rec, (float4*)rec2, (long long)b * rcap + slot, rk[u] & 1, i, fx[u], fy[u], fz[u]);
      else ((int*)rec)[(long long)b * rcap + slot] = i;
    }
  }
  cluster.sync();  // records complete (global, cluster scope); no CTA leaves while its table may still be read
  if (!TRI) {
    // ---- 6. mean plans: ascending point order inside small cells ----
    const int l1 = min(P[lmu], lcap);
    for (int ci = q * nt + tid; ci < l1; ci += kCl * nt) {
      const int c = (int)((unsigned)cells[(long long)b * lcap + ci] >> 24);
      if (c < 2 || c > kSortMax) continue;
      order_cell((int*)rec, (long long)b * rcap + (slot0[(long long)b * lcap + ci] & 0xffffff), c);
    }
  }
}

template <bool TRI>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kPlanThreads, 1)
    cp_cluster_kernel(const int* __restrict__ vox_coords, const float* __restrict__ fcoords, int r,
                      int* __restrict__ ind_out, int n, int t, int tq, int lm, int chunk,
                      int* __restrict__ cnt_out,
                      int* __restrict__ rank, int* __restrict__ cells, int* __restrict__ slot0,
                      int* __restrict__ meta, int* __restrict__ heavy, void* __restrict__ rec,
                      void* __restrict__ rec2, int lcap, int hcap, int rcap) {
  griddep_wait();               // chained PDL: this kernel may itself have started before its predecessor finished
  griddep_launch_dependents();  // the consumer kernel may start staging its rows now; it waits before reading the plan
  extern __shared__ int tab[];
  cp_cluster_body<TRI>(vox_coords, fcoords, r, ind_out, n, t, tq, lm, chunk, cnt_out, rank, cells, slot0, meta, heavy, rec,
                       rec2, lcap, hcap, rcap, blockIdx.x / kCl, tab);
}

// Both plans of a PVConv stage in ONE launch: clusters [0, B) build the mean plan (voxel coordinates), clusters
// [B, 2B) the trilinear plan (float coordinates).  One build is latency-bound on 8 CTAs per cloud (5-6 cluster
// barriers) and leaves most SMs idle; the second one runs beside it for free.
struct PlanOut {
  int *rank, *cells, *slot0, *meta, *heavy;
  void *rec, *rec2;
  int lcap, hcap, rcap, lm, chunk;
};
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kPlanThreads, 1)
    cp_cluster_dual_kernel(const int* __restrict__ vox_coords, const float* __restrict__ fcoords, int r,
                           int* __restrict__ ind_out, int* __restrict__ cnt_out, int n, int t, int tq, int nb,
                           PlanOut pm, PlanOut pt) {
  griddep_wait();
  griddep_launch_dependents();
  extern __shared__ int tab[];
  const int cl = blockIdx.x / kCl;  // uniform over the cluster
  if (cl < nb)
    cp_cluster_body<false>(vox_coords, nullptr, r, ind_out, n, t, tq, pm.lm, pm.chunk, cnt_out, pm.rank, pm.cells,
                           pm.slot0, pm.meta, pm.heavy, pm.rec, pm.rec2, pm.lcap, pm.hcap, pm.rcap, cl, tab);
  else
    cp_cluster_body<true>(nullptr, fcoords, r, nullptr, n, t, tq, pt.lm, pt.chunk, nullptr, pt.rank, pt.cells, pt.slot0,
                          pt.meta, pt.heavy, pt.rec, pt.rec2, pt.lcap, pt.hcap, pt.rcap, cl - nb, tab);
}

constexpr size_t kFusedMaxTable = 220 * 1024;  // per CTA; + ~2 KB static shared memory <= 227 KB
constexpr int kFusedMaxPoints = 1 << 18;       // beyond this 8 CTAs per cloud are slower than the multi-kernel build

static long long lcap_of(long long n, long long t) { return (((n < t ? n : t) + 31) & ~31LL) + 32; }
// record slots per cloud.  Trilinear: pair slots — at most one per point (every cell a singleton) + the padding of the
// ELL slices.  Mean: ids + up to 3 pad ids per heavy cell (> kLightMaxMean points each) + the ELL padding.
static long long rcap_of(long long n, bool tri) {
  return ((tri ? n + kEllPadTri : n + 3 * (n / (kLightMaxMean + 1) + 1) + kEllPad) + 3) & ~3LL;
}

bool cellplan_fits(long long n, long long t, bool trilinear) {
  return t <= (1LL << 24) && rcap_of(n, trilinear) < (1LL << 24) && 4 * n < 0x7fffffffLL;
}

size_t cellplan_workspace_bytes(int b, long long n, long long t, bool trilinear) {
  size_t s = 0;
  const int lmx = trilinear ? kLightMaxTri : kLightMaxMean, chk = trilinear ? kHeavyChunk : kHeavyChunkMean;
  const long long lcap = lcap_of(n, t), hcap = n / (lmx + 1) + n / chk + 2;
  const long long rcap = rcap_of(n, trilinear);
  s += align256((size_t)b * lcap * 4) * 2 + align256((size_t)b * kCtr * 4);  // cells (+counters), slot0
  s += align256((size_t)b * kMetaInts * 4);                    // meta
  s += align256((size_t)b * hcap * 4 * 4);                     // heavy
  s += align256((size_t)b * rcap * (trilinear ? 16 : 4)) * (trilinear ? 2 : 1);  // rec (+ rec2)
  s += align256((size_t)b * n * 4);                            // rank
  s += align256((size_t)b * t * 4) * 2;                        // aux, recbase
  if (trilinear) s += align256((size_t)b * t * 4);             // cnt (caller has none)
  return s + 256;
}

// Carves the plan arrays (and the build scratch) out of the workspace.  The carve is a pure function of
// (trilinear, b, n, t), so a workspace left behind by cellplan_build() can be re-attached later (cellplan_attach).
struct PlanScratch {
  int *zero_block, *ctr, *rank, *aux, *recbase;
  size_t zero_ints;
};
static bool cellplan_carve(bool tri, int b, int n, int t, CellPlan* plan, Workspace* ws, PlanScratch* sc) {
  plan->lcap = (int)lcap_of(n, t);
  plan->light_max = tri ? kLightMaxTri : kLightMaxMean;
  plan->chunk = tri ? kHeavyChunk : kHeavyChunkMean;
  plan->hcap = n / (plan->light_max + 1) + n / plan->chunk + 2;
  plan->rcap = (int)rcap_of(n, tri);
  // cells and the counters are zeroed by ONE memset: they are carved from one block
  sc->zero_ints = (size_t)b * plan->lcap + (size_t)b * kCtr;
  sc->zero_block = ws->take<int>(sc->zero_ints);
  plan->cells = sc->zero_block;
  sc->ctr = sc->zero_block ? sc->zero_block + (size_t)b * plan->lcap : nullptr;
  plan->slot0 = ws->take<int>((size_t)b * plan->lcap);
  plan->meta = ws->take<int>((size_t)b * kMetaInts);
  plan->heavy = (int*)ws->take<int4>((size_t)b * plan->hcap);
  plan->rec = tri ? (void*)ws->take<float4>((size_t)b * plan->rcap) : (void*)ws->take<int>((size_t)b * plan->rcap);
  plan->rec2 = tri ? (void*)ws->take<float4>((size_t)b * plan->rcap) : nullptr;
  sc->rank = ws->take<int>((size_t)b * n);
  sc->aux = ws->take<int>((size_t)b * t);
  sc->recbase = ws->take<int>((size_t)b * t);
  return ws->ok();
}

int cellplan_attach(bool tri, int b, int n, int t, CellPlan* plan, Workspace* ws) {
  PlanScratch sc;
  if (!cellplan_carve(tri, b, n, t, plan, ws, &sc)) return PCFM_E_WORKSPACE;
  plan->pdl = false;  // no plan kernel in front of the consumer
  return 0;
}

int cellplan_build(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                   int* cnt, CellPlan* plan, Workspace* ws, cudaStream_t st) {
  const bool tri = float_coords != nullptr;
  PlanScratch sc;
  cellplan_carve(tri, b, n, t, plan, ws, &sc);
  int* zero_block = sc.zero_block;
  int *ctr = sc.ctr, *rank = sc.rank, *aux = sc.aux, *recbase = sc.recbase;
  const size_t zero_ints = sc.zero_ints;
  if (!cnt) cnt = ws->take<int>((size_t)b * t);
  if (!ws->ok()) return PCFM_E_WORKSPACE;
  const int tq = (cdiv(cdiv(t, 32), kCl)) * 32;  // cells per CTA of the cluster
  const size_t tab_bytes = (size_t)tq * sizeof(int);
  if (g_opt_plan == 0 && n > 0 && n <= kFusedMaxPoints && tab_bytes <= kFusedMaxTable) {
    PCFM_CUDA(allow_dyn_smem((const void*)cp_cluster_kernel<true>, (int)kFusedMaxTable));
    PCFM_CUDA(allow_dyn_smem((const void*)cp_cluster_kernel<false>, (int)kFusedMaxTable));
    const bool chain = g_opt_pdl >= 2;  // see launch_dependent(): overlap the tail of the previous kernel on the stream
    if (tri)
      PCFM_CUDA(launch_dependent(cp_cluster_kernel<true>, dim3(b * kCl), dim3(kPlanThreads), tab_bytes, st, chain,
                                 (const int*)nullptr, float_coords, r, (int*)nullptr, n, t, tq, plan->light_max,
                                 plan->chunk, (int*)nullptr, rank, plan->cells, plan->slot0, plan->meta, plan->heavy,
                                 plan->rec, plan->rec2, plan->lcap, plan->hcap, plan->rcap));
    else
      PCFM_CUDA(launch_dependent(cp_cluster_kernel<false>, dim3(b * kCl), dim3(kPlanThreads), tab_bytes, st, chain,
                                 vox_coords, (const float*)nullptr, r, ind_out, n, t, tq, plan->light_max, plan->chunk,
                                 cnt, rank, plan->cells, plan->slot0, plan->meta, plan->heavy, plan->rec, plan->rec2,
                                 plan->lcap, plan->hcap, plan->rcap));
    PCFM_LAUNCHED();
    plan->pdl = g_opt_pdl != 0;
    return 0;
  }
  PCFM_CUDA(cudaMemsetAsync(cnt, 0, (size_t)b * t * sizeof(int), st));
  PCFM_CUDA(cudaMemsetAsync(zero_block, 0, zero_ints * sizeof(int), st));  // counters; holes in `cells` = empty
  dim3 gp(cdiv(n > 0 ? n : 1, 256), b), gt(cdiv(t, 256), b);
  if (n > 0) {
    if (tri) cp_count_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, nullptr, n, t, cnt, rank);
    else cp_count_kernel<false><<<gp, 256, 0, st>>>(vox_coords, nullptr, r, ind_out, n, t, cnt, rank);
    PCFM_LAUNCHED();
  }
  if (tri) cp_classify_kernel<true><<<gt, 256, 0, st>>>(cnt, t, plan->light_max, plan->chunk, ctr, aux, recbase);
  else cp_classify_kernel<false><<<gt, 256, 0, st>>>(cnt, t, plan->light_max, plan->chunk, ctr, aux, recbase);
  PCFM_LAUNCHED();
  if (tri)
    cp_layout_kernel<true><<<gt, 256, 0, st>>>(cnt, t, r, plan->light_max, plan->chunk, ctr, aux, recbase, plan->cells,
                                               plan->slot0, plan->meta, plan->heavy, plan->rec, plan->rec2, n,
                                               plan->lcap, plan->hcap, plan->rcap);
  else
    cp_layout_kernel<false><<<gt, 256, 0, st>>>(cnt, t, r, plan->light_max, plan->chunk, ctr, aux, recbase, plan->cells,
                                                plan->slot0, plan->meta, plan->heavy, plan->rec, plan->rec2, n,
                                                plan->lcap, plan->hcap, plan->rcap);
  PCFM_LAUNCHED();
  if (n > 0) {
    if (tri) {
      // no ordering pass for trilinear plans: their consumer adds cell sums with shared-memory atomics, whose order
      // is not fixed anyway
      cp_fill_kernel<true><<<gp, 256, 0, st>>>(nullptr, float_coords, r, n, t, rank, recbase, plan->rec, plan->rec2,
                                               plan->rcap);
      PCFM_LAUNCHED();
    } else {
      dim3 gl(cdiv(plan->lcap, 256), b);
      cp_fill_kernel<false><<<gp, 256, 0, st>>>(ind_out, nullptr, r, n, t, rank, recbase, plan->rec, plan->rec2,
                                                plan->rcap);
      PCFM_LAUNCHED();
      cp_order_kernel<<<gl, 256, 0, st>>>(plan->cells, plan->slot0, plan->meta, (int*)plan->rec, plan->lcap,
                                          plan->rcap);
      PCFM_LAUNCHED();
    }
  }
  return 0;
}


int cellplan_build_dual(const int* vox_coords, const float* float_coords, int r, int* ind_out, int b, int n, int t,
                        int* cnt, CellPlan* pm, Workspace* wsm, CellPlan* pt, Workspace* wst, cudaStream_t st) {
  const int tq = (cdiv(cdiv(t, 32), kCl)) * 32;
  const size_t tab_bytes = (size_t)tq * sizeof(int);
  const bool fused = g_opt_plan == 0 && n > 0 && n <= kFusedMaxPoints && tab_bytes <= kFusedMaxTable;
  if (!fused) {  // two separate builds
    int rc = cellplan_build(vox_coords, nullptr, r, ind_out, b, n, t, cnt, pm, wsm, st);
    if (rc) return rc;
    return cellplan_build(nullptr, float_coords, r, nullptr, b, n, t, nullptr, pt, wst, st);
  }
  PlanScratch sm, stt;
  if (!cellplan_carve(false, b, n, t, pm, wsm, &sm) || !cellplan_carve(true, b, n, t, pt, wst, &stt))
    return PCFM_E_WORKSPACE;
  PlanOut om = {sm.rank, pm->cells, pm->slot0, pm->meta, pm->heavy, pm->rec, pm->rec2, pm->lcap, pm->hcap, pm->rcap,
                pm->light_max, pm->chunk};
  PlanOut ot = {stt.rank, pt->cells, pt->slot0, pt->meta, pt->heavy, pt->rec, pt->rec2, pt->lcap, pt->hcap, pt->rcap,
                pt->light_max, pt->chunk};
  PCFM_CUDA(allow_dyn_smem((const void*)cp_cluster_dual_kernel, (int)kFusedMaxTable));
  PCFM_CUDA(launch_dependent(cp_cluster_dual_kernel, dim3(2 * b * kCl), dim3(kPlanThreads), tab_bytes, st,
                             g_opt_pdl >= 2, vox_coords, float_coords, r, ind_out, cnt, n, t, tq, b, om, ot));
  PCFM_LAUNCHED();
  pm->pdl = g_opt_pdl != 0;
  pt->pdl = false;  // its consumer runs much later (the backward pass), behind other kernels
  return 0;
}

}  // namespace pcfm
