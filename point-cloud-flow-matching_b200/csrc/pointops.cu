// pointops.cu — ball query, furthest point sampling, 3-nearest-neighbour search (+ interpolate forward).
#include "rows.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pcfm {

// =================================================================================================
// ball query (ball_query.cu:19-50): one WARP per centre.  Points are streamed through shared memory in
// tiles shared by the CTA's warps; 32 candidates are tested per step, hits are compacted with ballot +
// prefix-popcount so the output keeps the reference's "first U in index order" rule, with early exit.
// =================================================================================================
// packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each lane rounded exactly like the scalar instruction
__device__ __forceinline__ unsigned long long nn_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long nn_sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void nn_sq3x2(unsigned long long dx, unsigned long long dy, unsigned long long dz, float& lo,
                                         float& hi) {  // sq3(): FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
  unsigned long long t;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dx), "l"(t));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dz), "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(t));
}

constexpr int kBqWarps = 16;
constexpr int kBqTile = 2048;

__global__ void __launch_bounds__(kBqWarps * 32) ball_query_kernel(const float* __restrict__ centers,
                                                                   const float* __restrict__ points, int n,
                                                                   int m, float r2, int u,
                                                                   int* __restrict__ out) {
  __shared__ __align__(8) float px[kBqTile], py[kBqTile], pz[kBqTile];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * kBqWarps + wid;
  const bool valid = j < m;
  const float* cb = centers + (long long)b * 3 * m;
  const float* pb = points + (long long)b * 3 * n;
  int* ob = out + ((long long)b * m + (valid ? j : 0)) * u;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (valid) {
    cx = cb[j];
    cy = cb[(long long)j + m];
    cz = cb[(long long)j + 2LL * m];
  }
  int cnt = 0, first = 0;
  for (int base = 0; base < n; base += kBqTile) {
    const int len = min(kBqTile, n - base);
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
      px[k] = pb[base + k];
      py[k] = pb[(long long)base + k + n];
      pz[k] = pb[(long long)base + k + 2LL * n];
    }
    __syncthreads();
    if (valid && cnt < u) {
      // 64 points per step: lane L tests the points 2L and 2L + 1 of the step with one packed evaluation (the three
      // coordinate pairs arrive as LDS.64); hits are placed in index order: ranks from the two ballot masks
      const unsigned long long cxp = nn_pack2(cx, cx), cyp = nn_pack2(cy, cy), czp = nn_pack2(cz, cz);
      const unsigned lt = (1u << lane) - 1u;
      for (int k0 = 0; k0 < len; k0 += 64) {
        const int k = k0 + 2 * lane;                // kBqTile is a multiple of 64: the pair read stays inside the arrays
        float d0, d1;
        nn_sq3x2(nn_sub2(cxp, *reinterpret_cast<const unsigned long long*>(px + k)),
                 nn_sub2(cyp, *reinterpret_cast<const unsigned long long*>(py + k)),
                 nn_sub2(czp, *reinterpret_cast<const unsigned long long*>(pz + k)), d0, d1);  // centre - point
        const bool h0 = k < len && d0 < r2, h1 = k + 1 < len && d1 < r2;                       // strict <
        const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
        if (m0 | m1) {
          if (cnt == 0) {
            const int f0 = m0 ? 2 * (__ffs(m0) - 1) : 64, f1 = m1 ? 2 * (__ffs(m1) - 1) + 1 : 64;
            first = base + k0 + min(f0, f1);
          }
          const int below = cnt + __popc(m0 & lt) + __popc(m1 & lt);  // hits at lower indices than 2L
          if (h0 && below < u) ob[below] = base + k;
          const int below1 = below + (h0 ? 1 : 0);
          if (h1 && below1 < u) ob[below1] = base + k + 1;
          cnt += __popc(m0) + __popc(m1);
          if (cnt >= u) break;
        }
      }
    }
    const int done = (!valid || cnt >= u) ? 1 : 0;
    if (__syncthreads_and(done)) break;
  }
  if (valid) {
    // slots never reached keep the first hit (ball_query.cu:40-44); a centre without any hit stays 0
    const int filled = min(cnt, u);
    for (int v = filled + lane; v < u; v += 32) ob[v] = first;
  }
}

// =================================================================================================
// furthest point sampling (sampling.cu:86-167): one CTA per batch element, m dependent rounds.
// Winner of a round = arg-max of the running min-distance with the reference's tie-break: the reference
// runs 512 threads, thread t scans k = t, t+512, ... keeping the first strict maximum, then a pairwise tree
// keeps the LEFT (lower thread id) entry on ties => ties go to the smallest (k mod 512), then the smallest k.
// That order is encoded in a 64-bit key so any block size / reduction shape gives the same winner.
// =================================================================================================
__device__ __forceinline__ unsigned long long fps_key(float d, int k) {
  const unsigned lo = ((511u - ((unsigned)k & 511u)) << 22) | (0x3FFFFFu - ((unsigned)k >> 9));
  return ((unsigned long long)__float_as_uint(d) << 32) | lo;
}
__device__ __forceinline__ int fps_key_index(unsigned long long key) {
  const unsigned lo = (unsigned)key;
  return (int)(((0x3FFFFFu - (lo & 0x3FFFFFu)) << 9) | (511u - (lo >> 22)));
}

// warp maximum of 64-bit keys with two REDUX.MAX (high words, then the low words of the lanes that hold the maximum)
// instead of five 64-bit shuffle steps
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

template <bool SMEM>
__global__ void __launch_bounds__(1024, 1)
    fps_kernel(const float* __restrict__ coords, int n, int m, float* __restrict__ dist_g,
               int* __restrict__ indices) {
  extern __shared__ __align__(16) float sm[];  // SMEM: x[n] y[n] z[n] d[n]
  __shared__ unsigned long long wbest[32];
  __shared__ int s_old;
  if (m <= 0) return;
  const int b = blockIdx.x;
  const float* cb = coords + (long long)b * 3 * n;
  float* dist = SMEM ? sm + 3 * (size_t)n : dist_g + (long long)b * n;
  int* ib = indices + (long long)b * m;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    if (SMEM) {
      sm[k] = cb[k];
      sm[n + k] = cb[(long long)n + k];
      sm[2 * n + k] = cb[2LL * n + k];
    }
    dist[k] = 1e38f;  // sampling.cpp:53
  }
  if (threadIdx.x == 0) ib[0] = 0;
  __syncthreads();
  int old = 0;
  for (int j = 1; j < m; ++j) {
    float x1, y1, z1;
    if (SMEM) {
      x1 = sm[old]; y1 = sm[n + old]; z1 = sm[2 * n + old];
    } else {
      x1 = cb[old]; y1 = cb[(long long)n + old]; z1 = cb[2LL * n + old];
    }
    unsigned long long best = 0ull;  // below every real key (distances are >= 0)
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      float x2, y2, z2;
      if (SMEM) {
        x2 = sm[k]; y2 = sm[n + k]; z2 = sm[2 * n + k];
      } else {
        x2 = cb[k]; y2 = cb[(long long)n + k]; z2 = cb[2LL * n + k];
      }
      const float td = dist[k];
      const float d = sq3(x2 - x1, y2 - y1, z2 - z1);  // point - last selected
      const float d2 = fminf(d, td);
      if (d2 != td) dist[k] = d2;
      const unsigned long long key = fps_key(d2, k);
      best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    if (lane == 0) wbest[wid] = best;
    __syncthreads();
    if (wid == 0) {
      const unsigned long long v = warp_max_u64(lane < nw ? wbest[lane] : 0ull);
      if (lane == 0) {
        const int w = v ? fps_key_index(v) : 0;
        s_old = w;
        ib[j] = w;
      }
    }
    __syncthreads();
    old = s_old;
  }
}

// -------------------------------------------------------------------------------------------------
// Cluster version: one thread-block cluster of 8 CTAs per cloud, every point (x, y, z, running distance) lives in a
// REGISTER of exactly one thread for all m rounds.  Per round: register update -> warp / CTA arg-max -> the winning
// thread publishes (key, x, y, z) in its CTA's slot -> ONE cluster barrier -> every warp reads the 8 slots through
// distributed shared memory and reduces them: the next centre's coordinates arrive with the winner, no extra
// round trip.  Same 64-bit key => same winners as the one-CTA kernel and the reference.
// -------------------------------------------------------------------------------------------------
constexpr int kFpsCl = 8;
constexpr int kFpsClusterMin = 3072;  // measured crossover (B = 8, M = N / 4): 2048 points equal, 4096 points 1.44x for the cluster
struct __align__(16) FpsSlot {
  unsigned long long key;
  float x, y, z;
  int pad;
};

// XYZ_SMEM (clouds of 65537..131072 points, 16 points per thread): the coordinates live in shared memory
// ([3][points of this CTA], conflict-free strided reads), only the running distances stay in registers.
// PUSH: instead of publishing the slot locally and meeting at a cluster barrier (arrive + wait, then a remote read: two
// cluster round trips per round), the winner's warp WRITES the slot into the inbox of all 8 CTAs with st.async, whose
// completion is counted by an mbarrier in the receiving CTA (complete_tx): one one-way trip per round, and the slots are
// read from local shared memory.  Inboxes and barriers are double-buffered by round parity: a CTA can only be sent the
// slots of round j + 2 after it has itself sent round j + 1, i.e. after all its warps have consumed round j.
template <int PPT, bool XYZ_SMEM, bool PUSH>
__global__ void __cluster_dims__(kFpsCl, 1, 1) __launch_bounds__(1024, 1)
    fps_cluster_kernel(const float* __restrict__ coords, int n, int m, int* __restrict__ indices) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sxyz[];  // XYZ_SMEM: x[per] y[per] z[per]
  __shared__ unsigned long long wbest[32];
  __shared__ FpsSlot slot[2];
  __shared__ FpsSlot inbox[2][kFpsCl];
  __shared__ uint64_t inbar[2];
  if (m <= 0) return;
  const int q = (int)cluster.block_rank(), b = blockIdx.x / kFpsCl;
  const float* cb = coords + (long long)b * 3 * n;
  int* ib = indices + (long long)b * m;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int per = (n + kFpsCl - 1) / kFpsCl, k_lo = min(n, q * per), k_hi = min(n, k_lo + per);
  float px[XYZ_SMEM ? 1 : PPT], py[XYZ_SMEM ? 1 : PPT], pz[XYZ_SMEM ? 1 : PPT], pd[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = k_lo + tid + i * nt;
    const bool v = k < k_hi;
    const float x = v ? cb[k] : 0.f, y = v ? cb[(long long)n + k] : 0.f, z = v ? cb[2LL * n + k] : 0.f;
    if (XYZ_SMEM) {
      if (v) { sxyz[k - k_lo] = x; sxyz[per + k - k_lo] = y; sxyz[2 * per + k - k_lo] = z; }
    } else {
      px[i] = x; py[i] = y; pz[i] = z;
    }
    pd[i] = 1e38f;  // sampling.cpp:53
  }
  if (XYZ_SMEM) __syncthreads();
  if (q == 0 && tid == 0) ib[0] = 0;
  if (PUSH) {
    if (tid == 0) {
      mbar_init(&inbar[0], 1);
      mbar_init(&inbar[1], 1);
      fence_mbar_init();
    }
    cluster.sync();  // nobody sends before every barrier of the cluster exists
  }
  float x1 = cb[0], y1 = cb[n], z1 = cb[2LL * n];  // the first centre is point 0
  for (int j = 1; j < m; ++j) {
    if (PUSH && tid == 0) mbar_arrive_expect_tx(&inbar[j & 1], kFpsCl * 24u);  // 8 slots of 24 bytes will arrive
    unsigned long long best = 0ull;  // below every real key (distances are >= 0)
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = k_lo + tid + i * nt;
      float x, y, z;
      if (XYZ_SMEM) {
        const int l = min(k, k_hi - 1) - k_lo;  // clamped: an idle slot re-reads a valid point, its key is 0
        x = sxyz[l]; y = sxyz[per + l]; z = sxyz[2 * per + l];
      } else {
        x = px[i]; y = py[i]; z = pz[i];
      }
      const float d2 = fminf(sq3(x - x1, y - y1, z - z1), pd[i]);  // point - last selected
      pd[i] = d2;
      const unsigned long long key = k < k_hi ? fps_key(d2, k) : 0ull;
      if (key > best) { best = key; bx = x; by = y; bz = z; }
    }
    unsigned long long v = warp_max_u64(best);
    if (lane == 0) wbest[wid] = v;
    __syncthreads();
    v = warp_max_u64(lane < nw ? wbest[lane] : 0ull);  // every warp reduces the warp results itself: one barrier, no broadcast
    unsigned long long ck = 0ull;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (PUSH) {
      // the warp of the winning thread (warp 0 if this CTA owns no point: key 0) sends the slot to the 8 inboxes
      const unsigned who = __ballot_sync(0xffffffffu, v != 0ull && best == v);  // keys are unique: at most one lane
      if (v != 0ull ? who != 0u : wid == 0) {
        const int src = who ? __ffs(who) - 1 : 0;
        const float sx = v ? __shfl_sync(0xffffffffu, bx, src) : 0.f, sy = v ? __shfl_sync(0xffffffffu, by, src) : 0.f,
                    sz = v ? __shfl_sync(0xffffffffu, bz, src) : 0.f;
        if (lane < kFpsCl) {
          uint32_t dst, rbar;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(&inbox[j & 1][q])), "r"(lane));
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&inbar[j & 1])), "r"(lane));
          const unsigned long long xy = ((unsigned long long)__float_as_uint(sy) << 32) | __float_as_uint(sx);
          const unsigned long long zp = (unsigned long long)__float_as_uint(sz);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                       ::"r"(dst), "l"(v), "l"(xy), "r"(rbar) : "memory");
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                       ::"r"(dst + 16u), "l"(zp), "r"(rbar) : "memory");
        }
      }
      mbar_wait(&inbar[j & 1], (uint32_t)((j - 1) >> 1) & 1u);  // all 8 slots of this round have landed here (rounds start at 1)
      if (lane < kFpsCl) {
        const FpsSlot* rs = &inbox[j & 1][lane];
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(rs);    // key, (x, y)
        ck = a.x;
        cx = __uint_as_float((unsigned)a.y);
        cy = __uint_as_float((unsigned)(a.y >> 32));
        cz = rs->z;
      }
    } else {
      FpsSlot* mine = &slot[j & 1];
      if (v == 0ull) {
        if (tid == 0) { mine->key = 0ull; mine->x = 0.f; mine->y = 0.f; mine->z = 0.f; }  // this CTA owns no point
      } else if (best == v) {  // keys are unique per point: exactly one thread
        mine->key = v; mine->x = bx; mine->y = by; mine->z = bz;
      }
      cluster.sync();
      // lanes 0..7 of every warp fetch one CTA's slot each, 3 shuffle steps pick the cluster-wide winner
      if (lane < kFpsCl) {
        const FpsSlot* rs = cluster.map_shared_rank(mine, lane);
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(rs);            // key, (x, y)
        const float2 zz = *reinterpret_cast<const float2*>(&rs->z);                // z, pad
        ck = a.x;
        cx = __uint_as_float((unsigned)a.y);
        cy = __uint_as_float((unsigned)(a.y >> 32));
        cz = zz.x;
      }
    }
    // the cluster-wide winner among the 8 slots (lanes >= 8 hold key 0) and the lane that holds it
    const unsigned long long win = warp_max_u64(ck);
    const int wl = __ffs(__ballot_sync(0xffffffffu, ck == win)) - 1;
    ck = win;
    x1 = __shfl_sync(0xffffffffu, cx, wl);
    y1 = __shfl_sync(0xffffffffu, cy, wl);
    z1 = __shfl_sync(0xffffffffu, cz, wl);
    if (ck == 0ull) { x1 = cb[0]; y1 = cb[n]; z1 = cb[2LL * n]; }  // degenerate (all distances 0 at index 0): index 0
    if (q == 0 && tid == 0) ib[j] = ck ? fps_key_index(ck) : 0;
  }
  cluster.sync();  // no CTA leaves while its slots may still be read
}

// =================================================================================================
// three nearest neighbours (neighbor_interpolate.cu:20-75): one thread per point, centres streamed
// through shared memory; strict-< insertion keeps the earliest centre on ties.
// =================================================================================================
constexpr int kNnTile = 2048;

__global__ void __launch_bounds__(256) three_nn_kernel(const float* __restrict__ points,
                                                       const float* __restrict__ centers, int n, int m,
                                                       float* __restrict__ weights, int* __restrict__ indices) {
  griddep_launch_dependents();  // the interpolation gather behind this kernel may start staging its planes
  __shared__ __align__(16) float cx[kNnTile], cy[kNnTile], cz[kNnTile];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = j < n;
  const float* pb = points + (long long)b * 3 * n;
  const float* cb = centers + (long long)b * 3 * m;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (valid) {
    ux = pb[j];
    uy = pb[(long long)j + n];
    uz = pb[(long long)j + 2LL * n];
  }
  // the reference keeps the bests in double initialised to 1e40 (> FLT_MAX): +inf behaves identically
  float best0 = __int_as_float(0x7f800000), best1 = best0, best2 = best0;
  int i0 = 0, i1 = 0, i2 = 0;
  for (int base = 0; base < m; base += kNnTile) {
    const int len = min(kNnTile, m - base);
    __syncthreads();
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
      cx[k] = cb[base + k];
      cy[k] = cb[(long long)base + k + m];
      cz[k] = cb[(long long)base + k + 2LL * m];
    }
    __syncthreads();
    if (valid) {
      // four centres per step: one LDS.128 per coordinate, distances as packed pairs (FADD2 / FMUL2 / FFMA2, lanes
      // rounded like the scalar sq3); the exact ordered insertion runs only when one of the four beats the third best
      auto insert = [&](float d, int k) {
        if (d < best2) {
          best2 = d; i2 = k;
          if (d < best1) {
            best2 = best1; i2 = i1;
            best1 = d; i1 = k;
            if (d < best0) {
              best1 = best0; i1 = i0;
              best0 = d; i0 = k;
            }
          }
        }
      };
      const unsigned long long uxp = nn_pack2(ux, ux), uyp = nn_pack2(uy, uy), uzp = nn_pack2(uz, uz);
      int k = 0;
      for (; k + 4 <= len; k += 4) {
        const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(cx + k);  // (x0,x1) (x2,x3)
        const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(cy + k);
        const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(cz + k);
        float d0, d1, d2, d3;
        nn_sq3x2(nn_sub2(uxp, qx.x), nn_sub2(uyp, qy.x), nn_sub2(uzp, qz.x), d0, d1);
        nn_sq3x2(nn_sub2(uxp, qx.y), nn_sub2(uyp, qy.y), nn_sub2(uzp, qz.y), d2, d3);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) < best2) {
          insert(d0, base + k);
          insert(d1, base + k + 1);
          insert(d2, base + k + 2);
          insert(d3, base + k + 3);
        }
      }
      for (; k < len; ++k) insert(sq3(ux - cx[k], uy - cy[k], uz - cz[k]), base + k);
    }
  }
  if (!valid) return;
  best0 = fmaxf(fminf(1e10f, best0), 1e-10f);
  best1 = fmaxf(fminf(1e10f, best1), 1e-10f);
  best2 = fmaxf(fminf(1e10f, best2), 1e-10f);
  const float d0d1 = __fmul_rn(best0, best1);
  const float d0d2 = __fmul_rn(best0, best2);
  const float d1d2 = __fmul_rn(best1, best2);
  const float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
  float* wb = weights + (long long)b * 3 * n;
  int* ib = indices + (long long)b * 3 * n;
  wb[j] = __fmul_rn(d1d2, inv);
  ib[j] = i0;
  wb[(long long)j + n] = __fmul_rn(d0d2, inv);
  ib[(long long)j + n] = i1;
  wb[(long long)j + 2LL * n] = __fmul_rn(d0d1, inv);
  ib[(long long)j + 2LL * n] = i2;
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

int pcfm_ball_query(const float* centers, const float* points, int b, int n, int m, float radius, int u,
                    int* idx, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || u < 0) return PCFM_E_BADARG;
  if (b == 0 || m == 0 || u == 0) return 0;
  if (!centers || !idx || (n > 0 && !points)) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    PCFM_CUDA(cudaMemsetAsync(idx, 0, (size_t)b * m * u * sizeof(int), st));
    return 0;
  }
  dim3 grid(cdiv(m, kBqWarps), b);
  ball_query_kernel<<<grid, kBqWarps * 32, 0, st>>>(centers, points, n, m, radius * radius, u, idx);
  PCFM_LAUNCHED();
  return 0;
}

size_t pcfm_furthest_point_sampling_workspace(int b, int n, int m) {
  (void)m;
  if (b <= 0 || n <= 0) return 256;
  return align256((size_t)b * n * sizeof(float)) + 256;
}

int pcfm_furthest_point_sampling(const float* coords, int b, int n, int m, int* indices, void* workspace,
                                 size_t workspace_bytes, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || m == 0) return 0;
  if (!coords || !indices || n == 0) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  // Cluster kernel from 3072 points up (below, the ~0.55 us exchange per round loses to the one-CTA kernel):
  // points in registers (<= 8 per thread) up to 65536 points, coordinates in shared memory up to 131072.
  // option "fps": 0 = cluster kernel, slots pushed into the peers' inboxes (default), 2 = cluster kernel with a cluster
  // barrier per round, 1 = always one CTA per cloud
  if (g_opt_fps != 1 && n >= kFpsClusterMin && n <= kFpsCl * 1024 * 16) {
    const bool push = g_opt_fps == 0;
    const int per = cdiv(n, kFpsCl);
    dim3 grid(b * kFpsCl);
    if (per > 1024 * 8) {
      const size_t smem = (size_t)per * 3 * sizeof(float);
      PCFM_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     16384 * 3 * (int)sizeof(float)));
      PCFM_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     16384 * 3 * (int)sizeof(float)));
      if (push) fps_cluster_kernel<16, true, true><<<grid, 1024, smem, st>>>(coords, n, m, indices);
      else fps_cluster_kernel<16, true, false><<<grid, 1024, smem, st>>>(coords, n, m, indices);
    } else {
      const int nt = cdiv(per, 512) <= 8 ? 512 : 1024;
      const int ppt = cdiv(per, nt);  // 1..8 points per thread, no padded work beyond the last thread's
#define PCFM_FPS(P)                                                                                  \
  do {                                                                                               \
    if (push) fps_cluster_kernel<P, false, true><<<grid, nt, 0, st>>>(coords, n, m, indices);        \
    else fps_cluster_kernel<P, false, false><<<grid, nt, 0, st>>>(coords, n, m, indices);            \
  } while (0)
      switch (ppt) {
        case 1: PCFM_FPS(1); break;
        case 2: PCFM_FPS(2); break;
        case 3: PCFM_FPS(3); break;
        case 4: PCFM_FPS(4); break;
        case 5: PCFM_FPS(5); break;
        case 6: PCFM_FPS(6); break;
        case 7: PCFM_FPS(7); break;
        default: PCFM_FPS(8); break;
      }
#undef PCFM_FPS
    }
    PCFM_LAUNCHED();
    return 0;
  }
  const size_t smem = (size_t)n * 4 * sizeof(float);
  if (smem <= (size_t)kMaxStageFloats * sizeof(float)) {
    PCFM_CUDA(allow_dyn_smem((const void*)fps_kernel<true>, (int)(kMaxStageFloats * sizeof(float))));
    fps_kernel<true><<<b, 1024, smem, st>>>(coords, n, m, nullptr, indices);
  } else {
    Workspace ws(workspace, workspace_bytes);
    float* dist = ws.take<float>((size_t)b * n);
    if (!ws.ok()) return PCFM_E_WORKSPACE;
    fps_kernel<false><<<b, 1024, 0, st>>>(coords, n, m, dist, indices);
  }
  PCFM_LAUNCHED();
  return 0;
}

int pcfm_three_nn_interpolate_forward(const float* points, const float* centers, const float* centers_features,
                                      int b, int c, int n, int m, float* out, int* idx, float* wgt,
                                      pcfm_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0) return 0;
  if (!points || !idx || !wgt || (m > 0 && !centers)) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(n, 256), b);
  three_nn_kernel<<<grid, 256, 0, st>>>(points, centers, n, m, wgt, idx);
  PCFM_LAUNCHED();
  if (c == 0) return 0;
  if (!centers_features || !out || m == 0) return PCFM_E_BADARG;
  return launch_row_gather(3, centers_features, idx, wgt, out, b, c, m, n, st, true);  // three_nn_kernel releases it early
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// Voxelization.forward tail (voxelization.py:19-24): scale, shift, clamp, round — one pass, two outputs
// -------------------------------------------------------------------------------------------------
namespace pcfm {
__global__ void __launch_bounds__(256) voxel_coords_kernel(const float* __restrict__ centred,
                                                           const float* __restrict__ scale, long long per_batch,
                                                           float rf, float hi, float* __restrict__ nc,
                                                           int* __restrict__ vox) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_batch) return;
  const long long j = (long long)b * per_batch + i;
  float v = centred[j];
  if (scale) v = __fadd_rn(__fdiv_rn(v, scale[b]), 0.5f);  // torch: x / tensor is a true division
  else v = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);            // (x + 1) / 2.0
  v = __fmul_rn(v, rf);
  v = v < 0.f ? 0.f : (v > hi ? hi : v);                   // torch.clamp: NaN stays NaN
  nc[j] = v;
  vox[j] = __float2int_rn(v);                              // torch.round = half to even, then .to(int32)
}
}  // namespace pcfm

extern "C" int pcfm_voxel_coords(const float* centred, const float* scale, int b, int n, int r, float* norm_coords,
                                 int* vox_coords, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0) return 0;
  if (!centred || !norm_coords || !vox_coords) return PCFM_E_BADARG;
  const long long per_batch = 3LL * n;
  dim3 grid(pcfm::cdiv(per_batch, 256), b);
  pcfm::voxel_coords_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(centred, scale, per_batch, (float)r,
                                                                    (float)(r - 1), norm_coords, vox_coords);
  PCFM_LAUNCHED();
  return 0;
}
