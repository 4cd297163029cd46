// chamfer.cu — Chamfer-3D nearest neighbour (chamfer3D.cu:12-154), its gradient (:155-195) and the
// all-pairs block kernel used by the sharded evaluation.
//
// fp32-ALU bound: every thread keeps RA "own" points in registers and streams the other
// cloud from shared memory as broadcast LDS.128.  The distance is evaluated exactly as the reference's SASS
// (sq3 on other-self), so minima are bit-identical; arg-min = lowest index (strict < in ascending order).
#include "common.cuh"

namespace pcfm {

constexpr int kChThreads = 128;
constexpr int kChChunk = 2048;  // points of the other cloud per shared-memory chunk (32 KB as float4)

__device__ __forceinline__ float pos_inf() { return __int_as_float(0x7f800000); }

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2), each lane rounded exactly like the scalar instruction ----
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// sq3() of common.cuh on two lanes: FMUL2(dy,dy); FFMA2(dx,dx,.); FFMA2(dz,dz,.) -> bit patterns of the two results
__device__ __forceinline__ void sq3x2(unsigned long long dx, unsigned long long dy, unsigned long long dz,
                                      unsigned& lo, unsigned& hi) {
  unsigned long long t;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dx), "l"(t));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dz), "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
}


// one direction: for every point of `a` ([B,n,3]) the nearest point of `o` ([B,m,3])
template <int RA>
__global__ void __launch_bounds__(kChThreads)
    chamfer_nn_kernel(const float* __restrict__ a, const float* __restrict__ o, int n, int m,
                      float* __restrict__ dist, int* __restrict__ idx) {
  // the other cloud in SoA groups of four: one LDS.128 per coordinate yields two natural register PAIRS of columns,
  // so the distance arithmetic issues as packed FADD2 / FMUL2 / FFMA2 (two columns per instruction, lanes rounded
  // exactly like the scalar sq3)
  __shared__ __align__(16) float bx[kChChunk], by[kChChunk], bz[kChChunk];
  const int b = blockIdx.y;
  const float* ab = a + (long long)b * n * 3;
  const float* ob = o + (long long)b * m * 3;
  const int row0 = blockIdx.x * (kChThreads * RA) + threadIdx.x;
  float ax[RA], ay[RA], az[RA], best[RA];
  unsigned long long axp[RA], ayp[RA], azp[RA];
  int bk[RA];
#pragma unroll
  for (int r = 0; r < RA; ++r) {
    const int i = row0 + r * kChThreads;
    const bool v = i < n;
    ax[r] = v ? ab[(long long)i * 3 + 0] : 0.f;
    ay[r] = v ? ab[(long long)i * 3 + 1] : 0.f;
    az[r] = v ? ab[(long long)i * 3 + 2] : 0.f;
    axp[r] = pack2(ax[r], ax[r]);
    ayp[r] = pack2(ay[r], ay[r]);
    azp[r] = pack2(az[r], az[r]);
    best[r] = pos_inf();
    bk[r] = 0;
  }
  for (int k2 = 0; k2 < m; k2 += kChChunk) {
    const int len = min(kChChunk, m - k2);
    const int len4 = (len + 3) & ~3;
    __syncthreads();
    for (int k = threadIdx.x; k < len4; k += kChThreads) {
      const bool v = k < len;  // padding: distance +inf, never a strict minimum
      const float* p = ob + (long long)(k2 + (v ? k : 0)) * 3;
      bx[k] = v ? p[0] : pos_inf();
      by[k] = v ? p[1] : pos_inf();
      bz[k] = v ? p[2] : pos_inf();
    }
    __syncthreads();
    for (int k = 0; k < len4; k += 4) {
      const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(bx + k);  // (x0,x1) (x2,x3)
      const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(by + k);
      const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(bz + k);
#pragma unroll
      for (int r = 0; r < RA; ++r) {
        unsigned u0, u1, u2, u3;
        sq3x2(sub2(qx.x, axp[r]), sub2(qy.x, ayp[r]), sub2(qz.x, azp[r]), u0, u1);
        sq3x2(sub2(qx.y, axp[r]), sub2(qy.y, ayp[r]), sub2(qz.y, azp[r]), u2, u3);
        const float m4 = fminf(fminf(__uint_as_float(u0), __uint_as_float(u1)),
                               fminf(__uint_as_float(u2), __uint_as_float(u3)));
        if (m4 < best[r]) {  // strict: the first group of four that reaches the minimum wins
          best[r] = m4;
          bk[r] = k2 + k;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RA; ++r) {
    const int i = row0 + r * kChThreads;
    if (i >= n) continue;
    // lowest index inside the winning group of four: recompute the (bit-identical) distances
    int bi = bk[r];
    for (int t = 0; t < 4; ++t) {
      const int k = bk[r] + t;
      if (k >= m) break;
      const float* p = ob + (long long)k * 3;
      const float d = sq3(p[0] - ax[r], p[1] - ay[r], p[2] - az[r]);
      if (d == best[r]) {
        bi = k;
        break;
      }
    }
    dist[(long long)b * n + i] = best[r];
    idx[(long long)b * n + i] = bi;
  }
}


// =================================================================================================
// Single-pass operator: every squared distance is evaluated ONCE and feeds both directions (the reference and the
// kernel above evaluate it twice, once per direction: chamfer3D.cu:136-154 launches NmDistanceKernel twice).
//
// Tile = RA*128 rows of xyz1 (thread t owns the RA CONTIGUOUS rows t*RA .. t*RA+RA-1, in registers) x a range of
// columns of xyz2 (streamed through shared memory).  Row side: running minimum per row over groups of four columns
// (strict <: the first group that reaches the minimum wins, resolved to the lowest column at the end), exactly as
// chamfer_nn_kernel.  Column side: per column the minimum over a thread's rows is folded on the IEEE bit patterns
// (distances are >= +0, where unsigned order == float order; VIMNMX3 folds two rows per instruction), reduced over
// the warp with REDUX.MIN, and the LOWEST lane that holds the minimum is elected: because a thread's rows are
// contiguous and threads are numbered in row order, the lowest thread that attains the minimum contains the lowest
// row that attains it.  Results of different tiles meet in 64-bit keys  (bits(d) << 32) | index  combined with
// atomicMin: smallest distance first, lowest index on ties — the reference's tie rule (chamfer3D.cu:36,126) without any
// ordering between tiles.  A small kernel unpacks the keys and resolves the winning thread of a column to its row.
// Issue slots per distance: 3 (packed arithmetic) + ~1 (row minimum) + 0.5 (column minimum), against ~4.5 per
// direction before.
// =================================================================================================
constexpr int kOnceThreads = 128;
constexpr int kOnceWarps = kOnceThreads / 32;
constexpr int kOnceChunk = 1024;   // columns per shared-memory chunk: 12 KB of coordinates + 32 KB of per-warp column keys

template <int RA>
__global__ void __launch_bounds__(kOnceThreads)
    chamfer_once_kernel(const float* __restrict__ a, const float* __restrict__ o, int n, int m, int cols_per_tile,
                        unsigned long long* __restrict__ rowkey, unsigned long long* __restrict__ colkey) {
  __shared__ __align__(16) float bx[kOnceChunk], by[kOnceChunk], bz[kOnceChunk];
  // column keys of the current chunk, one row per warp: plain stores, no shared-memory atomics (a 64-bit shared
  // atomicMin is a CAS loop the warp waits for); the warps' rows are folded when the chunk is done
  __shared__ unsigned long long ckey[kOnceWarps][kOnceChunk];
  const int b = blockIdx.z;
  const float* ab = a + (long long)b * n * 3;
  const float* ob = o + (long long)b * m * 3;
  unsigned long long* rk = rowkey + (long long)b * n;
  unsigned long long* ck = colkey + (long long)b * m;
  const int gti = blockIdx.x * kOnceThreads + threadIdx.x;   // thread index in row order: rows gti*RA .. gti*RA+RA-1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ax[RA], ay[RA], az[RA], best[RA];
  unsigned long long axp[RA], ayp[RA], azp[RA];
  int bk[RA];
#pragma unroll
  for (int r = 0; r < RA; ++r) {
    const int i = gti * RA + r;
    const bool v = i < n;   // rows past the end sit at +inf: their distances never lower a minimum
    ax[r] = v ? ab[(long long)i * 3 + 0] : pos_inf();
    ay[r] = v ? ab[(long long)i * 3 + 1] : pos_inf();
    az[r] = v ? ab[(long long)i * 3 + 2] : pos_inf();
    axp[r] = pack2(ax[r], ax[r]);
    ayp[r] = pack2(ay[r], ay[r]);
    azp[r] = pack2(az[r], az[r]);
    best[r] = pos_inf();
    bk[r] = 0;
  }
  const int c_lo = blockIdx.y * cols_per_tile, c_hi = min(m, c_lo + cols_per_tile);
  for (int k2 = c_lo; k2 < c_hi; k2 += kOnceChunk) {
    const int len = min(kOnceChunk, c_hi - k2);
    const int len4 = (len + 3) & ~3;
    __syncthreads();
    for (int k = threadIdx.x; k < len4; k += kOnceThreads) {
      const bool v = k < len;  // padding: distance +inf, never a strict minimum
      const float* p = ob + (long long)(k2 + (v ? k : 0)) * 3;
      bx[k] = v ? p[0] : pos_inf();
      by[k] = v ? p[1] : pos_inf();
      bz[k] = v ? p[2] : pos_inf();
    }
    __syncthreads();
    for (int k = 0; k < len4; k += 4) {
      const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(bx + k);  // (x0,x1) (x2,x3)
      const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(by + k);
      const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(bz + k);
      unsigned cm0 = 0x7f800000u, cm1 = 0x7f800000u, cm2 = 0x7f800000u, cm3 = 0x7f800000u;
#pragma unroll
      for (int r = 0; r < RA; r += 2) {
        unsigned u0, u1, u2, u3, w0, w1, w2, w3;
        sq3x2(sub2(qx.x, axp[r]), sub2(qy.x, ayp[r]), sub2(qz.x, azp[r]), u0, u1);
        sq3x2(sub2(qx.y, axp[r]), sub2(qy.y, ayp[r]), sub2(qz.y, azp[r]), u2, u3);
        sq3x2(sub2(qx.x, axp[r + 1]), sub2(qy.x, ayp[r + 1]), sub2(qz.x, azp[r + 1]), w0, w1);
        sq3x2(sub2(qx.y, axp[r + 1]), sub2(qy.y, ayp[r + 1]), sub2(qz.y, azp[r + 1]), w2, w3);
        const float m4u = fminf(fminf(__uint_as_float(u0), __uint_as_float(u1)),
                                fminf(__uint_as_float(u2), __uint_as_float(u3)));
        const float m4w = fminf(fminf(__uint_as_float(w0), __uint_as_float(w1)),
                                fminf(__uint_as_float(w2), __uint_as_float(w3)));
        if (m4u < best[r]) {  // strict: the first group of four that reaches the minimum wins
          best[r] = m4u;
          bk[r] = k2 + k;
        }
        if (m4w < best[r + 1]) {
          best[r + 1] = m4w;
          bk[r + 1] = k2 + k;
        }
        cm0 = __vimin3_u32(cm0, u0, w0);
        cm1 = __vimin3_u32(cm1, u1, w1);
        cm2 = __vimin3_u32(cm2, u2, w2);
        cm3 = __vimin3_u32(cm3, u3, w3);
      }
      // column side: warp minimum (REDUX.MIN) and the lowest lane that holds it; lane t keeps column t's key and the
      // four keys leave with ONE 8-byte store instruction into this warp's key row
      const unsigned cmv[4] = {cm0, cm1, cm2, cm3};
      unsigned long long mine = ~0ull;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned wm = __reduce_min_sync(0xffffffffu, cmv[t]);
        const unsigned who = __ballot_sync(0xffffffffu, cmv[t] == wm);
        if (lane == t) mine = ((unsigned long long)wm << 32) | (unsigned)(gti - lane + (__ffs(who) - 1));
      }
      if (lane < 4) ckey[warp][k + lane] = mine;
    }
    __syncthreads();
    // fold the warps' key rows; the chunk's keys meet those of the other row tiles in global memory (one RED.MIN.64
    // per column and CTA)
    for (int k = threadIdx.x; k < len; k += kOnceThreads) {
      unsigned long long key = ckey[0][k];
#pragma unroll
      for (int w = 1; w < kOnceWarps; ++w) key = min(key, ckey[w][k]);
      if ((unsigned)(key >> 32) < 0x7f800000u) atomicMin(ck + k2 + k, key);
    }
  }
#pragma unroll
  for (int r = 0; r < RA; ++r) {
    const int i = gti * RA + r;
    if (i >= n) continue;
    // lowest index inside the winning group of four: recompute the (bit-identical) distances
    int bi = bk[r];
    for (int t = 0; t < 4; ++t) {
      const int k = bk[r] + t;
      if (k >= c_hi) break;
      const float* p = ob + (long long)k * 3;
      const float d = sq3(p[0] - ax[r], p[1] - ay[r], p[2] - az[r]);
      if (d == best[r]) {
        bi = k;
        break;
      }
    }
    const unsigned long long key = ((unsigned long long)__float_as_uint(best[r]) << 32) | (unsigned)bi;
    if (gridDim.y == 1) rk[i] = key;
    else atomicMin(rk + i, key);
  }
}

// keys -> (dist, idx); a column's key names the winning THREAD of chamfer_once_kernel: its lowest row at that distance
__global__ void __launch_bounds__(256)
    chamfer_once_finish_kernel(const float* __restrict__ a, const float* __restrict__ o, int n, int m, int ra,
                               const unsigned long long* __restrict__ rowkey,
                               const unsigned long long* __restrict__ colkey, float* __restrict__ dist1,
                               int* __restrict__ idx1, float* __restrict__ dist2, int* __restrict__ idx2) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const unsigned long long key = rowkey[(long long)b * n + j];
    dist1[(long long)b * n + j] = __uint_as_float((unsigned)(key >> 32));
    idx1[(long long)b * n + j] = (int)(unsigned)key;
  }
  if (j < m) {
    const unsigned long long key = colkey[(long long)b * m + j];
    const unsigned bits = (unsigned)(key >> 32);
    const float* p = o + ((long long)b * m + j) * 3;
    const float x = p[0], y = p[1], z = p[2];
    const int r0 = (int)(unsigned)key * ra;
    int bi = r0;
    for (int t = 0; t < ra; ++t) {
      const int i = r0 + t;
      if (i >= n) break;
      const float* q = a + ((long long)b * n + i) * 3;
      // same operand order as the row side of the kernel above (column - row): bit-identical
      if (__float_as_uint(sq3(x - q[0], y - q[1], z - q[2])) == bits) {
        bi = i;
        break;
      }
    }
    dist2[(long long)b * m + j] = __uint_as_float(bits);
    idx2[(long long)b * m + j] = bi;
  }
}

// gradient (chamfer3D.cu:155-174, both launches fused): accumulates into grad1 / grad2
__global__ void __launch_bounds__(256)
    chamfer_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                        const float* __restrict__ gd1, const float* __restrict__ gd2,
                        const int* __restrict__ idx1, const int* __restrict__ idx2, int n, int m,
                        float* __restrict__ g1, float* __restrict__ g2) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float* p1 = xyz1 + (long long)b * n * 3;
  const float* p2 = xyz2 + (long long)b * m * 3;
  float* o1 = g1 + (long long)b * n * 3;
  float* o2 = g2 + (long long)b * m * 3;
  if (j < n) {
    const int j2 = min(max(idx1[(long long)b * n + j], 0), m - 1);
    const float g = gd1[(long long)b * n + j] * 2;
    const float dx = p1[j * 3LL + 0] - p2[j2 * 3LL + 0];
    const float dy = p1[j * 3LL + 1] - p2[j2 * 3LL + 1];
    const float dz = p1[j * 3LL + 2] - p2[j2 * 3LL + 2];
    atomicAdd(o1 + j * 3LL + 0, __fmul_rn(g, dx));
    atomicAdd(o1 + j * 3LL + 1, __fmul_rn(g, dy));
    atomicAdd(o1 + j * 3LL + 2, __fmul_rn(g, dz));
    atomicAdd(o2 + j2 * 3LL + 0, -__fmul_rn(g, dx));
    atomicAdd(o2 + j2 * 3LL + 1, -__fmul_rn(g, dy));
    atomicAdd(o2 + j2 * 3LL + 2, -__fmul_rn(g, dz));
  }
  if (j < m) {
    const int j1 = min(max(idx2[(long long)b * m + j], 0), n - 1);
    const float g = gd2[(long long)b * m + j] * 2;
    const float dx = p2[j * 3LL + 0] - p1[j1 * 3LL + 0];
    const float dy = p2[j * 3LL + 1] - p1[j1 * 3LL + 1];
    const float dz = p2[j * 3LL + 2] - p1[j1 * 3LL + 2];
    atomicAdd(o2 + j * 3LL + 0, __fmul_rn(g, dx));
    atomicAdd(o2 + j * 3LL + 1, __fmul_rn(g, dy));
    atomicAdd(o2 + j * 3LL + 2, __fmul_rn(g, dz));
    atomicAdd(o1 + j1 * 3LL + 0, -__fmul_rn(g, dx));
    atomicAdd(o1 + j1 * 3LL + 1, -__fmul_rn(g, dy));
    atomicAdd(o1 + j1 * 3LL + 2, -__fmul_rn(g, dz));
  }
}

// =================================================================================================
// all-pairs Chamfer: one CTA per (S_i, R_j) cloud pair.  Each squared distance is evaluated ONCE and feeds
// the row minimum (registers) and the column minimum (warp shuffle min, then one shared-memory atomicMin on
// the float bit pattern per warp; non-negative floats order like unsigned ints).  No indices are needed for
// the metric, so the inner loop is 3 FADD + FMUL + 2 FFMA + 2 FMNMX per pair of points.
// =================================================================================================
// 128 threads x 16 own points (2048 rows per sweep).  64 x 32 needs 168 registers per thread: 3 warps per scheduler,
// 68 % issue-active with fixed-latency "wait" stalls (ncu); measured 0.82 / 1.04 / 0.96 M pairs/s for 64x32 / 128x16 /
// 256x8 (the column-min reduction per warp grows as the tile gets narrower).
constexpr int kApThreads = 128;
constexpr int kApRows = 16;
constexpr int kApMaxM = 8192;

// Minima are taken on the IEEE bit patterns with the 3-input integer min (VIMNMX3): squared distances are >= +0,
// where unsigned order == float order, so two distances are folded per instruction for the row side and two for
// the column side.  With the arithmetic packed two rows per instruction (below): 4 issue slots per evaluated pair of
// points, 6 cycles of the fp32 pipe.
__global__ void __launch_bounds__(kApThreads)
    chamfer_allpairs_kernel(const float* __restrict__ s, const float* __restrict__ rr, int nr, int n, int m,
                            float* __restrict__ cd) {
  extern __shared__ __align__(16) unsigned char ap_smem[];
  const int m2 = (m + 1) & ~1;                  // columns are walked two at a time; the pad column sits at +inf
  float4* col = (float4*)ap_smem;               // [m2]
  unsigned* colmin = (unsigned*)(col + m2);     // [m2]
  __shared__ float red[kApThreads / 32];
  __shared__ float red2[kApThreads / 32];
  const long long pair = blockIdx.x;
  const int i = (int)(pair / nr), j = (int)(pair % nr);
  const float* a = s + (long long)i * n * 3;
  const float* o = rr + (long long)j * m * 3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < m2; k += kApThreads) {
    if (k < m) {
      const float* p = o + (long long)k * 3;
      col[k] = make_float4(p[0], p[1], p[2], 0.f);
    } else {
      col[k] = make_float4(pos_inf(), pos_inf(), pos_inf(), 0.f);
    }
    colmin[k] = 0x7f800000u;
  }
  __syncthreads();
  float rowsum = 0.f;
  for (int base = 0; base < n; base += kApThreads * kApRows) {
    // rows are held in PAIRS (r, r+1) packed in 64-bit registers: the subtractions, the square and the two FMAs of a
    // distance are issued as FADD2 / FMUL2 / FFMA2 (sm_100 packed fp32: two IEEE-exact lanes per instruction), which
    // halves the issue slots of the arithmetic: 3 + 1 + 2 packed + 2 VIMNMX3 per TWO distances = 4 slots each
    unsigned long long ax[kApRows / 2], ay[kApRows / 2], az[kApRows / 2];
    unsigned rmin[kApRows];
#pragma unroll
    for (int r = 0; r < kApRows; r += 2) {
      float x[2], y[2], z[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = base + (r + h) * kApThreads + threadIdx.x;
        const bool v = row < n;
        // rows past the end sit at +inf: their distances are +inf (or NaN, whose bit pattern is above +inf's) and
        // never lower a minimum
        x[h] = v ? a[(long long)row * 3 + 0] : pos_inf();
        y[h] = v ? a[(long long)row * 3 + 1] : pos_inf();
        z[h] = v ? a[(long long)row * 3 + 2] : pos_inf();
        rmin[r + h] = 0x7f800000u;
      }
      ax[r / 2] = pack2(x[0], x[1]);
      ay[r / 2] = pack2(y[0], y[1]);
      az[r / 2] = pack2(z[0], z[1]);
    }
    for (int k = 0; k < m2; k += 2) {
      const float4 q0 = col[k], q1 = col[k + 1];
      const unsigned long long q0x = pack2(q0.x, q0.x), q0y = pack2(q0.y, q0.y), q0z = pack2(q0.z, q0.z);
      const unsigned long long q1x = pack2(q1.x, q1.x), q1y = pack2(q1.y, q1.y), q1z = pack2(q1.z, q1.z);
      unsigned cm0 = 0x7f800000u, cm1 = 0x7f800000u;
#pragma unroll
      for (int r = 0; r < kApRows; r += 2) {
        unsigned d00, d10, d01, d11;  // d<row half><column>
        sq3x2(sub2(q0x, ax[r / 2]), sub2(q0y, ay[r / 2]), sub2(q0z, az[r / 2]), d00, d10);
        sq3x2(sub2(q1x, ax[r / 2]), sub2(q1y, ay[r / 2]), sub2(q1z, az[r / 2]), d01, d11);
        rmin[r] = __vimin3_u32(rmin[r], d00, d01);
        rmin[r + 1] = __vimin3_u32(rmin[r + 1], d10, d11);
        cm0 = __vimin3_u32(cm0, d00, d10);
        cm1 = __vimin3_u32(cm1, d01, d11);
      }
      cm0 = __reduce_min_sync(0xffffffffu, cm0);  // REDUX.MIN: one instruction instead of 5 shuffles + 5 mins
      cm1 = __reduce_min_sync(0xffffffffu, cm1);
      if (lane == 0) {
        atomicMin(colmin + k, cm0);
        atomicMin(colmin + k + 1, cm1);
      }
    }
#pragma unroll
    for (int r = 0; r < kApRows; ++r) {
      const int row = base + r * kApThreads + threadIdx.x;
      if (row < n) rowsum += __uint_as_float(rmin[r]);
    }
  }
  __syncthreads();
  float colsum = 0.f;
  for (int k = threadIdx.x; k < m; k += kApThreads) colsum += __uint_as_float(colmin[k]);
  rowsum = warp_sum(rowsum);
  colsum = warp_sum(colsum);
  if (lane == 0) {
    red[wid] = rowsum;
    red2[wid] = colsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float rs = 0.f, cs = 0.f;
    for (int w = 0; w < kApThreads / 32; ++w) {
      rs += red[w];
      cs += red2[w];
    }
    cd[pair] = rs / (float)n + cs / (float)m;
  }
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

int pcfm_chamfer3d_forward(const float* xyz1, const float* xyz2, int b, int n, int m, float* dist1,
                           float* dist2, int* idx1, int* idx2, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;  // the reference's loops do not execute: outputs untouched
  if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int dir = 0; dir < 2; ++dir) {
    const float* a = dir ? xyz2 : xyz1;
    const float* o = dir ? xyz1 : xyz2;
    const int na = dir ? m : n, no = dir ? n : m;
    float* d = dir ? dist2 : dist1;
    int* ix = dir ? idx2 : idx1;
    const long long rows = (long long)na * b;
    const long long want = 2LL * num_sms() * kChThreads;
    if (rows >= want * 4) {
      dim3 grid(cdiv(na, kChThreads * 4), b);
      chamfer_nn_kernel<4><<<grid, kChThreads, 0, st>>>(a, o, na, no, d, ix);
    } else if (rows >= want * 2) {
      dim3 grid(cdiv(na, kChThreads * 2), b);
      chamfer_nn_kernel<2><<<grid, kChThreads, 0, st>>>(a, o, na, no, d, ix);
    } else {
      dim3 grid(cdiv(na, kChThreads), b);
      chamfer_nn_kernel<1><<<grid, kChThreads, 0, st>>>(a, o, na, no, d, ix);
    }
    PCFM_LAUNCHED();
  }
  return 0;
}


size_t pcfm_chamfer3d_forward_workspace(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 256;
  return align256((size_t)b * n * 8) + align256((size_t)b * m * 8) + 256;
}

int pcfm_chamfer3d_forward_ws(const float* xyz1, const float* xyz2, int b, int n, int m, float* dist1, float* dist2,
                              int* idx1, int* idx2, void* workspace, size_t workspace_bytes, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;  // the reference's loops do not execute: outputs untouched
  if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return PCFM_E_BADARG;
  // small problems: the two directional passes (one launch each, no key scratch to clear, no unpacking) are faster
  // (B=32, 2000 x 1000: 0.054 vs 0.104 ms); from ~3e8 distance evaluations on the single pass wins (B=8, 20000^2: 1.36 vs
  // 2.10 ms)
  if (b > 65535 || g_opt_chamfer_once == 0 || (g_opt_chamfer_once == 1 && (long long)b * n * m < (1LL << 28)))
    return pcfm_chamfer3d_forward(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, stream);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* rowkey = ws.take<unsigned long long>((size_t)b * n);
  unsigned long long* colkey = ws.take<unsigned long long>((size_t)b * m);
  if (!ws.ok()) return PCFM_E_WORKSPACE;
  // rows per thread: as many as keep the grid at >= 2 CTAs per SM together with the column split below
  const long long want = 2LL * num_sms();
  int ra = 16;
  while (ra > 4 && (long long)cdiv(n, kOnceThreads * ra) * b * cdiv(m, kOnceChunk) < want) ra >>= 1;
  const int rowtiles = cdiv(n, kOnceThreads * ra);
  int coltiles = (int)((want + (long long)rowtiles * b - 1) / ((long long)rowtiles * b));
  coltiles = max(1, min(coltiles, cdiv(m, kOnceChunk)));
  const int cols_per_tile = cdiv(cdiv(m, coltiles), 4) * 4;
  coltiles = cdiv(m, cols_per_tile);
  PCFM_CUDA(cudaMemsetAsync(colkey, 0xff, (size_t)b * m * 8, st));
  if (coltiles > 1) PCFM_CUDA(cudaMemsetAsync(rowkey, 0xff, (size_t)b * n * 8, st));
  dim3 grid(rowtiles, coltiles, b);
  if (ra == 16) chamfer_once_kernel<16><<<grid, kOnceThreads, 0, st>>>(xyz1, xyz2, n, m, cols_per_tile, rowkey, colkey);
  else if (ra == 8) chamfer_once_kernel<8><<<grid, kOnceThreads, 0, st>>>(xyz1, xyz2, n, m, cols_per_tile, rowkey, colkey);
  else chamfer_once_kernel<4><<<grid, kOnceThreads, 0, st>>>(xyz1, xyz2, n, m, cols_per_tile, rowkey, colkey);
  PCFM_LAUNCHED();
  dim3 g2(cdiv(n > m ? n : m, 256), b);
  chamfer_once_finish_kernel<<<g2, 256, 0, st>>>(xyz1, xyz2, n, m, ra, rowkey, colkey, dist1, idx1, dist2, idx2);
  PCFM_LAUNCHED();
  return 0;
}

int pcfm_chamfer3d_backward(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                            const int* idx1, const int* idx2, int b, int n, int m, float* gradxyz1,
                            float* gradxyz2, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;
  if (!xyz1 || !xyz2 || !graddist1 || !graddist2 || !idx1 || !idx2 || !gradxyz1 || !gradxyz2)
    return PCFM_E_BADARG;
  dim3 grid(cdiv(n > m ? n : m, 256), b);
  chamfer_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, n,
                                                              m, gradxyz1, gradxyz2);
  PCFM_LAUNCHED();
  return 0;
}

int pcfm_chamfer3d_allpairs(const float* s, const float* r, int ns, int nr, int n, int m, float* cd,
                            pcfm_stream_t stream) {
  if (ns < 0 || nr < 0 || n <= 0 || m <= 0 || m > kApMaxM) return PCFM_E_BADARG;
  if ((long long)ns * nr > 0x7fffffffLL) return PCFM_E_BADARG;
  if (ns == 0 || nr == 0) return 0;
  if (!s || !r || !cd) return PCFM_E_BADARG;
  PCFM_CUDA(allow_dyn_smem((const void*)chamfer_allpairs_kernel, (kApMaxM + 2) * 20));
  const size_t smem = (size_t)((m + 1) & ~1) * 20;
  chamfer_allpairs_kernel<<<(unsigned)((long long)ns * nr), kApThreads, smem, (cudaStream_t)stream>>>(
      s, r, nr, n, m, cd);
  PCFM_LAUNCHED();
  return 0;
}

}  // extern "C"
