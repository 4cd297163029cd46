// chamfer_nd.cu — nearest-neighbour (Chamfer) distance for D-dimensional points, D in {2, 5, 6}:
// third_party/ChamferDistancePytorch/chamfer{2D,5D,6D}/chamfer{2,5,6}D.cu (same algorithm as chamfer3D.cu with a
// wider point; 6D = xyz + rgb / normals).  A "next" row of the scope table.
//
//   dist1[b,i] = min_j |x1_i - x2_j|^2   idx1[b,i] = the smallest j that attains it     (and symmetric)
//
// The squared distance is evaluated in the reference's contracted order (SASS of chamfer6D.cu:32-38:
// FMUL(d1,d1); FFMA(d0,d0,.); FFMA(d2,d2,.); ... FFMA(d5,d5,.)), so distances and arg-mins are bit-identical.
// One thread owns RA points (D registers each); the other cloud is streamed through shared memory in tiles and
// broadcast to the whole CTA.
#include "common.cuh"

namespace pcfm {

constexpr int kNdThreads = 256;
constexpr int kNdTile = 1024;  // points of the other cloud per shared-memory tile (24 KB at D = 6)

template <int D>
__device__ __forceinline__ float sqd(const float (&a)[D], const float* __restrict__ o) {
  // the reference subtracts own from other: (o - a); then y*y, fma(x,x,.), fma(z,z,.), ...
  float acc = __fmul_rn(o[1] - a[1], o[1] - a[1]);
  acc = __fmaf_rn(o[0] - a[0], o[0] - a[0], acc);
#pragma unroll
  for (int k = 2; k < D; ++k) acc = __fmaf_rn(o[k] - a[k], o[k] - a[k], acc);
  return acc;
}

// packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each lane is rounded exactly like the scalar instruction
typedef unsigned long long nd_f32x2;
__device__ __forceinline__ nd_f32x2 nd_pack2(float lo, float hi) {
  nd_f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ nd_f32x2 nd_sub2(nd_f32x2 a, nd_f32x2 b) {
  nd_f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ nd_f32x2 nd_sq(nd_f32x2 d) {
  nd_f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(d));
  return r;
}
__device__ __forceinline__ nd_f32x2 nd_sqacc(nd_f32x2 d, nd_f32x2 acc) {
  nd_f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(d), "l"(acc));
  return r;
}

// The other cloud sits in shared memory as D coordinate rows (SoA), so one LDS.128 per coordinate yields two register
// PAIRS of columns and sqd() runs on packed fp32: two columns per instruction.  The running minimum is kept per group of
// four columns (strict <: the first group that reaches the minimum wins) and resolved to its lowest column at the end by
// recomputing the four (bit-identical) distances — the scheme of chamfer_nn_kernel (chamfer.cu).
template <int D, int RA>
__global__ void __launch_bounds__(kNdThreads)
    chamfer_nd_kernel(const float* __restrict__ own, const float* __restrict__ other, int n, int m,
                      float* __restrict__ dist, int* __restrict__ idx) {
  __shared__ __align__(16) float tile[D][kNdTile + 4];  // + 4: the D rows of one column fall into different banks (transposing fill)
  const int b = blockIdx.y;
  const float* pa = own + (long long)b * n * D;
  const float* po = other + (long long)b * m * D;
  float a[RA][D], best[RA];
  nd_f32x2 ap[RA][D];
  int bk[RA];
  int row[RA];
#pragma unroll
  for (int r = 0; r < RA; ++r) {
    row[r] = (blockIdx.x * RA + r) * kNdThreads + threadIdx.x;
    const int rr = min(row[r], n - 1);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      a[r][k] = pa[(long long)rr * D + k];
      ap[r][k] = nd_pack2(a[r][k], a[r][k]);
    }
    best[r] = __int_as_float(0x7f800000);  // +inf: the first candidate always wins, like `k==0 ||` in the reference
    bk[r] = 0;
  }
  for (int t0 = 0; t0 < m; t0 += kNdTile) {
    const int tn = min(kNdTile, m - t0), tn4 = (tn + 3) & ~3;
    __syncthreads();
    for (int i = threadIdx.x; i < tn4 * D; i += kNdThreads) {   // coalesced AoS read, transposed into the rows
      const int jj = i / D, k = i - jj * D;
      tile[k][jj] = jj < tn ? po[(long long)t0 * D + i] : __int_as_float(0x7f800000);  // padding: distance +inf
    }
    __syncthreads();
    for (int j = 0; j < tn4; j += 4) {
      ulonglong2 q[D];
#pragma unroll
      for (int k = 0; k < D; ++k) q[k] = *reinterpret_cast<const ulonglong2*>(&tile[k][j]);  // (o_j, o_j+1) (o_j+2, o_j+3)
#pragma unroll
      for (int r = 0; r < RA; ++r) {
        // the reference subtracts own from other, then y*y, fma(x,x,.), fma(z,z,.), ...  (sqd() above)
        nd_f32x2 acc0 = nd_sq(nd_sub2(q[1].x, ap[r][1])), acc1 = nd_sq(nd_sub2(q[1].y, ap[r][1]));
        acc0 = nd_sqacc(nd_sub2(q[0].x, ap[r][0]), acc0);
        acc1 = nd_sqacc(nd_sub2(q[0].y, ap[r][0]), acc1);
#pragma unroll
        for (int k = 2; k < D; ++k) {
          acc0 = nd_sqacc(nd_sub2(q[k].x, ap[r][k]), acc0);
          acc1 = nd_sqacc(nd_sub2(q[k].y, ap[r][k]), acc1);
        }
        float d0, d1, d2, d3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(acc0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d2), "=f"(d3) : "l"(acc1));
        const float m4 = fminf(fminf(d0, d1), fminf(d2, d3));
        if (m4 < best[r]) {  // strict: the lowest group keeps a tie (chamfer6D.cu:39-42 and the chunk merge :160-164)
          best[r] = m4;
          bk[r] = t0 + j;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RA; ++r)
    if (row[r] < n) {
      int bi = bk[r];
      for (int t = 0; t < 4; ++t) {   // lowest column inside the winning group of four
        const int k = bk[r] + t;
        if (k >= m) break;
        if (sqd<D>(a[r], po + (long long)k * D) == best[r]) {
          bi = k;
          break;
        }
      }
      dist[(long long)b * n + row[r]] = best[r];
      idx[(long long)b * n + row[r]] = bi;
    }
}

// gradient (chamfer6D.cu:185-216, both launches fused): accumulates into the caller-zeroed grad1 / grad2
template <int D>
__global__ void __launch_bounds__(256)
    chamfer_nd_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                           const float* __restrict__ gd1, const float* __restrict__ gd2,
                           const int* __restrict__ idx1, const int* __restrict__ idx2, int n, int m,
                           float* __restrict__ g1, float* __restrict__ g2) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float* p1 = xyz1 + (long long)b * n * D;
  const float* p2 = xyz2 + (long long)b * m * D;
  float* o1 = g1 + (long long)b * n * D;
  float* o2 = g2 + (long long)b * m * D;
  if (j < n) {
    const int j2 = min(max(idx1[(long long)b * n + j], 0), m - 1);
    const float g = gd1[(long long)b * n + j] * 2;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float v = __fmul_rn(g, p1[(long long)j * D + k] - p2[(long long)j2 * D + k]);
      atomicAdd(o1 + (long long)j * D + k, v);
      atomicAdd(o2 + (long long)j2 * D + k, -v);
    }
  }
  if (j < m) {
    const int j1 = min(max(idx2[(long long)b * m + j], 0), n - 1);
    const float g = gd2[(long long)b * m + j] * 2;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float v = __fmul_rn(g, p2[(long long)j * D + k] - p1[(long long)j1 * D + k]);
      atomicAdd(o2 + (long long)j * D + k, v);
      atomicAdd(o1 + (long long)j1 * D + k, -v);
    }
  }
}

template <int D>
static int chamfer_nd_forward(const float* xyz1, const float* xyz2, int b, int n, int m, float* dist1, float* dist2,
                              int* idx1, int* idx2, cudaStream_t st) {
  for (int dir = 0; dir < 2; ++dir) {
    const float* a = dir ? xyz2 : xyz1;
    const float* o = dir ? xyz1 : xyz2;
    const int na = dir ? m : n, no = dir ? n : m;
    float* d = dir ? dist2 : dist1;
    int* ix = dir ? idx2 : idx1;
    // two own points per thread once that still leaves two CTAs per SM
    if ((long long)na * b >= 4LL * num_sms() * kNdThreads) {
      dim3 grid(cdiv(na, kNdThreads * 2), b);
      chamfer_nd_kernel<D, 2><<<grid, kNdThreads, 0, st>>>(a, o, na, no, d, ix);
    } else {
      dim3 grid(cdiv(na, kNdThreads), b);
      chamfer_nd_kernel<D, 1><<<grid, kNdThreads, 0, st>>>(a, o, na, no, d, ix);
    }
    PCFM_LAUNCHED();
  }
  return 0;
}

template <int D>
static int chamfer_nd_backward(const float* xyz1, const float* xyz2, const float* gd1, const float* gd2,
                               const int* idx1, const int* idx2, int b, int n, int m, float* g1, float* g2,
                               cudaStream_t st) {
  dim3 grid(cdiv(n > m ? n : m, 256), b);
  chamfer_nd_grad_kernel<D><<<grid, 256, 0, st>>>(xyz1, xyz2, gd1, gd2, idx1, idx2, n, m, g1, g2);
  PCFM_LAUNCHED();
  return 0;
}

}  // namespace pcfm

using namespace pcfm;

extern "C" {

int pcfm_chamfer_nd_forward(const float* xyz1, const float* xyz2, int b, int n, int m, int d, float* dist1,
                            float* dist2, int* idx1, int* idx2, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || (d != 2 && d != 5 && d != 6)) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;  // the reference's loops do not execute: outputs untouched
  if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: return chamfer_nd_forward<2>(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, st);
    case 5: return chamfer_nd_forward<5>(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, st);
    default: return chamfer_nd_forward<6>(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, st);
  }
}

int pcfm_chamfer_nd_backward(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                             const int* idx1, const int* idx2, int b, int n, int m, int d, float* gradxyz1,
                             float* gradxyz2, pcfm_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || (d != 2 && d != 5 && d != 6)) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;
  if (!xyz1 || !xyz2 || !graddist1 || !graddist2 || !idx1 || !idx2 || !gradxyz1 || !gradxyz2)
    return PCFM_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: return chamfer_nd_backward<2>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, b, n, m, gradxyz1, gradxyz2, st);
    case 5: return chamfer_nd_backward<5>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, b, n, m, gradxyz1, gradxyz2, st);
    default: return chamfer_nd_backward<6>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, b, n, m, gradxyz1, gradxyz2, st);
  }
}

}  // extern "C"
