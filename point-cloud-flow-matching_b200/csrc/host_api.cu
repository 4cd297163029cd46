// host_api.cu — host-buffer convenience entry points (allocate, copy in, run, copy out, free).
#include "common.cuh"

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { return (int)cudaMalloc(&p, bytes ? bytes : 1); }
};
}  // namespace

extern "C" {

int pcfm_chamfer3d_forward_host(const float* xyz1_host, const float* xyz2_host, int b, int n, int m,
                                float* dist1_host, float* dist2_host, int* idx1_host, int* idx2_host) {
  if (b < 0 || n < 0 || m < 0) return PCFM_E_BADARG;
  if (b == 0 || n == 0 || m == 0) return 0;
  if (!xyz1_host || !xyz2_host || !dist1_host || !dist2_host || !idx1_host || !idx2_host) return PCFM_E_BADARG;
  const size_t s1 = (size_t)b * n, s2 = (size_t)b * m;
  DevBuf x1, x2, d1, d2, i1, i2;
  PCFM_CUDA((cudaError_t)x1.alloc(s1 * 12));
  PCFM_CUDA((cudaError_t)x2.alloc(s2 * 12));
  PCFM_CUDA((cudaError_t)d1.alloc(s1 * 4));
  PCFM_CUDA((cudaError_t)d2.alloc(s2 * 4));
  PCFM_CUDA((cudaError_t)i1.alloc(s1 * 4));
  PCFM_CUDA((cudaError_t)i2.alloc(s2 * 4));
  PCFM_CUDA(cudaMemcpy(x1.p, xyz1_host, s1 * 12, cudaMemcpyHostToDevice));
  PCFM_CUDA(cudaMemcpy(x2.p, xyz2_host, s2 * 12, cudaMemcpyHostToDevice));
  int rc = pcfm_chamfer3d_forward((const float*)x1.p, (const float*)x2.p, b, n, m, (float*)d1.p, (float*)d2.p,
                                  (int*)i1.p, (int*)i2.p, nullptr);
  if (rc) return rc;
  PCFM_CUDA(cudaMemcpy(dist1_host, d1.p, s1 * 4, cudaMemcpyDeviceToHost));
  PCFM_CUDA(cudaMemcpy(dist2_host, d2.p, s2 * 4, cudaMemcpyDeviceToHost));
  PCFM_CUDA(cudaMemcpy(idx1_host, i1.p, s1 * 4, cudaMemcpyDeviceToHost));
  PCFM_CUDA(cudaMemcpy(idx2_host, i2.p, s2 * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int pcfm_avg_voxelize_forward_host(const float* features_host, const int* coords_host, int b, int c, int n,
                                   int r, float* out_host, int* ind_host, int* cnt_host) {
  if (b < 0 || c < 0 || n < 0 || r <= 0) return PCFM_E_BADARG;
  if (b == 0) return 0;
  if (!features_host || !coords_host || !out_host || !ind_host || !cnt_host) return PCFM_E_BADARG;
  const size_t s = (size_t)r * r * r;
  const size_t wsb = pcfm_avg_voxelize_forward_workspace(b, c, n, r);
  DevBuf f, co, out, ind, cnt, ws;
  PCFM_CUDA((cudaError_t)f.alloc((size_t)b * c * n * 4));
  PCFM_CUDA((cudaError_t)co.alloc((size_t)b * 3 * n * 4));
  PCFM_CUDA((cudaError_t)out.alloc((size_t)b * c * s * 4));
  PCFM_CUDA((cudaError_t)ind.alloc((size_t)b * n * 4));
  PCFM_CUDA((cudaError_t)cnt.alloc((size_t)b * s * 4));
  PCFM_CUDA((cudaError_t)ws.alloc(wsb));
  PCFM_CUDA(cudaMemcpy(f.p, features_host, (size_t)b * c * n * 4, cudaMemcpyHostToDevice));
  PCFM_CUDA(cudaMemcpy(co.p, coords_host, (size_t)b * 3 * n * 4, cudaMemcpyHostToDevice));
  int rc = pcfm_avg_voxelize_forward((const float*)f.p, (const int*)co.p, b, c, n, r, (float*)out.p,
                                     (int*)ind.p, (int*)cnt.p, ws.p, wsb, nullptr);
  if (rc) return rc;
  PCFM_CUDA(cudaMemcpy(out_host, out.p, (size_t)b * c * s * 4, cudaMemcpyDeviceToHost));
  PCFM_CUDA(cudaMemcpy(ind_host, ind.p, (size_t)b * n * 4, cudaMemcpyDeviceToHost));
  PCFM_CUDA(cudaMemcpy(cnt_host, cnt.p, (size_t)b * s * 4, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
