// runtime.cu — library globals: options, launch counter, build info.
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace pcfm {
std::atomic<int64_t> g_launches{0};
int g_opt_devox_bwd = 0;
int g_opt_vox_fwd = 0;
int g_opt_stage = 0;
int g_opt_plan = 0;
int g_opt_devox_fwd = 0;
int g_opt_bwd_phases = 0;
int g_opt_emd_cluster = 1;
int g_opt_chamfer_once = 1;
int g_opt_gather_bwd = 0;
int g_opt_grouping_bwd = 0;
int g_opt_three_nn_bwd = 0;
int g_opt_fps = 0;
int g_opt_pdl = 2;
int g_opt_devox_il = 1;
int g_opt_devox_cl = 0;
int g_num_sms = 0;

int num_sms() {
  if (g_num_sms <= 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}
cudaError_t allow_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;  // (kernel, device) -> bytes already allowed
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  int& have = done[std::make_pair(func, dev)];
  if (have >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}
}  // namespace pcfm

extern "C" {

int pcfm_abi_version(void) { return PCFM_ABI_VERSION; }

const char* pcfm_build_info(void) {
  return "pcfm_b200 abi1 sm_100a nvcc " PCFM_STR(__CUDACC_VER_MAJOR__) "." PCFM_STR(__CUDACC_VER_MINOR__);
}

int64_t pcfm_launch_count(void) { return pcfm::g_launches.load(std::memory_order_relaxed); }
void pcfm_reset_launch_count(void) { pcfm::g_launches.store(0, std::memory_order_relaxed); }

unsigned long long pcfm_stream_capture_id(pcfm_stream_t stream) {
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo((cudaStream_t)stream, &status, &id) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return status == cudaStreamCaptureStatusActive ? id : 0;
}

int pcfm_set_option(const char* key, int value) {
  if (!key) return PCFM_E_BADARG;
  if (!strcmp(key, "devox_bwd")) { pcfm::g_opt_devox_bwd = value; return 0; }
  if (!strcmp(key, "vox_fwd")) { pcfm::g_opt_vox_fwd = value; return 0; }
  if (!strcmp(key, "stage")) { pcfm::g_opt_stage = value; return 0; }
  if (!strcmp(key, "devox_il")) { pcfm::g_opt_devox_il = value; return 0; }
  if (!strcmp(key, "devox_cl")) { pcfm::g_opt_devox_cl = value; return 0; }
  if (!strcmp(key, "pdl")) { pcfm::g_opt_pdl = value; return 0; }
  if (!strcmp(key, "fps")) { pcfm::g_opt_fps = value; return 0; }
  if (!strcmp(key, "gather_bwd")) { pcfm::g_opt_gather_bwd = value; return 0; }
  if (!strcmp(key, "grouping_bwd")) { pcfm::g_opt_grouping_bwd = value; return 0; }
  if (!strcmp(key, "three_nn_bwd")) { pcfm::g_opt_three_nn_bwd = value; return 0; }
  if (!strcmp(key, "plan")) { pcfm::g_opt_plan = value; return 0; }
  if (!strcmp(key, "devox_fwd")) { pcfm::g_opt_devox_fwd = value; return 0; }
  if (!strcmp(key, "bwd_phases")) { pcfm::g_opt_bwd_phases = value; return 0; }
  if (!strcmp(key, "emd_cluster")) { pcfm::g_opt_emd_cluster = value; return 0; }
  if (!strcmp(key, "chamfer_once")) { pcfm::g_opt_chamfer_once = value; return 0; }
  return PCFM_E_BADARG;
}

}  // extern "C"
