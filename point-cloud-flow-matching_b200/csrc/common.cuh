// common.cuh — shared device/host helpers for libpcfm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <stddef.h>

#include "../../include/pcfm_b200.h"

#define PCFM_STR2(x) #x
#define PCFM_STR(x) PCFM_STR2(x)

namespace pcfm {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
extern std::atomic<int64_t> g_launches;  // kernels launched by this library (pcfm_launch_count); the forward thread and
                                         // autograd's device threads both launch
extern int g_opt_devox_bwd;     // see pcfm_set_option
extern int g_opt_vox_fwd;
extern int g_opt_stage;
extern int g_opt_devox_il;      // 1 = channel-interleaved planes in devoxelize forward when G >= 4 (default)
extern int g_opt_devox_cl;      // devoxelize forward on small grids: 0 = point-lane kernels (default: measured faster),
                                // 1 = channel-lane kernel for r <= 8, 2 = up to r = 10 (one CTA per SM)
extern int g_opt_pdl;           // programmatic dependent launch: 0 = off, 1 = inside an operator (row kernel behind its
                                // plan / index kernel), 2 = also between operators (every PVConv kernel waits at its
                                // top, then releases the kernel behind it: launch + scheduling overlap the previous tail)
extern int g_opt_fps;           // 0 = cluster FPS (points in registers) when the cloud fits, 1 = one CTA per cloud
extern int g_opt_gather_bwd;    // 0 = gather backward with atomics (like the reference), 1 = sorted-segment sums
extern int g_opt_grouping_bwd;  // 0 = grouping backward as a scatter into a shared-memory row when it fits (default), 1 = sorted-segment sums
extern int g_opt_three_nn_bwd;  // the same choice for the 3-NN interpolate backward
extern int g_opt_devox_fwd;     // 0 = devoxelize forward as a walk over the trilinear cell plan when the caller offers a
                                // workspace (default), 1 = always the plane-staging point-lane kernels
extern int g_opt_bwd_phases;    // devoxelize backward, how light cells reach the plane: 0 = shared-memory atomics (default),
                                // 2 = deferred conflict-free corner phases: fixed summation order, not faster (tri.cu)
extern int g_opt_chamfer_once;  // pcfm_chamfer3d_forward_ws: 1 = single pass from ~3e8 distance evaluations on (default), 2 = always,
                                // 0 = never (pcfm/backend.py then calls the two-pass entry)
extern int g_opt_emd_cluster;   // approxmatch operator: 1 = cluster of 8 CTAs per batch element (default), 0 = one CTA
extern int g_opt_plan;          // 0 = fused cell-plan build when it fits, 1 = always the multi-kernel build
extern int g_num_sms;           // cached SM count of the current device (148 on B200)
int num_sms();
// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes) once per (kernel, device): function attributes are
// per device, so a process that drives several GPUs must set them on each one
cudaError_t allow_dyn_smem(const void* func, int bytes);

#define PCFM_LAUNCHED()                                   \
  do {                                                    \
    ::pcfm::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return (int)e__;              \
  } while (0)

#define PCFM_CUDA(x)                                      \
  do {                                                    \
    cudaError_t e__ = (x);                                \
    if (e__ != cudaSuccess) return (int)e__;              \
  } while (0)

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// <<<grid, block, smem, st>>> launch that, when `dependent`, lets the kernel overlap the tail of the previous kernel on
// the stream (see griddep_wait() below)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           bool dependent, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = dependent ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t cap, off;
  // blocks are handed out on 256-byte boundaries whatever the alignment of the caller's buffer (the *_workspace()
  // sizes include that slack)
  Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {
    const size_t mis = (size_t)((uintptr_t)p & 255);
    if (p && mis) {
      const size_t skip = 256 - mis;
      if (skip >= bytes) { base = nullptr; cap = 0; }
      else { base += skip; cap -= skip; }
    }
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align256(count * sizeof(T));
    if (off + bytes > cap || base == nullptr) { off = cap + 1; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
// The reference's `dx*dx+dy*dy+dz*dz` is contracted by nvcc (default -fmad=true) into
// FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)  (SASS of chamfer3D.cu:35, ball_query.cu:38, sampling.cu:134,
// neighbor_interpolate.cu:43).  Written with intrinsics so the order is fixed whatever the compiler does.
__device__ __forceinline__ float sq3(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D TMA bulk copies (cp.async.bulk; SASS: UBLKCP) -----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a bulk store / before TMA overwrites)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a byte-count bug must surface as a trap (error code), never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                   __cvta_generic_to_global(gmem_dst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Stage `count` floats from global into shared memory.  Called by ALL threads of the CTA.
// Uses one elected thread issuing 1-D TMA bulk copies when src/dst are 16-byte aligned (tail by plain loads),
// otherwise a plain coalesced copy.  `bar` is an initialised mbarrier (count 1), `phase` its current parity
// (flipped by the caller after each use).  Ends with a __syncthreads(): data visible to every thread.
constexpr uint32_t kBulkChunk = 32768;  // bytes per cp.async.bulk
__device__ __forceinline__ void stage_floats(float* smem_dst, const float* gmem_src, int count, uint64_t* bar,
                                             uint32_t phase, bool use_tma) {
  const bool aligned = use_tma && ((((uintptr_t)gmem_src) & 15) == 0) && ((smem_u32(smem_dst) & 15) == 0);
  int bulk_count = aligned ? (count & ~3) : 0;
  if (bulk_count > 0) {
    if (threadIdx.x == 0) {
      uint32_t bytes = (uint32_t)bulk_count * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      for (uint32_t off = 0; off < bytes; off += kBulkChunk) {
        uint32_t sz = min(kBulkChunk, bytes - off);
        bulk_g2s((char*)smem_dst + off, (const char*)gmem_src + off, sz, bar);
      }
    }
  }
  for (int i = bulk_count + threadIdx.x; i < count; i += blockDim.x) smem_dst[i] = __ldg(gmem_src + i);
  if (bulk_count > 0) mbar_wait(bar, phase);
  __syncthreads();
}

// ---- programmatic dependent launch (PDL): a kernel launched with launch_dependent() may start while the previous
// kernel on the stream is still running once that kernel has called griddep_launch_dependents(); it must call
// griddep_wait() before touching anything the previous kernel writes (the wait returns when that grid has completed
// and its memory is visible).  Both are no-ops for kernels launched normally.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Pointers to data the PREVIOUS grid writes (the cell plan) must be passed through this after griddep_wait(): loads
// through `const __restrict__` pointers are invariant loads (ld.global.nc) which the compiler may hoist above the wait —
// observed: the plan's `meta` load of a persistent kernel's first job moved in front of the wait and read the plan of
// the previous call.  The empty asm makes the pointer value (and so every load through it) depend on this program point.
template <typename T>
__device__ __forceinline__ const T* after_wait(const T* p) {
  asm volatile("" : "+l"(p) : : "memory");
  return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pcfm
