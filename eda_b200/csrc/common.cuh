// Shared helpers for the eda_b200 sm_100a kernels: error plumbing, PTX wrappers for
// clusters / DSMEM / mbarrier / bulk (TMA) copies.  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include "../../include/eda_b200.h"

namespace eda {

void set_last_cuda_error(cudaError_t e, const char *where);
void count_launches(int n);  // bookkeeping for eda_launch_count() (bench.py's gpu_launches)

// Call after every launch: records the error text and maps it to an ABI code (never exits,
// unlike the reference's CUDA_CHECK_ERRORS(), pointnet2/_ext_src/include/cuda_utils.h:35-44).
inline int check_launch(const char *where, int kernels = 1) {
  count_launches(kernels);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e, where);
    return EDA_ERR_CUDA_LAUNCH;
  }
  return EDA_OK;
}

#define EDA_CUDA_TRY(expr, where)                       \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) {                            \
      ::eda::set_last_cuda_error(_e, where);            \
      return EDA_ERR_CUDA_LAUNCH;                       \
    }                                                   \
  } while (0)

// The reference computes a*a + b*b + c*c as FMUL(b,b); FFMA(a,a,.); FFMA(c,c,.) (SASS of its
// sm_100a build).  Intrinsics pin that order so nvcc cannot re-associate.
__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Map a local shared-memory address to the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Same wait, but with acquire at cluster scope: the data was written by a peer CTA (st.async).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// One-sided 16-byte / 4-byte stores into a peer CTA's shared memory that complete_tx on the
// peer's mbarrier (DSMEM; no cluster-wide barrier needed on the critical path).
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t a, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(a), "r"(remote_bar)
               : "memory");
}

// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP); completes on `bar`.
// Requires 16-byte aligned src/dst and a byte count that is a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Counter-based dropout decision shared by the linear / attention kernels and eda_dropout_mask (the backward
// pass regenerates exactly the mask the forward kernel applied): element (a, b) of a call seeded `seed` is KEPT
// iff the top 24 bits of a 32-bit mix of (seed, a, b) are >= thresh = round(p * 2^24).
__host__ __device__ __forceinline__ uint32_t dropout_mix(uint32_t seed, uint32_t a, uint32_t b) {
  uint32_t h = seed ^ (a * 0x9E3779B1u);
  h ^= h >> 15; h *= 0x85EBCA77u;
  h ^= b * 0xC2B2AE3Du;
  h ^= h >> 13; h *= 0x27D4EB2Fu;
  h ^= h >> 16; h *= 0x165667B1u;
  h ^= h >> 15;
  return h;
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t a, uint32_t b, uint32_t thresh) {
  return (dropout_mix(seed, a, b) >> 8) >= thresh;
}
inline uint32_t dropout_thresh(float p) {
  if (!(p > 0.f)) return 0u;
  const double t = (double)p * 16777216.0;
  return t >= 16777216.0 ? 16777216u : (uint32_t)(t + 0.5);
}

// Optional device word added to every dropout seed at kernel run time (the `dropout_epoch` argument of every
// dropout-applying entry point): lets a CUDA graph of a training step draw fresh masks on every replay although the
// per-call seeds are frozen into the graph.
__device__ __forceinline__ uint32_t effective_seed(uint32_t seed, const uint32_t *epoch) {
  return epoch ? seed + __ldg(epoch) : seed;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function attribute: this records, per device, the
// largest value already applied for ONE kernel (one static instance per kernel / template instance) so the launch path
// pays an atomic load instead of a driver call; raising it is serialised by a mutex (threads, several devices in one
// process).  Devices beyond kMaxDev simply set the attribute on every launch.
struct SmemAttr {
  static constexpr int kMaxDev = 64;
  std::atomic<size_t> applied[kMaxDev];
  std::mutex mu;
  template <typename F>
  cudaError_t ensure(F *fn, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < kMaxDev;
    if (tracked && bytes <= applied[dev].load(std::memory_order_acquire)) return cudaSuccess;
    std::lock_guard<std::mutex> lock(mu);
    if (tracked && bytes <= applied[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && tracked) applied[dev].store(bytes, std::memory_order_release);
    return e;
  }
};

// Programmatic dependent launch (PDL): a kernel launched with the ProgrammaticStreamSerialization attribute may be
// scheduled while its predecessor in the stream is still running; `pdl_wait()` blocks until the predecessor grid has
// completed and its memory operations are visible (griddepcontrol.wait), `pdl_launch_dependents()` lets the successor
// be scheduled from now on.  With the wait placed after a kernel's private set-up (TMEM allocation, barrier
// initialisation, parameter loads) the launch latency and that set-up overlap the predecessor's tail — in a step of
// ~800 short kernels that is a measurable share.  Both are no-ops when the kernel was launched normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// EDA_PDL=0 switches the launch attribute off (debugging).
inline bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("EDA_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// Launch with the PDL attribute (see above).  The kernel must call pdl_wait() before it reads anything a predecessor in
// the stream may have written.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SM count of the current device (cached per device: cudaDeviceGetAttribute costs ~1 us per call on the launch path).
int sm_count();

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace eda
