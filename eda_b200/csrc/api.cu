// Version / error-string entry points of the C ABI (include/eda_b200.h).
#include <string.h>
#include <stdio.h>
#include "common.cuh"

namespace eda {
static thread_local char g_last_err[256] = "";
void set_last_cuda_error(cudaError_t e, const char *where) {
  snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int v = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
  if (v == 0) {
    v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    __atomic_store_n(&cached[dev], v, __ATOMIC_RELAXED);
  }
  return v;
}
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
}  // namespace eda

extern "C" {

unsigned long long eda_launch_count(void) { return __atomic_load_n(&eda::g_launches, __ATOMIC_RELAXED); }

int eda_version(void) { return 100; }

const char *eda_error_string(int code) {
  switch (code) {
    case EDA_OK: return "ok";
    case EDA_ERR_INVALID_ARGUMENT: return "invalid argument";
    case EDA_ERR_CUDA_LAUNCH: return "CUDA launch/runtime error (see eda_last_cuda_error)";
    case EDA_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case EDA_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error code";
  }
}

const char *eda_last_cuda_error(void) { return eda::g_last_err; }

}  // extern "C"
