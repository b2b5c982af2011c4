// Error reporting and version for the C ABI (include/cast_b200.h).
#include "cast_rt.cuh"
#include <stdio.h>
#include <string.h>

namespace cast {

static thread_local char g_err[256] = "";
static unsigned long long g_launches = 0;  // kernels enqueued through this library (host-side counter)

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s (code %d)", msg ? msg : "error", code);
  return code;
}

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d: %s", what, (int)e, cudaGetErrorString(e));
    return CAST_ERR_CUDA;
  }
  return CAST_OK;
}

}  // namespace cast

extern "C" int cast_version(void) { return CAST_ABI_VERSION; }
extern "C" const char* cast_last_error_string(void) { return cast::g_err; }
extern "C" unsigned long long cast_launch_count(void) { return __atomic_load_n(&cast::g_launches, __ATOMIC_RELAXED); }
