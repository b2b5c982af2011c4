// Error reporting and version for the C ABI (include/cast_b200.h).
#include "cast_rt.cuh"
#include <stdio.h>
#include <string.h>

namespace cast {

static thread_local char g_err[256] = "";
static unsigned long long g_launches = 0;  // kernels enqueued through this library (host-side counter)
#ifndef CAST_EMU
int g_pdl = 0;   // programmatic dependent launch of the step's main-chain kernels (cast_rt.cuh); cast_set_pdl(1) by the engine
#endif

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
extern "C" int cast_set_pdl(int on) {
#ifndef CAST_EMU
  cast::g_pdl = on ? 1 : 0;
#else
  (void)on;
#endif
  return CAST_OK;
}
extern "C" const char* cast_last_error_string(void) { return cast::g_err; }
extern "C" unsigned long long cast_launch_count(void) { return __atomic_load_n(&cast::g_launches, __ATOMIC_RELAXED); }

/* crc32c (Castagnoli, reflected 0x82F63B78) of a HOST buffer, slicing-by-8 — the per-tensor checksum of TensorFlow's
 * tensor-bundle format (checkpoint.py); host-side utility, no device work. */
extern "C" unsigned int cast_crc32c(const void* data, size_t n, unsigned int crc_in) {
  static unsigned int T[8][256];
  static bool init = false;
  if (!init) {
    for (unsigned i = 0; i < 256; ++i) {
      unsigned c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      T[0][i] = c;
    }
    for (unsigned i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xff];
    init = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  unsigned int crc = ~crc_in;
  while (n >= 8) {
    unsigned int lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = T[7][lo & 0xff] ^ T[6][(lo >> 8) & 0xff] ^ T[5][(lo >> 16) & 0xff] ^ T[4][lo >> 24] ^ T[3][hi & 0xff] ^
          T[2][(hi >> 8) & 0xff] ^ T[1][(hi >> 16) & 0xff] ^ T[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = T[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return ~crc;
}
