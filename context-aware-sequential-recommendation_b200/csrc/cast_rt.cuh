// Runtime glue shared by every kernel file: launch macro, error codes, warp reductions, the counter-based
// dropout generator.  Compiled by nvcc for sm_100a.  (tests/emu builds the same sources with -DCAST_EMU on
// the host to check indexing logic without a GPU; that build is test infrastructure and is never loaded by
// the package.)
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef CAST_EMU
#include "cuda_emu.h"
#define CAST_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cast_emu::launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); })
#define CAST_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cast_emu::dyn_smem_ptr())
#else
#include <cuda_runtime.h>
#define CAST_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define CAST_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char name##_raw_smem[]; \
  type* name = reinterpret_cast<type*>(name##_raw_smem)
#endif

// Programmatic dependent launch for the kernels of the training step's main chain: the next kernel's CTAs may be
// scheduled while the previous kernel drains (its launch latency and whatever it does before cast_pdl_wait() overlap
// the tail); every kernel launched this way calls cast_pdl_wait() before its first global-memory access, so the
// stream order of all data is unchanged.  cast_set_pdl(0) (env CAST_PDL=0) launches them the ordinary way.
#ifdef CAST_EMU
#define CAST_LAUNCH_DEP(kernel, grid, block, smem, stream, ...) CAST_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
static inline void cast_pdl_wait() {}
static inline void cast_pdl_trigger() {}
#else
namespace cast {
extern int g_pdl;
template <typename... KArgs, typename... Args>
static inline void launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  if (!g_pdl) {
    kernel<<<grid, block, smem, stream>>>(KArgs(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace cast
#define CAST_LAUNCH_DEP(kernel, grid, block, smem, stream, ...) \
  cast::launch_dep(kernel, grid, block, smem, stream, __VA_ARGS__)
__device__ __forceinline__ void cast_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cast_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#include "../../include/cast_b200.h"

#define CAST_NEG_FILL (-4294967296.0f)  // float32(-2**32+1), modules.py:227,239

namespace cast {

int set_error(int code, const char* msg);
int check_launch(const char* what);

static inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// The item table as the gather kernels see it: one [V, H] array, or n row shards owned by the n ranks of the box
// (SURVEY §8e "row-sharded item table": cyclic ownership, id -> shard id % n, local row id / n, +1 on shards 1..n-1
// whose local row 0 is a never-referenced pad so that EVERY shard keeps "row 0 is not an item").  `shards` is a device
// array of device pointers; entries of other ranks are peer mappings read over NVLink.
struct TableRef {
  const float* base;
  const float* const* shards;
  int n;
  __device__ __forceinline__ const float* row(int id, int H) const {
    if (!shards) return base + (long)id * H;
    const int o = id % n;
    return shards[o] + (long)(id / n + (o ? 1 : 0)) * H;
  }
};
static inline TableRef table_ref(const float* base) { return TableRef{base, nullptr, 1}; }
static inline TableRef table_ref(const float* const* shards, int n) { return TableRef{nullptr, shards, n}; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- counter-based dropout ------------------------------------------------------------------------------
// keep(site, idx) = field(idx & 1) of hash(seed, step, site, idx >> 1) >= thresh, thresh = floor(rate * 2^16);
// scale = 1/(1-rate).  One 32-bit hash word serves two adjacent elements (16 bits each): kernels whose threads own
// element pairs (the attention probabilities, B*h*T*T of them per block) pay half a hash per element.
// Masks are a pure function of (seed, step, site, flat element index in the TF-shaped tensor), so the
// backward kernels regenerate them and tests can hand the identical masks to the oracle.
struct Drop {
  unsigned k0, k1;   // mixed (seed, step, site)
  unsigned thresh;   // 16-bit threshold; 0 => dropout disabled
  float scale;
};

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

__device__ __forceinline__ Drop make_drop(float rate, unsigned long long seed, const unsigned long long* step,
                                          int site) {
  Drop d;
  unsigned long long st = step ? *step : 0ull;
  const unsigned long long key =
      mix64(seed + st * 0x9E3779B97F4A7C15ull) ^ mix64(0xD1B54A32D192ED03ull * (unsigned long long)(site + 1));
  d.k0 = (unsigned)key;
  d.k1 = (unsigned)(key >> 32);
  d.thresh = rate > 0.f ? (unsigned)fmin(65535.0, floor((double)rate * 65536.0)) : 0u;
  d.scale = rate > 0.f ? 1.0f / (1.0f - rate) : 1.0f;
  return d;
}

// 32-bit avalanche hash of the element index (two multiplies); the 64-bit key and the high index word are folded in
__device__ __forceinline__ unsigned drop_hash(const Drop& d, unsigned long long idx) {
  unsigned x = (unsigned)idx ^ d.k0;
  x += (unsigned)(idx >> 32) * 0x9E3779B1u;
  x ^= x >> 16;
  x *= 0x21F0AAADu;
  x ^= x >> 15;
  x *= 0x735A2D97u;
  x ^= x >> 15;
  x ^= d.k1;
  x *= 0x9E3779B1u;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ bool drop_keep(const Drop& d, unsigned long long idx) {
  const unsigned h = drop_hash(d, idx >> 1);
  return ((idx & 1ull) ? (h >> 16) : (h & 0xffffu)) >= d.thresh;
}

// multipliers of the adjacent elements idx, idx + 1 (idx even => one hash word)
__device__ __forceinline__ void drop_mul2(const Drop& d, unsigned long long idx, float& m0, float& m1) {
  if (d.thresh == 0u) { m0 = m1 = 1.0f; return; }
  if ((idx & 1ull) == 0ull) {
    const unsigned h = drop_hash(d, idx >> 1);
    m0 = (h & 0xffffu) >= d.thresh ? d.scale : 0.0f;
    m1 = (h >> 16) >= d.thresh ? d.scale : 0.0f;
  } else {
    m0 = drop_keep(d, idx) ? d.scale : 0.0f;
    m1 = drop_keep(d, idx + 1) ? d.scale : 0.0f;
  }
}

// the same for a caller that knows idx is even: no parity test, no branch (thresh == 0 keeps everything at scale 1)
__device__ __forceinline__ void drop_mul2_even(const Drop& d, unsigned long long idx, float& m0, float& m1) {
  const unsigned h = drop_hash(d, idx >> 1);
  m0 = (h & 0xffffu) >= d.thresh ? d.scale : 0.0f;
  m1 = (h >> 16) >= d.thresh ? d.scale : 0.0f;
}

// multiplier applied by tf.layers.dropout at element idx (0 or 1/(1-rate)); 1 when disabled
__device__ __forceinline__ float drop_mul(const Drop& d, unsigned long long idx) {
  if (d.thresh == 0u) return 1.0f;
  return drop_keep(d, idx) ? d.scale : 0.0f;
}

}  // namespace cast
