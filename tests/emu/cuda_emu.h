// TEST INFRASTRUCTURE ONLY — never built into or loaded by the product package.
//
// Minimal host emulation of the CUDA execution model so that the kernel sources under
// context-aware-sequential-recommendation_b200/csrc/ can be compiled with g++ (-DCAST_EMU) and their
// indexing / masking / reduction logic exercised in the GPU-less build container against the oracle.
// One CUDA thread == one OS thread; blocks run one after another; __syncthreads == pthread barrier;
// warp shuffles go through a per-warp exchange buffer.  Slow by design; tiny shapes only.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__ __restrict

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(d, s, n);
  return 0;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace cast_emu {
struct BlockCtx {
  pthread_barrier_t bar;
  std::vector<pthread_barrier_t> wbar;
  std::vector<uint64_t> slots;  // [nwarps][32]
  unsigned char* dyn_smem;
  int nthreads;
};
extern thread_local uint3_emu t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local BlockCtx* t_ctx;
extern thread_local int t_lin;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
inline unsigned char* dyn_smem_ptr() { return t_ctx->dyn_smem; }
inline void sync_block() { pthread_barrier_wait(&t_ctx->bar); }
inline void sync_warp() { pthread_barrier_wait(&t_ctx->wbar[t_lin / 32]); }
inline uint64_t exchange(uint64_t v, int src_lane) {
  uint64_t* s = &t_ctx->slots[(size_t)(t_lin / 32) * 32];
  s[t_lin % 32] = v;
  sync_warp();
  uint64_t r = s[src_lane & 31];
  sync_warp();
  return r;
}
// every lane publishes one pointer; `reader(ptrs)` runs while all 32 pointers (and what they point to) are stable
template <class F>
inline void warp_publish(const void* mine, F&& reader) {
  uint64_t* s = &t_ctx->slots[(size_t)(t_lin / 32) * 32];
  s[t_lin % 32] = (uint64_t)(uintptr_t)mine;
  sync_warp();
  reader(s);
  sync_warp();
}
}  // namespace cast_emu

#define threadIdx cast_emu::t_threadIdx
#define blockIdx cast_emu::t_blockIdx
#define blockDim cast_emu::t_blockDim
#define gridDim cast_emu::t_gridDim

static inline void __syncthreads() { cast_emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cast_emu::sync_warp(); }

template <class T>
static inline T emu_shfl(T v, int src) {
  uint64_t u = 0;
  memcpy(&u, &v, sizeof(T));
  u = cast_emu::exchange(u, src);
  T r;
  memcpy(&r, &u, sizeof(T));
  return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = cast_emu::t_lin % 32;
  return emu_shfl(v, (lane / width) * width + (src % width));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  int lane = cast_emu::t_lin % 32;
  return emu_shfl(v, lane ^ m);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
  int lane = cast_emu::t_lin % 32;
  int src = lane + d;
  if ((src / width) != (lane / width)) src = lane;
  return emu_shfl(v, src);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) {
  int lane = cast_emu::t_lin % 32;
  int src = lane - d;
  if (src < 0 || (src / width) != (lane / width)) src = lane;
  return emu_shfl(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (emu_shfl<unsigned>(pred ? 1u : 0u, l) << l);
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// ---- atomics
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline float atomicAdd(float* p, float v) {
  uint32_t* ip = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(ip, __ATOMIC_SEQ_CST), nw;
  float f;
  do {
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&nw, &f, 4);
  } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  memcpy(&f, &old, 4);
  return f;
}
static inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- math / misc intrinsics
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
using std::max;
using std::min;

#define __align__(n) __attribute__((aligned(n)))
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (emu_shfl<unsigned>(v, l) == v) r |= (1u << l);
  return r;
}
#define __expf(x) expf(x)
#define __logf(x) logf(x)
