// Peer memory for the row-sharded item table: one process per GPU; a rank's table shard, its sorted
// gradient entries and their source rows are read by the other ranks' kernels directly over NVLink / NVSwitch.  The
// buffers are ordinary device allocations of the owning process; their cudaIpcMemHandle travels through
// torch.distributed (dist.py) and is mapped here with peer access enabled.
#include "cast_rt.cuh"

using namespace cast;

/* a device allocation of its own (cudaMalloc: one IPC handle covers exactly this buffer) and its 64-byte handle */
extern "C" int cast_peer_alloc(size_t bytes, void** ptr, void* ipc_handle64) {
  if (!ptr || !ipc_handle64 || bytes == 0) return set_error(CAST_ERR_BAD_ARG, "peer_alloc");
#ifdef CAST_EMU
  return set_error(CAST_ERR_UNSUPPORTED, "peer_alloc: CUDA IPC is not emulated (tests use POSIX shared memory)");
#else
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(CAST_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    return set_error(CAST_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaMemset(p, 0, bytes);
  memcpy(ipc_handle64, &h, sizeof(h));
  *ptr = p;
  return CAST_OK;
#endif
}

extern "C" int cast_peer_free(void* ptr) {
  if (!ptr) return set_error(CAST_ERR_BAD_ARG, "peer_free");
#ifdef CAST_EMU
  return set_error(CAST_ERR_UNSUPPORTED, "peer_free");
#else
  if (cudaFree(ptr) != cudaSuccess) {
    cudaGetLastError();
    return set_error(CAST_ERR_CUDA, "peer_free");
  }
  return CAST_OK;
#endif
}

extern "C" int cast_peer_open(const void* ipc_handle64, void** base_ptr) {
  if (!ipc_handle64 || !base_ptr) return set_error(CAST_ERR_BAD_ARG, "peer_open");
#ifdef CAST_EMU
  return set_error(CAST_ERR_UNSUPPORTED, "peer_open: CUDA IPC is not emulated (tests use POSIX shared memory)");
#else
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, sizeof(h));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(CAST_ERR_CUDA, cudaGetErrorString(e));
  }
  *base_ptr = p;
  return CAST_OK;
#endif
}

extern "C" int cast_peer_close(void* base_ptr) {
  if (!base_ptr) return set_error(CAST_ERR_BAD_ARG, "peer_close");
#ifdef CAST_EMU
  return set_error(CAST_ERR_UNSUPPORTED, "peer_close");
#else
  if (cudaIpcCloseMemHandle(base_ptr) != cudaSuccess) {
    cudaGetLastError();
    return set_error(CAST_ERR_CUDA, "peer_close");
  }
  return CAST_OK;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// Data-parallel gradient exchange without a collective library: every rank leaves its flat gradient buffer in peer
// memory, a flag barrier over NVLink tells when all of them are final, and every rank sums the n buffers itself in
// rank order (fixed order: bit-reproducible, the same bits on every rank) — 4 B x #params x (n-1) of NVLink reads per
// rank, a few microseconds of barrier, all of it capturable in the step's CUDA graph.
namespace cast {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#ifndef CAST_EMU
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#ifndef CAST_EMU
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}

// flags[r]: rank r's arrival array (n x u64, in rank r's memory); state: {epoch, timeout flag} local to this rank.
// Thread r announces this rank's arrival in rank r's array and waits for rank r's arrival in its own.
__global__ void peer_barrier_kernel(unsigned long long* const* __restrict__ flags, int rank, int n,
                                    unsigned long long* __restrict__ state) {
  __shared__ unsigned long long epoch;
  if (threadIdx.x == 0) epoch = state[0] + 1ull;
  __syncthreads();
  const unsigned long long e = epoch;
  const int r = threadIdx.x;
  if (r < n) {
    __threadfence_system();
    st_release_sys(flags[r] + rank, e);
    const unsigned long long* mine = flags[rank] + r;
    unsigned long long spins = 0;
    while (ld_acquire_sys(mine) < e) {
      if (++spins > (1ull << 31)) {  // ~ tens of seconds: a peer died; report instead of hanging the GPU
        state[1] = 1ull;
        break;
      }
#ifdef CAST_EMU
      sched_yield();
#endif
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) state[0] = e;
}

// out[i] = sum_r g[r][i] (r ascending), i < count; the last 4 elements (loss_sum, auc_sum, count, pad) included
__global__ void peer_reduce_kernel(const float* const* __restrict__ g, int n, long count, float* __restrict__ out) {
  const long i4 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= count) return;
  if (i4 + 4 <= count) {
    float4 s = *reinterpret_cast<const float4*>(g[0] + i4);
    for (int r = 1; r < n; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(g[r] + i4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i4) = s;
  } else {
    for (long i = i4; i < count; ++i) {
      float s = g[0][i];
      for (int r = 1; r < n; ++r) s += g[r][i];
      out[i] = s;
    }
  }
}

}  // namespace cast

/* flags: DEVICE array of n pointers (rank r's n x u64 arrival array, zero-initialised peer memory); state: 2 x u64 of
 * this rank (zero-initialised): [0] epoch, [1] set to 1 if a peer never arrived.  One launch = one barrier. */
extern "C" int cast_peer_barrier(void* const* flags, int rank, int n, void* state, void* stream) {
  if (!flags || !state || n < 1 || n > 64 || rank < 0 || rank >= n) return set_error(CAST_ERR_BAD_ARG, "peer_barrier");
  CAST_LAUNCH(peer_barrier_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream,
              reinterpret_cast<unsigned long long* const*>(flags), rank, n, static_cast<unsigned long long*>(state));
  return check_launch("peer_barrier");
}

/* out[i] = sum over the n ranks (rank order) of grads[r][i]; grads: DEVICE array of n device pointers (16-byte aligned
 * buffers of `count` floats: own buffer + peer mappings). */
extern "C" int cast_peer_reduce(const void* const* grads, int n, long count, float* out, void* stream) {
  if (!grads || !out || n < 1 || count <= 0) return set_error(CAST_ERR_BAD_ARG, "peer_reduce");
  const long nthreads = cdiv(count, 4);
  CAST_LAUNCH(peer_reduce_kernel, dim3((unsigned)cdiv(nthreads, 256)), dim3(256), 0, (cudaStream_t)stream,
              reinterpret_cast<const float* const*>(grads), n, count, out);
  return check_launch("peer_reduce");
}
