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
