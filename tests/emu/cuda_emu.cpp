// TEST INFRASTRUCTURE ONLY — see cuda_emu.h.
#include "cuda_emu.h"

namespace cast_emu {
thread_local uint3_emu t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local BlockCtx* t_ctx = nullptr;
thread_local int t_lin = 0;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const int nwarps = (nthreads + 31) / 32;
  BlockCtx ctx;
  ctx.nthreads = nthreads;
  ctx.wbar.resize(nwarps);
  ctx.slots.assign((size_t)nwarps * 32, 0);
  // 16-byte aligned dynamic shared memory, poisoned so that reads of unwritten smem show up as NaNs
  void* mem = nullptr;
  if (posix_memalign(&mem, 128, smem + 128) != 0) abort();
  ctx.dyn_smem = static_cast<unsigned char*>(mem);
  pthread_barrier_init(&ctx.bar, nullptr, nthreads);
  for (int w = 0; w < nwarps; ++w) {
    int cnt = std::min(32, nthreads - w * 32);
    pthread_barrier_init(&ctx.wbar[w], nullptr, cnt);
  }
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        memset(ctx.dyn_smem, 0xff, smem);
        pool.clear();
        for (int t = 0; t < nthreads; ++t) {
          pool.emplace_back([&, t, bx, by, bz]() {
            t_ctx = &ctx;
            t_lin = t;
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            t_blockIdx.x = bx;
            t_blockIdx.y = by;
            t_blockIdx.z = bz;
            t_blockDim = block;
            t_gridDim = grid;
            body();
          });
        }
        for (auto& th : pool) th.join();
      }
  pthread_barrier_destroy(&ctx.bar);
  for (auto& b : ctx.wbar) pthread_barrier_destroy(&b);
  free(mem);
}
}  // namespace cast_emu
