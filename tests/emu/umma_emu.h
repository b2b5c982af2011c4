// TEST INFRASTRUCTURE ONLY — never built into or loaded by the product package.
//
// Host emulation of the slice of the tcgen05 / mbarrier / bulk-copy API that csrc/umma.cuh wraps, so that the
// tcgen05 row kernels (csrc/row_umma.cu) can be compiled with g++ -DCAST_EMU and their staging layout, descriptor
// arithmetic, barrier protocol and epilogue indexing checked against the oracle without a GPU.
//   * shared-memory "addresses" are byte offsets into the block's dynamic shared memory (cast_emu::dyn_smem_ptr());
//   * tensor memory is one global [128 lanes][512 columns] fp32 array (blocks run one after another);
//   * tcgen05.mma executes synchronously in the issuing thread: K-major, no-swizzle canonical layout only
//     (element (row, k) at (row/8)*SBO + (row%8)*16 + (k/4)*LBO + (k%4)*4), tf32 inputs truncated to 19 bits,
//     fp32 accumulation; tcgen05.commit therefore arrives on the mbarrier at once;
//   * mbarriers are packed into their 8 bytes (pending / expected arrivals, transaction bytes, phase) and updated
//     under one global mutex; waits spin with sched_yield.
#pragma once
#include <sched.h>
#include <assert.h>
#include <atomic>
#include <mutex>

namespace cast_emu {
inline std::mutex& mbar_mutex() {
  static std::mutex m;
  return m;
}
inline float (*tmem())[512] {
  static float t[128][512];
  return t;
}
struct MbarBits {
  uint32_t pending : 15, expected : 15, phase : 1, unused : 1;
  int32_t tx;
};
static_assert(sizeof(MbarBits) == 8, "mbarrier emulation must fit the 8-byte barrier object");
inline void mbar_try_complete(MbarBits* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1u;
    b->pending = b->expected;
  }
}
}  // namespace cast_emu

namespace cast {
namespace umma {

inline uint32_t smem_u32(const void* p) {
  const long off = static_cast<const unsigned char*>(p) - cast_emu::dyn_smem_ptr();
  assert(off >= 0 && off < (1 << 18));
  return (uint32_t)off;
}
inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
inline uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

inline void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp; a single allocation per kernel
  assert(ncols >= 32 && ncols <= 512 && (ncols & (ncols - 1)) == 0);
  if (cast_emu::t_lin % 32 == 0) {
    *dst_smem = 0u;
    float(*t)[512] = cast_emu::tmem();
    for (int l = 0; l < 128; ++l)
      for (int c = 0; c < 512; ++c) t[l][c] = __builtin_nanf("");  // uninitialised tensor memory
  }
  cast_emu::sync_warp();
}
inline void tmem_free(uint32_t, uint32_t) {}
inline void fence_before_sync() {}
inline void fence_after_sync() {}
inline void fence_smem_to_async() {}

inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> g(cast_emu::mbar_mutex());
  cast_emu::MbarBits* b = reinterpret_cast<cast_emu::MbarBits*>(bar);
  b->pending = count;
  b->expected = count;
  b->phase = 0;
  b->unused = 0;
  b->tx = 0;
}
inline bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t = 0) {
  for (;;) {
    {
      std::lock_guard<std::mutex> g(cast_emu::mbar_mutex());
      if (reinterpret_cast<cast_emu::MbarBits*>(bar)->phase != (parity & 1u)) return true;
    }
    sched_yield();
  }
}
inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> g(cast_emu::mbar_mutex());
  cast_emu::MbarBits* b = reinterpret_cast<cast_emu::MbarBits*>(bar);
  assert(b->pending > 0);
  b->pending -= 1;
  cast_emu::mbar_try_complete(b);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(cast_emu::mbar_mutex());
  cast_emu::MbarBits* b = reinterpret_cast<cast_emu::MbarBits*>(bar);
  assert(b->pending > 0);
  b->tx += (int32_t)bytes;
  b->pending -= 1;
  cast_emu::mbar_try_complete(b);
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  assert(bytes % 16 == 0 && ((uintptr_t)dst_smem & 15) == 0 && ((uintptr_t)src_gmem & 15) == 0);
  memcpy(dst_smem, src_gmem, bytes);
  std::lock_guard<std::mutex> g(cast_emu::mbar_mutex());
  cast_emu::MbarBits* b = reinterpret_cast<cast_emu::MbarBits*>(bar);
  b->tx -= (int32_t)bytes;
  cast_emu::mbar_try_complete(b);
}

inline float emu_tf32_in(const unsigned char* p) {
  uint32_t u;
  memcpy(&u, p, 4);
  u &= 0xffffe000u;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// D[tmem] (+)= A[smem] * B[smem]^T for one K = 8 step (K-major, no swizzle); issued by ONE thread
inline void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
  assert(M == 128 && N >= 16 && N <= 256 && N % 16 == 0);
  assert(((idesc >> 15) & 3u) == 0u);  // MN-major operands are not emulated (and need a swizzled layout on hardware)
  const unsigned char* sm = cast_emu::dyn_smem_ptr();
  const uint32_t a0 = (uint32_t)(desc_a & 0x3FFFu) << 4, alb = (uint32_t)((desc_a >> 16) & 0x3FFFu) << 4,
                 asb = (uint32_t)((desc_a >> 32) & 0x3FFFu) << 4;
  const uint32_t b0 = (uint32_t)(desc_b & 0x3FFFu) << 4, blb = (uint32_t)((desc_b >> 16) & 0x3FFFu) << 4,
                 bsb = (uint32_t)((desc_b >> 32) & 0x3FFFu) << 4;
  const int col0 = (int)(tmem_d & 0xFFFFu), lane0 = (int)(tmem_d >> 16);
  assert(lane0 == 0 && col0 + N <= 512);
  float(*t)[512] = cast_emu::tmem();
  float a[8], b[8];
  for (int m = 0; m < M; ++m) {
    for (int k = 0; k < 8; ++k) a[k] = emu_tf32_in(sm + a0 + (m >> 3) * asb + (m & 7) * 16 + (k >> 2) * alb + (k & 3) * 4);
    for (int n = 0; n < N; ++n) {
      for (int k = 0; k < 8; ++k) b[k] = emu_tf32_in(sm + b0 + (n >> 3) * bsb + (n & 7) * 16 + (k >> 2) * blb + (k & 3) * 4);
      float acc = accumulate ? t[m][col0 + n] : 0.f;
      for (int k = 0; k < 8; ++k) acc = fmaf(a[k], b[k], acc);
      t[m][col0 + n] = acc;
    }
  }
}
inline void mma_commit(uint64_t* bar) { mbar_arrive(bar); }

// 32 lanes x 32 consecutive columns: thread t of the warp receives lane (warp%4)*32 + t
inline void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xFFFFu);
  assert(lane0 == ((cast_emu::t_lin / 32) % 4) * 32 && col0 + 32 <= 512);
  float(*t)[512] = cast_emu::tmem();
  for (int i = 0; i < 32; ++i) v[i] = t[lane0 + cast_emu::t_lin % 32][col0 + i];
}

inline void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xFFFFu);
  assert(lane0 == ((cast_emu::t_lin / 32) % 4) * 32 && col0 + 16 <= 512);
  float(*t)[512] = cast_emu::tmem();
  for (int i = 0; i < 16; ++i) v[i] = t[lane0 + cast_emu::t_lin % 32][col0 + i];
}

// named barrier: generation-counting spin barrier (blocks run one after another, so one static set suffices)
inline void named_sync(int id, int nthreads) {
  static std::atomic<int> count[16], gen[16];
  assert(id > 0 && id < 16);
  const int g = gen[id].load();
  if (count[id].fetch_add(1) + 1 == nthreads) {
    count[id].store(0);
    gen[id].fetch_add(1);
  } else {
    while (gen[id].load() == g) sched_yield();
  }
}

}  // namespace umma
}  // namespace cast
