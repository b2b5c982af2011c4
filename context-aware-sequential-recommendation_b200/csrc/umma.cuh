// Thin inline-PTX layer over the Blackwell 5th-generation tensor-core path used by the scoring GEMM:
// tcgen05.mma (kind::tf32, operands in shared memory, accumulator in tensor memory), TMEM allocation, tcgen05.ld,
// tcgen05.commit -> mbarrier, and the proxy / thread-sync fences between them.  sm_100a only.
//
// Shared-memory operand layout (K-major, no swizzle; "interleaved" canonical layout of the UMMA descriptor):
//   a tile of R rows x KC fp32 is stored as KC/4 slabs; slab c holds elements [4c, 4c+4) of every row, 16 bytes per
//   row, rows contiguous.  8 consecutive rows x 16 bytes = one 128-byte core matrix; the next 8-row group follows at
//   SBO = 128 bytes; the next 4-element K slab follows at LBO = slab pitch.  The slab pitch is padded by 16 bytes
//   (R*16 + 16) so that threads writing consecutive slabs of one row hit different banks.
//   One tcgen05.mma (tf32) consumes K = 8 = two slabs; the descriptor start address advances by 2 slabs per k-step.
#pragma once
#include <stdint.h>

#ifdef CAST_EMU
// host emulation of this API (tests/emu, test infrastructure): same names and semantics, see the header
#include "umma_emu.h"
#else

namespace cast {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor: start address, leading (K) and stride (M/N) byte offsets in 16-byte units,
// descriptor version 1 (Blackwell), base offset 0, layout type 0 (no swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// the same descriptor `bytes` further into shared memory (bytes % 16 == 0; the 14-bit start-address field cannot
// overflow for addresses below 256 KB).  One 64-bit add instead of rebuilding the descriptor: the single MMA-issuing
// thread is latency-bound on its own instruction stream, so every instruction between two tcgen05.mma counts.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// 32-bit instruction descriptor for kind::tf32: D = f32, A = B = tf32, both K-major, dense, no negate.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait on a phase parity: returns false if the phase did not complete within ~max_tries polls (never hangs)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_tries = (1u << 24)) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t i = 0; i < max_tries; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival + announce `bytes` of asynchronous copies that will complete on this barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// TMA 1-D bulk copy global -> shared (16-byte aligned addresses, size multiple of 16); completion is signalled on the
// mbarrier as transaction bytes.  Issued by ONE thread.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T for one K = 8 step; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive columns: thread t of the warp receives lane (warp%4)*32 + t, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"  // same asm block: the registers are valid when the statement completes
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// named barrier over the first `nthreads` threads of the CTA (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// x = hi + lo + O(2^-22 |x|), hi and lo exactly representable in tf32 (10-bit mantissa)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float rem = x - hi;  // exact
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
  lo = __uint_as_float(l);
}

}  // namespace umma
}  // namespace cast

namespace cast {
namespace umma {

// Stage rows [row0, row0+R) x k in [k0, k0 + 4*slabs) of a strided fp32 matrix (element (r,k) at src[r*sr + k*sk]) as
// tf32 hi / lo slabs (layout in the header comment).  Elements with row >= rows_total or k >= kend are zero.  The
// thread <-> element mapping follows the unit stride of the source so that global reads coalesce: k-fast sources
// (sk == 1) walk the 4-element slabs of a row, row-fast sources (sr == 1; i.e. the operand is stored transposed)
// walk consecutive rows of a slab.  One thread owns ITEMS (row, slab) pairs; all of its global loads are issued
// before the first conversion so that 4*ITEMS loads are in flight per thread (the loop is latency-bound otherwise),
// then each pair becomes one 16-byte shared-memory store per half.
template <int NTHREADS, int R_MAX, int SLABS_MAX>
__device__ __forceinline__ void stage_split_strided(unsigned char* __restrict__ hi, unsigned char* __restrict__ lo,
                                                    int pitch, const float* __restrict__ src, long sr, long sk,
                                                    long row0, long rows_total, int R, long k0, long kend,
                                                    int slabs) {
  constexpr int ITEMS = (R_MAX * SLABS_MAX + NTHREADS - 1) / NTHREADS;
  const bool kfast = (sk == 1);
  const int total = R * slabs;
  float x[ITEMS][4];
  int off[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = threadIdx.x + i * NTHREADS;
    off[i] = -1;
#pragma unroll
    for (int e = 0; e < 4; ++e) x[i][e] = 0.f;
    if (idx < total) {
      int r, c;
      if (kfast) { r = idx / slabs; c = idx - r * slabs; } else { c = idx / R; r = idx - c * R; }
      off[i] = c * pitch + r * 16;
      const long row = row0 + r;
      const long k = k0 + 4 * c;
      if (row < rows_total) {
        const float* p = src + row * sr + k * sk;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k + e < kend) x[i][e] = __ldg(p + e * sk);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (off[i] < 0) continue;
    float4 h, l;
    split_tf32(x[i][0], h.x, l.x);
    split_tf32(x[i][1], h.y, l.y);
    split_tf32(x[i][2], h.z, l.z);
    split_tf32(x[i][3], h.w, l.w);
    *reinterpret_cast<float4*>(hi + off[i]) = h;
    *reinterpret_cast<float4*>(lo + off[i]) = l;
  }
}

}  // namespace umma
}  // namespace cast

#endif  // !CAST_EMU
