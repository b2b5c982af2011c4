// Warp-level tensor-core building blocks for the small-K fused kernels (attention, row-tile projections):
// mma.sync m16n8k8 TF32 with fp32 accumulation, used as 3xTF32 (x = hi + lo, hi = tf32(x), lo = x - hi;
// A*B ~= lo*hi + hi*lo + hi*hi) so that products keep ~2^-21 relative error and the fp32 1e-4 parity bar holds
// (single-pass TF32 does not).  Operands come from shared memory through ldmatrix (a 8x8 b16 matrix is a
// 8 rows x 4 floats block, and lane l receives element (l/4, l%4) -- exactly the tf32 A/B fragment order).
//
// Fragment maps (PTX ISA, mma.m16n8k8 .tf32), g = lane/4, t = lane%4:
//   A (16x8, row):  a0 (g, t)   a1 (g+8, t)   a2 (g, t+4)   a3 (g+8, t+4)
//   B ( 8x8, col):  b0 (k=t, n=g)   b1 (k=t+4, n=g)
//   C (16x8):       c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// A C fragment is reused directly as the A fragment of a following product by relabelling the contraction index
// inside the k-step: k-slot t <-> column 2t, k-slot t+4 <-> column 2t+1; the B fragment of that product is then
// loaded with the same relabelling (b0 = row 2t, b1 = row 2t+1 of the [k][n] matrix).
#pragma once
#include "cast_rt.cuh"

namespace cast {

// round-to-nearest (ties away) to tf32; finite inputs only (cvt.rna.tf32.f32 costs 4 SASS ops for its inf/nan path)
__device__ __forceinline__ unsigned tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
  hi = tf32_round(x);
  lo = __float_as_uint(x - __uint_as_float(hi));  // exact; the tensor core reads its leading 11 bits
}

template <int N>
__device__ __forceinline__ void tf32_split_n(const unsigned (&x)[N], unsigned (&hi)[N], unsigned (&lo)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) tf32_split(__uint_as_float(x[i]), hi[i], lo[i]);
}

#ifndef CAST_EMU

// (not volatile: a pure function of its register operands, so the scheduler may interleave independent chains)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// four 8x4-float blocks; lane l passes the address of row l%8 of block l/8 (16-byte aligned)
__device__ __forceinline__ void ldsm4(unsigned (&r)[4], const float* p) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

#else  // host emulation (tests/emu): same fragment maps, tf32 inputs truncated as the tensor core does

static inline float emu_tf32(unsigned u) { return __uint_as_float(u & 0xffffe000u); }

inline void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  struct Frag { unsigned a[4], b[2]; } mine{{a[0], a[1], a[2], a[3]}, {b0, b1}};
  const int lane = cast_emu::t_lin % 32, g = lane >> 2, t = lane & 3;
  float acc[4] = {c[0], c[1], c[2], c[3]};
  cast_emu::warp_publish(&mine, [&](const uint64_t* ptrs) {
    for (int k = 0; k < 8; ++k) {
      const Frag* fa = reinterpret_cast<const Frag*>((uintptr_t)ptrs[g * 4 + (k & 3)]);
      const float a_lo = emu_tf32(fa->a[(k < 4) ? 0 : 2]), a_hi = emu_tf32(fa->a[(k < 4) ? 1 : 3]);
      for (int cc = 0; cc < 2; ++cc) {
        const Frag* fb = reinterpret_cast<const Frag*>((uintptr_t)ptrs[(2 * t + cc) * 4 + (k & 3)]);
        const float bv = emu_tf32(fb->b[(k < 4) ? 0 : 1]);
        acc[cc] = fmaf(a_lo, bv, acc[cc]);
        acc[2 + cc] = fmaf(a_hi, bv, acc[2 + cc]);
      }
    }
  });
  for (int i = 0; i < 4; ++i) c[i] = acc[i];
}

inline void ldsm4(unsigned (&r)[4], const float* p) {
  const int lane = cast_emu::t_lin % 32;
  cast_emu::warp_publish(p, [&](const uint64_t* ptrs) {
    for (int j = 0; j < 4; ++j) {
      const float* row = reinterpret_cast<const float*>((uintptr_t)ptrs[8 * j + lane / 4]);
      r[j] = __float_as_uint(row[lane % 4]);
    }
  });
}

#endif

// ---- asynchronous global -> shared copies (LDGSTS): BYTES in {4, 8, 16}, both addresses BYTES-aligned; !valid => zeros
#ifndef CAST_EMU
template <int BYTES>
__device__ __forceinline__ void cp_async(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(src), "n"(BYTES), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else
template <int BYTES>
inline void cp_async(float* dst, const float* src, bool valid) {
  for (int i = 0; i < BYTES / 4; ++i) dst[i] = valid ? src[i] : 0.f;
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#endif

// c += A * B with 3xTF32 (small terms first)
__device__ __forceinline__ void mma_3x(float (&c)[4], const unsigned (&ah)[4], const unsigned (&al)[4], unsigned bh0,
                                       unsigned bh1, unsigned bl0, unsigned bl1) {
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}

// A fragment (rows 0..15, columns k0..k0+7) of a row-major matrix with row stride DS (DS % 4 == 0, DS/4 odd)
template <int DS>
__device__ __forceinline__ void ldsm_a(unsigned (&r)[4], const float* base, int k0, int lane) {
  const int blk = lane >> 3, rr = lane & 7;
  ldsm4(r, base + ((blk & 1) * 8 + rr) * DS + k0 + (blk >> 1) * 4);
}

// B fragments of two adjacent n-tiles from a [n][k] row-major matrix: r[0],r[1] = (b0,b1) of rows 0..7, r[2],r[3] of
// rows 8..15, contraction columns k0..k0+7
template <int DS>
__device__ __forceinline__ void ldsm_b2(unsigned (&r)[4], const float* base, int k0, int lane) {
  const int blk = lane >> 3, rr = lane & 7;
  ldsm4(r, base + ((blk >> 1) * 8 + rr) * DS + k0 + (blk & 1) * 4);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

}  // namespace cast
