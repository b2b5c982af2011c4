// Register-tiled FP32 building blocks shared by the attention and fused row kernels.  256 threads form a 16x16
// grid (ty = tid/16, tx = tid%16); operands live in shared memory, both read with LDS.128.  Row strides are
// template parameters wherever the caller knows them at compile time: with run-time strides ptxas re-derives every
// operand address with IMAD/LEA inside the inner loop and FFMA drops to ~1/3 of the issued instructions (ncu,
// profiles/r01_attn_fwd_v0.txt); with immediates the loop is LDS.128 + FFMA only.
#pragma once
#include "cast_rt.cuh"

namespace cast {

// acc[ii][jj] = sum_c As[(ty*RI+ii)][c] * Bs[(tx+16*jj)][c],  c < DPAD (multiple of 4), row stride DS floats
template <int RI, int RJ, int DPAD, int DS>
__device__ __forceinline__ void dot_tile(const float* __restrict__ As, const float* __restrict__ Bs,
                                         float (&acc)[RI][RJ], int ty, int tx) {
#pragma unroll
  for (int ii = 0; ii < RI; ++ii)
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) acc[ii][jj] = 0.f;
  const float* ap = As + ty * RI * DS;
  const float* bp = Bs + tx * DS;
#pragma unroll
  for (int c = 0; c < DPAD; c += 4) {
    float4 a[RI], b[RJ];
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) a[ii] = *reinterpret_cast<const float4*>(ap + ii * DS + c);
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) b[jj] = *reinterpret_cast<const float4*>(bp + 16 * jj * DS + c);
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        float s = acc[ii][jj];
        s = fmaf(a[ii].x, b[jj].x, s);
        s = fmaf(a[ii].y, b[jj].y, s);
        s = fmaf(a[ii].z, b[jj].z, s);
        s = fmaf(a[ii].w, b[jj].w, s);
        acc[ii][jj] = s;
      }
  }
}

// acc[ii][cc] += sum_{j<nj} Ps[(ty*RI+ii)*PS + j] * Vs[j*ds + col + cc]   (nj, col multiples of 4)
// DS > 0: compile-time stride of Vs; DS == 0: run-time stride ds_rt.  PS is a run-time stride (depends on maxlen).
template <int RI, int DS>
__device__ __forceinline__ void pv_tile(const float* __restrict__ Ps, int PS, const float* __restrict__ Vs, int ds_rt,
                                        int nj, int col, float (&acc)[RI][4], int ty) {
  const int ds = DS > 0 ? DS : ds_rt;
  const float* prow[RI];
#pragma unroll
  for (int ii = 0; ii < RI; ++ii) prow[ii] = Ps + (ty * RI + ii) * PS;
  const float* vp = Vs + col;
#pragma unroll 2
  for (int j = 0; j < nj; j += 4) {
    float4 p[RI], v[4];
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) p[ii] = *reinterpret_cast<const float4*>(prow[ii] + j);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) v[jj] = *reinterpret_cast<const float4*>(vp + jj * ds);
    vp += 4 * ds;
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) {
      acc[ii][0] = fmaf(p[ii].w, v[3].x, fmaf(p[ii].z, v[2].x, fmaf(p[ii].y, v[1].x, fmaf(p[ii].x, v[0].x, acc[ii][0]))));
      acc[ii][1] = fmaf(p[ii].w, v[3].y, fmaf(p[ii].z, v[2].y, fmaf(p[ii].y, v[1].y, fmaf(p[ii].x, v[0].y, acc[ii][1]))));
      acc[ii][2] = fmaf(p[ii].w, v[3].z, fmaf(p[ii].z, v[2].z, fmaf(p[ii].y, v[1].z, fmaf(p[ii].x, v[0].z, acc[ii][2]))));
      acc[ii][3] = fmaf(p[ii].w, v[3].w, fmaf(p[ii].z, v[2].w, fmaf(p[ii].y, v[1].w, fmaf(p[ii].x, v[0].w, acc[ii][3]))));
    }
  }
}

}  // namespace cast
