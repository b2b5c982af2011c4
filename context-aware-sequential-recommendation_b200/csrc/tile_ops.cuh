// Register-tiled FP32 building blocks shared by the attention and fused row kernels.  256 threads form a 16x16
// grid (ty = tid/16, tx = tid%16); operands live in shared memory, both read with LDS.128.
#pragma once
#include "cast_rt.cuh"

namespace cast {

// acc[ii][jj] = sum_c As[(ty*RI+ii)][c] * Bs[(tx+16*jj)][c]
template <int RI, int RJ>
__device__ __forceinline__ void dot_tile(const float* __restrict__ As, const float* __restrict__ Bs, int DS, int dpad,
                                         float (&acc)[RI][RJ], int ty, int tx) {
#pragma unroll
  for (int ii = 0; ii < RI; ++ii)
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) acc[ii][jj] = 0.f;
  for (int c = 0; c < dpad; c += 4) {
    float4 a[RI], b[RJ];
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) a[ii] = *reinterpret_cast<const float4*>(&As[(ty * RI + ii) * DS + c]);
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) b[jj] = *reinterpret_cast<const float4*>(&Bs[(tx + 16 * jj) * DS + c]);
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

// acc[ii][cc] += sum_{j<nj} Ps[(ty*RI+ii)*PS + j] * Vs[j*DS + col + cc]      (nj multiple of 4, col multiple of 4)
template <int RI>
__device__ __forceinline__ void pv_tile(const float* __restrict__ Ps, int PS, const float* __restrict__ Vs, int DS,
                                        int nj, int col, float (&acc)[RI][4], int ty) {
  for (int j = 0; j < nj; j += 4) {
    float4 p[RI], v[4];
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) p[ii] = *reinterpret_cast<const float4*>(&Ps[(ty * RI + ii) * PS + j]);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) v[jj] = *reinterpret_cast<const float4*>(&Vs[(j + jj) * DS + col]);
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
