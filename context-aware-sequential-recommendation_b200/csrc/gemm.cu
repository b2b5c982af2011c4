// K3 / K5 / K10 building block: strided fp32 GEMM  C[M,N] = epilogue( A[M,K] * B[K,N] )  on the FP32 pipe.
// One kernel serves the dense layers of modules.py (`tf.layers.dense` :203-205/:333-334, `conv1d(k=1)` :298-306)
// in all three roles:
//   forward   Y  = X  * W   (+bias, ReLU, dropout, residual, padding mask fused in the epilogue)
//   dgrad     dX = dY * W^T (B read through strides; ReLU/dropout backward and residual-gradient add fused)
//   wgrad     dW = X^T * dY (A read through strides; reduction over the B*T rows split across CTAs with
//                            fixed-order second stage -> deterministic)
// fp32 FFMA is deliberate: parity is fp32 1e-4 relative on LN/softmax chains with K as small as 25..50, where
// the 64x64x16 tiles already hold the whole K extent; tensor cores are kept for the catalog-scoring GEMM.
// Tile: 64x64x16, 256 threads, 4x4 register micro-tile, A tile stored k-major so both operands are LDS.128.
#include "cast_rt.cuh"

namespace cast {

int launch_reduce_partials(const float* partial, int nparts, long count, float* out0, long split, float* out1,
                           cudaStream_t stream);

constexpr int TM = 64, TN = 64, TK = 16, LDS_ = 68;

#ifndef CAST_EMU
// tensor-core path (gemm_umma.cu); its epilogue descriptor has the same fields as GemmEpi
struct UGemmEpi {
  const float* bias;
  int relu;
  float drop_rate;
  unsigned long long seed;
  const unsigned long long* step;
  int site;
  const float* act;
  long ld_act;
  float act_scale;
  const float* resid;
  long ldr;
  const int* row_ids;
};
int gemm_umma_launch(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc, long M,
                     int N, long K, const UGemmEpi& epi, int splits, float* partials, void* image, size_t image_bytes,
                     cudaStream_t stream);
size_t gemm_umma_image_bytes(int N, long K);
#endif
static int g_gemm_backend = 0;  // 0 = auto (tensor cores for wide shapes), 1 = FP32 FFMA tiles, 2 = tensor cores

struct GemmEpi {
  const float* bias;   // [N] or null
  int relu;            // max(0, .)
  float drop_rate;     // dropout on element (i*N + j)
  unsigned long long seed;
  const unsigned long long* step;
  int site;
  const float* act;    // [M, lda_act]: multiply by (act > 0 ? act_scale : 0)   (ReLU∘dropout backward)
  long ld_act;
  float act_scale;
  const float* resid;  // [M, ldr] added after everything above
  long ldr;
  const int* row_ids;  // multiply row i by (row_ids[i] != 0)
};

__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, long sam, long sak, const float* __restrict__ B, long sbk, long sbn,
            float* __restrict__ C, long ldc, long M, int N, long K, long klen, GemmEpi epi,
            float* __restrict__ partials) {
  __shared__ __align__(16) float As[TK * LDS_];
  __shared__ __align__(16) float Bs[TK * LDS_];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const long i0 = (long)blockIdx.x * TM;
  const int j0 = blockIdx.y * TN;
  const long kbeg = (long)blockIdx.z * klen;
  const long kend = (kbeg + klen < K) ? kbeg + klen : K;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const bool a_kfast = (sak == 1);
  const bool b_nfast = (sbn == 1);

  for (long k0 = kbeg; k0 < kend; k0 += TK) {
    // ---- stage A tile (As[k][i]) and B tile (Bs[k][j]); out-of-range elements are zero
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int ii, kk;
      if (a_kfast) { kk = t & 15; ii = (t >> 4) + 16 * p; } else { ii = t & 63; kk = (t >> 6) + 4 * p; }
      const long gi = i0 + ii, gk = k0 + kk;
      As[kk * LDS_ + ii] = (gi < M && gk < kend) ? A[gi * sam + gk * sak] : 0.f;
      int jj, kb;
      if (b_nfast) { jj = t & 63; kb = (t >> 6) + 4 * p; } else { kb = t & 15; jj = (t >> 4) + 16 * p; }
      const long gkb = k0 + kb;
      const int gj = j0 + jj;
      Bs[kb * LDS_ + jj] = (gj < N && gkb < kend) ? B[gkb * sbk + (long)gj * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk * LDS_ + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk * LDS_ + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
  }

  if (partials) {  // split-K: raw partial sums, dense [z][M][N]
    float* P = partials + (long)blockIdx.z * M * N;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const long gi = i0 + ty * 4 + x;
      if (gi >= M) continue;
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int gj = j0 + tx * 4 + y;
        if (gj < N) P[gi * N + gj] = acc[x][y];
      }
    }
    return;
  }

  const Drop d = make_drop(epi.drop_rate, epi.seed, epi.step, epi.site);
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const long gi = i0 + ty * 4 + x;
    if (gi >= M) continue;
    const float rm = epi.row_ids ? (epi.row_ids[gi] != 0 ? 1.f : 0.f) : 1.f;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int gj = j0 + tx * 4 + y;
      if (gj >= N) continue;
      float c = acc[x][y];
      if (epi.bias) c += epi.bias[gj];
      if (epi.relu) c = fmaxf(c, 0.f);
      c *= drop_mul(d, (unsigned long long)(gi * N + gj));
      if (epi.act) c *= (epi.act[gi * epi.ld_act + gj] > 0.f) ? epi.act_scale : 0.f;
      if (epi.resid) c += epi.resid[gi * epi.ldr + gj];
      c *= rm;
      C[gi * ldc + gj] = c;
    }
  }
}

// out[c] = sum over rows of X[r, c]; two fixed-order stages (rows split across gridDim.y CTAs).
__global__ void colsum_kernel(const float* __restrict__ X, long rows, long cols, long ld, long rlen,
                              float* __restrict__ partial) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long r0 = (long)blockIdx.y * rlen;
  const long r1 = (r0 + rlen < rows) ? r0 + rlen : rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four interleaved chains: the loads of 8 rows are in flight together
  long r = r0;
  for (; r + 8 <= r1; r += 8) {
    const float v0 = X[r * ld + c], v1 = X[(r + 1) * ld + c], v2 = X[(r + 2) * ld + c], v3 = X[(r + 3) * ld + c];
    const float v4 = X[(r + 4) * ld + c], v5 = X[(r + 5) * ld + c], v6 = X[(r + 6) * ld + c], v7 = X[(r + 7) * ld + c];
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    a0 += v4; a1 += v5; a2 += v6; a3 += v7;
  }
  for (; r < r1; ++r) a0 += X[r * ld + c];
  partial[(long)blockIdx.y * cols + c] = (a0 + a1) + (a2 + a3);
}

}  // namespace cast

using namespace cast;

extern "C" size_t cast_gemm_workspace_bytes(long M, int N, long K, int splits) {
  if (splits > 1) return (size_t)splits * (size_t)M * (size_t)N * sizeof(float);
#ifndef CAST_EMU
  return gemm_umma_image_bytes(N, K);   // pre-split weight image of the tensor-core path
#else
  (void)K;
  return 0;
#endif
}

extern "C" int cast_gemm(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                         long M, int N, long K, const float* bias, int relu, float drop_rate, unsigned long long seed,
                         const unsigned long long* step, int site, const float* act, long ld_act, float act_scale,
                         const float* resid, long ldr, const int* row_ids, int splits, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || splits < 1) return set_error(CAST_ERR_BAD_ARG, "gemm");
  if (drop_rate < 0.f || drop_rate >= 1.f) return set_error(CAST_ERR_BAD_ARG, "gemm: drop_rate");
  GemmEpi e;
  e.bias = bias; e.relu = relu; e.drop_rate = drop_rate; e.seed = seed; e.step = step; e.site = site;
  e.act = act; e.ld_act = ld_act; e.act_scale = act_scale; e.resid = resid; e.ldr = ldr; e.row_ids = row_ids;
  if (splits > 1) {
    if (bias || relu || drop_rate > 0.f || act || resid || row_ids || ldc != N)
      return set_error(CAST_ERR_BAD_ARG, "gemm: split-K supports no epilogue and needs ldc == N");
    if (!workspace || workspace_bytes < cast_gemm_workspace_bytes(M, N, K, splits))
      return set_error(CAST_ERR_WORKSPACE, "gemm: workspace too small");
  }
#ifndef CAST_EMU
  // measured on B200 (scripts/bench_gemm.py): the tcgen05 pipeline wins from ~0.5 GFLOP per call (25600 x 256 x 256:
  // 37-80 us vs 128-187 us); below that both paths sit on the ~20 us latency floor and the FFMA tiles are a bit faster
  const bool wide = N >= 64 && K >= 64 && M >= 64 && (double)M * N * K >= 2.5e8;
  if (g_gemm_backend == 2 || (g_gemm_backend == 0 && wide)) {
    UGemmEpi ue{bias, relu, drop_rate, seed, step, site, act, ld_act, act_scale, resid, ldr, row_ids};
    return gemm_umma_launch(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, ue, splits,
                            splits > 1 ? static_cast<float*>(workspace) : nullptr, splits > 1 ? nullptr : workspace,
                            splits > 1 ? 0 : workspace_bytes, (cudaStream_t)stream);
  }
#endif
  dim3 grid((unsigned)cdiv(M, TM), (unsigned)cdiv(N, TN), (unsigned)splits);
  if (splits == 1) {
    CAST_LAUNCH(gemm_kernel, grid, dim3(256), 0, (cudaStream_t)stream, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, K,
                e, (float*)nullptr);
    return check_launch("gemm");
  }
  long klen = cdiv(K, splits);
  klen = cdiv(klen, TK) * TK;
  float* part = static_cast<float*>(workspace);
  CAST_LAUNCH(gemm_kernel, grid, dim3(256), 0, (cudaStream_t)stream, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, klen,
              e, part);
  int rc = check_launch("gemm(split)");
  if (rc) return rc;
  return launch_reduce_partials(part, splits, M * (long)N, C, M * (long)N, (float*)nullptr, (cudaStream_t)stream);
}

extern "C" int cast_gemm_set_backend(int which) {
  if (which < 0 || which > 2) return set_error(CAST_ERR_BAD_ARG, "gemm_set_backend");
  g_gemm_backend = which;
  return CAST_OK;
}

extern "C" size_t cast_colsum_workspace_bytes(long rows, long cols) {
  long splits = rows >= 4096 ? 64 : 1;
  return (size_t)splits * (size_t)cols * sizeof(float);
}

extern "C" int cast_colsum(const float* X, long rows, long cols, long ld, float* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!X || !out || rows <= 0 || cols <= 0) return set_error(CAST_ERR_BAD_ARG, "colsum");
  const long splits = rows >= 4096 ? 64 : 1;
  if (!workspace || workspace_bytes < cast_colsum_workspace_bytes(rows, cols))
    return set_error(CAST_ERR_WORKSPACE, "colsum: workspace too small");
  float* part = static_cast<float*>(workspace);
  const long rlen = cdiv(rows, splits);
  CAST_LAUNCH(colsum_kernel, dim3((unsigned)cdiv(cols, 128), (unsigned)splits), dim3(128), 0, (cudaStream_t)stream, X,
              rows, cols, ld, rlen, part);
  int rc = check_launch("colsum");
  if (rc) return rc;
  return launch_reduce_partials(part, (int)splits, cols, out, cols, (float*)nullptr, (cudaStream_t)stream);
}
