// K2: layer normalisation exactly as modules.py:53-80 `normalize` — biased variance, eps inside the sqrt,
// gamma * xhat + beta — forward and backward, one warp per row with shuffle reductions (H <= 1024).
// Forward also emits the two "is this row exactly zero-sum" flags that multihead_attention derives from its
// inputs (modules.py:222 key mask from sum_H(keys), :248 query mask from sum_H(queries)).
// Backward: dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma; dgamma/dbeta are reduced in two
// fixed-order stages (per-CTA partials, then one pass over CTAs) so results are run-to-run bit-identical.
#include "cast_rt.cuh"

namespace cast {

constexpr int LN_MAXV = 32;  // values per lane -> H <= 1024
constexpr int LN_WARPS = 8;
constexpr int LN_ROWS_PER_CTA = 64;  // rows handled by one CTA in the backward kernel

template <int NV>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, long N, int H, float eps,
                                     float* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, float* __restrict__ xnz,
                                     float* __restrict__ ynz) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* xr = x + row * H;
  float v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < H ? xr[c] : 0.f;
    s += v[i];
  }
  s = warp_sum(s);
  const float mean = s / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dlt = (lane + 32 * i < H) ? v[i] - mean : 0.f;
    q += dlt * dlt;
  }
  q = warp_sum(q);
  const float var = q / (float)H;
  const float stdv = sqrtf(var + eps);  // (variance + epsilon) ** .5
  float ys = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      const float o = gamma[c] * ((v[i] - mean) / stdv) + beta[c];
      y[row * H + c] = o;
      ys += o;
    }
  }
  ys = warp_sum(ys);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = 1.0f / stdv;
    if (xnz) xnz[row] = (s != 0.f) ? 1.f : 0.f;
    if (ynz) ynz[row] = (ys != 0.f) ? 1.f : 0.f;
  }
}

// grid = ceil(N / LN_ROWS_PER_CTA); partial[cta][0..H) = sum dy*xhat, partial[cta][H..2H) = sum dy
template <int NV>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                     const float* __restrict__ gamma, long N, int H,
                                     const float* __restrict__ dx_add, float* __restrict__ dx,
                                     float* __restrict__ partial) {
  CAST_DYN_SMEM(float, sm);  // [LN_WARPS][2H]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long row0 = (long)blockIdx.x * LN_ROWS_PER_CTA;
  float dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = 0.f;
  // RB rows of a warp are in flight together (their loads are independent; one row at a time left the kernel
  // latency-bound at the small N of the short-sequence configs); rows are still accumulated in ascending order
  constexpr int RB = NV <= 4 ? 4 : (NV <= 8 ? 2 : 1);
  for (int r0 = warp; r0 < LN_ROWS_PER_CTA; r0 += LN_WARPS * RB) {
    float g[RB][NV], xh[RB][NV], dv[RB][NV];
    float mean[RB], rstd[RB], s1[RB], s2[RB];
    bool on[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int r = r0 + k * LN_WARPS;
      const long row = row0 + r;
      on[k] = r < LN_ROWS_PER_CTA && row < N;  // warp-uniform
      mean[k] = on[k] ? mean_in[row] : 0.f;
      rstd[k] = on[k] ? rstd_in[row] : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const bool in = on[k] && c < H;
        dv[k][i] = in ? dy[row * H + c] : 0.f;
        xh[k][i] = in ? x[row * H + c] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      s1[k] = s2[k] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        g[k][i] = 0.f;
        if (on[k] && c < H) {
          xh[k][i] = (xh[k][i] - mean[k]) * rstd[k];
          g[k][i] = dv[k][i] * gamma[c];
          s1[k] += g[k][i];
          s2[k] += g[k][i] * xh[k][i];
          dg[i] += dv[k][i] * xh[k][i];
          db[i] += dv[k][i];
        } else {
          xh[k][i] = 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      s1[k] = warp_sum(s1[k]) / (float)H;
      s2[k] = warp_sum(s2[k]) / (float)H;
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      if (!on[k]) continue;
      const long row = row0 + r0 + k * LN_WARPS;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < H) {
          float o = rstd[k] * (g[k][i] - s1[k] - xh[k][i] * s2[k]);
          if (dx_add) o += dx_add[row * H + c];
          dx[row * H + c] = o;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      sm[warp * 2 * H + c] = dg[i];
      sm[warp * 2 * H + H + c] = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < LN_WARPS; ++w) s += sm[w * 2 * H + c];
    partial[(long)blockIdx.x * 2 * H + c] = s;
  }
}

// out[c] = sum_{p < nparts} partial[p][c] in a fixed order.  Generic second stage used by LN, bias and wgrad.
// A CTA covers 32 columns with 8 part-groups (group g owns parts g, g+8, ...; 4 interleaved accumulators each, so 32
// independent loads are in flight per column instead of one dependent chain over all parts); the 8 group sums are
// then folded in group order through shared memory => same bits every run.
constexpr int RP_COLS = 32, RP_GROUPS = 8, RP_ILP = 4;
__global__ void __launch_bounds__(RP_COLS * RP_GROUPS)
reduce_partials_kernel(const float* __restrict__ partial, int nparts, long count, float* __restrict__ out0, long split,
                       float* __restrict__ out1) {
  __shared__ float red[RP_GROUPS][RP_COLS];
  const int cl = threadIdx.x % RP_COLS, g = threadIdx.x / RP_COLS;
  for (long c0 = (long)blockIdx.x * RP_COLS; c0 < count; c0 += (long)gridDim.x * RP_COLS) {
    const long c = c0 + cl;
    float acc[RP_ILP];
#pragma unroll
    for (int u = 0; u < RP_ILP; ++u) acc[u] = 0.f;
    if (c < count) {
      for (int p0 = g; p0 < nparts; p0 += RP_GROUPS * RP_ILP) {
#pragma unroll
        for (int u = 0; u < RP_ILP; ++u) {
          const int p = p0 + u * RP_GROUPS;
          if (p < nparts) acc[u] += partial[(long)p * count + c];
        }
      }
    }
    red[g][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (g == 0 && c < count) {
      float s = red[0][cl];
#pragma unroll
      for (int k = 1; k < RP_GROUPS; ++k) s += red[k][cl];
      if (c < split)
        out0[c] = s;
      else
        out1[c - split] = s;
    }
    __syncthreads();
  }
}

// several independent reductions in one launch (blockIdx.y = job): the per-CTA gradient partials of every fused
// backward kernel of a step are folded once, after the last of them, instead of by one small launch each
constexpr int RP_MAX_JOBS = 16;
struct ReduceJobs {
  const float* partial[RP_MAX_JOBS];
  float* out[RP_MAX_JOBS];
  long count[RP_MAX_JOBS];
  long pitch[RP_MAX_JOBS];  // distance between consecutive partial blocks (>= count)
  int nparts[RP_MAX_JOBS];
};
__global__ void __launch_bounds__(RP_COLS * RP_GROUPS) reduce_partials_batch_kernel(ReduceJobs jobs) {
  __shared__ float red[RP_GROUPS][RP_COLS];
  cast_pdl_wait();
  cast_pdl_trigger();
  const int job = blockIdx.y;
  const float* __restrict__ partial = jobs.partial[job];
  float* __restrict__ out = jobs.out[job];
  const long count = jobs.count[job], pitch = jobs.pitch[job];
  const int nparts = jobs.nparts[job];
  const int cl = threadIdx.x % RP_COLS, g = threadIdx.x / RP_COLS;
  for (long c0 = (long)blockIdx.x * RP_COLS; c0 < count; c0 += (long)gridDim.x * RP_COLS) {
    const long c = c0 + cl;
    float acc[RP_ILP];
#pragma unroll
    for (int u = 0; u < RP_ILP; ++u) acc[u] = 0.f;
    if (c < count) {
      for (int p0 = g; p0 < nparts; p0 += RP_GROUPS * RP_ILP) {
#pragma unroll
        for (int u = 0; u < RP_ILP; ++u) {
          const int p = p0 + u * RP_GROUPS;
          if (p < nparts) acc[u] += partial[(long)p * pitch + c];
        }
      }
    }
    red[g][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (g == 0 && c < count) {
      float s = red[0][cl];
#pragma unroll
      for (int k = 1; k < RP_GROUPS; ++k) s += red[k][cl];
      out[c] = s;
    }
    __syncthreads();
  }
}

int launch_reduce_partials(const float* partial, int nparts, long count, float* out0, long split, float* out1,
                           cudaStream_t stream) {
  long g = cdiv(count, RP_COLS);
  if (g > 2368) g = 2368;
  CAST_LAUNCH(reduce_partials_kernel, dim3((unsigned)g), dim3(RP_COLS * RP_GROUPS), 0, stream, partial, nparts, count,
              out0, split, out1);
  return check_launch("reduce_partials");
}

}  // namespace cast

using namespace cast;

extern "C" int cast_reduce_partials_batch(int njobs, const float* const* partials, const int* nparts,
                                          const long* counts, const long* pitches, float* const* outs,
                                          void* stream) {
  if (njobs < 0 || (njobs > 0 && (!partials || !nparts || !counts || !outs)))
    return set_error(CAST_ERR_BAD_ARG, "reduce_partials_batch");
  for (int j0 = 0; j0 < njobs; j0 += RP_MAX_JOBS) {
    ReduceJobs jobs{};
    const int n = njobs - j0 < RP_MAX_JOBS ? njobs - j0 : RP_MAX_JOBS;
    long maxc = 1;
    for (int j = 0; j < n; ++j) {
      if (!partials[j0 + j] || !outs[j0 + j] || nparts[j0 + j] <= 0 || counts[j0 + j] <= 0)
        return set_error(CAST_ERR_BAD_ARG, "reduce_partials_batch: job");
      jobs.partial[j] = partials[j0 + j];
      jobs.out[j] = outs[j0 + j];
      jobs.count[j] = counts[j0 + j];
      jobs.pitch[j] = pitches ? pitches[j0 + j] : counts[j0 + j];
      if (jobs.pitch[j] < jobs.count[j]) return set_error(CAST_ERR_BAD_ARG, "reduce_partials_batch: pitch");
      jobs.nparts[j] = nparts[j0 + j];
      if (counts[j0 + j] > maxc) maxc = counts[j0 + j];
    }
    long g = cdiv(maxc, RP_COLS);
    if (g > 2368) g = 2368;
    CAST_LAUNCH_DEP(reduce_partials_batch_kernel, dim3((unsigned)g, (unsigned)n), dim3(RP_COLS * RP_GROUPS), 0,
                (cudaStream_t)stream, jobs);
  }
  return check_launch("reduce_partials_batch");
}

extern "C" int cast_layernorm_fwd(const float* x, const float* gamma, const float* beta, long N, int H, float eps,
                                  float* y, float* mean, float* rstd, float* xnz, float* ynz, void* stream) {
  if (!x || !gamma || !beta || !y || H <= 0 || H > 32 * LN_MAXV || N < 0)
    return set_error(CAST_ERR_BAD_ARG, "layernorm_fwd");
  if (N == 0) return CAST_OK;
#define CAST_LN_FWD(NV)                                                                                   \
  CAST_LAUNCH(layernorm_fwd_kernel<NV>, dim3((unsigned)cdiv(N, LN_WARPS)), dim3(32 * LN_WARPS), 0,        \
              (cudaStream_t)stream, x, gamma, beta, N, H, eps, y, mean, rstd, xnz, ynz)
  if (H <= 64) CAST_LN_FWD(2);
  else if (H <= 128) CAST_LN_FWD(4);
  else if (H <= 256) CAST_LN_FWD(8);
  else if (H <= 512) CAST_LN_FWD(16);
  else CAST_LN_FWD(32);
#undef CAST_LN_FWD
  return check_launch("layernorm_fwd");
}

extern "C" int cast_layernorm_bwd_parts(long N) { return (int)cdiv(N, LN_ROWS_PER_CTA); }

extern "C" size_t cast_layernorm_bwd_workspace_bytes(long N, int H) {
  return (size_t)cdiv(N, LN_ROWS_PER_CTA) * 2 * H * sizeof(float);
}

extern "C" int cast_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, long N, int H, const float* dx_add, float* dx, float* dgamma,
                                  float* dbeta, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dx || (!dgamma != !dbeta) || H <= 0 || H > 32 * LN_MAXV || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "layernorm_bwd");
  if (!workspace || workspace_bytes < cast_layernorm_bwd_workspace_bytes(N, H))
    return set_error(CAST_ERR_WORKSPACE, "layernorm_bwd: workspace too small");
  const int ncta = (int)cdiv(N, LN_ROWS_PER_CTA);
  float* partial = static_cast<float*>(workspace);
  const size_t smem = (size_t)LN_WARPS * 2 * H * sizeof(float);
#define CAST_LN_BWD(NV)                                                                                         \
  {                                                                                                             \
    auto k = layernorm_bwd_kernel<NV>;                                                                          \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    CAST_LAUNCH(k, dim3(ncta), dim3(32 * LN_WARPS), smem, (cudaStream_t)stream, dy, x, mean, rstd, gamma, N, H, \
                dx_add, dx, partial);                                                                           \
  }
  if (H <= 64) CAST_LN_BWD(2)
  else if (H <= 128) CAST_LN_BWD(4)
  else if (H <= 256) CAST_LN_BWD(8)
  else if (H <= 512) CAST_LN_BWD(16)
  else CAST_LN_BWD(32)
#undef CAST_LN_BWD
  int rc = check_launch("layernorm_bwd");
  if (rc || !dgamma) return rc;  // dgamma == dbeta == null: partials [parts][gamma H | beta H] stay in the workspace
  return launch_reduce_partials(partial, ncta, 2L * H, dgamma, H, dbeta, (cudaStream_t)stream);
}
