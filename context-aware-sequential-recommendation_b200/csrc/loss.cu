// K6: positive / negative logits, literal BCE, AUC and the gradient w.r.t. the sequence embedding
// (models/sasrec.py:87-115).  One warp per position; the pos/neg rows are gathered from the zero-padded table
// (row 0 = zeros, unscaled: sasrec.py:89-90 look up `item_emb_table` directly).  Loss terms follow the
// reference literally:  -log(sigmoid(x)+1e-24) and -log(1-sigmoid(x)+1e-24)  (not softplus) and so do their
// derivatives  -s(1-s)/(s+1e-24)  and  s(1-s)/(1-s+1e-24).  Sums are un-normalised; two fixed-order stages.
#include "cast_rt.cuh"

namespace cast {

constexpr int LOSS_WARPS = 8;
constexpr int LOSS_ROWS_PER_CTA = 64;

__global__ void logits_loss_kernel(const float* __restrict__ seq, TableRef table, int V, int H,
                                   long N, const int* __restrict__ pos, const int* __restrict__ neg,
                                   float* __restrict__ pos_logits, float* __restrict__ neg_logits,
                                   float* __restrict__ dseq, float* __restrict__ gpos, float* __restrict__ gneg,
                                   float* __restrict__ partial) {
  __shared__ float red[LOSS_WARPS][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long row0 = (long)blockIdx.x * LOSS_ROWS_PER_CTA;
  float s_loss = 0.f, s_auc = 0.f, s_cnt = 0.f;
  for (int r = warp; r < LOSS_ROWS_PER_CTA; r += LOSS_WARPS) {
    const long n = row0 + r;
    if (n >= N) break;
    const int pi = pos[n], ni = neg[n];
    const bool pl = pi > 0 && pi < V, nl = ni > 0 && ni < V;
    const float* prow = table.row(pl ? pi : 0, H);
    const float* nrow = table.row(nl ? ni : 0, H);
    const float* srow = seq + n * H;
    float dp = 0.f, dn = 0.f;
    for (int c = lane; c < H; c += 32) {
      const float s = srow[c];
      if (pl) dp = fmaf(prow[c], s, dp);
      if (nl) dn = fmaf(nrow[c], s, dn);
    }
    dp = warp_sum(dp);
    dn = warp_sum(dn);
    const float ist = pi != 0 ? 1.f : 0.f;
    const float sp = 1.0f / (1.0f + expf(-dp));
    const float sn = 1.0f / (1.0f + expf(-dn));
    const float lterm = (-logf(sp + 1e-24f) - logf(1.0f - sn + 1e-24f)) * ist;
    const float df = dp - dn;
    const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    const float gp = -ist * sp * (1.0f - sp) / (sp + 1e-24f);
    const float gn = ist * sn * (1.0f - sn) / (1.0f - sn + 1e-24f);
    if (lane == 0) {
      if (pos_logits) pos_logits[n] = dp;
      if (neg_logits) neg_logits[n] = dn;
      if (gpos) gpos[n] = gp;
      if (gneg) gneg[n] = gn;
      s_loss += lterm;
      s_auc += ((sg + 1.0f) * 0.5f) * ist;
      s_cnt += ist;
    }
    if (dseq) {
      for (int c = lane; c < H; c += 32) {
        float g = 0.f;
        if (pl) g = gp * prow[c];
        if (nl) g = fmaf(gn, nrow[c], g);
        dseq[n * H + c] = g;
      }
    }
  }
  if (lane == 0) {
    red[warp][0] = s_loss;
    red[warp][1] = s_auc;
    red[warp][2] = s_cnt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < LOSS_WARPS; ++w) s += red[w][threadIdx.x];
    partial[(long)blockIdx.x * 3 + threadIdx.x] = s;
  }
}

__global__ void loss_final_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ sums) {
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(long)p * 3 + threadIdx.x];
    sums[threadIdx.x] = s;
  }
}

// Training tail in one pass (H <= 64): final LayerNorm of the main tower (modules.py:53-80) -> pos/neg logits, BCE, AUC
// (sasrec.py:87-115) -> d(seq_emb) -> LayerNorm backward.  Eight threads own a row end to end (thread `sub` holds the
// column pairs (j * 8 + sub) * 2, j < 4: a row's threads read 64 contiguous bytes per request), so the normalised row,
// its gradient and the LN-backward row sums never leave registers and the six row sums are 3-step shuffles over 8 lanes;
// a warp works on four rows at once.  Per-CTA partials of the loss sums and of dgamma / dbeta (fixed order over the
// CTA's 32 row slots) are left for the step's batched reduction.
constexpr int LNF_TPR = 8, LNF_NJ = 4, LNF_THREADS = 256, LNF_SLOTS = LNF_THREADS / LNF_TPR;
__device__ __forceinline__ float sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
// the pair (p[c], p[c + 1]) with c even; elements at or past H (or a dead row) read as 0
__device__ __forceinline__ void lnf_ld2(const float* __restrict__ p, int c, int H, bool on, bool vec, float& a, float& b) {
  a = b = 0.f;
  if (!on || c >= H) return;
  if (vec) {   // H even => c + 1 < H
    const float2 v = *reinterpret_cast<const float2*>(p + c);
    a = v.x;
    b = v.y;
  } else {
    a = p[c];
    if (c + 1 < H) b = p[c + 1];
  }
}
__device__ __forceinline__ void lnf_st2(float* __restrict__ p, int c, int H, bool vec, float a, float b) {
  if (c >= H) return;
  if (vec) {
    *reinterpret_cast<float2*>(p + c) = make_float2(a, b);
  } else {
    p[c] = a;
    if (c + 1 < H) p[c + 1] = b;
  }
}

__global__ void __launch_bounds__(LNF_THREADS)
lnf_loss_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                TableRef table, int V, int H, long N, const int* __restrict__ pos,
                const int* __restrict__ neg, float* __restrict__ seq, float* __restrict__ pos_logits,
                float* __restrict__ neg_logits, float* __restrict__ gpos, float* __restrict__ gneg,
                float* __restrict__ dx, float* __restrict__ partial_loss, float* __restrict__ partial_ln) {
  __shared__ float red[LNF_SLOTS][3];
  __shared__ float redg[LNF_SLOTS][2][64];
  cast_pdl_wait();
  cast_pdl_trigger();
  const int t = threadIdx.x, slot = t / LNF_TPR, sub = t % LNF_TPR;
  const long row0 = (long)blockIdx.x * LOSS_ROWS_PER_CTA;
  const bool heven = (H & 1) == 0;
  const bool vx = heven && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(seq) |
                             reinterpret_cast<uintptr_t>(dx)) & 7) == 0;
  float gam[LNF_NJ][2], bet[LNF_NJ][2];
#pragma unroll
  for (int j = 0; j < LNF_NJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = (j * LNF_TPR + sub) * 2 + e;
      gam[j][e] = c < H ? gamma[c] : 0.f;
      bet[j][e] = c < H ? beta[c] : 0.f;
    }
  float s_loss = 0.f, s_auc = 0.f, s_cnt = 0.f;
  float dg[LNF_NJ][2], db[LNF_NJ][2];
#pragma unroll
  for (int j = 0; j < LNF_NJ; ++j) dg[j][0] = dg[j][1] = db[j][0] = db[j][1] = 0.f;
  const float invH = 1.0f / (float)H;
  for (int r = slot; r < LOSS_ROWS_PER_CTA; r += LNF_SLOTS) {   // (no early exit: the shuffles need every lane)
    const long n = row0 + r;
    const bool on = n < N;
    const long nn = on ? n : 0;
    const int pi = on ? pos[nn] : 0, ni = on ? neg[nn] : 0;
    const bool pl = pi > 0 && pi < V, nl = ni > 0 && ni < V;
    const float* prow = table.row(pl ? pi : 0, H);
    const float* nrow = table.row(nl ? ni : 0, H);
    const bool vt = heven && ((reinterpret_cast<uintptr_t>(prow) | reinterpret_cast<uintptr_t>(nrow)) & 7) == 0;
    const float* xr = x + nn * H;
    float v[LNF_NJ][2], p[LNF_NJ][2], q[LNF_NJ][2];
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j) {
      const int c = (j * LNF_TPR + sub) * 2;
      lnf_ld2(xr, c, H, on, vx, v[j][0], v[j][1]);
      lnf_ld2(prow, c, H, pl, vt, p[j][0], p[j][1]);
      lnf_ld2(nrow, c, H, nl, vt, q[j][0], q[j][1]);
    }
    // ---- LayerNorm (biased variance, eps inside the sqrt)
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j) acc += v[j][0] + v[j][1];
    const float mean = sum8(acc) * invH;
    float xh[LNF_NJ][2], y[LNF_NJ][2];
    acc = 0.f;
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = (j * LNF_TPR + sub) * 2 + e;
        xh[j][e] = c < H ? v[j][e] - mean : 0.f;
        acc = fmaf(xh[j][e], xh[j][e], acc);
      }
    const float var = sum8(acc) * invH;
    const float stdv = sqrtf(var + eps);
    const float rs = 1.0f / stdv;
    float adp = 0.f, adn = 0.f;
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = (j * LNF_TPR + sub) * 2 + e;
        xh[j][e] *= rs;
        y[j][e] = c < H ? fmaf(gam[j][e], xh[j][e], bet[j][e]) : 0.f;
        adp = fmaf(p[j][e], y[j][e], adp);
        adn = fmaf(q[j][e], y[j][e], adn);
      }
      if (on) lnf_st2(seq + nn * H, (j * LNF_TPR + sub) * 2, H, vx, y[j][0], y[j][1]);
    }
    // ---- logits, loss terms, logit gradients (same expressions as logits_loss_kernel)
    const float dp = sum8(adp), dn = sum8(adn);
    const float ist = pi != 0 ? 1.f : 0.f;
    const float sp = 1.0f / (1.0f + expf(-dp));
    const float sn = 1.0f / (1.0f + expf(-dn));
    const float lterm = (-logf(sp + 1e-24f) - logf(1.0f - sn + 1e-24f)) * ist;
    const float df = dp - dn;
    const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    const float gp = -ist * sp * (1.0f - sp) / (sp + 1e-24f);
    const float gn = ist * sn * (1.0f - sn) / (1.0f - sn + 1e-24f);
    if (sub == 0 && on) {
      if (pos_logits) pos_logits[n] = dp;
      if (neg_logits) neg_logits[n] = dn;
      gpos[n] = gp;
      gneg[n] = gn;
      s_loss += lterm;
      s_auc += ((sg + 1.0f) * 0.5f) * ist;
      s_cnt += ist;
    }
    // ---- d(seq_emb) and LayerNorm backward of the row
    float ev[LNF_NJ][2], av[LNF_NJ][2];
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        ev[j][e] = fmaf(gn, q[j][e], gp * p[j][e]);
        av[j][e] = ev[j][e] * gam[j][e];
        a1 += av[j][e];
        a2 = fmaf(av[j][e], xh[j][e], a2);
      }
    const float s1 = sum8(a1) * invH, s2 = sum8(a2) * invH;
#pragma unroll
    for (int j = 0; j < LNF_NJ; ++j) {
      if (on)
        lnf_st2(dx + nn * H, (j * LNF_TPR + sub) * 2, H, vx, rs * (av[j][0] - s1 - xh[j][0] * s2),
                rs * (av[j][1] - s1 - xh[j][1] * s2));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (on) {
          dg[j][e] = fmaf(ev[j][e], xh[j][e], dg[j][e]);
          db[j][e] += ev[j][e];
        }
      }
    }
  }
  if (sub == 0) {
    red[slot][0] = s_loss;
    red[slot][1] = s_auc;
    red[slot][2] = s_cnt;
  }
#pragma unroll
  for (int j = 0; j < LNF_NJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = (j * LNF_TPR + sub) * 2 + e;
      redg[slot][0][c] = dg[j][e];
      redg[slot][1][c] = db[j][e];
    }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < LNF_SLOTS; ++w) s += red[w][threadIdx.x];
    partial_loss[(long)blockIdx.x * 3 + threadIdx.x] = s;
  }
  if (threadIdx.x < 128) {  // fixed-order fold over the row slots: threads 0..63 -> dgamma[c], 64..127 -> dbeta[c]
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    if (c < H) {
      float s = 0.f;
      for (int w = 0; w < LNF_SLOTS; ++w) s += redg[w][which][c];
      partial_ln[(long)blockIdx.x * 2 * H + which * H + c] = s;
    }
  }
}

}  // namespace cast

using namespace cast;

extern "C" int cast_lnf_loss_parts(long N) { return (int)cdiv(N, LOSS_ROWS_PER_CTA); }
extern "C" size_t cast_lnf_loss_workspace_bytes(long N, int H) {
  return (size_t)cdiv(N, LOSS_ROWS_PER_CTA) * (3 + 2 * (size_t)H) * sizeof(float);
}

extern "C" int cast_lnf_loss(const float* x, const float* gamma, const float* beta, float eps, const float* table, int V,
                             int H, long N, const int* pos, const int* neg, float* seq_emb, float* pos_logits,
                             float* neg_logits, float* gpos, float* gneg, float* dx, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!x || !gamma || !beta || !table || !pos || !neg || !seq_emb || !gpos || !gneg || !dx || V <= 0 || H <= 0 || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "lnf_loss");
  if (H > 64) return set_error(CAST_ERR_UNSUPPORTED, "lnf_loss: H > 64");
  if (!workspace || workspace_bytes < cast_lnf_loss_workspace_bytes(N, H))
    return set_error(CAST_ERR_WORKSPACE, "lnf_loss: workspace too small");
  const int ncta = (int)cdiv(N, LOSS_ROWS_PER_CTA);
  float* pl = static_cast<float*>(workspace);
  CAST_LAUNCH_DEP(lnf_loss_kernel, dim3(ncta), dim3(LNF_THREADS), 0, (cudaStream_t)stream, x, gamma, beta, eps,
              table_ref(table), V, H, N, pos, neg, seq_emb, pos_logits, neg_logits, gpos, gneg, dx, pl, pl + (size_t)ncta * 3);
  return check_launch("lnf_loss");
}

extern "C" int cast_logits_loss_parts(long N) { return (int)cdiv(N, LOSS_ROWS_PER_CTA); }

extern "C" size_t cast_logits_loss_workspace_bytes(long N) {
  return (size_t)cdiv(N, LOSS_ROWS_PER_CTA) * 3 * sizeof(float);
}

extern "C" int cast_logits_loss(const float* seq_emb, const float* table, int V, int H, long N, const int* pos,
                                const int* neg, float* pos_logits, float* neg_logits, float* sums, float* dseq,
                                float* gpos, float* gneg, void* workspace, size_t workspace_bytes, void* stream) {
  if (!seq_emb || !table || !pos || !neg || V <= 0 || H <= 0 || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "logits_loss");
  if (!workspace || workspace_bytes < cast_logits_loss_workspace_bytes(N))
    return set_error(CAST_ERR_WORKSPACE, "logits_loss: workspace too small");
  const int ncta = (int)cdiv(N, LOSS_ROWS_PER_CTA);
  float* partial = static_cast<float*>(workspace);
  CAST_LAUNCH(logits_loss_kernel, dim3(ncta), dim3(32 * LOSS_WARPS), 0, (cudaStream_t)stream, seq_emb,
              table_ref(table), V, H, N, pos, neg, pos_logits, neg_logits, dseq, gpos, gneg, partial);
  int rc = check_launch("logits_loss");
  if (rc || !sums) return rc;  // sums == null: the [parts][3] partials stay in the workspace (cast_reduce_partials_batch)
  CAST_LAUNCH(loss_final_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, partial, ncta, sums);
  return check_launch("loss_final");
}

/* cast_logits_loss with the item table row-sharded over the ranks of the box (TableRef, cast_rt.cuh) */
extern "C" int cast_logits_loss_sharded(const float* seq_emb, const float* const* shards, int nshards, int V, int H,
                                        long N, const int* pos, const int* neg, float* pos_logits, float* neg_logits,
                                        float* sums, float* dseq, float* gpos, float* gneg, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  if (!seq_emb || !shards || nshards < 1 || !pos || !neg || V <= 0 || H <= 0 || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "logits_loss_sharded");
  if (!workspace || workspace_bytes < cast_logits_loss_workspace_bytes(N))
    return set_error(CAST_ERR_WORKSPACE, "logits_loss_sharded: workspace too small");
  const int ncta = (int)cdiv(N, LOSS_ROWS_PER_CTA);
  float* partial = static_cast<float*>(workspace);
  CAST_LAUNCH(logits_loss_kernel, dim3(ncta), dim3(32 * LOSS_WARPS), 0, (cudaStream_t)stream, seq_emb,
              table_ref(shards, nshards), V, H, N, pos, neg, pos_logits, neg_logits, dseq, gpos, gneg, partial);
  int rc = check_launch("logits_loss_sharded");
  if (rc || !sums) return rc;
  CAST_LAUNCH(loss_final_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, partial, ncta, sums);
  return check_launch("loss_final");
}
