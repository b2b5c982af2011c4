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
// (sasrec.py:87-115) -> d(seq_emb) -> LayerNorm backward.  A warp owns a row end to end (lanes own columns c, c+32), so
// the normalised row, its gradient and the LN-backward row sums never leave registers; per-CTA partials of the loss
// sums and of dgamma / dbeta are left for the step's batched reduction.
__global__ void __launch_bounds__(32 * LOSS_WARPS)
lnf_loss_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                TableRef table, int V, int H, long N, const int* __restrict__ pos,
                const int* __restrict__ neg, float* __restrict__ seq, float* __restrict__ pos_logits,
                float* __restrict__ neg_logits, float* __restrict__ gpos, float* __restrict__ gneg,
                float* __restrict__ dx, float* __restrict__ partial_loss, float* __restrict__ partial_ln) {
  __shared__ float red[LOSS_WARPS][3];
  __shared__ float redg[LOSS_WARPS][2][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long row0 = (long)blockIdx.x * LOSS_ROWS_PER_CTA;
  const int c0 = lane, c1 = lane + 32;
  const bool h0 = c0 < H, h1 = c1 < H;
  const float g0 = h0 ? gamma[c0] : 0.f, g1 = h1 ? gamma[c1] : 0.f;
  const float b0 = h0 ? beta[c0] : 0.f, b1 = h1 ? beta[c1] : 0.f;
  float s_loss = 0.f, s_auc = 0.f, s_cnt = 0.f;
  float dg0 = 0.f, dg1 = 0.f, db0 = 0.f, db1 = 0.f;
  for (int r = warp; r < LOSS_ROWS_PER_CTA; r += LOSS_WARPS) {
    const long n = row0 + r;
    if (n >= N) break;
    const int pi = pos[n], ni = neg[n];
    const bool pl = pi > 0 && pi < V, nl = ni > 0 && ni < V;
    const float* prow = table.row(pl ? pi : 0, H);
    const float* nrow = table.row(nl ? ni : 0, H);
    const float* xr = x + n * H;
    const float v0 = h0 ? xr[c0] : 0.f, v1 = h1 ? xr[c1] : 0.f;
    const float p0 = (pl && h0) ? prow[c0] : 0.f, p1 = (pl && h1) ? prow[c1] : 0.f;
    const float q0 = (nl && h0) ? nrow[c0] : 0.f, q1 = (nl && h1) ? nrow[c1] : 0.f;
    // ---- LayerNorm (biased variance, eps inside the sqrt)
    const float mean = warp_sum(v0 + v1) / (float)H;
    const float d0 = h0 ? v0 - mean : 0.f, d1 = h1 ? v1 - mean : 0.f;
    const float var = warp_sum(d0 * d0 + d1 * d1) / (float)H;
    const float stdv = sqrtf(var + eps);
    const float rs = 1.0f / stdv;
    const float xh0 = d0 * rs, xh1 = d1 * rs;
    const float y0 = h0 ? g0 * xh0 + b0 : 0.f, y1 = h1 ? g1 * xh1 + b1 : 0.f;
    if (h0) seq[n * H + c0] = y0;
    if (h1) seq[n * H + c1] = y1;
    // ---- logits, loss terms, logit gradients (same expressions as logits_loss_kernel)
    const float dp = warp_sum(fmaf(p1, y1, p0 * y0));
    const float dn = warp_sum(fmaf(q1, y1, q0 * y0));
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
      gpos[n] = gp;
      gneg[n] = gn;
      s_loss += lterm;
      s_auc += ((sg + 1.0f) * 0.5f) * ist;
      s_cnt += ist;
    }
    // ---- d(seq_emb) and LayerNorm backward of the row
    const float e0 = fmaf(gn, q0, gp * p0), e1 = fmaf(gn, q1, gp * p1);
    const float a0 = e0 * g0, a1 = e1 * g1;
    const float s1 = warp_sum(a0 + a1) / (float)H;
    const float s2 = warp_sum(a0 * xh0 + a1 * xh1) / (float)H;
    if (h0) dx[n * H + c0] = rs * (a0 - s1 - xh0 * s2);
    if (h1) dx[n * H + c1] = rs * (a1 - s1 - xh1 * s2);
    dg0 = fmaf(e0, xh0, dg0);
    dg1 = fmaf(e1, xh1, dg1);
    db0 += e0;
    db1 += e1;
  }
  if (lane == 0) {
    red[warp][0] = s_loss;
    red[warp][1] = s_auc;
    red[warp][2] = s_cnt;
  }
  redg[warp][0][c0] = dg0;
  redg[warp][0][c1] = dg1;
  redg[warp][1][c0] = db0;
  redg[warp][1][c1] = db1;
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < LOSS_WARPS; ++w) s += red[w][threadIdx.x];
    partial_loss[(long)blockIdx.x * 3 + threadIdx.x] = s;
  }
  if (threadIdx.x < 128) {  // fixed-order fold over the warps: threads 0..63 -> dgamma[c], 64..127 -> dbeta[c]
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    if (c < H) {
      float s = 0.f;
      for (int w = 0; w < LOSS_WARPS; ++w) s += redg[w][which][c];
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
  CAST_LAUNCH(lnf_loss_kernel, dim3(ncta), dim3(32 * LOSS_WARPS), 0, (cudaStream_t)stream, x, gamma, beta, eps,
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
