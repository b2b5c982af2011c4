// K6: positive / negative logits, literal BCE, AUC and the gradient w.r.t. the sequence embedding
// (models/sasrec.py:87-115).  One warp per position; the pos/neg rows are gathered from the zero-padded table
// (row 0 = zeros, unscaled: sasrec.py:89-90 look up `item_emb_table` directly).  Loss terms follow the
// reference literally:  -log(sigmoid(x)+1e-24) and -log(1-sigmoid(x)+1e-24)  (not softplus) and so do their
// derivatives  -s(1-s)/(s+1e-24)  and  s(1-s)/(1-s+1e-24).  Sums are un-normalised; two fixed-order stages.
#include "cast_rt.cuh"

namespace cast {

constexpr int LOSS_WARPS = 8;
constexpr int LOSS_ROWS_PER_CTA = 64;

__global__ void logits_loss_kernel(const float* __restrict__ seq, const float* __restrict__ table, int V, int H,
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
    const float* prow = table + (long)(pl ? pi : 0) * H;
    const float* nrow = table + (long)(nl ? ni : 0) * H;
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

}  // namespace cast

using namespace cast;

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
  CAST_LAUNCH(logits_loss_kernel, dim3(ncta), dim3(32 * LOSS_WARPS), 0, (cudaStream_t)stream, seq_emb, table, V, H, N,
              pos, neg, pos_logits, neg_logits, dseq, gpos, gneg, partial);
  int rc = check_launch("logits_loss");
  if (rc || !sums) return rc;  // sums == null: the [parts][3] partials stay in the workspace (cast_reduce_partials_batch)
  CAST_LAUNCH(loss_final_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, partial, ncta, sums);
  return check_launch("loss_final");
}
