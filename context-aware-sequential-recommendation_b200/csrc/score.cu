// K9 (candidate mode): last-position vector . candidate rows -> logits, fused with the rank-of-target counts
// (models/sasrec.py:93-97 test_logits; util.py:318-321 rank = (-logits).argsort().argsort()[0]).
// The dot runs k = 0..H-1 in order with separately rounded multiply and add (no FMA contraction) so the oracle can
// reproduce the logits bit for bit; rank = count_greater when count_equal == 0, otherwise the host evaluates the
// reference's literal argsort expression on the returned logits (SURVEY A-12).
#include "cast_rt.cuh"

namespace cast {

__global__ void score_rank_cand_kernel(const float* __restrict__ seq_last, long ld, TableRef table,
                                       int V, int H, long U, const int* __restrict__ cand, int C,
                                       float* __restrict__ logits, int* __restrict__ cgt, int* __restrict__ ceq) {
  const int lane = threadIdx.x & 31;
  const long u = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= U) return;
  const float* s = seq_last + u * ld;
  // target logit (candidate 0), computed identically by every lane
  float l0;
  {
    const int id = cand[u * C];
    const bool live = id > 0 && id < V;
    const float* row = table.row(live ? id : 0, H);
    float acc = 0.f;
    for (int k = 0; k < H; ++k) acc = __fadd_rn(acc, __fmul_rn(s[k], live ? row[k] : 0.f));
    l0 = acc;
  }
  int gt = 0, eq = 0;
  for (int c = lane; c < C; c += 32) {
    const int id = cand[u * C + c];
    const bool live = id > 0 && id < V;
    const float* row = table.row(live ? id : 0, H);
    float acc = 0.f;
    for (int k = 0; k < H; ++k) acc = __fadd_rn(acc, __fmul_rn(s[k], live ? row[k] : 0.f));
    if (logits) logits[u * C + c] = acc;
    if (c > 0) {
      gt += acc > l0 ? 1 : 0;
      eq += acc == l0 ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
    eq += __shfl_xor_sync(0xffffffffu, eq, o);
  }
  if (lane == 0) {
    if (cgt) cgt[u] = gt;
    if (ceq) ceq[u] = eq;
  }
}

}  // namespace cast

using namespace cast;

extern "C" int cast_score_rank_cand(const float* seq_last, long ld, const float* table, int V, int H, long U,
                                    const int* cand, int C, float* logits, int* count_greater, int* count_equal,
                                    void* stream) {
  if (!seq_last || !table || !cand || V <= 0 || H <= 0 || U <= 0 || C <= 0)
    return set_error(CAST_ERR_BAD_ARG, "score_rank_cand");
  const int wpb = 4;
  CAST_LAUNCH(score_rank_cand_kernel, dim3((unsigned)cdiv(U, wpb)), dim3(32 * wpb), 0, (cudaStream_t)stream, seq_last,
              ld, table_ref(table), V, H, U, cand, C, logits, count_greater, count_equal);
  return check_launch("score_rank_cand");
}

/* cast_score_rank_cand with the item table row-sharded over the ranks of the box (TableRef, cast_rt.cuh) */
extern "C" int cast_score_rank_cand_sharded(const float* seq_last, long ld, const float* const* shards, int nshards,
                                            int V, int H, long U, const int* cand, int C, float* logits,
                                            int* count_greater, int* count_equal, void* stream) {
  if (!seq_last || !shards || nshards < 1 || !cand || V <= 0 || H <= 0 || U <= 0 || C <= 0)
    return set_error(CAST_ERR_BAD_ARG, "score_rank_cand_sharded");
  const int wpb = 4;
  CAST_LAUNCH(score_rank_cand_kernel, dim3((unsigned)cdiv(U, wpb)), dim3(32 * wpb), 0, (cudaStream_t)stream, seq_last,
              ld, table_ref(shards, nshards), V, H, U, cand, C, logits, count_greater, count_equal);
  return check_launch("score_rank_cand_sharded");
}
