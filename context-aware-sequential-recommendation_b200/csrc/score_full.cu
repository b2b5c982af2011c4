// K9 (full-catalog mode): rank of the target item among ALL items the user has not interacted with — the
// user-tile x item-tile GEMM of models/sasrec.py:93-97 (test_logits = seq_emb . item_emb^T, last position) fused with
// the rank-of-target count of util.py:318-321, with integer ranks that are exact.
//
//   canonical logit   s(u,j) = sum_k u[k]*e_j[k], k ascending, multiply and add rounded separately (fp32, no FMA) —
//                     the same arithmetic as cast_score_rank_cand and the oracle, so logits are bit-reproducible;
//   count_greater[u]  = #{ j in [1,V) : j != target(u), j not rated by u, s(u,j) >  s(u,target) }
//   count_equal[u]    = #{ ... same set ...                            s(u,j) == s(u,target) }
//
// Tensor-core path (mode 0).  The U x V x H contraction runs on the 5th-generation tensor cores: tcgen05.mma
// kind::tf32, a 128-user x 256-item fp32 accumulator tile in tensor memory, operands split in shared memory into
// tf32 hi + lo parts (3 MMAs per k-step: lo*hi + hi*lo + hi*hi, "3xTF32") so the tile is accurate to
// ~K * 2^-22 * |u||e_j|.  The epilogue (tcgen05.ld, one thread per user row) never trusts that approximation for a
// decision it could get wrong: with d = c_K * |u| * |e_j| a rigorous bound on its error, an item counts as greater
// when s~ > t + d, is dropped when s~ < t - d, and anything inside the band is re-scored with the canonical fp32 dot.
// Only integers leave the kernel (integer atomics => deterministic).  Rated items are subtracted afterwards by an
// exact gather kernel, so the tensor-core pass needs no per-user mask.
//
// Exact path (mode 1): the same counts by brute-force canonical dots (small catalogs, validation of mode 0).
#include "cast_rt.cuh"
#ifndef CAST_EMU
#include "umma.cuh"
#endif

namespace cast {

// canonical logit: sequential k, separately rounded multiply and add
__device__ __forceinline__ float canonical_dot(const float* __restrict__ u, const float* __restrict__ e, int H) {
  float acc = 0.f;
  for (int k = 0; k < H; ++k) acc = __fadd_rn(acc, __fmul_rn(u[k], e[k]));
  return acc;
}

constexpr float NORM_SLACK = 1.0001f;  // fp32 norm evaluation error, folded into the bound

// inorm[j] >= |e_j|_2 (row 0 is the zero-pad row of the lookup table: never a candidate)
__global__ void item_norm_kernel(const float* __restrict__ table, int V, int H, float* __restrict__ inorm) {
  const int lane = threadIdx.x & 31;
  const long j = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= V) return;
  float s = 0.f;
  for (int k = lane; k < H; k += 32) {
    const float x = table[j * H + k];
    s = fmaf(x, x, s);
  }
  s = warp_sum(s);
  if (lane == 0) inorm[j] = sqrtf(s) * NORM_SLACK;
}

// per user: |u|_2 bound and the canonical target score (target id outside [1,V) scores 0 like the zero-pad row)
__global__ void user_prep_kernel(const float* __restrict__ users, long ldu, const float* __restrict__ table, int V,
                                 int H, long U, const int* __restrict__ target, float* __restrict__ unorm,
                                 float* __restrict__ tscore) {
  const int lane = threadIdx.x & 31;
  const long u = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= U) return;
  const float* ur = users + u * ldu;
  float s = 0.f;
  for (int k = lane; k < H; k += 32) s = fmaf(ur[k], ur[k], s);
  s = warp_sum(s);
  if (lane == 0) {
    unorm[u] = sqrtf(s) * NORM_SLACK;
    const int t = target[u];
    tscore[u] = (t > 0 && t < V) ? canonical_dot(ur, table + (long)t * H, H) : 0.f;
  }
}

// mode 1 / emulation: brute force over an item range per CTA row (blockIdx.y = user)
__global__ void score_full_exact_kernel(const float* __restrict__ users, long ldu, const float* __restrict__ table,
                                        int V, int H, const int* __restrict__ target,
                                        const float* __restrict__ tscore, long items_per_cta, int* __restrict__ cgt,
                                        int* __restrict__ ceq) {
  CAST_DYN_SMEM(float, us);
  const long u = blockIdx.y;
  for (int k = threadIdx.x; k < H; k += blockDim.x) us[k] = users[u * ldu + k];
  __syncthreads();
  const float t = tscore[u];
  const int tid = target[u];
  long j0 = (long)blockIdx.x * items_per_cta;
  long j1 = j0 + items_per_cta < V ? j0 + items_per_cta : V;
  if (j0 < 1) j0 = 1;
  int gt = 0, eq = 0;
  for (long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    if (j == tid) continue;
    const float s = canonical_dot(us, table + j * H, H);
    gt += s > t ? 1 : 0;
    eq += s == t ? 1 : 0;
  }
  if (gt) atomicAdd(&cgt[u], gt);
  if (eq) atomicAdd(&ceq[u], eq);
}

// subtract the user's rated items (CSR, ids unique per user) that the catalog-wide pass counted
__global__ void rated_subtract_kernel(const float* __restrict__ users, long ldu, const float* __restrict__ table,
                                      int V, int H, long U, const int* __restrict__ target,
                                      const float* __restrict__ tscore, const int* __restrict__ rptr,
                                      const int* __restrict__ ridx, int* __restrict__ cgt, int* __restrict__ ceq) {
  const int lane = threadIdx.x & 31;
  const long u = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= U) return;
  const float t = tscore[u];
  const int tid = target[u];
  int gt = 0, eq = 0;
  for (int p = rptr[u] + lane; p < rptr[u + 1]; p += 32) {
    const int j = ridx[p];
    if (j < 1 || j >= V || j == tid) continue;
    const float s = canonical_dot(users + u * ldu, table + (long)j * H, H);
    gt += s > t ? 1 : 0;
    eq += s == t ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
    eq += __shfl_xor_sync(0xffffffffu, eq, o);
  }
  if (lane == 0) {
    cgt[u] -= gt;
    ceq[u] -= eq;
  }
}

#ifndef CAST_EMU
// ------------------------------------------------------------------------------------------------ tensor-core path
constexpr int SF_THREADS = 256;
constexpr int SF_M = 128;                      // users per tile  (UMMA M)
constexpr int SF_N = 256;                      // items per tile  (UMMA N) = TMEM columns
constexpr int SF_KC = 64;                      // K elements staged per chunk
constexpr int SF_A_PITCH = SF_M * 16 + 16;     // bytes between K slabs of the user tile
constexpr int SF_B_PITCH = SF_N * 16 + 16;     // bytes between K slabs of the item tile
constexpr int SF_SLABS = SF_KC / 4;
constexpr size_t SF_SMEM = 2 * (size_t)SF_SLABS * (SF_A_PITCH + SF_B_PITCH) + SF_N * sizeof(float) + 128;

struct ScoreFullArgs {
  const float* users;
  long ldu;
  const float* table;
  const float* inorm;
  const float* tscore;
  const float* unorm;
  const int* target;
  int V, H;
  long U;
  int nkc;               // K chunks of SF_KC elements (last one may be shorter)
  int kpad;              // H rounded up to 8
  long items_per_split;  // multiple of SF_N
  int nsplit;
  long nunits;
  float cbound;
  int* cgt;
  int* ceq;
  unsigned long long* stats;  // [0] band candidates re-scored exactly  (optional)
  int* err;
};

// rows [row0, row0+R) x elements [k0, k0 + 4*slabs) of a row-major fp32 matrix -> tf32 hi / lo slabs
template <int R_MAX>
__device__ __forceinline__ void stage_split(unsigned char* __restrict__ hi, unsigned char* __restrict__ lo, int pitch,
                                            const float* __restrict__ src, long ld, long row0, long rows_total, int R,
                                            int k0, int slabs, int H) {
  umma::stage_split_strided<SF_THREADS, R_MAX, SF_SLABS>(hi, lo, pitch, src, ld, 1, row0, rows_total, R, k0, H, slabs);
}

__global__ void __launch_bounds__(SF_THREADS, 1) score_full_umma_kernel(ScoreFullArgs a) {
  extern __shared__ __align__(128) unsigned char sf_smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  unsigned char* Ahi = sf_smem;
  unsigned char* Alo = Ahi + SF_SLABS * SF_A_PITCH;
  unsigned char* Bhi = Alo + SF_SLABS * SF_A_PITCH;
  unsigned char* Blo = Bhi + SF_SLABS * SF_B_PITCH;
  float* nrm = reinterpret_cast<float*>(Blo + SF_SLABS * SF_B_PITCH);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, SF_N);
  if (t == 0) umma::mbar_init(&mbar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma::idesc_tf32(SF_M, SF_N);
  const uint32_t q = warp & 3, half = warp >> 2;  // TMEM lane quarter this warp may read; column half it counts
  uint32_t parity = 0;
  bool failed = false;

  for (long unit = blockIdx.x; unit < a.nunits && !failed; unit += gridDim.x) {
    const long ut = unit / a.nsplit;
    const int sp = (int)(unit - ut * a.nsplit);
    const long u0 = ut * SF_M;
    const long jb = (long)sp * a.items_per_split;
    const long je = jb + a.items_per_split < a.V ? jb + a.items_per_split : a.V;
    const long me = u0 + q * 32 + lane;  // the user row whose accumulator lane this thread reads
    const bool live = me < a.U;
    const float tsc = live ? a.tscore[me] : 0.f;
    const float nu = live ? a.unorm[me] : 0.f;
    const int tgt = live ? a.target[me] : -1;
    int gt = 0, eq = 0;
    unsigned band = 0;
    if (a.nkc == 1) stage_split<SF_M>(Ahi, Alo, SF_A_PITCH, a.users, a.ldu, u0, a.U, SF_M, 0, a.kpad / 4, a.H);
    for (long j0 = jb; j0 < je && !failed; j0 += SF_N) {
      for (int kc = 0; kc < a.nkc; ++kc) {
        const int k0 = kc * SF_KC;
        const int slabs = (a.kpad - k0 < SF_KC ? a.kpad - k0 : SF_KC) / 4;
        if (a.nkc > 1) stage_split<SF_M>(Ahi, Alo, SF_A_PITCH, a.users, a.ldu, u0, a.U, SF_M, k0, slabs, a.H);
        stage_split<SF_N>(Bhi, Blo, SF_B_PITCH, a.table, a.H, j0, a.V, SF_N, k0, slabs, a.H);
        if (kc == 0)
          for (int c = t; c < SF_N; c += SF_THREADS) nrm[c] = (j0 + c < a.V) ? a.inorm[j0 + c] : 0.f;
        umma::fence_smem_to_async();
        __syncthreads();
        if (t == 0) {
          umma::fence_after_sync();
          const uint32_t ah = umma::smem_u32(Ahi), al = umma::smem_u32(Alo);
          const uint32_t bh = umma::smem_u32(Bhi), bl = umma::smem_u32(Blo);
          for (int s = 0; s < slabs / 2; ++s) {
            const uint64_t dah = umma::smem_desc(ah + 2 * s * SF_A_PITCH, SF_A_PITCH, 128);
            const uint64_t dal = umma::smem_desc(al + 2 * s * SF_A_PITCH, SF_A_PITCH, 128);
            const uint64_t dbh = umma::smem_desc(bh + 2 * s * SF_B_PITCH, SF_B_PITCH, 128);
            const uint64_t dbl = umma::smem_desc(bl + 2 * s * SF_B_PITCH, SF_B_PITCH, 128);
            umma::mma_tf32(tmem, dal, dbh, idesc, (kc > 0 || s > 0) ? 1u : 0u);
            umma::mma_tf32(tmem, dah, dbl, idesc, 1u);
            umma::mma_tf32(tmem, dah, dbh, idesc, 1u);
          }
          umma::mma_commit(&mbar);
        }
        // the MMAs read shared memory asynchronously: nobody restages (or reads the accumulator) before they finish
        const bool ok = umma::mbar_wait(&mbar, parity);
        parity ^= 1u;
        if (!__syncthreads_and(ok ? 1 : 0)) {
          failed = true;
          break;
        }
        umma::fence_after_sync();
      }
      if (failed) break;
      // ---- epilogue: thread <-> user row (TMEM lane), 128 of the 256 columns per warp half
#pragma unroll 1
      for (int cb = 0; cb < SF_N / 2; cb += 32) {
        const int col0 = (int)half * (SF_N / 2) + cb;
        float v[32];
        umma::tmem_ld32(tmem + ((q * 32u) << 16) + (uint32_t)col0, v);
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const long item = j0 + col0 + e;
            if (item < 1 || item >= je || item == tgt) continue;
            const float d = a.cbound * nu * nrm[col0 + e];
            const float s = v[e];
            if (s > tsc + d) {
              ++gt;
            } else if (s >= tsc - d) {  // inside the error band: decide with the canonical fp32 logit
              const float ex = canonical_dot(a.users + me * a.ldu, a.table + item * a.H, a.H);
              gt += ex > tsc ? 1 : 0;
              eq += ex == tsc ? 1 : 0;
              ++band;
            }
          }
        }
      }
      umma::fence_before_sync();
      __syncthreads();  // accumulator fully read before the next tile's first MMA overwrites it
      umma::fence_after_sync();
    }
    if (live && !failed) {
      if (gt) atomicAdd(&a.cgt[me], gt);
      if (eq) atomicAdd(&a.ceq[me], eq);
      if (a.stats && band) atomicAdd(&a.stats[0], (unsigned long long)band);
    }
  }
  if (failed && t == 0) atomicExch(a.err, 1);
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, SF_N);
}
#endif  // !CAST_EMU

}  // namespace cast

using namespace cast;

static inline size_t sf_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t cast_score_rank_full_workspace_bytes(long U, int V) {
  return sf_align((size_t)V * 4) + 2 * sf_align((size_t)U * 4) + 256;
}

extern "C" int cast_score_rank_full(const float* seq_last, long ld, const float* table, int V, int H, long U,
                                    const int* target, const int* rated_ptr, const int* rated_idx, int mode,
                                    int* count_greater, int* count_equal, unsigned long long* stats, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!seq_last || !table || !target || !count_greater || !count_equal || V <= 1 || H <= 0 || U <= 0 ||
      (mode != 0 && mode != 1) || (rated_ptr && !rated_idx))
    return set_error(CAST_ERR_BAD_ARG, "score_rank_full");
  if (!workspace || workspace_bytes < cast_score_rank_full_workspace_bytes(U, V))
    return set_error(CAST_ERR_WORKSPACE, "score_rank_full: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* inorm = reinterpret_cast<float*>(w);
  float* tscore = reinterpret_cast<float*>(w + sf_align((size_t)V * 4));
  float* unorm = reinterpret_cast<float*>(w + sf_align((size_t)V * 4) + sf_align((size_t)U * 4));
  int* err = reinterpret_cast<int*>(w + sf_align((size_t)V * 4) + 2 * sf_align((size_t)U * 4));
  int rc;
  cudaMemsetAsync(count_greater, 0, (size_t)U * sizeof(int), st);
  cudaMemsetAsync(count_equal, 0, (size_t)U * sizeof(int), st);
  cudaMemsetAsync(err, 0, sizeof(int), st);
  if (stats) cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), st);
  CAST_LAUNCH(user_prep_kernel, dim3((unsigned)cdiv(U, 4)), dim3(128), 0, st, seq_last, ld, table, V, H, U, target,
              unorm, tscore);
  if ((rc = check_launch("score_full(user_prep)"))) return rc;
#ifdef CAST_EMU
  mode = 1;  // the host emulation has no tensor cores; the exact path defines the same integers
#endif
  if (mode == 1) {
    const long per = 2048;
    CAST_LAUNCH(score_full_exact_kernel, dim3((unsigned)cdiv(V, per), (unsigned)U), dim3(128), H * sizeof(float), st,
                seq_last, ld, table, V, H, target, (const float*)tscore, per, count_greater, count_equal);
    if ((rc = check_launch("score_full(exact)"))) return rc;
  } else {
#ifndef CAST_EMU
    CAST_LAUNCH(item_norm_kernel, dim3((unsigned)cdiv(V, 8)), dim3(256), 0, st, table, V, H, inorm);
    if ((rc = check_launch("score_full(item_norm)"))) return rc;
    ScoreFullArgs a;
    a.users = seq_last; a.ldu = ld; a.table = table; a.inorm = inorm; a.tscore = tscore; a.unorm = unorm;
    a.target = target; a.V = V; a.H = H; a.U = U;
    a.kpad = (H + 7) & ~7;
    a.nkc = (int)cdiv(a.kpad, SF_KC);
    const long utiles = cdiv(U, SF_M);
    const long ntile_items = cdiv(V, SF_N);
    long nsplit = cdiv(148L * 2, utiles);          // ~2 work units per SM
    if (nsplit > ntile_items) nsplit = ntile_items;
    if (nsplit < 1) nsplit = 1;
    a.items_per_split = cdiv(ntile_items, nsplit) * SF_N;
    a.nsplit = (int)cdiv(V, a.items_per_split);
    a.nunits = utiles * a.nsplit;
    // |s~ - s| <= cbound * |u| * |e_j|: three dropped/rounded split terms (3 * 2^-22) plus one accumulation error of at
    // most 2^-22 * sum|u_k e_k| per MMA (3 * kpad/8 MMAs per tile); doubled for slack
    a.cbound = 2.0f * (3.0f + 3.0f * (float)a.kpad / 8.0f) * 2.384185791015625e-07f;
    a.cgt = count_greater; a.ceq = count_equal; a.stats = stats; a.err = err;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(score_full_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_SMEM);
      configured = true;
    }
    const long grid = a.nunits < 148 ? a.nunits : 148;
    score_full_umma_kernel<<<dim3((unsigned)grid), dim3(SF_THREADS), SF_SMEM, st>>>(a);
    if ((rc = check_launch("score_full(umma)"))) return rc;
#endif
  }
  if (rated_ptr) {
    CAST_LAUNCH(rated_subtract_kernel, dim3((unsigned)cdiv(U, 4)), dim3(128), 0, st, seq_last, ld, table, V, H, U,
                target, (const float*)tscore, rated_ptr, rated_idx, count_greater, count_equal);
    if ((rc = check_launch("score_full(rated)"))) return rc;
  }
  return CAST_OK;
}

// 0 = ok; non-zero = the tensor-core pass timed out waiting for its MMAs (never expected; results invalid)
extern "C" int cast_score_rank_full_status(const void* workspace, long U, int V, int* host_flag, void* stream) {
  if (!workspace || !host_flag) return set_error(CAST_ERR_BAD_ARG, "score_rank_full_status");
  const unsigned char* w = static_cast<const unsigned char*>(workspace);
  const int* err = reinterpret_cast<const int*>(w + sf_align((size_t)V * 4) + 2 * sf_align((size_t)U * 4));
#ifdef CAST_EMU
  *host_flag = *err;
  (void)stream;
#else
  if (cudaMemcpyAsync(host_flag, err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
    return set_error(CAST_ERR_CUDA, "score_rank_full_status");
  cudaStreamSynchronize((cudaStream_t)stream);
#endif
  return CAST_OK;
}
