// K4: causal multi-head attention with the reference's exact mask semantics (modules.py:208-269), forward and
// backward, scores never leaving shared memory.
//
//   S = Q K^T / sqrt(d)                                  :216-219
//   S = where(sum_H(keys)==0, -2^32+1, S)                :222-228   key mask from the *tensor*
//   S = where(k > q, -2^32+1, S)                         :232-241   causality
//   P = softmax(S) over ALL T keys                       :244       fully-masked rows => uniform 1/T, future incl.
//   P *= sign(|sum_H(queries)|)                          :248-253   query mask (queries = LN(x))
//   P = dropout(P)                                       :256       attention_weights captured here (:259)
//   O = P V ; heads merged ; O += queries                :262-269
//
// Layout: CTA = (query tile of 16*RI rows, head, batch element), 256 threads as a 16x16 grid, each thread a
// RI x RJ register micro-tile of the score chunk (RJ*16 keys per chunk).  Q/K/V tiles sit row-major in smem with
// a row stride DS (DS/4 odd => the 16 key rows a quarter-warp reads are bank-disjoint), reductions run along
// d with LDS.128 on both operands, the whole [tile, T] score block stays in smem (T <= ~800), softmax is one
// warp per row.  Backward is two deterministic kernels (no float atomics): dQ per query tile, dK/dV per key tile,
// both recomputing P from the saved row max / 1/rowsum and D_i = sum_j P_ij dP_ij.
#pragma once
#include "cast_rt.cuh"
#include "tile_ops.cuh"

namespace cast {

constexpr int ATT_THREADS = 256;

struct AttnDims {
  int B, T, H, h, d;  // d = H / h
  int dpad, DS;       // d rounded to 4; smem row stride
  int Tp4, PS;        // T rounded to 4; score row stride
  float inv_sqrt_d;
};

static inline AttnDims make_dims(int B, int T, int H, int h) {
  AttnDims a;
  a.B = B; a.T = T; a.H = H; a.h = h; a.d = H / h;
  a.dpad = (a.d + 3) & ~3;
  a.DS = ((a.dpad / 4) % 2 == 0) ? a.dpad + 4 : a.dpad;
  a.Tp4 = (T + 3) & ~3;
  a.PS = a.Tp4 + 4;
  a.inv_sqrt_d = 1.0f / sqrtf((float)a.d);  // outputs / (K_.get_shape()[-1] ** 0.5), as a multiply
  return a;
}

// dst[r][c] (stride DS) = src[(row0 + r) * ld + c] for r < nrows, c < dpad; zero outside (row >= rows_valid, c >= d)
__device__ __forceinline__ void load_tile(float* __restrict__ dst, int DS, int dpad, const float* __restrict__ src,
                                          long ld, int row0, int nrows, int rows_valid, int d) {
  const int lane = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < nrows; r += ATT_THREADS / 32) {
    const int row = row0 + r;
    const bool ok = row < rows_valid;
    const float* s = src + (long)row * ld;
    for (int c = lane; c < dpad; c += 32) dst[r * DS + c] = (ok && c < d) ? s[c] : 0.f;
  }
}

// index of the first key of batch element b whose mask is non-zero (T if none); all threads get the value
__device__ __forceinline__ int block_first_key(const float* __restrict__ kmask_b, int T, int* s_first) {
  if (threadIdx.x == 0) *s_first = T;
  __syncthreads();
  int mine = T;
  for (int j = threadIdx.x; j < T; j += ATT_THREADS)
    if (kmask_b[j] != 0.f) { mine = j; break; }
  if (mine < T) atomicMin(s_first, mine);
  __syncthreads();
  return *s_first;
}

// first_key as above, and qstart = index of the first row of batch element b whose item id is non-zero (0 when no ids
// are given): rows before qstart are padding whose block output is multiplied by 0 (`seq *= mask`, sasrec.py:83), so
// neither their attention row nor their gradient is ever observable and the kernels skip them.
__device__ __forceinline__ void block_first2(const float* __restrict__ kmask_b, const int* __restrict__ ids_b, int T,
                                             int* s2, int& first_key, int& qstart) {
  if (threadIdx.x == 0) { s2[0] = T; s2[1] = ids_b ? T : 0; }
  __syncthreads();
  int mk = T, mq = T;
  for (int j = threadIdx.x; j < T; j += ATT_THREADS) {
    if (mk == T && kmask_b[j] != 0.f) mk = j;
    if (ids_b && mq == T && ids_b[j] != 0) mq = j;
  }
  if (mk < T) atomicMin(&s2[0], mk);
  if (ids_b && mq < T) atomicMin(&s2[1], mq);
  __syncthreads();
  first_key = s2[0];
  qstart = s2[1];
}

struct AttnFwdArgs {
  const float *Q, *K, *V;  // [B*T, ld*]
  long ldq, ldk, ldv;
  const float* resid;      // queries = LN(x) [B*T, H]
  const float* kmask;      // [B*T]
  const float* qmask;      // [B*T]
  float* out;              // [B*T, H]
  const int* skip_ids;     // [B*T] item ids or null: leading rows with id 0 are skipped (out = queries)
  float* attn;             // [h*B, T, T] or null
  float* row_max;          // [B, h, T] or null
  float* row_linv;         // [B, h, T] or null
  float rate;
  unsigned long long seed;
  const unsigned long long* step;
  int site;
};

template <int RI, int RJ, int NB, int DPAD, int DS>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(AttnFwdArgs a, AttnDims dm) {
  constexpr int TQ = 16 * RI, TC = 16 * RJ;
  CAST_DYN_SMEM(float, sm);
  __shared__ int s_first[2];
  float* Qs = sm;
  float* KVs = Qs + TQ * DS;
  float* S = KVs + TC * DS;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, lane = t & 31, warp = t >> 5;
  const int q0 = blockIdx.x * TQ, hh = blockIdx.y, b = blockIdx.z;
  const int T = dm.T, d = dm.d, PS = dm.PS;
  constexpr int dpad = DPAD;
  const long rowbase = (long)b * T;
  int first_key, qstart;
  block_first2(a.kmask + rowbase, (a.skip_ids && !a.attn) ? a.skip_ids + rowbase : nullptr, T, s_first, first_key,
               qstart);
  if (q0 + TQ <= qstart) {  // every row of this tile is padding: output = residual only (block-uniform exit)
    for (int idx = t; idx < TQ * d; idx += ATT_THREADS) {
      const int i = q0 + idx / d, c = idx % d;
      if (i < T) {
        const long off = (rowbase + i) * dm.H + hh * d + c;
        a.out[off] = a.resid[off];
      }
    }
    for (int r = t; r < TQ; r += ATT_THREADS) {
      const int i = q0 + r;
      if (i < T) {
        const long si = ((long)b * dm.h + hh) * T + i;
        if (a.row_max) a.row_max[si] = 0.f;
        if (a.row_linv) a.row_linv[si] = 0.f;
      }
    }
    return;
  }
  const int qlo = q0 > qstart ? q0 : qstart;  // first computed row
  const bool uni = qlo < first_key;           // tile holds fully-masked rows: they are uniform over all T keys
  const int kend = uni ? T : (q0 + TQ < T ? q0 + TQ : T);
  const int kbeg = uni ? 0 : (first_key / TC) * TC;  // keys before first_key are masked for every computed row

  load_tile(Qs, DS, dpad, a.Q + rowbase * a.ldq + hh * d, a.ldq, q0, TQ, T, d);
  // ---- S = Q K^T (raw), chunk by chunk
  for (int j0 = kbeg; j0 < kend; j0 += TC) {
    __syncthreads();
    load_tile(KVs, DS, dpad, a.K + rowbase * a.ldk + hh * d, a.ldk, j0, TC, T, d);
    __syncthreads();
    float acc[RI][RJ];
    dot_tile<RI, RJ, DPAD, DS>(Qs, KVs, acc, ty, tx);
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int j = j0 + tx + 16 * jj;
        if (j < dm.Tp4) S[(ty * RI + ii) * PS + j] = acc[ii][jj];
      }
  }
  __syncthreads();
  // ---- masks, softmax, query mask, dropout: one warp per row
  const Drop dr = make_drop(a.rate, a.seed, a.step, a.site);
  for (int r = warp; r < TQ; r += ATT_THREADS / 32) {
    const int i = q0 + r;
    float* Sr = S + r * PS;
    if (i >= T || i < qstart) {
      for (int j = kbeg + lane; j < dm.Tp4; j += 32) Sr[j] = 0.f;
      if (i < T && lane == 0) {
        const long si = ((long)b * dm.h + hh) * T + i;
        if (a.row_max) a.row_max[si] = 0.f;
        if (a.row_linv) a.row_linv[si] = 0.f;
      }
      continue;
    }
    float mx = -INFINITY;
    for (int j = kbeg + lane; j < kend; j += 32) {
      const bool keep = (a.kmask[rowbase + j] != 0.f) && (j <= i);
      const float s = keep ? Sr[j] * dm.inv_sqrt_d : CAST_NEG_FILL;
      Sr[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = kbeg + lane; j < kend; j += 32) {
      const float e = __expf(Sr[j] - mx);
      Sr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float linv = 1.0f / sum;
    const float qm = a.qmask[rowbase + i] * linv;
    const unsigned long long ibase = ((unsigned long long)((long)hh * dm.B + b) * T + i) * T;
    float* arow = a.attn ? a.attn + ibase : nullptr;
    if (arow)  // keys before the first kept one are skipped by the tile loop: their weight is exactly 0
      for (int j = lane; j < kbeg; j += 32) arow[j] = 0.f;
    for (int j = kbeg + lane; j < dm.Tp4; j += 32) {
      float p = 0.f;
      if (j < kend) p = Sr[j] * qm * drop_mul(dr, ibase + j);
      Sr[j] = p;
      if (arow && j < T) arow[j] = p;
    }
    if (lane == 0) {
      const long si = ((long)b * dm.h + hh) * T + i;
      if (a.row_max) a.row_max[si] = mx;
      if (a.row_linv) a.row_linv[si] = linv;
    }
  }
  // ---- O = P V
  float o[NB][RI][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) o[nb][ii][cc] = 0.f;
  for (int j0 = kbeg; j0 < kend; j0 += TC) {
    __syncthreads();
    load_tile(KVs, DS, dpad, a.V + rowbase * a.ldv + hh * d, a.ldv, j0, TC, T, d);
    __syncthreads();
    int nj = dm.Tp4 - j0;
    if (nj > TC) nj = TC;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int col = nb * 64 + tx * 4;
      if (col < dpad) pv_tile<RI, DS>(S + j0, PS, KVs, DS, nj, col, o[nb], ty);
    }
  }
  // ---- heads merged + residual (outputs += queries)
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) {
      const int i = q0 + ty * RI + ii;
      if (i >= T) continue;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = nb * 64 + tx * 4 + cc;
        if (c < d) {
          const long off = (rowbase + i) * dm.H + hh * d + c;
          a.out[off] = o[nb][ii][cc] + a.resid[off];
        }
      }
    }
}

struct AttnBwdArgs {
  const float *Q, *K, *V;
  long ldq, ldk, ldv;
  const float* dO;  // gradient w.r.t. the attention output [B*T, H] (the residual branch is handled by the caller)
  const float* kmask;
  const float* qmask;
  const float* row_max;
  const float* row_linv;
  const int* skip_ids;  // as in the forward pass (must be the same pointer semantics)
  float* rowD;  // [B, h, T]  written by dq kernel, read by dkv kernel
  float *dQ, *dK, *dV;
  long lddq, lddk, lddv;
  float rate;
  unsigned long long seed;
  const unsigned long long* step;
  int site;
};

template <int RI, int RJ, int NB, int DPAD, int DS>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(AttnBwdArgs a, AttnDims dm) {
  constexpr int TQ = 16 * RI, TC = 16 * RJ;
  CAST_DYN_SMEM(float, sm);
  __shared__ int s_first[2];
  float* Qs = sm;
  float* dOs = Qs + TQ * DS;
  float* KVs = dOs + TQ * DS;
  float* P = KVs + TC * DS;
  float* dP = P + TQ * dm.PS;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, lane = t & 31, warp = t >> 5;
  const int q0 = blockIdx.x * TQ, hh = blockIdx.y, b = blockIdx.z;
  const int T = dm.T, d = dm.d, PS = dm.PS;
  constexpr int dpad = DPAD;
  const long rowbase = (long)b * T;
  int first_key, qstart;
  block_first2(a.kmask + rowbase, a.skip_ids ? a.skip_ids + rowbase : nullptr, T, s_first, first_key, qstart);
  if (q0 + TQ <= qstart) {  // padding-only tile: zero gradient
    for (int idx = t; idx < TQ * d; idx += ATT_THREADS) {
      const int i = q0 + idx / d, c = idx % d;
      if (i < T) a.dQ[(rowbase + i) * a.lddq + hh * d + c] = 0.f;
    }
    for (int r = t; r < TQ; r += ATT_THREADS)
      if (q0 + r < T) a.rowD[((long)b * dm.h + hh) * T + q0 + r] = 0.f;
    return;
  }
  const int qlo = q0 > qstart ? q0 : qstart;
  const bool uni = qlo < first_key;
  const int kend = uni ? T : (q0 + TQ < T ? q0 + TQ : T);
  const int kbeg = uni ? 0 : (first_key / TC) * TC;

  load_tile(Qs, DS, dpad, a.Q + rowbase * a.ldq + hh * d, a.ldq, q0, TQ, T, d);
  load_tile(dOs, DS, dpad, a.dO + rowbase * dm.H + hh * d, dm.H, q0, TQ, T, d);
  for (int j0 = kbeg; j0 < kend; j0 += TC) {
    __syncthreads();
    load_tile(KVs, DS, dpad, a.K + rowbase * a.ldk + hh * d, a.ldk, j0, TC, T, d);
    __syncthreads();
    float acc[RI][RJ];
    dot_tile<RI, RJ, DPAD, DS>(Qs, KVs, acc, ty, tx);
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int j = j0 + tx + 16 * jj;
        if (j < dm.Tp4) P[(ty * RI + ii) * PS + j] = acc[ii][jj];
      }
    __syncthreads();
    load_tile(KVs, DS, dpad, a.V + rowbase * a.ldv + hh * d, a.ldv, j0, TC, T, d);
    __syncthreads();
    dot_tile<RI, RJ, DPAD, DS>(dOs, KVs, acc, ty, tx);
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int j = j0 + tx + 16 * jj;
        if (j < dm.Tp4) dP[(ty * RI + ii) * PS + j] = acc[ii][jj];
      }
  }
  __syncthreads();
  const Drop dr = make_drop(a.rate, a.seed, a.step, a.site);
  for (int r = warp; r < TQ; r += ATT_THREADS / 32) {
    const int i = q0 + r;
    float* Pr = P + r * PS;
    float* dPr = dP + r * PS;
    if (i >= T || i < qstart) {
      for (int j = kbeg + lane; j < dm.Tp4; j += 32) dPr[j] = 0.f;
      if (i < T && lane == 0) a.rowD[((long)b * dm.h + hh) * T + i] = 0.f;
      continue;
    }
    const long si = ((long)b * dm.h + hh) * T + i;
    const float mx = a.row_max[si], linv = a.row_linv[si], qm = a.qmask[rowbase + i];
    const unsigned long long ibase = ((unsigned long long)((long)hh * dm.B + b) * T + i) * T;
    float D = 0.f;
    for (int j = kbeg + lane; j < kend; j += 32) {
      const bool keep = (a.kmask[rowbase + j] != 0.f) && (j <= i);
      const float s = keep ? Pr[j] * dm.inv_sqrt_d : CAST_NEG_FILL;
      const float p = __expf(s - mx) * linv;
      const float dp = dPr[j] * qm * drop_mul(dr, ibase + j);
      Pr[j] = p;
      dPr[j] = dp;
      D += p * dp;
    }
    D = warp_sum(D);
    for (int j = kbeg + lane; j < dm.Tp4; j += 32) {
      float ds = 0.f;
      if (j < kend) {
        const bool keep = (a.kmask[rowbase + j] != 0.f) && (j <= i);
        if (keep) ds = Pr[j] * (dPr[j] - D) * dm.inv_sqrt_d;
      }
      dPr[j] = ds;
    }
    if (lane == 0) a.rowD[si] = D;
  }
  float o[NB][RI][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) o[nb][ii][cc] = 0.f;
  for (int j0 = kbeg; j0 < kend; j0 += TC) {
    __syncthreads();
    load_tile(KVs, DS, dpad, a.K + rowbase * a.ldk + hh * d, a.ldk, j0, TC, T, d);
    __syncthreads();
    int nj = dm.Tp4 - j0;
    if (nj > TC) nj = TC;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int col = nb * 64 + tx * 4;
      if (col < dpad) pv_tile<RI, DS>(dP + j0, PS, KVs, DS, nj, col, o[nb], ty);
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) {
      const int i = q0 + ty * RI + ii;
      if (i >= T) continue;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = nb * 64 + tx * 4 + cc;
        if (c < d) a.dQ[(rowbase + i) * a.lddq + hh * d + c] = o[nb][ii][cc];
      }
    }
}

template <int RI, int RJ, int NB, int DPAD, int DS>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(AttnBwdArgs a, AttnDims dm) {
  constexpr int TKT = 16 * RI, TC = 16 * RJ, TCP = TC + 4;
  CAST_DYN_SMEM(float, sm);
  __shared__ int s_first[2];
  float* Ks = sm;
  float* Vs = Ks + TKT * DS;
  float* Qc = Vs + TKT * DS;
  float* dOc = Qc + TC * DS;
  float* St = dOc + TC * DS;   // [TKT][TCP]  -> P~^T (dropped, query-masked probabilities)
  float* dSt = St + TKT * TCP;    // [TKT][TCP]  -> dS^T
  float* stat = dSt + TKT * TCP;  // [4][TC]: row max, 1/rowsum, D, query mask
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int k0 = blockIdx.x * TKT, hh = blockIdx.y, b = blockIdx.z;
  const int T = dm.T, d = dm.d;
  constexpr int dpad = DPAD;
  const long rowbase = (long)b * T;
  int first_key, qstart;
  block_first2(a.kmask + rowbase, a.skip_ids ? a.skip_ids + rowbase : nullptr, T, s_first, first_key, qstart);
  const bool has_uniform = qstart < first_key;  // some computed (non-padding) row is fully masked
  if (k0 + TKT <= first_key && !has_uniform) {
    // every key of this tile is masked and no row is uniform: P == 0 on the whole tile => zero gradients
    for (int idx = t; idx < TKT * d; idx += ATT_THREADS) {
      const int j = k0 + idx / d, c = idx % d;
      if (j < T) {
        a.dK[(rowbase + j) * a.lddk + hh * d + c] = 0.f;
        a.dV[(rowbase + j) * a.lddv + hh * d + c] = 0.f;
      }
    }
    return;
  }
  const Drop dr = make_drop(a.rate, a.seed, a.step, a.site);

  load_tile(Ks, DS, dpad, a.K + rowbase * a.ldk + hh * d, a.ldk, k0, TKT, T, d);
  load_tile(Vs, DS, dpad, a.V + rowbase * a.ldv + hh * d, a.ldv, k0, TKT, T, d);
  float gk[NB][RI][4], gv[NB][RI][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) gk[nb][ii][cc] = gv[nb][ii][cc] = 0.f;

  for (int c0 = 0; c0 < T; c0 += TC) {
    // queries of this chunk matter if some are causally >= our keys, or are uniform (fully-masked) rows
    if (c0 + TC <= qstart) continue;                                      // padding rows only
    if (c0 + TC <= k0 && !(has_uniform && c0 < first_key)) continue;      // causally before our keys, not uniform
    __syncthreads();
    load_tile(Qc, DS, dpad, a.Q + rowbase * a.ldq + hh * d, a.ldq, c0, TC, T, d);
    load_tile(dOc, DS, dpad, a.dO + rowbase * dm.H + hh * d, dm.H, c0, TC, T, d);
    for (int ic = t; ic < TC; ic += ATT_THREADS) {
      const int i = c0 + ic;
      const bool ok = i < T;
      const long si = ((long)b * dm.h + hh) * T + i;
      stat[0 * TC + ic] = ok ? a.row_max[si] : 0.f;
      stat[1 * TC + ic] = ok ? a.row_linv[si] : 0.f;
      stat[2 * TC + ic] = ok ? a.rowD[si] : 0.f;
      stat[3 * TC + ic] = ok ? a.qmask[rowbase + i] : 0.f;
    }
    __syncthreads();
    {
      float acc[RI][RJ];
      dot_tile<RI, RJ, DPAD, DS>(Ks, Qc, acc, ty, tx);
#pragma unroll
      for (int ii = 0; ii < RI; ++ii)
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) St[(ty * RI + ii) * TCP + tx + 16 * jj] = acc[ii][jj];
      dot_tile<RI, RJ, DPAD, DS>(Vs, dOc, acc, ty, tx);
#pragma unroll
      for (int ii = 0; ii < RI; ++ii)
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) dSt[(ty * RI + ii) * TCP + tx + 16 * jj] = acc[ii][jj];
    }
    __syncthreads();
    for (int idx = t; idx < TKT * TC; idx += ATT_THREADS) {
      const int jr = idx / TC, ic = idx - jr * TC;
      const int j = k0 + jr, i = c0 + ic;
      float pd = 0.f, ds = 0.f;
      if (i < T && j < T && i >= qstart) {
        const bool keep = (a.kmask[rowbase + j] != 0.f) && (j <= i);
        const float s = keep ? St[jr * TCP + ic] * dm.inv_sqrt_d : CAST_NEG_FILL;
        const float p = __expf(s - stat[ic]) * stat[TC + ic];
        const unsigned long long eidx = (((unsigned long long)((long)hh * dm.B + b) * T + i) * T) + j;
        const float mul = stat[3 * TC + ic] * drop_mul(dr, eidx);
        pd = p * mul;
        const float dp = dSt[jr * TCP + ic] * mul;
        if (keep) ds = p * (dp - stat[2 * TC + ic]) * dm.inv_sqrt_d;
      }
      St[jr * TCP + ic] = pd;
      dSt[jr * TCP + ic] = ds;
    }
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int col = nb * 64 + tx * 4;
      if (col < dpad) {
        pv_tile<RI, DS>(St, TCP, dOc, DS, TC, col, gv[nb], ty);
        pv_tile<RI, DS>(dSt, TCP, Qc, DS, TC, col, gk[nb], ty);
      }
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) {
      const int j = k0 + ty * RI + ii;
      if (j >= T) continue;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = nb * 64 + tx * 4 + cc;
        if (c < d) {
          a.dK[(rowbase + j) * a.lddk + hh * d + c] = gk[nb][ii][cc];
          a.dV[(rowbase + j) * a.lddv + hh * d + c] = gv[nb][ii][cc];
        }
      }
    }
}

static inline size_t fwd_smem(const AttnDims& dm, int RI, int RJ) {
  return sizeof(float) * ((size_t)(16 * RI + 16 * RJ) * dm.DS + (size_t)16 * RI * dm.PS);
}
static inline size_t dq_smem(const AttnDims& dm, int RI, int RJ) {
  return sizeof(float) * ((size_t)(32 * RI + 16 * RJ) * dm.DS + (size_t)32 * RI * dm.PS);
}
static inline size_t dkv_smem(const AttnDims& dm, int RI, int RJ) {
  return sizeof(float) * ((size_t)(32 * RI + 32 * RJ) * dm.DS + (size_t)32 * RI * (16 * RJ + 4) + 4 * 16 * RJ);
}
constexpr size_t SMEM_MAX = 227 * 1024;

template <class K, class A>
static int launch_att(K kfn, size_t smem, int gridx, int h, int B, const A& args, const AttnDims& dm,
                      cudaStream_t stream, size_t* configured) {
  if (smem > SMEM_MAX) return set_error(CAST_ERR_UNSUPPORTED, "attention: tile does not fit shared memory");
  if (smem > *configured) {  // opt-in once per kernel instantiation (never during a graph replay)
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    *configured = smem;
  }
  CAST_LAUNCH(kfn, dim3((unsigned)gridx, (unsigned)h, (unsigned)B), dim3(ATT_THREADS), smem, stream, args, dm);
  return CAST_OK;
}

// One instantiation per padded head width (DPAD = d rounded up to 4; DS = its bank-friendly row stride).
// which: 0 = forward, 1 = backward dQ, 2 = backward dK/dV.
template <int DPAD>
int dispatch_att(int which, const AttnFwdArgs* fa, const AttnBwdArgs* ba, const AttnDims& dm,
                        cudaStream_t stream) {
  constexpr int DS = ((DPAD / 4) % 2 == 0) ? DPAD + 4 : DPAD;
  constexpr int NB = (DPAD + 63) / 64;
  static size_t cfg[3] = {48 * 1024, 48 * 1024, 48 * 1024};
  const int T = dm.T, h = dm.h, B = dm.B;
  if (NB == 1) {
    // small heads (measured at T=200, d=50 on B200, gpurun_out/tune_*.log): 32-query tiles for the forward and the
    // dQ kernels (46 / 79 KB of shared memory => 4 / 2 CTAs per SM instead of 2 / 1, and half the causal waste on
    // the diagonal tile), 64-key tiles for dK/dV (its 32-row variant re-reads Q/dO twice as often and was slower)
    if (which == 0)
      return launch_att(attn_fwd_kernel<2, 4, 1, DPAD, DS>, fwd_smem(dm, 2, 4), (int)cdiv(T, 32), h, B, *fa, dm, stream, &cfg[0]);
    if (which == 1)
      return launch_att(attn_bwd_dq_kernel<2, 4, 1, DPAD, DS>, dq_smem(dm, 2, 4), (int)cdiv(T, 32), h, B, *ba, dm, stream, &cfg[1]);
    return launch_att(attn_bwd_dkv_kernel<4, 4, 1, DPAD, DS>, dkv_smem(dm, 4, 4), (int)cdiv(T, 64), h, B, *ba, dm, stream, &cfg[2]);
  } else if (NB == 2) {
    if (which == 0)
      return launch_att(attn_fwd_kernel<2, 4, NB, DPAD, DS>, fwd_smem(dm, 2, 4), (int)cdiv(T, 32), h, B, *fa, dm, stream, &cfg[0]);
    if (which == 1)
      return launch_att(attn_bwd_dq_kernel<2, 4, NB, DPAD, DS>, dq_smem(dm, 2, 4), (int)cdiv(T, 32), h, B, *ba, dm, stream, &cfg[1]);
    return launch_att(attn_bwd_dkv_kernel<2, 4, NB, DPAD, DS>, dkv_smem(dm, 2, 4), (int)cdiv(T, 32), h, B, *ba, dm, stream, &cfg[2]);
  } else {
    if (which == 0)
      return launch_att(attn_fwd_kernel<2, 2, NB, DPAD, DS>, fwd_smem(dm, 2, 2), (int)cdiv(T, 32), h, B, *fa, dm, stream, &cfg[0]);
    if (which == 1)
      return launch_att(attn_bwd_dq_kernel<2, 2, NB, DPAD, DS>, dq_smem(dm, 2, 2), (int)cdiv(T, 32), h, B, *ba, dm, stream, &cfg[1]);
    return launch_att(attn_bwd_dkv_kernel<2, 2, NB, DPAD, DS>, dkv_smem(dm, 2, 2), (int)cdiv(T, 32), h, B, *ba, dm, stream, &cfg[2]);
  }
}


}  // namespace cast
