// Row kernels of one transformer block on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// tensor memory), for hidden_units <= RU_MAX_H — the GEMM-shaped [B*T, H] x [H, H] products of
//   ln_qkv_fwd : normalize (modules.py:74-78) -> Q = LN(x)Wq+bq, K = xWk+bk, V = xWv+bv (modules.py:203-205), key /
//                query zero-sum flags (:222, :248)
//   ln_ffn_fwd : normalize -> relu(.W1+b1) -> dropout -> .W2+b2 -> dropout -> + LN(y) -> *mask (modules.py:298-313,
//                sasrec.py:81-83)
// with the same arithmetic as the mma.sync kernels of fused_mma.cuh (3xTF32 operand split: x = hi + lo,
// A*B ~= lo*hi + hi*lo + hi*hi, fp32 accumulation) and the same saved tensors for backward.
//
// Structure (persistent CTAs, one per SM; 16 "row" warps + 1 MMA warp; every CTA owns a contiguous, balanced range of
// rows walked in 128-row tiles; row thread t owns row t & 127 = TMEM lane and the 16 columns of quarter t >> 7):
//   * the row tile [128, H] is ONE contiguous block of global memory: it is brought into a dense shared-memory tile with
//     16-byte cp.async (next tile prefetched while the current one is processed) and every output tile leaves through a
//     dense tile + coalesced 16-byte stores, so global traffic is fully coalesced although threads own rows;
//   * LayerNorm runs thread-per-(row, half) from registers (two partial sums per row meet in shared memory);
//   * the operand is split into tf32 hi / lo once per element and stored in the K-major no-swizzle UMMA layout of
//     umma.cuh (16-byte stores, conflict free); weights come pre-split as images made once per step by
//     cast_rowk_presplit and are bulk-copied (cp.async.bulk + mbarrier transaction bytes) once per CTA;
//   * a dedicated warp issues 3 x ceil(H/8) tcgen05.mma per product (M = 128, N = 64 or 128) as soon as the row warps
//     have arrived on the operand's mbarrier, and commits to a second mbarrier; the row warps never wait for the issue
//     loop, only for the product they need next.  In ln_qkv_fwd the K|V product (operand = raw x) is issued first and
//     runs under the LayerNorm; the Q product runs under the K and V epilogues;
//   * the epilogue reads the accumulator with tcgen05.ld (thread <-> row) and applies bias / ReLU / dropout / residual /
//     padding mask in registers.
#include "cast_rt.cuh"
#include "mma_tf32.cuh"
#include "umma.cuh"

namespace cast {

constexpr int RU_ROWS = 128;
constexpr int RU_ROW_THREADS = 512;          // 16 warps: 4 lane quarters x 4 column quarters
constexpr int RU_THREADS = RU_ROW_THREADS + 32;  // + the MMA warp
constexpr int RU_PA = RU_ROWS * 16 + 16;  // slab pitch of a 128-row operand tile (bytes)
constexpr int RU_NIMG = 9;                // weight images per block, see ru_img_offset
constexpr size_t RU_SMEM_MAX = 232448;    // 227 KB opt-in dynamic shared memory per CTA

struct RuShape {
  int H, KS, SL, NP;  // KS = ceil(H/8) k-steps, SL = 2*KS slabs of 4 floats, NP = 16*ceil(H/16) accumulator columns
  int img1, img2;     // bytes of a weight image with NP rows / 2*NP rows (hi then lo)
  int dense;          // bytes of a dense [128, H] tile, rounded up to 16
};

__host__ __device__ inline RuShape ru_shape(int H) {
  RuShape s;
  s.H = H;
  s.KS = (H + 7) / 8;
  s.SL = 2 * s.KS;
  s.NP = (H + 15) / 16 * 16;
  s.img1 = 2 * s.SL * (s.NP * 16 + 16);
  s.img2 = 2 * s.SL * (2 * s.NP * 16 + 16);
  s.dense = (RU_ROWS * H * 4 + 15) & ~15;
  return s;
}

// images of one block, in this order: WqT | [WkT;WvT] (2*NP rows) | W1T | W2T | Wq | Wk | Wv | W1 | W2
//   "T" images are the B operand of C = A W      : B[n][k] = W[k][n]
//   plain images are the B operand of dX = dY W^T : B[n][k] = W[n][k]
__host__ __device__ inline size_t ru_img_offset(const RuShape& s, int which) {
  size_t off = 0;
  for (int i = 0; i < which; ++i) off += (i == 1) ? s.img2 : s.img1;
  return off;
}
__host__ __device__ inline size_t ru_img_block_bytes(const RuShape& s) { return ru_img_offset(s, RU_NIMG); }

// ---------------------------------------------------------------------------------------------------------------------
struct RuPresplitArgs {
  const float* w[16][5];  // per block: Wq, Wk, Wv, W1, W2   ([H, H] row-major, [in, out])
};

// one thread per (image, row n, column k); grid = (ceil(2*NP*KP / 256), RU_NIMG, nblocks)
__global__ void __launch_bounds__(256) ru_presplit_kernel(RuPresplitArgs a, int H, unsigned char* __restrict__ images) {
  const RuShape s = ru_shape(H);
  const int blk = blockIdx.z, which = blockIdx.y;
  unsigned char* img = images + (size_t)blk * ru_img_block_bytes(s) + ru_img_offset(s, which);
  const int rows = (which == 1) ? 2 * s.NP : s.NP;
  const int pitch = rows * 16 + 16;
  const int half = s.SL * pitch;
  const int KP = 4 * s.SL;
  const bool transposed = which < 4;
  const int idx = blockIdx.x * (int)blockDim.x + threadIdx.x;
  if (idx >= rows * KP) return;
  int n, k;  // the fast index follows the unit stride of the source so that the global reads coalesce
  if (transposed) { k = idx / rows; n = idx - k * rows; } else { n = idx / KP; k = idx - n * KP; }
  const float* W;
  int nn = n;
  switch (which) {
    case 0: case 4: W = a.w[blk][0]; break;
    case 1: W = a.w[blk][n < s.NP ? 1 : 2]; nn = n < s.NP ? n : n - s.NP; break;
    case 2: case 7: W = a.w[blk][3]; break;
    case 3: case 8: W = a.w[blk][4]; break;
    case 5: W = a.w[blk][1]; break;
    default: W = a.w[blk][2]; break;
  }
  float v = 0.f;
  if (nn < H && k < H) v = transposed ? W[k * H + nn] : W[nn * H + k];
  unsigned hi, lo;
  tf32_split(v, hi, lo);
  const int off = (k >> 2) * pitch + n * 16 + (k & 3) * 4;
  *reinterpret_cast<unsigned*>(img + off) = hi;
  *reinterpret_cast<unsigned*>(img + half + off) = lo;
}

// ---------------------------------------------------------------------------------------------------------------------
// device helpers.  Row thread t: row r = t & 127, column quarter qd = t >> 7, columns c0 + i (i < 16, c0 = 16 qd),
// slabs 4 qd + i (i < 4).

#ifndef CAST_EMU
__device__ __forceinline__ void ru_cp_async16(float* dst, const float* src, int bytes) {  // bytes in [0,16]: rest zero
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
#else
inline void ru_cp_async16(float* dst, const float* src, int bytes) {
  for (int i = 0; i < 4; ++i) dst[i] = (4 * i < bytes) ? src[i] : 0.f;
}
#endif

__device__ __forceinline__ void ru_row_sync() { umma::named_sync(1, RU_ROW_THREADS); }

// dense tile <- nvalid floats starting at src (a contiguous row block of a [N, H] tensor), zero up to ntotal floats
__device__ __forceinline__ void ru_tile_load_async(float* __restrict__ tile, const float* __restrict__ src, long nvalid,
                                                   int ntotal) {
  const int t = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = 4 * t; i < ntotal; i += 4 * RU_ROW_THREADS) {
      long left = (nvalid - i) * 4;
      const int bytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
      ru_cp_async16(tile + i, bytes > 0 ? src + i : src, bytes);
    }
  } else {  // unaligned base pointer: plain loads (never the case for the engine's buffers)
    for (int i = t; i < ntotal; i += RU_ROW_THREADS) tile[i] = i < nvalid ? src[i] : 0.f;
  }
  cp_async_commit();
}

// nvalid floats of a dense tile -> contiguous global block
__device__ __forceinline__ void ru_tile_store(float* __restrict__ dst, const float* __restrict__ tile, long nvalid) {
  const int t = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int n4 = (int)(nvalid >> 2);
    for (int i = t; i < n4; i += RU_ROW_THREADS)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(tile)[i];
    for (int i = 4 * n4 + t; i < nvalid; i += RU_ROW_THREADS) dst[i] = tile[i];
  } else {
    for (int i = t; i < nvalid; i += RU_ROW_THREADS) dst[i] = tile[i];
  }
}

// this thread's 16 columns of its row: x[i] = tile[r][c0 + i] (0 for columns >= H)
__device__ __forceinline__ void ru_row_load(const float* __restrict__ tile, int r, int c0, int H, float (&x)[16]) {
  const float* p = tile + r * H + c0;
  if ((H & 1) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (c0 + i < H) {
        const float2 v = *reinterpret_cast<const float2*>(p + i);
        x[i] = v.x;
        x[i + 1] = v.y;
      } else {
        x[i] = x[i + 1] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (c0 + i < H) ? p[i] : 0.f;
  }
}

__device__ __forceinline__ void ru_row_store(float* __restrict__ tile, int r, int c0, int H, const float (&x)[16]) {
  float* p = tile + r * H + c0;
  if ((H & 1) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 2)
      if (c0 + i < H) *reinterpret_cast<float2*>(p + i) = make_float2(x[i], x[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < H) p[i] = x[i];
  }
}

// split this thread's columns into tf32 hi / lo and store them as its slabs of the operand tile (x is 0 beyond H)
__device__ __forceinline__ void ru_stage(unsigned char* __restrict__ hi, unsigned char* __restrict__ lo, int r, int qd,
                                         int SL, const float (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = 4 * qd + i;
    if (s < SL) {
      uint4 h, l;
      tf32_split(x[4 * i + 0], h.x, l.x);
      tf32_split(x[4 * i + 1], h.y, l.y);
      tf32_split(x[4 * i + 2], h.z, l.z);
      tf32_split(x[4 * i + 3], h.w, l.w);
      *reinterpret_cast<uint4*>(hi + s * RU_PA + r * 16) = h;
      *reinterpret_cast<uint4*>(lo + s * RU_PA + r * 16) = l;
    }
  }
}

// the row warps' "operand tile is staged" signal: every writer makes its generic-proxy stores visible to the async
// proxy, the warp converges, one lane arrives (barrier count = 16 warps)
__device__ __forceinline__ void ru_operand_ready(uint64_t* bar) {
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}

// D[128 x N] (tensor memory, columns from tmem_d) = A[128 x 8KS] * B[N x 8KS]^T with the 3xTF32 split; ONE thread.
// a_*, b_* are shared-memory addresses of the hi / lo halves; the small terms are accumulated first.
__device__ __forceinline__ void ru_issue(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                         uint32_t b_pitch, int KS, uint32_t idesc, uint32_t accumulate) {
  uint64_t dah = umma::smem_desc(a_hi, RU_PA, 128), dal = umma::smem_desc(a_lo, RU_PA, 128);
  uint64_t dbh = umma::smem_desc(b_hi, b_pitch, 128), dbl = umma::smem_desc(b_lo, b_pitch, 128);
  for (int ks = 0; ks < KS; ++ks) {
    umma::mma_tf32(tmem_d, dal, dbh, idesc, (ks > 0) ? 1u : accumulate);
    umma::mma_tf32(tmem_d, dah, dbl, idesc, 1u);
    umma::mma_tf32(tmem_d, dah, dbh, idesc, 1u);
    dah = umma::desc_advance(dah, 2 * RU_PA);
    dal = umma::desc_advance(dal, 2 * RU_PA);
    dbh = umma::desc_advance(dbh, 2 * b_pitch);
    dbl = umma::desc_advance(dbl, 2 * b_pitch);
  }
}

// optional phase trace (tuning): row thread 0 of every CTA stamps clock64() at phase boundaries, [cta][tile slot < 4][16]
__device__ long long* g_rowk_trace = nullptr;
#ifndef CAST_EMU
#define RU_STAMP(slot, k)                                                                      \
  if (trace && threadIdx.x == 0 && (slot) < 4) trace[((long)blockIdx.x * 4 + (slot)) * 16 + (k)] = clock64()
#else
#define RU_STAMP(slot, k)
#endif

// a watchdog instead of a hang: a barrier phase that never completes is reported through cast_rowk_status
__device__ int g_rowk_timeout = 0;
__device__ __forceinline__ void ru_wait(uint64_t* bar, uint32_t parity) {
  if (!umma::mbar_wait(bar, parity, 1u << 22)) g_rowk_timeout = 1;
}

// balanced partition of the N rows over the persistent grid: every CTA owns rpc consecutive rows (a multiple of 4, so
// that tile base addresses stay 16-byte aligned), walked in 128-row tiles; only the last tile of a CTA is partial
struct RuRange {
  long lo, hi;  // rows [lo, hi)
  int ntiles;
};
__device__ __forceinline__ RuRange ru_range(long N, long rpc) {
  RuRange g;
  g.lo = (long)blockIdx.x * rpc;
  g.hi = g.lo + rpc < N ? g.lo + rpc : N;
  g.ntiles = g.hi > g.lo ? (int)((g.hi - g.lo + RU_ROWS - 1) / RU_ROWS) : 0;
  return g;
}

// LayerNorm of the row held as four column quarters by threads r, r+128, r+256, r+384 (modules.py:74-78: biased
// variance, eps inside the sqrt).  y = gamma * ((x - mean) * rstd) + beta for columns < H, 0 beyond.
// red: [3][128][4] floats.  Two row barriers inside; the partial sums of y are left in red[2] for the caller to read
// after ITS next row barrier.  `between` runs right after the first barrier (every thread has finished reading its
// input row by then).
struct RuLn {
  float mean, rstd, xsum;
};
template <class F>
__device__ __forceinline__ RuLn ru_layernorm(const float (&x)[16], float (&y)[16], const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, int r, int qd, int c0, int H,
                                             float* __restrict__ red, F&& between) {
  float sp = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) sp += x[i];  // columns >= H hold zeros
  red[r * 4 + qd] = sp;
  ru_row_sync();
  between();
  RuLn o;
  {
    const float4 p = *reinterpret_cast<const float4*>(red + r * 4);
    o.xsum = (p.x + p.y) + (p.z + p.w);
  }
  o.mean = o.xsum / (float)H;
  float qp = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = (c0 + i < H) ? x[i] - o.mean : 0.f;
    y[i] = d;
    qp += d * d;
  }
  red[512 + r * 4 + qd] = qp;
  ru_row_sync();
  {
    const float4 p = *reinterpret_cast<const float4*>(red + 512 + r * 4);
    o.rstd = 1.0f / sqrtf(((p.x + p.y) + (p.z + p.w)) / (float)H + eps);
  }
  float yp = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + c0 + i);
    const float4 b = *reinterpret_cast<const float4*>(beta + c0 + i);
    y[i + 0] = (c0 + i + 0 < H) ? g.x * (y[i + 0] * o.rstd) + b.x : 0.f;
    y[i + 1] = (c0 + i + 1 < H) ? g.y * (y[i + 1] * o.rstd) + b.y : 0.f;
    y[i + 2] = (c0 + i + 2 < H) ? g.z * (y[i + 2] * o.rstd) + b.z : 0.f;
    y[i + 3] = (c0 + i + 3 < H) ? g.w * (y[i + 3] * o.rstd) + b.w : 0.f;
    yp += (y[i] + y[i + 1]) + (y[i + 2] + y[i + 3]);
  }
  red[1024 + r * 4 + qd] = yp;
  return o;
}

struct RuLnQkvArgs {
  const float *x, *gamma, *beta, *bq, *bk, *bv;
  const unsigned char* img;  // this block's weight images
  float eps;
  float *qn, *Q, *K, *V, *mean, *rstd, *kmask, *qmask;
  long N, rpc;  // rows, rows per CTA
  int H;
};

__global__ void __launch_bounds__(RU_THREADS, 1) ru_ln_qkv_fwd_kernel(RuLnQkvArgs a) {
  CAST_DYN_SMEM(unsigned char, sm);
  // weights landed | x staged | LN(x) staged | K,V product done | Q product done
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  const RuShape s = ru_shape(a.H);
  const int H = a.H;
  unsigned char* wq = sm;
  unsigned char* wkv = wq + s.img1;
  unsigned char* a_hi = wkv + s.img2;
  unsigned char* a_lo = a_hi + s.SL * RU_PA;
  float* tin = reinterpret_cast<float*>(a_lo + s.SL * RU_PA);
  float* out0 = tin + s.dense / 4;
  float* out1 = out0 + s.dense / 4;
  float* vec = out1 + s.dense / 4;  // gamma | beta | bq | bk | bv, 64 floats each
  float* red = vec + 5 * 64;        // [3][128][4]
  const int t = threadIdx.x, warp = t >> 5, r = t & 127, qd = (t >> 7) & 3, c0 = 16 * qd;
  const bool have_cols = c0 < H;
  long long* trace = g_rowk_trace;
  RU_STAMP(0, 0);

  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], RU_ROW_THREADS / 32);
    umma::mbar_init(&bars[2], RU_ROW_THREADS / 32);
    umma::mbar_init(&bars[3], 1);
    umma::mbar_init(&bars[4], 1);
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  for (int i = t; i < 5 * 64; i += RU_THREADS) {
    const int which = i >> 6, c = i & 63;
    const float* src = which == 0 ? a.gamma : which == 1 ? a.beta : which == 2 ? a.bq : which == 3 ? a.bk : a.bv;
    vec[i] = c < H ? src[c] : 0.f;
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const RuRange rg = ru_range(a.N, a.rpc);
  const uint32_t idesc_q = umma::idesc_tf32(128, s.NP), idesc_kv = umma::idesc_tf32(128, 2 * s.NP);

  if (warp == RU_ROW_THREADS / 32) {  // ---------------------------------------------------------------- MMA warp
    if ((t & 31) == 0 && rg.ntiles > 0) {
      umma::mbar_arrive_expect_tx(&bars[0], (uint32_t)(s.img1 + s.img2));
      umma::bulk_g2s(wq, a.img + ru_img_offset(s, 0), (uint32_t)s.img1, &bars[0]);
      umma::bulk_g2s(wkv, a.img + ru_img_offset(s, 1), (uint32_t)s.img2, &bars[0]);
      cast_pdl_wait();
      ru_wait(&bars[0], 0);
      for (int it = 0; it < rg.ntiles; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        ru_wait(&bars[1], par);  // raw x staged: K | V = x [Wk | Wv]
        umma::fence_after_sync();
        ru_issue(tmem + (uint32_t)s.NP, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(wkv),
                 umma::smem_u32(wkv + s.img2 / 2), (uint32_t)(2 * s.NP * 16 + 16), s.KS, idesc_kv, 0u);
        umma::mma_commit(&bars[3]);
        ru_wait(&bars[2], par);  // LN(x) staged (the row warps waited for the K|V product first): Q = LN(x) Wq
        umma::fence_after_sync();
        ru_issue(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(wq), umma::smem_u32(wq + s.img1 / 2),
                 (uint32_t)(s.NP * 16 + 16), s.KS, idesc_q, 0u);
        umma::mma_commit(&bars[4]);
      }
    }
    __syncwarp();
  } else {  // ------------------------------------------------------------------------------------------ row warps
    const int ntot = RU_ROWS * H;
    auto rows_of = [&](int it) {
      const long left = rg.hi - (rg.lo + (long)it * RU_ROWS);
      return left < RU_ROWS ? left : (long)RU_ROWS;
    };
    cast_pdl_wait();   // weights, LayerNorm parameters and biases above do not come from the previous kernel; x does
    if (rg.ntiles > 0) ru_tile_load_async(tin, a.x + rg.lo * H, rows_of(0) * H, ntot);
    const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
    for (int it = 0; it < rg.ntiles; ++it) {
      const long row0 = rg.lo + (long)it * RU_ROWS, row = row0 + r;
      const long vrows = rows_of(it), nvalid = vrows * H;
      const uint32_t par = (uint32_t)(it & 1);
      cp_async_wait<0>();
      ru_row_sync();
      RU_STAMP(it, 1);
      float x[16], q[16];
      ru_row_load(tin, r, c0, H, x);
      ru_stage(a_hi, a_lo, r, qd, s.SL, x);  // (the previous tile's Q product, the last reader of the operand, is done)
      ru_operand_ready(&bars[1]);
      RU_STAMP(it, 2);
      const RuLn ln = ru_layernorm(x, q, vec, vec + 64, a.eps, r, qd, c0, H, red, [&]() {
        if (it + 1 < rg.ntiles) ru_tile_load_async(tin, a.x + (row0 + RU_ROWS) * H, rows_of(it + 1) * H, ntot);
      });
      RU_STAMP(it, 3);
      ru_row_store(out0, r, c0, H, q);
      ru_wait(&bars[3], par);  // K | V accumulated; the operand tile is free again
      umma::fence_after_sync();
      RU_STAMP(it, 4);
      ru_stage(a_hi, a_lo, r, qd, s.SL, q);
      ru_operand_ready(&bars[2]);
      ru_row_sync();
      RU_STAMP(it, 5);
      if (qd == 0 && row < rg.hi) {
        const float4 p = *reinterpret_cast<const float4*>(red + 1024 + r * 4);
        const float ysum = (p.x + p.y) + (p.z + p.w);
        a.mean[row] = ln.mean;
        a.rstd[row] = ln.rstd;
        if (a.kmask) a.kmask[row] = ln.xsum != 0.f ? 1.f : 0.f;
        if (a.qmask) a.qmask[row] = ysum != 0.f ? 1.f : 0.f;
      }
      ru_tile_store(a.qn + row0 * H, out0, nvalid);
      float v[16];
      if (have_cols) {  // K = x Wk + bk
        umma::tmem_ld16(tmem + lane_base + (uint32_t)(s.NP + c0), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += vec[3 * 64 + c0 + i];
        ru_row_store(out1, r, c0, H, v);
      }
      ru_row_sync();
      RU_STAMP(it, 6);
      ru_tile_store(a.K + row0 * H, out1, nvalid);
      if (have_cols) {  // V = x Wv + bv
        umma::tmem_ld16(tmem + lane_base + (uint32_t)(2 * s.NP + c0), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += vec[4 * 64 + c0 + i];
        ru_row_store(out0, r, c0, H, v);
      }
      ru_row_sync();
      RU_STAMP(it, 7);
      ru_tile_store(a.V + row0 * H, out0, nvalid);
      ru_wait(&bars[4], par);
      umma::fence_after_sync();
      RU_STAMP(it, 8);
      if (have_cols) {  // Q = LN(x) Wq + bq
        umma::tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += vec[2 * 64 + c0 + i];
        ru_row_store(out1, r, c0, H, v);
      }
      umma::fence_before_sync();
      ru_row_sync();
      ru_tile_store(a.Q + row0 * H, out1, nvalid);
      RU_STAMP(it, 9);
    }
    cp_async_wait<0>();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 256);
}

struct RuLnFfnArgs {
  const float *y, *gamma, *beta, *b1, *b2;
  const unsigned char* img;
  const int* ids;
  float eps, rate;
  unsigned long long seed;
  const unsigned long long* step;
  int site_h, site_o;
  float *zn, *h1d, *xout, *mean, *rstd;
  long N, rpc;
  int H;
};

__global__ void __launch_bounds__(RU_THREADS, 1) ru_ln_ffn_fwd_kernel(RuLnFfnArgs a) {
  CAST_DYN_SMEM(unsigned char, sm);
  // weights landed | LN(y) staged | hidden staged | first product done | second product done
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  const RuShape s = ru_shape(a.H);
  const int H = a.H;
  unsigned char* w1 = sm;
  unsigned char* w2 = w1 + s.img1;
  unsigned char* a_hi = w2 + s.img1;
  unsigned char* a_lo = a_hi + s.SL * RU_PA;
  float* tin = reinterpret_cast<float*>(a_lo + s.SL * RU_PA);
  float* out0 = tin + s.dense / 4;
  float* out1 = out0 + s.dense / 4;
  float* vec = out1 + s.dense / 4;  // gamma | beta | b1 | b2
  float* red = vec + 4 * 64;
  const int t = threadIdx.x, warp = t >> 5, r = t & 127, qd = (t >> 7) & 3, c0 = 16 * qd;
  const bool have_cols = c0 < H;
  long long* trace = g_rowk_trace;
  RU_STAMP(0, 0);

  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], RU_ROW_THREADS / 32);
    umma::mbar_init(&bars[2], RU_ROW_THREADS / 32);
    umma::mbar_init(&bars[3], 1);
    umma::mbar_init(&bars[4], 1);
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
  for (int i = t; i < 4 * 64; i += RU_THREADS) {
    const int which = i >> 6, c = i & 63;
    const float* src = which == 0 ? a.gamma : which == 1 ? a.beta : which == 2 ? a.b1 : a.b2;
    vec[i] = c < H ? src[c] : 0.f;
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const RuRange rg = ru_range(a.N, a.rpc);
  const uint32_t idesc = umma::idesc_tf32(128, s.NP);
  const uint32_t wpitch = (uint32_t)(s.NP * 16 + 16);

  if (warp == RU_ROW_THREADS / 32) {  // ---------------------------------------------------------------- MMA warp
    if ((t & 31) == 0 && rg.ntiles > 0) {
      umma::mbar_arrive_expect_tx(&bars[0], (uint32_t)(2 * s.img1));
      umma::bulk_g2s(w1, a.img + ru_img_offset(s, 2), (uint32_t)(2 * s.img1), &bars[0]);  // W1T and W2T are adjacent
      cast_pdl_wait();
      ru_wait(&bars[0], 0);
      for (int it = 0; it < rg.ntiles; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        ru_wait(&bars[1], par);
        umma::fence_after_sync();
        ru_issue(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w1), umma::smem_u32(w1 + s.img1 / 2),
                 wpitch, s.KS, idesc, 0u);
        umma::mma_commit(&bars[3]);
        ru_wait(&bars[2], par);
        umma::fence_after_sync();
        ru_issue(tmem + (uint32_t)s.NP, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(w2),
                 umma::smem_u32(w2 + s.img1 / 2), wpitch, s.KS, idesc, 0u);
        umma::mma_commit(&bars[4]);
      }
    }
    __syncwarp();
  } else {  // ------------------------------------------------------------------------------------------ row warps
    const int ntot = RU_ROWS * H;
    auto rows_of = [&](int it) {
      const long left = rg.hi - (rg.lo + (long)it * RU_ROWS);
      return left < RU_ROWS ? left : (long)RU_ROWS;
    };
    cast_pdl_wait();   // weights, LayerNorm parameters and biases above do not come from the previous kernel; y does
    if (rg.ntiles > 0) ru_tile_load_async(tin, a.y + rg.lo * H, rows_of(0) * H, ntot);
    const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
    const Drop dh = make_drop(a.rate, a.seed, a.step, a.site_h);
    const Drop dout = make_drop(a.rate, a.seed, a.step, a.site_o);
    for (int it = 0; it < rg.ntiles; ++it) {
      const long row0 = rg.lo + (long)it * RU_ROWS, row = row0 + r;
      const long vrows = rows_of(it), nvalid = vrows * H;
      const uint32_t par = (uint32_t)(it & 1);
      const float m = (row < rg.hi && a.ids) ? (a.ids[row] != 0 ? 1.f : 0.f) : 1.f;  // fetched early, used last
      cp_async_wait<0>();
      ru_row_sync();
      RU_STAMP(it, 1);
      float x[16], zn[16];
      ru_row_load(tin, r, c0, H, x);
      const RuLn ln = ru_layernorm(x, zn, vec, vec + 64, a.eps, r, qd, c0, H, red, [&]() {
        if (it + 1 < rg.ntiles) ru_tile_load_async(tin, a.y + (row0 + RU_ROWS) * H, rows_of(it + 1) * H, ntot);
      });
      RU_STAMP(it, 2);
      ru_stage(a_hi, a_lo, r, qd, s.SL, zn);  // (the previous tile's second product is done: operand tile free)
      ru_operand_ready(&bars[1]);
      ru_row_store(out0, r, c0, H, zn);
      if (qd == 0 && row < rg.hi) {
        a.mean[row] = ln.mean;
        a.rstd[row] = ln.rstd;
      }
      ru_row_sync();
      RU_STAMP(it, 3);
      ru_tile_store(a.zn + row0 * H, out0, nvalid);
      ru_wait(&bars[3], par);
      umma::fence_after_sync();
      RU_STAMP(it, 4);
      float v[16];
      // this thread's 16 flat indices row*H + c0 + i start even or odd for the whole tile row: the pair-per-hash-word
      // path is taken without a per-pair parity test (the odd-start loop stays out of the hot instruction stream)
      const unsigned long long fi0 = (unsigned long long)(row * H + c0);
      const bool fi_even = (fi0 & 1ull) == 0ull;
      if (have_cols) {  // hidden = dropout(relu(zn W1 + b1))
        umma::tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
        float dm[16];
        if (fi_even) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) drop_mul2_even(dh, fi0 + i, dm[i], dm[i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) drop_mul2(dh, fi0 + i, dm[i], dm[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          v[i] = (c < H) ? fmaxf(v[i] + vec[2 * 64 + c], 0.f) * dm[i] : 0.f;
        }
        ru_row_store(out1, r, c0, H, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      ru_stage(a_hi, a_lo, r, qd, s.SL, v);
      ru_operand_ready(&bars[2]);
      ru_row_sync();
      RU_STAMP(it, 5);
      ru_tile_store(a.h1d + row0 * H, out1, nvalid);
      ru_wait(&bars[4], par);
      umma::fence_after_sync();
      RU_STAMP(it, 6);
      if (have_cols) {  // xout = (dropout(hidden W2 + b2) + zn) * mask
        umma::tmem_ld16(tmem + lane_base + (uint32_t)(s.NP + c0), v);
        float dm[16];
        if (fi_even) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) drop_mul2_even(dout, fi0 + i, dm[i], dm[i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) drop_mul2(dout, fi0 + i, dm[i], dm[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ((v[i] + vec[3 * 64 + ((c0 + i) & 63)]) * dm[i] + zn[i]) * m;
        ru_row_store(out0, r, c0, H, v);
      }
      umma::fence_before_sync();
      ru_row_sync();
      ru_tile_store(a.xout + row0 * H, out0, nvalid);
      RU_STAMP(it, 7);
    }
    cp_async_wait<0>();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 128);
}

static size_t ru_ln_qkv_smem(const RuShape& s) {
  return (size_t)s.img1 + s.img2 + 2 * (size_t)s.SL * RU_PA + 3 * (size_t)s.dense + (5 * 64 + 3 * 512) * sizeof(float);
}
static size_t ru_ln_ffn_smem(const RuShape& s) {
  return 2 * (size_t)s.img1 + 2 * (size_t)s.SL * RU_PA + 3 * (size_t)s.dense + (4 * 64 + 3 * 512) * sizeof(float);
}

constexpr int RU_NUM_SMS = 148;

}  // namespace cast

using namespace cast;

#define CAST_RU_SMEM(kernel, bytes)                                                            \
  {                                                                                            \
    static size_t configured = 48 * 1024;                                                      \
    if ((bytes) > configured) {                                                                \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
      configured = (bytes);                                                                    \
    }                                                                                          \
  }

// persistent grid size: one CTA per SM; cast_rowk_set_grid lowers it (tests: several tiles per CTA at small N)
static int g_rowk_max_ctas = RU_NUM_SMS;
extern "C" int cast_rowk_set_grid(int max_ctas) {
  if (max_ctas < 1) return set_error(CAST_ERR_BAD_ARG, "rowk_set_grid");
  g_rowk_max_ctas = max_ctas < RU_NUM_SMS ? max_ctas : RU_NUM_SMS;
  return CAST_OK;
}

extern "C" int cast_rowk_supported(int H) {
  if (H < 8 || H > 64) return 0;
  const RuShape s = ru_shape(H);
  return ru_ln_qkv_smem(s) <= RU_SMEM_MAX && ru_ln_ffn_smem(s) <= RU_SMEM_MAX;
}

extern "C" size_t cast_rowk_image_bytes(int H) {
  if (H <= 0 || H > 64) return 0;
  return ru_img_block_bytes(ru_shape(H));
}

extern "C" int cast_rowk_presplit(const float* const* weights, int nblocks, int H, void* images, size_t image_bytes,
                                  void* stream) {
  if (!weights || !images || nblocks <= 0) return set_error(CAST_ERR_BAD_ARG, "rowk_presplit");
  if (!cast_rowk_supported(H)) return set_error(CAST_ERR_UNSUPPORTED, "rowk_presplit: hidden_units not supported");
  const size_t per = cast_rowk_image_bytes(H);
  if (image_bytes < per * (size_t)nblocks) return set_error(CAST_ERR_WORKSPACE, "rowk_presplit: image buffer too small");
  for (int b0 = 0; b0 < nblocks; b0 += 16) {
    const int nb = nblocks - b0 < 16 ? nblocks - b0 : 16;
    RuPresplitArgs a;
    for (int b = 0; b < 16; ++b)
      for (int j = 0; j < 5; ++j) a.w[b][j] = b < nb ? weights[(size_t)(b0 + b) * 5 + j] : nullptr;
    for (int b = 0; b < nb; ++b)
      for (int j = 0; j < 5; ++j)
        if (!a.w[b][j]) return set_error(CAST_ERR_BAD_ARG, "rowk_presplit: null weight pointer");
    const RuShape s = ru_shape(H);
    CAST_LAUNCH(ru_presplit_kernel, dim3((unsigned)cdiv(2L * s.NP * 4 * s.SL, 256), RU_NIMG, (unsigned)nb), dim3(256), 0,
                (cudaStream_t)stream, a, H, static_cast<unsigned char*>(images) + (size_t)b0 * per);
  }
  return check_launch("rowk_presplit");
}

extern "C" int cast_rowk_ln_qkv_fwd(const float* x, const float* gamma, const float* beta, const float* bq,
                                    const float* bk, const float* bv, const void* images, long N, int H, float eps,
                                    float* qn, float* Q, float* K, float* V, float* mean, float* rstd, float* kmask,
                                    float* qmask, void* stream) {
  if (!x || !gamma || !beta || !bq || !bk || !bv || !images || !qn || !Q || !K || !V || !mean || !rstd || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "rowk_ln_qkv_fwd");
  if (!cast_rowk_supported(H)) return set_error(CAST_ERR_UNSUPPORTED, "rowk_ln_qkv_fwd: hidden_units not supported");
  const RuShape s = ru_shape(H);
  const long ntiles = cdiv(N, RU_ROWS);
  const int grid = (int)(ntiles < g_rowk_max_ctas ? ntiles : g_rowk_max_ctas);
  const long rpc = cdiv(cdiv(N, grid), 4) * 4;
  RuLnQkvArgs a{x, gamma, beta, bq, bk, bv, static_cast<const unsigned char*>(images), eps, qn, Q, K, V, mean, rstd,
                kmask, qmask, N, rpc, H};
  const size_t smem = ru_ln_qkv_smem(s);
  CAST_RU_SMEM(ru_ln_qkv_fwd_kernel, smem)
  CAST_LAUNCH_DEP(ru_ln_qkv_fwd_kernel, dim3(grid), dim3(RU_THREADS), smem, (cudaStream_t)stream, a);
  return check_launch("rowk_ln_qkv_fwd");
}

extern "C" int cast_rowk_ln_ffn_fwd(const float* y, const float* gamma, const float* beta, const float* b1,
                                    const float* b2, const void* images, const int* ids, float drop_rate,
                                    unsigned long long seed, const unsigned long long* step, int site_hidden,
                                    int site_out, long N, int H, float eps, float* zn, float* h1d, float* xout,
                                    float* mean, float* rstd, void* stream) {
  if (!y || !gamma || !beta || !b1 || !b2 || !images || !zn || !h1d || !xout || !mean || !rstd || N <= 0)
    return set_error(CAST_ERR_BAD_ARG, "rowk_ln_ffn_fwd");
  if (!cast_rowk_supported(H)) return set_error(CAST_ERR_UNSUPPORTED, "rowk_ln_ffn_fwd: hidden_units not supported");
  if (drop_rate < 0.f || drop_rate >= 1.f) return set_error(CAST_ERR_BAD_ARG, "rowk_ln_ffn_fwd: drop_rate");
  const RuShape s = ru_shape(H);
  const long ntiles = cdiv(N, RU_ROWS);
  const int grid = (int)(ntiles < g_rowk_max_ctas ? ntiles : g_rowk_max_ctas);
  const long rpc = cdiv(cdiv(N, grid), 4) * 4;
  RuLnFfnArgs a{y, gamma, beta, b1, b2, static_cast<const unsigned char*>(images), ids, eps, drop_rate, seed, step,
                site_hidden, site_out, zn, h1d, xout, mean, rstd, N, rpc, H};
  const size_t smem = ru_ln_ffn_smem(s);
  CAST_RU_SMEM(ru_ln_ffn_fwd_kernel, smem)
  CAST_LAUNCH_DEP(ru_ln_ffn_fwd_kernel, dim3(grid), dim3(RU_THREADS), smem, (cudaStream_t)stream, a);
  return check_launch("rowk_ln_ffn_fwd");
}

/* tuning hook: device buffer of 148*4*16 int64 that thread 0 of every CTA fills with clock64() phase stamps (NULL: off) */
extern "C" int cast_rowk_set_trace(void* device_buffer) {
#ifndef CAST_EMU
  long long* p = static_cast<long long*>(device_buffer);
  if (cudaMemcpyToSymbol(g_rowk_trace, &p, sizeof(p)) != cudaSuccess) return set_error(CAST_ERR_CUDA, "rowk_set_trace");
#endif
  return CAST_OK;
}

/* 1 if a tcgen05 row kernel gave up waiting on one of its barriers since the last call (synchronises the device) */
extern "C" int cast_rowk_status(int* timed_out) {
  if (!timed_out) return set_error(CAST_ERR_BAD_ARG, "rowk_status");
#ifdef CAST_EMU
  *timed_out = g_rowk_timeout;
  g_rowk_timeout = 0;
#else
  int v = 0, zero = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return set_error(CAST_ERR_CUDA, "rowk_status: device error");
  cudaMemcpyFromSymbol(&v, g_rowk_timeout, sizeof(int));
  cudaMemcpyToSymbol(g_rowk_timeout, &zero, sizeof(int));
  *timed_out = v;
#endif
  return CAST_OK;
}
