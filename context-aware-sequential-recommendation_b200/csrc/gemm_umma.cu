// Dense projections on the 5th-generation tensor cores: the strided GEMM of cast_gemm (tf.layers.dense / conv1d(k=1)
// of modules.py:203-205, :298-306, :333-334 and their gradients) as tcgen05.mma kind::tf32 with the 3xTF32 operand
// split (hi*hi + hi*lo + lo*hi => ~2^-22 relative to sum|a||b|: fp32-grade results, parity tolerance 1e-4 holds),
// fp32 accumulator tile of 128 rows x up to 256 columns in tensor memory, epilogue fused out of TMEM with one thread
// per output row (bias, ReLU, dropout, ReLU-backward mask, residual, padding-row mask — same order as gemm.cu).
// Used for the wide models (hidden_units > 64: BASELINE configs 4 and 5); H <= 64 runs the fused row kernels.
#include "cast_rt.cuh"
#ifndef CAST_EMU
#include "umma.cuh"

namespace cast {

int launch_reduce_partials(const float* partial, int nparts, long count, float* out0, long split, float* out1,
                           cudaStream_t stream);

constexpr int UG_THREADS = 256;
constexpr int UG_M = 128, UG_N = 256;
constexpr int UG_KC = 32;                      // K elements per pipeline stage
constexpr int UG_STAGES = 2;
constexpr int UG_A_PITCH = UG_M * 16 + 16, UG_B_PITCH = UG_N * 16 + 16, UG_SLABS = UG_KC / 4;
constexpr int UG_STAGE_BYTES = 2 * UG_SLABS * (UG_A_PITCH + UG_B_PITCH);
constexpr int UG_CPITCH = UG_N + 4;            // floats per row of the epilogue tile (conflict-free 16-byte stores)
constexpr size_t UG_SMEM = (size_t)UG_STAGES * UG_STAGE_BYTES + 128;
static_assert((size_t)UG_M * UG_CPITCH * 4 <= (size_t)UG_STAGES * UG_STAGE_BYTES, "epilogue tile must fit the stages");

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

struct UGemmArgs {
  const float* A;
  long sam, sak;
  const float* B;
  long sbk, sbn;
  float* C;
  long ldc;
  long M;
  int N;
  long K, klen;
  int ntile;        // output columns per CTA (<= UG_N, multiple of 16)
  UGemmEpi epi;
  float* partials;  // split-K: raw sums [z][M][N]
  int* err;
};

// Per-thread staging plan of one operand: ITEMS (row, 4-element K slab) pairs whose source pointer advances by one
// K chunk per pipeline step.  All index arithmetic happens once per CTA; a chunk costs the loads, 8 cvt.rna, 4 subs
// and two 16-byte shared-memory stores per pair.  KFAST: the source is contiguous along K (pairs of a row are
// neighbours => coalesced float4 / scalar loads along K); otherwise it is contiguous along the operand's rows
// (transposed operand: the four K elements are four loads, each coalesced across the warp's consecutive rows).
template <int R_MAX, bool KFAST>
struct StagePlan {
  static constexpr int ITEMS = R_MAX * UG_SLABS / UG_THREADS;
  const float* ptr[ITEMS];
  int off[ITEMS];        // byte offset inside the hi / lo slab array
  int mode[ITEMS];       // 0 = row beyond the staged tile (nothing stored), 1 = row beyond the operand (zeros), 2 = load
  int kofs[ITEMS];       // first K index of the pair relative to the chunk start
  long sk;

  __device__ __forceinline__ void init(const float* src, long sr, long sk_, long row0, long rows_total, int R,
                                       long kbeg, int pitch) {
    sk = sk_;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int idx = threadIdx.x + i * UG_THREADS;
      int r, c;
      if (KFAST) { r = idx / UG_SLABS; c = idx % UG_SLABS; } else { c = idx / R_MAX; r = idx % R_MAX; }
      const long row = row0 + r;
      kofs[i] = 4 * c;
      off[i] = c * pitch + r * 16;
      mode[i] = (r >= R) ? 0 : (row >= rows_total ? 1 : 2);
      ptr[i] = src + (row < rows_total ? row : 0) * sr + (kbeg + 4 * c) * sk;
    }
  }

  // kleft = number of valid K elements from the chunk start (>= UG_KC for a full chunk)
  template <bool VEC4>
  __device__ __forceinline__ void run(unsigned char* __restrict__ hi, unsigned char* __restrict__ lo, long kleft) {
    float x[ITEMS][4];
    const bool full = kleft >= UG_KC;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) x[i][e] = 0.f;
      if (mode[i] == 2) {
        if (full) {
          if (KFAST && VEC4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(ptr[i]));
            x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[i][e] = __ldg(ptr[i] + e * sk);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (kofs[i] + e < kleft) x[i][e] = __ldg(ptr[i] + e * sk);
        }
      }
      ptr[i] += UG_KC * sk;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (mode[i] == 0) continue;
      const int o = off[i];
      float4 h, l;
      umma::split_tf32(x[i][0], h.x, l.x);
      umma::split_tf32(x[i][1], h.y, l.y);
      umma::split_tf32(x[i][2], h.z, l.z);
      umma::split_tf32(x[i][3], h.w, l.w);
      *reinterpret_cast<float4*>(hi + o) = h;
      *reinterpret_cast<float4*>(lo + o) = l;
    }
  }
};

// Pipeline per CTA (one 128 x <=256 output tile, one K range): while the tensor core works on stage s (three
// tcgen05.mma per 8 K-elements, issued by thread 0, completion tracked by mbar[s]) all 256 threads split-and-stage the
// next K chunk into stage s^1.  A stage is refilled only after the MMAs that read it have arrived on its mbarrier.
template <bool A_KFAST, bool B_KFAST, bool VEC4>
__global__ void __launch_bounds__(UG_THREADS, 1) gemm_umma_kernel(UGemmArgs a) {
  extern __shared__ __align__(128) unsigned char ug_smem[];
  __shared__ __align__(8) uint64_t mbar[UG_STAGES];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, UG_N);
  if (t == 0) {
    umma::mbar_init(&mbar[0], 1);
    umma::mbar_init(&mbar[1], 1);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const long i0 = (long)blockIdx.x * UG_M;
  const int j0 = blockIdx.y * a.ntile;
  const int nN = a.N - j0 < a.ntile ? a.N - j0 : a.ntile;  // live columns of this tile
  const int npad = (nN + 15) & ~15;                   // UMMA N (multiple of 16 for M = 128)
  const long kbeg = (long)blockIdx.z * a.klen;
  const long kend = kbeg + a.klen < a.K ? kbeg + a.klen : a.K;
  const int nchunks = (int)((kend - kbeg + UG_KC - 1) / UG_KC);
  const uint32_t idesc = umma::idesc_tf32(UG_M, npad);
  uint32_t parity[UG_STAGES] = {0u, 0u};
  bool failed = false;

  StagePlan<UG_M, A_KFAST> pa;
  StagePlan<UG_N, B_KFAST> pb;
  pa.init(a.A, a.sam, a.sak, i0, a.M, UG_M, kbeg, UG_A_PITCH);
  // operand B as N x K: row n = output column, element (n,k) at B[k*sbk + n*sbn]
  pb.init(a.B, a.sbn, a.sbk, j0, a.N, npad, kbeg, UG_B_PITCH);

  auto stage = [&](int chunk) {
    unsigned char* base = ug_smem + (size_t)(chunk & 1) * UG_STAGE_BYTES;
    unsigned char* Ahi = base;
    unsigned char* Alo = Ahi + UG_SLABS * UG_A_PITCH;
    unsigned char* Bhi = Alo + UG_SLABS * UG_A_PITCH;
    unsigned char* Blo = Bhi + UG_SLABS * UG_B_PITCH;
    const long kleft = kend - (kbeg + (long)chunk * UG_KC);
    pa.template run<VEC4>(Ahi, Alo, kleft);
    pb.template run<VEC4>(Bhi, Blo, kleft);
    umma::fence_smem_to_async();
  };

  stage(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    __syncthreads();  // stage `ch` is complete in shared memory (and fenced towards the async proxy)
    if (t == 0) {
      umma::fence_after_sync();
      unsigned char* base = ug_smem + (size_t)(ch & 1) * UG_STAGE_BYTES;
      const uint32_t ah = umma::smem_u32(base), al = ah + UG_SLABS * UG_A_PITCH;
      const uint32_t bh = al + UG_SLABS * UG_A_PITCH, bl = bh + UG_SLABS * UG_B_PITCH;
#pragma unroll
      for (int s = 0; s < UG_KC / 8; ++s) {  // a short last chunk is zero padded to UG_KC
        const uint64_t dah = umma::smem_desc(ah + 2 * s * UG_A_PITCH, UG_A_PITCH, 128);
        const uint64_t dal = umma::smem_desc(al + 2 * s * UG_A_PITCH, UG_A_PITCH, 128);
        const uint64_t dbh = umma::smem_desc(bh + 2 * s * UG_B_PITCH, UG_B_PITCH, 128);
        const uint64_t dbl = umma::smem_desc(bl + 2 * s * UG_B_PITCH, UG_B_PITCH, 128);
        umma::mma_tf32(tmem, dal, dbh, idesc, (ch == 0 && s == 0) ? 0u : 1u);
        umma::mma_tf32(tmem, dah, dbl, idesc, 1u);
        umma::mma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      umma::mma_commit(&mbar[ch & 1]);
    }
    if (ch + 1 < nchunks) {
      const int nb = (ch + 1) & 1;
      if (ch >= 1) {  // stage nb was read by the MMAs of chunk ch-1: wait for them before overwriting it
        const bool ok = umma::mbar_wait(&mbar[nb], parity[nb]);
        parity[nb] ^= 1u;
        if (!__syncthreads_and(ok ? 1 : 0)) { failed = true; break; }
      }
      stage(ch + 1);
    }
  }
  if (!failed) {  // drain: the last one or two commits
    const int last = (nchunks - 1) & 1;
    if (nchunks >= 2) {
      const bool ok = umma::mbar_wait(&mbar[last ^ 1], parity[last ^ 1]);
      if (!__syncthreads_and(ok ? 1 : 0)) failed = true;
    }
    if (!failed) {
      const bool ok = umma::mbar_wait(&mbar[last], parity[last]);
      if (!__syncthreads_and(ok ? 1 : 0)) failed = true;
    }
  }
  if (!failed) {
    umma::fence_after_sync();
    // ---- epilogue 1: accumulator rows out of TMEM (thread <-> row, warp halves take alternate 32-column groups)
    // into a row-major shared-memory tile (the pipeline stages are free now)
    float* Ct = reinterpret_cast<float*>(ug_smem);
    const uint32_t q = warp & 3, half = warp >> 2;
    const int r = (int)q * 32 + lane;
#pragma unroll 1
    for (int cb = (int)half * 32; cb < npad; cb += 64) {
      float v[32];
      umma::tmem_ld32(tmem + ((q * 32u) << 16) + (uint32_t)cb, v);
#pragma unroll
      for (int x = 0; x < 32; x += 4)
        *reinterpret_cast<float4*>(&Ct[r * UG_CPITCH + cb + x]) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
    }
    umma::fence_before_sync();
    __syncthreads();
    // ---- epilogue 2: one warp per row, lanes over columns (coalesced reads of bias / act / resid, coalesced stores)
    const UGemmEpi& e = a.epi;
    const Drop d = make_drop(e.drop_rate, e.seed, e.step, e.site);
    float* P = a.partials ? a.partials + (long)blockIdx.z * a.M * a.N : nullptr;
    const int rows = a.M - i0 < UG_M ? (int)(a.M - i0) : UG_M;
    for (int rr = warp; rr < rows; rr += UG_THREADS / 32) {
      const long gi = i0 + rr;
      const float* crow = Ct + rr * UG_CPITCH;
      if (P) {
        float* prow = P + gi * a.N + j0;
        for (int cc = lane; cc < nN; cc += 32) prow[cc] = crow[cc];
        continue;
      }
      const float rm = e.row_ids ? (e.row_ids[gi] != 0 ? 1.f : 0.f) : 1.f;
      const float* actrow = e.act ? e.act + gi * e.ld_act + j0 : nullptr;
      const float* resrow = e.resid ? e.resid + gi * e.ldr + j0 : nullptr;
      float* orow = a.C + gi * a.ldc + j0;
      const unsigned long long dbase = (unsigned long long)(gi * a.N + j0);
      for (int cc = lane; cc < nN; cc += 32) {
        float c = crow[cc];
        if (e.bias) c += e.bias[j0 + cc];
        if (e.relu) c = fmaxf(c, 0.f);
        c *= drop_mul(d, dbase + cc);
        if (actrow) c *= (actrow[cc] > 0.f) ? e.act_scale : 0.f;
        if (resrow) c += resrow[cc];
        orow[cc] = c * rm;
      }
    }
  } else if (t == 0) {
    atomicExch(a.err, 1);
  }
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, UG_N);
}

// device-side watchdog flag shared by all tensor-core GEMM launches of the process (0 = ok)
__device__ int g_umma_gemm_err = 0;

template <bool AK, bool BK, bool V4>
static void launch_inst(const UGemmArgs& a, dim3 grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(gemm_umma_kernel<AK, BK, V4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UG_SMEM);
    configured = true;
  }
  gemm_umma_kernel<AK, BK, V4><<<grid, dim3(UG_THREADS), UG_SMEM, stream>>>(a);
}

int gemm_umma_launch(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc, long M,
                     int N, long K, const UGemmEpi& epi, int splits, float* partials, cudaStream_t stream) {
  static int* err_ptr = nullptr;
  if (!err_ptr) cudaGetSymbolAddress(reinterpret_cast<void**>(&err_ptr), g_umma_gemm_err);
  UGemmArgs a;
  a.A = A; a.sam = sam; a.sak = sak; a.B = B; a.sbk = sbk; a.sbn = sbn; a.C = C; a.ldc = ldc; a.M = M; a.N = N;
  a.K = K; a.epi = epi; a.partials = partials; a.err = err_ptr;
  // column tile: 256 wide unless that leaves most of the 148 SMs idle (short M: C4-sized batches, weight gradients)
  int ntile = UG_N;
  if (splits == 1)
    while (ntile > 64 && cdiv(M, UG_M) * cdiv(N, ntile) < 148) ntile /= 2;
  a.ntile = ntile;
  if (splits > 1) {  // about two waves of CTAs over the reduction dimension are enough (fewer partials to reduce)
    const long tiles = cdiv(M, UG_M) * cdiv(N, ntile);
    long want = cdiv(296, tiles);
    if (want < 1) want = 1;
    if (want < splits) splits = (int)want;
  }
  long klen = cdiv(K, splits);
  klen = cdiv(klen, UG_KC) * UG_KC;
  a.klen = klen;
  const int zs = (int)cdiv(K, klen);
  dim3 grid((unsigned)cdiv(M, UG_M), (unsigned)cdiv(N, ntile), (unsigned)zs);
  const bool ak = (sak == 1), bk = (sbk == 1);
  // float4 loads along K need 16-byte aligned rows in every K-contiguous operand (klen is a multiple of 32)
  const bool v4 = (!ak || (sam % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0)) &&
                  (!bk || (sbn % 4 == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0)) && (ak || bk);
  if (ak && bk) { if (v4) launch_inst<true, true, true>(a, grid, stream); else launch_inst<true, true, false>(a, grid, stream); }
  else if (ak)  { if (v4) launch_inst<true, false, true>(a, grid, stream); else launch_inst<true, false, false>(a, grid, stream); }
  else if (bk)  { if (v4) launch_inst<false, true, true>(a, grid, stream); else launch_inst<false, true, false>(a, grid, stream); }
  else launch_inst<false, false, false>(a, grid, stream);
  int rc = check_launch("gemm(umma)");
  if (rc) return rc;
  if (partials)
    return launch_reduce_partials(partials, zs, M * (long)N, C, M * (long)N, (float*)nullptr, stream);
  return CAST_OK;
}

}  // namespace cast

extern "C" int cast_gemm_tensor_status(int* host_flag, void* stream) {
  if (!host_flag) return cast::set_error(CAST_ERR_BAD_ARG, "gemm_tensor_status");
  cudaStreamSynchronize((cudaStream_t)stream);
  if (cudaMemcpyFromSymbol(host_flag, cast::g_umma_gemm_err, sizeof(int)) != cudaSuccess)
    return cast::set_error(CAST_ERR_CUDA, "gemm_tensor_status");
  return CAST_OK;
}
#else
#include <stddef.h>
extern "C" int cast_gemm_tensor_status(int* host_flag, void* stream) {
  (void)stream;
  if (host_flag) *host_flag = 0;
  return 0;
}
#endif
