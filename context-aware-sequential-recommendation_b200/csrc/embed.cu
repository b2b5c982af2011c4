// K1 / K11: zero-padded table gather x sqrt(H) + position row + context add -> dropout -> padding mask
// (reference: modules.py:148-160 `embedding`, models/sasrec.py:27-62, models/cast_1.py:62-91), and the
// element-wise backward companions (dropout/mask backward, concat+dropout for the CAST merges,
// models/cast_4.py:111-127).  All HBM-bound: one warp per [H] row, lanes stride the row.
#include "cast_rt.cuh"

namespace cast {

// A warp walks rows (grid-stride); lane l owns the column pairs 2l, 2l + 64, ... of each: 8-byte loads / stores when H
// is even and the rows are 8-byte aligned, one dropout hash word per pair, and the per-thread setup (dropout key) is
// paid once per EMB_ROWS_PER_WARP rows instead of once per two elements.
constexpr int EMB_ROWS_PER_WARP = 4;
__global__ void embed_fwd_kernel(const int* __restrict__ ids, TableRef table, int V, int H, long N,
                                 int T, float scale, const float* __restrict__ pos, const float* __restrict__ add,
                                 float rate, unsigned long long seed, const unsigned long long* step, int site,
                                 const int* __restrict__ mask_ids, float* __restrict__ out) {
  cast_pdl_wait();
  cast_pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const Drop d = make_drop(rate, seed, step, site);
  const bool heven = (H & 1) == 0;
  const bool vio = heven && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pos) |
                              reinterpret_cast<uintptr_t>(add)) & 7) == 0;
#pragma unroll
  for (int k = 0; k < EMB_ROWS_PER_WARP; ++k) {
    const long row = warp * EMB_ROWS_PER_WARP + k;
    if (row >= N) return;
    const int id = ids[row];
    const bool live = (id > 0) && (id < V);  // row 0 is the zero pad (modules.py:154-156)
    const float m = mask_ids ? (mask_ids[row] != 0 ? 1.f : 0.f) : 1.f;
    const float* trow = table.row(live ? id : 0, H);
    const float* prow = pos ? pos + (long)(row % T) * H : nullptr;
    const float* arow = add ? add + row * H : nullptr;
    const bool vt = vio && (reinterpret_cast<uintptr_t>(trow) & 7) == 0;
    for (int c = 2 * lane; c < H; c += 64) {
      const bool two = c + 1 < H;
      float v0 = 0.f, v1 = 0.f;
      if (live) {
        if (vt) {
          const float2 tv = *reinterpret_cast<const float2*>(trow + c);
          v0 = tv.x * scale;
          v1 = tv.y * scale;
        } else {
          v0 = trow[c] * scale;
          if (two) v1 = trow[c + 1] * scale;
        }
      }
      if (prow) {
        if (vio) {
          const float2 pv = *reinterpret_cast<const float2*>(prow + c);
          v0 += pv.x;
          v1 += pv.y;
        } else {
          v0 += prow[c];
          if (two) v1 += prow[c + 1];
        }
      }
      if (arow) {
        if (vio) {
          const float2 av = *reinterpret_cast<const float2*>(arow + c);
          v0 += av.x;
          v1 += av.y;
        } else {
          v0 += arow[c];
          if (two) v1 += arow[c + 1];
        }
      }
      const unsigned long long idx = (unsigned long long)(row * H + c);
      float m0, m1;
      if ((idx & 1ull) == 0ull) {
        drop_mul2_even(d, idx, m0, m1);
      } else {
        drop_mul2(d, idx, m0, m1);
      }
      v0 = v0 * m0 * m;
      v1 = v1 * m1 * m;
      if (vio) {
        *reinterpret_cast<float2*>(out + row * H + c) = make_float2(v0, v1);
      } else {
        out[row * H + c] = v0;
        if (two) out[row * H + c + 1] = v1;
      }
    }
  }
}

// out_m = in * rowmask ; out_md = in * rowmask * dropout_multiplier   (either output may be null)
__global__ void mask_dropout_kernel(const float* __restrict__ in, const int* __restrict__ mask_ids, float rate,
                                    unsigned long long seed, const unsigned long long* step, int site, long N,
                                    int H, float* __restrict__ out_m, float* __restrict__ out_md) {
  const Drop d = make_drop(rate, seed, step, site);
  const long total = N * H;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / H;
    const float m = mask_ids ? (mask_ids[row] != 0 ? 1.f : 0.f) : 1.f;
    const float g = in[i] * m;
    if (out_m) out_m[i] = g;
    if (out_md) out_md[i] = g * drop_mul(d, (unsigned long long)i);
  }
}

// cat[n, s*H + c] = src_s[n, c] * dropA(n*wa*H + col) (s < wa) * dropB(n*k*H + col)
struct CatSrc {
  const float* p[4];
};
struct CatDst {
  float* p[4];
};

__global__ void concat_dropout_fwd_kernel(CatSrc src, int k, int wa, long N, int H, float rateA, int siteA,
                                          float rateB, int siteB, unsigned long long seed,
                                          const unsigned long long* step, float* __restrict__ cat) {
  const Drop dA = make_drop(rateA, seed, step, siteA);
  const Drop dB = make_drop(rateB, seed, step, siteB);
  const int W = k * H;
  const long total = N * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / W;
    const int col = (int)(i - n * W);
    const int s = col / H;
    float v = src.p[s][n * H + (col - s * H)];
    if (s < wa) v *= drop_mul(dA, (unsigned long long)(n * (long)(wa * H) + col));
    v *= drop_mul(dB, (unsigned long long)i);
    cat[i] = v;
  }
}

__global__ void concat_dropout_bwd_kernel(const float* __restrict__ dcat, int k, int wa, long N, int H, float rateA,
                                          int siteA, float rateB, int siteB, unsigned long long seed,
                                          const unsigned long long* step, CatDst dst) {
  const Drop dA = make_drop(rateA, seed, step, siteA);
  const Drop dB = make_drop(rateB, seed, step, siteB);
  const int W = k * H;
  const long total = N * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / W;
    const int col = (int)(i - n * W);
    const int s = col / H;
    if (!dst.p[s]) continue;
    float g = dcat[i];
    if (s < wa) g *= drop_mul(dA, (unsigned long long)(n * (long)(wa * H) + col));
    g *= drop_mul(dB, (unsigned long long)i);
    dst.p[s][n * H + (col - s * H)] = g;
  }
}

__global__ void dropout_keep_kernel(float rate, unsigned long long seed, const unsigned long long* step, int site,
                                    long n, unsigned char* __restrict__ keep) {
  const Drop d = make_drop(rate, seed, step, site);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    keep[i] = (d.thresh == 0u || drop_keep(d, (unsigned long long)i)) ? 1 : 0;
}

__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                             long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

// out = dy * (act > 0 ? scale : 0)   (ReLU backward; with scale = 1/(1-rate) also dropout∘ReLU backward)
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ act, float scale,
                                float* __restrict__ out, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = act[i] > 0.f ? dy[i] * scale : 0.f;
}

static inline int ew_grid(long total) {
  long g = cdiv(total, 256);
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace cast

using namespace cast;

extern "C" int cast_embed_fwd(const int* ids, const float* table, int V, int H, long N, int T, float scale,
                              const float* pos, const float* add, float drop_rate, unsigned long long seed,
                              const unsigned long long* step, int site, const int* mask_ids, float* out,
                              void* stream) {
  if (!ids || !table || !out || H <= 0 || N < 0 || T <= 0 || V <= 0) return set_error(CAST_ERR_BAD_ARG, "embed_fwd");
  if (drop_rate < 0.f || drop_rate >= 1.f) return set_error(CAST_ERR_BAD_ARG, "embed_fwd: drop_rate");
  if (N == 0) return CAST_OK;
  const int wpb = 8;
  CAST_LAUNCH_DEP(embed_fwd_kernel, dim3((unsigned)cdiv(N, wpb * EMB_ROWS_PER_WARP)), dim3(32 * wpb), 0,
              (cudaStream_t)stream, ids, table_ref(table), V, H, N, T, scale, pos, add, drop_rate, seed, step, site, mask_ids, out);
  return check_launch("embed_fwd");
}

extern "C" int cast_embed_fwd_sharded(const int* ids, const float* const* shards, int nshards, int V, int H, long N,
                                      int T, float scale, const float* pos, const float* add, float drop_rate,
                                      unsigned long long seed, const unsigned long long* step, int site,
                                      const int* mask_ids, float* out, void* stream) {
  if (!ids || !shards || nshards < 1 || !out || H <= 0 || N < 0 || T <= 0 || V <= 0)
    return set_error(CAST_ERR_BAD_ARG, "embed_fwd_sharded");
  if (drop_rate < 0.f || drop_rate >= 1.f) return set_error(CAST_ERR_BAD_ARG, "embed_fwd_sharded: drop_rate");
  if (N == 0) return CAST_OK;
  const int wpb = 8;
  CAST_LAUNCH_DEP(embed_fwd_kernel, dim3((unsigned)cdiv(N, wpb * EMB_ROWS_PER_WARP)), dim3(32 * wpb), 0,
              (cudaStream_t)stream, ids, table_ref(shards, nshards), V, H, N, T, scale, pos, add, drop_rate, seed, step, site, mask_ids, out);
  return check_launch("embed_fwd_sharded");
}

extern "C" int cast_mask_dropout(const float* in, const int* mask_ids, float drop_rate, unsigned long long seed,
                                 const unsigned long long* step, int site, long N, int H, float* out_masked,
                                 float* out_masked_dropped, void* stream) {
  if (!in || N < 0 || H <= 0) return set_error(CAST_ERR_BAD_ARG, "mask_dropout");
  if (N == 0) return CAST_OK;
  CAST_LAUNCH(mask_dropout_kernel, dim3(ew_grid(N * H)), dim3(256), 0, (cudaStream_t)stream, in, mask_ids,
              drop_rate, seed, step, site, N, H, out_masked, out_masked_dropped);
  return check_launch("mask_dropout");
}

extern "C" int cast_concat_dropout_fwd(const float* const* srcs, int k, int width_a, long N, int H, float rate_a,
                                       int site_a, float rate_b, int site_b, unsigned long long seed,
                                       const unsigned long long* step, float* cat, void* stream) {
  if (!srcs || k < 1 || k > 4 || width_a < 0 || width_a > k || !cat) return set_error(CAST_ERR_BAD_ARG, "concat_fwd");
  CatSrc s;
  for (int i = 0; i < 4; ++i) s.p[i] = i < k ? srcs[i] : nullptr;
  if (N == 0) return CAST_OK;
  CAST_LAUNCH(concat_dropout_fwd_kernel, dim3(ew_grid(N * H * k)), dim3(256), 0, (cudaStream_t)stream, s, k,
              width_a, N, H, rate_a, site_a, rate_b, site_b, seed, step, cat);
  return check_launch("concat_dropout_fwd");
}

extern "C" int cast_concat_dropout_bwd(const float* dcat, int k, int width_a, long N, int H, float rate_a, int site_a,
                                       float rate_b, int site_b, unsigned long long seed,
                                       const unsigned long long* step, float* const* dsts, void* stream) {
  if (!dsts || k < 1 || k > 4 || width_a < 0 || width_a > k || !dcat) return set_error(CAST_ERR_BAD_ARG, "concat_bwd");
  CatDst d;
  for (int i = 0; i < 4; ++i) d.p[i] = i < k ? dsts[i] : nullptr;
  if (N == 0) return CAST_OK;
  CAST_LAUNCH(concat_dropout_bwd_kernel, dim3(ew_grid(N * H * k)), dim3(256), 0, (cudaStream_t)stream, dcat, k,
              width_a, N, H, rate_a, site_a, rate_b, site_b, seed, step, d);
  return check_launch("concat_dropout_bwd");
}

extern "C" int cast_dropout_keep(float drop_rate, unsigned long long seed, const unsigned long long* step, int site,
                                 long n, unsigned char* keep, void* stream) {
  if (!keep || n < 0) return set_error(CAST_ERR_BAD_ARG, "dropout_keep");
  if (n == 0) return CAST_OK;
  CAST_LAUNCH(dropout_keep_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, drop_rate, seed, step,
              site, n, keep);
  return check_launch("dropout_keep");
}

extern "C" int cast_add(const float* a, const float* b, float* out, long n, void* stream) {
  if (!a || !b || !out || n < 0) return set_error(CAST_ERR_BAD_ARG, "add");
  if (n == 0) return CAST_OK;
  CAST_LAUNCH(axpby_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, a, b, out, n);
  return check_launch("add");
}

extern "C" int cast_relu_bwd(const float* dy, const float* act, float scale, float* out, long n, void* stream) {
  if (!dy || !act || !out || n < 0) return set_error(CAST_ERR_BAD_ARG, "relu_bwd");
  if (n == 0) return CAST_OK;
  CAST_LAUNCH(relu_bwd_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, dy, act, scale, out, n);
  return check_launch("relu_bwd");
}
