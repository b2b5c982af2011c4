// K8: tf.train.AdamOptimizer(learning_rate=lr, beta2=0.98) (models/sasrec.py:120-121) as one fused pass over the
// flat parameter buffer.  TF-1.15 semantics: lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t); m,v updated on EVERY element of
// every variable (the embedding tables too: the zero-pad concat makes their gradient dense), epsilon
// added to sqrt(v) outside the root.  beta powers are fp32 state multiplied after the update, like TF's
// beta1_power / beta2_power variables.  HBM-bound: 16 B read + 12 B written per parameter.
#include "cast_rt.cuh"

namespace cast {

struct AdamState {
  float b1p, b2p;
  unsigned long long step;
};

__global__ void adam_init_kernel(AdamState* st, float b1, float b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->b1p = b1;
    st->b2p = b2;
    st->step = 0ull;
  }
}

__global__ void adam_step_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m,
                                 float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                                 const float* __restrict__ gdenom, float l2, long l2_lo, long l2_hi,
                                 const AdamState* __restrict__ st) {
  const float gs = gdenom ? 1.0f / *gdenom : 1.0f;
  const float lr_t = lr * sqrtf(1.0f - st->b2p) / (1.0f - st->b1p);
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float wi = w[i];
    float g = grad[i] * gs;
    if (i >= l2_lo && i < l2_hi) g = fmaf(l2, wi, g);
    const float mi = b1 * m[i] + omb1 * g;
    const float vi = b2 * v[i] + omb2 * g * g;
    m[i] = mi;
    v[i] = vi;
    w[i] = wi - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// same update, four elements per thread (16-byte loads/stores); n4 = n/4, pointers 16-byte aligned
__global__ void adam_step_vec4_kernel(float4* __restrict__ w, const float4* __restrict__ grad, float4* __restrict__ m,
                                      float4* __restrict__ v, long n4, float lr, float b1, float b2, float eps,
                                      const float* __restrict__ gdenom, float l2, long l2_lo, long l2_hi,
                                      const AdamState* __restrict__ st) {
  const float gs = gdenom ? 1.0f / *gdenom : 1.0f;
  const float lr_t = lr * sqrtf(1.0f - st->b2p) / (1.0f - st->b1p);
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 wi = w[i], gi = grad[i], mi = m[i], vi = v[i];
    float wv[4] = {wi.x, wi.y, wi.z, wi.w}, gv[4] = {gi.x, gi.y, gi.z, gi.w};
    float mv[4] = {mi.x, mi.y, mi.z, mi.w}, vv[4] = {vi.x, vi.y, vi.z, vi.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long idx = 4 * i + e;
      float g = gv[e] * gs;
      if (idx >= l2_lo && idx < l2_hi) g = fmaf(l2, wv[e], g);
      mv[e] = b1 * mv[e] + omb1 * g;
      vv[e] = b2 * vv[e] + omb2 * g * g;
      wv[e] = wv[e] - lr_t * mv[e] / (sqrtf(vv[e]) + eps);
    }
    m[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
    v[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    w[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
}

__global__ void adam_advance_kernel(AdamState* st, float b1, float b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->b1p *= b1;
    st->b2p *= b2;
    st->step += 1ull;
  }
}

}  // namespace cast

using namespace cast;

extern "C" int cast_adam_init_state(void* state, float beta1, float beta2, void* stream) {
  if (!state) return set_error(CAST_ERR_BAD_ARG, "adam_init_state");
  CAST_LAUNCH(adam_init_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, static_cast<AdamState*>(state), beta1,
              beta2);
  return check_launch("adam_init");
}

// the update of one contiguous range; `advance` != 0 also steps the beta powers / step counter afterwards (callers that
// update several ranges of one model — a row shard of the item table plus the replicated weights — advance once)
extern "C" int cast_adam_tf_range(float* w, const float* grad, float* m, float* v, long n, float lr, float beta1,
                                  float beta2, float eps, const float* gdenom, float l2, long l2_lo, long l2_hi,
                                  void* state, int advance, void* stream) {
  if (!w || !grad || !m || !v || !state || n <= 0) return set_error(CAST_ERR_BAD_ARG, "adam_tf_range");
  const bool vec = (n % 4 == 0) && ((((uintptr_t)w | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  int rc;
  if (vec) {
    long g = cdiv(n / 4, 256);
    if (g > 148 * 8) g = 148 * 8;
    CAST_LAUNCH(adam_step_vec4_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream,
                reinterpret_cast<float4*>(w), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m),
                reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps, gdenom, l2, l2_lo, l2_hi,
                static_cast<const AdamState*>(state));
  } else {
    long g = cdiv(n, 256);
    if (g > 148 * 8) g = 148 * 8;
    CAST_LAUNCH(adam_step_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, w, grad, m, v, n, lr, beta1,
                beta2, eps, gdenom, l2, l2_lo, l2_hi, static_cast<const AdamState*>(state));
  }
  if ((rc = check_launch("adam_step"))) return rc;
  if (advance) {
    CAST_LAUNCH(adam_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, static_cast<AdamState*>(state),
                beta1, beta2);
    return check_launch("adam_advance");
  }
  return CAST_OK;
}

extern "C" int cast_adam_tf_step(float* w, const float* grad, float* m, float* v, long n, float lr, float beta1,
                                 float beta2, float eps, const float* gdenom, float l2, long l2_lo, long l2_hi,
                                 void* state, void* stream) {
  return cast_adam_tf_range(w, grad, m, v, n, lr, beta1, beta2, eps, gdenom, l2, l2_lo, l2_hi, state, 1, stream);
}
