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

// One element of the update.  The contractions are written out (fmaf) so that every kernel below — scalar, 16-byte
// and the peer-summing form — rounds identically: a data-parallel run stays bit-comparable with a single-process one.
__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, bool decay, float l2, float b1,
                                          float omb1, float b2, float omb2, float lr_t, float eps) {
  if (decay) g = fmaf(l2, w, g);
  m = fmaf(b1, m, omb1 * g);
  v = fmaf(b2, v, (omb2 * g) * g);
  w = w - (lr_t * m) / (sqrtf(v) + eps);
}

__global__ void adam_step_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m,
                                 float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                                 const float* __restrict__ gdenom, float l2, long l2_lo, long l2_hi,
                                 const AdamState* __restrict__ st) {
  cast_pdl_wait();
  cast_pdl_trigger();
  const float gs = gdenom ? 1.0f / *gdenom : 1.0f;
  const float lr_t = lr * sqrtf(1.0f - st->b2p) / (1.0f - st->b1p);
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float wi = w[i], mi = m[i], vi = v[i];
    adam_elem(wi, mi, vi, grad[i] * gs, i >= l2_lo && i < l2_hi, l2, b1, omb1, b2, omb2, lr_t, eps);
    m[i] = mi;
    v[i] = vi;
    w[i] = wi;
  }
}

// same update, four elements per thread (16-byte loads/stores); n4 = n/4, pointers 16-byte aligned
__global__ void adam_step_vec4_kernel(float4* __restrict__ w, const float4* __restrict__ grad, float4* __restrict__ m,
                                      float4* __restrict__ v, long n4, float lr, float b1, float b2, float eps,
                                      const float* __restrict__ gdenom, float l2, long l2_lo, long l2_hi,
                                      const AdamState* __restrict__ st) {
  cast_pdl_wait();
  cast_pdl_trigger();
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
      adam_elem(wv[e], mv[e], vv[e], gv[e] * gs, idx >= l2_lo && idx < l2_hi, l2, b1, omb1, b2, omb2, lr_t, eps);
    }
    m[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
    v[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    w[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
}

// Data-parallel form: the gradient of element i is the sum over the n ranks' buffers in rank order (own buffer + peer
// mappings read over NVLink), formed inside the optimizer pass instead of by a separate reduction launch.  The loads
// of one element group from all ranks are issued together (a serial walk pays one NVLink round trip per rank).  The
// buffers hold `np` gradient numerators followed by `tail` >= 3 sums [loss_sum, auc_sum, count, ...]: the reduced
// count is the gradient denominator; every reduced element is also written to g_red (loss read-back, diagnostics).
constexpr int ADAM_PEER_BATCH = 8;
__global__ void adam_step_peers_kernel(float* __restrict__ w, const float* const* __restrict__ g, int n,
                                       float* __restrict__ g_red, float* __restrict__ m, float* __restrict__ v,
                                       long np, long tail, float lr, float b1, float b2, float eps, float l2,
                                       long l2_lo, long l2_hi, const AdamState* __restrict__ st) {
  __shared__ float s_count;
  cast_pdl_wait();
  cast_pdl_trigger();
  if (threadIdx.x == 0) {
    float c = g[0][np + 2];
    for (int r = 1; r < n; ++r) c += g[r][np + 2];
    s_count = c;
  }
  __syncthreads();
  const float gs = 1.0f / s_count;
  const float lr_t = lr * sqrtf(1.0f - st->b2p) / (1.0f - st->b1p);
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  const long total = np + tail;
  for (long i0 = 4 * ((long)blockIdx.x * blockDim.x + threadIdx.x); i0 < total; i0 += 4 * (long)gridDim.x * blockDim.x) {
    float gv[4] = {0.f, 0.f, 0.f, 0.f};
    const int cnt = total - i0 < 4 ? (int)(total - i0) : 4;
    if (cnt == 4) {
      for (int r0 = 0; r0 < n; r0 += ADAM_PEER_BATCH) {
        float4 q[ADAM_PEER_BATCH];
#pragma unroll
        for (int u = 0; u < ADAM_PEER_BATCH; ++u)
          if (r0 + u < n) q[u] = *reinterpret_cast<const float4*>(g[r0 + u] + i0);
#pragma unroll
        for (int u = 0; u < ADAM_PEER_BATCH; ++u)
          if (r0 + u < n) {
            if (r0 + u == 0) {
              gv[0] = q[u].x; gv[1] = q[u].y; gv[2] = q[u].z; gv[3] = q[u].w;
            } else {
              gv[0] += q[u].x; gv[1] += q[u].y; gv[2] += q[u].z; gv[3] += q[u].w;
            }
          }
      }
      *reinterpret_cast<float4*>(g_red + i0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    } else {
      for (int e = 0; e < cnt; ++e) {
        float sacc = g[0][i0 + e];
        for (int r = 1; r < n; ++r) sacc += g[r][i0 + e];
        gv[e] = sacc;
        g_red[i0 + e] = sacc;
      }
    }
    const int na = np - i0 >= 4 ? 4 : (np - i0 > 0 ? (int)(np - i0) : 0);  // elements of this group that are parameters
    if (na == 4) {
      const float4 wi = *reinterpret_cast<const float4*>(w + i0), mi = *reinterpret_cast<const float4*>(m + i0),
                   vi = *reinterpret_cast<const float4*>(v + i0);
      float wv[4] = {wi.x, wi.y, wi.z, wi.w}, mv[4] = {mi.x, mi.y, mi.z, mi.w}, vv[4] = {vi.x, vi.y, vi.z, vi.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long idx = i0 + e;
        adam_elem(wv[e], mv[e], vv[e], gv[e] * gs, idx >= l2_lo && idx < l2_hi, l2, b1, omb1, b2, omb2, lr_t, eps);
      }
      *reinterpret_cast<float4*>(m + i0) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + i0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      *reinterpret_cast<float4*>(w + i0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
    } else {
      for (int e = 0; e < na; ++e) {
        const long idx = i0 + e;
        float wi = w[idx], mi = m[idx], vi = v[idx];
        adam_elem(wi, mi, vi, gv[e] * gs, idx >= l2_lo && idx < l2_hi, l2, b1, omb1, b2, omb2, lr_t, eps);
        m[idx] = mi;
        v[idx] = vi;
        w[idx] = wi;
      }
    }
  }
}

__global__ void adam_advance_kernel(AdamState* st, float b1, float b2) {
  cast_pdl_wait();
  cast_pdl_trigger();
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
    CAST_LAUNCH_DEP(adam_step_vec4_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream,
                reinterpret_cast<float4*>(w), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m),
                reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps, gdenom, l2, l2_lo, l2_hi,
                static_cast<const AdamState*>(state));
  } else {
    long g = cdiv(n, 256);
    if (g > 148 * 8) g = 148 * 8;
    CAST_LAUNCH_DEP(adam_step_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, w, grad, m, v, n, lr, beta1,
                beta2, eps, gdenom, l2, l2_lo, l2_hi, static_cast<const AdamState*>(state));
  }
  if ((rc = check_launch("adam_step"))) return rc;
  if (advance) {
    CAST_LAUNCH_DEP(adam_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, static_cast<AdamState*>(state),
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

/* Data-parallel TF-Adam: grads is a DEVICE array of n_ranks pointers (this rank's buffer and the peer mappings of the
 * others, each 16-byte aligned, n + tail floats: gradient numerators then [loss_sum, auc_sum, count, ...]); the update
 * uses sum_r grads[r][i] (rank order) / sum_r count_r and the reduced buffer is left in g_red[n + tail].  Replaces
 * cast_peer_reduce + cast_adam_tf_step; the caller brackets it with cast_peer_barrier as for cast_peer_reduce. */
extern "C" int cast_adam_tf_step_peers(float* w, const void* const* grads, int n_ranks, float* g_red, float* m, float* v,
                                       long n, long tail, float lr, float beta1, float beta2, float eps, float l2,
                                       long l2_lo, long l2_hi, void* state, void* stream) {
  if (!w || !grads || !g_red || !m || !v || !state || n <= 0 || tail < 3 || n_ranks < 1)
    return set_error(CAST_ERR_BAD_ARG, "adam_tf_step_peers");
  if ((((uintptr_t)w | (uintptr_t)g_red | (uintptr_t)m | (uintptr_t)v) & 15) != 0)
    return set_error(CAST_ERR_BAD_ARG, "adam_tf_step_peers: buffers must be 16-byte aligned");
  long g = cdiv(cdiv(n + tail, 4), 256);
  if (g > 148 * 8) g = 148 * 8;
  CAST_LAUNCH_DEP(adam_step_peers_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, w,
              reinterpret_cast<const float* const*>(grads), n_ranks, g_red, m, v, n, tail, lr, beta1, beta2, eps, l2,
              l2_lo, l2_hi, static_cast<const AdamState*>(state));
  int rc;
  if ((rc = check_launch("adam_step_peers"))) return rc;
  CAST_LAUNCH_DEP(adam_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, static_cast<AdamState*>(state), beta1,
              beta2);
  return check_launch("adam_advance");
}
