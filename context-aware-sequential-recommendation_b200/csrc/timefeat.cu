// Time-context ids on the device, straight from raw int64 timestamps (reference util.py:24-43 hour / weekday of a
// record, util.py:73-120 get_timedelta_bin, sampler.py:61-72 / util.py:276-289 which apply them per position):
//
//   bin[b,t]  = get_timedelta_bin(ref[b] - ts[b,t])      ref[b] = ts[b,T-1] (the newest event of the left-padded row)
//   hour[b,t] = (ts mod 86400) / 3600 + 1                 unless the caller passes its own reference times
//   day[b,t]  = (ts / 86400 + 3) mod 7 + 1                Monday = 1 (1970-01-01 was a Thursday)
//
// The bin rule is a non-decreasing step function of the integer delta, so the host tabulates it ONCE per dataset as
// `edges[k] = smallest delta whose bin is > k` with the reference's own scalar expression (float64 log included,
// data.time_bin_edges) and the kernel only counts edges <= delta: bit-exact for the linear and the log scale without
// evaluating log() on the device.  Padding positions (id == 0) get 0 / 0 / 0 like the zero-initialised host arrays.
#include "cast_rt.cuh"

namespace cast {

__global__ void time_features_kernel(const long long* __restrict__ ts, const long long* __restrict__ ref,
                                     const int* __restrict__ ids, int B, int T, const long long* __restrict__ edges,
                                     int n_edges, int* __restrict__ bins, int* __restrict__ hours,
                                     int* __restrict__ days) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long)B * T) return;
  const int b = (int)(n / T);
  int bin = 0, hour = 0, day = 0;
  if (ids[n] != 0) {
    const long long t = ts[n];
    const long long r = ref ? ref[b] : ts[(long)b * T + T - 1];
    const long long delta = r - t;
    int lo = 0, hi = n_edges;  // number of edges <= delta
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (edges[mid] <= delta) lo = mid + 1; else hi = mid;
    }
    bin = lo;
    const long long sod = ((t % 86400) + 86400) % 86400;
    hour = (int)(sod / 3600) + 1;
    long long dd = (t - sod) / 86400 + 3;
    day = (int)(((dd % 7) + 7) % 7) + 1;
  }
  bins[n] = bin;
  hours[n] = hour;
  days[n] = day;
}

}  // namespace cast

using namespace cast;

extern "C" int cast_time_features(const long long* ts, const long long* ref, const int* ids, int B, int T,
                                  const long long* edges, int n_edges, int* bins, int* hours, int* days,
                                  void* stream) {
  if (!ts || !ids || !edges || !bins || !hours || !days || B <= 0 || T <= 0 || n_edges < 0)
    return set_error(CAST_ERR_BAD_ARG, "time_features");
  const long n = (long)B * T;
  CAST_LAUNCH(time_features_kernel, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, ts, ref, ids, B,
              T, edges, n_edges, bins, hours, days);
  return check_launch("time_features");
}
