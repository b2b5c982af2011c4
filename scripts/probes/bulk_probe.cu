// Probe: cp.async.bulk global->shared throughput per SM (one issuing thread, `depth` copies of `bytes` in flight),
// all CTAs reading the SAME source region or disjoint regions.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../context-aware-sequential-recommendation_b200/csrc/umma.cuh"
using namespace cast;

__global__ void probe(const unsigned char* src, size_t cta_stride, int bytes, int depth, int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[8];
  const int t = threadIdx.x;
  if (t == 0) for (int i = 0; i < depth; ++i) umma::mbar_init(&bar[i], 1);
  __syncthreads();
  if (t == 0) {
    const unsigned char* s = src + (size_t)blockIdx.x * cta_stride;
    long long t0 = clock64();
    for (int i = 0; i < depth; ++i) {
      umma::mbar_arrive_expect_tx(&bar[i], bytes);
      umma::bulk_g2s(sm + (size_t)i * bytes, s + (size_t)i * bytes, bytes, &bar[i]);
    }
    for (int it = 0; it < iters; ++it) {
      const int i = it % depth;
      umma::mbar_wait(&bar[i], (it / depth) & 1);
      if (it + depth < iters) {
        umma::mbar_arrive_expect_tx(&bar[i], bytes);
        umma::bulk_g2s(sm + (size_t)i * bytes, s + (size_t)((it + depth) % 64) * bytes, bytes, &bar[i]);
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  unsigned char* src; size_t total = (size_t)148 * 64 * 65792 + 1024; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  const int iters = 256;
  for (int ctas : {1, 148}) for (int same : {1, 0}) for (int bytes : {16384, 32896, 65792}) for (int depth : {1, 2, 4}) {
    if ((size_t)bytes * depth > 199000) continue;
    probe<<<ctas, 32, 200000>>>(src, same ? 0 : (size_t)64 * 65792, bytes, depth, iters, d);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double us = h / 1965.0;  // approx at 1.965 GHz
    printf("ctas=%3d %s bytes=%6d depth=%d: %8.1f clk/copy  %6.1f GB/s per SM  (%s)\n", ctas, same ? "same-src" : "disjoint", bytes, depth,
           (double)h / iters, (double)bytes * iters / (us * 1e3), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
