// Probe: cycles per tcgen05.mma (kind::tf32, M=128) issued back to back from one thread on static shared-memory
// operands, and the latency of one commit -> mbarrier round trip.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../context-aware-sequential-recommendation_b200/csrc/umma.cuh"
using namespace cast;

__global__ void probe(int n_mma, int N, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x;
  for (int i = t; i < 200000 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.001f * (i % 97);
  if (t < 32) umma::tmem_alloc(&slot, 512);
  if (t == 0) umma::mbar_init(&bar, 1);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot;
  if (t == 0) {
    const uint32_t idesc = umma::idesc_tf32(128, N);
    const uint32_t a = umma::smem_u32(sm), b = a + 40000;
    uint32_t par = 0;
    // round-trip latency of a single MMA + commit + wait
    long long t0 = clock64();
    umma::mma_tf32(tmem, umma::smem_desc(a, 2064, 128), umma::smem_desc(b, 4112, 128), idesc, 0);
    umma::mma_commit(&bar);
    umma::mbar_wait(&bar, par); par ^= 1;
    long long t1 = clock64();
    out[0] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i)
      umma::mma_tf32(tmem, umma::smem_desc(a + (i & 3) * 4128, 2064, 128), umma::smem_desc(b + (i & 3) * 8224, 4112, 128), idesc, 1);
    long long tiss = clock64();
    umma::mma_commit(&bar);
    umma::mbar_wait(&bar, par); par ^= 1;
    t1 = clock64();
    out[1] = tiss - t0;   // issue time
    out[2] = t1 - t0;     // until completion
  }
  __syncthreads();
  if (t < 32) umma::tmem_free(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  for (int N : {64, 128, 256}) for (int n : {1, 12, 96, 960}) {
    probe<<<1, 128, 200000>>>(n, N, d);
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("N=%3d n_mma=%4d: single round trip %lld clk | issue %lld clk (%.1f/mma) | done %lld clk (%.1f/mma) %s\n", N, n, h[0], h[1],
           (double)h[1] / n, h[2], (double)h[2] / n, cudaGetErrorString(e));
  }
  // all SMs at once (power / clock effects): 148 CTAs
  probe<<<148, 128, 200000>>>(960, 256, d);
  long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("148 CTAs N=256 n=960: done %lld clk (%.1f/mma)\n", h[2], (double)h[2] / 960);
  return 0;
}
