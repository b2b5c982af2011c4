// Probe (hardware facts the row kernels of csrc/row_umma.cu rely on): tcgen05.mma kind::tf32 with MN-major operands
// (contraction over the ROWS of row-staged tiles: weight gradients), instruction shapes N = 112 / 176, M = 64 and where
// its rows land in tensor memory, cycles per MMA for those shapes, and the 1-D bulk store shared -> global.
// Host-driven: the host builds byte images of both operands, descriptors and the expected result for each hypothesis.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_mn_probe umma_mn_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../context-aware-sequential-recommendation_b200/csrc/umma.cuh"
using namespace cast;

struct Cfg {
  uint32_t a_off, b_off;          // byte offsets of the operand images in dynamic shared memory
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_step, b_step;        // descriptor start-address advance per k-step (bytes)
  uint32_t idesc;
  int ksteps, ncols, reps;
  uint32_t img_bytes;
};

__global__ void probe(const unsigned char* img, Cfg c, float* out, long long* clk, int* status) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (uint32_t i = t; i < c.img_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm)[i] = reinterpret_cast<const uint32_t*>(img)[i];
  if (warp == 0) umma::tmem_alloc(&slot, 512);
  if (t == 0) umma::mbar_init(&bar, 1);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot;
  uint32_t par = 0;
  bool ok = true;
  if (t == 0) {
    const uint32_t a = umma::smem_u32(sm) + c.a_off, b = umma::smem_u32(sm) + c.b_off;
    for (int k = 0; k < c.ksteps; ++k)
      umma::mma_tf32(tmem, umma::smem_desc(a + k * c.a_step, c.a_lbo, c.a_sbo),
                     umma::smem_desc(b + k * c.b_step, c.b_lbo, c.b_sbo), c.idesc, k > 0);
    umma::mma_commit(&bar);
    ok = umma::mbar_wait(&bar, par);
    par ^= 1;
    if (!ok) *status = 1;
  }
  __syncthreads();
  umma::fence_after_sync();
  if (*status == 0) {
    for (int cb = 0; cb < c.ncols; cb += 32) {
      float v[32];
      umma::tmem_ld32(tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)cb, v);
      for (int j = 0; j < 32; ++j) out[(size_t)t * c.ncols + cb + j] = v[j];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t == 0 && c.reps > 0 && *status == 0) {
    umma::fence_after_sync();
    const uint32_t a = umma::smem_u32(sm) + c.a_off, b = umma::smem_u32(sm) + c.b_off;
    const long long t0 = clock64();
    for (int r = 0; r < c.reps; ++r)
      for (int k = 0; k < c.ksteps; ++k)
        umma::mma_tf32(tmem + 256, umma::smem_desc(a + k * c.a_step, c.a_lbo, c.a_sbo),
                       umma::smem_desc(b + k * c.b_step, c.b_lbo, c.b_sbo), c.idesc, 1);
    const long long t1 = clock64();
    umma::mma_commit(&bar);
    if (!umma::mbar_wait(&bar, par)) *status = 2;
    const long long t2 = clock64();
    clk[0] = t1 - t0;
    clk[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 512);
}

// ---- bulk store probe: thread-per-row float2 stores into a dense [128][50] tile, then ONE cp.async.bulk s2g
__global__ void bulk_store_probe(float* dst, int* status) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* tile = reinterpret_cast<float*>(sm);
  const int t = threadIdx.x;
  for (int c = 0; c < 50; c += 2) *reinterpret_cast<float2*>(tile + t * 50 + c) = make_float2(t * 100.f + c, t * 100.f + c + 1);
  umma::fence_smem_to_async();
  __syncthreads();
  if (t == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(umma::smem_u32(tile)),
                 "r"(128 * 50 * 4)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    *status = 0;
  }
}


// ---- chain timing: 7 k-steps x 3 passes (lo*hi, hi*lo, hi*hi) as the row kernels issue them, N = 64 or 128,
// accumulating into ONE tensor-memory tile (dependent chain) or into THREE tiles round-robin (one per pass)
__global__ void chain_probe(int N, int nacc, int reps, long long* clk, int* status) {
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
    const uint32_t pa = 2064, pb = N * 16 + 16;
    const uint32_t a_hi = umma::smem_u32(sm), a_lo = a_hi + 14 * pa, b_hi = a_lo + 14 * pa, b_lo = b_hi + 14 * pb;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint64_t dah = umma::smem_desc(a_hi, pa, 128), dal = umma::smem_desc(a_lo, pa, 128);
      uint64_t dbh = umma::smem_desc(b_hi, pb, 128), dbl = umma::smem_desc(b_lo, pb, 128);
      for (int ks = 0; ks < 7; ++ks) {
        umma::mma_tf32(tmem, dal, dbh, idesc, 1u);
        umma::mma_tf32(tmem + (nacc == 3 ? N : 0), dah, dbl, idesc, 1u);
        umma::mma_tf32(tmem + (nacc == 3 ? 2 * N : 0), dah, dbh, idesc, 1u);
        dah = umma::desc_advance(dah, 2 * pa); dal = umma::desc_advance(dal, 2 * pa);
        dbh = umma::desc_advance(dbh, 2 * pb); dbl = umma::desc_advance(dbl, 2 * pb);
      }
    }
    const long long t1 = clock64();
    umma::mma_commit(&bar);
    if (!umma::mbar_wait(&bar, 0)) *status = 3;
    const long long t2 = clock64();
    clk[0] = t1 - t0; clk[1] = t2 - t0;
  }
  __syncthreads();
  if (t < 32) umma::tmem_free(tmem, 512);
}

static uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Run {
  std::vector<float> out;
  long long clk[2];
  int status;
  cudaError_t err;
};

static Run run(const std::vector<unsigned char>& img, Cfg c) {
  static unsigned char* dimg = nullptr;
  static float* dout = nullptr;
  static long long* dclk = nullptr;
  static int* dst = nullptr;
  if (!dimg) {
    cudaMalloc(&dimg, 220000); cudaMalloc(&dout, 128 * 512 * 4); cudaMalloc(&dclk, 16); cudaMalloc(&dst, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
  }
  c.img_bytes = (uint32_t)((img.size() + 3) & ~3u);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 128 * 512 * 4);
  cudaMemset(dst, 0, 4);
  cudaMemset(dclk, 0, 16);
  probe<<<1, 128, 220000>>>(dimg, c, dout, dclk, dst);
  Run r;
  r.err = cudaDeviceSynchronize();
  r.out.resize((size_t)128 * c.ncols);
  cudaMemcpy(r.out.data(), dout, r.out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(r.clk, dclk, 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(&r.status, dst, 4, cudaMemcpyDeviceToHost);
  return r;
}

static float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 13) - 6); }
static float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); }

// K-major image: element (row, k) at (k/4)*pitch + row*16 + (k%4)*4
static void put_kmajor(std::vector<unsigned char>& img, uint32_t off, int rows, int K, uint32_t pitch, float (*f)(int, int)) {
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      float v = f(r, k);
      memcpy(&img[off + (k / 4) * pitch + r * 16 + (k % 4) * 4], &v, 4);
    }
}
// "row tile" image of the transposed operand: logical (mn, k) stored as row k, feature mn of a row-staged tile:
// (mn/4)*pitch + k*16 + (mn%4)*4   (pitch = Krows*16 + 16)
static void put_rowtile(std::vector<unsigned char>& img, uint32_t off, int MN, int K, uint32_t pitch, float (*f)(int, int)) {
  for (int m = 0; m < MN; ++m)
    for (int k = 0; k < K; ++k) {
      float v = f(m, k);
      memcpy(&img[off + (m / 4) * pitch + k * 16 + (m % 4) * 4], &v, 4);
    }
}

static void check(const char* name, const Run& r, int Mv, int Nv, int ncols, int K, const int* lane_of_row = nullptr) {
  if (r.err != cudaSuccess || r.status) {
    printf("%-46s : FAILED (cuda: %s, status %d)\n", name, cudaGetErrorString(r.err), r.status);
    return;
  }
  int bad = 0, first_m = -1, first_n = -1;
  float got0 = 0, exp0 = 0;
  for (int m = 0; m < Mv; ++m)
    for (int n = 0; n < Nv; ++n) {
      float e = 0;
      for (int k = 0; k < K; ++k) e += aval(m, k) * bval(n, k);
      const int lane = lane_of_row ? lane_of_row[m] : m;
      const float g = r.out[(size_t)lane * ncols + n];
      if (!(g == e)) {
        if (!bad) { first_m = m; first_n = n; got0 = g; exp0 = e; }
        ++bad;
      }
    }
  if (bad) printf("%-46s : MISMATCH %d of %d (first at m=%d n=%d: got %g expected %g)\n", name, bad, Mv * Nv, first_m, first_n, got0, exp0);
  else printf("%-46s : MATCH (%d x %d, K=%d)\n", name, Mv, Nv, K);
}

int main() {
  // ---------------- test 1: K-major baseline, NaN in unused A rows / B rows must not leak
  {
    const int K = 16; const uint32_t pa = 128 * 16 + 16, pb = 64 * 16 + 16;
    std::vector<unsigned char> img(100000, 0);
    Cfg c{}; c.a_off = 0; c.b_off = 40000;
    put_kmajor(img, c.a_off, 128, K, pa, aval);
    put_kmajor(img, c.b_off, 64, K, pb, bval);
    const uint32_t nanbits = 0x7fc00000u;
    for (int r = 100; r < 128; ++r) for (int k = 0; k < K; ++k) memcpy(&img[c.a_off + (k / 4) * pa + r * 16 + (k % 4) * 4], &nanbits, 4);
    for (int r = 56; r < 64; ++r) for (int k = 0; k < K; ++k) memcpy(&img[c.b_off + (k / 4) * pb + r * 16 + (k % 4) * 4], &nanbits, 4);
    c.a_lbo = pa; c.a_sbo = 128; c.b_lbo = pb; c.b_sbo = 128; c.a_step = 2 * pa; c.b_step = 2 * pb;
    c.idesc = idesc(128, 64, 0, 0); c.ksteps = K / 8; c.ncols = 64; c.reps = 64;
    Run r = run(img, c);
    check("K-major A,B (NaN in A rows>=100, B rows>=56)", r, 100, 56, 64, K);
    printf("    timing M=128 N=64 : issue %.1f clk/mma, done %.1f clk/mma\n", r.clk[0] / (64.0 * c.ksteps), r.clk[1] / (64.0 * c.ksteps));
  }
  // ---------------- tests 2-4: MN-major operands from row-staged tiles (K = tile rows), both descriptor hypotheses
  for (int hyp = 0; hyp < 2; ++hyp)
    for (int which = 1; which <= 3; ++which) {  // 1: A MN-major, 2: B MN-major, 3: both
      const int K = 32;  // tile rows
      const uint32_t prow = K * 16 + 16;        // pitch of a row-staged tile with K rows
      const uint32_t pa = 128 * 16 + 16, pb = 64 * 16 + 16;
      std::vector<unsigned char> img(200000, 0);
      Cfg c{}; c.a_off = 0; c.b_off = 100000;
      const bool amn = which & 1, bmn = which & 2;
      if (amn) put_rowtile(img, c.a_off, 128, K, prow, aval); else put_kmajor(img, c.a_off, 128, K, pa, aval);
      if (bmn) put_rowtile(img, c.b_off, 64, K, prow, bval); else put_kmajor(img, c.b_off, 64, K, pb, bval);
      const uint32_t mn_lbo = hyp == 0 ? 128 : prow, mn_sbo = hyp == 0 ? prow : 128;
      c.a_lbo = amn ? mn_lbo : pa; c.a_sbo = amn ? mn_sbo : 128; c.a_step = amn ? 128 : 2 * pa;
      c.b_lbo = bmn ? mn_lbo : pb; c.b_sbo = bmn ? mn_sbo : 128; c.b_step = bmn ? 128 : 2 * pb;
      c.idesc = idesc(128, 64, amn, bmn); c.ksteps = K / 8; c.ncols = 64; c.reps = 0;
      char name[128];
      snprintf(name, sizeof name, "MN-major %s, hypothesis %s", which == 1 ? "A" : which == 2 ? "B" : "A and B",
               hyp == 0 ? "LBO=128 (K groups), SBO=pitch (MN groups)" : "LBO=pitch, SBO=128");
      Run r = run(img, c);
      check(name, r, 128, 64, 64, K);
    }
  // ---------------- test 5: N = 112 and N = 176 (both MN-major, K = 128 rows, 14-slab operands back to back), timing
  for (int hyp = 0; hyp < 2; ++hyp)
  for (int N : {112, 176, 64, 128}) {
    const int K = 128; const uint32_t prow = K * 16 + 16;
    std::vector<unsigned char> img(220000, 0);
    Cfg c{}; c.a_off = 0; c.b_off = 32 * prow;  // A: 32 slabs (M = 128), B: N/4 slabs
    if (c.b_off + (N / 4) * prow > 220000) { printf("N=%d: image too large, skipped\n", N); continue; }
    put_rowtile(img, c.a_off, 128, K, prow, aval);
    put_rowtile(img, c.b_off, N, K, prow, bval);
    c.a_lbo = hyp ? prow : 128; c.a_sbo = hyp ? 128 : prow; c.a_step = 128; c.b_lbo = c.a_lbo; c.b_sbo = c.a_sbo; c.b_step = 128;
    c.idesc = idesc(128, N, 1, 1); c.ksteps = K / 8; c.ncols = (N + 31) / 32 * 32; c.reps = 8;
    char name[128];
    snprintf(name, sizeof name, "MN-major A and B, M=128 N=%d K=128 (hyp %d)", N, hyp);
    Run r = run(img, c);
    check(name, r, 128, N, c.ncols, K);
    printf("    timing M=128 N=%d MN-major: issue %.1f clk/mma, done %.1f clk/mma\n", N, r.clk[0] / (8.0 * c.ksteps), r.clk[1] / (8.0 * c.ksteps));
  }
  // ---------------- test 6: M = 64 (K-major): where do the rows land, and how fast is it
  {
    const int K = 16; const uint32_t pa = 64 * 16 + 16, pb = 64 * 16 + 16;
    std::vector<unsigned char> img(100000, 0);
    Cfg c{}; c.a_off = 0; c.b_off = 40000;
    put_kmajor(img, c.a_off, 64, K, pa, aval);
    put_kmajor(img, c.b_off, 64, K, pb, bval);
    c.a_lbo = pa; c.a_sbo = 128; c.b_lbo = pb; c.b_sbo = 128; c.a_step = 2 * pa; c.b_step = 2 * pb;
    c.idesc = idesc(64, 64, 0, 0); c.ksteps = K / 8; c.ncols = 64; c.reps = 64;
    Run r = run(img, c);
    if (r.err != cudaSuccess || r.status) printf("M=64: FAILED (%s, status %d)\n", cudaGetErrorString(r.err), r.status);
    else {
      int lane_of_row[64];
      bool all = true;
      for (int m = 0; m < 64; ++m) {
        lane_of_row[m] = -1;
        for (int lane = 0; lane < 128 && lane_of_row[m] < 0; ++lane) {
          bool eq = true;
          for (int n = 0; n < 64 && eq; ++n) {
            float e = 0;
            for (int k = 0; k < K; ++k) e += aval(m, k) * bval(n, k);
            eq = r.out[(size_t)lane * 64 + n] == e;
          }
          if (eq) lane_of_row[m] = lane;
        }
        all = all && lane_of_row[m] >= 0;
      }
      printf("M=64 N=64 K-major: rows found %s; row->lane:", all ? "ALL" : "NOT ALL");
      for (int m = 0; m < 64; m += 1) printf(" %d", lane_of_row[m]);
      printf("\n    timing M=64 N=64 : issue %.1f clk/mma, done %.1f clk/mma\n", r.clk[0] / (64.0 * c.ksteps), r.clk[1] / (64.0 * c.ksteps));
    }
  }
  // ---------------- test 7: bulk store shared -> global
  {
    float* d; int* st; cudaMalloc(&d, 128 * 50 * 4); cudaMalloc(&st, 4); cudaMemset(d, 0, 128 * 50 * 4); cudaMemset(st, 0xff, 4);
    cudaFuncSetAttribute(bulk_store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    bulk_store_probe<<<1, 128, 32768>>>(d, st);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * 50);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t) for (int c = 0; c < 50; ++c) bad += h[t * 50 + c] != t * 100.f + c;
    printf("bulk store shared->global (25600 B): %s, %d bad elements\n", cudaGetErrorString(e), bad);
  }
  // ---------------- test 8: realistic 21-MMA chains, one accumulator vs three
  {
    long long* dclk; int* dst; cudaMalloc(&dclk, 16); cudaMalloc(&dst, 4);
    cudaFuncSetAttribute(chain_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    for (int N : {64, 128})
      for (int nacc : {1, 3})
        for (int reps : {1, 8}) {
          cudaMemset(dst, 0, 4);
          chain_probe<<<1, 128, 200000>>>(N, nacc, reps, dclk, dst);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; int st; cudaMemcpy(h, dclk, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
          printf("chain N=%3d accumulators=%d reps=%d: issue %.1f clk/mma, done %.1f clk/mma (%lld total) %s status %d\n", N, nacc, reps,
                 h[0] / (21.0 * reps), h[1] / (21.0 * reps), h[1], cudaGetErrorString(e), st);
        }
  }
  return 0;
}
