// Microbenchmark (developer tool): latency of tcgen05.ld -> tcgen05.wait::ld as a function of how many loads are in flight.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

__device__ __forceinline__ void ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// MODE n (1..4): n loads x32 into distinct registers, then one wait.  MODE 5: 4 loads x8 + wait.  MODE 6: 1 load x32 + 4 dummy x1 loads + wait
// MODE 7: ld x32 + wait + clock-based measure of latency (prints cycles)
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((warp % 4) * 32u) << 16) + (warp / 4) * 128;
  uint32_t a[32], b[32], c[32], d[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = b[i] = c[i] = d[i] = i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 1 && MODE <= 4) {
      tmem_ld_32x32b_x32(base, a);
      if (MODE >= 2) tmem_ld_32x32b_x32(base + 32, b);
      if (MODE >= 3) tmem_ld_32x32b_x32(base + 64, c);
      if (MODE >= 4) tmem_ld_32x32b_x32(base + 96, d);
      tmem_wait_ld();
      acc += a[it & 31] + b[it & 31] + c[it & 31] + d[it & 31];
    }
    if (MODE == 5) { ld_x8(base, a); ld_x8(base + 8, a + 8); ld_x8(base + 16, a + 16); ld_x8(base + 24, a + 24); tmem_wait_ld(); acc += a[it & 31]; }
    if (MODE == 6) { tmem_ld_32x32b_x32(base, a); ld_x1(base + 40, b[0]); ld_x1(base + 41, b[1]); ld_x1(base + 42, b[2]); ld_x1(base + 43, b[3]); tmem_wait_ld(); acc += a[it & 31] + b[it & 3]; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / iters;
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
void run(int warps, const char* name) {
  float* d; cudaMalloc(&d, 4096); long long* c; cudaMalloc(&c, 8);
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps * 32>>>(d, 100, c);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148, warps * 32>>>(d, iters, c);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  printf("%-30s warps/SM=%d  %.0f ns/iter  %lld cycles/iter\n", name, warps, ms * 1e6 / iters, hc);
  cudaFree(d);
}

int main() {
  for (int w : {1, 4, 8}) run<1>(w, "1 x (ld x32) + wait");
  for (int w : {4, 8}) run<2>(w, "2 x (ld x32) + wait");
  for (int w : {4, 8}) run<3>(w, "3 x (ld x32) + wait");
  for (int w : {4, 8}) run<4>(w, "4 x (ld x32) + wait");
  for (int w : {4, 8}) run<5>(w, "4 x (ld x8) + wait");
  for (int w : {4, 8}) run<6>(w, "ld x32 + 4 x (ld x1) + wait");
  return 0;
}
