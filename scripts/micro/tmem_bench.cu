// Microbenchmark (developer tool): tcgen05.ld / tcgen05.st throughput per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((warp % 4) * 32u) << 16) + (warp / 4) * 128;
  uint32_t r[32], q[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // ld x32, wait each
      tmem_ld_32x32b_x32(base + (it & 3) * 32, r); tmem_wait_ld_dep(r); acc += r[it & 31];
    } else if (MODE == 1) {   // two ld x32 in flight, one wait
      tmem_ld_32x32b_x32(base + (it & 1) * 64, r); tmem_ld_32x32b_x32(base + (it & 1) * 64 + 32, q);
      tmem_wait_ld_dep(r); tmem_wait_ld_dep(q); acc += r[it & 31] + q[it & 31];
    } else if (MODE == 2) {   // st x32
      r[0] = it; tmem_st_32x32b_x32(base + (it & 3) * 32, r); tmem_wait_st();
    } else {                  // st x16 (P store)
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = r[i] + it;
      tmem_st_32x32b_x16(base + (it & 7) * 16, h); tmem_wait_st();
    }
  }
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
void run(int warps, const char* name, double bytes_per_iter_per_warp) {
  float* d; cudaMalloc(&d, 4096);
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps * 32>>>(d, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  double bytes = bytes_per_iter_per_warp * warps * iters;  // per SM
  printf("%s warps/SM=%d  %.3f ms  %.1f B/clk/SM (at 1.965 GHz)  %s\n", name, warps, ms, bytes / (ms * 1e-3) / 1.965e9,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int w : {4, 8}) run<0>(w, "ld x32 wait-each ", 4096.0);
  for (int w : {4, 8}) run<1>(w, "ld 2x x32 in flight", 8192.0);
  for (int w : {4, 8}) run<2>(w, "st x32           ", 4096.0);
  for (int w : {4, 8}) run<3>(w, "st x16           ", 2048.0);
  return 0;
}
