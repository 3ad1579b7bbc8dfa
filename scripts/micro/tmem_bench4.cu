// Microbenchmark (developer tool, round 2): tcgen05.ld throughput per SM WITHOUT local-memory contamination.
// tmem_bench*.cu of round 1 consumed the loaded registers with a dynamically indexed read (r[it & 31]); ptxas then keeps
// the array in local memory and every iteration also stores 4 KB per warp to L1 (8 STL.128 per LDTM in the SASS) — the
// 47 B/clk/SM those tools reported is not a clean TMEM number.  Here every loaded register is consumed by a statically
// indexed XOR chain (cuobjdump -sass: no STL / LDL in any kernel).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

__device__ __forceinline__ void ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, uint32_t* r) {  // 16 lanes x 256 bit x 4 = 16 registers
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}

template <int N>
__device__ __forceinline__ uint32_t xor_all(const uint32_t (&r)[N]) {
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) a ^= r[i];
  return a;
}

// MODE 0: ld x32 + wait, per iteration            (4 KB / warp-iteration)
// MODE 1: 2 x ld x32 in flight + one wait          (8 KB)
// MODE 2: 4 x ld x32 in flight + one wait          (16 KB)
// MODE 3: ld x16 + wait                            (2 KB)
// MODE 4: 4 x ld x8 + one wait                     (4 KB)
// MODE 5: 16x256b.x4 + wait                        (2 KB: 16 lanes x 32 B x 4)
// MODE 6: ld x32, NO wait inside the loop (wait every 8th iteration), registers overwritten: pure issue rate
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((warp % 4) * 32u) << 16) + ((warp / 4) % 4) * 128;
  uint32_t a[32], b[32], c[32], d[32], acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { tmem_ld_32x32b_x32(base, a); tmem_wait_ld_dep(a); acc ^= xor_all(a); }
    if (MODE == 1) { tmem_ld_32x32b_x32(base, a); tmem_ld_32x32b_x32(base + 32, b); tmem_wait_ld_dep(a); reg_fence(b); acc ^= xor_all(a) ^ xor_all(b); }
    if (MODE == 2) {
      tmem_ld_32x32b_x32(base, a); tmem_ld_32x32b_x32(base + 32, b); tmem_ld_32x32b_x32(base + 64, c); tmem_ld_32x32b_x32(base + 96, d);
      tmem_wait_ld_dep(a); reg_fence(b); reg_fence(c); reg_fence(d);
      acc ^= xor_all(a) ^ xor_all(b) ^ xor_all(c) ^ xor_all(d);
    }
    if (MODE == 3) { uint32_t h[16]; tmem_ld_32x32b_x16(base, h); tmem_wait_ld(); acc ^= xor_all(h); }
    if (MODE == 4) { ld_x8(base, a); ld_x8(base + 8, a + 8); ld_x8(base + 16, a + 16); ld_x8(base + 24, a + 24); tmem_wait_ld_dep(a); acc ^= xor_all(a); }
    if (MODE == 5) { uint32_t h[16]; ld_16x256b_x4(base, h); tmem_wait_ld(); acc ^= xor_all(h); }
    if (MODE == 6) { tmem_ld_32x32b_x32(base, a); if ((it & 7) == 7) { tmem_wait_ld_dep(a); acc ^= xor_all(a); } }
  }
  if (MODE == 6) { tmem_wait_ld_dep(a); acc ^= xor_all(a); }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
void run(int warps, const char* name, double bytes_per_warp_iter) {
  uint32_t* d; cudaMalloc(&d, 4096); long long* c; cudaMalloc(&c, 8);
  const int iters = 20000;
  k<MODE><<<148, warps * 32>>>(d, 100, c);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(d, iters, c);
  cudaDeviceSynchronize();
  long long hc = 0; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-34s warps/SM=%2d  %8.1f cycles/iter  %7.1f B/clk/SM %s\n", name, warps, (double)hc / iters,
         bytes_per_warp_iter * warps * iters / (double)hc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(c);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<0>(w, "32x32b.x32 + wait", 4096.0);
  for (int w : {4, 8, 16}) run<1>(w, "2 x 32x32b.x32 + wait", 8192.0);
  for (int w : {4, 8}) run<2>(w, "4 x 32x32b.x32 + wait", 16384.0);
  for (int w : {4, 8, 16}) run<3>(w, "32x32b.x16 + wait", 2048.0);
  for (int w : {4, 8}) run<4>(w, "4 x 32x32b.x8 + wait", 4096.0);
  for (int w : {4, 8, 16}) run<5>(w, "16x256b.x4 + wait", 2048.0);
  for (int w : {4, 8, 16}) run<6>(w, "32x32b.x32, wait every 8th", 4096.0);
  return 0;
}
