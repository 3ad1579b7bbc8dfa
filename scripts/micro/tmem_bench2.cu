// Microbenchmark (developer tool): does tcgen05.ld overlap with MUFU work of the same warp? which ld shape is fastest?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

__device__ __forceinline__ void ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}
// 16x256b.x8: 16 lanes x 256 bit x 8 repeats = 16 lanes x 64 columns; 32 regs per thread
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: ld x32 + wait only. 1: 32 MUFU only. 2: ld x32 issued, then 32 MUFU on other regs, then wait (overlap?)
// 3: ld 32x32b.x64 + wait. 4: ld 16x256b.x8 + wait (2 KB)  5: ld x32, NO dependency/wait per iter (wait every 8)
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((warp % 4) * 32u) << 16) + (warp / 4) * 128;
  uint32_t r[32], acc = 0; float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r[i] = i; v[i] = 0.3f + i * 0.001f; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { tmem_ld_32x32b_x32(base + (it & 3) * 32, r); tmem_wait_ld_dep(r); acc += r[it & 31]; }
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]) * 0.5f - 1.0f;
    }
    if (MODE == 2) {
      tmem_ld_32x32b_x32(base + (it & 3) * 32, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]) * 0.5f - 1.0f;
      tmem_wait_ld_dep(r); acc += r[it & 31];
    }
    if (MODE == 3) { uint32_t q[64]; ld_32x32b_x64(base + (it & 1) * 64, q); tmem_wait_ld(); acc += q[it & 63]; }
    if (MODE == 4) { ld_16x256b_x8(base + (it & 1) * 64, r); tmem_wait_ld_dep(r); acc += r[it & 31]; }
    if (MODE == 5) { tmem_ld_32x32b_x32(base + (it & 3) * 32, r); if ((it & 7) == 7) { tmem_wait_ld_dep(r); acc += r[it & 31]; } }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  if (acc == 0x12345678u || s == 1.2345f) out[threadIdx.x] = acc + s;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
void run(int warps, const char* name, double bytes) {
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
  printf("%-34s warps/SM=%d  %.3f ms  %.0f ns/iter  %.1f B/clk/SM(@1.965GHz) %s\n", name, warps, ms, ms * 1e6 / iters,
         bytes * warps * iters / (ms * 1e-3) / 1.965e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int w : {4, 8}) run<0>(w, "ld 32x32b.x32 + wait", 4096.0);
  for (int w : {4, 8}) run<1>(w, "32 MUFU only", 0.0);
  for (int w : {4, 8}) run<2>(w, "ld x32 || 32 MUFU, then wait", 4096.0);
  for (int w : {4, 8}) run<3>(w, "ld 32x32b.x64 + wait", 8192.0);
  for (int w : {4, 8}) run<4>(w, "ld 16x256b.x8 + wait", 4096.0);
  for (int w : {4, 8}) run<5>(w, "ld x32, wait every 8", 4096.0);
  return 0;
}
