// Microbenchmark (developer tool): MUFU.EX2 throughput per SM as a function of resident warps per SMSP and of the
// surrounding instruction mix (the attention softmax inner loop: FFMA + EX2 + FADD + FMNMX + 0.5 F2FP per element).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float s0 = 0, s1 = 0, s2 = 0, s3 = 0, m0 = -1e30f, m1 = -1e30f;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (MODE == 0) {  // pure MUFU
        v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]);
      } else {          // softmax mix
        float a = fmaf(v[i], 1.4426950408889634f, -seed), b = fmaf(v[i + 1], 1.4426950408889634f, -seed);
        m0 = fmaxf(m0, fmaxf(v[i], v[i + 1]));
        float p0 = ex2(a), p1 = ex2(b);
        if (i & 2) s0 += p0 + p1; else s1 += p0 + p1;
        __half2 h = __floats2half2_rn(p0, p1);
        acc ^= *reinterpret_cast<unsigned*>(&h);
        v[i] = p0 * 0.5f - 1.0f; v[i + 1] = p1 * 0.5f - 1.0f;
      }
    }
  }
  float r = s0 + s1 + s2 + s3 + m0 + m1;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += v[i];
  if (r == 12345.678f) out[threadIdx.x] = r + acc;
}

template <int MODE>
void run(int warps_per_sm, const char* name) {
  float* d; cudaMalloc(&d, 4096);
  int iters = 2000;
  // one block per SM with warps_per_sm warps
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps_per_sm * 32>>>(d, 10, 0.3f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148, warps_per_sm * 32>>>(d, iters, 0.3f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ex = 148.0 * warps_per_sm * 32 * 32.0 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s warps/SM=%2d  %.3f ms  %.2f Gex2/s  -> %.2f ex2/clk/SM at %.0f MHz nominal\n", name, warps_per_sm, ms,
         ex / ms / 1e6, ex / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1e3);
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w, "pure-mufu ");
  for (int w : {4, 8, 16, 32}) run<1>(w, "softmaxmix");
  return 0;
}
