// Microbenchmark (developer tool), follow-up of umma_bench.cu: is the ~128-cycle cost of a tcgen05.mma (M = 128,
// K = 16, N = 64) a limit per CTA, per issuing thread, or per SM?  NTHR issuing threads (different warps) per CTA, each
// with its own accumulator and mbarrier; 128 TMEM columns per CTA so that up to 4 CTAs fit on an SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

template <int NTHR, bool TS>
__global__ void __launch_bounds__(128) k(int iters, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x % 32 == 0 && warp >= 1 && warp <= NTHR) {
    const uint32_t w = warp - 1;
    constexpr uint32_t idesc = umma_idesc_f16(128, 32, true);
    const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem), 1024, 0);
    const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + 16384), 1024, 8192);
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (TS) umma_ts(tmem + 32 * w, tmem + 64 + 32 * w + 8 * kk, bdesc + 128 * kk, idesc, 1u);
        else umma_ss(tmem + 32 * w, adesc + 2 * kk, bdesc + 128 * kk, idesc, 1u);
      }
      if ((it + 1) % 16 == 0) {
        tc_commit(&bar[w]);
        mbar_wait(&bar[w], phase);
        phase ^= 1;
      }
    }
    tc_commit(&bar[w]);
    mbar_wait(&bar[w], phase);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && w == 0) cyc[0] = (t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

template <int NTHR, bool TS>
void run(int ctas_per_sm, const char* name) {
  long long* c; cudaMalloc(&c, 8);
  const int iters = 4000;
  const int smem = 16384 + 8192 + 1024;
  cudaFuncSetAttribute(k<NTHR, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<NTHR, TS><<<148 * ctas_per_sm, 128, smem>>>(10, c);
  cudaDeviceSynchronize();
  k<NTHR, TS><<<148 * ctas_per_sm, 128, smem>>>(iters, c);
  cudaError_t e = cudaDeviceSynchronize();
  long long hc = 0; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  const double per_mma = (double)hc / (iters * 4.0);
  printf("%-14s N=32 issuing threads/CTA=%d CTAs/SM=%d : %7.1f cycles per MMA per thread -> %6.0f flop/clk/SM (%s)\n",
         name, NTHR, ctas_per_sm, per_mma, 2.0 * 128 * 32 * 16 * ctas_per_sm * NTHR / per_mma, cudaGetErrorString(e));
  cudaFree(c);
}

int main() {
  for (int c : {1, 2, 3, 4}) run<1, false>(c, "SS");
  for (int c : {1, 2, 4}) run<2, false>(c, "SS");
  for (int c : {1, 2, 4}) run<1, true>(c, "TS");
  for (int c : {1, 2}) run<2, true>(c, "TS");
  return 0;
}
