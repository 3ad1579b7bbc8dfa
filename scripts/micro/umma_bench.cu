// Microbenchmark (developer tool): cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of the
// operand form (A from shared memory "SS" / from tensor memory "TS") and of the number of CTAs sharing an SM.
// One thread per CTA issues `iters` batches of 4 MMAs (one 64-wide K tile) followed by a tcgen05.commit, and waits on
// the mbarrier every `depth` batches. Operand contents are irrelevant (shared memory is left uninitialised).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../esm_b200/csrc/common.cuh"
using namespace esmb200;

template <int N, bool TS, bool B_MN, int NACC = 1>
__global__ void __launch_bounds__(128) k(int iters, int depth, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x / 32;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, NACC > 2 ? 512 : 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, B_MN);
    const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem), 1024, 0);
    const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + 16384), 1024, B_MN ? 8192 : 0);
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t d = tmem + (kk % NACC) * 64;  // NACC independent accumulators (N = 64 only when NACC > 1)
        if (TS) umma_ts(d, tmem + (NACC > 2 ? 256 : 192) + 8 * kk, bdesc + (B_MN ? 128 : 2) * kk, idesc, 1u);
        else umma_ss(d, adesc + 2 * kk, bdesc + (B_MN ? 128 : 2) * kk, idesc, 1u);
      }
      if ((it + 1) % depth == 0) {
        tc_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, phase);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cyc[0] = (t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, NACC > 2 ? 512 : 256); }
}

template <int N, bool TS, bool B_MN, int NACC = 1>
void run(int ctas_per_sm, int depth, const char* name) {
  long long* c; cudaMalloc(&c, 8);
  const int iters = 4000;
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k<N, TS, B_MN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS, B_MN, NACC><<<148 * ctas_per_sm, 128, smem>>>(10, depth, c);
  cudaDeviceSynchronize();
  k<N, TS, B_MN, NACC><<<148 * ctas_per_sm, 128, smem>>>(iters, depth, c);
  cudaError_t e = cudaDeviceSynchronize();
  long long hc = 0; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  const double per_mma = (double)hc / (iters * 4.0);
  printf("%-26s N=%3d CTAs/SM=%d depth=%2d : %7.1f cycles per MMA per CTA -> %6.0f flop/clk/SM  (%s)\n", name, N,
         ctas_per_sm, depth, per_mma, 2.0 * 128 * N * 16 * ctas_per_sm / per_mma, cudaGetErrorString(e));
  cudaFree(c);
}

int main() {
  for (int c : {1, 2}) {
    run<64, false, false>(c, 16, "SS, B K-major");
    run<128, false, false>(c, 16, "SS, B K-major");
    run<256, false, false>(c, 16, "SS, B K-major");
    run<64, false, true>(c, 16, "SS, B MN-major");
    run<64, true, true>(c, 16, "TS (A in TMEM), B MN-major");
    run<128, true, false>(c, 16, "TS (A in TMEM), B K-major");
  }
  run<64, false, false, 2>(1, 16, "SS, 2 accumulators");
  run<64, false, false, 4>(1, 16, "SS, 4 accumulators");
  run<64, true, true, 2>(1, 16, "TS, 2 accumulators");
  run<64, false, false, 2>(2, 16, "SS, 2 accumulators");
  run<64, true, true, 2>(2, 16, "TS, 2 accumulators");
  run<64, false, false>(2, 1, "SS, commit+wait every 4");
  run<64, true, true>(2, 1, "TS, commit+wait every 4");
  return 0;
}
