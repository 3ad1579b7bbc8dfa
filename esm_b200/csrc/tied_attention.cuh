// esm_b200 — MSA tied row attention (sm_100a, head_dim 64).
//
// Replaces /root/reference/esm/axial_attention.py:71-111 (RowSelfAttention.compute_attention_weights /
// compute_attention_update): the attention logits are SUMMED over the R alignment rows,
//     S[h,b,i,j] = sum_r sum_e q[b,r,i,h,e] * k[b,r,j,h,e]                      (einsum "rinhd,rjnhd->hnij", :87)
//     P = softmax_j(S)  (padded key columns filled with -10000, :94-97)          (:105)
//     ctx[b,r,i,h,:] = sum_j P[h,b,i,j] * v[b,r,j,h,:]                           (einsum "hnij,rjnhd->rinhd", :108)
// Both contractions read q/k/v IN PLACE from the fused projection output qkv [B*R*C, 3E] (fp16) through 2-D TMA boxes:
// for the logits the K loop simply walks over the rows r (row offset r*C, 64 K-elements per step), for the update the
// V tile of row r is the MN-major B operand, exactly like V in the flash-attention kernels.  No regrouping copies.
//
//   tied_scores_kernel : one CTA per (b, h, 128 query columns, 256 key columns); K = R*64; fp32 logits -> S [H,B,C,C]
//   tied_softmax_kernel: one warp per logits row; fp32 softmax; writes fp16 P [H*B*C, Cp] (Cp = C rounded up to 64,
//                        zero filled) and, on request, the fp32 probabilities in place of the logits
//   tied_pv_kernel     : one CTA per (b, h, 128 query columns, 4 alignment rows): D[128, 4 x 64] += P_tile V_r tile
//                        over the key columns; fp16 context -> ctx [B*R*C, E]
// Roles inside the MMA kernels (192 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 allocates
// TMEM, warps 2-5 = epilogue (one TMEM lane = one output row per thread).
#pragma once

#include "common.cuh"

namespace esmb200 {

struct TiedParams {
  int B, R, C, H, E;   // E = 64 * H
  int Cp;              // C rounded up to 64: row pitch of P
  float* S;            // [H, B, C, C] fp32 logits (tied_scores) / probabilities (tied_softmax, optional)
  __half* P;           // [H*B*C, Cp] fp16 probabilities
  __half* ctx;         // [B*R*C, E]
  const uint8_t* key_pad;  // optional: key_pad[b * key_pad_stride + c] = 1 <=> key column c of alignment b is padding
                           // (filled with -10000 before the softmax)
  long long key_pad_stride;
  int write_probs;     // tied_softmax: also write the fp32 probabilities over S
};

namespace tied_cfg {
constexpr int NUM_THREADS = 192;
// scores
constexpr int S_BM = 128, S_BN = 256, S_STAGES = 4;
constexpr int S_A_BYTES = S_BM * 128, S_B_BYTES = S_BN * 128;
constexpr int S_STAGE_BYTES = S_A_BYTES + S_B_BYTES;                 // 48 KB
constexpr int S_SMEM_BYTES = S_STAGES * S_STAGE_BYTES + 1024 + 256;
// update
constexpr int V_BM = 128, V_ROWS = 4, V_STAGES = 2;  // 2 stages = 96 KB: two CTAs per SM, one's epilogue under the other's MMAs
constexpr int V_P_BYTES = V_BM * 128, V_V_BYTES = 64 * 128;          // 16 KB + 4 x 8 KB
constexpr int V_STAGE_BYTES = V_P_BYTES + V_ROWS * V_V_BYTES;        // 48 KB
constexpr int V_SMEM_BYTES = V_STAGES * V_STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 256;
}  // namespace tied_cfg

// ---------------------------------------------------------------------------------------------------------------
// S = sum_r Q_r K_r^T
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(tied_cfg::NUM_THREADS, 1)
tied_scores_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const TiedParams p) {
  using namespace tied_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_STAGES * S_STAGE_BYTES);
  uint64_t* full = bars;                 // [S_STAGES]
  uint64_t* empty = bars + S_STAGES;     // [S_STAGES]
  uint64_t* done = bars + 2 * S_STAGES;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S_STAGES + 1);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;  // warp: uniform for ptxas
  const int m0 = blockIdx.x * S_BM, n0 = blockIdx.y * S_BN;
  const int b = blockIdx.z / p.H, h = blockIdx.z % p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_d = *tmem_slot;

  // Both control warps run their loops warp-convergent with operands derived from warp-uniform values; only the TMA /
  // tcgen05 instructions sit under elect_one() (no per-instruction ELECT / R2UR / BRA.U.ANY waterfall: the R-deep loop of
  // this kernel was issue-bound, 5 x ~94 cycles per alignment row against 256 cycles of tensor work).
  const uint32_t u_smem = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
  const uint32_t u_bars = u_smem + S_STAGES * S_STAGE_BYTES;
  if (warp == 0) {
    for (int r = 0; r < p.R; ++r) {
      const uint32_t s = r % S_STAGES;
      mbar_wait_relaxed(&empty[s], ((r / S_STAGES) & 1) ^ 1);
      const int row = (b * p.R + r) * p.C;
      const uint32_t st = u_smem + s * S_STAGE_BYTES, fb = u_bars + s * 8;
      if (elect_one()) {
        mbar_arrive_expect_tx_addr(fb, S_STAGE_BYTES);
        tma_load_2d_addr(st, &tmap_q, fb, h * 64, row + m0);
        tma_load_2d_addr(st + S_A_BYTES, &tmap_k, fb, p.E + h * 64, row + n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(S_BM, S_BN, false);
    const uint32_t u_tmem = __shfl_sync(0xffffffffu, tmem_d, 0);
    for (int r = 0; r < p.R; ++r) {
      const uint32_t s = r % S_STAGES;
      mbar_wait(&full[s], (r / S_STAGES) & 1);
      tc_fence_after();
      const uint64_t adesc = umma_smem_desc_sw128(u_smem + s * S_STAGE_BYTES, 1024, 0);
      const uint64_t bdesc = umma_smem_desc_sw128(u_smem + s * S_STAGE_BYTES + S_A_BYTES, 1024, 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(u_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (r | k) != 0 ? 1u : 0u);
        tc_commit_addr(u_bars + (S_STAGES + s) * 8);
        if (r + 1 == p.R) tc_commit_addr(u_bars + 2 * S_STAGES * 8);
      }
      __syncwarp();
    }
  } else {
    const uint32_t quarter = warp % 4;
    const int ci = m0 + quarter * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    float* dst = p.S + ((size_t)(h * p.B + b) * p.C + ci) * p.C;
    const bool vec_ok = (p.C % 4) == 0;
#pragma unroll 1
    for (int c = 0; c < S_BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_d + ((quarter * 32u) << 16) + c * 32, v);
      tmem_wait_ld_dep(v);
      const int cj0 = n0 + c * 32;
      if (ci < p.C && cj0 < p.C) {
        if (vec_ok && cj0 + 32 <= p.C) {
          float4* d4 = reinterpret_cast<float4*>(dst + cj0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                __uint_as_float(v[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cj0 + i < p.C) dst[cj0 + i] = __uint_as_float(v[i]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// row softmax over the key columns: fp32 in, fp16 P out (+ fp32 probabilities in place when asked)
// ---------------------------------------------------------------------------------------------------------------
constexpr int TIED_MAX_C = 1024;  // MSA Transformer max_positions (msa_transformer.py:57-60)

__global__ void __launch_bounds__(256)
tied_softmax_kernel(const TiedParams p) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long row = (long long)blockIdx.x * 8 + warp;  // (h*B + b)*C + ci
  const long long rows = (long long)p.H * p.B * p.C;
  if (row >= rows) return;
  pdl_launch_dependents();
  pdl_wait();
  const int b = (int)((row / p.C) % p.B);
  float* s = p.S + row * p.C;
  const uint8_t* pad = p.key_pad ? p.key_pad + (size_t)b * p.key_pad_stride : nullptr;
  float v[TIED_MAX_C / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < TIED_MAX_C / 32; ++i) {
    const int c = i * 32 + lane;
    float x = -INFINITY;
    if (c < p.C) {
      x = s[c];
      if (pad && pad[c]) x = -10000.f;  // axial_attention.py:94-97
    }
    v[i] = x;
    mx = fmaxf(mx, x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < TIED_MAX_C / 32; ++i) {
    v[i] = (i * 32 + lane < p.C) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __half* pr = p.P + row * p.Cp;
#pragma unroll
  for (int i = 0; i < TIED_MAX_C / 32; ++i) {
    const int c = i * 32 + lane;
    if (c < p.Cp) {
      const float q = v[i] * inv;
      pr[c] = __float2half_rn(q);
      if (p.write_probs && c < p.C) s[c] = q;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ctx_r = P V_r for 4 alignment rows r per CTA
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(tied_cfg::NUM_THREADS, 2)
tied_pv_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_v,
               const TiedParams p) {
  using namespace tied_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + V_STAGES * V_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + V_STAGES;
  uint64_t* done = bars + 2 * V_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * V_STAGES + 1);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;  // warp: uniform for ptxas
  const int m0 = blockIdx.x * V_BM;
  const int r0 = blockIdx.y * V_ROWS;
  const int nr = min(V_ROWS, p.R - r0);
  const int b = blockIdx.z / p.H, h = blockIdx.z % p.H;
  const int nk = p.Cp / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_p);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < V_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_d = *tmem_slot;

  // warp-convergent control warps, see tied_scores_kernel (16 MMAs per 64-key slab were 16 waterfalls here)
  const uint32_t u_smem = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
  const uint32_t u_bars = u_smem + (uint32_t)(reinterpret_cast<uint8_t*>(bars) - smem);
  if (warp == 0) {
    for (int j = 0; j < nk; ++j) {
      const uint32_t s = j % V_STAGES;
      mbar_wait_relaxed(&empty[s], ((j / V_STAGES) & 1) ^ 1);
      const uint32_t st = u_smem + s * V_STAGE_BYTES, fb = u_bars + s * 8;
      if (elect_one()) {
        mbar_arrive_expect_tx_addr(fb, V_P_BYTES + nr * V_V_BYTES);
        tma_load_2d_addr(st, &tmap_p, fb, j * 64, (h * p.B + b) * p.C + m0);
        for (int i = 0; i < nr; ++i)
          tma_load_2d_addr(st + V_P_BYTES + i * V_V_BYTES, &tmap_v, fb, 2 * p.E + h * 64,
                           (b * p.R + r0 + i) * p.C + j * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(V_BM, 64, true);
    const uint32_t u_tmem = __shfl_sync(0xffffffffu, tmem_d, 0);
    for (int j = 0; j < nk; ++j) {
      const uint32_t s = j % V_STAGES;
      mbar_wait(&full[s], (j / V_STAGES) & 1);
      tc_fence_after();
      const uint32_t base = u_smem + s * V_STAGE_BYTES;
      const uint64_t pdesc = umma_smem_desc_sw128(base, 1024, 0);
      if (elect_one()) {
        for (int i = 0; i < nr; ++i) {
          const uint64_t vdesc = umma_smem_desc_sw128(base + V_P_BYTES + i * V_V_BYTES, 1024, 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(u_tmem + 64 * i, pdesc + 2 * k, vdesc + 128 * k, idesc, (j | k) != 0 ? 1u : 0u);
        }
        tc_commit_addr(u_bars + (V_STAGES + s) * 8);
        if (j + 1 == nk) tc_commit_addr(u_bars + 2 * V_STAGES * 8);
      }
      __syncwarp();
    }
  } else {
    const uint32_t quarter = warp % 4;
    const int ci = m0 + quarter * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int i = 0; i < nr; ++i) {
      uint32_t out[32];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_d + ((quarter * 32u) << 16) + 64 * i + 32 * hlf, v);
        tmem_wait_ld_dep(v);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          out[hlf * 16 + e] = pack_half2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
      }
      if (ci < p.C) {
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + ((size_t)(b * p.R + r0 + i) * p.C + ci) * p.E + h * 64);
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = make_uint4(out[4 * e], out[4 * e + 1], out[4 * e + 2], out[4 * e + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, TMEM_COLS);
  }
}

inline cudaError_t launch_tied_scores(const CUtensorMap& tq, const CUtensorMap& tk, const TiedParams& p,
                                      cudaStream_t st) {
  using namespace tied_cfg;
  cudaError_t e = cudaFuncSetAttribute(tied_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  dim3 grid((p.C + S_BM - 1) / S_BM, (p.C + S_BN - 1) / S_BN, p.B * p.H);
  return launch_pdl(tied_scores_kernel, grid, dim3(NUM_THREADS), S_SMEM_BYTES, st, tq, tk, p);
}

inline cudaError_t launch_tied_softmax(const TiedParams& p, cudaStream_t st) {
  const long long rows = (long long)p.H * p.B * p.C;
  return launch_pdl(tied_softmax_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, p);
}

inline cudaError_t launch_tied_pv(const CUtensorMap& tp, const CUtensorMap& tv, const TiedParams& p, cudaStream_t st) {
  using namespace tied_cfg;
  cudaError_t e = cudaFuncSetAttribute(tied_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  dim3 grid((p.C + V_BM - 1) / V_BM, (p.R + V_ROWS - 1) / V_ROWS, p.B * p.H);
  return launch_pdl(tied_pv_kernel, grid, dim3(NUM_THREADS), V_SMEM_BYTES, st, tp, tv, p);
}

}  // namespace esmb200
