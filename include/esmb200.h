/* esmb200.h — C ABI of the B200-native ESM-2 transformer-layer forward path (libesmb200.so).
 *
 * The reference (facebookresearch/esm, fair-esm 2.0.1) is pure Python and has no FFI for this path; the seam this
 * library sits behind is the Python method
 *     esm.modules.TransformerLayer.forward(x, self_attn_mask, self_attn_padding_mask, need_head_weights)
 *                                                      /root/reference/esm/modules.py:120-142
 * called from ESM2.forward's layer loop                 /root/reference/esm/model/esm2.py:111-121
 * Each entry point below names the reference code it replaces. The reference-side binding (a ctypes stub) is shown in
 * INTEGRATION.md; esm_b200/_lib.py is the shipped copy of that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch), borrowed for the duration of the call;
 *     the library owns only its packed fp16 weight copies inside esmb200_layer
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise
 *   - return 0 on success, a negative ESMB200_E* code on failure; esmb200_last_error() gives the message of the last
 *     failure on the calling thread. A device out-of-memory message starts with "CUDA out of memory" so that
 *     scripts/fold.py:165-178's handler keeps working
 *   - activations: residual stream x is fp32 [B, T, E] row-major (batch-major, i.e. the reference's (T,B,E)
 *     transposed); MMA operands are fp16 with fp32 accumulation; LayerNorm / softmax / residual adds are fp32
 *   - head_dim <= 64 (every esm.pretrained.esm2_* model except the 15B one, whose heads are 128 wide); no CPU fallback
 */
#ifndef ESMB200_H_
#define ESMB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESMB200_OK 0
#define ESMB200_EINVAL -1   /* bad argument / unsupported shape */
#define ESMB200_ECUDA -2    /* CUDA runtime or driver error */
#define ESMB200_ENOMEM -3   /* device allocation failed ("CUDA out of memory ...") */
#define ESMB200_EWORKSPACE -4 /* workspace too small */

#define ESMB200_ABI_VERSION 2

typedef struct esmb200_layer esmb200_layer; /* opaque: packed weights + TMA descriptors of one TransformerLayer */

/* fp32 device pointers to one layer's parameters, nn.Linear layout weight[out,in] (state-dict names in comments,
 * prefix "layers.{i}."; /root/reference/esm/modules.py:99-118, multihead_attention.py:109-113) */
typedef struct esmb200_layer_weights {
  int32_t embed_dim;   /* E */
  int32_t num_heads;   /* H, E == head_dim * H */
  int32_t ffn_dim;     /* F = 4E */
  float ln_eps;        /* 1e-5 */
  const float* ln1_weight; /* self_attn_layer_norm.weight [E] */
  const float* ln1_bias;   /* self_attn_layer_norm.bias   [E] */
  const float* q_weight;   /* self_attn.q_proj.weight [E,E] */
  const float* q_bias;     /* self_attn.q_proj.bias   [E]   */
  const float* k_weight;   /* self_attn.k_proj.weight [E,E] */
  const float* k_bias;
  const float* v_weight;   /* self_attn.v_proj.weight [E,E] */
  const float* v_bias;
  const float* out_weight; /* self_attn.out_proj.weight [E,E] */
  const float* out_bias;
  const float* ln2_weight; /* final_layer_norm.weight [E] */
  const float* ln2_bias;
  const float* fc1_weight; /* fc1.weight [F,E]; NULL = attention-only layer (no ln2/fc1/fc2; esmb200_axial_stack_forward) */
  const float* fc1_bias;   /* fc1.bias   [F]   */
  const float* fc2_weight; /* fc2.weight [E,F] */
  const float* fc2_bias;   /* fc2.bias   [E]   */
  int32_t head_dim;        /* 0 = E / H. Even values <= 128. 16 / 24 / 32 (ESM-2 8M / 35M / 150M) run in zero-padded
                            * 64-wide head slots of the attention-side tensors; 64 = 650M / 3B / MSA Transformer;
                            * 65..128 (15B: 128) take two slots per head and 64-column rope tables */
  int32_t precision;       /* 0 = fp16 MMA operands (default). 1 = "fp32x3": every MMA operand (activations, weights,
                            * q, k, v, P) is an fp16 hi | lo pair and every product runs hi*hi + lo*hi + hi*lo into the
                            * fp32 accumulator (22 significand bits per operand) — fp32-grade results at ~3x the tensor
                            * work; needs E % 64 == 0 and head_dim <= 64; not available on the MSA axial path */
} esmb200_layer_weights;

int esmb200_abi_version(void);
const char* esmb200_last_error(void);

/* Packs one TransformerLayer's weights (fp32 -> fp16, [Wq;Wk;Wv] concatenated) on `stream`.
 * Replaces TransformerLayer.__init__/_init_submodules state, modules.py:87-118. */
int esmb200_layer_create(const esmb200_layer_weights* w, void* stream, esmb200_layer** out);
int esmb200_layer_destroy(esmb200_layer* layer);

/* Scratch bytes needed by esmb200_layer_forward / esmb200_stack_forward for a [B,T] batch. */
size_t esmb200_workspace_bytes(int32_t embed_dim, int32_t num_heads, int32_t ffn_dim, int32_t B, int32_t T,
                               int32_t precision);

/* One TransformerLayer.forward (modules.py:120-142), in place on x:
 *     x += out_proj(attention(rope(q_proj(LN1 x) * d^-1/2), rope(k_proj(LN1 x)), v_proj(LN1 x)));
 *     x += fc2(gelu(fc1(LN2 x)))
 *   x          fp32 [B,T,E], updated in place
 *   pad_mask   uint8/bool [B,T], nonzero = padding key (self_attn_padding_mask, esm2.py:82), or NULL
 *   rope_cos/sin fp32 [T,32] (head_dim <= 64) or [T,64] (head_dim <= 128): cos/sin(t * inv_freq[j]) for j < head_dim/2
 *              (rotary_embedding.py:47-61), built by the caller; columns >= head_dim/2 are ignored
 *   attn_probs fp32 [B,H,T,T] or NULL: softmax probabilities per head (need_head_weights=True,
 *              multihead_attention.py:397-400, batch-major i.e. already transposed as esm2.py:121 does) */
int esmb200_layer_forward(esmb200_layer* layer, float* x, const uint8_t* pad_mask, int32_t B, int32_t T,
                          const float* rope_cos, const float* rope_sin, float* attn_probs, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Optional contact-head accumulation fused into the need_head_weights pass of esmb200_stack_forward
 * (ContactPredictionHead.forward esm/modules.py:338-357, restated as in esmb200_contact_accumulate): every layer's
 * probabilities are folded into the accumulators while they are written, the stacked attention tensor is never read
 * back. S = hi - lo, nt = ceil(T / 128). Ignored (the separate probability kernel runs) for fp32x3 layers. */
typedef struct esmb200_contact_job {
  const float* weights; /* [n_layers, H] fp32: contact_head.regression.weight */
  const uint8_t* keep;  /* [B,T] 1 = not <eos>, or NULL */
  float* acc;           /* [B,S,S]  += sum_{l,h} w[l,h] A_{l,h}; zeroed by the caller */
  float* row_part;      /* [n_layers,B,H,4*nt,S] row sums of A_{l,h} per 32-key quarter tile (sum over 4*nt = rowsum) */
  float* col_part;      /* [n_layers,B,H,4*nt,S] column sums per 32-query quarter tile (sum over 4*nt = colsum) */
  int32_t lo, hi;       /* cropped positions [lo,hi): 1 .. T-1 for <cls> ... <eos> */
} esmb200_contact_job;

/* The layer loop of ESM2.forward (esm2.py:111-121): runs n_layers layers in place on x.
 *   repr_out[i]  NULL or fp32 [B,T,E]: copy of x after layer i (hidden_representations[i+1], esm2.py:117-118)
 *   attn_out[i]  NULL or fp32 [B,H,T,T]: attention probabilities of layer i (esm2.py:119-121); batch b starts at
 *                attn_out[i] + b * attn_batch_stride elements (0 = contiguous H*T*T), so the caller can point layer i
 *                into its slice of the stacked [B,L,H,T,T] result (esm2.py:134) and skip the torch.stack copy;
 *                attn_flags bit 0: write the rows of padded QUERY tokens as zeros (esm2.py:135-139; padded key
 *                columns are zero anyway), so the caller needs no masking pass over the stack
 * either array pointer itself may be NULL. */
int esmb200_stack_forward(esmb200_layer* const* layers, int32_t n_layers, float* x, const uint8_t* pad_mask,
                          int32_t B, int32_t T, const float* rope_cos, const float* rope_sin,
                          float* const* repr_out, float* const* attn_out, int64_t attn_batch_stride,
                          int32_t attn_flags, const esmb200_contact_job* contact /* nullable */, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Embedding prologue of ESM2.forward (esm2.py:84-95): gather from table [V,E], zero <mask> rows and rescale by
 * 0.88/(1 - n_mask/n_nonpad) when token_dropout, zero pad rows. tokens int64 [B,T] -> x fp32 [B,T,E]. */
int esmb200_embed_tokens(const int64_t* tokens, const float* table, float* x, int32_t B, int32_t T, int32_t E,
                         int32_t padding_idx, int32_t mask_idx, int32_t token_dropout, void* stream);

/* torch.nn.LayerNorm over the last dim (ESM1bLayerNorm, modules.py:68-81; emb_layer_norm_after, esm2.py:123):
 * fp32 [M,E] -> fp32 [M,E]. out may alias x. */
int esmb200_layernorm(const float* x, const float* weight, const float* bias, float* out, int32_t M, int32_t E,
                      float eps, void* stream);

/* Per-sequence mean representation, scripts/extract.py:116-119: out[b] = mean_t x[b, 1 : 1+lengths[b]] (residues only,
 * <cls> at position 0 excluded). x fp32 [B,T,E], lengths int32 [B] (device), out fp32 [B,E]. */
int esmb200_mean_pool(const float* x, const int32_t* lengths, float* out, int32_t B, int32_t T, int32_t E,
                      void* stream);

/* ---- single-kernel entry points (used by the parity tests and profiles; same kernels as above) ---- */

/* out = epilogue(A[M,K] fp16 x W[N,K]^T fp16 + bias[N]);  epilogue: 0 qkv+rope -> fp16, 1 residual-add into fp32 out,
 * 2 gelu -> fp16, 3 fp32, 4 gelu -> fp32. rope_* / T / E only for epilogue 0. K % 64 == 0, N % 64 == 0. */
int esmb200_gemm_f16(int32_t epilogue, const void* a_f16, const void* w_f16, const float* bias, void* out, int32_t M,
                     int32_t N, int32_t K, const float* rope_cos, const float* rope_sin, int32_t T, int32_t E,
                     void* stream);

/* qkv[M,3E] fp16 = A[M,E] fp16 x [Wq;Wk;Wv]^T + bias, q columns scaled by q_scale, q/k rotated when rope tables are
 * given (ESM-2: multihead_attention.py:258-261,354-355) or left unrotated when rope_cos == rope_sin == NULL (MSA axial
 * attention: axial_attention.py:79-81 with q_scale = d^-1/2 / sqrt(rows), :199-202 with q_scale = d^-1/2). */
int esmb200_gemm_qkv_f16(const void* a_f16, const void* w_qkv_f16, const float* bias_qkv, void* out_f16, int32_t M,
                         int32_t E, float q_scale, const float* rope_cos, const float* rope_sin, int32_t T,
                         void* stream);

/* ctx[B*T,E] fp16 = softmax(q k^T + key padding mask) v per head, from qkv fp16 [B*T,3E] (q pre-scaled, q/k rotated).
 * scratch: at least esmb200_attention_scratch_bytes(B,T). attn_probs as in esmb200_layer_forward. */
size_t esmb200_attention_scratch_bytes(int32_t B, int32_t T);
int esmb200_attention(const void* qkv_f16, const uint8_t* pad_mask, void* ctx_f16, float* attn_probs, int32_t B,
                      int32_t T, int32_t H, void* scratch, void* stream);
/* The same for head_dim 128 (esm2_t48_15B, esm/pretrained.py:390-397): qkv fp16 [B*T, 3*128*H], ctx fp16 [B*T, 128*H];
 * a head is two adjacent 64-wide column slots, any fixed permutation of the 128 dimensions that is
 * shared by q, k and v gives the same result. */
int esmb200_attention128(const void* qkv_f16, const uint8_t* pad_mask, void* ctx_f16, float* attn_probs, int32_t B,
                         int32_t T, int32_t H, void* scratch, void* stream);

/* ---- MSA Transformer axial attention on qkv fp16 [B*R*C, 3E] (B alignments, R rows, C columns; q pre-scaled) ----
 * esmb200_tied_row_attention: RowSelfAttention.compute_attention_weights / compute_attention_update,
 *   esm/axial_attention.py:71-111 — logits summed over the R rows (:87), key_pad [B,C] (1 = padded key column, filled
 *   with -10000, :94-97; NULL = none), softmax over the key columns (:105), ctx[B*R*C,E] fp16 = P v per row (:108).
 *   The caller zeroes q at padded positions (:82-85). attn_probs: optional fp32 [H,B,C,C] (the reference's return
 *   layout), NULL to skip. scratch: esmb200_tied_row_attention_scratch_bytes(B,C,H). C <= 1024.
 * esmb200_column_attention: ColumnSelfAttention.compute_attention_update, esm/axial_attention.py:182-222 — per
 *   alignment column, attention over the R rows; pad_mask [B*C, R] (1 = padded key; such keys get probability 0 where
 *   the reference fills -10000, identical unless every key of a column is padded), ctx[B*R*C,E] fp16.
 *   scratch: esmb200_attention_scratch_bytes(B*C, R). */
size_t esmb200_tied_row_attention_scratch_bytes(int32_t B, int32_t C, int32_t H);
int esmb200_tied_row_attention(const void* qkv_f16, const uint8_t* key_pad, void* ctx_f16, float* attn_probs, int32_t B,
                               int32_t R, int32_t C, int32_t H, void* scratch, size_t scratch_bytes, void* stream);
int esmb200_column_attention(const void* qkv_f16, const uint8_t* pad_mask, void* ctx_f16, int32_t B, int32_t R,
                             int32_t C, int32_t H, void* scratch, void* stream);

/* n_layers x AxialTransformerLayer.forward (esm/modules.py:195-221) = the layer loop of MSATransformer.forward
 * (esm/model/msa_transformer.py:190-201), in place on the batch-major residual stream x [B,R,C,E] fp32:
 *   x += out_proj(tied_row_attention(LN(x)));  x += out_proj(column_attention(LN(x)));  x += fc2(gelu(fc1(LN(x))))
 * row_layers[i]: an esmb200_layer created with fc1_weight == NULL (attention-only) from row_self_attention's
 *   layer_norm + q/k/v/out projections; col_layers[i]: column_self_attention's layer_norm + projections as ln1/q/k/v/out
 *   and feed_forward_layer's layer_norm + fc1/fc2 as ln2/fc1/fc2.
 * pad_mask [B,R,C] and col_pad_mask [B,C,R] (its transpose): 1 = padding, both NULL for unpadded alignments.
 * row_attn_out: NULL, or n_layers pointers (NULL entries allowed) to fp32 [H,B,C,C] buffers (the reference's
 *   row-attention return layout, axial_attention.py:87,105). Column attention maps are not produced by this call.
 * workspace: esmb200_axial_workspace_bytes(E,F,B,R,C) bytes. */
size_t esmb200_axial_workspace_bytes(int32_t E, int32_t F, int32_t B, int32_t R, int32_t C);
int esmb200_axial_stack_forward(esmb200_layer* const* row_layers, esmb200_layer* const* col_layers, int32_t n_layers,
                                float* x, const uint8_t* pad_mask, const uint8_t* col_pad_mask, int32_t B, int32_t R,
                                int32_t C, float* const* row_attn_out, void* workspace, size_t workspace_bytes,
                                void* stream);

/* MSA Transformer embedding prologue, esm/model/msa_transformer.py:155-172 (+ LearnedPositionalEmbedding.forward,
 * esm/modules.py:241-257): x[B,R,C,E] fp32 = LayerNorm(embed_tokens[tok] + embed_positions[pos] +
 * msa_position_embedding[r]) * (1 - is_pad). tokens int64 [B,R,C]; pos_table [max_positions + padding_idx + 1, E];
 * msa_pos [>=R, msa_pos_dim] or NULL, msa_pos_dim = E or 1 (the initial esm_msa1 release, pretrained.py:123-125). */
int esmb200_msa_embed(const int64_t* tokens, const float* embed_table, const float* pos_table, const float* msa_pos,
                      int32_t msa_pos_dim, const float* ln_weight, const float* ln_bias, float eps, float* x,
                      int32_t B, int32_t R, int32_t C, int32_t E, int32_t padding_idx, void* stream);

/* Contact head, one layer's share (ContactPredictionHead.forward esm/modules.py:338-357, symmetrize :27-29, apc :32-41):
 * attn = that layer's attention maps fp32 [B,H,T,T] (batch_stride floats between batch elements, so a slice of a stacked
 * [B,L,H,T,T] tensor works), cropped to positions [lo,hi) and multiplied by keep[b,i]*keep[b,j] (keep [B,T], 1 = not
 * <eos>; NULL = no masking). S = hi - lo.
 *   acc      [B,S,S]            += sum_h w[h] * A_h   (zeroed by the caller before the first layer)
 *   row_sum  [B,H,S]             = rowsum(A_h)
 *   col_part [B,H,ceil(S/16),S]  = column sums of each 16-row stripe; colsum(A_h) = sum over the stripe axis
 * No atomics: results are bit-reproducible. a1_c = row_sum + colsum feeds esmb200_contact_finalize. */
int esmb200_contact_accumulate(const float* attn, int64_t batch_stride, const float* w, const uint8_t* keep, float* acc,
                               float* row_sum, float* col_part, int32_t B, int32_t H, int32_t T, int32_t lo, int32_t hi,
                               void* stream);

/* Contact head tail (modules.py:33-41,352-357): out[b,i,j] = sigmoid(acc[b,i,j] + acc[b,j,i] - sum_c u[b,c,i]*a1[b,c,j] + bias)
 * with a1 [B,C,S] (C = layers*heads channels) and u = a1 * w_c / sum_i a1_c[i] prepared by the caller; bias = device
 * pointer to one float or NULL; out [B,S,S]. */
int esmb200_contact_finalize(const float* acc, const float* u, const float* a1, const float* bias, float* out, int32_t B,
                             int32_t C, int32_t S, void* stream);

/* fp32 [M,E] -> LayerNorm -> fp16 [M,E] (the GEMM A operand) */
int esmb200_layernorm_f16(const float* x, const float* weight, const float* bias, void* out_f16, int32_t M, int32_t E,
                          float eps, void* stream);

/* ---- launch accounting and per-launch timing (bench.py roofline numbers) ----
 * esmb200_launch_count: kernels launched by this library since it was loaded.
 * esmb200_profile_enable(n): n > 0 brackets each of the next n launches with CUDA events on the launch stream
 *   (0 disables and frees the events); esmb200_profile_read returns up to max_records (tag, milliseconds) pairs,
 *   synchronising on the recorded events, and resets the record list.
 *   tags: 0 LN1->f16, 1 QKV+RoPE GEMM, 2 attention, 3 out-proj GEMM, 4 LN2->f16, 5 fc1+GELU GEMM, 6 fc2 GEMM,
 *         7 key bits, 8 embed, 9 LayerNorm fp32, 10 attention probs, 11 convert, 12 other GEMM, 13 mean pool,
 *         14 tied row logits, 15 tied row softmax, 16 tied row update */
long long esmb200_launch_count(void);
int esmb200_profile_enable(int32_t max_launches);
int esmb200_profile_read(int32_t* tags, float* ms, int32_t max_records);

/* fp32 -> fp16 elementwise */
int esmb200_convert_f16(const float* src, void* dst_f16, size_t n, void* stream);

/* ---- fp32x3 precision building blocks (operands as fp16 hi | lo pairs along K): the LM head and kernel-level tests.
 * esmb200_layernorm_split: fp32 [M,E] -> LayerNorm -> fp16 [M,2E] (hi in columns [0,E), lo = rn(y - hi) in [E,2E)).
 * esmb200_convert_split:   fp32 [rows,K] -> fp16 [rows,2K] the same way (weights).
 * esmb200_gemm_split:      esmb200_gemm_f16 with a [M,2K], w [N,2K]; fp16 outputs (QKV_ROPE, BIAS_GELU) are written as
 *                          [M,2N] hi | lo, fp32 outputs as [M,N].  K % 64 == 0.
 * esmb200_attention_split: esmb200_attention on qkv [B*T, 6E] = [q k v]_hi | [q k v]_lo -> ctx [B*T, 2E] hi | lo. */
int esmb200_layernorm_split(const float* x, const float* weight, const float* bias, void* out_f16, int32_t M, int32_t E,
                            float eps, void* stream);
int esmb200_convert_split(const float* src, void* dst_f16, int64_t rows, int32_t K, void* stream);
int esmb200_gemm_split(int32_t epilogue, const void* a, const void* w, const float* bias, void* out, int32_t M,
                       int32_t N, int32_t K, const float* rope_cos, const float* rope_sin, int32_t T, int32_t E,
                       void* stream);
int esmb200_attention_split(const void* qkv, const uint8_t* pad_mask, void* ctx, float* attn_probs, int32_t B, int32_t T,
                            int32_t H, void* scratch, void* stream);

/* ---- process-wide kernel selection knobs (A/B measurements; the defaults are the product configuration) ----
 * "attn"      8 (attention8.cuh, 4 CTAs/SM) | 7 (attention7.cuh, the round-1 kernel: only in a library built with
 *             -DESMB200_EXPERIMENTS, A/B measurements)                                          env ESMB200_ATTN
 * "attn_poly" 0 | 2 | 3 | 4 (default): every n-th pair of softmax exponentials on the FMA pipe   env ESMB200_ATTN_POLY
 * "pdl"       0 (default) | 1: programmatic dependent launch between the layer's kernels        env ESMB200_PDL
 * Returns ESMB200_EINVAL for an unknown name or value. Not thread-safe against concurrent launches. */
int esmb200_set_option(const char* name, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* ESMB200_H_ */
