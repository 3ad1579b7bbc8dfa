// esm_b200 — C ABI (include/esmb200.h): host-side orchestration of the sm_100a kernels.
//
// Everything here is plain C-callable: device pointers in, launches on the caller's stream, no torch types.
// Host work per call is limited to encoding a handful of TMA descriptors (cuTensorMapEncodeTiled) and
// launching 7 kernels per TransformerLayer:
//   LN1->fp16 | QKV GEMM (+bias, q scale, RoPE) | attention | out-proj GEMM (+bias, residual)
//   LN2->fp16 | fc1 GEMM (+bias, erf-GELU)      | fc2 GEMM (+bias, residual)
#include "../../include/esmb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <mutex>

#ifdef ESMB200_EXPERIMENTS
#include "attention7.cuh"  // round-1 kernel, A/B only
#endif
#include "attention8.cuh"
#include "attention_contact.cuh"
#include "attention_probs.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm2.cuh"
#include "tied_attention.cuh"

using namespace esmb200;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int fail_cuda(cudaError_t e, const char* what) {
  if (e == cudaErrorMemoryAllocation)
    return fail(ESMB200_ENOMEM, std::string("CUDA out of memory. (") + what + ")");
  return fail(ESMB200_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CK(expr)                                        \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return fail_cuda(_e, #expr); \
  } while (0)


// ---- launch accounting + optional per-launch CUDA-event timing (bench.py's roofline numbers) --------------------
enum ProfTag : int { T_LN1 = 0, T_QKV, T_ATTN, T_OUT, T_LN2, T_FC1, T_FC2, T_KEYBITS, T_EMBED, T_LN_F32, T_PROBS,
                     T_CONVERT, T_GEMM_OTHER, T_MEANPOOL, T_TIED_SCORES, T_TIED_SOFTMAX, T_TIED_PV, T_COUNT };
struct Profiler {  // process-wide, guarded by `mu`: launches may come from several host threads / streams
  std::mutex mu;
  bool on = false;
  std::vector<cudaEvent_t> ev;  // pairs (start, stop)
  std::vector<int> tag;
  size_t used = 0;              // events used
  long long launches = 0;       // kernels launched by this library since load
};
Profiler g_prof;

struct ProfScope {
  cudaStream_t st;
  bool rec;
  size_t slot = 0;
  ProfScope(int tag, cudaStream_t s) : st(s), rec(false) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    ++g_prof.launches;
    if (g_prof.on && g_prof.used + 2 <= g_prof.ev.size()) {
      rec = true;
      slot = g_prof.used;
      g_prof.used += 2;
      g_prof.tag.push_back(tag);
      cudaEventRecord(g_prof.ev[slot], st);
    }
  }
  ~ProfScope() {
    if (rec) cudaEventRecord(g_prof.ev[slot + 1], st);
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2D row-major [rows, cols] (cols contiguous) of fp16 (esize 2) or fp32 (esize 4);
// box = {128 bytes of columns, box_rows}, 128B swizzle.
int make_tmap_2d(CUtensorMap* map, const void* ptr, int esize, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ESMB200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld_elems * esize) % 16 != 0)
    return fail(ESMB200_EINVAL, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * (uint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
    return fail(ESMB200_ECUDA, buf);
  }
  return ESMB200_OK;
}

int make_tmap_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                  uint32_t box_rows) {
  return make_tmap_2d(map, ptr, 2, rows, cols, ld_elems, box_rows);
}

// attention implementation: 8 = attention8.cuh (4 CTAs/SM, default), 7 = attention7.cuh (2 CTAs/SM, two MMA issuing
// threads; kept for A/B runs).  ESMB200_ATTN_POLY = n sends every n-th pair of exponentials of v8 to the FMA pipe.
int g_attn_version = -1, g_attn_poly = -1;  // -1: take the environment / default on first use

int attn_version() {
  if (g_attn_version < 0) {
    const char* e = getenv("ESMB200_ATTN");
#ifdef ESMB200_EXPERIMENTS
    g_attn_version = (e && e[0] == '7') ? 7 : 8;
#else
    g_attn_version = 8;
#endif
  }
  return g_attn_version;
}

int attn_poly() {
  if (g_attn_poly < 0) {
    const char* e = getenv("ESMB200_ATTN_POLY");
    g_attn_poly = (e && (e[0] == '0' || e[0] == '2' || e[0] == '3' || e[0] == '4')) ? (e[0] - '0') : 4;
  }
  return g_attn_poly;
}

cudaError_t launch_attention_fwd(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnParams& ap, int sms,
                                 cudaStream_t st) {
  if (ap.lo_off > 0) return launch_attention_v8_poly<0, true>(tq, tkv, ap, sms, st);  // fp32x3: all exponentials on MUFU
  if (ap.slots == 2) return launch_attention_v8_poly<4, false, 2>(tq, tkv, ap, sms, st);  // head_dim <= 128 (15B)
#ifdef ESMB200_EXPERIMENTS
  if (attn_version() == 7) return launch_attention_v7(tq, tkv, ap, sms, st);
#endif
  switch (attn_poly()) {
    case 0: return launch_attention_v8_poly<0>(tq, tkv, ap, sms, st);
    case 2: return launch_attention_v8_poly<2>(tq, tkv, ap, sms, st);
    case 3: return launch_attention_v8_poly<3>(tq, tkv, ap, sms, st);
    default: return launch_attention_v8_poly<4>(tq, tkv, ap, sms, st);
  }
}

// 64-wide column slots per head on the attention side: 1 for head_dim <= 64, 2 up to 128 (elementwise.cuh head_slot)
inline int head_slots(int E, int H) { return (H > 0 && E / H > 64) ? 2 : 1; }

int qkv_chunked() {  // tile walk of the QKV GEMM (gemm2.cuh): 1 = one contiguous run of tiles per cluster (default)
  static const int v = [] {
    const char* e = getenv("ESMB200_QKV_CHUNKED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

constexpr int kMaxDevices = 64;

int num_sms() {  // per device: one process may drive several GPUs
  static int n[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

int check_device() {
  static int ok[kMaxDevices] = {};  // 0 unknown, 1 sm_100, -1 other
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(ESMB200_ECUDA, "no CUDA device");
  if (dev < 0 || dev >= kMaxDevices) return fail(ESMB200_ECUDA, "device ordinal out of range");
  if (ok[dev] == 0) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok[dev] = (major == 10) ? 1 : -1;
  }
  if (ok[dev] < 0)
    return fail(ESMB200_ECUDA, "esmb200 requires an sm_100a (Blackwell B200) device; there is no fallback path");
  return ESMB200_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const GemmParams& p,
                cudaStream_t st, int tag = T_GEMM_OTHER, bool split = false) {
  ProfScope ps(tag, st);
  cudaError_t e;
  if (split) {  // fp32x3 precision: operands stored as fp16 hi | lo along K (gemm2.cuh)
    if (p.K % 64 != 0) return fail(ESMB200_EINVAL, "fp32x3 precision needs K % 64 == 0");
    switch (epi) {
      case EPI_QKV_ROPE: e = launch_gemm2_epi<EPI_QKV_ROPE, true>(ta, tb, tout, p, num_sms(), st); break;
      case EPI_BIAS_RESIDUAL: e = launch_gemm2_epi<EPI_BIAS_RESIDUAL, true>(ta, tb, tout, p, num_sms(), st); break;
      case EPI_BIAS_GELU: e = launch_gemm2_epi<EPI_BIAS_GELU, true>(ta, tb, tout, p, num_sms(), st); break;
      case EPI_BIAS_F32: e = launch_gemm2_epi<EPI_BIAS_F32, true>(ta, tb, tout, p, num_sms(), st); break;
      case EPI_BIAS_GELU_F32: e = launch_gemm2_epi<EPI_BIAS_GELU_F32, true>(ta, tb, tout, p, num_sms(), st); break;
      default: return fail(ESMB200_EINVAL, "unknown GEMM epilogue");
    }
    if (e != cudaSuccess) return fail_cuda(e, "gemm launch (fp32x3)");
    return ESMB200_OK;
  }
  switch (epi) {
    case EPI_QKV_ROPE: e = launch_gemm2_epi<EPI_QKV_ROPE>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_BIAS_RESIDUAL: e = launch_gemm2_epi<EPI_BIAS_RESIDUAL>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_BIAS_GELU: e = launch_gemm2_epi<EPI_BIAS_GELU>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_BIAS_F32: e = launch_gemm2_epi<EPI_BIAS_F32>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_BIAS_GELU_F32: e = launch_gemm2_epi<EPI_BIAS_GELU_F32>(ta, tb, tout, p, num_sms(), st); break;
#ifdef ESMB200_EXPERIMENTS  // profiling-only epilogues (profiles/r01_epilogue_experiments.txt)
    case EPI_NONE: e = launch_gemm2_epi<EPI_NONE>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_LDONLY: e = launch_gemm2_epi<EPI_LDONLY>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_LD_X16: e = launch_gemm2_epi<EPI_LD_X16>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_LD_4WARPS: e = launch_gemm2_epi<EPI_LD_4WARPS>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_LD_BATCH: e = launch_gemm2_epi<EPI_LD_BATCH>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_GELU_MATHONLY: e = launch_gemm2_epi<EPI_GELU_MATHONLY>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_F16_STOREONLY: e = launch_gemm2_epi<EPI_F16_STOREONLY>(ta, tb, tout, p, num_sms(), st); break;
    case EPI_FMA_MATHONLY: e = launch_gemm2_epi<EPI_FMA_MATHONLY>(ta, tb, tout, p, num_sms(), st); break;
#endif
    default: return fail(ESMB200_EINVAL, "unknown GEMM epilogue");
  }
  if (e != cudaSuccess) return fail_cuda(e, "gemm launch");
  return ESMB200_OK;
}

// scratch layout of the attention kernels
struct AttnScratch {
  uint32_t* keybits;
  int* kvlen;
  float* row_max;
  float* row_sum;
  int words;
};

size_t attn_scratch_bytes(int B, int T, int H) {
  const int words = (int)align_up((size_t)(T + 31) / 32, 4);
  return align_up((size_t)B * words * 4, 256) + align_up((size_t)B * 4, 256) + 2 * align_up((size_t)B * H * T * 4, 256);
}

AttnScratch carve_attn_scratch(void* scratch, int B, int T, int H) {
  AttnScratch s;
  s.words = (int)align_up((size_t)(T + 31) / 32, 4);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  s.keybits = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * s.words * 4, 256);
  s.kvlen = reinterpret_cast<int*>(p);
  p += align_up((size_t)B * 4, 256);
  s.row_max = reinterpret_cast<float*>(p);
  p += align_up((size_t)B * H * T * 4, 256);
  s.row_sum = reinterpret_cast<float*>(p);
  return s;
}

int run_key_bits(const uint8_t* pad_mask, const AttnScratch& s, int B, int T, cudaStream_t st) {
  const int wpb = 4;
  ProfScope ps(T_KEYBITS, st);
  key_bits_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, st>>>(pad_mask, s.keybits, s.kvlen, B, T, s.words);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

// contact-head accumulators of ONE layer (esmb200_contact_job resolved for layer i)
struct ContactLayer {
  const float* w;
  const uint8_t* keep;
  float* acc;
  float* row_part;
  float* col_part;
  int lo, S;
};

int run_attention(const void* qkv, void* ctx, float* probs, long long probs_batch_stride, int attn_flags,
                  const AttnScratch& s, int B, int T, int H, cudaStream_t st, bool split = false,
                  const ContactLayer* contact = nullptr, int slots = 1) {
  const int E = H * 64 * slots;
  if (slots == 2 && split) return fail(ESMB200_EINVAL, "head_dim > 64: fp32x3 precision is not available");
  const uint64_t qcols = (uint64_t)(split ? 6 : 3) * E;  // fp32x3: [q k v]_hi | [q k v]_lo
  CUtensorMap tq;
  int rc = make_tmap_f16(&tq, qkv, (uint64_t)B * T, qcols, qcols, 128);
  if (rc) return rc;
  AttnParams ap;
  ap.B = B; ap.T = T; ap.H = H; ap.E = E;
  ap.lo_off = split ? 3 * E : 0;
  ap.slots = slots;
  ap.keybits = s.keybits; ap.kvlen = s.kvlen; ap.words = s.words;
  ap.ctx = static_cast<__half*>(ctx);
  ap.row_max = probs ? s.row_max : nullptr;
  ap.row_sum = probs ? s.row_sum : nullptr;
  cudaError_t e;
  {
    CUtensorMap tkv;
    rc = make_tmap_f16(&tkv, qkv, (uint64_t)B * T, qcols, qcols, attn8_cfg::BLOCK_KV);
    if (rc) return rc;
    ProfScope ps(T_ATTN, st);
    e = launch_attention_fwd(tq, tkv, ap, num_sms(), st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "attention launch");
  if (probs && contact && !split) {
    // probabilities written once and folded into the contact accumulators in the same pass (attention_contact.cuh)
    if (B > 65535) return fail(ESMB200_EINVAL, "return_contacts: B must be <= 65535");
    ContactFuseParams cp;
    cp.B = B; cp.T = T; cp.H = H; cp.E = E; cp.slots = slots;
    cp.keybits = s.keybits; cp.kvlen = s.kvlen; cp.words = s.words;
    cp.row_max = s.row_max; cp.row_sum = s.row_sum; cp.probs = probs;
    cp.batch_stride = probs_batch_stride > 0 ? probs_batch_stride : (long long)H * T * T;
    cp.zero_pad_rows = attn_flags & 1;
    cp.w = contact->w; cp.keep = contact->keep; cp.acc = contact->acc;
    cp.row_part = contact->row_part; cp.col_part = contact->col_part; cp.lo = contact->lo; cp.S = contact->S;
    {
      ProfScope ps(T_PROBS, st);
      e = launch_attention_probs_contact(tq, cp, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "attention probs+contact launch");
    return ESMB200_OK;
  }
  if (probs) {
    if ((size_t)B * H > 65535) return fail(ESMB200_EINVAL, "need_head_weights: B*H must be <= 65535");
    ProbsParams pp;
    pp.B = B; pp.T = T; pp.H = H; pp.E = E;
    pp.keybits = s.keybits; pp.kvlen = s.kvlen; pp.words = s.words;
    pp.row_max = s.row_max; pp.row_sum = s.row_sum; pp.probs = probs;
    pp.batch_stride = probs_batch_stride > 0 ? probs_batch_stride : (long long)H * T * T;
    pp.zero_pad_rows = attn_flags & 1;
    pp.lo_off = split ? 3 * E : 0;
    pp.slots = slots;
    {
      ProfScope ps(T_PROBS, st);
      e = launch_attention_probs(tq, pp, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "attention probs launch");
  }
  return ESMB200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// layer object
// ---------------------------------------------------------------------------------------------------------------
struct esmb200_layer {
  int E, H, F;
  int d;          // head_dim (<= 128); every head occupies `slots` 64-wide slots of the attention-side tensors
  int slots;      // 1 (d <= 64) or 2 (64 < d <= 128, ESM-2 15B)
  int Ea;         // 64 * slots * H: width of q / k / v / ctx
  float q_scale;  // d^-1/2 (multihead_attention.py:100)
  int split;      // 1: fp32x3 precision — weights packed as fp16 hi | lo along K, activations likewise
  float eps;
  // borrowed fp32 parameters (owned by the caller, must outlive the layer)
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *out_b, *fc1_b, *fc2_b;
  // owned packed copies
  __half* w_qkv;  // [3*Ea, E]: row s*Ea + head_slot(h*d + j) <- W_s row h*d + j, zero rows elsewhere
  __half* w_out;  // [E, Ea]: column head_slot(h*d + j) <- out_proj.weight column h*d + j
  __half* w_fc1;  // [F,E]
  __half* w_fc2;  // [E,F]
  float* b_qkv;   // [3*Ea]
  CUtensorMap tm_qkv, tm_out, tm_fc1, tm_fc2;  // B operands, box {64, 128 rows}
};

extern "C" {

int esmb200_abi_version(void) { return ESMB200_ABI_VERSION; }

const char* esmb200_last_error(void) { return g_last_error.c_str(); }

int esmb200_convert_f16(const float* src, void* dst, size_t n, void* stream) {
  if (n == 0) return ESMB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps(T_CONVERT, st);
  convert_f32_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, static_cast<__half*>(dst), n);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

int esmb200_layer_destroy(esmb200_layer* L) {
  if (!L) return ESMB200_OK;
  cudaFree(L->w_qkv);
  cudaFree(L->w_out);
  cudaFree(L->w_fc1);
  cudaFree(L->w_fc2);
  cudaFree(L->b_qkv);
  delete L;
  return ESMB200_OK;
}

int esmb200_layer_create(const esmb200_layer_weights* w, void* stream, esmb200_layer** out) {
  if (!w || !out) return fail(ESMB200_EINVAL, "null argument");
  int rc = check_device();
  if (rc) return rc;
  const int E = w->embed_dim, H = w->num_heads, F = w->ffn_dim;
  if (E <= 0 || H <= 0 || E % H != 0) return fail(ESMB200_EINVAL, "embed_dim must be a positive multiple of num_heads");
  const int d = w->head_dim > 0 ? w->head_dim : E / H;
  if (d * H != E || d > 128 || d % 2 != 0)
    return fail(ESMB200_EINVAL, "esmb200 supports even head_dim <= 128 (every ESM-2 model, MSA Transformer)");
  if (E % 16 != 0) return fail(ESMB200_EINVAL, "embed_dim must be a multiple of 16");
  const bool has_ffn = w->fc1_weight != nullptr;  // NULL fc1_weight: attention-only layer (MSA row-attention sub-layer)
  if (has_ffn && (F <= 0 || F % 64 != 0)) return fail(ESMB200_EINVAL, "ffn_dim must be a positive multiple of 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  esmb200_layer* L = new esmb200_layer();
  memset(static_cast<void*>(L), 0, sizeof(*L));
  const int slots = head_slots(E, H);
  const int Ea = 64 * slots * H;
  L->E = E; L->H = H; L->F = has_ffn ? F : 0; L->d = d; L->slots = slots; L->Ea = Ea; L->eps = w->ln_eps;
  L->q_scale = 1.0f / sqrtf((float)d);
  L->ln1_w = w->ln1_weight; L->ln1_b = w->ln1_bias; L->ln2_w = w->ln2_weight; L->ln2_b = w->ln2_bias;
  L->out_b = w->out_bias; L->fc1_b = w->fc1_bias; L->fc2_b = w->fc2_bias;
  if (w->precision != 0 && w->precision != 1) {
    delete L;
    return fail(ESMB200_EINVAL, "precision must be 0 (fp16 operands) or 1 (fp32x3: fp16 hi|lo operands)");
  }
  const int split = w->precision;
  if (split && slots == 2) {
    delete L;
    return fail(ESMB200_EINVAL, "fp32x3 precision is not available for head_dim > 64");
  }
  if (split && (E % 64 != 0 || (has_ffn && F % 64 != 0))) {
    delete L;
    return fail(ESMB200_EINVAL, "fp32x3 precision needs embed_dim % 64 == 0 (all ESM-2 models except 35M)");
  }
  L->split = split;
  const size_t pf = split ? 2 : 1;  // fp32x3: every K extent doubles (hi | lo)
  const size_t EaE = (size_t)Ea * E, EF = (size_t)E * F;
  cudaError_t e;
#define ALLOC(ptr, bytes)                                              \
  if ((e = cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes))) != cudaSuccess) { \
    esmb200_layer_destroy(L);                                          \
    return fail_cuda(e, "cudaMalloc(packed weights)");                 \
  }
  ALLOC(L->w_qkv, 3 * EaE * 2 * pf);
  ALLOC(L->w_out, EaE * 2 * pf);
  if (has_ffn) {
    ALLOC(L->w_fc1, EF * 2 * pf);
    ALLOC(L->w_fc2, EF * 2 * pf);
  }
  ALLOC(L->b_qkv, (size_t)3 * Ea * 4);
#undef ALLOC
  auto convert = [&](const float* src, __half* dst, size_t rows, int K) -> int {  // [rows,K] fp32 -> fp16 (hi | lo)
    if (!split) return esmb200_convert_f16(src, dst, rows * (size_t)K, stream);
    size_t blocks = (rows * (size_t)K + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope ps(T_CONVERT, st);
    convert_f32_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, rows, K);
    cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? ESMB200_OK : fail_cuda(ce, "convert_f32_split");
  };
  if (d == 64 && !split) {  // slots are full: plain conversion
    rc = esmb200_convert_f16(w->q_weight, L->w_qkv, EaE, stream);
    if (!rc) rc = esmb200_convert_f16(w->k_weight, L->w_qkv + EaE, EaE, stream);
    if (!rc) rc = esmb200_convert_f16(w->v_weight, L->w_qkv + 2 * EaE, EaE, stream);
    if (!rc) rc = esmb200_convert_f16(w->out_weight, L->w_out, EaE, stream);
    if (!rc) {
      e = cudaMemcpyAsync(L->b_qkv, w->q_bias, (size_t)E * 4, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(L->b_qkv + E, w->k_bias, (size_t)E * 4, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(L->b_qkv + 2 * E, w->v_bias, (size_t)E * 4, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) rc = fail_cuda(e, "bias pack");
    }
  } else {  // head_dim < 64 (scatter every head into its zero-padded 64-wide slot) and / or hi | lo operands
    e = cudaMemsetAsync(L->w_qkv, 0, 3 * EaE * 2 * pf, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(L->w_out, 0, EaE * 2 * pf, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(L->b_qkv, 0, (size_t)3 * Ea * 4, st);
    if (e != cudaSuccess) rc = fail_cuda(e, "memset(packed weights)");
    const float* ws3[3] = {w->q_weight, w->k_weight, w->v_weight};
    const float* bs3[3] = {w->q_bias, w->k_bias, w->v_bias};
    const unsigned blocks = (unsigned)(((size_t)E * E + 255) / 256);
    for (int s3 = 0; s3 < 3 && !rc; ++s3) {
      ProfScope ps(T_CONVERT, st);
      pack_head_rows_kernel<<<blocks, 256, 0, st>>>(ws3[s3], bs3[s3], L->w_qkv + (size_t)s3 * EaE * pf, L->b_qkv + s3 * Ea,
                                                    E, d, split);
      if ((e = cudaGetLastError()) != cudaSuccess) rc = fail_cuda(e, "pack_head_rows");
    }
    if (!rc) {
      ProfScope ps(T_CONVERT, st);
      pack_head_cols_kernel<<<blocks, 256, 0, st>>>(w->out_weight, L->w_out, E, Ea, d, split);
      if ((e = cudaGetLastError()) != cudaSuccess) rc = fail_cuda(e, "pack_head_cols");
    }
  }
  if (!rc && has_ffn) rc = convert(w->fc1_weight, L->w_fc1, F, E);
  if (!rc && has_ffn) rc = convert(w->fc2_weight, L->w_fc2, E, F);
  const uint32_t wbox = gemm2_cfg::HALF_N;
  if (!rc) rc = make_tmap_f16(&L->tm_qkv, L->w_qkv, 3 * (uint64_t)Ea, pf * E, pf * E, wbox);
  if (!rc) rc = make_tmap_f16(&L->tm_out, L->w_out, E, pf * Ea, pf * Ea, wbox);
  if (!rc && has_ffn) rc = make_tmap_f16(&L->tm_fc1, L->w_fc1, F, pf * E, pf * E, wbox);
  if (!rc && has_ffn) rc = make_tmap_f16(&L->tm_fc2, L->w_fc2, E, pf * F, pf * F, wbox);
  if (rc) {
    esmb200_layer_destroy(L);
    return rc;
  }
  *out = L;
  return ESMB200_OK;
}

size_t esmb200_attention_scratch_bytes(int32_t B, int32_t T) {
  // H is bounded by E/64; the stats arrays are sized by the caller-visible worst case through workspace_bytes,
  // this standalone entry sizes them for H <= 64.
  return attn_scratch_bytes(B, T, 64);
}

size_t esmb200_workspace_bytes(int32_t E, int32_t H, int32_t F, int32_t B, int32_t T, int32_t precision) {
  const size_t M = (size_t)B * T, Ea = (size_t)64 * head_slots(E, H) * H, pf = precision ? 2 : 1;
  const size_t a = align_up(M * E * 2 * pf, 1024);                  // xn fp16 [M,E] (hi | lo)
  const size_t big_qkv_ctx = align_up(M * 3 * Ea * 2 * pf, 1024) + align_up(M * Ea * 2 * pf, 1024);
  const size_t big_h = align_up(M * F * 2 * pf, 1024);
  const size_t big = big_qkv_ctx > big_h ? big_qkv_ctx : big_h;    // h aliases qkv+ctx
  return a + big + attn_scratch_bytes(B, T, H) + 1024;
}

namespace {
struct Workspace {
  __half* xn;
  __half* qkv;
  __half* ctx;
  __half* h;
  AttnScratch as;
};

int carve_workspace(Workspace* ws, void* workspace, size_t bytes, int E, int H, int F, int B, int T, int split) {
  if (bytes < esmb200_workspace_bytes(E, H, F, B, T, split)) return fail(ESMB200_EWORKSPACE, "workspace too small");
  const size_t M = (size_t)B * T, Ea = (size_t)64 * head_slots(E, H) * H, pf = split ? 2 : 1;
  uint8_t* p = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  ws->xn = reinterpret_cast<__half*>(p);
  p += align_up(M * E * 2 * pf, 1024);
  ws->qkv = reinterpret_cast<__half*>(p);
  ws->h = reinterpret_cast<__half*>(p);
  ws->ctx = reinterpret_cast<__half*>(p + align_up(M * 3 * Ea * 2 * pf, 1024));
  const size_t big_qkv_ctx = align_up(M * 3 * Ea * 2 * pf, 1024) + align_up(M * Ea * 2 * pf, 1024);
  const size_t big_h = align_up(M * F * 2 * pf, 1024);
  p += big_qkv_ctx > big_h ? big_qkv_ctx : big_h;
  ws->as = carve_attn_scratch(p, B, T, H);
  return ESMB200_OK;
}

struct ActMaps {
  CUtensorMap xn, ctx, h;       // A operands (fp16, box {64,128}); h doubles as fc1's output map
  CUtensorMap qkv_out, x_out;   // epilogue outputs: qkv fp16 [M,3E] box {64,128}; x fp32 [M,E] box {32,128}
};

int layer_forward_impl(esmb200_layer* L, float* x, int B, int T, const float* rope_cos, const float* rope_sin,
                       float* attn_probs, long long attn_batch_stride, int attn_flags, const Workspace& ws,
                       const ActMaps& am, cudaStream_t st, const ContactLayer* contact = nullptr) {
  const int E = L->E, F = L->F, H = L->H, Ea = L->Ea;
  const int M = B * T;
  const bool split = L->split != 0;
  cudaError_t e = cudaSuccess;
  // LN1 -> fp16 (modules.py:124)
  {
    ProfScope ps(T_LN1, st);
    e = split ? launch_layernorm<2>(x, L->ln1_w, L->ln1_b, ws.xn, M, E, L->eps, st)
              : launch_layernorm<1>(x, L->ln1_w, L->ln1_b, ws.xn, M, E, L->eps, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "layernorm1");
  // q,k,v projections + bias + q scale + RoPE (multihead_attention.py:258-261,354-355)
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = 3 * Ea; g.K = E; g.bias = L->b_qkv; g.out = ws.qkv; g.ldo = 3 * Ea;
  g.rope_cos = rope_cos; g.rope_sin = rope_sin; g.T = T; g.E = Ea; g.q_scale = L->q_scale; g.chunked = qkv_chunked();
  g.lo_col_off = 3 * Ea;
  g.rope_ld = 32 * L->slots;
  int rc = launch_gemm(EPI_QKV_ROPE, am.xn, L->tm_qkv, am.qkv_out, g, st, T_QKV, split);
  if (rc) return rc;
  // attention (multihead_attention.py:357-394)
  rc = run_attention(ws.qkv, ws.ctx, attn_probs, attn_batch_stride, attn_flags, ws.as, B, T, H, st, split, contact,
                     L->slots);
  if (rc) return rc;
  // out_proj + residual (multihead_attention.py:395, modules.py:134)
  memset(&g, 0, sizeof g);
  g.M = M; g.N = E; g.K = Ea; g.bias = L->out_b; g.out = x; g.ldo = E;
  rc = launch_gemm(EPI_BIAS_RESIDUAL, am.ctx, L->tm_out, am.x_out, g, st, T_OUT, split);
  if (rc) return rc;
  // LN2 -> fp16 (modules.py:137)
  {
    ProfScope ps(T_LN2, st);
    e = split ? launch_layernorm<2>(x, L->ln2_w, L->ln2_b, ws.xn, M, E, L->eps, st)
              : launch_layernorm<1>(x, L->ln2_w, L->ln2_b, ws.xn, M, E, L->eps, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "layernorm2");
  // fc1 + GELU (modules.py:138)
  memset(&g, 0, sizeof g);
  g.M = M; g.N = F; g.K = E; g.bias = L->fc1_b; g.out = ws.h; g.ldo = F; g.lo_col_off = F;
  rc = launch_gemm(EPI_BIAS_GELU, am.xn, L->tm_fc1, am.h, g, st, T_FC1, split);
  if (rc) return rc;
  // fc2 + residual (modules.py:139-140)
  memset(&g, 0, sizeof g);
  g.M = M; g.N = E; g.K = F; g.bias = L->fc2_b; g.out = x; g.ldo = E;
  rc = launch_gemm(EPI_BIAS_RESIDUAL, am.h, L->tm_fc2, am.x_out, g, st, T_FC2, split);
  return rc;
}

int make_act_maps(ActMaps* am, const Workspace& ws, const float* x, int E, int H, int F, int M, int split = 0) {
  const uint64_t Ea = (uint64_t)64 * head_slots(E, H) * H, pf = split ? 2 : 1;  // fp32x3: activations are [rows, 2 * width]
  int rc = make_tmap_f16(&am->xn, ws.xn, M, pf * E, pf * E, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&am->ctx, ws.ctx, M, pf * Ea, pf * Ea, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&am->h, ws.h, M, pf * F, pf * F, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&am->qkv_out, ws.qkv, M, pf * 3 * Ea, pf * 3 * Ea, 128);
  if (!rc) rc = make_tmap_2d(&am->x_out, x, 4, M, E, E, 128);
  return rc;
}
}  // namespace

int esmb200_stack_forward(esmb200_layer* const* layers, int32_t n_layers, float* x, const uint8_t* pad_mask,
                          int32_t B, int32_t T, const float* rope_cos, const float* rope_sin,
                          float* const* repr_out, float* const* attn_out, int64_t attn_batch_stride,
                          int32_t attn_flags, const esmb200_contact_job* contact, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (!layers || n_layers <= 0 || !x || !rope_cos || !rope_sin || !workspace)
    return fail(ESMB200_EINVAL, "null argument");
  if (contact) {
    if (!attn_out) return fail(ESMB200_EINVAL, "a contact job needs attn_out for every layer");
    for (int i = 0; i < n_layers; ++i)
      if (!attn_out[i]) return fail(ESMB200_EINVAL, "a contact job needs attn_out for every layer");
    if (!contact->weights || !contact->acc || !contact->row_part || !contact->col_part || contact->lo < 0 ||
        contact->hi > T || contact->hi <= contact->lo)
      return fail(ESMB200_EINVAL, "bad contact job");
  }
  if (B <= 0 || T <= 0) return fail(ESMB200_EINVAL, "empty batch");
  if ((long long)B * T > 0x7fffffffLL / 8) return fail(ESMB200_EINVAL, "B*T too large for one call; split the batch");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int E = layers[0]->E, F = layers[0]->F, H = layers[0]->H;
  if (F <= 0) return fail(ESMB200_EINVAL, "attention-only layers belong to esmb200_axial_stack_forward");
  for (int i = 1; i < n_layers; ++i)
    if (layers[i]->E != E || layers[i]->F != F || layers[i]->H != H || layers[i]->split != layers[0]->split)
      return fail(ESMB200_EINVAL, "layers of one stack must share E, H, F and precision");
  const int split = layers[0]->split;
  Workspace ws;
  rc = carve_workspace(&ws, workspace, workspace_bytes, E, H, F, B, T, split);
  if (rc) return rc;
  ActMaps am;
  rc = make_act_maps(&am, ws, x, E, H, F, B * T, split);
  if (rc) return rc;
  rc = run_key_bits(pad_mask, ws.as, B, T, st);
  if (rc) return rc;
  const int nt128 = (T + 127) / 128;
  for (int i = 0; i < n_layers; ++i) {
    ContactLayer cl;
    if (contact) {
      const int S = contact->hi - contact->lo;
      const size_t part = (size_t)B * H * nt128 * S;
      cl.w = contact->weights + (size_t)i * H; cl.keep = contact->keep; cl.acc = contact->acc;
      cl.row_part = contact->row_part + (size_t)i * 4 * part; cl.col_part = contact->col_part + (size_t)i * 4 * part;
      cl.lo = contact->lo; cl.S = S;
    }
    rc = layer_forward_impl(layers[i], x, B, T, rope_cos, rope_sin, attn_out ? attn_out[i] : nullptr, attn_batch_stride,
                            attn_flags, ws, am, st, contact ? &cl : nullptr);
    if (rc) return rc;
    if (repr_out && repr_out[i])
      CK(cudaMemcpyAsync(repr_out[i], x, (size_t)B * T * E * 4, cudaMemcpyDeviceToDevice, st));
  }
  return ESMB200_OK;
}

int esmb200_layer_forward(esmb200_layer* layer, float* x, const uint8_t* pad_mask, int32_t B, int32_t T,
                          const float* rope_cos, const float* rope_sin, float* attn_probs, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (!layer) return fail(ESMB200_EINVAL, "null layer");
  float* attn_arr[1] = {attn_probs};
  esmb200_layer* arr[1] = {layer};
  return esmb200_stack_forward(arr, 1, x, pad_mask, B, T, rope_cos, rope_sin, nullptr, attn_probs ? attn_arr : nullptr,
                               0, 0, nullptr, workspace, workspace_bytes, stream);
}

int esmb200_embed_tokens(const int64_t* tokens, const float* table, float* x, int32_t B, int32_t T, int32_t E,
                         int32_t padding_idx, int32_t mask_idx, int32_t token_dropout, void* stream) {
  if (!tokens || !table || !x) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || B > 65535 || T <= 0 || E % 4 != 0) return fail(ESMB200_EINVAL, "bad shape");
  ProfScope ps(T_EMBED, static_cast<cudaStream_t>(stream));
  int chunks = (8 * num_sms() + B - 1) / B;  // >= 8 blocks per SM over the batch, at least 4 rows per block
  if (chunks > (T + 3) / 4) chunks = (T + 3) / 4;
  if (chunks < 1) chunks = 1;
  embed_tokens_kernel<<<dim3(chunks, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(tokens, table, x, T, E, padding_idx,
                                                                                     mask_idx, token_dropout);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

int esmb200_layernorm(const float* x, const float* weight, const float* bias, float* out, int32_t M, int32_t E,
                      float eps, void* stream) {
  if (!x || !weight || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  ProfScope ps(T_LN_F32, static_cast<cudaStream_t>(stream));
  cudaError_t e = launch_layernorm<0>(x, weight, bias, out, M, E, eps, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "layernorm");
  return ESMB200_OK;
}

int esmb200_layernorm_f16(const float* x, const float* weight, const float* bias, void* out, int32_t M, int32_t E,
                          float eps, void* stream) {
  if (!x || !weight || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  ProfScope ps(T_LN1, static_cast<cudaStream_t>(stream));
  cudaError_t e = launch_layernorm<1>(x, weight, bias, out, M, E, eps, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "layernorm_f16");
  return ESMB200_OK;
}

int esmb200_gemm_f16(int32_t epilogue, const void* a, const void* w, const float* bias, void* out, int32_t M,
                     int32_t N, int32_t K, const float* rope_cos, const float* rope_sin, int32_t T, int32_t E,
                     void* stream) {
  if (!a || !w || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  const bool f16_out = (epilogue == EPI_QKV_ROPE || epilogue == EPI_BIAS_GELU);
  if (M <= 0 || N <= 0 || K <= 0 || K % 8 != 0 || N % (f16_out ? 64 : 32) != 0)
    return fail(ESMB200_EINVAL, "gemm needs K % 8 == 0 and N % 64 == 0 (fp16 output) / N % 32 == 0 (fp32 output)");
  int rc = check_device();
  if (rc) return rc;
  if (epilogue == EPI_QKV_ROPE && (!rope_cos || !rope_sin || T <= 0 || E <= 0 || E % 64 != 0 || N != 3 * E))
    return fail(ESMB200_EINVAL, "qkv epilogue needs rope tables, T and N == 3E");
  CUtensorMap ta, tb, tout;
  const bool out_f16 = (epilogue == EPI_QKV_ROPE || epilogue == EPI_BIAS_GELU || epilogue == EPI_F16_STOREONLY ||
                        epilogue == EPI_GELU_MATHONLY || epilogue == EPI_FMA_MATHONLY);
  rc = make_tmap_f16(&ta, a, M, K, K, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&tb, w, N, K, K, gemm2_cfg::HALF_N);
  if (!rc) rc = make_tmap_2d(&tout, out, out_f16 ? 2 : 4, M, N, N, 128);
  if (rc) return rc;
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = N;
  g.rope_cos = rope_cos; g.rope_sin = rope_sin; g.T = T; g.E = E; g.q_scale = 0.125f;
  return launch_gemm(epilogue, ta, tb, tout, g, static_cast<cudaStream_t>(stream));
}

// ---- fp32x3 precision building blocks (hi | lo fp16 operands): used by the LM head and the kernel-level parity tests
int esmb200_layernorm_split(const float* x, const float* weight, const float* bias, void* out, int32_t M, int32_t E,
                            float eps, void* stream) {
  if (!x || !weight || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  ProfScope ps(T_LN1, static_cast<cudaStream_t>(stream));
  cudaError_t e = launch_layernorm<2>(x, weight, bias, out, M, E, eps, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "layernorm_split");
  return ESMB200_OK;
}

int esmb200_convert_split(const float* src, void* dst, int64_t rows, int32_t K, void* stream) {
  if (!src || !dst) return fail(ESMB200_EINVAL, "null argument");
  if (rows <= 0 || K <= 0) return ESMB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t blocks = ((size_t)rows * K + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps(T_CONVERT, st);
  convert_f32_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, static_cast<__half*>(dst), (size_t)rows, K);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

int esmb200_gemm_split(int32_t epilogue, const void* a, const void* w, const float* bias, void* out, int32_t M,
                       int32_t N, int32_t K, const float* rope_cos, const float* rope_sin, int32_t T, int32_t E,
                       void* stream) {
  if (!a || !w || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  const bool f16_out = (epilogue == EPI_QKV_ROPE || epilogue == EPI_BIAS_GELU);
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || N % (f16_out ? 64 : 32) != 0)
    return fail(ESMB200_EINVAL, "split gemm needs K % 64 == 0 and N % 64 == 0 (fp16 output) / N % 32 == 0 (fp32 output)");
  if (epilogue < 0 || epilogue > EPI_BIAS_GELU_F32) return fail(ESMB200_EINVAL, "unknown GEMM epilogue");
  int rc = check_device();
  if (rc) return rc;
  if (epilogue == EPI_QKV_ROPE && (!rope_cos || !rope_sin || T <= 0 || E <= 0 || E % 64 != 0 || N != 3 * E))
    return fail(ESMB200_EINVAL, "qkv epilogue needs rope tables, T and N == 3E");
  CUtensorMap ta, tb, tout;
  rc = make_tmap_f16(&ta, a, M, 2 * (uint64_t)K, 2 * (uint64_t)K, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&tb, w, N, 2 * (uint64_t)K, 2 * (uint64_t)K, gemm2_cfg::HALF_N);
  if (!rc) rc = f16_out ? make_tmap_2d(&tout, out, 2, M, 2 * (uint64_t)N, 2 * (uint64_t)N, 128)
                        : make_tmap_2d(&tout, out, 4, M, N, N, 128);
  if (rc) return rc;
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = N; g.lo_col_off = N;
  g.rope_cos = rope_cos; g.rope_sin = rope_sin; g.T = T; g.E = E; g.q_scale = 0.125f;
  return launch_gemm(epilogue, ta, tb, tout, g, static_cast<cudaStream_t>(stream), T_GEMM_OTHER, true);
}

int esmb200_attention_split(const void* qkv, const uint8_t* pad_mask, void* ctx, float* attn_probs, int32_t B, int32_t T,
                            int32_t H, void* scratch, void* stream) {
  if (!qkv || !ctx || !scratch) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || H <= 0 || H > 64) return fail(ESMB200_EINVAL, "bad shape");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AttnScratch s = carve_attn_scratch(scratch, B, T, H);
  rc = run_key_bits(pad_mask, s, B, T, st);
  if (rc) return rc;
  return run_attention(qkv, ctx, attn_probs, 0, 0, s, B, T, H, st, true);
}

int esmb200_gemm_qkv_f16(const void* a, const void* w, const float* bias, void* out, int32_t M, int32_t E, float q_scale,
                         const float* rope_cos, const float* rope_sin, int32_t T, void* stream) {
  if (!a || !w || !bias || !out) return fail(ESMB200_EINVAL, "null argument");
  if (M <= 0 || E <= 0 || E % 64 != 0) return fail(ESMB200_EINVAL, "qkv gemm needs E % 64 == 0");
  if ((rope_cos == nullptr) != (rope_sin == nullptr) || (rope_cos && T <= 0))
    return fail(ESMB200_EINVAL, "rope tables must be given together with T, or not at all");
  int rc = check_device();
  if (rc) return rc;
  CUtensorMap ta, tb, tout;
  rc = make_tmap_f16(&ta, a, M, E, E, gemm2_cfg::BOX_M);
  if (!rc) rc = make_tmap_f16(&tb, w, 3 * (uint64_t)E, E, E, gemm2_cfg::HALF_N);
  if (!rc) rc = make_tmap_2d(&tout, out, 2, M, 3 * (uint64_t)E, 3 * (uint64_t)E, 128);
  if (rc) return rc;
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = 3 * E; g.K = E; g.bias = bias; g.out = out; g.ldo = 3 * E;
  g.rope_cos = rope_cos; g.rope_sin = rope_sin; g.T = T > 0 ? T : 1; g.E = E; g.q_scale = q_scale;
  return launch_gemm(EPI_QKV_ROPE, ta, tb, tout, g, static_cast<cudaStream_t>(stream));
}

int esmb200_attention(const void* qkv, const uint8_t* pad_mask, void* ctx, float* attn_probs, int32_t B, int32_t T,
                      int32_t H, void* scratch, void* stream) {
  if (!qkv || !ctx || !scratch) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || H <= 0 || H > 64) return fail(ESMB200_EINVAL, "bad shape");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AttnScratch s = carve_attn_scratch(scratch, B, T, H);
  rc = run_key_bits(pad_mask, s, B, T, st);
  if (rc) return rc;
  return run_attention(qkv, ctx, attn_probs, 0, 0, s, B, T, H, st);
}

int esmb200_attention128(const void* qkv, const uint8_t* pad_mask, void* ctx, float* attn_probs, int32_t B, int32_t T,
                         int32_t H, void* scratch, void* stream) {
  if (!qkv || !ctx || !scratch) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || H <= 0 || H > 64) return fail(ESMB200_EINVAL, "bad shape");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AttnScratch s = carve_attn_scratch(scratch, B, T, H);
  rc = run_key_bits(pad_mask, s, B, T, st);
  if (rc) return rc;
  return run_attention(qkv, ctx, attn_probs, 0, 0, s, B, T, H, st, false, nullptr, 2);
}

// ---- MSA axial attention (esm/axial_attention.py) -----------------------------------------------------------------
size_t esmb200_tied_row_attention_scratch_bytes(int32_t B, int32_t C, int32_t H) {
  const size_t Cp = align_up((size_t)C, 64);
  return align_up((size_t)H * B * C * C * 4, 1024) + align_up((size_t)H * B * C * Cp * 2, 1024) + 2048;
}

static int tied_row_impl(const void* qkv, const uint8_t* key_pad, long long key_pad_stride, void* ctx,
                         float* attn_probs, int32_t B, int32_t R, int32_t C, int32_t H, void* scratch,
                         size_t scratch_bytes, void* stream) {
  if (!qkv || !ctx || !scratch) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || R <= 0 || C <= 0 || H <= 0 || H > 64 || (long long)B * H > 65535)
    return fail(ESMB200_EINVAL, "bad shape");
  if (C > TIED_MAX_C) return fail(ESMB200_EINVAL, "tied row attention supports at most 1024 alignment columns");
  if (scratch_bytes < esmb200_tied_row_attention_scratch_bytes(B, C, H))
    return fail(ESMB200_EWORKSPACE, "tied row attention scratch too small");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int E = H * 64;
  const size_t Cp = align_up((size_t)C, 64);
  uint8_t* sp = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(scratch), 1024));
  TiedParams tp;
  tp.B = B; tp.R = R; tp.C = C; tp.H = H; tp.E = E; tp.Cp = (int)Cp;
  tp.S = attn_probs ? attn_probs : reinterpret_cast<float*>(sp);
  tp.P = reinterpret_cast<__half*>(sp + align_up((size_t)H * B * C * C * 4, 1024));
  tp.ctx = static_cast<__half*>(ctx);
  tp.key_pad = key_pad;
  tp.key_pad_stride = key_pad_stride;
  tp.write_probs = attn_probs ? 1 : 0;
  const uint64_t rows = (uint64_t)B * R * C;
  CUtensorMap tq, tk, tv, tpm;
  if ((rc = make_tmap_f16(&tq, qkv, rows, (uint64_t)3 * E, (uint64_t)3 * E, tied_cfg::S_BM))) return rc;
  if ((rc = make_tmap_f16(&tk, qkv, rows, (uint64_t)3 * E, (uint64_t)3 * E, tied_cfg::S_BN))) return rc;
  if ((rc = make_tmap_f16(&tv, qkv, rows, (uint64_t)3 * E, (uint64_t)3 * E, 64))) return rc;
  if ((rc = make_tmap_f16(&tpm, tp.P, (uint64_t)H * B * C, Cp, Cp, tied_cfg::V_BM))) return rc;
  cudaError_t e;
  {
    ProfScope ps(T_TIED_SCORES, st);
    e = launch_tied_scores(tq, tk, tp, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "tied scores launch");
  {
    ProfScope ps(T_TIED_SOFTMAX, st);
    e = launch_tied_softmax(tp, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "tied softmax launch");
  {
    ProfScope ps(T_TIED_PV, st);
    e = launch_tied_pv(tpm, tv, tp, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "tied update launch");
  return ESMB200_OK;
}

int esmb200_tied_row_attention(const void* qkv, const uint8_t* key_pad, void* ctx, float* attn_probs, int32_t B,
                               int32_t R, int32_t C, int32_t H, void* scratch, size_t scratch_bytes, void* stream) {
  return tied_row_impl(qkv, key_pad, C, ctx, attn_probs, B, R, C, H, scratch, scratch_bytes, stream);
}

int esmb200_column_attention(const void* qkv, const uint8_t* pad_mask, void* ctx, int32_t B, int32_t R, int32_t C,
                             int32_t H, void* scratch, void* stream) {
  if (!qkv || !ctx || !scratch) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || R <= 0 || C <= 0 || H <= 0 || H > 64) return fail(ESMB200_EINVAL, "bad shape");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int E = H * 64;
  const int S = B * C;  // one "sequence" of R tokens per alignment column
  AttnScratch s = carve_attn_scratch(scratch, S, R, H);
  rc = run_key_bits(pad_mask, s, S, R, st);
  if (rc) return rc;
  CUtensorMap tq, tkv;
  const uint64_t wide = (uint64_t)C * 3 * E;  // qkv viewed as [B*R, C*3E]: token r of column c at row r, x = c*3E
  if ((rc = make_tmap_f16(&tq, qkv, (uint64_t)B * R, wide, wide, attn8_cfg::BLOCK_Q))) return rc;
  if ((rc = make_tmap_f16(&tkv, qkv, (uint64_t)B * R, wide, wide, attn8_cfg::BLOCK_KV))) return rc;
  AttnParams ap;
  ap.B = S; ap.T = R; ap.H = H; ap.E = E;
  ap.keybits = s.keybits; ap.kvlen = s.kvlen; ap.words = s.words;
  ap.ctx = static_cast<__half*>(ctx);
  ap.row_max = nullptr; ap.row_sum = nullptr;
  ap.cols = C;
  cudaError_t e;
  {
    ProfScope ps(T_ATTN, st);
    e = launch_attention_fwd(tq, tkv, ap, num_sms(), st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "column attention launch");
  return ESMB200_OK;
}


size_t esmb200_axial_workspace_bytes(int32_t E, int32_t F, int32_t B, int32_t R, int32_t C) {
  return esmb200_workspace_bytes(E, E / 64, F, B * C, R, 0) + esmb200_tied_row_attention_scratch_bytes(B, C, E / 64) + 1024;
}

// (A CUDA-graph replay of this launch sequence was measured: 20.70 vs 20.77 ms per 128 x 512 MSA — the ~2 ms between the
// sum of the kernel times and the wall time are not host launch overhead, so the calls stay plain stream launches.)
int esmb200_axial_stack_forward(esmb200_layer* const* row_layers, esmb200_layer* const* col_layers, int32_t n_layers,
                                float* x, const uint8_t* pad_mask, const uint8_t* col_pad_mask, int32_t B, int32_t R,
                                int32_t C, float* const* row_attn_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (!row_layers || !col_layers || n_layers <= 0 || !x || !workspace) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || R <= 0 || C <= 0) return fail(ESMB200_EINVAL, "empty alignment");
  if ((pad_mask == nullptr) != (col_pad_mask == nullptr))
    return fail(ESMB200_EINVAL, "pad_mask [B,R,C] and col_pad_mask [B,C,R] must be given together");
  if ((long long)B * R * C > 0x7fffffffLL / 8) return fail(ESMB200_EINVAL, "B*R*C too large for one call");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int E = col_layers[0]->E, F = col_layers[0]->F, H = col_layers[0]->H;
  if (F <= 0) return fail(ESMB200_EINVAL, "col_layers carry the feed-forward weights");
  for (int i = 0; i < n_layers; ++i)
    if (row_layers[i]->E != E || col_layers[i]->E != E || col_layers[i]->F != F)
      return fail(ESMB200_EINVAL, "layers of one stack must share E and F");
  if (workspace_bytes < esmb200_axial_workspace_bytes(E, F, B, R, C))
    return fail(ESMB200_EWORKSPACE, "workspace too small");
  const int M = B * R * C;
  Workspace ws;
  if (E != 64 * H) return fail(ESMB200_EINVAL, "the MSA axial path needs head_dim 64");
  for (int i = 0; i < n_layers; ++i)
    if (row_layers[i]->split || col_layers[i]->split)
      return fail(ESMB200_EINVAL, "the MSA axial path runs with fp16 operands only (precision 0)");
  const size_t base_bytes = esmb200_workspace_bytes(E, H, F, B * C, R, 0);
  rc = carve_workspace(&ws, workspace, base_bytes, E, H, F, B * C, R, 0);
  if (rc) return rc;
  uint8_t* tied_scratch = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024)) + base_bytes;
  const size_t tied_bytes = esmb200_tied_row_attention_scratch_bytes(B, C, H);
  ActMaps am;
  rc = make_act_maps(&am, ws, x, E, H, F, M);
  if (rc) return rc;
  rc = run_key_bits(col_pad_mask, ws.as, B * C, R, st);  // column attention: B*C sequences of R keys
  if (rc) return rc;
  CUtensorMap tcq, tckv;
  const uint64_t wide = (uint64_t)C * 3 * E;
  if ((rc = make_tmap_f16(&tcq, ws.qkv, (uint64_t)B * R, wide, wide, attn8_cfg::BLOCK_Q))) return rc;
  if ((rc = make_tmap_f16(&tckv, ws.qkv, (uint64_t)B * R, wide, wide, attn8_cfg::BLOCK_KV))) return rc;
  const float row_scale = 0.125f / sqrtf((float)R);  // axial_attention.py:36-38
  cudaError_t e;
  GemmParams g;
  for (int i = 0; i < n_layers; ++i) {
    // ---------------- tied row attention (modules.py:202-207; axial_attention.py:71-130) ----------------
    esmb200_layer* L = row_layers[i];
    {
      ProfScope ps(T_LN1, st);
      e = launch_layernorm<1>(x, L->ln1_w, L->ln1_b, ws.xn, M, E, L->eps, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "row layernorm");
    memset(&g, 0, sizeof g);
    g.M = M; g.N = 3 * E; g.K = E; g.bias = L->b_qkv; g.out = ws.qkv; g.ldo = 3 * E;
    g.T = 1; g.E = E; g.q_scale = row_scale;
    rc = launch_gemm(EPI_QKV_ROPE, am.xn, L->tm_qkv, am.qkv_out, g, st, T_QKV);
    if (rc) return rc;
    if (pad_mask) {
      ProfScope ps(T_KEYBITS, st);
      zero_q_at_pads_kernel<<<(M + 7) / 8, 256, 0, st>>>(ws.qkv, pad_mask, M, E);
      CK(cudaGetLastError());
    }
    rc = tied_row_impl(ws.qkv, pad_mask, (long long)R * C, ws.ctx, row_attn_out ? row_attn_out[i] : nullptr, B, R, C, H,
                       tied_scratch, tied_bytes, stream);
    if (rc) return rc;
    memset(&g, 0, sizeof g);
    g.M = M; g.N = E; g.K = E; g.bias = L->out_b; g.out = x; g.ldo = E;
    rc = launch_gemm(EPI_BIAS_RESIDUAL, am.ctx, L->tm_out, am.x_out, g, st, T_OUT);
    if (rc) return rc;
    // ---------------- column attention (modules.py:208-212; axial_attention.py:182-239) ----------------
    L = col_layers[i];
    {
      ProfScope ps(T_LN1, st);
      e = launch_layernorm<1>(x, L->ln1_w, L->ln1_b, ws.xn, M, E, L->eps, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "column layernorm");
    memset(&g, 0, sizeof g);
    g.M = M; g.N = 3 * E; g.K = E; g.bias = L->b_qkv; g.out = ws.qkv; g.ldo = 3 * E;
    g.T = 1; g.E = E; g.q_scale = 0.125f;
    rc = launch_gemm(EPI_QKV_ROPE, am.xn, L->tm_qkv, am.qkv_out, g, st, T_QKV);
    if (rc) return rc;
    {
      AttnParams ap;
      ap.B = B * C; ap.T = R; ap.H = H; ap.E = E;
      ap.keybits = ws.as.keybits; ap.kvlen = ws.as.kvlen; ap.words = ws.as.words;
      ap.ctx = ws.ctx; ap.row_max = nullptr; ap.row_sum = nullptr; ap.cols = C;
      ProfScope ps(T_ATTN, st);
      e = launch_attention_fwd(tcq, tckv, ap, num_sms(), st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "column attention launch");
    memset(&g, 0, sizeof g);
    g.M = M; g.N = E; g.K = E; g.bias = L->out_b; g.out = x; g.ldo = E;
    rc = launch_gemm(EPI_BIAS_RESIDUAL, am.ctx, L->tm_out, am.x_out, g, st, T_OUT);
    if (rc) return rc;
    // ---------------- feed-forward (modules.py:213-214, 413-418) ----------------
    {
      ProfScope ps(T_LN2, st);
      e = launch_layernorm<1>(x, L->ln2_w, L->ln2_b, ws.xn, M, E, L->eps, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "ffn layernorm");
    memset(&g, 0, sizeof g);
    g.M = M; g.N = F; g.K = E; g.bias = L->fc1_b; g.out = ws.h; g.ldo = F;
    rc = launch_gemm(EPI_BIAS_GELU, am.xn, L->tm_fc1, am.h, g, st, T_FC1);
    if (rc) return rc;
    memset(&g, 0, sizeof g);
    g.M = M; g.N = E; g.K = F; g.bias = L->fc2_b; g.out = x; g.ldo = E;
    rc = launch_gemm(EPI_BIAS_RESIDUAL, am.h, L->tm_fc2, am.x_out, g, st, T_FC2);
    if (rc) return rc;
  }
  return ESMB200_OK;
}



int esmb200_msa_embed(const int64_t* tokens, const float* embed_table, const float* pos_table, const float* msa_pos,
                      int32_t msa_pos_dim, const float* ln_weight, const float* ln_bias, float eps, float* x,
                      int32_t B, int32_t R, int32_t C, int32_t E, int32_t padding_idx, void* stream) {
  if (!tokens || !embed_table || !pos_table || !ln_weight || !ln_bias || !x) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || R <= 0 || C <= 0 || E <= 0 || E % 4 != 0 || E > 20 * 128) return fail(ESMB200_EINVAL, "bad shape");
  if (msa_pos && msa_pos_dim != E && msa_pos_dim != 1)
    return fail(ESMB200_EINVAL, "msa_position_embedding width must be E or 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(T_EMBED, st);
  const size_t smem = (size_t)C * sizeof(int);
  const int grid = B * R;
  if (E <= 4 * 128)
    msa_embed_kernel<4><<<grid, 256, smem, st>>>(tokens, embed_table, pos_table, msa_pos, msa_pos_dim, ln_weight,
                                                  ln_bias, eps, x, R, C, E, padding_idx);
  else if (E <= 10 * 128)
    msa_embed_kernel<10><<<grid, 256, smem, st>>>(tokens, embed_table, pos_table, msa_pos, msa_pos_dim, ln_weight,
                                                   ln_bias, eps, x, R, C, E, padding_idx);
  else
    msa_embed_kernel<20><<<grid, 256, smem, st>>>(tokens, embed_table, pos_table, msa_pos, msa_pos_dim, ln_weight,
                                                   ln_bias, eps, x, R, C, E, padding_idx);
  CK(cudaGetLastError());
  return ESMB200_OK;
}


int esmb200_contact_accumulate(const float* attn, int64_t batch_stride, const float* w, const uint8_t* keep, float* acc,
                               float* row_sum, float* col_part, int32_t B, int32_t H, int32_t T, int32_t lo, int32_t hi,
                               void* stream) {
  if (!attn || !w || !acc || !row_sum || !col_part) return fail(ESMB200_EINVAL, "null argument");
  const int S = hi - lo;
  if (B <= 0 || H <= 0 || T <= 0 || lo < 0 || hi > T || S <= 0 || B > 65535) return fail(ESMB200_EINVAL, "bad shape");
  if (S > 1024) return fail(ESMB200_EINVAL, "contact head supports at most 1024 positions");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(T_PROBS, st);
  const size_t smem = (size_t)8 * S * sizeof(float);
  dim3 grid((S + 15) / 16, B);  // 8 warps x 2 rows
  if (S <= 512)
    contact_accumulate_kernel<2, 16, 2><<<grid, 256, smem, st>>>(attn, batch_stride, w, keep, acc, row_sum, col_part, H, T,
                                                                  lo, S);
  else
    contact_accumulate_kernel<2, 32, 1><<<grid, 256, smem, st>>>(attn, batch_stride, w, keep, acc, row_sum, col_part, H, T,
                                                                  lo, S);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

int esmb200_contact_finalize(const float* acc, const float* u, const float* a1, const float* bias, float* out, int32_t B,
                             int32_t C, int32_t S, void* stream) {
  if (!acc || !u || !a1 || !out) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || C <= 0 || S <= 0 || B > 65535) return fail(ESMB200_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(T_PROBS, st);
  dim3 grid((S + 63) / 64, (S + 63) / 64, B);
  contact_finalize_kernel<<<grid, 256, 0, st>>>(acc, u, a1, bias, out, C, S);
  CK(cudaGetLastError());
  return ESMB200_OK;
}


int esmb200_mean_pool(const float* x, const int32_t* lengths, float* out, int32_t B, int32_t T, int32_t E,
                      void* stream) {
  if (!x || !lengths || !out) return fail(ESMB200_EINVAL, "null argument");
  if (B <= 0 || T < 2 || E <= 0 || B > 65535) return fail(ESMB200_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(T_MEANPOOL, st);
  if (E % 4 != 0) return fail(ESMB200_EINVAL, "mean_pool needs E % 4 == 0");
  dim3 grid((E + 127) / 128, B);
  mean_pool_kernel<<<grid, 256, 0, st>>>(x, lengths, out, T, E);
  CK(cudaGetLastError());
  return ESMB200_OK;
}

#ifdef ESMB200_TRACE
int esmb200_debug_read_attn_trace(long long* out, int32_t n) {
  CK(cudaMemcpyFromSymbol(out, esmb200::g_attn_trace, sizeof(long long) * (size_t)n));
  return ESMB200_OK;
}
#endif

int esmb200_set_option(const char* name, int32_t value) {
  if (!name) return fail(ESMB200_EINVAL, "null option name");
#ifdef ESMB200_EXPERIMENTS
  if (!strcmp(name, "attn") && (value == 7 || value == 8)) { g_attn_version = value; return ESMB200_OK; }
#else
  if (!strcmp(name, "attn") && value == 8) { g_attn_version = value; return ESMB200_OK; }
#endif
  if (!strcmp(name, "attn_poly") && (value == 0 || value == 2 || value == 3 || value == 4)) {
    g_attn_poly = value;
    return ESMB200_OK;
  }
  if (!strcmp(name, "pdl") && (value == 0 || value == 1)) { pdl_flag() = value; return ESMB200_OK; }
  return fail(ESMB200_EINVAL, std::string("unknown option or value: ") + name);
}

long long esmb200_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  return g_prof.launches;
}

int esmb200_profile_enable(int32_t max_launches) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.tag.clear();
  g_prof.used = 0;
  g_prof.on = max_launches > 0;
  for (int i = 0; i < 2 * max_launches; ++i) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    g_prof.ev.push_back(e);
  }
  return ESMB200_OK;
}

int esmb200_profile_read(int32_t* tags, float* ms, int32_t max_records) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  const int n = (int)(g_prof.used / 2);
  int out = 0;
  for (int i = 0; i < n && out < max_records; ++i, ++out) {
    CK(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    tags[out] = g_prof.tag[i];
    ms[out] = t;
  }
  g_prof.used = 0;
  g_prof.tag.clear();
  return out;
}

}  // extern "C"
