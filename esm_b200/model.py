"""Host-side mirror of the reference's ESM-2 interface, dispatching the transformer-layer path to libesmb200.so.

Mirrors (same names, argument meaning, state-dict keys and result dict):
    esm.modules.TransformerLayer.forward       /root/reference/esm/modules.py:120-142
    esm.model.esm2.ESM2.__init__ / forward     /root/reference/esm/model/esm2.py:15-144
PyTorch owns parameters, activations and the workspace (torch tensors) and provides the CUDA stream; every layer's
compute goes through the C ABI in include/esmb200.h.  There is no CPU or eager fallback: on a non-CUDA tensor, or
when libesmb200.so is missing, the forward raises.

Documented deviations from the reference:
  * `TransformerLayer.forward` returns `attn=None` unless `need_head_weights=True` (the reference always computes a
    head-averaged (B,T,T) map that ESM2.forward discards, modules.py:130 / esm2.py:112-121).
  * MMA operands are fp16 (fp32 accumulate, fp32 residual stream / LayerNorm / softmax); tolerance in DESIGN.md.
  * head_dim <= 128 (every ESM-2 checkpoint); heads other than 64 wide run in zero-padded 64-wide slots (two per head
    above 64: 15B).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .alphabet import Alphabet


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class RotaryEmbedding(nn.Module):
    """Holds the `inv_freq` buffer under the reference's key (rotary_embedding.py:37-41); tables are built by
    ESM2._rope_tables with the same torch ops as rotary_embedding.py:47-61."""

    def __init__(self, dim: int):
        super().__init__()
        inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)


class MultiheadAttention(nn.Module):
    """Parameter container with the reference's names (multihead_attention.py:109-113,130-132)."""

    def __init__(self, embed_dim: int, num_heads: int):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.rot_emb = RotaryEmbedding(self.head_dim)


def rope_tables(inv_freq: torch.Tensor, seq_len: int):
    """cos/sin [T, 32] fp32 (head_dim <= 64) or [T, 64] (head_dim <= 128) — rotary_embedding.py:53-59 (the reference's
    table is the first d/2 columns duplicated on the last dim); columns >= d/2 are padding the kernels never use."""
    inv_freq = inv_freq.float()
    t = torch.arange(seq_len, device=inv_freq.device).type_as(inv_freq)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    cos, sin = freqs.cos(), freqs.sin()
    width = 32 if cos.shape[1] <= 32 else 64
    if cos.shape[1] < width:
        pad = width - cos.shape[1]
        cos, sin = F.pad(cos, (0, pad), value=1.0), F.pad(sin, (0, pad), value=0.0)
    return cos.contiguous(), sin.contiguous()


def _f32(p: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view of a parameter (the tensor itself when it already is one)."""
    t = p.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


class LayerBinding:
    """The esmb200_layer handle of one transformer block, built from a module that carries the reference's attribute
    names (self_attn.{q,k,v,out}_proj, self_attn_layer_norm, fc1, fc2, final_layer_norm: modules.py:99-118) — this
    repo's TransformerLayer or the reference's own esm.modules.TransformerLayer (esm_b200.integration).  Parameters that
    are not fp32 (model.half(), esmfold.py:59-62) are mirrored to fp32 copies owned by the binding; the handle is
    re-packed when a parameter is replaced or modified in place."""

    def __init__(self, module: nn.Module):
        self.module = module
        a = module.self_attn
        self.embed_dim = a.q_proj.weight.shape[1]
        self.attention_heads = a.num_heads
        self.head_dim = self.embed_dim // self.attention_heads
        self.ffn_embed_dim = module.fc1.weight.shape[0]
        if self.head_dim * self.attention_heads != self.embed_dim or self.head_dim > 128 or self.head_dim % 2:
            raise ValueError("esm_b200 supports even head_dim <= 128 (every ESM-2 model); "
                             f"got embed_dim={self.embed_dim}, heads={self.attention_heads}")
        self.precision = 0  # 0 = fp16 MMA operands, 1 = "fp32x3" (esmb200.h: esmb200_layer_weights.precision)
        self._handle = None
        self._key = None
        self._keep = None

    def _params(self) -> List[torch.Tensor]:
        m, a = self.module, self.module.self_attn
        return [m.self_attn_layer_norm.weight, m.self_attn_layer_norm.bias, a.q_proj.weight, a.q_proj.bias,
                a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias, a.out_proj.weight, a.out_proj.bias,
                m.final_layer_norm.weight, m.final_layer_norm.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]

    def handle(self):
        ps = self._params()
        key = (self.precision,) + tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        if self._handle is not None and key == self._key:
            return self._handle
        self.release()
        for p in ps:
            if not p.is_cuda:
                raise _lib.Esmb200Error("esm_b200 runs on CUDA (sm_100a) only: move the model with .cuda(); "
                                        "there is no CPU fallback")
        lib = _lib.load()
        keep = [_f32(p) for p in ps]
        w = _lib.LayerWeights()
        w.embed_dim, w.num_heads, w.ffn_dim = self.embed_dim, self.attention_heads, self.ffn_embed_dim
        w.head_dim = self.head_dim
        w.precision = self.precision
        w.ln_eps = self.module.self_attn_layer_norm.eps
        names = [f[0] for f in _lib.LayerWeights._fields_[4:20]]
        for n, p in zip(names, keep):
            setattr(w, n, p.data_ptr())
        out = ctypes.c_void_p()
        with torch.cuda.device(keep[0].device):
            _lib.check(lib.esmb200_layer_create(ctypes.byref(w), _stream(), ctypes.byref(out)))
        self._handle, self._key, self._keep = out, key, keep  # the library borrows LN weights and biases from `keep`
        return out

    def release(self):
        if self._handle is not None:
            _lib.load().esmb200_layer_destroy(self._handle)
            self._handle, self._key, self._keep = None, None, None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class TransformerLayer(nn.Module):
    """Pre-LN transformer block of ESM-2 (modules.py:84-142) executed by libesmb200.so."""

    def __init__(self, embed_dim: int, ffn_embed_dim: int, attention_heads: int):
        super().__init__()
        self.embed_dim = embed_dim
        self.ffn_embed_dim = ffn_embed_dim
        self.attention_heads = attention_heads
        self.self_attn = MultiheadAttention(embed_dim, attention_heads)
        self.self_attn_layer_norm = nn.LayerNorm(embed_dim)
        self.fc1 = nn.Linear(embed_dim, ffn_embed_dim)
        self.fc2 = nn.Linear(ffn_embed_dim, embed_dim)
        self.final_layer_norm = nn.LayerNorm(embed_dim)
        self._binding = LayerBinding(self)

    def handle(self):
        """esmb200_layer* for the current parameters."""
        return self._binding.handle()

    def release(self):
        self._binding.release()

    @property
    def precision(self) -> int:
        return self._binding.precision

    @precision.setter
    def precision(self, value: int):
        self._binding.precision = int(value)

    # ---- reference-facing forward ------------------------------------------------------------------------------
    def forward(self, x, self_attn_mask=None, self_attn_padding_mask=None, need_head_weights=False):
        """x: (T, B, E) like the reference (modules.py:120-122). Returns (x (T,B,E), attn (H,B,T,T) or None)."""
        return layer_forward(self._binding, x, self_attn_mask, self_attn_padding_mask, need_head_weights)


def layer_forward(binding: LayerBinding, x, self_attn_mask=None, self_attn_padding_mask=None, need_head_weights=False):
    """TransformerLayer.forward (modules.py:120-142) through the C ABI, for any module a LayerBinding wraps."""
    if self_attn_mask is not None:
        raise NotImplementedError("ESM-2 never passes self_attn_mask (esm2.py:112-116)")
    T, B, E = x.shape
    # (B,T,E) batch-major PRIVATE copy, updated in place.  Layer 0 of the reference receives a transposed view of the
    # tensor it also returns as representations[0] (esm2.py:99-106): contiguous()/float() would hand that storage back.
    xb = torch.empty((B, T, E), dtype=torch.float32, device=x.device)
    xb.copy_(x.transpose(0, 1))
    cos, sin = rope_tables(binding.module.self_attn.rot_emb.inv_freq, T)
    attn = run_stack([binding], xb, self_attn_padding_mask, cos, sin, None, [0] if need_head_weights else [])
    out = xb.transpose(0, 1).to(x.dtype)
    if need_head_weights:
        return out, attn[0].transpose(0, 1).contiguous().to(x.dtype)  # (B,H,T,T) -> (H,B,T,T), multihead_attention.py:398-400
    return out, None


_workspaces: Dict[tuple, torch.Tensor] = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Scratch for one stack call, cached per (device, CUDA stream): calls on different streams never share it, calls
    on one stream are ordered by the stream (the library is re-entrant across handles and streams)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def run_stack(layers: Sequence, x: torch.Tensor, padding_mask: Optional[torch.Tensor],
              rope_cos: torch.Tensor, rope_sin: torch.Tensor, repr_out: Optional[Dict[int, torch.Tensor]],
              attn_layers: Sequence[int], zero_pad_rows: bool = False, contact_job=None):
    """esmb200_stack_forward on x fp32 (B,T,E) in place. repr_out: {layer index (0-based): (B,T,E) tensor to fill}.
    Returns {layer index: (B,H,T,T) fp32} for the indices in attn_layers."""
    if not x.is_cuda:
        raise _lib.Esmb200Error("esm_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    assert x.dtype == torch.float32 and x.is_contiguous()
    lib = _lib.load()
    B, T, E = x.shape
    n = len(layers)
    Fdim, H = layers[0].ffn_embed_dim, layers[0].attention_heads
    with torch.cuda.device(x.device):
        handles = (ctypes.c_void_p * n)(*[l.handle() for l in layers])
        nbytes = lib.esmb200_workspace_bytes(E, H, Fdim, B, T, getattr(layers[0], "precision", 0))
        ws = _workspace(nbytes, x.device)
        mask = None
        if padding_mask is not None:
            mask = padding_mask.to(device=x.device, dtype=torch.uint8).contiguous()
            assert mask.shape == (B, T)
        reprs = (ctypes.c_void_p * n)()
        keep = []
        if repr_out:
            for i, t in repr_out.items():
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == x.shape
                reprs[i] = t.data_ptr()
        attns = (ctypes.c_void_p * n)()
        attn_t = {}
        stacked = None
        if attn_layers:
            # one [B, n_attn, H, T, T] allocation, layer i written straight into its slice (esm2.py:134's stack)
            stacked = torch.empty((B, len(attn_layers), H, T, T), dtype=torch.float32, device=x.device)
            for pos, i in enumerate(attn_layers):
                a = stacked[:, pos]
                attn_t[i] = a
                attns[i] = a.data_ptr()
        keep.append(mask)
        _lib.check(lib.esmb200_stack_forward(handles, n, _ptr(x), _ptr(mask), B, T, _ptr(rope_cos), _ptr(rope_sin),
                                             reprs if repr_out else None, attns if attn_layers else None,
                                             (len(attn_layers) * H * T * T) if attn_layers else 0,
                                             1 if zero_pad_rows else 0,
                                             ctypes.byref(contact_job) if contact_job is not None else None,
                                             _ptr(ws), ws.numel(), _stream()))
    if stacked is not None:
        attn_t["stacked"] = stacked
    return attn_t


def gelu(x):
    """modules.py:17-24"""
    return x * 0.5 * (1.0 + torch.erf(x / 1.4142135623730951))


class RobertaLMHead(nn.Module):
    """modules.py:298-314: dense -> gelu -> LayerNorm -> tied-embedding projection + bias.

    `forward(features)` is the plain PyTorch evaluation (used on already layer-normed features, e.g. by callers that
    hold a representation).  `forward_native(x_pre_ln, ln_w, ln_b, eps)` is what ESM2.forward uses on the GPU: it
    starts from the residual stream BEFORE emb_layer_norm_after and runs the whole tail through libesmb200.so
    (LayerNorm->fp16 | tcgen05 GEMM + bias + erf-GELU | LayerNorm->fp16 | tcgen05 GEMM onto the 33 tokens, padded to 64
    output columns), replacing ~25 ms of fp32 cuBLAS + elementwise passes per 256x1024-token batch by ~2 ms."""

    def __init__(self, embed_dim, output_dim, weight):
        super().__init__()
        self.dense = nn.Linear(embed_dim, embed_dim)
        self.layer_norm = nn.LayerNorm(embed_dim)
        self.weight = weight
        self.bias = nn.Parameter(torch.zeros(output_dim))
        self._packed = None
        self._packed_key = None
        self._packed_split = None

    def forward(self, features):
        x = self.dense(features)
        x = gelu(x)
        x = self.layer_norm(x)
        return F.linear(x, self.weight) + self.bias

    def _pack(self):
        ps = [self.dense.weight, self.weight, self.bias, self.dense.bias, self.layer_norm.weight, self.layer_norm.bias]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        if self._packed is None or key != self._packed_key:
            E = self.dense.weight.shape[1]
            V = self.weight.shape[0]
            npad = (V + 63) // 64 * 64  # GEMM N must be a multiple of 64
            w_out = torch.zeros((npad, E), dtype=torch.float16, device=self.weight.device)
            w_out[:V] = self.weight.detach().half()
            b_out = torch.zeros((npad,), dtype=torch.float32, device=self.weight.device)
            b_out[:V] = self.bias.detach().float()
            self._packed = (self.dense.weight.detach().half().contiguous(), w_out, b_out, V, npad,
                            _f32(self.dense.bias), _f32(self.layer_norm.weight), _f32(self.layer_norm.bias))
            self._packed_key = key
        return self._packed

    def _pack_split(self):
        """fp32x3 operands of the two GEMMs: fp16 hi | lo halves along K (esmb200_convert_split)."""
        ps = [self.dense.weight, self.weight]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        if self._packed_split is None or key != self._packed_split[0]:
            lib = _lib.load()
            E = self.dense.weight.shape[1]
            V = self.weight.shape[0]
            npad = (V + 63) // 64 * 64
            dev = self.weight.device
            wd32 = _f32(self.dense.weight)
            wo32 = torch.zeros((npad, E), dtype=torch.float32, device=dev)
            wo32[:V] = self.weight.detach().float()
            wd = torch.empty((E, 2 * E), dtype=torch.float16, device=dev)
            wo = torch.empty((npad, 2 * E), dtype=torch.float16, device=dev)
            _lib.check(lib.esmb200_convert_split(_ptr(wd32), _ptr(wd), E, E, _stream()))
            _lib.check(lib.esmb200_convert_split(_ptr(wo32), _ptr(wo), npad, E, _stream()))
            self._packed_split = (key, wd, wo)
        return self._packed_split[1], self._packed_split[2]

    def forward_native(self, x_pre: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor, eps: float,
                       precision: int = 0) -> torch.Tensor:
        """x_pre: fp32 [B,T,E] residual stream before emb_layer_norm_after (esm2.py:123). Returns logits [B,T,V]."""
        lib = _lib.load()
        B, T, E = x_pre.shape
        M = B * T
        dev = x_pre.device
        w_dense, w_out, b_out, V, npad, b_dense, ln2_w, ln2_b = self._pack()
        if precision:
            wd, wo = self._pack_split()
            a16 = torch.empty((M, 2 * E), dtype=torch.float16, device=dev)
            _lib.check(lib.esmb200_layernorm_split(_ptr(x_pre), _ptr(ln_w), _ptr(ln_b), _ptr(a16), M, E, eps, _stream()))
            h = torch.empty((M, E), dtype=torch.float32, device=dev)
            _lib.check(lib.esmb200_gemm_split(_lib.EPI_BIAS_GELU_F32, _ptr(a16), _ptr(wd), _ptr(b_dense), _ptr(h), M, E, E,
                                              None, None, 0, 0, _stream()))
            _lib.check(lib.esmb200_layernorm_split(_ptr(h), _ptr(ln2_w), _ptr(ln2_b), _ptr(a16), M, E,
                                                   self.layer_norm.eps, _stream()))
            logits = torch.empty((M, npad), dtype=torch.float32, device=dev)
            _lib.check(lib.esmb200_gemm_split(_lib.EPI_BIAS_F32, _ptr(a16), _ptr(wo), _ptr(b_out), _ptr(logits), M, npad, E,
                                              None, None, 0, 0, _stream()))
            return logits.view(B, T, npad)[:, :, :V].contiguous()
        a16 = torch.empty((M, E), dtype=torch.float16, device=dev)
        _lib.check(lib.esmb200_layernorm_f16(_ptr(x_pre), _ptr(ln_w), _ptr(ln_b), _ptr(a16), M, E, eps, _stream()))
        h = torch.empty((M, E), dtype=torch.float32, device=dev)
        _lib.check(lib.esmb200_gemm_f16(_lib.EPI_BIAS_GELU_F32, _ptr(a16), _ptr(w_dense), _ptr(b_dense),
                                        _ptr(h), M, E, E, None, None, 0, 0, _stream()))
        _lib.check(lib.esmb200_layernorm_f16(_ptr(h), _ptr(ln2_w), _ptr(ln2_b),
                                             _ptr(a16), M, E, self.layer_norm.eps, _stream()))
        logits = torch.empty((M, npad), dtype=torch.float32, device=dev)
        _lib.check(lib.esmb200_gemm_f16(_lib.EPI_BIAS_F32, _ptr(a16), _ptr(w_out), _ptr(b_out), _ptr(logits), M, npad, E,
                                        None, None, 0, 0, _stream()))
        return logits.view(B, T, npad)[:, :, :V].contiguous()  # [B,T,V] packed like the reference's (esm2.py:129)


class ContactPredictionHead(nn.Module):
    """modules.py:317-357 (symmetrize :27-29, apc :32-41) — PyTorch; SURVEY §8f #1 lists it as a next row."""

    def __init__(self, in_features: int, prepend_bos: bool, append_eos: bool, bias=True, eos_idx: Optional[int] = None):
        super().__init__()
        self.in_features = in_features
        self.prepend_bos = prepend_bos
        self.append_eos = append_eos
        if append_eos and eos_idx is None:
            raise ValueError("Using an alphabet with eos token, but no eos token was passed in.")
        self.eos_idx = eos_idx
        self.regression = nn.Linear(in_features, 1, bias)
        self.activation = nn.Sigmoid()

    def forward(self, tokens, attentions):
        """modules.py:338-357 evaluated without the [B, L*H, S, S] temporaries of symmetrize/apc (which need ~6x the
        24 GB attention stack of configs[3]): with A_c the eos-masked, cropped map of channel c = (layer, head),
            logit_ij = sum_c w_c (A_c + A_c^T)_ij - sum_c (w_c / a12_c) a1_c[i] a1_c[j] + b,
            a1_c = rowsum(A_c) + colsum(A_c),  a12_c = sum(a1_c).
        On the GPU every layer's maps are read once by esmb200_contact_accumulate (sum over heads + row/column sums,
        no atomics: bit-reproducible) and esmb200_contact_finalize fuses the rank-(L*H) correction, the symmetrisation,
        the bias and the sigmoid."""
        B, L, H, T, _ = attentions.shape
        lo = 1 if self.prepend_bos else 0
        hi = T - 1 if self.append_eos else T
        S = hi - lo
        w = self.regression.weight.view(L, H).to(attentions.dtype)
        if not (attentions.is_cuda and attentions.dtype == torch.float32 and attentions.is_contiguous()):
            return self._forward_torch(tokens, attentions, w, lo, hi)
        lib = _lib.load()
        dev = attentions.device
        keep8 = tokens.ne(self.eos_idx).to(torch.uint8).contiguous() if self.append_eos else None
        nt = (S + 15) // 16
        acc = torch.zeros((B, S, S), dtype=torch.float32, device=dev)
        a1 = torch.empty((B, L, H, S), dtype=torch.float32, device=dev)
        row = torch.empty((B, H, S), dtype=torch.float32, device=dev)
        col = torch.empty((B, H, nt, S), dtype=torch.float32, device=dev)
        wl = w.float().contiguous()
        with torch.cuda.device(dev):
            for l in range(L):
                _lib.check(lib.esmb200_contact_accumulate(
                    ctypes.c_void_p(attentions.data_ptr() + l * H * T * T * 4), L * H * T * T,
                    ctypes.c_void_p(wl.data_ptr() + l * H * 4), _ptr(keep8), _ptr(acc), _ptr(row), _ptr(col),
                    B, H, T, lo, hi, _stream()))
                torch.add(row, col.sum(2), out=a1[:, l])               # a1_c = rowsum + colsum, fixed summation order
        return self._finalize(acc, a1.view(B, L * H, S), wl)

    def _finalize(self, acc: torch.Tensor, a1f: torch.Tensor, wl: torch.Tensor) -> torch.Tensor:
        """acc [B,S,S], a1f [B,L*H,S] (rowsum + colsum per channel), wl [L,H] -> contacts [B,S,S]."""
        lib = _lib.load()
        B, C, S = a1f.shape
        with torch.cuda.device(acc.device):
            a1f = a1f.contiguous()
            a12 = a1f.sum(-1, keepdim=True)                               # [B, L*H, 1]
            u = (a1f * (wl.reshape(1, C, 1) / a12)).contiguous()
            bias = _f32(self.regression.bias) if self.regression.bias is not None else None
            out = torch.empty((B, S, S), dtype=torch.float32, device=acc.device)
            _lib.check(lib.esmb200_contact_finalize(_ptr(acc), _ptr(u), _ptr(a1f), _ptr(bias), _ptr(out), B, C, S, _stream()))
        return out

    # ---- fused path: the accumulators are filled by esmb200_stack_forward while the probabilities are written -------
    def begin_job(self, tokens: torch.Tensor, num_layers: int, num_heads: int):
        """Buffers + esmb200_contact_job for a [B,T] batch (attention_contact.cuh); finish_job() turns them into contacts."""
        B, T = tokens.shape
        lo = 1 if self.prepend_bos else 0
        hi = T - 1 if self.append_eos else T
        S = hi - lo
        dev = tokens.device
        nt = (T + 127) // 128
        st = {
            "keep": tokens.ne(self.eos_idx).to(torch.uint8).contiguous() if self.append_eos else None,
            "acc": torch.zeros((B, S, S), dtype=torch.float32, device=dev),
            "row": torch.empty((num_layers, B, num_heads, 4 * nt, S), dtype=torch.float32, device=dev),
            "col": torch.empty((num_layers, B, num_heads, 4 * nt, S), dtype=torch.float32, device=dev),
            "w": _f32(self.regression.weight).view(num_layers, num_heads).contiguous(),
        }
        job = _lib.ContactJob()
        job.weights, job.keep = st["w"].data_ptr(), (st["keep"].data_ptr() if st["keep"] is not None else None)
        job.acc, job.row_part, job.col_part = st["acc"].data_ptr(), st["row"].data_ptr(), st["col"].data_ptr()
        job.lo, job.hi = lo, hi
        st["job"] = job
        return st

    def finish_job(self, st) -> torch.Tensor:
        L, B, H, _, S = st["row"].shape
        a1 = st["row"].sum(3) + st["col"].sum(3)                          # [L,B,H,S], fixed summation order
        return self._finalize(st["acc"], a1.permute(1, 0, 2, 3).reshape(B, L * H, S), st["w"])

    def _forward_torch(self, tokens, attentions, w, lo, hi):
        """The same formula with PyTorch ops (non-CUDA or non-fp32 inputs; cross-check in the tests)."""
        B, L, H, T, _ = attentions.shape
        S = hi - lo
        keep = None
        if self.append_eos:
            keep = tokens.ne(self.eos_idx).to(attentions)[:, lo:hi]  # [B,S]
        acc = torch.zeros((B, S, S), dtype=attentions.dtype, device=attentions.device)
        corr = torch.zeros_like(acc)
        for l in range(L):
            a = attentions[:, l, :, lo:hi, lo:hi]  # [B,H,S,S] view
            if keep is not None:
                a = a * (keep[:, None, :, None] * keep[:, None, None, :])
            acc += torch.einsum("bhij,h->bij", a, w[l])
            a1 = a.sum(-1) + a.sum(-2)  # [B,H,S]
            a12 = a1.sum(-1)            # [B,H]
            corr += torch.einsum("bhi,bhj->bij", a1 * (w[l][None, :] / a12)[:, :, None], a1)
        logits = acc + acc.transpose(-1, -2) - corr
        if self.regression.bias is not None:
            logits = logits + self.regression.bias
        return self.activation(logits)


class ESM2(nn.Module):
    """Drop-in for esm.model.esm2.ESM2 (esm2.py:14-147): same constructor, same state-dict keys (so
    `load_state_dict(reference_model.state_dict())` and the esm2_t*.pt checkpoints load), same forward contract."""

    def __init__(self, num_layers: int = 33, embed_dim: int = 1280, attention_heads: int = 20,
                 alphabet: Union[Alphabet, str] = "ESM-1b", token_dropout: bool = True):
        super().__init__()
        self.num_layers = num_layers
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        if isinstance(alphabet, str):
            alphabet = Alphabet.from_architecture(alphabet)
        self.alphabet = alphabet
        self.alphabet_size = len(alphabet)
        self.padding_idx = alphabet.padding_idx
        self.mask_idx = alphabet.mask_idx
        self.cls_idx = alphabet.cls_idx
        self.eos_idx = alphabet.eos_idx
        self.prepend_bos = alphabet.prepend_bos
        self.append_eos = alphabet.append_eos
        self.token_dropout = token_dropout
        self.embed_scale = 1
        self.embed_tokens = nn.Embedding(self.alphabet_size, embed_dim, padding_idx=self.padding_idx)
        self.layers = nn.ModuleList(
            [TransformerLayer(embed_dim, 4 * embed_dim, attention_heads) for _ in range(num_layers)])
        self.contact_head = ContactPredictionHead(num_layers * attention_heads, self.prepend_bos, self.append_eos,
                                                  eos_idx=self.eos_idx)
        self.emb_layer_norm_after = nn.LayerNorm(embed_dim)
        self.lm_head = RobertaLMHead(embed_dim, self.alphabet_size, self.embed_tokens.weight)
        self._rope_cache = None
        self._mirrors: Dict[str, tuple] = {}
        self.precision = "fp16"

    PRECISIONS = {"fp16": 0, "fp32x3": 1}

    def set_precision(self, name: str) -> "ESM2":
        """"fp16" (default): fp16 MMA operands, fp32 accumulation — the fast path bench.py measures.
        "fp32x3": every MMA operand (LayerNorm output, weights, q, k, v, softmax probabilities, context, FFN hidden) is
        an fp16 hi + lo pair and every product runs hi*hi + lo*hi + hi*lo into the fp32 accumulator: 22 significand bits
        per operand, fp32-grade parity with the reference at ~3x the tensor work.  Needs
        embed_dim % 64 == 0."""
        if name not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(self.PRECISIONS)}")
        if name == "fp32x3" and (self.embed_dim % 64 != 0 or self.embed_dim // self.attention_heads > 64):
            raise ValueError("fp32x3 precision needs embed_dim % 64 == 0 and head_dim <= 64")
        self.precision = name
        for layer in self.layers:
            layer.precision = self.PRECISIONS[name]
        return self

    def _rope_tables(self, T: int):
        inv = self.layers[0].self_attn.rot_emb.inv_freq
        key = (T, inv.device, inv.data_ptr())
        if self._rope_cache is None or self._rope_cache[0] != key:
            self._rope_cache = (key,) + rope_tables(inv, T)
        return self._rope_cache[1], self._rope_cache[2]

    def _mirror(self, name: str, p: torch.Tensor) -> torch.Tensor:
        """fp32 mirror of a non-fp32 parameter (model.half()), cached until the parameter changes."""
        if p.dtype == torch.float32 and p.is_contiguous():
            return p.detach()
        key = (p.data_ptr(), p._version, p.dtype)
        hit = self._mirrors.get(name)
        if hit is None or hit[0] != key:
            hit = (key, _f32(p))
            self._mirrors[name] = hit
        return hit[1]

    @torch.no_grad()
    def forward(self, tokens, repr_layers=[], need_head_weights=False, return_contacts=False):
        if return_contacts:
            need_head_weights = True
        assert tokens.ndim == 2
        if not tokens.is_cuda:
            raise _lib.Esmb200Error("esm_b200 runs on CUDA (sm_100a) only: pass tokens.cuda(); no CPU fallback")
        lib = _lib.load()
        if tokens.dtype != torch.int64:
            if tokens.dtype.is_floating_point or tokens.dtype == torch.bool:
                raise TypeError(f"tokens must be an integer tensor, got {tokens.dtype}")
            tokens = tokens.long()
        tokens = tokens.contiguous()
        # nn.Embedding raises a device-side assert on an out-of-range id (esm2.py:84); same here, without a host sync
        torch._assert_async(((tokens >= 0) & (tokens < self.alphabet_size)).all())
        B, T = tokens.shape
        E, N = self.embed_dim, self.num_layers
        dtype = self.embed_tokens.weight.dtype  # fp32, or fp16/bf16 after model.half() (esmfold.py:59-62)
        padding_mask = tokens.eq(self.padding_idx)  # esm2.py:82
        repr_layers = set(repr_layers)
        hidden: Dict[int, torch.Tensor] = {}
        cast = (lambda t: t) if dtype == torch.float32 else (lambda t: t.to(dtype))

        with torch.cuda.device(tokens.device):
            # esm2.py:84-95 embedding prologue
            x = torch.empty((B, T, E), dtype=torch.float32, device=tokens.device)
            table = self._mirror("embed_tokens", self.embed_tokens.weight)
            _lib.check(lib.esmb200_embed_tokens(_ptr(tokens), _ptr(table), _ptr(x), B, T, E,
                                                self.padding_idx, self.mask_idx, int(self.token_dropout), _stream()))
            if 0 in repr_layers:
                hidden[0] = cast(x.clone())
            # esm2.py:108-109 drops the mask when the batch has no padding; that test is a device->host sync, which
            # would serialise back-to-back forwards (bulk extraction). The kernels take the all-false mask at no cost
            # (one uniform compare per 32 keys), so the mask is always passed.
            mask = padding_mask
            # esm2.py:111-121 layer loop (intermediate representations are copied out by the library)
            repr_out = {i - 1: torch.empty_like(x) for i in repr_layers if 0 < i < N}
            cos, sin = self._rope_tables(T)
            # contacts: folded into the probability pass (fp16 mode; fp32x3 runs the separate kernels afterwards)
            cjob = (self.contact_head.begin_job(tokens, N, self.attention_heads)
                    if return_contacts and self.precision == "fp16" else None)
            attn_t = run_stack(list(self.layers), x, mask, cos, sin, repr_out,
                               list(range(N)) if need_head_weights else [], zero_pad_rows=True,
                               contact_job=cjob["job"] if cjob else None)
            for i, t in repr_out.items():
                hidden[i + 1] = cast(t)
            # esm2.py:129 LM head, from the pre-LN stream (its first step is the same emb_layer_norm_after)
            ln = self.emb_layer_norm_after
            ln_w, ln_b = self._mirror("ln_after.w", ln.weight), self._mirror("ln_after.b", ln.bias)
            logits = cast(self.lm_head.forward_native(x, ln_w, ln_b, ln.eps, self.PRECISIONS[self.precision]))
            # esm2.py:123-128 final LayerNorm; the last representation is post-LN
            _lib.check(lib.esmb200_layernorm(_ptr(x), _ptr(ln_w), _ptr(ln_b), _ptr(x), B * T, E, ln.eps, _stream()))
        if N in repr_layers:
            hidden[N] = cast(x)
        result = {"logits": logits, "representations": hidden}
        if need_head_weights:
            # B x L x H x T x T (esm2.py:134), each layer written in place by the library; rows/columns of padded
            # tokens are already zero (esm2.py:135-139: padded keys have probability 0, padded query rows are zeroed
            # by the probability kernel), so no masking pass over the stack is needed
            attentions = attn_t["stacked"]
            result["attentions"] = cast(attentions)
            if return_contacts:
                contacts = self.contact_head.finish_job(cjob) if cjob else self.contact_head(tokens, attentions)
                result["contacts"] = cast(contacts)
        return result

    def predict_contacts(self, tokens):
        return self(tokens, return_contacts=True)["contacts"]
