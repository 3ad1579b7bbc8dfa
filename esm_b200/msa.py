"""MSA-Transformer axial block on the sm_100a kernels — host-side mirror of
/root/reference/esm/modules.py:145-221 (AxialTransformerLayer, NormalizedResidualBlock :360-392,
FeedForwardNetwork :395-418) and /root/reference/esm/axial_attention.py (RowSelfAttention :11-130,
ColumnSelfAttention :133-239), with the reference's parameter names so its state dicts load.

CUDA path for BASELINE.json configs[4].  All the arithmetic of the block runs in libesmb200.so:
  * LayerNorm -> fp16, q/k/v projection (+ bias, q scale; no rotary embedding), out-projection + residual,
    fc1 + erf-GELU, fc2 + residual: the same tcgen05 GEMM / LayerNorm kernels as the ESM-2 path;
  * tied row attention (logits summed over the R rows, axial_attention.py:87): esmb200_tied_row_attention — a tcgen05
    contraction over K = R*64 that walks the alignment rows with TMA boxes taken straight from the projection output,
    a row softmax (with the reference's -10000 fill on padded key columns), and the P.V update with V tiles as the
    MN-major operand (csrc/tied_attention.cuh);
  * column attention: esmb200_column_attention — the flash-attention kernel with strided TMA boxes, one "sequence" of
    R rows per alignment column, no regrouping copy (csrc/attention4.cuh, AttnParams::cols).
PyTorch is used for the buffers, for zeroing q at padded positions (axial_attention.py:82-85, one masked_fill_) and, only
when the column attention MAPS are requested, for regrouping qkv column-major.

Padding: the reference fills padded keys with -10000, this path gives them probability exactly 0 in the column
attention — identical unless every key of a column is padded (there the reference averages v uniformly, this path
returns 0); such positions are themselves padding.  head_dim 64, inference only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .model import _ptr, _stream


class _AttnParams(nn.Module):
    """q/k/v/out projections under the reference's names (axial_attention.py:30-34, 151-155)."""

    def __init__(self, embed_dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)


class _FFNParams(nn.Module):
    def __init__(self, embed_dim: int, ffn_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, ffn_dim)
        self.fc2 = nn.Linear(ffn_dim, embed_dim)


class _ResidualBlock(nn.Module):
    """`layer` + `layer_norm` containers of NormalizedResidualBlock (modules.py:360-373)."""

    def __init__(self, layer: nn.Module, embed_dim: int):
        super().__init__()
        self.layer = layer
        self.layer_norm = nn.LayerNorm(embed_dim)


def _gemm(epi: int, a16: torch.Tensor, w16: torch.Tensor, bias: torch.Tensor, out: torch.Tensor) -> None:
    M, K = a16.shape
    N = w16.shape[0]
    _lib.check(_lib.load().esmb200_gemm_f16(epi, _ptr(a16), _ptr(w16), _ptr(bias), _ptr(out), M, N, K, None, None, 0, 0,
                                            _stream()))


def _ln16(x: torch.Tensor, ln: nn.LayerNorm, out16: torch.Tensor) -> None:
    M, E = x.shape
    _lib.check(_lib.load().esmb200_layernorm_f16(_ptr(x), _ptr(ln.weight), _ptr(ln.bias), _ptr(out16), M, E, ln.eps,
                                                 _stream()))


class AxialTransformerLayer(nn.Module):
    """Drop-in for esm.modules.AxialTransformerLayer (modules.py:145-221) at inference."""

    def __init__(self, embedding_dim: int = 768, ffn_embedding_dim: int = 3072, num_attention_heads: int = 8,
                 dropout: float = 0.1, attention_dropout: float = 0.1, activation_dropout: float = 0.1,
                 max_tokens_per_msa: int = 2 ** 14) -> None:
        super().__init__()
        if embedding_dim != 64 * num_attention_heads:
            raise ValueError("esm_b200 supports head_dim == 64 only")
        self.embedding_dim = embedding_dim
        self.ffn_embedding_dim = ffn_embedding_dim
        self.num_heads = num_attention_heads
        self.row_self_attention = _ResidualBlock(_AttnParams(embedding_dim, num_attention_heads), embedding_dim)
        self.column_self_attention = _ResidualBlock(_AttnParams(embedding_dim, num_attention_heads), embedding_dim)
        self.feed_forward_layer = _ResidualBlock(_FFNParams(embedding_dim, ffn_embedding_dim), embedding_dim)
        self._packed = None
        self._packed_key = None

    # ---- fp16 operand copies (re-made when a parameter changes) ------------------------------------------------
    def _pack(self):
        ps = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or key != self._packed_key:
            def qkv(a):
                w = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).detach().half().contiguous()
                b = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).detach().float().contiguous()
                return w, b
            row, col, ffn = self.row_self_attention.layer, self.column_self_attention.layer, self.feed_forward_layer.layer
            self._packed = {
                "row_qkv": qkv(row), "col_qkv": qkv(col),
                "row_out": row.out_proj.weight.detach().half().contiguous(),
                "col_out": col.out_proj.weight.detach().half().contiguous(),
                "fc1": ffn.fc1.weight.detach().half().contiguous(), "fc2": ffn.fc2.weight.detach().half().contiguous(),
            }
            self._packed_key = key
        return self._packed

    @torch.no_grad()
    def forward(self, x: torch.Tensor, self_attn_mask: Optional[torch.Tensor] = None,
                self_attn_padding_mask: Optional[torch.Tensor] = None, need_head_weights: bool = False):
        """x: (R, C, B, E) like the reference (modules.py:195-221); self_attn_padding_mask: (B, R, C) bool.
        Returns x, or (x, column_attn, row_attn)."""
        if self_attn_mask is not None:
            raise NotImplementedError
        if not x.is_cuda:
            raise _lib.Esmb200Error("esm_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        R, C, B, E = x.shape
        xb = x.permute(2, 0, 1, 3).contiguous().float()  # [B,R,C,E], updated in place by the residual epilogues
        row_probs, col_probs = self.forward_batch_major(xb, self_attn_padding_mask, need_head_weights)
        out = xb.permute(1, 2, 0, 3).to(x.dtype)
        if need_head_weights:
            return out, col_probs, row_probs
        return out

    @torch.no_grad()
    def forward_batch_major(self, xb: torch.Tensor, padding_mask: Optional[torch.Tensor] = None,
                            need_probs: bool = False):
        """In-place layer on the batch-major residual stream xb [B,R,C,E] fp32 (what MSATransformer keeps between
        layers).  padding_mask [B,R,C] bool or None.  Returns (row_attn [H,B,C,C], column_attn [H,C,B,R,R]) or
        (None, None)."""
        lib = _lib.load()
        B, R, C, E = xb.shape
        H, d, Fd = self.num_heads, 64, self.ffn_embedding_dim
        M = B * R * C
        dev = xb.device
        pk = self._pack()
        x2 = xb.view(M, E)
        xn = torch.empty((M, E), dtype=torch.float16, device=dev)
        qkv = torch.empty((M, 3 * E), dtype=torch.float16, device=dev)
        ctx = torch.empty((M, E), dtype=torch.float16, device=dev)
        key_pad = col_pad = None
        if padding_mask is not None:
            pm = padding_mask.to(device=dev, dtype=torch.bool)
            key_pad = pm[:, 0].contiguous().to(torch.uint8)                    # [B,C]   axial_attention.py:94-97
            col_pad = pm.permute(0, 2, 1).contiguous().to(torch.uint8)         # [B*C,R] axial_attention.py:212-215
        with torch.cuda.device(dev):
            # ================= tied row attention (axial_attention.py:71-111) =================
            blk = self.row_self_attention
            _ln16(x2, blk.layer_norm, xn)
            w, b = pk["row_qkv"]
            _lib.check(lib.esmb200_gemm_qkv_f16(_ptr(xn), _ptr(w), _ptr(b), _ptr(qkv), M, E,
                                                (d ** -0.5) / math.sqrt(R), None, None, 0, _stream()))
            if padding_mask is not None:  # q zeroed at padded positions (:82-85)
                qkv.view(B, R, C, 3, E)[:, :, :, 0].masked_fill_(pm[..., None], 0)
            row_probs = torch.empty((H, B, C, C), dtype=torch.float32, device=dev) if need_probs else None
            nbytes = lib.esmb200_tied_row_attention_scratch_bytes(B, C, H)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.esmb200_tied_row_attention(_ptr(qkv), _ptr(key_pad), _ptr(ctx), _ptr(row_probs), B, R, C, H,
                                                      _ptr(scratch), nbytes, _stream()))
            _gemm(_lib.EPI_BIAS_RESIDUAL, ctx, pk["row_out"], blk.layer.out_proj.bias, x2)            # :110 + residual

            # ================= column attention (axial_attention.py:182-222) =================
            blk = self.column_self_attention
            _ln16(x2, blk.layer_norm, xn)
            w, b = pk["col_qkv"]
            _lib.check(lib.esmb200_gemm_qkv_f16(_ptr(xn), _ptr(w), _ptr(b), _ptr(qkv), M, E, d ** -0.5, None, None, 0,
                                                _stream()))
            scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B * C, R), dtype=torch.uint8, device=dev)
            col_probs = None
            if not need_probs:  # q/k/v tiles are fetched straight from the row-major qkv (strided TMA boxes)
                _lib.check(lib.esmb200_column_attention(_ptr(qkv), _ptr(col_pad), _ptr(ctx), B, R, C, H, _ptr(scratch),
                                                        _stream()))
            else:               # probabilities requested: regroup column-major and use the probs-writing path
                qkv_t = qkv.view(B, R, C, 3 * E).permute(0, 2, 1, 3).contiguous()                  # [B*C, R, 3E]
                ctx_t = torch.empty((B * C * R, E), dtype=torch.float16, device=dev)
                col_probs = torch.empty((B * C, H, R, R), dtype=torch.float32, device=dev)
                _lib.check(lib.esmb200_attention(_ptr(qkv_t), _ptr(col_pad), _ptr(ctx_t), _ptr(col_probs), B * C, R, H,
                                                 _ptr(scratch), _stream()))
                ctx.view(B, R, C, E).copy_(ctx_t.view(B, C, R, E).permute(0, 2, 1, 3))
            _gemm(_lib.EPI_BIAS_RESIDUAL, ctx, pk["col_out"], blk.layer.out_proj.bias, x2)

            # ================= feed-forward (modules.py:413-418) =================
            blk = self.feed_forward_layer
            _ln16(x2, blk.layer_norm, xn)
            hbuf = torch.empty((M, Fd), dtype=torch.float16, device=dev)
            _gemm(_lib.EPI_BIAS_GELU, xn, pk["fc1"], blk.layer.fc1.bias, hbuf)
            _gemm(_lib.EPI_BIAS_RESIDUAL, hbuf, pk["fc2"], blk.layer.fc2.bias, x2)
        if need_probs:
            # reference shapes: column_attn [H, C, B, R, R] (axial_attention.py:206), row_attn [H, B, C, C] (:87)
            col_probs = col_probs.view(B, C, H, R, R).permute(2, 1, 0, 3, 4).contiguous()
        return row_probs, col_probs
