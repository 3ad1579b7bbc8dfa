"""ORACLE (test infrastructure only) — CPU restatement of the MSA-Transformer axial block (BASELINE.json configs[4]).

Follows /root/reference/esm/axial_attention.py (RowSelfAttention :71-130, ColumnSelfAttention :182-222) and
/root/reference/esm/modules.py (AxialTransformerLayer :195-221, NormalizedResidualBlock :375-392,
FeedForwardNetwork :413-418), batch-major and functional over a state dict with the reference's key names.
Pinned against outputs of the reference's own AxialTransformerLayer (tests/golden/msa_*.pt, make_golden_msa.py).
x layout here: [B, R, C, E] (the reference uses [R, C, B, E]).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _ln(x, sd, pre):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], 1e-5)


def row_attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, num_heads: int,
                  padding_mask: Optional[torch.Tensor] = None):
    """axial_attention.py:71-130 — tied row attention: logits summed over the R alignment rows, scale d^-1/2 / sqrt(R)
    (:36-38), padded positions zeroed in q (:82-85) and -10000 on padded key columns of row 0 (:94-97).
    x [B,R,C,E], padding_mask [B,R,C] bool. Returns (out [B,R,C,E], probs [H,B,C,C])."""
    B, R, C, E = x.shape
    d = E // num_heads
    scaling = (d ** -0.5) / math.sqrt(R)
    q = F.linear(x, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"]).view(B, R, C, num_heads, d) * scaling
    k = F.linear(x, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"]).view(B, R, C, num_heads, d)
    v = F.linear(x, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"]).view(B, R, C, num_heads, d)
    if padding_mask is not None:
        q = q * (1 - padding_mask[..., None, None].to(q))
    logits = torch.einsum("brihd,brjhd->hbij", q, k)
    if padding_mask is not None:
        logits = logits.masked_fill(padding_mask[:, 0][None, :, None, :], -10000)
    probs = logits.softmax(-1)
    ctx = torch.einsum("hbij,brjhd->brihd", probs, v).reshape(B, R, C, E)
    return F.linear(ctx, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"]), probs


def column_attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, num_heads: int,
                     padding_mask: Optional[torch.Tensor] = None):
    """axial_attention.py:182-222 — per-column attention over the R rows, scale d^-1/2, -10000 on padded keys.
    Returns (out [B,R,C,E], probs [H,C,B,R,R])."""
    B, R, C, E = x.shape
    d = E // num_heads
    q = F.linear(x, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"]).view(B, R, C, num_heads, d) * (d ** -0.5)
    k = F.linear(x, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"]).view(B, R, C, num_heads, d)
    v = F.linear(x, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"]).view(B, R, C, num_heads, d)
    logits = torch.einsum("bichd,bjchd->hcbij", q, k)
    if padding_mask is not None:
        logits = logits.masked_fill(padding_mask.permute(2, 0, 1)[None, :, :, None, :], -10000)
    probs = logits.softmax(-1)
    ctx = torch.einsum("hcbij,bjchd->bichd", probs, v).reshape(B, R, C, E)
    return F.linear(ctx, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"]), probs


@torch.no_grad()
def axial_layer(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, num_heads: int,
                padding_mask: Optional[torch.Tensor] = None, need_head_weights: bool = False):
    """modules.py:195-221 with each sub-layer wrapped as x + f(LN(x)) (modules.py:375-392, dropout = identity)."""
    p = pre + "row_self_attention."
    a, row_probs = row_attention(_ln(x, sd, p), sd, p + "layer.", num_heads, padding_mask)
    x = x + a
    p = pre + "column_self_attention."
    a, col_probs = column_attention(_ln(x, sd, p), sd, p + "layer.", num_heads, padding_mask)
    x = x + a
    p = pre + "feed_forward_layer."
    h = F.gelu(F.linear(_ln(x, sd, p), sd[p + "layer.fc1.weight"], sd[p + "layer.fc1.bias"]))  # nn.GELU(): exact erf
    x = x + F.linear(h, sd[p + "layer.fc2.weight"], sd[p + "layer.fc2.bias"])
    if need_head_weights:
        return x, col_probs, row_probs
    return x


def make_axial_state_dict(embed_dim: int, ffn_dim: int, seed: int = 0, n_layers: int = 1) -> Dict[str, torch.Tensor]:
    """Deterministic weights with the reference's AxialTransformerLayer key names (prefix "layers.{i}.")."""
    g = torch.Generator().manual_seed(seed)
    E, Fd = embed_dim, ffn_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    for i in range(n_layers):
        for blk in ("row_self_attention", "column_self_attention"):
            p = f"layers.{i}.{blk}."
            for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sd[p + f"layer.{name}.weight"] = rn(E, E, std=E ** -0.5 * (1.5 if name in ("q_proj", "k_proj") else 1.0))
                sd[p + f"layer.{name}.bias"] = rn(E, std=0.1)
            sd[p + "layer_norm.weight"] = 1.0 + rn(E, std=0.2)
            sd[p + "layer_norm.bias"] = rn(E, std=0.1)
        p = f"layers.{i}.feed_forward_layer."
        sd[p + "layer.fc1.weight"] = rn(Fd, E, std=E ** -0.5)
        sd[p + "layer.fc1.bias"] = rn(Fd, std=0.1)
        sd[p + "layer.fc2.weight"] = rn(E, Fd, std=Fd ** -0.5)
        sd[p + "layer.fc2.bias"] = rn(E, std=0.1)
        sd[p + "layer_norm.weight"] = 1.0 + rn(E, std=0.2)
        sd[p + "layer_norm.bias"] = rn(E, std=0.1)
    return sd
