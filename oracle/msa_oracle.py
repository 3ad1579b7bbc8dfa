"""ORACLE (test infrastructure only) — CPU restatement of the MSA-Transformer axial block (BASELINE.json configs[4]).

Follows /root/reference/esm/axial_attention.py (RowSelfAttention :71-130, ColumnSelfAttention :182-222) and
/root/reference/esm/modules.py (AxialTransformerLayer :195-221, NormalizedResidualBlock :375-392,
FeedForwardNetwork :413-418), batch-major and functional over a state dict with the reference's key names.
and /root/reference/esm/model/msa_transformer.py:146-220 (the whole MSATransformer.forward: embedding prologue with
LearnedPositionalEmbedding modules.py:241-257, layer loop, final LayerNorm, LM head, contact head on the row attentions).
Pinned against outputs of the reference's own AxialTransformerLayer and MSATransformer (tests/golden/msa_*.pt,
make_golden_msa.py).  x layout here: [B, R, C, E] (the reference uses [R, C, B, E]).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

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


PAD = 1  # <pad> of the "MSA Transformer" alphabet (data.py:158-164); <cls> = 0 is prepended, no <eos>


def make_msa_state_dict(n_layers: int, embed_dim: int, ffn_dim: int, num_heads: int, seed: int = 0,
                        max_positions: int = 1024, vocab: int = 33, msa_pos_dim: Optional[int] = None):
    """Deterministic weights for the whole MSATransformer under the reference's state-dict names."""
    sd = make_axial_state_dict(embed_dim, ffn_dim, seed=seed, n_layers=n_layers)
    g = torch.Generator().manual_seed(seed + 7919)
    E = embed_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd["embed_tokens.weight"] = rn(vocab, E)
    sd["embed_tokens.weight"][PAD].zero_()
    sd["embed_positions.weight"] = rn(max_positions + PAD + 1, E, std=0.5)
    sd["embed_positions.weight"][PAD].zero_()
    sd["msa_position_embedding"] = rn(1, 1024, 1, E if msa_pos_dim is None else msa_pos_dim, std=0.5)
    for name in ("emb_layer_norm_before", "emb_layer_norm_after", "lm_head.layer_norm"):
        sd[name + ".weight"] = 1.0 + rn(E, std=0.2)
        sd[name + ".bias"] = rn(E, std=0.1)
    sd["lm_head.dense.weight"] = rn(E, E, std=E ** -0.5)
    sd["lm_head.dense.bias"] = rn(E, std=0.1)
    sd["lm_head.weight"] = sd["embed_tokens.weight"]  # tied (modules.py:305)
    sd["lm_head.bias"] = rn(vocab, std=0.1)
    sd["contact_head.regression.weight"] = rn(1, n_layers * num_heads, std=2.0)
    sd["contact_head.regression.bias"] = rn(1, std=0.5)
    return sd


def make_msa_tokens(B: int, R: int, C: int, seed: int = 1234, pad_cols: int = 0, pad_rows_last: int = 0):
    """tokens [B,R,C]: column 0 = <cls>, residues uniform over ids 4..23 and the gap symbol 30 ("-"); optional trailing
    padding columns (all MSAs) and padding rows (last MSA), the way MSABatchConverter pads (data.py:310-338)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(4, 24, (B, R, C), generator=g)
    gaps = torch.rand(B, R, C, generator=g) < 0.1
    t[gaps] = 30
    t[:, :, 0] = 0
    if pad_cols:
        t[:, :, C - pad_cols:] = PAD
    if pad_rows_last:
        t[B - 1, R - pad_rows_last:, :] = PAD
    return t


@torch.no_grad()
def msa_transformer_forward(sd: Dict[str, torch.Tensor], n_layers: int, num_heads: int, tokens: torch.Tensor,
                            repr_layers: Iterable[int] = (), need_head_weights: bool = False,
                            return_contacts: bool = False):
    """msa_transformer.py:146-220 — same result dict as MSATransformer.forward."""
    if return_contacts:
        need_head_weights = True
    B, R, C = tokens.shape
    pad = tokens.eq(PAD)
    padding_mask = pad if bool(pad.any()) else None
    x = sd["embed_tokens.weight"][tokens]
    nonpad = tokens.ne(PAD).view(B * R, C).int()
    positions = (torch.cumsum(nonpad, 1).int() * nonpad).long() + PAD          # modules.py:249-250
    x = x + sd["embed_positions.weight"][positions].view(B, R, C, -1)
    if "msa_position_embedding" in sd:
        x = x + sd["msa_position_embedding"][:, :R]
    x = F.layer_norm(x, (x.shape[-1],), sd["emb_layer_norm_before.weight"], sd["emb_layer_norm_before.bias"], 1e-5)
    if padding_mask is not None:
        x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))
    repr_layers = set(repr_layers)
    hidden = {}
    if 0 in repr_layers:
        hidden[0] = x
    rows, cols = [], []
    for i in range(n_layers):
        out = axial_layer(x, sd, f"layers.{i}.", num_heads, padding_mask, need_head_weights)
        if need_head_weights:
            x, col, row = out
            cols.append(col.permute(2, 0, 1, 3, 4))   # H,C,B,R,R -> B,H,C,R,R
            rows.append(row.permute(1, 0, 2, 3))      # H,B,C,C -> B,H,C,C
        else:
            x = out
        if (i + 1) in repr_layers:
            hidden[i + 1] = x
    x = F.layer_norm(x, (x.shape[-1],), sd["emb_layer_norm_after.weight"], sd["emb_layer_norm_after.bias"], 1e-5)
    if n_layers in repr_layers:
        hidden[n_layers] = x
    h = F.gelu(F.linear(x, sd["lm_head.dense.weight"], sd["lm_head.dense.bias"]))
    h = F.layer_norm(h, (h.shape[-1],), sd["lm_head.layer_norm.weight"], sd["lm_head.layer_norm.bias"], 1e-5)
    logits = F.linear(h, sd["embed_tokens.weight"]) + sd["lm_head.bias"]
    result = {"logits": logits, "representations": hidden}
    if need_head_weights:
        result["col_attentions"] = torch.stack(cols, 1)
        result["row_attentions"] = torch.stack(rows, 1)
        if return_contacts:
            # modules.py:338-357 with prepend_bos, no eos: strip the <cls> row/column, symmetrize, APC, regression
            a = result["row_attentions"][..., 1:, 1:]
            Bq, L, H, S, _ = a.shape
            a = a.reshape(Bq, L * H, S, S)
            a = a + a.transpose(-1, -2)
            a1, a2, a12 = a.sum(-1, keepdim=True), a.sum(-2, keepdim=True), a.sum((-1, -2), keepdim=True)
            a = a - a1 * a2 / a12
            logit = F.linear(a.permute(0, 2, 3, 1), sd["contact_head.regression.weight"],
                             sd["contact_head.regression.bias"])
            result["contacts"] = torch.sigmoid(logit.squeeze(3))
    return result
