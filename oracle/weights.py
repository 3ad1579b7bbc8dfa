"""ORACLE support — deterministic ESM-2 weights and tokens shared by the reference, the oracle and the CUDA path.

Pretrained checkpoints are unreachable offline (esm/pretrained.py:53), so parity is established on seeded random
weights with the reference's state-dict keys and shapes (SURVEY §7 data-layout notes).  Following SURVEY §7.1 the
LayerNorm gains/biases and the zero-initialised biases are randomised (the defaults 1/0/0 would hide epilogue bugs)
and q/k projections are scaled up (x1.5) so the softmax is peaked (max probabilities ~0.9) but not saturated.
"""
from __future__ import annotations

from typing import Dict

import torch

VOCAB = 33


def make_state_dict(num_layers: int, embed_dim: int, num_heads: int, seed: int = 0,
                    qk_gain: float = 1.5) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    E, F, d = embed_dim, 4 * embed_dim, embed_dim // num_heads

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    emb = rn(VOCAB, E, std=1.0)
    emb[1].zero_()  # padding_idx row (nn.Embedding(padding_idx=1), esm2.py:43-47)
    sd["embed_tokens.weight"] = emb
    w_std = E ** -0.5
    for i in range(num_layers):
        p = f"layers.{i}."
        for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
            gain = qk_gain if name in ("q_proj", "k_proj") else 1.0
            sd[p + f"self_attn.{name}.weight"] = rn(E, E, std=w_std * gain)
            sd[p + f"self_attn.{name}.bias"] = rn(E, std=0.1)
        sd[p + "self_attn.rot_emb.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))
        sd[p + "self_attn_layer_norm.weight"] = 1.0 + rn(E, std=0.2)
        sd[p + "self_attn_layer_norm.bias"] = rn(E, std=0.1)
        sd[p + "fc1.weight"] = rn(F, E, std=w_std)
        sd[p + "fc1.bias"] = rn(F, std=0.1)
        sd[p + "fc2.weight"] = rn(E, F, std=F ** -0.5)
        sd[p + "fc2.bias"] = rn(E, std=0.1)
        sd[p + "final_layer_norm.weight"] = 1.0 + rn(E, std=0.2)
        sd[p + "final_layer_norm.bias"] = rn(E, std=0.1)
    sd["contact_head.regression.weight"] = rn(1, num_layers * num_heads, std=1.0)
    sd["contact_head.regression.bias"] = rn(1, std=0.1)
    sd["emb_layer_norm_after.weight"] = 1.0 + rn(E, std=0.2)
    sd["emb_layer_norm_after.bias"] = rn(E, std=0.1)
    sd["lm_head.weight"] = sd["embed_tokens.weight"]  # tied, esm2.py:71-75
    sd["lm_head.bias"] = rn(VOCAB, std=0.1)
    sd["lm_head.dense.weight"] = rn(E, E, std=w_std)
    sd["lm_head.dense.bias"] = rn(E, std=0.1)
    sd["lm_head.layer_norm.weight"] = 1.0 + rn(E, std=0.2)
    sd["lm_head.layer_norm.bias"] = rn(E, std=0.1)
    return sd


def make_tokens(lengths, total_len: int, seed: int = 1234, n_mask: int = 0) -> torch.Tensor:
    """tokens [B, total_len]: <cls>, `length` residues drawn from the 20 standard amino acids (ids 4..23),
    <eos>, then <pad>. `n_mask` residues of the first sequence are replaced by <mask> (id 32)."""
    g = torch.Generator().manual_seed(seed)
    B = len(lengths)
    tok = torch.full((B, total_len), 1, dtype=torch.int64)
    for b, n in enumerate(lengths):
        assert n + 2 <= total_len
        tok[b, 0] = 0
        tok[b, 1: n + 1] = torch.randint(4, 24, (n,), generator=g)
        tok[b, n + 1] = 2
    for j in range(n_mask):
        tok[0, 2 + 3 * j] = 32
    return tok
