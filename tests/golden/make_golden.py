"""Generates tests/golden/*.pt by running the UNMODIFIED reference (facebookresearch/esm imported from
/root/reference) on the deterministic weights/tokens of oracle/weights.py.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [case ...]
The fixtures hold tokens + reference outputs; the weights are re-created from (num_layers, E, H, seed) by
oracle.weights.make_state_dict on whichever machine runs the tests (same torch version => same CPU generator stream;
a checksum of the state dict is stored in the fixture and verified by the tests).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import esm  # noqa: E402  (the reference)
from oracle.weights import make_state_dict, make_tokens  # noqa: E402

CASES = {
    # name: (layers, E, H, lengths, total_len, n_mask, repr_layers)
    # d=64 cases run on the CUDA path; t6_8M_like (d=16) is BASELINE.json configs[0], oracle/plumbing only.
    "tiny_L2_E128_H2": (2, 128, 2, [38, 21, 30], 40, 2, [0, 1, 2]),
    "mid_L3_E256_H4": (3, 256, 4, [198, 150], 200, 1, [0, 2, 3]),
    "t6_8M_like_L6_E320_H20": (6, 320, 20, [64, 64, 64, 64], 66, 0, [6]),
    "nopad_L2_E128_H2": (2, 128, 2, [126, 126], 128, 0, [2]),
    # head_dim 128 = esm2_t48_15B's head width (pretrained.py:390-397) at a small embedding width
    "t48_15B_like_L2_E256_H2": (2, 256, 2, [140, 97], 142, 1, [0, 1, 2]),
}


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items())))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for name, (L, E, H, lengths, total, n_mask, repr_layers) in CASES.items():
        if only and name not in only:
            continue
        sd = make_state_dict(L, E, H, seed=0)
        model = esm.model.esm2.ESM2(num_layers=L, embed_dim=E, attention_heads=H, alphabet="ESM-1b")
        missing = model.load_state_dict(sd, strict=True)
        model.eval()
        tokens = make_tokens(lengths, total, seed=1234, n_mask=n_mask)
        with torch.no_grad():
            out = model(tokens, repr_layers=repr_layers, need_head_weights=True, return_contacts=True)
        fixture = {
            "config": {"num_layers": L, "embed_dim": E, "attention_heads": H, "seed": 0},
            "state_dict_checksum": checksum(sd),
            "tokens": tokens,
            "repr_layers": repr_layers,
            "logits": out["logits"].clone(),
            "representations": {k: v.clone() for k, v in out["representations"].items()},
            # full [B,L,H,T,T] only when small; otherwise first/last layer x first/last head (indices recorded)
            "attentions": out["attentions"].clone() if out["attentions"].numel() <= 500_000 else None,
            "attentions_sub_layers": [0, L - 1],
            "attentions_sub_heads": [0, H - 1],
            "attentions_sub": out["attentions"][:, [0, L - 1]][:, :, [0, H - 1]].clone(),
            "contacts": out["contacts"].clone(),
            "reference": "facebookresearch/esm @ 2b36991 (fair-esm 2.0.1), torch %s, CPU fp32" % torch.__version__,
        }
        path = os.path.join(HERE, name + ".pt")
        torch.save(fixture, path)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", missing)


if __name__ == "__main__":
    main()
