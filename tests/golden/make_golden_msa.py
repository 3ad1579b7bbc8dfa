"""Generates tests/golden/msa_*.pt by running the UNMODIFIED reference esm.modules.AxialTransformerLayer and
esm.model.msa_transformer.MSATransformer (/root/reference) on the deterministic weights of
oracle/msa_oracle.py::make_axial_state_dict / make_msa_state_dict.
Run in the build container only:  python tests/golden/make_golden_msa.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from esm.modules import AxialTransformerLayer  # noqa: E402  (the reference)
from argparse import Namespace  # noqa: E402

import esm  # noqa: E402  (the reference)
from oracle.msa_oracle import make_axial_state_dict, make_msa_state_dict, make_msa_tokens  # noqa: E402

MODEL_CASES = {  # name: (layers, E, F, H, B, R, C, pad_cols, pad_rows_last, max_tokens_per_msa)
    "msa_model_L2_E128_H2": (2, 128, 512, 2, 2, 5, 24, 3, 2, 2 ** 14),
    "msa_model_L3_E256_H4_nopad": (3, 256, 1024, 4, 1, 9, 70, 0, 0, 2 ** 8),  # chunked reference path
}


def model_cases():
    for name, (L, E, Fd, H, B, R, C, pc, pr, mt) in MODEL_CASES.items():
        sd = make_msa_state_dict(L, E, Fd, H, seed=0)
        alphabet = esm.data.Alphabet.from_architecture("msa_transformer")
        args = Namespace(layers=L, embed_dim=E, ffn_embed_dim=Fd, attention_heads=H, dropout=0.0, attention_dropout=0.0,
                         activation_dropout=0.0, max_tokens_per_msa=mt, max_tokens=mt, max_positions=1024,
                         embed_positions_msa=True)
        model = esm.MSATransformer(args, alphabet).eval()
        model.load_state_dict(sd, strict=True)
        tokens = make_msa_tokens(B, R, C, seed=1234, pad_cols=pc, pad_rows_last=pr)
        with torch.no_grad():
            out = model(tokens, repr_layers=[0, 1, L], return_contacts=True)
        fx = {"config": {"layers": L, "E": E, "F": Fd, "H": H, "seed": 0, "B": B, "R": R, "C": C, "tok_seed": 1234,
                         "pad_cols": pc, "pad_rows_last": pr},
              "tokens": tokens, "logits": out["logits"].clone(),
              "representations": {k: v.clone() for k, v in out["representations"].items()},
              "row_attentions": out["row_attentions"].clone(), "contacts": out["contacts"].clone(),
              "col_attentions_sample": out["col_attentions"][:, :, :, :3].clone(),
              "reference": "facebookresearch/esm @ 2b36991 esm.MSATransformer, torch %s CPU fp32" % torch.__version__}
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB")

CASES = {  # name: (E, F, H, B, R, C, padded, max_tokens_per_msa)
    "msa_small_E128_H2": (128, 512, 2, 2, 6, 20, True, 2 ** 14),
    "msa_mid_E256_H4": (256, 1024, 4, 1, 12, 136, False, 2 ** 10),   # > max_tokens: exercises the reference's chunked path
}


def main():
    for name, (E, Fd, H, B, R, C, padded, mt) in CASES.items():
        sd = make_axial_state_dict(E, Fd, seed=0)
        layer = AxialTransformerLayer(E, Fd, H, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                                      max_tokens_per_msa=mt).eval()
        layer.load_state_dict({k[len("layers.0."):]: v for k, v in sd.items()}, strict=True)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, R, C, E, generator=g)
        mask = None
        if padded:
            mask = torch.zeros(B, R, C, dtype=torch.bool)
            mask[:, :, C - 3:] = True  # the last alignment columns are padding in every row (as MSABatchConverter pads)
            mask[1, R - 2:, :] = True  # and the last rows of the second MSA
        with torch.no_grad():
            xr = x.permute(1, 2, 0, 3).contiguous()  # reference layout [R, C, B, E]
            out, col_attn, row_attn = layer(xr, self_attn_padding_mask=mask, need_head_weights=True)
        # x is regenerated from the seed by the tests (torch.Generator().manual_seed(11), randn(B,R,C,E)): only its checksum is stored
        fx = {"config": {"E": E, "F": Fd, "H": H, "seed": 0, "B": B, "R": R, "C": C, "x_seed": 11},
              "x_checksum": float(x.double().abs().sum()), "mask": mask,
              "out": out.permute(2, 0, 1, 3).contiguous(), "row_attn": row_attn.clone(),
              "col_attn_sample": col_attn[:, :4].clone(),
              "reference": "facebookresearch/esm @ 2b36991 esm.modules.AxialTransformerLayer, torch %s CPU fp32" % torch.__version__}
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
    model_cases()
