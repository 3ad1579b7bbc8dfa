"""Model factories with the reference's names and return convention (/root/reference/esm/pretrained.py:24-28,
164-183, 344-397): each returns `(model, alphabet)`.

Weights: the reference downloads `https://dl.fbaipublicfiles.com/fair-esm/models/{name}.pt` (pretrained.py:53).
Here a checkpoint is loaded when it is found locally (torch hub cache or an explicit path, same file format and
key-upgrade rule as pretrained.py:164-183); otherwise — there is no network in this environment — the model is
returned with seeded random initialisation and `model.random_init = True` so that callers can tell.
Only head_dim == 64 architectures run on the CUDA path (650M, 3B); the others raise at construction.
"""
from __future__ import annotations

import os
import re
from typing import Optional, Tuple

import torch

from .alphabet import Alphabet
from .model import ESM2

# name -> (num_layers, embed_dim, attention_heads)   (README.md:477-482 + checkpoint cfg)
ESM2_ARCH = {
    "esm2_t6_8M_UR50D": (6, 320, 20),
    "esm2_t12_35M_UR50D": (12, 480, 20),
    "esm2_t30_150M_UR50D": (30, 640, 20),
    "esm2_t33_650M_UR50D": (33, 1280, 20),
    "esm2_t36_3B_UR50D": (36, 2560, 40),
    "esm2_t48_15B_UR50D": (48, 5120, 40),
}


def _hub_path(name: str) -> str:
    return os.path.join(torch.hub.get_dir(), "checkpoints", f"{name}.pt")


def _upgrade_state_dict(state_dict):
    """pretrained.py:164-170: strip the fairseq prefixes."""
    prefixes = ["encoder.sentence_encoder.", "encoder."]
    pattern = re.compile("^" + "|".join(prefixes))
    return {pattern.sub("", k): v for k, v in state_dict.items()}


def load_model_and_alphabet_local(model_location: str) -> Tuple[ESM2, Alphabet]:
    """pretrained.py:67-77 / 164-183 for ESM-2 ("esm2*" file names): reads {"cfg": {"model": ...}, "model": sd}."""
    data = torch.load(str(model_location), map_location="cpu", weights_only=False)
    cfg = data["cfg"]["model"]
    get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    model = ESM2(num_layers=get("encoder_layers"), embed_dim=get("encoder_embed_dim"),
                 attention_heads=get("encoder_attention_heads"), alphabet="ESM-1b",
                 token_dropout=get("token_dropout"))
    sd = _upgrade_state_dict(data["model"])
    reg = str(model_location)[:-3] + "-contact-regression.pt"
    if os.path.exists(reg):
        sd.update(torch.load(reg, map_location="cpu", weights_only=False)["model"])
    model.load_state_dict(sd, strict=os.path.exists(reg))
    model.random_init = False
    return model.eval(), model.alphabet


def load_model_and_alphabet(model_name: str, seed: int = 0) -> Tuple[ESM2, Alphabet]:
    if model_name.endswith(".pt"):
        return load_model_and_alphabet_local(model_name)
    if model_name not in ESM2_ARCH:
        raise ValueError(f"esm_b200 covers the ESM-2 family only; unknown model {model_name!r}")
    path = _hub_path(model_name)
    if os.path.exists(path):
        return load_model_and_alphabet_local(path)
    L, E, H = ESM2_ARCH[model_name]
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = ESM2(num_layers=L, embed_dim=E, attention_heads=H, alphabet="ESM-1b")
    torch.random.set_rng_state(gen_state)
    model.random_init = True
    return model.eval(), model.alphabet


def esm2_t33_650M_UR50D():
    return load_model_and_alphabet("esm2_t33_650M_UR50D")


def esm2_t36_3B_UR50D():
    return load_model_and_alphabet("esm2_t36_3B_UR50D")


def esm2_t6_8M_UR50D():
    return load_model_and_alphabet("esm2_t6_8M_UR50D")


def esm2_t12_35M_UR50D():
    return load_model_and_alphabet("esm2_t12_35M_UR50D")


def esm2_t30_150M_UR50D():
    return load_model_and_alphabet("esm2_t30_150M_UR50D")


def esm2_t48_15B_UR50D():
    return load_model_and_alphabet("esm2_t48_15B_UR50D")
