"""Developer tool (GPU box): cost of the fp32x3 precision mode — the configs[1] model, 64 x 1024 tokens, both modes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import pretrained  # noqa: E402


def main():
    model, _ = pretrained.load_model_and_alphabet("esm2_t33_650M_UR50D", allow_random_init=True, device="cuda")
    g = torch.Generator().manual_seed(1)
    B = 64
    tok = torch.randint(4, 24, (B, 1024), generator=g)
    tok[:, 0] = 0
    tok[:, -1] = 2
    tok = tok.cuda()
    for prec in ("fp16", "fp32x3"):
        model.set_precision(prec)
        for _ in range(2):
            model(tok, repr_layers=[33])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            model(tok, repr_layers=[33])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{prec}: {ms:.1f} ms per {B} x 1024 tokens = {B / ms * 1e3:.1f} sequences/s")


if __name__ == "__main__":
    main()
