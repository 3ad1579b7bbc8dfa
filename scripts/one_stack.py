"""Developer tool (GPU box, under ncu): a few forwards of the configs[1] model at a small batch (so that an ncu
--set full capture of one layer's kernels is short), or one contacts forward at the 3B width."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import ESM2, pretrained  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "32"
    g = torch.Generator().manual_seed(1)
    if what == "contacts":
        with torch.device("cuda"):
            model = ESM2(num_layers=2, embed_dim=2560, attention_heads=40).eval()
        tok = torch.randint(4, 24, (4, 512), generator=g)
        tok[:, 0] = 0
        tok[:, -1] = 2
        model(tok.cuda(), return_contacts=True)
        torch.cuda.synchronize()
        return
    B = int(what)
    model, _ = pretrained.load_model_and_alphabet("esm2_t33_650M_UR50D", allow_random_init=True, device="cuda")
    tok = torch.randint(4, 24, (B, 1024), generator=g)
    tok[:, 0] = 0
    tok[:, -1] = 2
    tok = tok.cuda()
    for _ in range(3):
        model(tok, repr_layers=[33])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
