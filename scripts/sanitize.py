"""Developer tool (GPU box, under compute-sanitizer): one small invocation of every kernel family — ESM-2 forward with
attentions + contacts (fp16 and fp32x3), a narrow-head model, the MSA axial stack with padding."""
import os
import sys
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import ESM2  # noqa: E402
from esm_b200.msa import MSATransformer  # noqa: E402


def main():
    g = torch.Generator().manual_seed(0)
    with torch.device("cuda"):
        m = ESM2(num_layers=2, embed_dim=256, attention_heads=4).eval()
        m8 = ESM2(num_layers=1, embed_dim=320, attention_heads=20).eval()
        msa = MSATransformer(Namespace(layers=1, embed_dim=128, ffn_embed_dim=512, attention_heads=2, max_positions=1024,
                                       embed_positions_msa=True)).eval()
    tok = torch.randint(4, 24, (3, 150), generator=g)
    tok[:, 0] = 0
    tok[0, -1] = 2
    tok[1, 100] = 2
    tok[1, 101:] = 1
    tok[2, 30] = 2
    tok[2, 31:] = 1
    tok = tok.cuda()
    for prec in ("fp16", "fp32x3"):
        m.set_precision(prec)
        out = m(tok, repr_layers=[0, 1, 2], return_contacts=True)
        assert torch.isfinite(out["contacts"]).all()
    out = m8(tok, repr_layers=[1], need_head_weights=True)
    assert torch.isfinite(out["logits"]).all()
    mt = torch.randint(4, 24, (2, 6, 70), generator=g)
    mt[:, :, 0] = 0
    mt[:, :, 60:] = 1
    mt[1, 4:, :] = 1
    out = msa(mt.cuda(), return_contacts=True)
    torch.cuda.synchronize()
    print("sanitize workload ok")


if __name__ == "__main__":
    main()
