"""Developer helper (GPU box): prints the measured error of the CUDA path against the committed reference outputs
(tests/golden) and the CPU oracle, so that the tolerances stated in tests/test_gpu_model.py / DESIGN.md are measured
numbers with margin rather than guesses. Writes gpurun_out/parity_report.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import ESM2  # noqa: E402
from oracle import esm2_oracle  # noqa: E402
from oracle.weights import make_state_dict, make_tokens  # noqa: E402


def metrics(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    d = (got - ref).abs()
    return {"rel_fro": float((got - ref).norm() / ref.norm()), "max_abs": float(d.max()),
            "ref_absmax": float(ref.abs().max()), "ref_rms": float(ref.pow(2).mean().sqrt())}


def main():
    rep = {}
    for name in ["tiny_L2_E128_H2", "mid_L3_E256_H4", "nopad_L2_E128_H2"]:
        fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
        cfg = fx["config"]
        L, E, H = cfg["num_layers"], cfg["embed_dim"], cfg["attention_heads"]
        model = ESM2(L, E, H)
        model.load_state_dict(make_state_dict(L, E, H, seed=cfg["seed"]))
        model = model.eval().cuda()
        out = model(fx["tokens"].cuda(), repr_layers=fx["repr_layers"], return_contacts=True)
        r = {f"repr{k}": metrics(out["representations"][k], v) for k, v in fx["representations"].items()}
        r["logits"] = metrics(out["logits"], fx["logits"])
        r["contacts"] = metrics(out["contacts"], fx["contacts"])
        sub = out["attentions"][:, [0, L - 1]][:, :, [0, H - 1]]
        r["attentions_sub"] = metrics(sub, fx["attentions_sub"])
        rep[name] = r
    for qk_gain in (1.0, 1.5, 3.0):
        L, E, H = 6, 1280, 20
        sd = make_state_dict(L, E, H, seed=0, qk_gain=qk_gain)
        model = ESM2(L, E, H)
        model.load_state_dict(sd)
        model = model.eval().cuda()
        tokens = make_tokens([298, 140, 5], 300, seed=7, n_mask=2)
        ref = esm2_oracle.esm2_forward(sd, L, H, tokens, repr_layers=list(range(L + 1)), need_head_weights=True)
        out = model(tokens.cuda(), repr_layers=list(range(L + 1)), need_head_weights=True)
        r = {f"repr{k}": metrics(out["representations"][k], ref["representations"][k]) for k in range(L + 1)}
        r["logits"] = metrics(out["logits"], ref["logits"])
        r["attentions"] = metrics(out["attentions"], ref["attentions"])
        rep[f"oracle_L6_E1280_qkgain{qk_gain}"] = r
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
    for k, v in rep.items():
        print(k)
        for kk, vv in v.items():
            print("   %-16s rel_fro %.2e  max_abs %.2e  (ref absmax %.2f rms %.3f)" % (kk, vv["rel_fro"], vv["max_abs"], vv["ref_absmax"], vv["ref_rms"]))


if __name__ == "__main__":
    main()
