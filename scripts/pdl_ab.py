"""Developer tool (GPU box): programmatic dependent launch on/off, alternated inside ONE process on the same model and
tokens (box-to-box and run-to-run clock differences are larger than the effect), at two per-GPU batch sizes
(256 = configs[1] on one GPU, 32 = the per-GPU share at 8 GPUs)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import _lib, pretrained  # noqa: E402


def main():
    lib = _lib.load()
    model, _ = pretrained.load_model_and_alphabet("esm2_t33_650M_UR50D", allow_random_init=True, device="cuda")
    res = {}
    for B in (32, 256):
        g = torch.Generator().manual_seed(1)
        tok = torch.randint(4, 24, (B, 1024), generator=g)
        tok[:, 0] = 0
        tok[:, -1] = 2
        tok = tok.cuda()
        times = {0: [], 1: []}
        for rep in range(4):
            for pdl in (1, 0):
                _lib.check(lib.esmb200_set_option(b"pdl", pdl))
                for _ in range(2):
                    model(tok, repr_layers=[33])
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 6 if B == 32 else 3
                a.record()
                for _ in range(n):
                    model(tok, repr_layers=[33])
                b.record()
                torch.cuda.synchronize()
                times[pdl].append(a.elapsed_time(b) / n)
        res[f"B{B}"] = {"pdl_on_ms": sorted(times[1]), "pdl_off_ms": sorted(times[0])}
        print(f"B={B}: PDL on {min(times[1]):.3f} ms (median {sorted(times[1])[2]:.3f}), off {min(times[0]):.3f} ms "
              f"(median {sorted(times[0])[2]:.3f})")
    del model
    # configs[4]: the MSA axial stack is ~180 SHORT, non-persistent launches per forward -- the one place a tail exists
    msa, _ = pretrained.load_msa_model_and_alphabet("esm_msa1b_t12_100M_UR50S", allow_random_init=True, device="cuda")
    g = torch.Generator().manual_seed(2)
    mtok = torch.randint(4, 24, (1, 128, 512), generator=g)
    mtok[:, :, 0] = 0
    mtok = mtok.cuda()
    times = {0: [], 1: []}
    for rep in range(4):
        for pdl in (1, 0):
            _lib.check(lib.esmb200_set_option(b"pdl", pdl))
            for _ in range(2):
                msa(mtok, repr_layers=[12])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                msa(mtok, repr_layers=[12])
            b.record()
            torch.cuda.synchronize()
            times[pdl].append(a.elapsed_time(b) / 5)
    res["msa_128x512"] = {"pdl_on_ms": sorted(times[1]), "pdl_off_ms": sorted(times[0])}
    print(f"MSA 128x512: PDL on {min(times[1]):.3f} ms (median {sorted(times[1])[2]:.3f}), off {min(times[0]):.3f} ms "
          f"(median {sorted(times[0])[2]:.3f})")
    _lib.check(lib.esmb200_set_option(b"pdl", 0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pdl_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
