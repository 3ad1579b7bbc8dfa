"""Developer tool (GPU box): speed of the two-slot (head_dim 128) path at the esm2_t48_15B layer shape — a 4-layer
5120 x 40-head model on 16 x 1024 tokens, per-kernel timings from the library's profiler."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import ESM2, _lib  # noqa: E402


def main():
    L, E, H, B, T = 4, 5120, 40, 16, 1024
    if "attn_only" in sys.argv:
        L = 0
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    tok = torch.randint(4, 24, (B, T), generator=g)
    tok[:, 0], tok[:, -1] = 0, 2
    tok = tok.cuda()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if L > 0:
      model = ESM2(num_layers=L, embed_dim=E, attention_heads=H).eval().cuda()
      with torch.no_grad():
        for _ in range(3):
            model(tok, repr_layers=[L])
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            model(tok, repr_layers=[L])
        b.record()
        torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 5
      F = 4 * E
      fl_layer = B * (8 * T * E * E + 4 * T * T * E + 4 * T * E * F)
      res = {"ms_per_forward": round(ms, 3), "layers": L, "model_tflops_layers_only": round(L * fl_layer / ms / 1e9, 1)}
      print(json.dumps(res))
      del model
    # per kernel: attention alone through the standalone entry
    lib = _lib.load()
    qkv = (torch.randn(B * T, 3 * E, device="cuda") * 0.3).half()
    ctx = torch.empty(B * T, E, dtype=torch.float16, device="cuda")
    scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(3):
        _lib.check(lib.esmb200_attention128(P(qkv), None, P(ctx), None, B, T, H, P(scratch), st))
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        _lib.check(lib.esmb200_attention128(P(qkv), None, P(ctx), None, B, T, H, P(scratch), st))
    b.record()
    torch.cuda.synchronize()
    ms_a = a.elapsed_time(b) / 10
    print(json.dumps({"attention128_ms": round(ms_a, 4), "TFLOP/s": round(4 * B * H * T * T * 128 / ms_a / 1e9, 1)}))


if __name__ == "__main__":
    main()
