"""Developer tool (GPU box): attention kernel A/B — v7 vs v8 x FMA-pipe exponential share, correctness against a PyTorch
fp32 evaluation on a ragged batch and timing at the configs[1] shape.  Writes gpurun_out/attn_sweep.json."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import _lib as L  # noqa: E402


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def reference(qkv, lens, B, T, H):
    E = 64 * H
    q, k, v = (qkv[:, i * E:(i + 1) * E].float().view(B, T, H, 64).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2)
    key = torch.arange(T, device=qkv.device)[None, :] < lens[:, None]
    s = s.masked_fill(~key[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ v).transpose(1, 2).reshape(B * T, E)


def main():
    lib = L.load()
    dev = torch.device("cuda:0")
    H = 20
    E = 64 * H
    out = {"correctness": {}, "timing": {}}
    variants = [("v7", 7, 3, 4)] + [(f"v8_poly{p}", 8, p, 4) for p in (0, 2, 3, 4)]
    # ---- correctness: ragged batch, sharp logits (gain), partial last blocks
    g = torch.Generator().manual_seed(3)
    for gain in (1.0, 4.0):
        B, T = 5, 333
        lens = torch.tensor([333, 200, 64, 1, 129], device=dev)
        qkv = torch.randn(B * T, 3 * E, generator=g).to(dev)
        qkv[:, :E] *= 0.125 * gain
        qkv[:, E:2 * E] *= gain
        qkv = qkv.half()
        mask = (torch.arange(T, device=dev)[None, :] >= lens[:, None]).to(torch.uint8).contiguous()
        ref = reference(qkv, lens, B, T, H)
        scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device=dev)
        for name, ver, poly, ctas in variants:
            L.check(lib.esmb200_set_option(b"attn", ver))
            L.check(lib.esmb200_set_option(b"attn_poly", poly))
            ctx = torch.zeros(B * T, E, dtype=torch.float16, device=dev)
            L.check(lib.esmb200_attention(P(qkv), P(mask), P(ctx), None, B, T, H, P(scratch), S()))
            torch.cuda.synchronize()
            valid = (torch.arange(T, device=dev)[None, :] < lens[:, None]).reshape(-1)
            err = (ctx.float() - ref)[valid].abs().max().item()
            out["correctness"][f"{name}_gain{gain}"] = err
            print("correctness", name, gain, err, flush=True)
    # ---- timing at B=64 and B=256, T=1024
    for B in (64, 256):
        T = 1024
        qkv = torch.randn(B * T, 3 * E, device=dev)
        qkv[:, :E] *= 0.125
        qkv = qkv.half()
        ctx = torch.empty(B * T, E, dtype=torch.float16, device=dev)
        scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device=dev)
        for name, ver, poly, ctas in variants:
            L.check(lib.esmb200_set_option(b"attn", ver))
            L.check(lib.esmb200_set_option(b"attn_poly", poly))
            ms = timeit(lambda: L.check(lib.esmb200_attention(P(qkv), None, P(ctx), None, B, T, H, P(scratch), S())))
            tf = 4.0 * B * H * T * T * 64 / ms / 1e9
            out["timing"][f"{name}_B{B}"] = {"ms": ms, "TFLOP/s": tf}
            print("timing", name, B, round(ms, 4), round(tf, 1), flush=True)
    # MSA column-attention shape: 512 sequences x 128 tokens x 12 heads
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("SWEEP_TAG", "default")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"attn_sweep_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
