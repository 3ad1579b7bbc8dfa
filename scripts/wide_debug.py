"""Developer tool (GPU box): stress the two-slot (head_dim 128) attention kernel with many tiles per CTA; reports where the
result differs from a PyTorch evaluation.  ESMB200_LIB_PATH selects a library variant."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import _lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


def ref(qkv, B, T, H, D):
    y = qkv.float().view(B, T, 3, H, D)
    q, k, v = (y[:, :, i].transpose(1, 2) for i in range(3))
    p = torch.softmax(q @ k.transpose(-1, -2), -1)
    return (p @ v).transpose(1, 2).reshape(B * T, H * D)


def run(kind, B, T, H, reps=4):
    lib = _lib.load()
    D = 128
    E = D * H
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * T, 3 * E, generator=g)
    if kind == "raise":
        u = torch.randn(D, generator=g)
        u = u / u.norm() * (8.0 ** 0.5)
        blk = (torch.arange(B * T).float() % T / 128).floor()
        for h in range(H):
            qkv[:, h * D:(h + 1) * D] = u + 0.1 * torch.randn(B * T, D, generator=g)
            qkv[:, E + h * D:E + (h + 1) * D] = u * (0.8 * blk[:, None]) + 0.3 * torch.randn(B * T, D, generator=g)
    else:
        qkv[:, :E] *= 0.35
    qkv = qkv.half().cuda()
    want = ref(qkv, B, T, H, D)
    scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for r in range(reps):
        ctx = torch.full((B * T, E), float("nan"), dtype=torch.float16, device="cuda")
        try:
            _lib.check(lib.esmb200_attention128(P(qkv), None, P(ctx), None, B, T, H, P(scratch), st))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(kind, (B, T, H), "rep", r, "CUDA ERROR", str(e)[:200])
            return
        bad = ~torch.isclose(ctx.float(), want, atol=4e-3, rtol=4e-3)
        nan = torch.isnan(ctx.float())
        msg = f"{kind} B{B} T{T} H{H} rep {r}: bad {int(bad.sum())} nan {int(nan.sum())}"
        if bad.any():
            idx = bad.nonzero()
            rows, cols = idx[:, 0], idx[:, 1]
            tiles = sorted({(int(r_) // T, int(c_) // D, (int(r_) % T) // 128) for r_, c_ in zip(rows[:4000].tolist(), cols[:4000].tolist())})
            msg += f" tiles(b,h,qt) {tiles[:6]} rows {int(rows.min())}..{int(rows.max())} cols {int(cols.min())}..{int(cols.max())}"
            w = [(b * H + h) * ((T + 127) // 128) + qt for b, h, qt in tiles[:6]]
            msg += f" work_idx {w}"
        print(msg, flush=True)


if __name__ == "__main__":
    print("lib", _lib.LIB_PATH)
    run("random", 1, 640, 2)
    run("random", 3, 640, 40)
    run("raise", 3, 640, 40)
    run("random", 16, 1024, 40, reps=2)
