"""Developer tool: per-phase timeline of the attention kernel (CTA 0) from clock64 stamps.
Build first: nvcc ... -DESMB200_TRACE -o /tmp/libesmb200_trace.so api.cu ; run with ESMB200_LIB_PATH=/tmp/libesmb200_trace.so"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import _lib as L  # noqa: E402


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def main():
    lib = L.load()
    dev = torch.device("cuda:0")
    B, T, H = 64, 1024, 20
    E = 64 * H
    qkv = torch.randn(B * T, 3 * E, device=dev)
    qkv[:, :E] *= 0.125
    qkv = qkv.half()
    ctx = torch.empty(B * T, E, dtype=torch.float16, device=dev)
    scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        L.check(lib.esmb200_attention(P(qkv), None, P(ctx), None, B, T, H, P(scratch), st))
    torch.cuda.synchronize()
    n = 4000
    buf = (ctypes.c_longlong * n)()
    lib.esmb200_debug_read_attn_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int32]
    L.check(lib.esmb200_debug_read_attn_trace(buf, n))
    t = [[buf[s * 400 + i] for i in range(400)] for s in range(10)]
    names = {0: "mma: before wait p_full", 1: "mma: p_full seen", 2: "mma: PV issued (+QK ahead next iter)",
             4: "sm: block start", 5: "sm: s_full seen", 6: "sm: pv_done(g-2) seen", 7: "sm: S in registers",
             8: "sm: exps + P stores issued", 9: "sm: s_free/p_full arrived"}
    base = t[4][32]
    print("softmax warp 2, CTA 0, blocks 32..47 (cycles relative to block 32 start):")
    for g in range(32, 48):
        row = [t[s][g] - base for s in (4, 5, 6, 7, 8, 9)]
        d = [row[i + 1] - row[i] for i in range(5)]
        print(f"g={g:3d} start {row[0]:7d} | wait s_full {d[0]:5d} | wait pv_done {d[1]:5d} | tmem ld+wait {d[2]:5d} | exp pass {d[3]:5d} | arrive {d[4]:5d} | block total {t[4][g + 1] - t[4][g]:6d}")
    print("mma thread:")
    for g in range(32, 48):
        print(f"g={g:3d} wait p_full {t[1][g] - t[0][g]:6d} | issue PV {t[2][g] - t[1][g]:5d} | loop total {t[0][g + 1] - t[0][g]:6d}")


if __name__ == "__main__":
    main()
