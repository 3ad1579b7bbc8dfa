"""Developer tool: merged timeline of attention v7's three roles (CTA 0) from clock64 stamps.
Build: nvcc ... -DESMB200_TRACE -o esm_b200/libesmb200_trace.so api.cu ; run with ESMB200_ATTN=7
ESMB200_LIB_PATH=esm_b200/libesmb200_trace.so"""
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
    names = ["QK  : loop top", "QK  : waits done (kv_full, pv_done g-2)", "QK  : 4 MMAs + commits issued",
             "PV  : loop top", "PV  : waits done (kv_full, p_full)", "PV  : 4 MMAs + commits issued",
             "soft: block top", "soft: s_full seen", "soft: S in registers", "soft: exps done, before P store"]
    base = t[6][34]
    ev = []
    for g in range(34, 42):
        for s in range(10):
            ev.append((t[s][g] - base, g, names[s]))
    ev.sort()
    for c, g, nm in ev:
        print(f"{c:8d}  g={g:3d}  {nm}")


if __name__ == "__main__":
    main()
