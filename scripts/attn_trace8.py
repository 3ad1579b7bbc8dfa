"""Developer tool: where CTA 0's MMA thread and one softmax warp of attention v8 spend a block (clock64 stamps).
Build:  python -m esm_b200.build -DESMB200_TRACE --out=build_variants/lib_trace.so ;  run with ESMB200_LIB_PATH set to it."""
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
    wide = "d128" in sys.argv  # the two-slot kernel (head_dim 128): S double buffered, QK^T(j+1) issued before P.V(j)
    B, T, H = (16, 1024, 40) if wide else (64, 1024, 20)
    E = (128 if wide else 64) * H
    fn = lib.esmb200_attention128 if wide else lib.esmb200_attention
    qkv = torch.randn(B * T, 3 * E, device=dev)
    qkv[:, :E] *= 0.125
    qkv = qkv.half()
    ctx = torch.empty(B * T, E, dtype=torch.float16, device=dev)
    scratch = torch.empty(lib.esmb200_attention_scratch_bytes(B, T), dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        L.check(fn(P(qkv), None, P(ctx), None, B, T, H, P(scratch), st))
    torch.cuda.synchronize()
    n = 4000
    buf = (ctypes.c_longlong * n)()
    lib.esmb200_debug_read_attn_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int32]
    L.check(lib.esmb200_debug_read_attn_trace(buf, n))
    t = [[buf[s * 400 + i] for i in range(400)] for s in range(10)]
    lo, hi = 40, 270
    blocks = [g for g in range(lo, hi) if g % 16 not in (0, 15)]  # skip tile boundaries (16 blocks per tile at T=1024)

    def avg(f):
        v = [f(g) for g in blocks]
        return sum(v) / len(v)

    nb = max(g for g in range(400) if t[0][g] > 0)
    print("CTA 0: %d blocks in %d cycles = %.0f per block incl. tile boundaries" % (nb, t[0][nb] - t[0][0], (t[0][nb] - t[0][0]) / nb))
    bnd = [t[0][g + 1] - t[0][g] for g in range(15, nb - 1, 16)]
    print("tile-boundary block periods (loop top of block 15 -> loop top of block 0'):", bnd[:8], "avg %.0f" % (sum(bnd) / len(bnd)))
    print("block period (MMA thread, loop top to loop top)      %7.0f" % avg(lambda g: t[0][g + 1] - t[0][g]))
    print("MMA: wait for P_j (loop top -> p_full seen)          %7.0f" % avg(lambda g: t[1][g] - t[0][g]))
    print("MMA: issue P.V(j) 4 MMA + commit                     %7.0f" % avg(lambda g: t[2][g] - t[1][g]))
    if wide:  # order inside an iteration: [wait kv_full(j+1); QK^T(j+1)]; wait P_j; P.V(j)
        print("MMA: loop top -> kv_full(j+1) seen                   %7.0f" % avg(lambda g: t[3][g + 1] - t[2][g - 1]))
        print("MMA: issue QK^T(j+1) 8 MMA + commit                  %7.0f" % avg(lambda g: t[4][g + 1] - t[3][g + 1]))
        print("MMA: QK^T(j+1) issued -> loop top marker             %7.0f" % avg(lambda g: t[0][g] - t[4][g + 1]))
    else:
        print("MMA: wait kv_full(j+1)                               %7.0f" % avg(lambda g: t[3][g + 1] - t[2][g]))
        print("MMA: issue QK^T(j+1) 4 MMA + commit                  %7.0f" % avg(lambda g: t[4][g + 1] - t[3][g + 1]))
    print("soft: block period                                   %7.0f" % avg(lambda g: t[5][g + 1] - t[5][g]))
    print("soft: wait for S_j (block top -> s_full seen)        %7.0f" % avg(lambda g: t[6][g] - t[5][g]))
    print("soft: loads + exponentials (s_full seen -> P ready)  %7.0f" % avg(lambda g: t[8][g] - t[6][g]))
    print("soft: P store + fence + arrive                       %7.0f" % avg(lambda g: t[9][g] - t[8][g]))
    print("QK^T(j+1) issued (MMA) -> S_{j+1} seen by softmax    %7.0f" % avg(lambda g: t[6][g + 1] - t[4][g + 1]))
    print("P_j arrive (softmax) -> p_full seen by the MMA thread %6.0f" % avg(lambda g: t[1][g] - t[9][g]))


if __name__ == "__main__":
    main()
