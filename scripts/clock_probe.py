"""Developer tool: run one GEMM epilogue variant for ~2 s and report TFLOP/s together with the SM clock and board power
sampled through NVML while it runs (is the epilogue's cost a clock/power effect?)."""
import ctypes
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esm_b200 import _lib as L  # noqa: E402
import pynvml  # noqa: E402


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def main():
    lib = L.load()
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    M, E, F = 65536, 1280, 5120
    a = torch.randn(M, E, device=dev).half()
    w = (torch.randn(F, E, device=dev) * E ** -0.5).half()
    bias = torch.zeros(F, device=dev)
    out16 = torch.zeros(M, F, dtype=torch.float16, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for epi, name in ((5, "no epilogue"), (9, "tmem loads (batched)"), (10, "loads + gelu math"), (12, "loads + 15 FMA/elem"), (11, "loads + f16 store"),
                      (2, "full gelu epilogue")):
        samples = []
        stop = threading.Event()

        def sampler():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.02)

        def run(n):
            for _ in range(n):
                L.check(lib.esmb200_gemm_f16(epi, P(a), P(w), P(bias), P(out16), M, F, E, None, None, 0, 0, st))
        run(20)
        torch.cuda.synchronize()
        th = threading.Thread(target=sampler)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 2500
        run(n)
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / n
        clk = sorted(s[0] for s in samples)[len(samples) // 2]
        pw = sorted(s[1] for s in samples)[len(samples) // 2]
        tf = 2.0 * M * F * E / ms / 1e9
        print(f"{name:24s} {ms:.4f} ms  {tf:7.1f} TFLOP/s  median SM clock {clk} MHz  power {pw:.0f} W  "
              f"-> {tf / clk * 1000:.1f} TFLOP/s per GHz")


if __name__ == "__main__":
    main()
