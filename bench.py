#!/usr/bin/env python
"""bench.py — sequences/sec for ESM-2 650M bulk embedding extraction at L=1024 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

A "step" is one pass of the hot path (embed -> 33 x TransformerLayer -> final LayerNorm -> per-sequence mean, and for
N > 1 one NCCL all-gather of the per-sequence representations) over one synthetic batch:
BASELINE.json configs[1] = esm2_t33_650M_UR50D, 256 sequences of 1024 tokens (<cls> + 1022 residues + <eos>, no
padding, generator seed 1234), seeded random-init weights (checkpoints are unreachable offline).  For N > 1 the same
256-sequence batch is sharded over the ranks (configs[2], strong scaling).

`value`  : device-timed (CUDA events), tokens already resident in HBM, result left in HBM.
`e2e`    : the same workload through the public host-facing call esm_b200.extract.BulkEmbedder.embed(): tokens start
           in pinned HOST memory, per-token [B,T,E] fp32 and per-sequence mean representations end in pinned HOST
           memory; H2D and D2H copies are inside the timed region.
`roofline`: the dominant kernel (by time inside the timed steps, measured with CUDA events on the launch stream via
           esmb200_profile_enable) against the measured cuBLAS bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline`: the reference algorithm (oracle port, PyTorch fp32 ATen ops = what the reference executes) on the
           box's host cores, on a bounded sample (N=1, rank 0 only).
Only the cpu_baseline / --impl reference legs import oracle/.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "esm2_t33_650M_UR50D"
L_LAYERS, E, H, F = 33, 1280, 20, 5120
GLOBAL_BATCH, SEQ_LEN = 256, 1024
TAGS = ["ln1_f16", "gemm_qkv_rope", "attention", "gemm_out_residual", "ln2_f16", "gemm_fc1_gelu", "gemm_fc2_residual",
        "key_bits", "embed", "layernorm_f32", "attention_probs", "convert", "gemm_other", "mean_pool",
        "tied_row_logits", "tied_row_softmax", "tied_row_update"]


def flops_per_seq(T=SEQ_LEN):
    """per layer 8TE^2 + 4T^2E + 4TEF, plus the LM head (not executed for embedding extraction)."""
    return L_LAYERS * (8 * T * E * E + 4 * T * T * E + 4 * T * E * F)


def make_tokens(B, T, seed=1234):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(4, 24, (B, T), generator=g, dtype=torch.int64)
    tok[:, 0] = 0
    tok[:, -1] = 2
    return tok


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tensor_burst": d["bf16_tflops"], "tensor_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "source": "MEASURED_PEAKS.json (of measured)"}
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md (of fallback)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def pick_cpu_threads(state_dict, T=SEQ_LEN):
    """All host threads are available to the CPU arm; PyTorch's intra-op scaling is not monotonic on many-core hosts
    (on the 128-thread B200 host 128 threads run this model SLOWER than 32), so time one TransformerLayer per candidate
    count and keep the fastest — the reference gets its best configuration."""
    from oracle import esm2_oracle
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= c <= ncpu})
    x = torch.randn(1, T, E)
    best, best_t = cands[-1], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            esm2_oracle.transformer_layer(x, state_dict, "layers.0.", H, None, False)
            t0 = time.perf_counter()
            esm2_oracle.transformer_layer(x, state_dict, "layers.0.", H, None, False)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    return best


def cpu_reference_seq_per_s(state_dict, n_seq, steps, warmup, T=SEQ_LEN):
    """The reference algorithm (oracle port) on the host cores; returns (seq/s, ms per step, cores)."""
    from oracle import esm2_oracle
    cores = pick_cpu_threads(state_dict)
    torch.set_num_threads(cores)
    tok = make_tokens(n_seq, T, seed=1234)
    with torch.no_grad():
        esm2_oracle.esm2_forward(state_dict, L_LAYERS, H, tok[:1, :128], repr_layers=[L_LAYERS])  # thread-pool warm-up
        for _ in range(warmup):
            esm2_oracle.esm2_forward(state_dict, L_LAYERS, H, tok, repr_layers=[L_LAYERS])
        t0 = time.perf_counter()
        for _ in range(steps):
            esm2_oracle.esm2_forward(state_dict, L_LAYERS, H, tok, repr_layers=[L_LAYERS])
        dt = time.perf_counter() - t0
    return n_seq * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from esm_b200 import pretrained
    model, _ = pretrained.load_model_and_alphabet(MODEL)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    n_seq = args.ref_seqs
    v, ms, cores = cpu_reference_seq_per_s(sd, n_seq, args.steps, args.warmup)
    sample = (f"{n_seq} of the {GLOBAL_BATCH} sequences (L={SEQ_LEN}) per step, fp32, torch {torch.__version__} CPU, "
              f"{cores} threads (fastest of the counts tried on {os.cpu_count()} logical cores)")
    print(json.dumps({
        "impl": "reference", "metric": "sequences/sec ESM-2 650M L=1024 embedding extract", "value": v,
        "unit": "sequences/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} bulk embedding, batch={GLOBAL_BATCH} synthetic L={SEQ_LEN} (configs[1])",
                   "weights": "seeded random init", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--micro-batch", type=int, default=128)
    ap.add_argument("--ref-seqs", type=int, default=2, help="sequences per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-seqs", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"  # no NCCL version banner on stdout: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    from esm_b200 import _lib, pretrained
    from esm_b200.extract import BulkEmbedder, all_gather_rows, mean_pool, residue_lengths, shard_range
    lib = _lib.load()

    model, alphabet = pretrained.load_model_and_alphabet(MODEL)
    model = model.to(dev)
    tokens_host = make_tokens(args.batch, SEQ_LEN, seed=1234)
    s, e = shard_range(args.batch, world, rank)
    local_host = tokens_host[s:e].contiguous().pin_memory()
    local_dev = local_host.to(dev)
    n_local = e - s

    def step_device():
        # micro-batches keep the workspace at a few GB; per-token output of each micro-batch stays in HBM
        means = []
        for i in range(0, n_local, args.mb_dev):
            tk = local_dev[i:i + args.mb_dev]
            out = model(tk, repr_layers=[L_LAYERS])["representations"][L_LAYERS]
            means.append(mean_pool(out, residue_lengths(tk, alphabet)))
        m = torch.cat(means, 0) if len(means) > 1 else means[0]
        if world > 1:
            m = all_gather_rows(m, args.batch)
        return m

    args.mb_dev = n_local  # one call per step on the device-resident path
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    launches0 = lib.esmb200_launch_count()
    per_step_launches = None
    sampler = ClockSampler(local_rank)
    max_rec = 260 * args.steps
    _lib.check(lib.esmb200_profile_enable(max_rec))
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.esmb200_launch_count() - launches0
    tags = (ctypes.c_int32 * max_rec)()
    mss = (ctypes.c_float * max_rec)()
    nrec = lib.esmb200_profile_read(tags, mss, max_rec)
    _lib.check(lib.esmb200_profile_enable(0))

    # ---- e2e: host tokens -> host representations through the public API
    emb = BulkEmbedder(model, include=("mean", "per_tok"), micro_batch=args.micro_batch)
    for _ in range(2):
        emb.embed(local_host)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = emb.embed(local_host)
        if world > 1:
            all_gather_rows(res["mean"].to(dev, non_blocking=True), args.batch)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = args.batch / (ms_step / 1e3)
    e2e_value = args.batch / (ms_e2e / args.steps / 1e3)

    # ---- per-kernel table + roofline of the dominant kernel
    per = {}
    for i in range(nrec):
        d = per.setdefault(TAGS[tags[i]], [0, 0.0])
        d[0] += 1
        d[1] += mss[i]
    M = n_local * SEQ_LEN
    work = {  # algorithmic FLOPs (tensor) or bytes (hbm) per launch, SURVEY §8(d)
        "gemm_qkv_rope": ("tensor", 2.0 * M * E * 3 * E), "gemm_out_residual": ("tensor", 2.0 * M * E * E),
        "gemm_fc1_gelu": ("tensor", 2.0 * M * E * F), "gemm_fc2_residual": ("tensor", 2.0 * M * F * E),
        "attention": ("tensor", 4.0 * n_local * H * SEQ_LEN * SEQ_LEN * 64),
        "ln1_f16": ("hbm", 6.0 * M * E), "ln2_f16": ("hbm", 6.0 * M * E), "layernorm_f32": ("hbm", 8.0 * M * E),
        "mean_pool": ("hbm", 4.0 * M * E), "embed": ("hbm", 4.0 * M * E),
    }
    peaks = measured_peaks()
    kernels = {}
    for name, (cnt, tot) in per.items():
        avg = tot / cnt
        row = {"launches": cnt, "avg_ms": round(avg, 4), "share": round(tot / ms_total, 4)}
        if name in work:
            kind, amount = work[name]
            if kind == "tensor":
                row["TFLOP/s"] = round(amount / avg / 1e9, 1)
                row["frac_of_peak"] = round(amount / avg / 1e9 / peaks["tensor_sustained"], 3)
            else:
                row["GB/s"] = round(amount / avg / 1e6, 1)
                row["frac_of_peak"] = round(amount / avg / 1e6 / peaks["hbm"], 3)
        kernels[name] = row
    dom = max(per.items(), key=lambda kv: kv[1][1])[0] if per else None
    roofline = None
    if dom and dom in work:
        kind, amount = work[dom]
        avg = per[dom][1] / per[dom][0]
        if kind == "tensor":
            ach, peak, unit = amount / avg / 1e9, peaks["tensor_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = amount / avg / 1e6, peaks["hbm"], "GB/s"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):  # DRAM bytes per token of this kernel from a committed ncu --set full capture
            bpt = json.load(open(tpath))["bytes_per_token"].get(dom)
            traffic = round(bpt * M) if bpt else None
        roofline = {"kernel": dom, "bound": kind, "achieved": round(ach, 1), "peak": peak, "unit": unit,
                    "frac": round(ach / peak, 4), "traffic": traffic,
                    "peak_source": peaks["source"] + (", sustained (kernel timed inside a long step)" if kind == "tensor" else ""),
                    "avg_launch_ms": round(avg, 4), "algorithmic_per_launch": amount}

    out = {
        "metric": "sequences/sec ESM-2 650M L=1024 embedding extract", "value": round(value, 2),
        "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate/residual/LayerNorm/softmax", "data": "synthetic",
        "config": {"workload": f"{MODEL} bulk embedding, batch={args.batch} synthetic L={SEQ_LEN} (BASELINE.json "
                               f"configs[{1 if world == 1 else 2}])",
                   "global_batch": args.batch, "seq_len": SEQ_LEN, "per_gpu_batch": n_local,
                   "parallelism": f"dp{world} (sequence sharding, one all-gather of [B,E] means)",
                   "weights": "seeded random init (no checkpoints offline)", "repr_layers": [L_LAYERS],
                   "l2": "activations per step (>1 GB/GPU) exceed the 126 MB L2; no explicit flush"},
        "model_tflops": round(value * flops_per_seq() / 1e12, 1),
        "tensor_frac_whole_step": round(value * flops_per_seq() / 1e12 / world / peaks["tensor_sustained"], 4),
        "e2e": {"value": round(e2e_value, 2), "unit": "sequences/s", "h2d_bytes_per_step": emb.h2d_bytes * world,
                "d2h_bytes_per_step": emb.d2h_bytes * world, "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": "esm_b200.extract.BulkEmbedder.embed (mean + per_tok to pinned host memory)"},
        "gpu_launches": int(launches), "kernels": kernels, "roofline": roofline, "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        n = args.cpu_baseline_seqs
        v, ms, cores = cpu_reference_seq_per_s(sd, n, steps=1, warmup=0)
        out["cpu_baseline"] = {"value": round(v, 4), "unit": "sequences/s", "cores": cores, "kind": "port",
                               "sample": f"one pass over {n} of the {args.batch} sequences (L={SEQ_LEN}), fp32 oracle "
                                         f"port of the reference on the host CPU, {ms / 1e3:.1f} s"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
