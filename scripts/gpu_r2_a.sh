#!/bin/bash
# round-2 GPU session A: correctness of the new kernel set, attention A/B, GEMM table, bench with/without PDL
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== attn sweep (in-order WAR)"; timeout 300 python scripts/attn_sweep.py > gpurun_out/a_attn_sweep.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/a_attn_sweep.log
echo "== attn sweep (SAFE_WAR build)"; SWEEP_TAG=safewar ESMB200_LIB_PATH=$PWD/esm_b200/libesmb200_safewar.so timeout 300 python scripts/attn_sweep.py > gpurun_out/a_attn_sweep_safewar.log 2>&1; echo "rc=$?"; grep timing gpurun_out/a_attn_sweep_safewar.log | tail -12
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/a_test_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/a_test_kernels.log
echo "== model + msa tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_msa.py -x -q -m gpu > gpurun_out/a_test_model.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/a_test_model.log
echo "== kernel bench"; KB_B=256 timeout 300 python scripts/kernel_bench.py > gpurun_out/a_kernel_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/a_kernel_bench.log | tail -12
echo "== bench PDL on"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_pdl1.json 2> gpurun_out/a_bench_pdl1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/a_bench_pdl1.json
echo "== bench PDL off"; ESMB200_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_pdl0.json 2> gpurun_out/a_bench_pdl0.err; echo "rc=$?"; cut -c1-600 gpurun_out/a_bench_pdl0.json
echo "== bench v7"; ESMB200_ATTN=7 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_v7.json 2> gpurun_out/a_bench_v7.err; echo "rc=$?"; cut -c1-600 gpurun_out/a_bench_v7.json
