#!/bin/bash
# round-2 GPU session AB (1 GPU): attention v8 timeline of the MMA thread / a softmax warp, sweep, tests
mkdir -p gpurun_out
echo "=== trace"; ESMB200_LIB_PATH=$PWD/build_variants/lib_trace.so timeout 300 python scripts/attn_trace8.py 2>&1 | grep -v Warn | tail -14
echo "=== trace d128"; ESMB200_LIB_PATH=$PWD/build_variants/lib_trace.so timeout 300 python scripts/attn_trace8.py d128 2>&1 | grep -v Warn | tail -15
SWEEP_TAG=chain timeout 600 python scripts/attn_sweep.py 2>&1 | grep -v Warn | grep -E "timing v|gain4.0" | tail -16
echo "== kernel + model tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_parity_full.py tests/test_gpu_precision.py tests/test_gpu_msa.py -q -m gpu -x > gpurun_out/ab_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/ab_tests.log | tail -6
