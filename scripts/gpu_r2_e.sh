#!/bin/bash
mkdir -p gpurun_out
echo "== failed tests again"; timeout 1200 python -m pytest tests/test_gpu_precision.py tests/test_gpu_reference_dropin.py tests/test_gpu_parity_full.py tests/test_gpu_kernels.py -q -m gpu -s > gpurun_out/e_tests.log 2>&1; echo "rc=$?"; grep -E "PARITY|passed|failed|FAILED" gpurun_out/e_tests.log | tail -40
bash scripts/gpu_r2_d.sh
