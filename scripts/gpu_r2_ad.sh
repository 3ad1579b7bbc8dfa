#!/bin/bash
# round-2 GPU session AD (1 GPU): tied (MSA) kernels with warp-convergent control warps — MSA tests, then the final verification script
mkdir -p gpurun_out
echo "== msa tests"; timeout 900 python -m pytest tests/test_gpu_msa.py tests/test_gpu_parity_full.py -q -m gpu -x -k "msa or MSA" > gpurun_out/ad_msa.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/ad_msa.log | tail -5
bash scripts/gpu_r2_final.sh
