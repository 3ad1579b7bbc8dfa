#!/bin/bash
# round-2 GPU session AA (1 GPU): MMA-issuer role swap across co-resident CTAs (A/B against a build without it)
mkdir -p gpurun_out
for v in "" build_variants/lib_noswap.so; do
  echo "=== variant '${v}'"
  if [ -n "$v" ]; then export ESMB200_LIB_PATH=$PWD/$v; else unset ESMB200_LIB_PATH; fi
  SWEEP_TAG=swap timeout 600 python scripts/attn_sweep.py 2>&1 | grep -v Warn | grep -E "timing|gain4" | tail -12
  timeout 300 python scripts/wide_head_speed.py attn_only 2>&1 | tail -1
done
unset ESMB200_LIB_PATH
echo "== kernel + model tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_parity_full.py tests/test_gpu_precision.py -q -m gpu -x > gpurun_out/aa_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/aa_tests.log | tail -6
