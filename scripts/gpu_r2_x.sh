#!/bin/bash
# round-2 GPU session X (1 GPU): multi-tile stress of the double-buffered two-slot attention kernel (early vs late QK^T issue), tests, speed
mkdir -p gpurun_out
for v in "" build_variants/lib_late.so; do
  echo "=== variant '${v}'"
  if [ -n "$v" ]; then export ESMB200_LIB_PATH=$PWD/$v; else unset ESMB200_LIB_PATH; fi
  timeout 300 python scripts/wide_debug.py 2>&1 | grep -v Warning | tail -16
  timeout 300 python scripts/wide_head_speed.py 2>&1 | tail -2
done
unset ESMB200_LIB_PATH
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/x_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|PARITY reference_eager" gpurun_out/x_tests.log | tail -12
