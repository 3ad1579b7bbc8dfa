#!/bin/bash
# round-2 GPU session V (1 GPU): head_dim 128 (two-slot heads) — new tests first, then the whole suite, speed at the 15B layer shape,
# PDL A/B incl. the MSA stack, reference arm
mkdir -p gpurun_out
echo "== wide-head tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_full.py tests/test_gpu_model.py tests/test_gpu_reference_dropin.py -q -m gpu -s -k "128 or wide or 15B or golden or dropin or reference" > gpurun_out/v_wide.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|PARITY" gpurun_out/v_wide.log | tail -25
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/v_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/v_tests.log | tail -12
echo "== 15B layer shape speed"; timeout 600 python scripts/wide_head_speed.py 2>&1 | tail -3
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
echo "== PDL A/B incl. MSA"; timeout 900 python scripts/pdl_ab.py 2>/dev/null | tail -4
