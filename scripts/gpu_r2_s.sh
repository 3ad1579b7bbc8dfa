#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests (new padding test)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/s_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" gpurun_out/r02_sanitizer_$tool.log | head -12
done
