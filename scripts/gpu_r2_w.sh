#!/bin/bash
# round-2 GPU session W (1 GPU): double-buffered two-slot attention, two-slot fused contact pass, row-chunked embed, new mean_pool
mkdir -p gpurun_out
echo "== kernel + wide-head tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_full.py tests/test_gpu_model.py -q -m gpu -s -x > gpurun_out/w_wide.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|PARITY" gpurun_out/w_wide.log | tail -25
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/w_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/w_tests.log | tail -12
echo "== 15B layer shape speed"; timeout 600 python scripts/wide_head_speed.py 2>&1 | tail -3
echo "== bench (no extras)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/w_bench.json 2>gpurun_out/w_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/w_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['clocks'])
for k in ('embed','mean_pool','attention','gemm_qkv_rope'): print('  ',k,d['kernels'][k])
PY
