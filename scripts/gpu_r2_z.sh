#!/bin/bash
# round-2 GPU session Z (1 GPU): attention v9 (S released early, P in its own TMEM columns, 3 CTAs/SM) vs v8
mkdir -p gpurun_out
echo "== sweep"; SWEEP_TAG=v9 timeout 600 python scripts/attn_sweep.py 2>&1 | grep -v Warn | tail -24
echo "== kernel + model tests with ESMB200_ATTN=9"; ESMB200_ATTN=9 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_parity_full.py -q -m gpu -x > gpurun_out/z_tests9.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/z_tests9.log | tail -6
for v in 8 9; do
echo "== bench attn=$v"; ESMB200_ATTN=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/z_bench$v.json 2>gpurun_out/z_bench$v.err; python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/z_bench{v}.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['clocks'])
for k in ('attention','gemm_qkv_rope','gemm_fc1_gelu'): print('  ',k,d['kernels'][k])
PY
done
