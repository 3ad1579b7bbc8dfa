#!/bin/bash
# round-2 GPU session C: the whole GPU suite (all failures listed), then the bench with the extra legs
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/c_tests.log 2>&1; echo "rc=$?"; grep -E "PARITY|passed|failed|FAILED|Error" gpurun_out/c_tests.log | tail -60
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "rc=$?"; tail -3 gpurun_out/c_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],d['clocks'])
for k,v in d['kernels'].items(): print('  ',k,v)
print('eager',d.get('gpu_eager_baseline'))
print('configs',json.dumps(d.get('configs'),indent=1))
print('cpu',d.get('cpu_baseline'))
PY
echo "== PDL A/B in one process"; timeout 600 python scripts/pdl_ab.py 2>&1 | tail -8
