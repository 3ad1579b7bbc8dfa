#!/bin/bash
# round-2 GPU session L: full suite after the kernel clean-up, configs[3] with the barrier-free fused contact kernel, the bench
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/l_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/l_tests.log | tail -12
cp gpurun_out/l_tests.log gpurun_out/r02_gpu_tests.log
echo "== configs[3]"; C4_PROFILE=1 timeout 600 python scripts/config4_bench.py 2>/dev/null | tail -1; cut -c1-150 gpurun_out/config4_profile.txt | sed -n 4,12p
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/l_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],d['clocks'])
for k,v in d['kernels'].items(): print('  ',k,v)
print('eager',d.get('gpu_eager_baseline'))
c=d.get('configs',{})
print('configs3', c.get('3B_L512_contacts',{}).get('value'), c.get('3B_L512_contacts',{}).get('embed_only',{}).get('value'), 'msa', c.get('msa_128x512',{}).get('ms_per_msa'))
print('cpu',d.get('cpu_baseline'))
PY
