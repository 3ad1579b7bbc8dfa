#!/bin/bash
# round-2 final verification (1 GPU): what the driver runs at round end — full GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/final_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/final_tests.log | tail -8
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],d['clocks'])
for k,v in d['kernels'].items(): print('  ',k,v)
print('roofline',d['roofline'])
print('eager',d.get('gpu_eager_baseline'))
c=d.get('configs',{})
print('configs3', c.get('3B_L512_contacts',{}).get('value'), c.get('3B_L512_contacts',{}).get('embed_only',{}).get('value'), 'msa', c.get('msa_128x512',{}).get('ms_per_msa'))
print('cpu',d.get('cpu_baseline'))
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
