#!/bin/bash
# round-2 GPU session B: whole GPU test suite (new parity / drop-in / narrow-head tests), bench with the extra legs
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/b_tests.log 2>&1; echo "rc=$?"; grep -E "PARITY|passed|failed|Error|error" gpurun_out/b_tests.log | tail -40
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; tail -3 gpurun_out/b_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/b_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],d['clocks'])
for k,v in d['kernels'].items(): print('  ',k,v)
print('roofline',d['roofline'])
print('eager',d.get('gpu_eager_baseline'))
print('configs',json.dumps(d.get('configs'),indent=1))
print('cpu',d.get('cpu_baseline'))
PY
for poly in 0 4; do echo "== bench poly $poly"; ESMB200_ATTN_POLY=$poly timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/b_bench_poly$poly.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b_bench_poly$poly.json')); print(d['value'], d['ms_per_step'], d['kernels']['attention'])"; done
echo "== bench PDL off"; ESMB200_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/b_bench_pdl0.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b_bench_pdl0.json')); print(d['value'], d['ms_per_step'])"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/b_bench_ref.json 2>/dev/null; cut -c1-700 gpurun_out/b_bench_ref.json
