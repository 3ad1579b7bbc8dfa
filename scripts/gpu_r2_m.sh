#!/bin/bash
mkdir -p gpurun_out
echo "== attn sweep"; SWEEP_TAG=r02 timeout 400 python scripts/attn_sweep.py > gpurun_out/m_attn_sweep.log 2>&1; echo "rc=$?"; grep -E "timing|correctness.*3cta" gpurun_out/m_attn_sweep.log
for c in 4 3 4 3; do echo "== bench attn_ctas=$c"; ESMB200_ATTN_CTAS=$c timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['value'], d['ms_per_step'], 'attn',k['attention']['avg_ms'], d['clocks']['sm_mhz'])"; done
