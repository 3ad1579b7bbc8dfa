#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -q -m gpu -x > gpurun_out/j_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/j_tests.log | tail -12
for f in 1 0; do echo "== bench B=32 fuse_ln=$f"; ESMB200_FUSE_LN=$f timeout 600 python bench.py --batch 32 --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['value'], d['ms_per_step'], 'out',k['gemm_out_residual']['avg_ms'],'fc2',k['gemm_fc2_residual']['avg_ms'])"; done
for f in 1 0 1 0; do echo "== bench fuse_ln=$f"; ESMB200_FUSE_LN=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['value'], d['ms_per_step'], 'out',k['gemm_out_residual']['avg_ms'],'fc2',k['gemm_fc2_residual']['avg_ms'],'ln1',k['ln1_f16']['avg_ms'],k['ln1_f16']['launches'],'ln2',k.get('ln2_f16',{}).get('avg_ms'), d['clocks']['sm_mhz'])"; done
echo "== configs[3]"; C4_PROFILE=1 timeout 600 python scripts/config4_bench.py 2>/dev/null | tail -1; cut -c1-150 gpurun_out/config4_profile.txt | sed -n 4,10p
