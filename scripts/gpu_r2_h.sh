#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 8 12 15; do echo "== LN_DEBUG=$dbg"; ESMB200_LN_DEBUG=$dbg ESMB200_FUSE_LN=1 timeout 300 python bench.py --batch 32 --steps 2 --warmup 1 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['ms_per_step'], 'out',k['gemm_out_residual']['avg_ms'],'fc2',k['gemm_fc2_residual']['avg_ms'])"; done
echo "== fuse off"; ESMB200_FUSE_LN=0 timeout 300 python bench.py --batch 32 --steps 2 --warmup 1 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['ms_per_step'], 'out',k['gemm_out_residual']['avg_ms'],'fc2',k['gemm_fc2_residual']['avg_ms'])"
echo "== ncu fused contact kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_probs_contact" -c 1 -o gpurun_out/r02_prof_cfuse python scripts/one_stack.py contacts > gpurun_out/r02_prof_cfuse.log 2>&1; echo "rc=$?"
