#!/bin/bash
# round-2 GPU session Z2 (1 GPU): ncu --set full of attention v9 and v8 (stand-alone, B=64) for a side-by-side
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel_v9 -s 12 -c 1 -o gpurun_out/r02_prof_attn_v9 -f python scripts/attn_sweep.py > gpurun_out/z2_ncu9.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel_v8 -s 12 -c 1 -o gpurun_out/r02_prof_attn_v8 -f python scripts/attn_sweep.py > gpurun_out/z2_ncu8.log 2>&1; echo "rc=$?"
for v in 8 9; do ncu -i gpurun_out/r02_prof_attn_v$v.ncu-rep --page raw --csv > gpurun_out/r02_prof_attn_v${v}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r02_prof_attn_v*
