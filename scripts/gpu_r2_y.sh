#!/bin/bash
# round-2 GPU session Y (1 GPU): ncu --set full of the two-slot attention kernel at the 15B shape
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel_v8 -s 2 -c 1 -o gpurun_out/r02_prof_attn128 -f python scripts/wide_head_speed.py attn_only > gpurun_out/y_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/y_ncu.log
ncu -i gpurun_out/r02_prof_attn128.ncu-rep --page raw --csv > gpurun_out/r02_prof_attn128_raw.csv 2>/dev/null; wc -c gpurun_out/r02_prof_attn128_raw.csv
