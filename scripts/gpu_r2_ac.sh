#!/bin/bash
# round-2 GPU session AC (1 GPU): two-slot kernel after the issue-order change (stress + trace), then the final verification script
mkdir -p gpurun_out
echo "=== d128 stress"; timeout 300 python scripts/wide_debug.py 2>&1 | grep -v Warning | tail -15
echo "=== trace d128"; ESMB200_LIB_PATH=$PWD/build_variants/lib_trace.so timeout 300 python scripts/attn_trace8.py d128 2>&1 | grep -v Warn | tail -15
timeout 300 python scripts/wide_head_speed.py 2>&1 | tail -2
bash scripts/gpu_r2_final.sh
