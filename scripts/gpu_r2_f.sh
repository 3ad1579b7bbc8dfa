#!/bin/bash
mkdir -p gpurun_out
echo "== tests (fused contacts, tolerances)"; timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/f_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/f_tests.log | tail -12
cp gpurun_out/f_tests.log gpurun_out/r02_gpu_tests.log
echo "== configs[3]"; timeout 600 python scripts/config4_bench.py 2>/dev/null | tail -1
echo "== bench chunked=1"; timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels']['gemm_qkv_rope'], d['kernels']['gemm_fc1_gelu'], d['clocks']['sm_mhz'])"
echo "== bench chunked=0"; ESMB200_QKV_CHUNKED=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels']['gemm_qkv_rope'], d['clocks']['sm_mhz'])"
echo "== bench chunked=1 again"; timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels']['gemm_qkv_rope'], d['clocks']['sm_mhz'])"
echo "== fp32x3 speed"; timeout 600 python scripts/precision_speed.py 2>/dev/null | tail -3
