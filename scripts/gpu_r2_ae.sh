#!/bin/bash
# round-2 GPU session AE (1 GPU): ncu evidence for the FINAL build — launch list of the bench command, full in-step capture of one layer
mkdir -p gpurun_out
echo "== launch list (bench, one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
echo "rc=$? lines: $(wc -l < gpurun_out/r02_launches_final.csv)"
echo "== full capture: in-step kernels at B=32 (one layer's worth after warm-up)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm2_f16_kernel|attention_fwd_kernel_v8|layernorm_rows" \
  -s 240 -c 8 -f -o gpurun_out/r02_prof_step_final python scripts/one_stack.py 32 > gpurun_out/r02_prof_step_final.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r02_prof_step_final.ncu-rep
