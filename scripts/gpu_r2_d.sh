#!/bin/bash
# round-2 GPU session D: ncu evidence — launch list of the bench command and full captures of the hot kernels
mkdir -p gpurun_out
echo "== launch list (bench, one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
echo "rc=$? lines: $(wc -l < gpurun_out/r02_launches.csv)"
echo "== full capture: in-step kernels at B=32 (one layer's worth after warm-up)"
# 1 stack call = key_bits + 33 x 7 launches; skip the first call (warm-up) then capture 8 launches = LN1..fc2 of one layer + LN1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm2_f16_kernel|attention_fwd_kernel_v8|layernorm_rows" \
  -s 240 -c 8 -o gpurun_out/r02_prof_step python scripts/one_stack.py 32 > gpurun_out/r02_prof_step.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r02_prof_step.ncu-rep
echo "== full capture: contact path kernels (3B width, 2 layers)"
timeout 900 ncu --set full --clock-control none -k regex:"attention_probs_kernel|contact_accumulate|contact_finalize" -c 4 \
  -o gpurun_out/r02_prof_contacts python scripts/one_stack.py contacts > gpurun_out/r02_prof_contacts.log 2>&1
echo "rc=$?"
