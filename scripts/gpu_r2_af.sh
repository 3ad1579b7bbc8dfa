#!/bin/bash
# round-2 GPU session AF (1 GPU): spin vs try_wait for the MMA warp's wait on P (in-step A/B: the spinning warp takes issue slots)
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" build_variants/lib_nospin.so; do
  if [ -n "$v" ]; then export ESMB200_LIB_PATH=$PWD/$v; else unset ESMB200_LIB_PATH; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/af_bench.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/af_bench.json') if l.startswith('{')][-1])
print(f"variant '{sys.argv[1]}': value {d['value']} ms {d['ms_per_step']} attention {d['kernels']['attention']['avg_ms']} ms ({d['kernels']['attention']['frac_of_peak']}) clocks {d['clocks']['sm_mhz']}")
PY
done; done
unset ESMB200_LIB_PATH
SWEEP_TAG=spin timeout 600 python scripts/attn_sweep.py 2>&1 | grep -E "timing v8_poly(0|4)" 
ESMB200_LIB_PATH=$PWD/build_variants/lib_nospin.so SWEEP_TAG=nospin timeout 600 python scripts/attn_sweep.py 2>&1 | grep -E "timing v8_poly(0|4)"
