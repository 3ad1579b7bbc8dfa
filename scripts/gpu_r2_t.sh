#!/bin/bash
# round-2 GPU session T (8 GPUs): same-box N=1 / N=8 / N=4 / N=2 scaling of the contract bench, then the MSA PDL A/B on one GPU
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { n=$1
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/t_bench_n$n.json 2> gpurun_out/t_bench_n$n.err; fi
  echo "N=$n rc=$?"; tail -2 gpurun_out/t_bench_n$n.err | cut -c1-300
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    lines=[l for l in open(f'gpurun_out/t_bench_n{n}.json') if l.startswith('{')]
    d=json.loads(lines[-1])
    print(' N',n,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['clocks'], 'share_sum', round(sum(v.get('share',0) for v in d['kernels'].values()),4), 'stdout_lines', len(open(f'gpurun_out/t_bench_n{n}.json').read().splitlines()))
    for k,v in d['kernels'].items(): print('    ',k,v)
except Exception as e: print(' parse failed',e)
PY
}
run 1; run 8
if [ "$1" != full ]; then exit 0; fi
run 4; run 2
echo "== reference arm under torchrun N=2 (rank 0 only works)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -2 | cut -c1-400
echo "== PDL A/B incl. MSA"; timeout 600 python scripts/pdl_ab.py 2>/dev/null | tail -4
