#!/bin/bash
# 2 GPUs: the multi-rank extraction test and the N=2 bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== multirank extract test"; timeout 900 python -m pytest tests/test_gpu_multirank_extract.py -q -m gpu > gpurun_out/n2_test.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/n2_test.log
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/n2_bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'], d['clocks'])"
echo "== bench N=2 reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-300
echo "== bench N=1 on the same box"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
