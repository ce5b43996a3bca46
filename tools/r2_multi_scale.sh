#!/bin/bash
N=$1; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29503 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > $out/r2m${N}s_bench.json; python -c "import json; d=json.load(open('$out/r2m${N}s_bench.json')); print('bf16x3 N=$N', d['value'], d['ms_per_step'])"
timeout 400 $TR --master-port 29505 bench.py --gpus $N --steps 10 --warmup 3 --compute bf16 --no-cpu-baseline 2>/dev/null | grep '^{' > $out/r2m${N}s_bench_bf16.json; python -c "import json; d=json.load(open('$out/r2m${N}s_bench_bf16.json')); print('bf16 (cfg3) N=$N', d['value'], d['ms_per_step'])"
