#!/bin/bash
N=$1; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29502 tools/nccl_busbw.py 2>/dev/null | grep '^{' > $out/r2m${N}_busbw.json; cat $out/r2m${N}_busbw.json
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | grep '^{' > $out/r2m${N}_bench.json; python -c "import json; d=json.load(open('$out/r2m${N}_bench.json')); print('bf16x3 N=$N', d['value'], d['ms_per_step'])"
timeout 600 $TR --master-port 29504 bench.py --gpus $N --steps 10 --warmup 3 --overlap 2>/dev/null | grep '^{' > $out/r2m${N}_bench_overlap.json; python -c "import json; d=json.load(open('$out/r2m${N}_bench_overlap.json')); print('bf16x3 overlap N=$N', d['value'], d['ms_per_step'])"
timeout 600 $TR --master-port 29505 bench.py --gpus $N --steps 10 --warmup 3 --compute bf16 2>/dev/null | grep '^{' > $out/r2m${N}_bench_bf16.json; python -c "import json; d=json.load(open('$out/r2m${N}_bench_bf16.json')); print('bf16 (cfg3) N=$N', d['value'], d['ms_per_step'])"
timeout 600 $TR --master-port 29506 bench.py --gpus $N --check 2>/dev/null | grep '^{' > $out/r2m${N}_check.json; cat $out/r2m${N}_check.json
timeout 900 $TR --master-port 29507 bench.py --gpus $N --workload varlen 2>/dev/null | grep '^{' > $out/r2m${N}_varlen.json; python -c "import json; d=json.load(open('$out/r2m${N}_varlen.json')); print('varlen N=$N', d['value'], d['ms_per_step'], d['steps'])"
