#!/bin/bash
# 2 x B200 at the final commit: data-parallel correctness check and one cfg2 bench line
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29506 bench.py --gpus 2 --check 2>$out/r2f_n2_check.err | grep '^{' > $out/r2f_n2_check.json; cut -c1-900 $out/r2f_n2_check.json
timeout 300 $TR --master-port 29503 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>$out/r2f_n2_bench.err | grep '^{' > $out/r2f_n2_bench.json; python -c "import json; d=json.load(open('$out/r2f_n2_bench.json')); print('bf16x3 N=2', d['value'], d['ms_per_step'])"
