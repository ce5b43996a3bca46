#!/bin/bash
N=$1; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for c in 2 4 8; do
  NCCL_MAX_CTAS=$c timeout 600 $TR --master-port 2951$c bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('overlap NCCL_MAX_CTAS=$c N=$N', round(d['ms_per_step'],2), d['kernel_ms_per_step'])"
done
timeout 600 $TR --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-overlap 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no-overlap N=$N', round(d['ms_per_step'],2), d['kernel_ms_per_step'])"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --check 2>/dev/null | grep '^{'
