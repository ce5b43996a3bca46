#!/bin/bash
# usage: bash tools/r2_multi.sh N   (on a box with N GPUs)
N=$1; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > $out/r2m${N}_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2m${N}_pytest.log; tail -6 $out/r2m${N}_pytest.log | cut -c1-400
fi
timeout 600 $TR --master-port 29501 bench.py --gpus $N --check > $out/r2m${N}_check.json 2> $out/r2m${N}_check.err; cat $out/r2m${N}_check.json; tail -2 $out/r2m${N}_check.err
timeout 300 $TR --master-port 29502 tools/nccl_busbw.py > $out/r2m${N}_busbw.json 2> $out/r2m${N}_busbw.err; cat $out/r2m${N}_busbw.json
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 10 --warmup 3 > $out/r2m${N}_bench.json 2> $out/r2m${N}_bench.err; python -c "import json; d=json.load(open('$out/r2m${N}_bench.json')); print('bf16x3 N=$N', d['value'], d['ms_per_step'])" || tail -3 $out/r2m${N}_bench.err
timeout 600 $TR --master-port 29504 bench.py --gpus $N --steps 10 --warmup 3 --no-overlap > $out/r2m${N}_bench_nooverlap.json 2> $out/r2m${N}_bench_nooverlap.err; python -c "import json; d=json.load(open('$out/r2m${N}_bench_nooverlap.json')); print('bf16x3 no-overlap N=$N', d['value'], d['ms_per_step'])"
timeout 600 $TR --master-port 29505 bench.py --gpus $N --steps 10 --warmup 3 --compute bf16 > $out/r2m${N}_bench_bf16.json 2> $out/r2m${N}_bench_bf16.err; python -c "import json; d=json.load(open('$out/r2m${N}_bench_bf16.json')); print('bf16 N=$N', d['value'], d['ms_per_step'])"
