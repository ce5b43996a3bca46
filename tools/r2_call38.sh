#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_recurrence.py tests/test_gpu_parity.py -x -q -k "one_gate or other_cells or rnn or birnn" 2>&1 | tail -5 | cut -c1-300
run() { env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$* $EXTRA', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
EXTRA="--cell rnn_relu" run A=1
EXTRA="--cell rnn_tanh" run A=1
