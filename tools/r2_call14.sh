#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
run CTCASR_LSTM_PF=0
run CTCASR_LSTM_PF=4
run CTCASR_LSTM_PF=8
run CTCASR_LSTM_PF=14
run CTCASR_LSTM_PF=8 CTCASR_LSTM_L2_KEEP_MB=40
run CTCASR_LSTM_PF=8 CTCASR_LSTM_L2_KEEP_MB=80
timeout 300 python -m pytest tests/test_gpu_recurrence.py -x -q 2>&1 | tail -2
