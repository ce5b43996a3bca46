#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -15 | cut -c1-300
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$* $EXTRA', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
EXTRA="--model ds2" run CTCASR_CONV_IMPLICIT_DGRAD=0
EXTRA="--model ds2" run CTCASR_CONV_IMPLICIT_DGRAD=1
