#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_beam.py -x -q 2>&1 | tail -2 | cut -c1-200
run() { env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$* $EXTRA', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
EXTRA="" run CTCASR_GEMM_STREAM_C=0
EXTRA="" run CTCASR_GEMM_STREAM_C=1
EXTRA="--compute bf16" run CTCASR_GEMM_STREAM_C=0
EXTRA="--compute bf16" run CTCASR_GEMM_STREAM_C=1
