#!/bin/bash
# CTA-pair GEMM: parity + effect on the cfg2 / cfg3 step
timeout 600 python -m pytest tests/test_gpu_gemm_pair.py -x -q -s 2>&1 | tail -15 | cut -c1-300
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$* $EXTRA', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
EXTRA="--compute bf16" run CTCASR_GEMM_PAIR=0
EXTRA="--compute bf16" run CTCASR_GEMM_PAIR=1
EXTRA="" run CTCASR_GEMM_PAIR=0
EXTRA="" run CTCASR_GEMM_PAIR=1
