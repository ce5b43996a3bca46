#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_gemm_pair.py tests/test_gpu_parity.py tests/test_gpu_bf16_oracle.py tests/test_gpu_varlen.py -x -q 2>&1 | tail -6 | cut -c1-300
run() { env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$* $EXTRA', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
EXTRA="" run CTCASR_NARROW_TC=0
EXTRA="" run CTCASR_NARROW_TC=1
EXTRA="--compute bf16" run CTCASR_NARROW_TC=1
