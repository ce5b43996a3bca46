#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_model_api.py tests/test_gpu_bf16_oracle.py tests/test_gpu_recurrence.py -x -q -s > $out/r2c5_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c5_pytest.log
grep -E "rounding oracle|passed|failed|Error|error|assert" $out/r2c5_pytest.log | tail -30
timeout 300 python bench.py --cell rnn_relu --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('relu', d['ms_per_step'], d['kernel_ms_per_step'])"
