#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_recurrence.py tests/test_gpu_bf16_oracle.py -x -q > $out/r2c11_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c11_pytest.log
tail -4 $out/r2c11_pytest.log | cut -c1-300
timeout 120 python tools/lstm_trace.py lstm bf16 | tail -7
for args in "--cell lstm --compute bf16" "--cell gru --compute bf16"; do
  timeout 300 python bench.py $args --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$args', round(d['ms_per_step'],2), d['kernel_ms_per_step'])"
done
