#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_recurrence.py tests/test_gpu_bf16_oracle.py tests/test_gpu_parity.py -x -q > $out/r2c12_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c12_pytest.log
tail -4 $out/r2c12_pytest.log | cut -c1-300
timeout 120 python tools/lstm_trace.py lstm bf16x3 | tail -7 | head -4
timeout 120 python tools/lstm_trace.py rnn_relu bf16x3 | tail -7 | head -4
for args in "--cell lstm" "--cell lstm --compute bf16" "--cell rnn_relu" "--cell gru"; do
  timeout 300 python bench.py $args --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$args', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
