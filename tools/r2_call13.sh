#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_recurrence.py tests/test_gpu_parity.py tests/test_gpu_varlen.py -x -q > $out/r2c13_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c13_pytest.log
tail -4 $out/r2c13_pytest.log | cut -c1-300
for b in 32 64; do
  timeout 300 python bench.py --batch $b --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=$b', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
CTCASR_LSTM_NO_WIDE=1 timeout 300 python bench.py --batch 64 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B=64 (two 32-row launches)', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
