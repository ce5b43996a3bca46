#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 60 tools/ubench/mma_loop | head -3
timeout 900 python -m pytest tests/test_gpu_recurrence.py tests/test_gpu_parity.py -x -q > $out/r2c4_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c4_pytest.log
tail -4 $out/r2c4_pytest.log
for args in "--cell lstm" "--cell lstm --compute bf16" "--cell rnn_relu" "--cell gru"; do
  tag=$(echo $args | tr -d ' -')
  timeout 300 python bench.py $args --steps 4 --warmup 3 --no-cpu-baseline > $out/r2c4_bench_$tag.json 2> $out/r2c4_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2c4_bench_$tag.json")); print("$args", d["ms_per_step"], d["kernel_ms_per_step"])
except Exception as e:
    print("$args failed", e); print(open("$out/r2c4_bench_$tag.err").read()[-1500:])
PY
done
