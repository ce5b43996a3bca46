#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2c6_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c6_pytest.log
tail -12 $out/r2c6_pytest.log
timeout 300 python bench.py --workload ctc --sweep > $out/r2c6_bench_ctc.json 2> $out/r2c6_bench_ctc.err; python -c "
import json; d=json.load(open('$out/r2c6_bench_ctc.json')); print('ctc', d['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline'))
for r in d.get('sweep', []): print(r)" || tail -5 $out/r2c6_bench_ctc.err
timeout 600 python bench.py --steps 4 --warmup 3 > $out/r2c6_bench_train.json 2> $out/r2c6_bench_train.err; cat $out/r2c6_bench_train.json; tail -3 $out/r2c6_bench_train.err
