#!/bin/bash
# round 2, call 1: parity of the one-gate persistent kernel + a first rnn_relu bench line
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_recurrence.py -x -q > $out/r2c1_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c1_pytest.log
tail -15 $out/r2c1_pytest.log
timeout 300 python bench.py --cell rnn_relu --steps 4 --warmup 3 --no-cpu-baseline > $out/r2c1_bench_relu.json 2> $out/r2c1_bench_relu.err
cat $out/r2c1_bench_relu.json; tail -3 $out/r2c1_bench_relu.err
timeout 300 python bench.py --cell rnn_tanh --steps 4 --warmup 3 --no-cpu-baseline > $out/r2c1_bench_tanh.json 2> $out/r2c1_bench_tanh.err
cat $out/r2c1_bench_tanh.json; tail -3 $out/r2c1_bench_tanh.err
