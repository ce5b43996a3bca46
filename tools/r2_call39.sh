#!/bin/bash
# checkpoint of the round: all GPU tests, every bench line, launch list, ncu captures of the kernels changed in this session
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2c39_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c39_pytest.log; tail -3 $out/r2c39_pytest.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > $out/r2c39_bench_$name.json 2> $out/r2c39_bench_$name.err; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$out/r2c39_bench_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],3), round(d['value']), d.get('kernel_ms_per_step'))
except Exception as e: print('$name FAILED', e)"; }
b cfg2
b cfg2_bf16 --compute bf16 --no-cpu-baseline
b ds2 --model ds2 --no-cpu-baseline
b relu --cell rnn_relu --no-cpu-baseline
b tanh --cell rnn_tanh --no-cpu-baseline
b gru --cell gru --no-cpu-baseline
b ctc --workload ctc --sweep
b varlen --workload varlen --no-cpu-baseline
b ds2_bf16 --model ds2 --compute bf16 --no-cpu-baseline
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
