#!/bin/bash
# checkpoint of the round: all GPU tests, every bench line, launch list, ncu captures of the kernels changed in this session
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2c22_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c22_pytest.log; tail -3 $out/r2c22_pytest.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > $out/r2c22_bench_$name.json 2> $out/r2c22_bench_$name.err; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$out/r2c22_bench_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],3), round(d['value']), d.get('kernel_ms_per_step'))
except Exception as e: print('$name FAILED', e)"; }
b cfg2
b cfg2_bf16 --compute bf16 --no-cpu-baseline
b ds2 --model ds2 --no-cpu-baseline
b relu --cell rnn_relu --no-cpu-baseline
b tanh --cell rnn_tanh --no-cpu-baseline
b gru --cell gru --no-cpu-baseline
b ctc --workload ctc --sweep
b varlen --workload varlen --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/r2c22_ncu_launches_cfg2_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/r2c22_ncu_bench.log 2>&1; tail -1 $out/r2c22_ncu_launches_cfg2_step.csv | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_tc_pair -s 3 -c 3 -o $out/r2c22_gemm_pair python tools/profile_target.py gemm bf16 > $out/r2c22_ncu_gemm.log 2>&1; tail -1 $out/r2c22_ncu_gemm.log
timeout 300 $NCU -k regex:ctc_warp -s 1 -c 1 -o $out/r2c22_ctc_warp python tools/profile_target.py ctc > $out/r2c22_ncu_ctc.log 2>&1; tail -1 $out/r2c22_ncu_ctc.log
