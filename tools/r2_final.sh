#!/bin/bash
# final checkpoint of round 2 (1 GPU): all GPU tests, smoke, default bench (+ bf16, ds2), accumulation probe at the shipped
# chunk length, ncu launch list of the default bench command, ncu --set full of one chunked weight-gradient GEMM
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2f_pytest.log; tail -3 $out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
b() { name=$1; shift; timeout 400 python bench.py "$@" > $out/r2f_bench_$name.json 2> $out/r2f_bench_$name.err; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$out/r2f_bench_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],3), round(d['value']), d.get('kernel_ms_per_step'))
except Exception as e: print('$name FAILED', e)"; }
b cfg2
b cfg2_bf16 --compute bf16 --no-cpu-baseline
b ds2 --model ds2 --no-cpu-baseline
timeout 300 python tools/accum_probe.py > $out/r2f_accum.log 2>&1; tail -2 $out/r2f_accum.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r2f_ncu_launches_cfg2_step.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/r2f_ncu_bench.log 2>&1; python tools/launch_summary.py $out/r2f_ncu_launches_cfg2_step.csv 2>&1 | head -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $out/r2f_gemm_wgrad_chained python tools/chain_ubench.py > $out/r2f_ncu_gemm.log 2>&1; ls -la $out/r2f_gemm_wgrad_chained.ncu-rep
