#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_varlen.py -x -q > $out/r2c9_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c9_pytest.log
tail -4 $out/r2c9_pytest.log | cut -c1-300
timeout 300 python bench.py --workload ctc --sweep --no-cpu-baseline > $out/r2c9_bench_ctc.json 2> $out/r2c9_bench_ctc.err; python -c "
import json; d=json.load(open('$out/r2c9_bench_ctc.json')); print('ctc', d['ms_per_step'], d['roofline']['frac'])
for r in d.get('sweep', []): print(r['B'], r['T'], round(r['ms'],3), round(r['frac_of_hbm'],4))" || tail -5 $out/r2c9_bench_ctc.err
timeout 600 python bench.py --workload varlen --steps 12 > $out/r2c9_varlen.json 2> $out/r2c9_varlen.err; cat $out/r2c9_varlen.json | cut -c1-1500; tail -5 $out/r2c9_varlen.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_warp -s 1 -c 1 -o $out/r2_ctc_warp2 python tools/profile_target.py ctc > $out/r2c9_ncu_ctc.log 2>&1; tail -1 $out/r2c9_ncu_ctc.log
