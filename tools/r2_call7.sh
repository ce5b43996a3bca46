#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2c7_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2c7_pytest.log
tail -12 $out/r2c7_pytest.log | cut -c1-300
timeout 300 python bench.py --workload ctc --sweep --no-cpu-baseline > $out/r2c7_bench_ctc.json 2> $out/r2c7_bench_ctc.err; python -c "
import json; d=json.load(open('$out/r2c7_bench_ctc.json')); print('ctc', d['ms_per_step'], d['roofline']['frac'])
for r in d.get('sweep', []): print(r['B'], r['T'], round(r['ms'],3), round(r['frac_of_hbm'],4))" || tail -5 $out/r2c7_bench_ctc.err
CTCASR_CTC_BLOCK=1 timeout 300 python bench.py --workload ctc --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('block kernel ctc', d['ms_per_step'])"
