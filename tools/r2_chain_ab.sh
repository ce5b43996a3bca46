#!/bin/bash
# same-box A/B of the chained-accumulation chunk length on the full cfg2 step
out=gpurun_out
for ch in 0 8192 16384 0 8192 16384; do
  CTCASR_GEMM_CHAIN=$ch timeout 300 python bench.py --no-cpu-baseline --steps 10 > $out/r2ab_$ch.json 2> $out/r2ab_$ch.err
  python -c "
import json
d=json.loads([l for l in open('$out/r2ab_$ch.json') if l.startswith('{')][-1]); print('chain $ch', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['kernel_ms_per_step']['gemm_tc'],2), round(d['kernel_ms_per_step']['rec_fwd']+d['kernel_ms_per_step']['rec_bwd'],2))"
done
