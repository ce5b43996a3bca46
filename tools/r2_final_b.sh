#!/bin/bash
# final commit: the reference's default cell (rnn_relu) and the cfg4 line
out=gpurun_out
timeout 200 python bench.py --cell rnn_relu --no-cpu-baseline > $out/r2f_bench_relu.json 2> $out/r2f_bench_relu.err
timeout 200 python bench.py --workload varlen --no-cpu-baseline > $out/r2f_bench_varlen.json 2> $out/r2f_bench_varlen.err
python - <<'PY'
import json
for n in ("relu", "varlen"):
    try:
        d = json.loads([l for l in open("gpurun_out/r2f_bench_%s.json" % n) if l.startswith("{")][-1]); print(n, round(d["ms_per_step"], 2), round(d["value"]))
    except Exception as e:
        print(n, "failed", e)
PY
