#!/bin/bash
# One GPU-box call: parity tests, the bench lines, and the ncu launch list of the bench command.
# Usage (from the repo root, on the box): bash tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench_train.json 2> $out/${tag}_bench_train.err
cat $out/${tag}_bench_train.json
timeout 120 python bench.py --workload ctc > $out/${tag}_bench_ctc.json 2> $out/${tag}_bench_ctc.err
cat $out/${tag}_bench_ctc.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_ncu_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
tail -2 $out/${tag}_ncu_launches.csv
