#!/bin/bash
out=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r2c30_ncu_launches_ds2_step.csv python bench.py --model ds2 --steps 1 --warmup 3 --no-cpu-baseline > $out/r2c30_ncu_bench.log 2>&1; tail -1 $out/r2c30_ncu_launches_ds2_step.csv | cut -c1-100
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_conv.py -x -q -k "conv2d_layer or implicit" > $out/r2c30_memcheck_conv.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" $out/r2c30_memcheck_conv.log | tr '\n' ' '
