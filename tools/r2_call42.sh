#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -k "implicit_conv_kernels" 2>&1 | grep -v "^$" | tail -40 | cut -c1-1500 > gpurun_out/r2c42_test_alone.log
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -k "implicit" 2>&1 | grep -v "^$" | tail -40 | cut -c1-1500 > gpurun_out/r2c42_test_pair.log
timeout 300 python tools/r2_diag_conv.py 2>&1 | tail -20 | cut -c1-1200 > gpurun_out/r2c42_diag.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/r2_diag_conv.py one 2>&1 | tail -60 | cut -c1-400 > gpurun_out/r2c42_memcheck.log
tail -5 gpurun_out/r2c42_test_alone.log
