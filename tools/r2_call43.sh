#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_parity.py -x -q -s -k "implicit_conv_kernels or contraction_at_cfg2" 2>&1 | grep -v "^$" | tail -12 | cut -c1-900
timeout 600 python tools/accum_probe.py 2>&1 | tail -14 | cut -c1-1200
