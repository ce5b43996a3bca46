#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py tests/test_gpu_gemm_pair.py tests/test_gpu_bf16_oracle.py tests/test_golden.py -x -q 2>&1 | tail -6 | cut -c1-300
