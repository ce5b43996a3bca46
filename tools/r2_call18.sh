#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gemm_pair.py -x -q -s 2>&1 | tail -8 | cut -c1-300
for p in 0 1; do PROBE_MODE=bf16 CTCASR_GEMM_PAIR=$p timeout 200 python tools/gemm_shapes.py 2>&1 | tail -8; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16_oracle.py -x -q 2>&1 | tail -4 | cut -c1-300
