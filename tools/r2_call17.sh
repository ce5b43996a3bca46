#!/bin/bash
for m in bf16 bf16x3; do for p in 0 1; do PROBE_MODE=$m CTCASR_GEMM_PAIR=$p timeout 200 python tools/gemm_shapes.py 2>&1 | tail -8; done; done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv | tail -1
