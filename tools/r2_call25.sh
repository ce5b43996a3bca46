#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model_api.py tests/test_gpu_gemm_pair.py -x -q 2>&1 | tail -5 | cut -c1-300
