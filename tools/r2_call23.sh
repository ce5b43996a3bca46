#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -x -q -k "dropout or ds2 or conv2d" 2>&1 | tail -12 | cut -c1-400
