#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -s -k "implicit" 2>&1 | tail -6 | cut -c1-400
