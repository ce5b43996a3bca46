#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_beam.py tests/test_golden.py -x -q -s 2>&1 | tail -15 | cut -c1-300
