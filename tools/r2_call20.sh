#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_model_api.py -x -q -k "ctc or golden or label or status" 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --workload ctc --steps 8 --warmup 3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ctc ms', round(d['ms_per_step'],4), d['roofline'])"
