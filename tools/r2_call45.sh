#!/bin/bash
for ch in 0 4096 8192; do CTCASR_GEMM_CHAIN=$ch timeout 200 python tools/chain_ubench.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_pair.py tests/test_gpu_conv.py -x -q -k "gemm or contraction or conv or dense" 2>&1 | tail -3 | cut -c1-400
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err
CTCASR_GEMM_CHAIN=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c45_bench_nochain.json 2> gpurun_out/r2c45_bench_nochain.err
python - <<'PY'
import json
for n in ("", "_nochain"):
    try:
        d = json.load(open("gpurun_out/r2c45_bench%s.json" % n))
        print(n or "chain", d["ms_per_step"], d.get("kernel_ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
PY
