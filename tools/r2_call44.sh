#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_pair.py tests/test_gpu_conv.py tests/test_gpu_bf16_oracle.py -x -q -s 2>&1 | grep -v "^$" | tail -8 | cut -c1-600
timeout 600 python tools/accum_probe.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l[:300]); continue
    print(r['K'], r['terms'], ' '.join('%s=%.2e' % (k, r[k]['max_rel_to_max']) for k in ('bf16x3','bf16','tf32','bf16x3_8_partial_sums','torch_fp32_matmul')))
"
cp gpurun_out/r2_accum_error.json gpurun_out/r2_accum_error_chunked.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c44_bench.json 2> gpurun_out/r2c44_bench.err
CTCASR_GEMM_CHAIN=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c44_bench_nochain.json 2> gpurun_out/r2c44_bench_nochain.err
timeout 300 python bench.py --no-cpu-baseline --compute bf16 > gpurun_out/r2c44_bench_bf16.json 2> gpurun_out/r2c44_bench_bf16.err
timeout 300 python bench.py --no-cpu-baseline --model ds2 > gpurun_out/r2c44_bench_ds2.json 2> gpurun_out/r2c44_bench_ds2.err
python - <<'PY'
import json
for n in ("", "_nochain", "_bf16", "_ds2"):
    try:
        d = json.load(open("gpurun_out/r2c44_bench%s.json" % n))
        print(n or "chain", d["ms_per_step"], d.get("kernel_ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
PY
