"""Time of one weight-gradient GEMM of cfg2 (dWx of the second BiLSTM layer: [32000, 4096]^T [32000, 16384]) with the
chained accumulation of gemm_tc.cu at the chunk length CTCASR_GEMM_CHAIN names (run once per setting, on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from ctc_asr_b200 import _lib, ops  # noqa: E402

K, M, N = 32000, 4096, 16384
a = torch.randn(K, M, device="cuda")
b = torch.randn(K, N, device="cuda")
c = torch.empty(M, N, device="cuda")
for mode in ("bf16x3",):
    cid = _lib.COMPUTE_ID[mode]
    for _ in range(2):
        ops.gemm(a, b, ta=True, out=c, compute=cid)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.gemm(a, b, ta=True, out=c, compute=cid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("CTCASR_GEMM_CHAIN=%s %s: %.3f ms per GEMM incl. operand split (%.0f TFLOP/s algorithmic)" % (
        os.environ.get("CTCASR_GEMM_CHAIN", "default"), mode, ms, 2.0 * K * M * N / ms / 1e9))
