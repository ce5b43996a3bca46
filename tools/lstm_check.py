"""Determinism + timing check of the persistent LSTM kernels at cfg2 size (one BiLSTM-2048 layer, B=32).
  python tools/lstm_check.py [T] [repeats]
Runs the layer forward + backward `repeats` times on identical inputs; every run must reproduce the
first one bit for bit (a data race in the cross-CTA state exchange shows up here long before it
shows up as a hang).  Prints the kernel times from the library's CUDA-event profiler.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_asr_b200 import _lib, ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, nin, H = 32, 2048, 2048
C = _lib.COMPUTE_BF16X3
torch.manual_seed(0)
x = torch.randn(T, B, nin, device="cuda")
wx = torch.randn(nin, 8 * H, device="cuda") * 0.02
wh = torch.randn(2, H, 4 * H, device="cuda") * 0.02
bias = torch.zeros(8 * H, device="cuda")
sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
rb, _ = ops.birnn_sizes(T, B, nin, H, 2)
reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
y = torch.empty(T, B, 2 * H, device="cuda")
dy = torch.randn(T, B, 2 * H, device="cuda") * 0.01
dx = torch.empty(T, B, nin, device="cuda")
dwx, dwh, db = torch.empty_like(wx), torch.empty_like(wh), torch.empty_like(bias)
lib = _lib.load()
first = None
bad = 0
for r in range(R):
    if r == 1:
        lib.ctcasr_profile_enable(1)
    ops.birnn_fwd(x, sl, wx, wh, bias, y, reserve, 2, True, compute=C)
    ops.birnn_bwd(x, sl, wx, wh, y, reserve, dy, dx, dwx, dwh, db, 2, True, compute=C)
    torch.cuda.synchronize()
    got = [t.clone() for t in (y, dx, dwh)]
    if first is None:
        first = got
    elif not all(torch.equal(a, b) for a, b in zip(first, got)):
        bad += 1
        print("run %d differs: max |dy| %.3e |ddx| %.3e" % (r, (first[0] - got[0]).abs().max().item(), (first[1] - got[1]).abs().max().item()))
ms = (ctypes.c_double * 5)()
n = (ctypes.c_int * 5)()
lib.ctcasr_profile_collect(ms, n, 5)
lib.ctcasr_profile_enable(0)
print("T=%d repeats=%d nondeterministic=%d | lstm_fwd %.3f ms/launch (%.2f us/step)  lstm_bwd %.3f ms/launch (%.2f us/step)" % (
    T, R, bad, ms[0] / max(n[0], 1), 1e3 * ms[0] / max(n[0], 1) / T, ms[1] / max(n[1], 1), 1e3 * ms[1] / max(n[1], 1) / T))
sys.exit(1 if bad else 0)
