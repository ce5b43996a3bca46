"""Per-phase timeline of the persistent LSTM forward kernel (globaltimer stamps of a few CTAs)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ctc_asr_b200 import _lib, ops
lib = _lib.load()
T, B, nin, H = 64, 32, 2048, 2048
grid = 2 * H // 32
trace = torch.zeros(grid * 64 * 8, dtype=torch.int64, device="cuda")
lib.ctcasr_debug_lstm_trace.argtypes = [ctypes.c_void_p]
lib.ctcasr_debug_lstm_trace(ctypes.c_void_p(trace.data_ptr()))
x = torch.randn(T, B, nin, device="cuda"); wx = torch.randn(nin, 8 * H, device="cuda") * 0.02
wh = torch.randn(2, H, 4 * H, device="cuda") * 0.02; bias = torch.zeros(8 * H, device="cuda")
sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
rb, _ = ops.birnn_sizes(T, B, nin, H, 2); reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
y = torch.empty(T, B, 2 * H, device="cuda")
for _ in range(2):
    trace.zero_()
    ops.birnn_fwd(x, sl, wx, wh, bias, y, reserve, 2, True, compute=_lib.COMPUTE_BF16X3)
torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(grid, 64, 8).astype(np.float64)
names = ["B: barrier passed", "B: all h tiles issued", "MMA: all issued+commit", "EPI: accumulator ready", "EPI: cells+stores done", "EPI: fenced+signalled", "A: last weight tile issued"]
for cta in (0, 37, 64, 127):
    print("CTA", cta)
    for step in (20, 21, 22):
        t0 = tr[cta, step, 0]
        print("  step %d:" % step, "  ".join("%s %+.2fus" % (names[k].split(":")[0] + str(k), (tr[cta, step, k] - t0) / 1e3) for k in range(7)),
              " | step period %.2fus" % ((tr[cta, step + 1, 0] - t0) / 1e3))
print("slots:", {k: n for k, n in enumerate(names)})
# skew of barrier passing across CTAs of direction 0 at step 21
b = tr[:64, 21, 0]; print("barrier-pass skew across dir-0 CTAs at step 21: %.2f us" % ((b.max() - b.min()) / 1e3))
s5 = tr[:64, 20, 5]; print("signal time spread at step 20: %.2f us; last signal -> next barrier pass (CTA0): %.2f us" % ((s5.max() - s5.min()) / 1e3, (tr[0, 21, 0] - s5.max()) / 1e3))
