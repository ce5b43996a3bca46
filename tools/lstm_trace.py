"""Per-phase timeline of the persistent forward recurrence kernels (globaltimer stamps of a few CTAs).
  python tools/lstm_trace.py [cell=lstm|gru|rnn_relu] [compute=bf16x3|bf16]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ctc_asr_b200 import _lib, ops
lib = _lib.load()
cell = sys.argv[1] if len(sys.argv) > 1 else "lstm"
compute = _lib.COMPUTE_ID[sys.argv[2] if len(sys.argv) > 2 else "bf16x3"]
cid = {"rnn_tanh": 0, "rnn_relu": 1, "lstm": 2, "gru": 3}[cell]
G = {0: 1, 1: 1, 2: 4, 3: 3}[cid]
T, B, nin, H = 64, 32, 2048, 2048
grid = 128
trace = torch.zeros(grid * 64 * 8, dtype=torch.int64, device="cuda")
lib.ctcasr_debug_lstm_trace.argtypes = [ctypes.c_void_p]
lib.ctcasr_debug_lstm_trace(ctypes.c_void_p(trace.data_ptr()))
x = torch.randn(T, B, nin, device="cuda"); wx = torch.randn(nin, 2 * G * H, device="cuda") * 0.02
wh = torch.randn(2, H, G * H, device="cuda") * 0.02; bias = torch.zeros(2 * G * H + (2 * H if cid == 3 else 0), device="cuda")
sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
rb, _ = ops.birnn_sizes(T, B, nin, H, cid); reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
y = torch.empty(T, B, 2 * H, device="cuda")
for _ in range(2):
    trace.zero_()
    ops.birnn_fwd(x, sl, wx, wh, bias, y, reserve, cid, True, compute=compute)
torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(grid, 64, 8).astype(np.float64)
if cid >= 2:
    names = ["B: barrier passed", "B: all h tiles issued", "MMA: all issued+commit", "EPI: accumulator ready", "EPI: cells+stores done", "EPI: fenced+signalled", "A: last weight tile issued"]
else:
    names = ["B: barrier passed", "B: all h tiles issued", "MMA: all issued+commit", "EPI: accumulator ready", "EPI: partials landed", "EPI: cells+stores issued", "EPI: fenced+signalled"]
print("cell", cell, "compute", sys.argv[2] if len(sys.argv) > 2 else "bf16x3")
for cta in (0, 37, 64, 127):
    print("CTA", cta)
    for step in (20, 21, 22):
        t0 = tr[cta, step, 0]
        print("  step %d:" % step, "  ".join("%s %+.2fus" % (names[k].split(":")[0] + str(k), (tr[cta, step, k] - t0) / 1e3) for k in range(7)),
              " | step period %.2fus" % ((tr[cta, step + 1, 0] - t0) / 1e3))
print("slots:", {k: n for k, n in enumerate(names)})
sig = 5 if cid >= 2 else 6
b = tr[:64, 21, 0]; print("barrier-pass skew across dir-0 CTAs at step 21: %.2f us" % ((b.max() - b.min()) / 1e3))
s5 = tr[:64, 20, sig]; print("signal time spread at step 20: %.2f us; last signal -> next barrier pass (CTA0): %.2f us" % ((s5.max() - s5.min()) / 1e3, (tr[0, 21, 0] - s5.max()) / 1e3))
