"""Times the stepwise recurrent path (one Bi-RNN layer, B=32, H=2048) per frame: python tools/step_probe.py [cell] [T]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_asr_b200 import _lib, ops
cell = sys.argv[1] if len(sys.argv) > 1 else "rnn_relu"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cid = {"rnn_tanh": 0, "rnn_relu": 1, "lstm": 2, "gru": 3}[cell]
G = {0: 1, 1: 1, 2: 4, 3: 3}[cid]
B, nin, H = 32, 2048, 2048
x = torch.randn(T, B, nin, device="cuda")
wx = torch.randn(nin, 2 * G * H, device="cuda") * 0.01
wh = torch.randn(2, H, G * H, device="cuda") * 0.01
bias = torch.zeros(2 * G * H + (2 * H if cid == 3 else 0), device="cuda")
sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
rb, _ = ops.birnn_sizes(T, B, nin, H, cid)
reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
y = torch.empty(T, B, 2 * H, device="cuda"); dy = torch.randn(T, B, 2 * H, device="cuda") * 0.01
dx = torch.empty(T, B, nin, device="cuda")
dwx, dwh, db = torch.empty_like(wx), torch.empty_like(wh), torch.empty_like(bias)
C = _lib.COMPUTE_BF16X3
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ops.birnn_fwd(x, sl, wx, wh, bias, y, reserve, cid, False, compute=C)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ops.birnn_bwd(x, sl, wx, wh, y, reserve, dy, dx, dwx, dwh, db, cid, False, compute=C)
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print("%s T=%d: fwd host-enqueue %.1f us/frame, fwd total %.1f us/frame | bwd host-enqueue %.1f, bwd total %.1f us/frame"
          % (cell, T, (t1 - t0) / T * 1e6, (t2 - t0) / T * 1e6, (t3 - t2) / T * 1e6, (t4 - t2) / T * 1e6))
