"""Small single-purpose GPU workloads for `ncu --set full` captures (one GPU, a few launches).
  python tools/profile_target.py lstm [T]     one BiLSTM-2048 layer, B=32: recurrence fwd + bwd kernels
  python tools/profile_target.py gemm [mode]  the cfg2 GEMM shapes in the default arithmetic (or bf16 / tf32)
  python tools/profile_target.py ctc          cfg5 CTC forward-backward
  python tools/profile_target.py convgemm     the conv layers' GEMM shapes
  python tools/profile_target.py conv         second conv layer forward + backward (im2col / col2im kernels)
  python tools/profile_target.py beam         prefix beam search, 32 x 300 frames, width 1024
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ctc_asr_b200 import _lib, ops, synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "lstm"
C = _lib.COMPUTE_BF16X3
if what == "lstm":
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    B, nin, H = 32, 2048, 2048
    x = torch.randn(T, B, nin, device="cuda")
    wx = torch.randn(nin, 8 * H, device="cuda") * 0.02
    wh = torch.randn(2, H, 4 * H, device="cuda") * 0.02
    bias = torch.zeros(8 * H, device="cuda")
    sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
    rb, _ = ops.birnn_sizes(T, B, nin, H, 2)
    reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
    y = torch.empty(T, B, 2 * H, device="cuda")
    dy = torch.randn(T, B, 2 * H, device="cuda")
    dx = torch.empty(T, B, nin, device="cuda")
    dwx, dwh, db = torch.empty_like(wx), torch.empty_like(wh), torch.empty_like(bias)
    for _ in range(2):
        ops.birnn_fwd(x, sl, wx, wh, bias, y, reserve, 2, True, compute=C)
        ops.birnn_bwd(x, sl, wx, wh, y, reserve, dy, dx, dwx, dwh, db, 2, True, compute=C)
    torch.cuda.synchronize()
elif what == "convgemm":
    # the conv layers' GEMM shapes (second layer, B=32 x 10 s): dgrad (K = 64), forward (N = 64), wgrad (split-K)
    for (M, N, K, ta, tb) in [(320000, 7392, 64, 0, 1), (320000, 64, 7392, 0, 0), (7392, 64, 320000, 1, 0)]:
        a = torch.randn((K, M) if ta else (M, K), device="cuda")
        b = torch.randn((N, K) if tb else (K, N), device="cuda")
        c = torch.empty(M, N, device="cuda")
        for _ in range(2):
            ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=C)
        del a, b, c
    torch.cuda.synchronize()
elif what == "conv":
    # second conv layer of the ds2 front-end at B=32 x 10 s: [500,32,40,32 (pitch 64)] -> [500,32,20,32 (pitch 64)], 11x21 / (1,2)
    T, B, F, Cc, pitch, N, kt, kf, st, sf = 500, 32, 40, 32, 64, 64, 11, 21, 1, 2
    x = torch.randn(T * B * F, pitch, device="cuda")
    w = torch.randn(kt * kf * Cc, N, device="cuda") * 0.01
    bias = torch.zeros(N, device="cuda")
    y = torch.empty(T * B * 20, N, device="cuda")
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(bias)
    for _ in range(2):
        ops.conv2d_fwd(x, pitch, w, bias, y, T, B, F, Cc, kt, kf, st, sf, compute=C)
        dy = torch.randn_like(y)
        ops.conv2d_bwd(x, pitch, w, y, dy, dx, dw, db, T, B, F, Cc, kt, kf, st, sf, compute=C)
    torch.cuda.synchronize()
elif what == "beam":
    T, B, V = 300, 32, 29
    rng = np.random.default_rng(0)
    logits = torch.from_numpy((rng.standard_normal((T, B, V)) * 3).astype(np.float32)).cuda()
    sl = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ops.beam_search(logits, sl, beam_width=1024)
    torch.cuda.synchronize()
elif what == "gemm":
    if len(sys.argv) > 2:
        C = _lib.COMPUTE_ID[sys.argv[2]]            # e.g. bf16: the CTA-pair kernel with 64-wide k-blocks
    for (M, N, K, ta, tb) in [(32000, 16384, 4096, 0, 0), (4096, 16384, 32000, 1, 0), (32000, 4096, 16384, 0, 1)]:
        a = torch.randn((K, M) if ta else (M, K), device="cuda")
        b = torch.randn((N, K) if tb else (K, N), device="cuda")
        c = torch.empty(M, N, device="cuda")
        for _ in range(2):
            ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=C)
    torch.cuda.synchronize()
else:
    B, T, L, V = 512, 1700, 84, 29
    rng = np.random.default_rng(0)
    logits = torch.from_numpy((rng.standard_normal((T, B, V)) * 3).astype(np.float32)).cuda()
    lab, ll = synthetic.make_labels(rng, B, L, T)
    dsl = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ops.ctc_loss(logits, torch.from_numpy(lab).cuda(), torch.from_numpy(ll).cuda(), dsl)
    torch.cuda.synchronize()
