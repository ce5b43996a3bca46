"""GPU probe: kernel-only time (library CUDA events, tag 2 = tcgen05 GEMM) of the cfg2 GEMM shapes in one compute mode.
  PROBE_MODE=bf16|bf16x3 [CTCASR_GEMM_PAIR=0] python tools/gemm_shapes.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_asr_b200 import ops, _lib

mode = os.environ.get("PROBE_MODE", "bf16")
MODE = _lib.COMPUTE_ID[mode]
lib = _lib.load()
shapes = [(32000, 16384, 4096, 0, 0, "rnn L2 proj fwd"), (32000, 16384, 2048, 0, 0, "rnn L1 proj fwd"),
          (32000, 2048, 2048, 0, 0, "dense fwd"), (32000, 2048, 4096, 0, 0, "dense4 fwd"),
          (4096, 16384, 32000, 1, 0, "rnn L2 wgrad"), (32000, 4096, 16384, 0, 1, "rnn L2 dgrad"),
          (2048, 8192, 32000, 1, 0, "rnn wh wgrad")]
for (M, N, K, ta, tb, name) in shapes:
    a = torch.randn((K, M) if ta else (M, K), device="cuda")
    b = torch.randn((N, K) if tb else (K, N), device="cuda")
    c = torch.empty((M, N), device="cuda")
    for _ in range(2):
        ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=MODE)
    torch.cuda.synchronize()
    lib.ctcasr_profile_enable(1)
    for _ in range(5):
        ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=MODE)
    ms, n = (ctypes.c_double * 5)(), (ctypes.c_int * 5)()
    lib.ctcasr_profile_collect(ms, n, 5)
    lib.ctcasr_profile_enable(0)
    t = ms[2] / max(n[2], 1)
    print("%-6s pair=%s %-16s M=%5d N=%5d K=%5d ta=%d tb=%d: %7.3f ms  %7.1f TFLOP/s (algorithmic)" % (
        mode, os.environ.get("CTCASR_GEMM_PAIR", "1"), name, M, N, K, ta, tb, t, 2.0 * M * N * K / t / 1e9))
