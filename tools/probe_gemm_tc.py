"""GPU probe: tcgen05 GEMM (CTCASR_COMPUTE_TF32) against fp64 numpy for every operand orientation,
tile-tail shapes and epilogues.  Prints one line per case; never raises, so a single gpurun call
shows which descriptor/orientation combinations work."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ctc_asr_b200 import ops, _lib

TF32, FP32 = _lib.COMPUTE_TF32, _lib.COMPUTE_FP32
MODE = {'tf32': _lib.COMPUTE_TF32, 'bf16x3': _lib.COMPUTE_BF16X3}[os.environ.get('PROBE_MODE', 'bf16x3')]
print('mode', os.environ.get('PROBE_MODE', 'bf16x3'))


def case(M, N, K, ta, tb, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    try:
        c = ops.gemm(A, B, ta=ta, tb=tb, compute=MODE)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print("M=%d N=%d K=%d ta=%d tb=%d  EXCEPTION %s" % (M, N, K, ta, tb, e)); return
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    got = c.cpu().numpy().astype(np.float64)
    err = np.abs(got - want).max() / np.abs(want).max()
    # signed bias: mean of (got - want) * sign(want) relative to mean |want|
    bias = ((got - want) * np.sign(want)).mean() / np.abs(want).mean()
    ref32 = torch.matmul((A.T if ta else A), (B.T if tb else B)).cpu().numpy()
    err32 = np.abs(ref32 - want).max() / np.abs(want).max()
    print("M=%5d N=%5d K=%5d ta=%d tb=%d  max rel err %.3e  signed bias %.3e  (torch fp32 err %.1e)" % (M, N, K, ta, tb, err, bias, err32))


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    for ta in (0, 1):
        for tb in (0, 1):
            case(128, 256, 64, ta, tb)
    for ta in (0, 1):
        for tb in (0, 1):
            case(384, 512, 256, ta, tb)
            case(200, 320, 80, ta, tb)        # tails in M, N and K
    case(4096, 2048, 2048, 0, 0)
    case(2048, 4096, 4096, 1, 0)
    case(4096, 2048, 4096, 0, 1)
    # timing of the cfg2 shapes
    for (M, N, K, ta, tb, name) in [(32000, 16384, 4096, 0, 0, "rnn L2 input fwd"), (32000, 2048, 2048, 0, 0, "dense fwd"),
                                    (4096, 16384, 32000, 1, 0, "rnn L2 wgrad"), (32000, 4096, 16384, 0, 1, "rnn L2 dgrad")]:
        a = torch.randn((K, M) if ta else (M, K), device="cuda")
        b = torch.randn((N, K) if tb else (K, N), device="cuda")
        c = torch.empty((M, N), device="cuda")
        try:
            for _ in range(2):
                ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=MODE)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gemm(a, b, ta=bool(ta), tb=bool(tb), out=c, compute=MODE)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            torch.backends.cuda.matmul.allow_tf32 = True
            aa, bb = (a.T if ta else a), (b.T if tb else b)
            for _ in range(2):
                torch.matmul(aa, bb, out=c)
            e0.record()
            for _ in range(5):
                torch.matmul(aa, bb, out=c)
            e1.record(); torch.cuda.synchronize()
            ms_cublas = e0.elapsed_time(e1) / 5
            torch.backends.cuda.matmul.allow_tf32 = False
            print("%-18s M=%d N=%d K=%d: %.3f ms = %.1f TFLOP/s   (cuBLAS tf32 %.3f ms = %.1f TFLOP/s)" % (
                name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9, ms_cublas, 2.0 * M * N * K / ms_cublas / 1e9))
        except Exception as e:  # noqa
            print(name, "EXCEPTION", e)
