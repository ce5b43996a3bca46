"""Reported bars next to the bench numbers (not optimisation targets; SURVEY.md 2.2 / BASELINE.md 3, 5):
  * torch.nn.LSTM (cuDNN 9 on B200) at the cfg2 recurrent shape: 2 layers, bidirectional, H = 2048, input 2048,
    T = 1000, B = 32, forward + backward, fp32 (TF32 off / on) and bf16;
  * TF32 and bf16 torch.matmul peaks measured like MEASURED_PEAKS.json (8192^3, best of 10);
  * torch.nn.LSTM on the host cores (oneDNN), forward + backward, on a T = 100 sample.
Writes one JSON object to stdout (commit it under profiles/)."""
import json
import os
import sys
import time

import torch


def cuda_ms(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def lstm_gpu(dtype, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    T, B, D, H = 1000, 32, 2048, 2048
    m = torch.nn.LSTM(D, H, num_layers=2, bidirectional=True).cuda().to(dtype)
    x = torch.randn(T, B, D, device="cuda", dtype=dtype, requires_grad=True)
    dy = torch.randn(T, B, 2 * H, device="cuda", dtype=dtype)

    def fwd():
        with torch.no_grad():
            m(x)

    def step():
        y, _ = m(x)
        y.backward(dy)
        m.zero_grad(set_to_none=True); x.grad = None

    return {"fwd_ms": cuda_ms(fwd), "fwd_bwd_ms": cuda_ms(step)}


def matmul_peak(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a, b = torch.randn(n, n, device="cuda", dtype=dtype), torch.randn(n, n, device="cuda", dtype=dtype)
    ms = cuda_ms(lambda: torch.matmul(a, b), n=10, warm=3)
    return 2.0 * n ** 3 / (ms * 1e-3) / 1e12


def lstm_cpu():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, B, D, H = 100, 32, 2048, 2048
    m = torch.nn.LSTM(D, H, num_layers=2, bidirectional=True)
    x = torch.randn(T, B, D, requires_grad=True)
    dy = torch.randn(T, B, 2 * H)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        y, _ = m(x)
        y.backward(dy)
        ts.append(time.perf_counter() - t0)
    return {"T": T, "B": B, "fwd_bwd_s": ts[-1], "frames_per_s_recurrent_layers_only": T * B / ts[-1], "cores": cores}


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    out["matmul_peak_tflops"] = {"tf32": matmul_peak(torch.float32, True), "fp32_no_tf32": matmul_peak(torch.float32, False),
                                 "bf16": matmul_peak(torch.bfloat16, False)}
    out["cudnn_lstm_cfg2_shape"] = {"shape": "2 layers, bidirectional, input 2048, H 2048, T 1000, B 32",
                                    "fp32": lstm_gpu(torch.float32, False), "tf32": lstm_gpu(torch.float32, True),
                                    "bf16": lstm_gpu(torch.bfloat16, False)}
    if "--no-cpu" not in sys.argv:
        out["onednn_cpu_lstm"] = lstm_cpu()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
