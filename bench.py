#!/usr/bin/env python
"""bench.py — audio-frames/s of one full training step of the hot path (BASELINE.json metric).

A "step" = inference_fn (dense stack + stacked BiLSTM) + CTC loss forward-backward + backward pass
+ (N>1: NCCL all-reduce of the flat gradient) + Adam, on one synthetic batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (sm_100a kernels)
  python bench.py --impl reference ...                         the reference path on host cores
  python bench.py --workload ctc ...                           cfg5: CTC forward-backward alone

At N=1 the workload is BASELINE.json configs[1] ("DS2": 3 dense + 2 BiLSTM-2048 + 2 dense, batch
32 x 10 s, 80-bin features, fp32 storage).  N>1: one process per GPU (torchrun), the same per-GPU
batch on every rank (weak scaling), one all-reduce of the gradient per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(used_model="ds1", num_layers_dense=3, num_units_dense=2048, num_layers_rnn=2, num_units_rnn=2048, rnn_cell="lstm",
            cudnn=False, dense_dropout_rate=0.1)
CFG2_B, CFG2_T, CFG2_L = 32, 1000, 160
CFG5 = dict(B=512, T=1700, L=84, V=29)
# CPU arm: frames per utterance in the bounded sample (of 1000).  32 frames x B=32 is ~3.5 s per step on 16-24
# host cores: 1 warm-up + 2 timed steps stay near 10 s, and a driver-chosen --steps 20 --warmup 5 under two minutes.
CPU_SAMPLE_FRAMES = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "ctc"])
    ap.add_argument("--compute", default="bf16x3", choices=["bf16x3", "tf32", "fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=CFG2_B)
    ap.add_argument("--frames", type=int, default=CFG2_T)
    ap.add_argument("--units", type=int, default=2048, help="debug: shrink D and H")
    ap.add_argument("--model", default="ds1", choices=["ds1", "ds2"],
                    help="front-end: ds1 = 3 dense layers (BASELINE configs[1], the default), ds2 = 3 conv layers")
    ap.add_argument("--cell", default="lstm", choices=["lstm", "gru", "rnn_relu", "rnn_tanh"],
                    help="rnn_cell of the reference's menu (asr/params.py:48); the metric is quoted on lstm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames per utterance in the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference_frames_per_s(cfg, B, sample_T, L, steps=2, warmup=1):
    """The reference's path restated on torch-CPU (oracle/torch_ref.py), timed on the host cores on
    a bounded sample: the same batch size and model, `sample_T` frames per utterance instead of the
    full length (every op on the path is linear in T)."""
    import torch
    from ctc_asr_b200 import synthetic
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synthetic.init_params(cfg, seed=1)
    p = torch_ref.params_to_torch(params, torch.float32)
    x, sl, lab, ll = synthetic.fixed_batch(B, sample_T, L, seed=0)
    xs, sls, labs, lls = (torch.from_numpy(a) for a in (x, sl, lab, ll))
    opt = torch.optim.Adam(list(p.values()), lr=cfg.learning_rate, betas=(cfg.adam_beta1, cfg.adam_beta2),
                           eps=cfg.adam_epsilon)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = torch_ref.train_step_grads(cfg, p, xs, sls, labs, lls)
        opt.step()
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times[warmup:]))
    return B * sample_T / dt, dt, cores


def run_reference(args):
    from ctc_asr_b200.params import ModelConfig
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = ModelConfig(**dict(CFG2, num_units_dense=args.units, num_units_rnn=args.units, dense_dropout_rate=0.0))
    # bounded sample: keep the whole --steps K --warmup W run near two minutes (~0.11 s per frame of a B=32 step on
    # 16-24 host cores), between 4 and CPU_SAMPLE_FRAMES frames per utterance
    auto_T = max(4, min(CPU_SAMPLE_FRAMES, int(120.0 / (0.11 * (args.steps + args.warmup)))))
    sample_T = args.cpu_sample_frames or auto_T
    L = max(1, min(CFG2_L, sample_T // 4))
    t0 = time.perf_counter()
    fps, dt, cores = cpu_reference_frames_per_s(cfg, args.batch, sample_T, L, steps=args.steps, warmup=args.warmup)
    sample = "B=%d utterances x %d frames per step (of %d), %d warm-up + %d timed steps, torch-CPU restatement of the TF graph" % (
        args.batch, sample_T, args.frames, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": "audio-frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "cfg2: 3 dense + 2 BiLSTM-%d + 2 dense, B=%d x T=%d frames x 80 features, L=%d, full train step "
                               "(TF 1.12 not installable offline: TF-equivalent restatement on torch-CPU)" % (
                                   args.units, args.batch, args.frames, CFG2_L)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ctc_asr_b200 import _lib, ops, synthetic
    from ctc_asr_b200.model import CTCModel
    from ctc_asr_b200.params import ModelConfig, flops_per_frame_fwd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    peaks = load_peaks()
    K, W = args.steps, max(args.warmup, 3)

    if args.workload == "ctc":
        return run_ctc(args, torch, lib, ops, synthetic, peaks, rank, world)

    cfg = ModelConfig(**dict(CFG2, num_units_dense=args.units, num_units_rnn=args.units, compute=args.compute,
                             used_model=args.model, rnn_cell=args.cell, cudnn=args.cell != "lstm"))
    B, T, L = args.batch, args.frames, min(CFG2_L, max(1, args.frames // 4))
    from ctc_asr_b200.params import conv_out_frames
    T_rnn = conv_out_frames(cfg, T)                      # ds2: the conv stack halves the frame rate
    model = CTCModel(cfg, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(B, T, L, seed=rank)
    hx, hsl = torch.from_numpy(x).pin_memory(), torch.from_numpy(sl).pin_memory()
    hlab, hll = torch.from_numpy(lab).pin_memory(), torch.from_numpy(ll).pin_memory()
    dx, dsl, dlab, dll = hx.cuda(), hsl.cuda(), hlab.cuda(), hll.cuda()
    gb = B * world
    allreduce = (lambda g: dist.all_reduce(g)) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop_body, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loop_body()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        model.train_step(dx, dsl, (dlab, dll), global_batch=gb, allreduce=allreduce)

    def step_e2e():
        a = hx.cuda(non_blocking=True); b = hsl.cuda(non_blocking=True)
        c = hlab.cuda(non_blocking=True); d = hll.cuda(non_blocking=True)
        loss = model.train_step(a, b, (c, d), global_batch=gb, allreduce=allreduce)
        return float(loss)          # D2H read of the step's result

    import ctypes
    for _ in range(W):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.ctcasr_launch_count()
    lib.ctcasr_profile_enable(1)            # CUDA events around the hot kernels, on the launching stream
    ms = timed(step_resident, K)
    prof_ms, prof_n = (ctypes.c_double * 5)(), (ctypes.c_int * 5)()
    lib.ctcasr_profile_collect(prof_ms, prof_n, 5)
    lib.ctcasr_profile_enable(0)
    launches = lib.ctcasr_launch_count() - l0
    clocks = sampler.finish()
    ms_e2e = timed(step_e2e, K)

    frames = B * T * world
    value = frames * K / (ms * 1e-3)
    e2e = frames * K / (ms_e2e * 1e-3)
    H, D = cfg.num_units_rnn, cfg.num_units_dense
    flops_step = 3.0 * flops_per_frame_fwd(cfg) * B * T            # per GPU
    # ---- roofline of the dominant kernel: the persistent LSTM recurrence (fwd + bwd launches) ----------
    # algorithmic bytes per time step and layer = the recurrent weights of both directions, which the
    # kernel has to stream once per step because they do not fit on chip (two bf16 pieces = 4 B per
    # weight) + the step's slice of P/gates (read + write) and c, y / dy
    w_bytes = 2 * H * 4 * H * 4
    act_bytes = B * 8 * H * 4 * 2 + B * 2 * H * 4 * 2
    lstm_launches = prof_n[0] + prof_n[1]
    lstm_ms = prof_ms[0] + prof_ms[1]
    lstm_bytes = (w_bytes + act_bytes) * T_rnn * lstm_launches
    lstm_gbs = lstm_bytes / (lstm_ms * 1e-3) / 1e9 if lstm_ms > 0 else 0.0
    # ---- second class: the tcgen05 GEMMs (everything GEMM-shaped but the recurrence and the 29-class layer)
    rec_flops = 2 * (cfg.num_layers_rnn * 2 * 2 * H * 4 * H) * T_rnn / T     # recurrent matvec fwd + bwd, per input frame
    tc_flops = (3.0 * flops_per_frame_fwd(cfg) - rec_flops - 3 * 2 * D * cfg.num_classes * T_rnn / T) * B * T * K
    gemm_tflops = tc_flops / (prof_ms[2] * 1e-3) / 1e12 if prof_ms[2] > 0 else 0.0
    mma_per_mac = {"bf16x3": 3.0, "tf32": 2.0, "fp32": 1.0, "bf16": 1.0}[args.compute]
    gemm_peak = peaks["bf16_tflops_sustained"] / mma_per_mac
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "lstm_dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    line = {
        "metric": "audio-frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.compute, "data": "synthetic",
        "config": {"workload": "cfg2: %s + 2 Bi%s-%d + 2 dense (3%s2r2d), per-GPU B=%d x T=%d frames x 80 features, "
                               "L=%d labels, fwd + CTC + bwd + %sAdam, dense dropout 0.1" % (
                                   "3 dense" if args.model == "ds1" else "3 conv (ds2 front-end, RNN at %d frames)" % T_rnn,
                                   {"lstm": "LSTM", "gru": "GRU", "rnn_relu": "RNN(relu)", "rnn_tanh": "RNN(tanh)"}[args.cell],
                                   args.units, "d" if args.model == "ds1" else "c", B, T, L,
                                   "NCCL all-reduce + " if world > 1 else ""),
                   "global_batch": gb, "params": model.num_params,
                   "arithmetic": {"bf16x3": "fp32 storage; GEMMs and recurrence as 3 (6 for ReLU-kinked layers) bf16 tcgen05 products "
                                            "of split operands, fp32 TMEM accumulation (fp32-level accuracy)",
                                  "tf32": "fp32 storage; tcgen05 kind::tf32", "fp32": "SIMT FFMA",
                                  "bf16": "BASELINE cfg3 arithmetic: GEMM operands rounded to bf16, one tcgen05 product, fp32 accumulation, "
                                          "fp32 master weights and CTC; LSTM recurrence bf16x3"}[args.compute],
                   "l2_policy": "inputs larger than L2 (>=7 GB of activations per step), no explicit flush"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(x.nbytes + sl.nbytes + lab.nbytes + ll.nbytes),
                "d2h_bytes_per_step": 4 + 4 * B},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "lstm_fwd_kernel + lstm_bwd_cluster_kernel (persistent recurrence, %d launches)" % lstm_launches,
                     "bound": "hbm", "achieved": lstm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": lstm_gbs / peaks["hbm_gbs"], "traffic": traffic,
                     "ms_per_launch": lstm_ms / max(lstm_launches, 1), "share_of_step": lstm_ms / ms,
                     "note": "algorithmic bytes = per time step the recurrent weights of both directions (streamed: 128 MiB "
                             "does not fit on chip) + gate/state slices; peak of %s" % peaks["source"]},
        "roofline_gemm": {"kernel": "gemm_tc_kernel (%d launches)" % prof_n[2], "bound": "tensor", "achieved": gemm_tflops,
                          "peak": gemm_peak, "unit": "TFLOP/s", "frac": gemm_tflops / gemm_peak if gemm_peak else None,
                          "share_of_step": prof_ms[2] / ms,
                          "note": "algorithmic GEMM FLOPs / event time; peak = %s bf16 sustained / %g MMAs per MAC" % (
                              peaks["source"], mma_per_mac)},
        "kernel_ms_per_step": {"lstm_fwd": prof_ms[0] / K, "lstm_bwd": prof_ms[1] / K, "gemm_tc": prof_ms[2] / K,
                               "ctc": prof_ms[3] / K, "operand_split": prof_ms[4] / K},
        "step_tflops": flops_step * K / (ms * 1e-3) / 1e12,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ccfg = cfg.replace(dense_dropout_rate=0.0)
        sT = args.cpu_sample_frames or CPU_SAMPLE_FRAMES
        fps, dt, cores = cpu_reference_frames_per_s(ccfg, B, sT, max(1, min(L, sT // 4)), steps=2, warmup=1)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "B=%d x %d frames per step (of %d), 1 warm-up + 2 timed steps, torch-CPU "
                                          "restatement of the TF graph (oracle/torch_ref.py)" % (B, sT, T)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ctc(args, torch, lib, ops, synthetic, peaks, rank, world):
    """cfg5: CTC forward-backward in isolation, judged on HBM GB/s."""
    B, T, L, V = CFG5["B"], CFG5["T"], CFG5["L"], CFG5["V"]
    rng = np.random.default_rng(rank)
    logits = torch.from_numpy((rng.standard_normal((T, B, V)) * 3).astype(np.float32))
    lab, ll = synthetic.make_labels(rng, B, L, T)
    hl = logits.pin_memory()
    dl, dlab, dll = hl.cuda(), torch.from_numpy(lab).cuda(), torch.from_numpy(ll).cuda()
    dsl = torch.full((B,), T, dtype=torch.int32, device="cuda")
    grad = torch.empty_like(dl)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        ops.ctc_loss(dl, dlab, dll, dsl, out_grad=grad)
    torch.cuda.synchronize()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    l0 = lib.ctcasr_launch_count()
    tot = 0.0
    for _ in range(K):
        flush.zero_()                                       # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.ctc_loss(dl, dlab, dll, dsl, out_grad=grad)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    launches = lib.ctcasr_launch_count() - l0
    clocks = sampler.finish()
    ms = tot / K
    t0 = time.perf_counter()
    for _ in range(K):
        d = hl.cuda(non_blocking=True)
        loss, g, st = ops.ctc_loss(d, dlab, dll, dsl, out_grad=grad)
        g.cpu()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / K
    alg_bytes = B * (2 * T * V * 4 + L * 4 + 12)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ctc_dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    line = {
        "metric": "audio-frames/s (CTC forward-backward only)", "value": B * T / (ms * 1e-3), "unit": "frames/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "cfg5: CTC alpha/beta, B=%d, T=%d, L=%d, V=%d" % (B, T, L, V), "l2_policy": "256 MiB flush between iterations"},
        "e2e": {"value": B * T / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(logits.numel() * 4),
                "d2h_bytes_per_step": int(logits.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": gbs / peaks["hbm_gbs"], "traffic": traffic, "kernel": "ctc_loss_kernel",
                     "note": "algorithmic bytes = logits in + gradient out + labels (SURVEY 8d); the kernel is bound by instruction "
                             "issue of the alpha/beta recursion (3 sweeps of 2L+1 states x T frames), not by HBM; peak of %s" % peaks["source"]},
    }
    if rank == 0:
        print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
