#!/usr/bin/env python
"""bench.py — audio-frames/s of one full training step of the hot path (BASELINE.json metric).

A "step" = inference_fn (dense stack + stacked BiRNN) + CTC loss forward-backward + backward pass
+ (N>1: NCCL all-reduce of the gradient, one bucket per layer group, under the backward pass) + Adam, on one synthetic batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (sm_100a kernels), BASELINE configs[1]
  python bench.py --impl reference ...                         the reference path on the host cores, same config
  python bench.py --compute bf16                               configs[2]'s arithmetic (bf16 operands)
  python bench.py --workload varlen                            configs[3]: bucketed variable-length batches of 64
  python bench.py --workload ctc [--sweep]                     configs[4]: CTC forward-backward alone (+ B x T sweep)
  python bench.py --check  (under torchrun, N > 1)             data-parallel correctness on the GPUs

At N=1 the workload is BASELINE.json configs[1] ("DS2": 3 dense + 2 BiLSTM-2048 + 2 dense, batch
32 x 10 s, 80-bin features, fp32 storage).  N>1: one process per GPU (torchrun), the same per-GPU
batch on every rank (weak scaling).  Prints ONE JSON line (rank 0).
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
# CPU arm: frames per utterance in the bounded sample (of 1000), BASELINE.md section 5: "time T=250 and state the
# reduction".  B=32 x 250 frames is ~22 s per step on 16 host cores.
CPU_SAMPLE_FRAMES = 250
CELL_NAMES = {"lstm": "LSTM", "gru": "GRU", "rnn_relu": "RNN(relu)", "rnn_tanh": "RNN(tanh)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "ctc", "varlen"])
    ap.add_argument("--compute", default="bf16x3", choices=["bf16x3", "tf32", "fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=CFG2_B)
    ap.add_argument("--frames", type=int, default=CFG2_T)
    ap.add_argument("--units", type=int, default=2048, help="debug: shrink D and H")
    ap.add_argument("--model", default="ds1", choices=["ds1", "ds2"],
                    help="front-end: ds1 = 3 dense layers (BASELINE configs[1], the default), ds2 = 3 conv layers")
    ap.add_argument("--cell", default="lstm", choices=["lstm", "gru", "rnn_relu", "rnn_tanh"],
                    help="rnn_cell of the reference's menu (asr/params.py:48); the metric is quoted on lstm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames per utterance in the CPU sample (0 = %d)" % CPU_SAMPLE_FRAMES)
    ap.add_argument("--overlap", action="store_true", help="N>1: one asynchronous all-reduce per gradient bucket under the backward pass "
                    "instead of one all-reduce after it (measured: no gain, the collective's CTAs displace the persistent recurrence kernels)")
    ap.add_argument("--no-overlap", action="store_true", help="(default) one all-reduce of the whole gradient after backward")
    ap.add_argument("--sweep", action="store_true", help="--workload ctc: B x T sweep of SURVEY 8(d)")
    ap.add_argument("--check", action="store_true", help="N>1: all-reduced gradient == 1-GPU gradient of the global batch; "
                                                         "parameters bit-identical across ranks after 5 steps")
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
        peaks = {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                 "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    else:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    # L2 -> SM read bandwidth of an L2-resident working set, measured with tools/ubench/l2_peak.cu on this pool
    l2 = _profile_json("r2_l2_peak.json")
    peaks["l2_gbs"] = max(l2["l2_resident_gbs"].values()) if l2 else None
    return peaks


def _profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    return json.load(open(path)) if os.path.exists(path) else None


# ------------------------------------------------------------------------------ configuration (both arms)
def make_cfg(args, **over):
    from ctc_asr_b200.params import ModelConfig
    kw = dict(CFG2, num_units_dense=args.units, num_units_rnn=args.units, compute=args.compute, used_model=args.model,
              rnn_cell=args.cell, cudnn=args.cell != "lstm")
    kw.update(over)
    return ModelConfig(**kw)


def workload_config(args, cfg, world):
    """The `config` object of the JSON line: identical for our arm and for --impl reference."""
    from ctc_asr_b200.params import conv_out_frames, param_specs
    B, T = args.batch, args.frames
    L = min(CFG2_L, max(1, T // 4))
    T_rnn = conv_out_frames(cfg, T)
    front = "3 dense" if args.model == "ds1" else "3 conv (ds2 front-end, RNN at %d frames)" % T_rnn
    return {"workload": "cfg2: %s + 2 Bi%s-%d + 2 dense (3%s2r2d), per-GPU B=%d x T=%d frames x 80 features, L=%d labels, "
                        "fwd + CTC + bwd + Adam (N>1: + NCCL all-reduce), dense dropout 0.1" % (
                            front, CELL_NAMES[args.cell], args.units, "d" if args.model == "ds1" else "c", B, T, L),
            "global_batch": B * world,
            "params": int(sum(int(np.prod(s)) for _, s, _ in param_specs(cfg))),
            "arithmetic": {"bf16x3": "fp32 storage; GEMMs and recurrence as 3 (6 for ReLU-kinked layers) bf16 tcgen05 products "
                                     "of split operands, fp32 TMEM accumulation in chains of <= 8192 (1e-5 .. 3e-5 of fp64)",
                           "tf32": "fp32 storage; tcgen05 kind::tf32 GEMMs, bf16x3 recurrence", "fp32": "SIMT FFMA",
                           "bf16": "BASELINE cfg3 arithmetic: GEMM and recurrence operands rounded to bf16, one tcgen05 product, fp32 "
                                   "accumulation, fp32 master weights, state and CTC"}[args.compute],
            "l2_policy": "inputs larger than L2 (>=7 GB of activations per step), no explicit flush",
            "allreduce": None if world == 1 else ("one asynchronous NCCL all-reduce per gradient bucket (dense4+logits, each RNN layer, "
                                                  "front-end) under the backward pass" if args.overlap else "one NCCL all-reduce of the flat gradient after backward")}


# ------------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference_frames_per_s(cfg, B, sample_T, L, steps=1, warmup=0, prime=True):
    """The reference's path restated on torch-CPU (oracle/torch_ref.py), timed on the host cores on a bounded sample:
    the same batch size and model, `sample_T` frames per utterance instead of the full length (every op on the path is
    linear in T; the fixed per-step costs — Adam over all parameters, Python dispatch — are included)."""
    import torch
    from ctc_asr_b200 import synthetic
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synthetic.init_params(cfg, seed=1)
    p = torch_ref.params_to_torch(params, torch.float32)
    opt = torch.optim.Adam(list(p.values()), lr=cfg.learning_rate, betas=(cfg.adam_beta1, cfg.adam_beta2),
                           eps=cfg.adam_epsilon)

    def batch(T):
        x, sl, lab, ll = synthetic.fixed_batch(B, T, max(1, min(L, T // 4)), seed=0)
        return tuple(torch.from_numpy(a) for a in (x, sl, lab, ll))

    if prime:                                           # thread pool, allocator, optimizer state: a 4-frame step, untimed
        torch_ref.train_step_grads(cfg, p, *batch(4))
        opt.step()
    xs, sls, labs, lls = batch(sample_T)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_ref.train_step_grads(cfg, p, xs, sls, labs, lls)
        opt.step()
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times[warmup:]))
    return B * sample_T / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = make_cfg(args)
    ccfg = cfg.replace(dense_dropout_rate=0.0)
    t0 = time.perf_counter()
    # bounded sample: B utterances x sample_T frames per step.  BASELINE.md section 5 asks for T = 250; the whole
    # --steps K --warmup W run has to end within a few minutes, so sample_T is what a 5-minute budget allows at the speed
    # a 16-frame calibration step shows, between 64 and 250 frames (250 whenever K + W <= ~11 steps on 16 cores)
    sample_T = args.cpu_sample_frames
    if not sample_T:
        _, dt16, _ = cpu_reference_frames_per_s(ccfg, args.batch, 16, CFG2_L, steps=1, warmup=0)
        per_frame = dt16 / 16.0
        sample_T = int(max(64, min(CPU_SAMPLE_FRAMES, 300.0 / ((args.steps + args.warmup) * per_frame))))
    fps, dt, cores = cpu_reference_frames_per_s(ccfg, args.batch, sample_T, CFG2_L, steps=args.steps, warmup=args.warmup)
    sample = ("each step = B=%d utterances x %d frames (of the workload's %d; per-frame cost is independent of T, the per-step fixed costs "
              "- Adam over all parameters, Python dispatch - are included), %d warm-up + %d timed steps, torch-CPU restatement of the "
              "reference's TF graph (oracle/torch_ref.py; TF 1.12 is not installable offline), all %d host threads, no dropout" % (
                  args.batch, sample_T, args.frames, args.warmup, args.steps, cores))
    line = {
        "impl": "reference", "metric": "audio-frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, cfg, world),
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the persistent recurrence kernels occupy 128 of the 148 SMs: keep the collective's CTAs inside the other 20 so
        # that the bucketed all-reduce really runs under them
        if "--overlap" in sys.argv:
            os.environ.setdefault("NCCL_MAX_CTAS", "4")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    peaks = load_peaks()
    ctx = dict(args=args, torch=torch, dist=dist, lib=lib, ops=ops, synthetic=synthetic, peaks=peaks, rank=rank, world=world,
               local_rank=local_rank)
    if args.check:
        run_check(ctx)
    elif args.workload == "ctc":
        run_ctc(ctx)
    elif args.workload == "varlen":
        run_varlen(ctx)
    else:
        run_train(ctx)
    if world > 1:
        dist.destroy_process_group()


def _timers(ctx):
    torch, dist, world = ctx["torch"], ctx["dist"], ctx["world"]

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
    return barrier, timed


def _allreduce(ctx):
    dist = ctx["dist"]
    if ctx["world"] == 1:
        return None
    return lambda g, async_op=False: dist.all_reduce(g, async_op=async_op)


def run_train(ctx):
    import ctypes
    args, torch, lib, synthetic, peaks = ctx["args"], ctx["torch"], ctx["lib"], ctx["synthetic"], ctx["peaks"]
    rank, world = ctx["rank"], ctx["world"]
    from ctc_asr_b200.model import CTCModel
    from ctc_asr_b200.params import CELL_ID, NUM_GATES, conv_out_frames, flops_per_frame_fwd
    K, W = args.steps, max(args.warmup, 3)
    cfg = make_cfg(args)
    B, T, L = args.batch, args.frames, min(CFG2_L, max(1, args.frames // 4))
    T_rnn = conv_out_frames(cfg, T)                      # ds2: the conv stack halves the frame rate
    model = CTCModel(cfg, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(B, T, L, seed=rank)
    hx, hsl = torch.from_numpy(x).pin_memory(), torch.from_numpy(sl).pin_memory()
    hlab, hll = torch.from_numpy(lab).pin_memory(), torch.from_numpy(ll).pin_memory()
    dx, dsl, dlab, dll = hx.cuda(), hsl.cuda(), hlab.cuda(), hll.cuda()
    gb = B * world
    allreduce, overlap = _allreduce(ctx), args.overlap
    barrier, timed = _timers(ctx)

    def step_resident():
        model.train_step(dx, dsl, (dlab, dll), global_batch=gb, allreduce=allreduce, overlap=overlap)

    def step_e2e():
        a = hx.cuda(non_blocking=True); b = hsl.cuda(non_blocking=True)
        c = hlab.cuda(non_blocking=True); d = hll.cuda(non_blocking=True)
        loss = model.train_step(a, b, (c, d), global_batch=gb, allreduce=allreduce, overlap=overlap)
        return float(loss)          # D2H read of the step's result

    for _ in range(W):
        step_resident()
    model.check_step()
    sampler = ClockSampler(ctx["local_rank"])
    sampler.start()
    l0 = lib.ctcasr_launch_count()
    lib.ctcasr_profile_enable(1)            # CUDA events around the hot kernels, on the launching stream
    ms = timed(step_resident, K)
    prof_ms, prof_n = (ctypes.c_double * 5)(), (ctypes.c_int * 5)()
    lib.ctcasr_profile_collect(prof_ms, prof_n, 5)
    lib.ctcasr_profile_enable(0)
    launches = lib.ctcasr_launch_count() - l0
    clocks = sampler.finish()
    model.check_step()
    ms_e2e = timed(step_e2e, K)

    frames = B * T * world
    value = frames * K / (ms * 1e-3)
    e2e = frames * K / (ms_e2e * 1e-3)
    H, D = cfg.num_units_rnn, cfg.num_units_dense
    G = NUM_GATES[cfg.rnn_cell]
    flops_step = 3.0 * flops_per_frame_fwd(cfg) * B * T            # per GPU
    # ---- roofline of the dominant kernel class: the persistent recurrence (one forward + one backward launch per layer) ----
    # SURVEY 8(d) puts the RNN on the tensor roofline: algorithmic FLOPs of one launch = both directions' recurrent
    # products  2 dirs x 2 B H GH  per time step (the same in the backward pass: dh = dz Wh^T), over T steps.
    rec_launches = prof_n[0] + prof_n[1]
    rec_ms = prof_ms[0] + prof_ms[1]
    rec_flops_launch = 2.0 * 2.0 * B * H * G * H * T_rnn
    rec_tflops = rec_flops_launch * rec_launches / (rec_ms * 1e-3) / 1e12 if rec_ms > 0 else 0.0
    mma_per_mac = {"bf16x3": 3.0, "tf32": 3.0, "fp32": 0.0, "bf16": 1.0}[args.compute]
    cid, comp = CELL_ID[cfg.rnn_cell], _lib_compute(args.compute)
    stream_bytes = 0.5 * (lib.ctcasr_birnn_stream_bytes(T_rnn, B, H, cid, comp, 0) + lib.ctcasr_birnn_stream_bytes(T_rnn, B, H, cid, comp, 1))
    ms_launch = rec_ms / max(rec_launches, 1)
    traffic = None
    tr = _profile_json("r2_rec_dram_traffic.json")
    if tr and tr.get("cell") == args.cell and tr.get("compute") == args.compute and args.units == 2048:
        traffic = tr.get("dram_bytes_per_launch")
    roofline = {
        "kernel": "persistent recurrence: gated_fwd_kernel + gated_bwd_cluster_kernel (lstm / gru) or rnn_rec_kernel (rnn_relu / rnn_tanh), "
                  "%d launches of %d time steps" % (rec_launches, T_rnn),
        "bound": "tensor", "achieved": rec_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": rec_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic,
        "ms_per_launch": ms_launch, "us_per_time_step": 1e3 * ms_launch / T_rnn, "share_of_step": rec_ms / ms,
        "mma_per_mac": mma_per_mac, "frac_of_issued_mma": rec_tflops * mma_per_mac / peaks["bf16_tflops_sustained"],
        "l2_stream": {"bytes_per_launch": stream_bytes, "achieved_gbs": stream_bytes / (ms_launch * 1e-3) / 1e9 if ms_launch else None,
                      "peak_gbs": peaks["l2_gbs"],
                      "frac": (stream_bytes / (ms_launch * 1e-3) / 1e9 / peaks["l2_gbs"]) if (ms_launch and peaks["l2_gbs"]) else None,
                      "note": "bytes TMA pulls from L2 / HBM per launch (weight tiles not resident in tensor memory + the state tiles of "
                              "every step, all CTAs) against the L2 read peak measured with tools/ubench/l2_peak.cu"},
        "dram": {"achieved_gbs": traffic / (ms_launch * 1e-3) / 1e9 if (traffic and ms_launch) else None, "peak_gbs": peaks["hbm_gbs"],
                 "frac": traffic / (ms_launch * 1e-3) / 1e9 / peaks["hbm_gbs"] if (traffic and ms_launch) else None,
                 "note": "ncu dram__bytes_read + write per launch from the round-2 capture of these kernels, unchanged since (profiles/r2_rec_dram_traffic.json)"},
        "note": "achieved = algorithmic recurrent FLOPs (2 directions x 2 B H GH per time step) / CUDA-event time; peak = %s bf16 sustained. "
                "At B = 32 the recurrence is a T-step chain of skinny products: the tensor fraction is low by construction, the step "
                "period (us_per_time_step) against its latency chain is what DESIGN.md section 4 analyses" % peaks["source"]}
    # ---- second class: the tcgen05 GEMMs (everything GEMM-shaped but the recurrence and the 29-class layer)
    rec_flops = 2 * (cfg.num_layers_rnn * 2 * 2 * H * G * H) * T_rnn / T     # recurrent matvec fwd + bwd, per input frame
    tc_flops = (3.0 * flops_per_frame_fwd(cfg) - rec_flops - 3 * 2 * D * cfg.num_classes * T_rnn / T) * B * T * K
    gemm_tflops = tc_flops / (prof_ms[2] * 1e-3) / 1e12 if prof_ms[2] > 0 else 0.0
    gemm_mma = {"bf16x3": 3.0, "tf32": 2.0, "fp32": 1.0, "bf16": 1.0}[args.compute]
    gemm_peak = peaks["bf16_tflops_sustained"] / gemm_mma
    config = workload_config(args, cfg, world)
    line = {
        "metric": "audio-frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.compute, "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(x.nbytes + sl.nbytes + lab.nbytes + ll.nbytes),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_gemm": {"kernel": "gemm_tc_kernel + gemm_tc_pair_kernel (%d launches)" % prof_n[2], "bound": "tensor", "achieved": gemm_tflops,
                          "peak": gemm_peak, "unit": "TFLOP/s", "frac": gemm_tflops / gemm_peak if gemm_peak else None,
                          "share_of_step": prof_ms[2] / ms,
                          "frac_of_bf16_sustained": gemm_tflops / peaks["bf16_tflops_sustained"],
                          "note": "algorithmic GEMM FLOPs / event time; peak = %s bf16 sustained / %g MMAs per MAC" % (
                              peaks["source"], gemm_mma)},
        "kernel_ms_per_step": {"rec_fwd": prof_ms[0] / K, "rec_bwd": prof_ms[1] / K, "gemm_tc": prof_ms[2] / K,
                               "ctc": prof_ms[3] / K, "operand_split": prof_ms[4] / K},
        "step_tflops": flops_step * K / (ms * 1e-3) / 1e12,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ccfg = cfg.replace(dense_dropout_rate=0.0)
        sT = args.cpu_sample_frames or CPU_SAMPLE_FRAMES
        fps, dt, cores = cpu_reference_frames_per_s(ccfg, B, sT, L, steps=1, warmup=0)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "one step of B=%d x %d frames (of %d) after a 4-frame priming step, torch-CPU "
                                          "restatement of the TF graph (oracle/torch_ref.py), %.1f s" % (B, sT, T, dt)}
    if rank == 0:
        print(json.dumps(line))


def _lib_compute(name):
    from ctc_asr_b200 import _lib
    return _lib.COMPUTE_ID[name]


# ------------------------------------------------------------------------------------ cfg4: variable length
def run_varlen(ctx):
    """BASELINE configs[3]: the 6,144-utterance synthetic corpus (0.7-17 s), bucketed into batches of 64 like
    asr/input_functions.py:90-98, every batch one training step; true frames / s (padding does not count).
    N>1: every rank takes an interleaved shard of each batch (same count and length mix per GPU)."""
    args, torch, lib, synthetic = ctx["args"], ctx["torch"], ctx["lib"], ctx["synthetic"]
    rank, world = ctx["rank"], ctx["world"]
    from ctc_asr_b200 import parallel
    from ctc_asr_b200.model import CTCModel
    cfg = make_cfg(args)
    model = CTCModel(cfg, seed=1)
    Bg = 64 * world if world > 1 else 64
    batches = list(synthetic.variable_batches(n_utts=6144, batch_size=Bg, seed=4))
    K = min(len(batches), args.steps if args.steps != 8 else len(batches))
    pick = np.linspace(0, len(batches) - 1, K).round().astype(int)           # spread over the buckets when K < all
    batches = [batches[i] for i in pick]
    dev = []
    for x, sl, lab, ll in batches:
        idx = parallel.shard_indices(Bg, rank, world, interleave=True)
        dev.append(tuple(torch.from_numpy(np.ascontiguousarray(a[idx])).pin_memory() for a in (x, sl, lab, ll)))
    allreduce = _allreduce(ctx)
    barrier, timed = _timers(ctx)
    res = [tuple(a.cuda() for a in b) for b in dev]

    def epoch(batches_, from_host):
        for b in batches_:
            x, sl, lab, ll = (a.cuda(non_blocking=True) for a in b) if from_host else b
            loss = model.train_step(x, sl, (lab, ll), global_batch=Bg, allreduce=allreduce)
            if from_host:
                float(loss)

    for b in (res[0], res[-1], res[len(res) // 2]):                          # warm-up on the extreme shapes
        model.train_step(b[0], b[1], (b[2], b[3]), global_batch=Bg, allreduce=allreduce)
    model.check_step()
    sampler = ClockSampler(ctx["local_rank"])
    sampler.start()
    l0 = lib.ctcasr_launch_count()
    ms = timed(lambda: epoch(res, False), 1)
    launches = lib.ctcasr_launch_count() - l0
    clocks = sampler.finish()
    model.check_step()
    ms_e2e = timed(lambda: epoch(dev, True), 1)
    true_frames = int(sum(int(b[1].sum()) for b in batches))
    padded_frames = int(sum(b[0].shape[0] * b[0].shape[1] for b in batches))
    line = {
        "metric": "audio-frames/s", "value": true_frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": 3,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": args.compute, "data": "synthetic",
        "config": {"workload": "cfg4: %d bucketed variable-length batches of %d utterances (0.7-17 s, 69-1699 frames, 96 buckets, padded to the "
                               "longest of the batch), 3 dense + 2 Bi%s-%d + 2 dense, true sequence lengths in the RNN and CTC, full train step; "
                               "value counts TRUE frames (%d of %d padded)" % (K, Bg, CELL_NAMES[args.cell], args.units, true_frames, padded_frames),
                   "global_batch": Bg, "l2_policy": "every batch a different shape and > L2"},
        "e2e": {"value": true_frames / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": int(sum(sum(a.numel() * a.element_size() for a in b) for b in dev) / K), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks,
        "padding_efficiency": true_frames / padded_frames,
    }
    if rank == 0:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------ cfg5: CTC in isolation
def run_ctc(ctx):
    """cfg5: CTC forward-backward in isolation, judged on HBM GB/s."""
    args, torch, lib, ops, synthetic, peaks = ctx["args"], ctx["torch"], ctx["lib"], ctx["ops"], ctx["synthetic"], ctx["peaks"]
    L, V = CFG5["L"], CFG5["V"]
    K, W = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def point(B, T, k, e2e=False):
        rng = np.random.default_rng(B * 7 + T)
        logits = torch.from_numpy((rng.standard_normal((T, B, V)) * 3).astype(np.float32))
        lab, ll = synthetic.make_labels(rng, B, min(L, T // 4), T)
        hl = logits.pin_memory() if e2e else logits
        dl, dlab, dll = hl.cuda(), torch.from_numpy(lab).cuda(), torch.from_numpy(ll).cuda()
        dsl = torch.full((B,), T, dtype=torch.int32, device="cuda")
        grad = torch.empty_like(dl)
        for _ in range(W):
            ops.ctc_loss(dl, dlab, dll, dsl, out_grad=grad)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(k):
            flush.zero_()                                       # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.ctc_loss(dl, dlab, dll, dsl, out_grad=grad)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / k
        alg = B * (2 * T * V * 4 + min(L, T // 4) * 4 + 12)
        out = {"B": B, "T": T, "ms": ms, "algorithmic_gbs": alg / (ms * 1e-3) / 1e9, "frac_of_hbm": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
               "frames_per_s": B * T / (ms * 1e-3)}
        if e2e:
            t0 = time.perf_counter()
            for _ in range(k):
                d = hl.cuda(non_blocking=True)
                loss, g, st = ops.ctc_loss(d, dlab, dll, dsl, out_grad=grad)
                g.cpu()
            torch.cuda.synchronize()
            out["ms_e2e"] = (time.perf_counter() - t0) * 1e3 / k
            out["logits"] = logits
            out["labels"] = (lab, ll)
        return out

    sampler = ClockSampler(ctx["local_rank"])
    sampler.start()
    l0 = lib.ctcasr_launch_count()
    main = point(CFG5["B"], CFG5["T"], K, e2e=True)
    launches = lib.ctcasr_launch_count() - l0
    clocks = sampler.finish()
    B, T = CFG5["B"], CFG5["T"]
    traffic = None
    tr = _profile_json("r2_ctc_dram_traffic.json") or _profile_json("ctc_dram_traffic.json")
    if tr:
        traffic = tr.get("dram_bytes_per_launch")
    line = {
        "metric": "audio-frames/s (CTC forward-backward only)", "value": main["frames_per_s"], "unit": "frames/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": main["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "cfg5: CTC alpha/beta, B=%d, T=%d, L=%d, V=%d" % (B, T, L, V), "l2_policy": "256 MiB flush between iterations"},
        "e2e": {"value": B * T / (main["ms_e2e"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(T * B * V * 4),
                "d2h_bytes_per_step": int(T * B * V * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": main["algorithmic_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": main["frac_of_hbm"], "traffic": traffic, "kernel": "ctc_warp_kernel<6>",
                     "note": "algorithmic bytes = logits in + gradient out + labels (SURVEY 8d); peak of %s" % peaks["source"]},
    }
    if args.sweep:
        line["sweep"] = [{k: v for k, v in point(b, t, 3).items() if k in ("B", "T", "ms", "algorithmic_gbs", "frac_of_hbm", "frames_per_s")}
                         for t in (425, 850, 1700) for b in (64, 128, 256, 512, 1024, 2048)]
    if not args.no_cpu_baseline:
        # the reference's CTCLoss is a CPU op sharded over the batch on the intra-op pool; torch's CPU ctc_loss is
        # an independent implementation of the same recursion with the same threading model
        import torch.nn.functional as F
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        lab, ll = main["labels"]
        lg = main["logits"].clone().requires_grad_(True)
        targets = torch.from_numpy(np.concatenate([lab[b, :ll[b]] for b in range(B)]).astype(np.int64))
        tl, il = torch.from_numpy(ll.astype(np.int64)), torch.full((B,), T, dtype=torch.int64)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            loss = F.ctc_loss(F.log_softmax(lg, 2), targets, il, tl, blank=V - 1, reduction="sum")
            loss.backward()
            ts.append(time.perf_counter() - t0)
            lg.grad = None
        line["cpu_baseline"] = {"value": B * T / ts[-1], "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "the full cfg5 batch once (after one warm-up), torch-CPU log_softmax + ctc_loss forward + backward, %.2f s" % ts[-1]}
    if ctx["rank"] == 0:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------ multi-GPU correctness
def run_check(ctx):
    """SURVEY 8(e) on the hardware: (1) the all-reduced data-parallel gradient equals the 1-GPU gradient of the concatenated global batch
    (1e-5 of its largest entry; summation order differs), for the overlapped bucketed all-reduce and for the single one;
    (2) after 5 training steps the parameters are bit-identical on every rank."""
    args, torch, dist, synthetic = ctx["args"], ctx["torch"], ctx["dist"], ctx["synthetic"]
    rank, world = ctx["rank"], ctx["world"]
    from ctc_asr_b200.model import CTCModel
    assert world > 1, "--check needs torchrun with more than one rank"
    cfg = make_cfg(args, dense_dropout_rate=0.0)
    B, T = args.batch, args.frames
    L = min(CFG2_L, max(1, T // 4))
    model = CTCModel(cfg, seed=1)
    gx, gsl, glab, gll = synthetic.fixed_batch(B * world, T, L, seed=123)       # the same global batch on every rank
    gsl[1::3] = np.maximum(2 * L + 2, T - 7 * np.arange(len(gsl[1::3])) - 5)
    for b in range(B * world):
        gx[b, gsl[b]:] = 0
    full = tuple(torch.from_numpy(a).cuda() for a in (gx, gsl, glab, gll))
    lo, hi = rank * B, (rank + 1) * B
    mine = tuple(a[lo:hi].contiguous() for a in full)
    allreduce = _allreduce(ctx)
    # 1-GPU gradient of the whole global batch
    logits, sl = model.inference_fn(full[0], full[1], training=False)
    loss_full = float(model.loss_fn(logits, sl, (full[2], full[3])))
    model.backward()
    g_full = model.grad_flat.clone()
    errs = {}
    for name, overlap in (("bucketed_overlapped", True), ("single_allreduce", False)):
        logits, sl = model.inference_fn(mine[0], mine[1], training=False)
        loss = model.loss_fn(logits, sl, (mine[2], mine[3]), global_batch=B * world)
        if overlap:
            works = []
            model.backward(on_bucket=lambda a, b: works.append(allreduce(model.grad_flat[a:b], async_op=True)))
            for w in works:
                w.wait()
        else:
            model.backward()
            allreduce(model.grad_flat)
        lsum = loss.detach().clone().reshape(1)
        dist.all_reduce(lsum)
        per = {k: float((model.g[k] - g_full[o:o + model.g[k].numel()].view_as(model.g[k])).abs().max() /
                        g_full[o:o + model.g[k].numel()].abs().max().clamp_min(1e-30)) for k, (o, _) in model.offsets.items()
               if cfg.used_model == "ds1"}
        worst = max(per, key=per.get) if per else None
        errs[name] = {"grad_rel_err": float((model.grad_flat - g_full).abs().max() / g_full.abs().max()),
                      "loss_rel_err": abs(float(lsum) - loss_full) / abs(loss_full),
                      "worst_tensor": worst, "worst_tensor_rel_err": per.get(worst) if worst else None}
    # 5 training steps, then compare the parameters across ranks bit for bit
    tcfg = cfg.replace(dense_dropout_rate=0.1, learning_rate=1e-4)
    tmodel = CTCModel(tcfg, seed=1)
    for _ in range(5):
        tmodel.train_step(mine[0], mine[1], (mine[2], mine[3]), global_batch=B * world, allreduce=allreduce, overlap=args.overlap)
    tmodel.check_step()
    bits = tmodel.flat.view(torch.int32)
    mx, mn = bits.clone(), bits.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(mx, mn))
    moved = float((tmodel.flat - model.flat).abs().max())
    # 1e-5 while the split wgrad sums are short; at the full cfg2 size the two sides are sums over K = T*B = 32,000 .. 256,000
    # bf16x3 products taken in different orders (measured 5e-5 at N = 2, 3e-4 at N = 8): the path's own 1e-3 bar
    tol = 1e-5 if (args.units <= 512 and T <= 256) else 1e-3
    ok = identical and moved > 0 and all(e["grad_rel_err"] < tol and e["loss_rel_err"] < 1e-5 for e in errs.values())
    if rank == 0:
        print(json.dumps({"check": "data-parallel correctness", "n_gpus": world, "ok": ok, "global_batch": B * world, "frames": T,
                          "units": args.units, "cell": args.cell, "compute": args.compute,
                          "gradient_vs_single_gpu": errs, "gradient_tolerance": tol, "params_bit_identical_after_5_steps": identical,
                          "max_param_change": moved}))
    if not ok:
        sys.exit(1)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
