"""Seeded synthetic inputs with the shapes/lengths contract of the reference's input pipeline.

The real pipeline (asr/input_functions.py) is host-side I/O and out of scope; what the hot path
sees is restated here: standardised features (asr/input_functions.py:331) padded
with 0.0, label ids in 1..27 padded with 0 (asr/model.py:71 strips them), frame counts from the
25 ms / 10 ms framing (asr/params.py:146-147) and bucket boundaries as in
asr/util/csv_helper.py:29-38.
"""
import math

import numpy as np

from .params import ModelConfig, param_specs, MIN_EXAMPLE_LENGTH, MAX_EXAMPLE_LENGTH


def num_frames(seconds, sampling_rate=16000, win_len=0.025, win_step=0.010):
    """Frames produced by python_speech_features framing for a clip of `seconds`."""
    n, wl, ws = int(round(seconds * sampling_rate)), int(round(win_len * sampling_rate)), int(round(win_step * sampling_rate))
    return 1 if n <= wl else 1 + int(math.ceil((n - wl) / ws))


def init_params(cfg: ModelConfig, seed=1, dtype=np.float32):
    """Random-init weights with the reference's initialisers (asr/model.py:146, :209-212)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, init in param_specs(cfg):
        if init == "zeros":
            a = np.zeros(shape, dtype)
        elif init == "truncnorm":
            a = rng.standard_normal(shape)
            bad = np.abs(a) > 2.0
            while bad.any():                      # tf.truncated_normal re-draws beyond 2 sigma
                a[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(a) > 2.0
            a = (a * 0.046875).astype(dtype)
        elif init[0] == "glorot_normal":
            _, fan_in, fan_out = init
            a = rng.standard_normal(shape)
            bad = np.abs(a) > 2.0
            while bad.any():
                a[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(a) > 2.0
            a = (a * (math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978)).astype(dtype)
        else:
            _, fan_in, fan_out = init
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            a = rng.uniform(-lim, lim, shape).astype(dtype)
        out[name] = a
    return out


def make_labels(rng, B, L, T=None, lmax=None):
    """Label ids uniform in 1..27, 0-padded to lmax.  L: int or per-utterance array.
    If T is given, labels are re-drawn until L + #adjacent repeats <= T (CTC feasibility)."""
    L = np.broadcast_to(np.asarray(L, np.int32), (B,)).copy()
    lmax = int(L.max()) if lmax is None else lmax
    lab = np.zeros((B, max(lmax, 1)), np.int32)
    T = None if T is None else np.broadcast_to(np.asarray(T, np.int32), (B,))
    for b in range(B):
        while True:
            row = rng.integers(1, 28, size=L[b]).astype(np.int32)
            rep = int((row[1:] == row[:-1]).sum()) if L[b] > 1 else 0
            if T is None or L[b] + rep <= T[b]:
                break
        lab[b, :L[b]] = row
    return lab, L


def fixed_batch(B, T, L, F=80, seed=0):
    """cfg1/cfg2-style batch: all utterances T frames, L labels."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    lab, ll = make_labels(rng, B, L, T)
    return x, np.full(B, T, np.int32), lab, ll


def bucket_boundaries(lengths_frames, num_buckets=96):
    """asr/util/csv_helper.py:29-38: every (N // num_buckets)-th sorted length, de-duplicated."""
    ls = sorted(int(v) for v in lengths_frames)
    step = max(len(ls) // num_buckets, 1)
    return sorted(set(ls[i] for i in range(step, len(ls), step)))


def variable_batches(n_utts=6144, batch_size=64, F=80, seed=0, num_buckets=96, max_batches=None):
    """cfg4: durations U(0.7, 17) s, bucketed by length, padded to the longest in the batch
    (asr/input_functions.py:90-98).  Yields (x[B,Tmax,F], seq_len, labels, label_len)."""
    rng = np.random.default_rng(seed)
    dur = np.sort(rng.uniform(MIN_EXAMPLE_LENGTH, MAX_EXAMPLE_LENGTH, n_utts))
    frames = np.array([num_frames(d) for d in dur], np.int32)
    bounds = bucket_boundaries((dur / 0.010).astype(np.int64), num_buckets)
    edges = [0] + [int(np.searchsorted((dur / 0.010).astype(np.int64), b)) for b in bounds] + [n_utts]
    produced = 0
    for lo, hi in zip(edges[:-1], edges[1:]):
        for s in range(lo, hi - batch_size + 1, batch_size):
            T = frames[s:s + batch_size]
            L = np.clip(np.round(16 * dur[s:s + batch_size]).astype(np.int32), 2, T // 2)
            x = rng.standard_normal((batch_size, int(T.max()), F)).astype(np.float32)
            for b in range(batch_size):
                x[b, T[b]:] = 0.0
            lab, ll = make_labels(rng, batch_size, L, T)
            yield x, T.copy(), lab, ll
            produced += 1
            if max_batches is not None and produced >= max_batches:
                return
