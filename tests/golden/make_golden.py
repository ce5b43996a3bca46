"""Generates tests/golden/path_golden.npz — small committed input/output vectors of the whole path.

The reference itself cannot run here (TensorFlow 1.12 is not installable offline), so the
vectors come from the two implementations that do NOT share code with the CUDA path or the C oracle:
  * the torch-CPU restatement of the TF graph (oracle/torch_ref.py: per-timestep loop + autograd +
    F.ctc_loss) in float64 for cfg1 (2 dense + 1 BiRNN-128 tanh, one 1 s utterance) and a small ds2 /
    LSTM model — loss, logits and every gradient tensor;
  * the numpy restatement of python_speech_features (oracle/features_ref.py) for MFCC / log-mel features
    of a synthetic clip;
plus the beam-search oracle's transcripts (pinned separately by enumeration, tests/test_oracle_beam.py).
Inputs are regenerated from seeds by the tests; only outputs (and the seeds) are stored.

  python tests/golden/make_golden.py        (run from the repo root; deterministic)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ctc_asr_b200 import synthetic                              # noqa: E402
from ctc_asr_b200.params import ModelConfig                      # noqa: E402
from oracle import features_ref, ref, torch_ref                  # noqa: E402

CASES = {
    "cfg1": dict(cfg=dict(used_model="ds1", num_layers_dense=2, num_units_dense=128, num_layers_rnn=1, num_units_rnn=128, rnn_cell="rnn_tanh",
                          cudnn=False, dense_dropout_rate=0.0), B=1, T=99, L=16),
    "lstm": dict(cfg=dict(used_model="ds1", num_layers_dense=3, num_units_dense=64, num_layers_rnn=2, num_units_rnn=32, rnn_cell="lstm",
                          cudnn=False, dense_dropout_rate=0.0), B=4, T=60, L=8),
    "ds2": dict(cfg=dict(used_model="ds2", conv_filters=(8, 8, 64), num_units_dense=64, num_layers_rnn=1, num_units_rnn=32,
                         rnn_cell="lstm", num_features=20, cudnn=False, dense_dropout_rate=0.0), B=3, T=41, L=5),
}


def case_inputs(name):
    c = CASES[name]
    cfg = ModelConfig(**c["cfg"])
    params = synthetic.init_params(cfg, seed=1)
    rng = np.random.default_rng(7)
    for k in params:
        if k.endswith("bias"):
            params[k] = (rng.standard_normal(params[k].shape) * 0.05).astype(np.float32)
    x, sl, lab, ll = synthetic.fixed_batch(c["B"], c["T"], c["L"], F=cfg.num_features, seed=3)
    if c["B"] > 1 and cfg.used_model == "ds1":
        sl = np.maximum(c["T"] - 9 * np.arange(c["B"]), 2 * c["L"] + 2).astype(np.int32)
        for b in range(c["B"]):
            x[b, sl[b]:] = 0
    return cfg, params, x, sl, lab, ll


def golden_clip(n=24000, seed=5):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    sig = 2500 * np.sin(2 * np.pi * 210 * t) + 1200 * np.sin(2 * np.pi * 1710 * t + 0.3) + 250 * rng.standard_normal(n)
    sig *= 0.6 + 0.4 * np.sin(2 * np.pi * 2.5 * t) ** 2
    return np.clip(np.round(sig), -32768, 32767).astype(np.int16)


def beam_logits(seed=11, T=50, B=3, V=29):
    return (np.random.default_rng(seed).standard_normal((T, B, V)) * 3).astype(np.float32)


def fingerprint(name, g):
    """||g|| and 8 fixed random projections of a gradient tensor (kept instead of the tensor: small fixture,
    and any element off by more than the tolerance moves a projection)."""
    g = np.asarray(g, np.float64).ravel()
    seed = int.from_bytes(name.encode()[-4:].rjust(4, b"0"), "little") % (2 ** 31)
    r = np.random.default_rng(seed).standard_normal((8, g.size))
    return np.concatenate([[np.linalg.norm(g)], r @ g])


def main():
    out = {}
    for name in CASES:
        cfg, params, x, sl, lab, ll = case_inputs(name)
        p = torch_ref.params_to_torch(params, torch.float64)
        loss, grads, logits = torch_ref.train_step_grads(cfg, p, torch.tensor(x, dtype=torch.float64), torch.tensor(sl),
                                                         torch.tensor(lab), torch.tensor(ll))
        out[name + "/loss"] = np.float64(loss)
        out[name + "/logits"] = logits.numpy().astype(np.float32)
        for k, g in grads.items():
            out[name + "/grad/" + k] = fingerprint(k, g.numpy())
    clip = golden_clip()
    for ft in ("mfcc", "mel"):
        f, n = features_ref.load_sample(clip, ft, "local")
        out["features/" + ft] = f.astype(np.float32)
    x = beam_logits()
    sl = np.array([50, 37, 12], np.int32)
    for W in (8, 1024):
        ids, n, lp = ref.ctc_beam_search(x, sl, beam_width=W)
        out["beam/%d/ids" % W], out["beam/%d/len" % W] = ids, n
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "path_golden.npz"), **out)
    print("wrote %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(ROOT, "tests", "golden", "path_golden.npz")) / 1024))


if __name__ == "__main__":
    main()
