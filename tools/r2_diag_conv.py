"""Diagnostic: the ds2 whole path with 32-filter conv layers, stage by stage against the oracle (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ctc_asr_b200 import synthetic  # noqa: E402
from ctc_asr_b200.model import CTCModel  # noqa: E402
from ctc_asr_b200.params import ModelConfig  # noqa: E402
from oracle import model_ref  # noqa: E402


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(tag, **kw):
    base = dict(used_model="ds2", conv_filters=(32, 32, 64), num_units_dense=64, num_layers_rnn=1, num_units_rnn=64,
                rnn_cell="lstm", num_features=20, cudnn=True, dense_dropout_rate=0.1, conv_dropout_rate=0.1,
                rnn_dropout_rate=0.0, compute="bf16x3", random_seed=5)
    training = kw.pop("training", True)
    base.update(kw)
    cfg = ModelConfig(**base)
    B, T, L = 4, 61, 6
    params = synthetic.init_params(cfg, seed=1)
    rng = np.random.default_rng(0)
    for k in params:
        if k.endswith("bias"):
            params[k] = (rng.standard_normal(params[k].shape) * 0.05).astype(np.float32)
    x, sl, lab, ll = synthetic.fixed_batch(B, T, L, F=cfg.num_features, seed=0)
    sl = np.maximum(T - 7 * np.arange(B), 2 * L + 2).astype(np.int32)
    for b in range(B):
        x[b, sl[b]:] = 0
    model = CTCModel(cfg, params=params)
    logits, sl_out = model.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=training)
    torch.cuda.synchronize()
    ologits, cache = model_ref.forward(cfg, params, x, sl, training=training, seed=int(cfg.random_seed))
    out = [tag]
    crate = cfg.conv_dropout_rate
    from oracle import ref
    for li, (xin, pitch, y, d) in enumerate(model._saved["conv"]):
        x4, oy = cache["conv"][li]
        oyd = ref.dropout(oy, crate, int(cfg.random_seed) + 200 + li, pitch=max(64, (oy.shape[-1] + 7) // 8 * 8))
        got = y.cpu().numpy().reshape(d["To"], B, d["Fo"], d["N"])[..., :oy.shape[-1]]
        e = rel_err(got, oyd)
        mism = int(((got == 0) != (oyd == 0)).sum())
        out.append("conv%d err %.2e zero-mismatch %d/%d" % (li, e, mism, got.size))
    xin, oy, _, _ = cache["rnn"][0]
    h, y, _ = model._saved["rnn"][0]
    out.append("rnn_in err %.2e" % rel_err(h.cpu().numpy().reshape(xin.shape), xin))
    out.append("rnn_y err %.2e" % rel_err(y.cpu().numpy(), oy))
    oh, oy4 = cache["d4"]
    out.append("d4 err %.2e" % rel_err(model._saved["d4"][1].cpu().numpy(), oy4))
    out.append("logits err %.2e" % rel_err(logits.cpu().numpy(), ologits))
    before = logits.clone()
    loss = model.loss_fn(logits, sl_out, (torch.from_numpy(lab), torch.from_numpy(ll)))
    torch.cuda.synchronize()
    out.append("logits changed by loss_fn: %s" % (not torch.equal(before, logits)))
    model.backward()
    torch.cuda.synchronize()
    out.append("logits changed by backward: %s (err now %.2e)" % (not torch.equal(before, logits), rel_err(logits.cpu().numpy(), ologits)))
    oloss, ograds, _, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll, training=training, seed=int(cfg.random_seed))
    out.append("loss err %.2e" % (abs(float(loss) - oloss) / abs(oloss)))
    got = model.grads_numpy()
    out.append("grads " + " ".join("%s=%.1e" % (k.split("/")[-2][-8:] + "/" + k.split("/")[-1][:1], rel_err(got[k], w)) for k, w in ograds.items()))
    print(" | ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run("as-is")
        sys.exit(0)
    run("as-is")
    run("no-conv-drop", conv_dropout_rate=0.0)
    run("no-dense-drop", dense_dropout_rate=0.0)
    run("eval", training=False, conv_dropout_rate=0.0)
    run("fp32", compute="fp32")
    run("bf16x3-filters(8,8,64)", conv_filters=(8, 8, 64))
    run("tanh-tf", rnn_cell="rnn_tanh", cudnn=False)
    run("seed77", random_seed=77)
