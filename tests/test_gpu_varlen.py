"""BASELINE cfg4: variable-length batches of 64 utterances (0.7-17 s), bucketed by length as the
reference's input pipeline does (asr/input_functions.py:90-103, asr/util/csv_helper.py:29-38), through
the BiRNN with dynamic_rnn(sequence_length) semantics and CTC with the true lengths."""
import numpy as np
import pytest
import torch

from ctc_asr_b200 import ops, synthetic
from ctc_asr_b200.model import CTCModel
from ctc_asr_b200.params import ModelConfig
from oracle import model_ref, ref

from test_gpu_parity import RTOL, dev, rel_err

pytestmark = pytest.mark.gpu


def _bucketed(n_utts, batch_size, which):
    batches = list(synthetic.variable_batches(n_utts=n_utts, batch_size=batch_size, seed=4))
    return batches[which]


@pytest.mark.parametrize("compute,which", [("fp32", 0), ("bf16x3", -1)])
def test_cfg4_bucketed_batch_vs_oracle(compute, which):
    """One batch of 64 from the shortest / the longest bucket, small layers (oracle speed): loss, every
    gradient tensor and the greedy transcripts.  Frames past each utterance's length must not matter."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=64, num_layers_rnn=2, num_units_rnn=64, rnn_cell="lstm",
                      cudnn=False, dense_dropout_rate=0.0, compute=compute)
    x, sl, lab, ll = _bucketed(6144, 64, which)
    assert x.shape[0] == 64 and sl.min() < sl.max() and x.shape[1] == sl.max()
    assert (69 <= sl).all() and (sl <= 1699).all()                     # 0.7 s .. 17 s at 10 ms steps
    params = synthetic.init_params(cfg, seed=1)
    model = CTCModel(cfg, params=params)
    batch = (torch.from_numpy(x), torch.from_numpy(sl), (torch.from_numpy(lab), torch.from_numpy(ll)))
    logits, _ = model.inference_fn(batch[0], batch[1], training=False)
    loss = model.loss_fn(logits, batch[1], batch[2])
    model.backward()
    got = model.grads_numpy()
    oloss, ograds, ologits, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll)
    assert abs(float(loss) - oloss) / abs(oloss) < RTOL
    errs = {k: rel_err(got[k], want) for k, want in ograds.items()}
    assert max(errs.values()) < RTOL, errs
    ids, n = ops.greedy_decode(logits, dev(sl))
    oi, on = ref.greedy_decode(logits.cpu().numpy(), sl)
    assert (n.cpu().numpy() == on).all() and (ids.cpu().numpy() == oi).all()
    # garbage in the padding must not change anything (the input pipeline pads with zeros, but the
    # contract is the length vector)
    x2 = x.copy()
    for b in range(64):
        x2[b, sl[b]:] = 1e3
    logits2, _ = model.inference_fn(torch.from_numpy(x2), batch[1], training=False)
    loss2 = model.loss_fn(logits2, batch[1], batch[2])
    model.backward()
    assert float(loss2) == float(loss)
    got2 = model.grads_numpy()
    assert all(np.array_equal(got2[k], got[k]) for k in got if not k.startswith("dense/dense/"))
    assert all(rel_err(got2[k], got[k]) < 1e-6 for k in got)


def test_cfg4_full_size_longest_bucket():
    """B=64 x up to 17 s on the cfg2 layers (two 32-row slices per LSTM launch): the step runs, finite,
    deterministic; per-utterance losses equal those of the same utterances in a batch of 32."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=2048, num_layers_rnn=2, num_units_rnn=2048, rnn_cell="lstm",
                      cudnn=False, dense_dropout_rate=0.0, compute="bf16x3")
    model = CTCModel(cfg, seed=1)
    x, sl, lab, ll = _bucketed(6144, 64, -1)
    b = tuple(torch.from_numpy(a).cuda() for a in (x, sl, lab, ll))
    logits, _ = model.inference_fn(b[0], b[1], training=False)
    loss = float(model.loss_fn(logits, b[1], (b[2], b[3])))
    model.backward()
    per64, g64 = model.last_per_utterance_loss.clone(), model.grad_flat.clone()
    assert np.isfinite(loss) and bool(torch.isfinite(g64).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    logits, _ = model.inference_fn(b[0], b[1], training=False)
    loss_again = float(model.loss_fn(logits, b[1], (b[2], b[3])))
    model.backward()
    e1.record()
    torch.cuda.synchronize()
    assert loss_again == loss and torch.equal(model.grad_flat, g64)
    ms = e0.elapsed_time(e1)
    print("cfg4 longest bucket: T=%d, %d true frames, fwd+CTC+bwd %.1f ms -> %.0f true frames/s"
          % (x.shape[1], int(sl.sum()), ms, sl.sum() / ms * 1e3))
    half = slice(32, 64)
    logits, _ = model.inference_fn(b[0][half], b[1][half], training=False)
    model.loss_fn(logits, b[1][half], (b[2][half], b[3][half]))
    assert torch.allclose(model.last_per_utterance_loss, per64[half], rtol=1e-5)
