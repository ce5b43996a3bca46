"""GPU parity of the CTC prefix beam search (decode_fn's decoder, asr/model.py:292-296) against the
oracle's restatement of TF's CTCBeamSearchDecoder in its TF-faithful mode (reoffer_wipe=True: including the
order-dependent side effect of TF's Step(), oracle/beam_search.h).  Integer work: transcripts and lengths
bit-exact (the kernel and the oracle evaluate log-sum-exp with the same IEEE operation sequence); the best
path's log-score to 1e-5."""
import numpy as np
import pytest
import torch

from ctc_asr_b200 import ops
from ctc_asr_b200.params import ModelConfig
from oracle import ref

from test_gpu_parity import dev

pytestmark = pytest.mark.gpu


def _check(x, sl, W, merge=False):
    ids, n, lp = ops.beam_search(dev(x), dev(sl, torch.int32), beam_width=W, merge_repeated=merge)
    torch.cuda.synchronize()
    oi, on, olp = ref.ctc_beam_search(x, sl, beam_width=W, merge_repeated=merge, reoffer_wipe=True)
    ids, n, lp = ids.cpu().numpy(), n.cpu().numpy(), lp.cpu().numpy()
    assert np.array_equal(n, on), (n, on)
    assert np.array_equal(ids, oi)
    np.testing.assert_allclose(lp, olp, rtol=1e-5, atol=1e-5)
    return ids, n


@pytest.mark.parametrize("W", [1, 7, 64, 300, 1024])
@pytest.mark.parametrize("scale", [1.0, 5.0])
def test_beam_search_vs_oracle(W, scale):
    rng = np.random.default_rng(W)
    T, B, V = 60, 6, 29
    x = (rng.standard_normal((T, B, V)) * scale).astype(np.float32)
    sl = np.array([60, 41, 0, 1, 17, 60], np.int32)
    ids, n = _check(x, sl, W)
    assert n[2] == 0 and (ids[2] == -1).all()
    assert ((ids >= -1) & (ids < V - 1)).all()
    _check(x, sl, W, merge=True)


def test_beam_search_wide_beam_small_alphabet_and_long_utterance():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((9, 4, 5)) * 2).astype(np.float32)        # every prefix fits in the beam for a while
    _check(x, np.array([9, 5, 3, 9], np.int32), 1024)
    x = (rng.standard_normal((400, 2, 29)) * 4).astype(np.float32)     # many frames: node table, re-entering prefixes
    _check(x, np.array([400, 333], np.int32), 128)
    x = (rng.standard_normal((150, 2, 29)) * 2).astype(np.float32)     # the reference's width
    _check(x, np.array([150, 150], np.int32), 1024)


def test_decode_fn_uses_the_reference_decoder_at_full_size():
    """decode_fn on a cfg2-sized batch (B=32, T=1000, beam_width 1024): runs, is deterministic, never
    scores below the greedy path's own prefix, and equals greedy decoding on sharply peaked logits."""
    from ctc_asr_b200.model import CTCModel
    rng = np.random.default_rng(5)
    T, B, V = 1000, 32, 29
    cls = rng.integers(0, V, (T, B))
    cls[rng.random((T, B)) < 0.6] = V - 1                              # mostly blanks, like a trained model
    x = (rng.standard_normal((T, B, V))).astype(np.float32)
    np.put_along_axis(x, cls[..., None], 12.0, axis=2)                 # winner 12 above the noise
    sl = np.full(B, T, np.int32); sl[3] = 517
    logits, seq = dev(x), dev(sl, torch.int32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ids, n, lp = ops.beam_search(logits, seq, beam_width=1024)
    e0.record()
    ids2, n2, lp2 = ops.beam_search(logits, seq, beam_width=1024)
    e1.record()
    torch.cuda.synchronize()
    print("beam search B=32 x T=1000, width 1024: %.1f ms" % e0.elapsed_time(e1))
    assert torch.equal(ids, ids2) and torch.equal(n, n2) and torch.equal(lp, lp2)
    gi, gn = ops.greedy_decode(logits, seq)
    assert torch.equal(n, gn)
    for b in range(B):
        assert torch.equal(ids[b, :n[b]], gi[b, :gn[b]])
    cfg = ModelConfig(used_model="ds1", num_layers_rnn=1, num_units_rnn=64, num_units_dense=64, compute="fp32")
    model = CTCModel(cfg, seed=1)
    decoded, plaintext, summary = model.decode_fn(logits, seq, originals=["x"] * B)
    assert len(decoded) == B and all(torch.equal(decoded[b], ids[b, :n[b]].cpu()) for b in range(B))
    assert all(isinstance(s, str) for s in plaintext) and summary[0][1] == "x"
    d2, _, _ = model.decode_fn(logits, seq, decoder="greedy")
    assert all(torch.equal(a, c) for a, c in zip(decoded, d2))


def test_tf_step_order_artifact_is_reproduced():
    """TF r1.12's Step() wipes a prefix that is pushed out of the beam and re-offered inside the grow loop (oracle
    reoffer_wipe=True).  On flat random logits, where thousands of prefixes score within a few nats, that changes a few
    transcripts relative to the order-independent beam (reoffer_wipe=False); the kernel detects those frames, replays
    TF's sequential loop for them and must equal the TF-faithful mode everywhere — and this test only means something
    if the two oracle modes do differ on its inputs."""
    rng = np.random.default_rng(11)
    T, B, V = 50, 32, 29
    differ = 0
    for W, scale in ((16, 1.0), (256, 1.0), (256, 3.0), (64, 0.5)):
        x = (rng.standard_normal((T, B, V)) * scale).astype(np.float32)
        sl = np.full(B, T, np.int32)
        ids, n = _check(x, sl, W)                                   # == reoffer_wipe=True, bit for bit
        pi, pn, _ = ref.ctc_beam_search(x, sl, beam_width=W, reoffer_wipe=False)
        differ += sum(int(n[b] != pn[b] or not np.array_equal(ids[b, :n[b]], pi[b, :pn[b]])) for b in range(B))
    print("utterances on which TF's artifact changes the transcript: %d of %d" % (differ, 4 * B))
    assert differ > 0
