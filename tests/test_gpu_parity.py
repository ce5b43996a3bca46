"""GPU parity tests proper: every call goes through the C-ABI (libctcasr.so) and is checked against
the CPU oracle (oracle/) on the same seeded inputs, against the committed golden vectors, and — at
BASELINE.json's full sizes — through size-independent properties.

Tolerances (north_star): integer/index work bit-exact; floating point within 1e-3 relative
(gradients: max abs error normalised by max |oracle gradient|)."""
import json
import os

import numpy as np
import pytest
import torch

from ctc_asr_b200 import _lib, ops, synthetic
from ctc_asr_b200.params import ModelConfig
from oracle import model_ref, ref

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ------------------------------------------------------------------------------------------ CTC
def _ctc_gpu(logits, labels, ll, sl, blank=None, scale=1.0):
    loss, grad, status = ops.ctc_loss(dev(logits, torch.float32), dev(labels, torch.int32), dev(ll, torch.int32),
                                      dev(sl, torch.int32), blank=blank, grad_scale=scale)
    torch.cuda.synchronize()
    return loss.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()


def test_ctc_tf_known_answer(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ctc_tf_known_answer.json")))
    T, V, utts = g["T"], g["V"], g["utterances"]
    B = len(utts)
    logits = np.zeros((T, B, V), np.float32)
    labels = np.zeros((B, 5), np.int32)
    ll = np.zeros(B, np.int32)
    for b, u in enumerate(utts):
        logits[:, b] = np.log(np.asarray(u["probs"]))
        labels[b, :len(u["labels"])] = u["labels"]
        ll[b] = len(u["labels"])
    loss, grad, status = _ctc_gpu(logits, labels, ll, np.full(B, T, np.int32), blank=g["blank"])
    assert (status == 0).all()
    for b, u in enumerate(utts):
        assert abs(loss[b] - u["loss"]) < 1e-4 * u["loss"]
    np.testing.assert_allclose(grad[0, 0], utts[0]["grad_row0"], atol=2e-5)


def test_ctc_host_buffer_entry_point(golden_dir):
    """ctcasr_ctc_loss_host: the drop-in for TF's CPU CTCLoss op (host logits in, host grad out)."""
    rng = np.random.default_rng(1)
    T, B, V = 40, 5, 29
    logits = (rng.standard_normal((T, B, V)) * 2).astype(np.float32)
    labels, ll = synthetic.make_labels(rng, B, np.array([3, 9, 1, 12, 5]), 40)
    sl = np.array([40, 33, 7, 40, 21], np.int32)
    loss, grad, status = np.zeros(B, np.float32), np.zeros_like(logits), np.zeros(B, np.int32)
    lib = _lib.load()
    vp = lambda a: a.ctypes.data_as(__import__("ctypes").c_void_p)
    rc = lib.ctcasr_ctc_loss_host(vp(logits), T, B, V, V - 1, vp(labels), labels.shape[1], vp(ll), vp(sl),
                                  vp(loss), vp(grad), 1.0, vp(status))
    assert rc == 0, lib.ctcasr_last_error()
    ol, og, _ = ref.ctc_loss(logits.astype(np.float64), labels, ll, sl)
    assert rel_err(loss, ol) < RTOL and rel_err(grad, og) < RTOL


@pytest.mark.parametrize("T,B,L", [(50, 7, 20), (200, 9, 30), (333, 4, 100)])
def test_ctc_ragged_vs_oracle(T, B, L):
    rng = np.random.default_rng(T)
    V = 29
    logits = (rng.standard_normal((T, B, V)) * 3).astype(np.float32)
    sl = rng.integers(1, T + 1, B).astype(np.int32)
    sl[0] = T
    ll = np.minimum(rng.integers(0, L + 1, B), sl // 2).astype(np.int32)
    labels, ll = synthetic.make_labels(rng, B, ll, sl, lmax=L)
    loss, grad, status = _ctc_gpu(logits, labels, ll, sl, scale=0.25)
    ol, og, ost = ref.ctc_loss(logits.astype(np.float64), labels, ll, sl)
    assert (status == ost).all() and (status == 0).all()
    assert rel_err(loss, ol) < RTOL
    assert rel_err(grad, og * 0.25) < RTOL
    for b in range(B):
        assert not grad[sl[b]:, b].any()            # zero gradient past the utterance's length


def test_ctc_long_labels_multi_state_per_thread():
    """L = 422 is the corpus maximum (README.md:169): S = 845 > 480 threads per group."""
    rng = np.random.default_rng(9)
    T, B, V, L = 900, 2, 29, 422
    logits = (rng.standard_normal((T, B, V)) * 2).astype(np.float32)
    labels, ll = synthetic.make_labels(rng, B, np.array([422, 300]), T)
    sl = np.array([900, 700], np.int32)
    loss, grad, status = _ctc_gpu(logits, labels, ll, sl)
    ol, og, _ = ref.ctc_loss(logits.astype(np.float64), labels, ll, sl)
    assert (status == 0).all() and rel_err(loss, ol) < RTOL and rel_err(grad, og) < RTOL


def test_ctc_bench_length_stays_within_tolerance():
    """cfg5 geometry (T=1700, L=84) with flat random logits, the hardest case for a log-domain fp32
    recursion (TF's own: 1e-2 of gradient error, tests/test_oracle_ctc.py).  The kernel carries
    alpha/beta as (hi, lo) float pairs with error-free additions and stays within 1e-3."""
    rng = np.random.default_rng(5)
    T, B, V = 1700, 3, 29
    logits = (rng.standard_normal((T, B, V)) * 3).astype(np.float32)
    labels, ll = synthetic.make_labels(rng, B, 84, T)
    sl = np.full(B, T, np.int32)
    loss, grad, status = _ctc_gpu(logits, labels, ll, sl)
    ol, og, _ = ref.ctc_loss(logits.astype(np.float64), labels, ll, sl)
    assert (status == 0).all() and rel_err(loss, ol) < 1e-5 and rel_err(grad, og) < RTOL


def test_ctc_error_statuses_and_empty():
    rng = np.random.default_rng(0)
    T, B, V = 4, 5, 5
    logits = rng.standard_normal((T, B, V)).astype(np.float32)
    labels = np.array([[1, 1, 1], [4, 0, 0], [1, 2, 0], [1, 0, 0], [0, 0, 0]], np.int32)
    ll = np.array([3, 1, 2, 1, 0], np.int32)
    sl = np.array([4, 4, 9, 0, 0], np.int32)
    loss, grad, status = _ctc_gpu(logits, labels, ll, sl, blank=4)
    ol, og, ost = ref.ctc_loss(logits.astype(np.float64), labels, ll, sl, blank=4)
    assert status.tolist() == ost.tolist() == [1, 2, 3, 1, 0]
    assert np.isinf(loss[:4]).all() and loss[4] == 0.0 and not grad.any()


def test_ctc_full_size_properties():
    """cfg5 at full size (B=512, T=1700, L=84): per-frame gradient rows sum to 0 (softmax minus a
    posterior that sums to 1) and a seeded sample of utterances matches the oracle."""
    rng = np.random.default_rng(0)
    T, B, V, L = 1700, 512, 29, 84
    logits = (rng.standard_normal((T, B, V)) * 3).astype(np.float32)
    labels, ll = synthetic.make_labels(rng, B, L, T)
    sl = np.full(B, T, np.int32)
    lg = dev(logits)
    loss, grad, status = ops.ctc_loss(lg, dev(labels), dev(ll), dev(sl))
    assert int(status.abs().sum()) == 0
    rowsum = grad.sum(2).abs().max().item()
    assert rowsum < 1e-4, rowsum
    pick = [0, 17, 255, 511]
    ol, og, _ = ref.ctc_loss(logits[:, pick].astype(np.float64), labels[pick], ll[pick], sl[pick])
    assert rel_err(loss.cpu().numpy()[pick], ol) < 1e-5
    assert rel_err(grad.cpu().numpy()[:, pick], og) < RTOL


def test_greedy_decode_bit_exact():
    rng = np.random.default_rng(2)
    T, B, V = 300, 33, 29
    logits = rng.standard_normal((T, B, V)).astype(np.float32)
    logits[5, 0, 3] = logits[5, 0, 7] = 9.0          # exact tie: first max wins
    logits[10:40, 1, :] = 0.0                        # all-equal frames -> class 0 repeated
    sl = rng.integers(0, T + 1, B).astype(np.int32)
    sl[0] = T
    ids, n = ops.greedy_decode(dev(logits), dev(sl))
    oi, on = ref.greedy_decode(logits, sl)
    assert (n.cpu().numpy() == on).all()
    assert (ids.cpu().numpy() == oi).all()


# ------------------------------------------------------------------------------------ GEMM / dense
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_simt_all_orientations(ta, tb):
    rng = np.random.default_rng(3)
    M, N, K = 77, 45, 130
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), ta=ta, tb=tb).cpu().numpy()
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    assert rel_err(c, want) < 1e-5


@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_simt_split_k_long_contraction(accumulate):
    """The logits-layer weight gradient shape: few output tiles, a very long K.  With a scratch arena
    present the SIMT kernel slices K over CTAs and adds the slices in a fixed order (bit-reproducible)."""
    rng = np.random.default_rng(31)
    ops.gemm(dev(rng.standard_normal((256, 256)).astype(np.float32)), dev(rng.standard_normal((256, 256)).astype(np.float32)),
             compute=_lib.COMPUTE_BF16X3)                     # makes ops allocate the arena
    M, N, K = 200, 29, 9000
    a = rng.standard_normal((K, M)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    outs = []
    for _ in range(2):
        out = dev(c0.copy())
        ops.gemm(dev(a), dev(b), ta=True, out=out, accumulate=accumulate)
        outs.append(out.cpu().numpy())
    want = a.T.astype(np.float64) @ b.astype(np.float64) + (c0 if accumulate else 0.0)
    assert rel_err(outs[0], want) < 1e-5
    assert np.array_equal(outs[0], outs[1])


def test_gemm_bf16x3_scratch_arena_grows_for_both_operands():
    """Starting without an arena, one call must report the need of BOTH operands' pieces (a need reported
    operand by operand made the single grow-and-retry of the host wrapper fail on large first calls)."""
    import ctypes
    lib = _lib.load()
    _lib.check(lib.ctcasr_set_scratch(ctypes.c_void_p(0), 0), "set_scratch")
    ops._scratch.clear()
    rng = np.random.default_rng(32)
    a = rng.standard_normal((1024, 1536)).astype(np.float32)
    b = rng.standard_normal((1536, 1024)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), compute=_lib.COMPUTE_BF16X3).cpu().numpy()
    assert rel_err(c, a.astype(np.float64) @ b.astype(np.float64)) < 2e-5


@pytest.mark.parametrize("rate", [0.0, 0.3])
def test_dense_fwd_bwd_vs_oracle(rate):
    rng = np.random.default_rng(4)
    M, K, N = 150, 80, 96
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    y = ops.dense_fwd(dev(x), dev(w), dev(b), act=1, cutoff=2.0, drop_rate=rate, seed=7)
    oy = ref.dense_fwd(x.astype(np.float64), w, b, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    assert rel_err(y.cpu().numpy(), oy) < 1e-5
    assert ((y.cpu().numpy() == 0) == (oy == 0)).mean() > 0.9999      # identical keep / clip masks
    dw, db, dx = torch.empty(K, N).cuda(), torch.empty(N).cuda(), torch.empty(M, K).cuda()
    ops.dense_bwd(dev(x), dev(w), y, dev(dy), dw, db, dx=dx, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    odx, odw, odb = ref.dense_bwd(x.astype(np.float64), w, oy, dy, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    assert rel_err(dw.cpu().numpy(), odw) < 1e-4
    assert rel_err(db.cpu().numpy(), odb) < 1e-4
    assert rel_err(dx.cpu().numpy(), odx) < 1e-4


# ------------------------------------------------------------------------------------------ BiRNN
@pytest.mark.parametrize("cell", ["rnn_tanh", "rnn_relu", "lstm", "gru"])
@pytest.mark.parametrize("use_len", [True, False])
@pytest.mark.parametrize("shape", [(23, 5, 12, 20), (17, 7, 16, 64), (9, 40, 8, 128)])
def test_birnn_layer_vs_oracle(cell, use_len, shape):
    """Stepwise recurrent path (fp32): H = 20 runs the generic SIMT product per frame, H = 64 / 128 the
    double-buffered small-batch kernel (step_gemm.cu); B = 40 spans two 32-row blocks."""
    rng = np.random.default_rng(6)
    T, B, nin, H = shape
    cid = ref.CELL_IDS[cell]
    G = ref.NUM_GATES[cid]
    x = rng.standard_normal((T, B, nin)).astype(np.float32)
    sl = np.maximum(1, T - (5 * np.arange(B)) % T).astype(np.int32)
    sl[-1] = 1
    wx = (rng.standard_normal((nin, 2 * G * H)) * 0.3).astype(np.float32)
    wh = (rng.standard_normal((2, H, G * H)) * (0.3 if H <= 32 else 1.0 / np.sqrt(H))).astype(np.float32)
    bias = (rng.standard_normal(2 * G * H + (2 * H if cell == "gru" else 0)) * 0.1).astype(np.float32)
    dy = rng.standard_normal((T, B, 2 * H)).astype(np.float32)
    rb, _ = ops.birnn_sizes(T, B, nin, H, cid)
    reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
    y = torch.empty((T, B, 2 * H), device="cuda")
    X, SL, WX, WH, BI = dev(x), dev(sl), dev(wx), dev(wh), dev(bias)
    ops.birnn_fwd(X, SL, WX, WH, BI, y, reserve, cid, use_len)
    oy, og, oc = ref.birnn_fwd(x.astype(np.float64), sl, wx, wh, bias, cid, use_len=use_len)
    assert rel_err(y.cpu().numpy(), oy) < 1e-5
    dx = torch.empty((T, B, nin), device="cuda")
    dwx, dwh, db = torch.empty_like(WX), torch.empty_like(WH), torch.empty_like(BI)
    ops.birnn_bwd(X, SL, WX, WH, y, reserve, dev(dy), dx, dwx, dwh, db, cid, use_len)
    odx, odwx, odwh, odb = ref.birnn_bwd(x.astype(np.float64), sl, wx, wh, oy, og, oc, dy, cid, use_len=use_len)
    for got, want, name in [(dx, odx, "dx"), (dwx, odwx, "dwx"), (dwh, odwh, "dwh"), (db, odb, "dbias")]:
        assert rel_err(got.cpu().numpy(), want) < 1e-4, name


def test_adam_vs_oracle():
    rng = np.random.default_rng(8)
    n = 1003
    p, g = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    P, M, V, G = dev(p), dev(m), dev(v), dev(g)
    po, mo, vo = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for step in (1, 2, 3):
        ops.adam(P, M, V, G, step, 1e-3, 0.9, 0.999, 1e-8)
        ref.adam(po, mo, vo, g.astype(np.float64), step, lr=1e-3)
    # (1 - beta2) evaluated in fp32 (as TF does) differs from the fp64 oracle by 1.3e-5 relative
    assert rel_err(P.cpu().numpy(), po) < 1e-6 and rel_err(V.cpu().numpy(), vo) < 1e-4


# --------------------------------------------------------------------------------- whole hot path
def _whole_path(cfg, B, T, L, ragged, seed=0, gtol=RTOL, ltol=RTOL, training=False):
    from ctc_asr_b200.model import CTCModel
    params = synthetic.init_params(cfg, seed=1)
    rng = np.random.default_rng(seed)
    for k in params:                                   # biases away from 0 so every term matters
        if k.endswith("bias"):
            params[k] = (rng.standard_normal(params[k].shape) * 0.05).astype(np.float32)
    x, sl, lab, ll = synthetic.fixed_batch(B, T, L, F=cfg.num_features, seed=seed)
    if ragged:
        sl = np.maximum(T - 7 * np.arange(B), 2 * L + 2).astype(np.int32)
        for b in range(B):
            x[b, sl[b]:] = 0
    model = CTCModel(cfg, params=params)
    logits, sl_out = model.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=training)
    loss = model.loss_fn(logits, sl_out, (torch.from_numpy(lab), torch.from_numpy(ll)))
    model.backward()
    torch.cuda.synchronize()
    # training: the dropout keep-masks are the counter hash both sides share, seeded by FLAGS.random_seed at step 0
    oloss, ograds, ologits, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll, training=training,
                                                         seed=int(cfg.random_seed))
    sl = sl_out.cpu().numpy()                          # ds2: the conv length for every utterance
    assert rel_err(logits.cpu().numpy(), ologits) < ltol
    assert abs(float(loss) - oloss) / abs(oloss) < ltol
    got = model.grads_numpy()
    errs = {k: rel_err(got[k], want) for k, want in ograds.items()}
    print("gradient max-rel-err per tensor:", {k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) < gtol, errs
    # greedy ids on the SAME logits are bit-exact (integer work)
    ids, n = ops.greedy_decode(logits, dev(sl))
    oi, on = ref.greedy_decode(logits.cpu().numpy(), sl)
    assert (n.cpu().numpy() == on).all() and (ids.cpu().numpy() == oi).all()
    return model


@pytest.mark.parametrize("rate", [0.0, 0.25])
def test_dropout_op_vs_oracle(rate):
    """ctcasr_dropout: the keep-mask is integer work (bit-exact against the oracle's hash), kept values are x / (1 - rate)."""
    rng = np.random.default_rng(12)
    x = (rng.standard_normal((333, 40)) + 3.0).astype(np.float32)
    y = ops.dropout(dev(x), rate, 987654321)
    want = ref.dropout(x.astype(np.float64), rate, 987654321)
    got = y.cpu().numpy()
    assert np.array_equal(got == 0, want == 0)
    assert rel_err(got, want) < 1e-6
    if rate:
        assert 0.2 < (got == 0).mean() < 0.3
    z = dev(x)
    ops.dropout(z, rate, 987654321, out=z)                                # in place: the backward pass on a gradient
    assert torch.equal(z, y)


@pytest.mark.parametrize("cudnn,cell,layers", [(False, "rnn_tanh", 2), (True, "rnn_relu", 3), (True, "lstm", 2)])
def test_whole_path_with_rnn_and_dense_dropout(cudnn, cell, layers):
    """Training mode with every dropout of the ds1 path on: dense (asr/util/tf_contrib.py:58), the RNN stack's
    DropoutWrapper in / out masks (TF path, :190-194) or inter-layer dropout (cuDNN path, asr/model.py:201-206)."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=64, num_layers_rnn=layers, num_units_rnn=32,
                      rnn_cell=cell, cudnn=cudnn, dense_dropout_rate=0.1, rnn_dropout_rate=0.2, compute="fp32",
                      random_seed=4321)
    _whole_path(cfg, B=4, T=50, L=6, ragged=True, training=True)


def test_rnn_dropout_is_off_in_eval_and_changes_training():
    cfg = ModelConfig(used_model="ds1", num_layers_dense=1, num_units_dense=64, num_layers_rnn=2, num_units_rnn=32,
                      rnn_cell="rnn_relu", cudnn=True, dense_dropout_rate=0.0, rnn_dropout_rate=0.5, compute="fp32")
    from ctc_asr_b200.model import CTCModel
    x, sl, lab, ll = synthetic.fixed_batch(3, 40, 5, F=cfg.num_features, seed=3)
    m = CTCModel(cfg, params=synthetic.init_params(cfg, seed=1))
    ev = m.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=False)[0].clone()
    m0 = CTCModel(cfg.replace(rnn_dropout_rate=0.0), params=synthetic.init_params(cfg, seed=1))
    assert torch.equal(ev, m0.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=True)[0])
    tr = m.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=True)[0]
    assert not torch.allclose(ev, tr)


def test_cfg1_ds1_tanh_single_utterance():
    """BASELINE cfg1: 2 dense + 1 BiRNN-128 (tanh), one 1 s utterance (99 frames), 80 mel bins."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=128, num_layers_rnn=1, num_units_rnn=128,
                      rnn_cell="rnn_tanh", cudnn=False, dense_dropout_rate=0.0, compute="fp32")
    _whole_path(cfg, B=1, T=99, L=16, ragged=False)


@pytest.mark.parametrize("cudnn", [False, True])
def test_small_3d2r2d_lstm_ragged(cudnn):
    """The cfg2 layout (3 dense + 2 BiLSTM + 2 dense) at a size the oracle finishes in seconds."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=64, num_layers_rnn=2, num_units_rnn=32,
                      rnn_cell="lstm", cudnn=cudnn, dense_dropout_rate=0.0, compute="fp32")
    _whole_path(cfg, B=4, T=60, L=8, ragged=True)


@pytest.mark.parametrize("cell", ["gru", "rnn_relu"])
def test_whole_path_other_cells(cell):
    """The rest of the reference's rnn_cell menu (asr/params.py:48-50): GRU and the default ReLU RNN."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=64, num_layers_rnn=2, num_units_rnn=24,
                      rnn_cell=cell, cudnn=True, dense_dropout_rate=0.0, compute="fp32")
    _whole_path(cfg, B=3, T=30, L=5, ragged=False)


def test_train_step_decreases_loss_and_matches_oracle_adam():
    from ctc_asr_b200.model import CTCModel
    cfg = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=48, num_layers_rnn=1, num_units_rnn=24,
                      rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, learning_rate=1e-3, compute="fp32")
    params = synthetic.init_params(cfg, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(3, 40, 6, seed=3)
    model = CTCModel(cfg, params=params)
    batch = (torch.from_numpy(x), torch.from_numpy(sl), (torch.from_numpy(lab), torch.from_numpy(ll)))
    l0 = float(model.train_step(*batch))
    # one oracle step: same gradients -> same Adam update
    _, og, _, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll)
    for k in ("logits/dense/kernel", "rnn/l0/wh"):
        p = params[k].astype(np.float64).ravel().copy()
        m, v = np.zeros_like(p), np.zeros_like(p)
        ref.adam(p, m, v, og[k].ravel().copy(), 1, lr=1e-3)
        got = model.params_numpy()[k].ravel()
        assert np.mean(np.abs(got - p) < 2e-5) > 0.99, k     # |g| ~ eps elements may differ
    for _ in range(5):
        l1 = float(model.train_step(*batch))
    assert l1 < l0


# ------------------------------------------------------------------- tcgen05 path (compute = tf32)
TF32 = _lib.COMPUTE_TF32


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(384, 512, 256), (200, 320, 80), (1000, 264, 1048)])
def test_gemm_tcgen05_all_orientations_and_tails(ta, tb, M, N, K):
    """TMA + tcgen05.mma kind::tf32 + TMEM epilogue, K-major and MN-major operands, partial tiles in
    M, N and K.  TF32 operands are rounded to nearest by the TMA unit: noise ~3e-4 of max, no bias."""
    rng = np.random.default_rng(M + N)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), ta=ta, tb=tb, compute=TF32).cpu().numpy().astype(np.float64)
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    assert rel_err(c, want) < 1e-3
    bias = ((c - want) * np.sign(want)).mean() / np.abs(want).mean()
    assert abs(bias) < 5e-5, bias


def test_gemm_tcgen05_accumulate_and_strided_views():
    rng = np.random.default_rng(1)
    M, N, K = 300, 256, 512
    big_a = rng.standard_normal((M, 2 * K)).astype(np.float32)          # A = right half: lda = 2K
    b = rng.standard_normal((K, N)).astype(np.float32)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    A = dev(big_a)[:, K:]
    C = dev(c0)
    lib = _lib.load()
    import ctypes
    _lib.check(lib.ctcasr_gemm(ops.ptr(A), ops.ptr(dev(b)), ops.ptr(C), M, N, K, 0, 0, 2 * K, N, N, 1, TF32,
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "gemm")
    want = c0 + big_a[:, K:].astype(np.float64) @ b.astype(np.float64)
    assert rel_err(C.cpu().numpy(), want) < 1e-3


@pytest.mark.parametrize("rate", [0.0, 0.3])
def test_dense_tcgen05_epilogues_vs_oracle(rate):
    rng = np.random.default_rng(4)
    M, K, N = 640, 256, 512
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.1).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    y = ops.dense_fwd(dev(x), dev(w), dev(b), act=1, cutoff=2.0, drop_rate=rate, seed=7, compute=TF32)
    oy = ref.dense_fwd(x.astype(np.float64), w, b, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    assert rel_err(y.cpu().numpy(), oy) < 2e-3
    dw, db, dx = torch.empty(K, N).cuda(), torch.empty(N).cuda(), torch.empty(M, K).cuda()
    ops.dense_bwd(dev(x), dev(w), y, dev(dy), dw, db, dx=dx, act=1, cutoff=2.0, drop_rate=rate, seed=7, compute=TF32)
    # the oracle gets the GPU's forward output so both sides use the same activation mask
    odx, odw, odb = ref.dense_bwd(x.astype(np.float64), w, y.cpu().numpy().astype(np.float64), dy, act=1, cutoff=2.0,
                                  drop_rate=rate, seed=7)
    assert rel_err(dw.cpu().numpy(), odw) < 1e-3
    assert rel_err(db.cpu().numpy(), odb) < 1e-4
    assert rel_err(dx.cpu().numpy(), odx) < 1e-3


def test_whole_path_tf32_lstm():
    """3d2r2d LSTM at a size where every GEMM but the logits layer runs on tcgen05 (tf32)."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=256, num_layers_rnn=2, num_units_rnn=64,
                      rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, compute="tf32")
    # TF32 operands carry 2^-11 relative rounding noise per GEMM input (measured 3e-4 of max per
    # GEMM output); seven chained GEMMs plus the CTC posterior's sensitivity to the logits put the
    # gradients at a few 1e-3 of max, and ReLU-mask flips in the dense stack put single columns of
    # the lowest layers at a few 1e-2 (measured 3.4e-2).  tf32 is the optional fast mode; the 1e-3
    # bar of north_star is met by compute='bf16x3' (default, benchmarked) and compute='fp32'.
    _whole_path(cfg, B=8, T=64, L=10, ragged=True, gtol=6e-2)


def test_whole_path_bf16_cfg3_arithmetic():
    """BASELINE cfg3 ("DS2 bf16"): GEMM operands rounded to bf16 (one tcgen05 product, fp32 accumulation), fp32
    master weights, fp32 CTC, bf16x3 recurrence.  This is reduced precision by construction: the tolerance
    written here is bf16's (2^-9 operand rounding per GEMM input, mask flips in the dense stack), not the
    1e-3 bar, which compute='bf16x3' and 'fp32' meet."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=256, num_layers_rnn=2, num_units_rnn=64,
                      rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, compute="bf16")
    _whole_path(cfg, B=8, T=64, L=10, ragged=True, gtol=2e-1, ltol=2e-2)     # measured: 1.2e-1 (first dense kernel), 4e-3 (RNN)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, True)])
def test_gemm_bf16_single_product(ta, tb):
    rng = np.random.default_rng(9)
    M, N, K = 1000, 264, 1048
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), ta=ta, tb=tb, compute=_lib.COMPUTE_BF16).cpu().numpy().astype(np.float64)
    # exact reference of what the kernel computes: operands rounded to bf16, products summed in high precision
    ar = torch.from_numpy(a).to(torch.bfloat16).to(torch.float64).numpy()
    br = torch.from_numpy(b).to(torch.bfloat16).to(torch.float64).numpy()
    want = (ar.T if ta else ar) @ (br.T if tb else br)
    assert rel_err(c, want) < 1e-5                       # only the fp32 accumulation order differs
    full = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    assert 1e-4 < rel_err(c, full) < 1e-2                # and it really is bf16 arithmetic


# ------------------------------------------------- tcgen05 path, fp32-accurate (compute = bf16x3)
BF16X3 = _lib.COMPUTE_BF16X3


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(384, 512, 256), (200, 320, 80), (1000, 264, 1048)])
def test_gemm_bf16x3_all_orientations_and_tails(ta, tb, M, N, K):
    """Operands split into two bf16 pieces, three kind::f16 products accumulated in fp32 TMEM:
    error ~2^-16 per product, i.e. ~30x below TF32."""
    rng = np.random.default_rng(M + N + 1)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), ta=ta, tb=tb, compute=BF16X3).cpu().numpy().astype(np.float64)
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    assert rel_err(c, want) < 3e-5


@pytest.mark.parametrize("rate", [0.0, 0.3])
def test_dense_bf16x3_vs_oracle(rate):
    """Forward through a ReLU kink uses the 3-piece / 6-product split: fp32-level pre-activations,
    so the keep/clip masks agree with the oracle's."""
    rng = np.random.default_rng(4)
    M, K, N = 640, 256, 512
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.1).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    y = ops.dense_fwd(dev(x), dev(w), dev(b), act=1, cutoff=2.0, drop_rate=rate, seed=7, compute=BF16X3)
    oy = ref.dense_fwd(x.astype(np.float64), w, b, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    assert rel_err(y.cpu().numpy(), oy) < 1e-5
    assert ((y.cpu().numpy() == 0) == (oy == 0)).mean() > 0.9999
    dw, db, dx = torch.empty(K, N).cuda(), torch.empty(N).cuda(), torch.empty(M, K).cuda()
    ops.dense_bwd(dev(x), dev(w), y, dev(dy), dw, db, dx=dx, act=1, cutoff=2.0, drop_rate=rate, seed=7, compute=BF16X3)
    odx, odw, odb = ref.dense_bwd(x.astype(np.float64), w, oy, dy, act=1, cutoff=2.0, drop_rate=rate, seed=7)
    assert rel_err(dw.cpu().numpy(), odw) < RTOL
    assert rel_err(db.cpu().numpy(), odb) < RTOL
    assert rel_err(dx.cpu().numpy(), odx) < RTOL


@pytest.mark.parametrize("cudnn", [False, True])
def test_whole_path_bf16x3_lstm(cudnn):
    """The benchmarked arithmetic (bf16x3 on tcgen05) meets the same 1e-3 bar as the fp32 SIMT path."""
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=256, num_layers_rnn=2, num_units_rnn=64,
                      rnn_cell="lstm", cudnn=cudnn, dense_dropout_rate=0.0, compute="bf16x3")
    _whole_path(cfg, B=8, T=64, L=10, ragged=True)


@pytest.mark.parametrize("use_len", [True, False])
@pytest.mark.parametrize("T,B,nin,H", [(23, 5, 24, 64), (40, 32, 64, 128), (7, 1, 16, 192), (19, 70, 32, 128)])
def test_lstm_persistent_tcgen05_layer_vs_oracle(use_len, T, B, nin, H):
    """The persistent cooperative LSTM kernels (lstm_tc.cu): all T steps in one launch, weights and
    state as bf16-split tcgen05 operands, per-direction step barrier across CTAs."""
    rng = np.random.default_rng(T * 7 + B)
    x = rng.standard_normal((T, B, nin)).astype(np.float32)
    sl = np.maximum(1, T - (3 * np.arange(B)) % T).astype(np.int32)       # ragged; B=70 runs as 3 batch slices
    wx = (rng.standard_normal((nin, 8 * H)) * 0.2).astype(np.float32)
    wh = (rng.standard_normal((2, H, 4 * H)) * 0.1).astype(np.float32)
    bias = (rng.standard_normal(8 * H) * 0.1).astype(np.float32)
    dy = rng.standard_normal((T, B, 2 * H)).astype(np.float32)
    rb, _ = ops.birnn_sizes(T, B, nin, H, 2)
    reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
    y = torch.empty((T, B, 2 * H), device="cuda")
    X, SL, WX, WH, BI = dev(x), dev(sl), dev(wx), dev(wh), dev(bias)
    ops.birnn_fwd(X, SL, WX, WH, BI, y, reserve, 2, use_len, compute=BF16X3)
    oy, og, oc = ref.birnn_fwd(x.astype(np.float64), sl, wx, wh, bias, 2, use_len=use_len)
    assert rel_err(y.cpu().numpy(), oy) < 1e-4
    dx = torch.empty((T, B, nin), device="cuda")
    dwx, dwh, db = torch.empty_like(WX), torch.empty_like(WH), torch.empty_like(BI)
    ops.birnn_bwd(X, SL, WX, WH, y, reserve, dev(dy), dx, dwx, dwh, db, 2, use_len, compute=BF16X3)
    odx, odwx, odwh, odb = ref.birnn_bwd(x.astype(np.float64), sl, wx, wh, oy, og, oc, dy, 2, use_len=use_len)
    for got, want, name in [(dx, odx, "dx"), (dwx, odwx, "dwx"), (dwh, odwh, "dwh"), (db, odb, "dbias")]:
        assert rel_err(got.cpu().numpy(), want) < RTOL, name


# ----------------------------------------------- full-size properties (BASELINE cfg2: B=32, T=1000)
@pytest.fixture(scope="module")
def cfg2_model():
    from ctc_asr_b200.model import CTCModel
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=2048, num_layers_rnn=2, num_units_rnn=2048,
                      rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, compute="bf16x3")
    model = CTCModel(cfg, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(32, 1000, 160, seed=0)
    sl[5], sl[17] = 700, 431                     # two shorter utterances exercise the masking at scale
    x[5, 700:] = 0
    x[17, 431:] = 0
    return model, tuple(torch.from_numpy(a).cuda() for a in (x, sl, lab, ll))


def _loss_and_grad(model, batch, rows=slice(None), global_batch=None):
    x, sl, lab, ll = batch
    logits, _ = model.inference_fn(x[rows], sl[rows], training=False)
    loss = model.loss_fn(logits, sl[rows], (lab[rows], ll[rows]), global_batch=global_batch)
    model.backward()
    return float(loss), model.grad_flat.clone(), logits.clone()      # logits live in a buffer the next call reuses


def test_cfg2_full_size_gradient_is_the_derivative_of_the_loss(cfg2_model):
    """Size-independent property at the benchmarked size: the backward pass (LSTM cluster kernel,
    dgrad / wgrad GEMMs, CTC gradient) is the derivative of the forward pass —
    (L(p + e d) - L(p - e d)) / 2e == <grad, d> along a random direction d."""
    model, batch = cfg2_model
    loss0, grad, _ = _loss_and_grad(model, batch)
    assert np.isfinite(loss0)
    gen = torch.Generator(device="cuda").manual_seed(3)
    d = torch.randn(model.flat.shape, device="cuda", generator=gen)
    d *= grad.abs().mean() / (d.abs().mean() + 1e-30)
    d = grad + d                                   # mostly along the gradient so the signal is large
    d /= d.norm()
    analytic = float((grad.double() * d.double()).sum())
    p0 = model.flat.clone()
    eps = 2e-4                                     # |grad| ~ 2e4: stay in the linear regime, above fp32 loss resolution
    model.flat.copy_(p0 + eps * d)
    lp, _, _ = _loss_and_grad(model, batch)
    model.flat.copy_(p0 - eps * d)
    lm, _, _ = _loss_and_grad(model, batch)
    model.flat.copy_(p0)
    fd = (lp - lm) / (2 * eps)
    print("cfg2 directional derivative: finite difference %.4f, analytic %.4f, loss %.4f" % (fd, analytic, loss0))
    assert abs(fd - analytic) / abs(analytic) < 2e-2, (fd, analytic)


def test_cfg2_full_size_shards_add_up(cfg2_model):
    """Data-parallel property on one GPU: gradients of two half batches, each scaled by 1/global_batch,
    sum to the full-batch gradient (what the NCCL all-reduce relies on); per-utterance results do not
    depend on which other utterances share the batch."""
    model, batch = cfg2_model
    loss, grad, logits = _loss_and_grad(model, batch)
    per_utt = model.last_per_utterance_loss.clone()
    la, ga, logits_a = _loss_and_grad(model, batch, slice(0, 16), global_batch=32)
    pa = model.last_per_utterance_loss.clone()
    lb, gb, _ = _loss_and_grad(model, batch, slice(16, 32), global_batch=32)
    assert abs((la + lb) - loss) / abs(loss) < 1e-5
    err = (ga + gb - grad).abs().max().item() / grad.abs().max().item()
    print("cfg2 shard sum: loss err %.2e, grad err %.2e" % (abs((la + lb) - loss) / abs(loss), err))
    assert err < 1e-3, err
    assert torch.allclose(pa, per_utt[:16], rtol=1e-5)
    # same utterance, different batch composition: identical greedy transcript (integer work)
    ids_full, n_full = ops.greedy_decode(logits, batch[1])
    ids_half, n_half = ops.greedy_decode(logits_a, batch[1][:16].contiguous())
    assert (n_full[:16] == n_half).all()
    for b in range(16):
        assert torch.equal(ids_full[b, :n_full[b]], ids_half[b, :n_half[b]])


@pytest.mark.parametrize("compute,tol", [("bf16x3", 6e-5)])
def test_weight_gradient_contraction_at_cfg2_length(compute, tol):
    """Full-size accuracy of the longest contraction of the path: a weight gradient dW = x^T dz sums over every frame of
    the batch (K = T x B = 32,000 at cfg2; 256 output tiles: no split over K).  The tensor core truncates each accumulation
    into tensor memory, so ONE chain of K / 16 x 3 MMAs is off by 1.2e-4 at this length (linear in K; profiles/
    r2_accum_error.json, tools/accum_probe.py); the GEMM sums chunks of <= 8192 in separate accumulators and adds them in
    fp32 (gemm_tc.cu, chained accumulation): 3e-5 whatever K.  Against fp64 on the same device."""
    torch.manual_seed(11)
    K, M, N = 32000, 2048, 4096
    a = torch.randn(K, M, device="cuda")
    b = torch.randn(K, N, device="cuda")
    c = ops.gemm(a, b, ta=True, compute=_lib.COMPUTE_ID[compute])
    want = a.double().t() @ b.double()
    err = float((c.double() - want).abs().max() / want.abs().max())
    print("dW contraction K = %d (%s): max err / max = %.2e" % (K, compute, err))
    assert err < tol, err


def test_edit_distance_bit_exact():
    """tf.edit_distance(decoded, labels) (asr/model.py:338): integer work, bit-exact vs the oracle,
    up to the corpus' longest label (422, README.md:169)."""
    from ctc_asr_b200 import metrics
    rng = np.random.default_rng(1)
    B, Lh, Lt = 40, 430, 422
    hl, tl = rng.integers(0, Lh + 1, B).astype(np.int32), rng.integers(0, Lt + 1, B).astype(np.int32)
    hl[0], tl[0] = 0, 0
    hl[1], tl[1] = 7, 0
    hl[2], tl[2] = Lh, Lt
    hyp, truth = rng.integers(1, 28, (B, Lh)).astype(np.int32), rng.integers(1, 28, (B, Lt)).astype(np.int32)
    for norm in (False, True):
        got = metrics.edit_distance(dev(hyp), dev(hl), dev(truth), dev(tl), normalize=norm).cpu().numpy()
        want = ref.edit_distance(hyp, hl, truth, tl, normalize=norm)
        assert np.array_equal(got, want)
