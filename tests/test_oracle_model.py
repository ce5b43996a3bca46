"""CPU tests pinning the RNN / dense / whole-path oracle (oracle/oracle.c + oracle/model_ref.py).

Independent implementations used as pins (the reference has no tests):
  * oracle/torch_ref.py — per-timestep torch restatement of the TF graph + autograd
  * torch.nn.LSTM / torch.nn.RNN with pack_padded_sequence — library RNN with the TF fused
    kernel mapped to torch's (W_ih, W_hh) layout
"""
import numpy as np
import pytest
import torch

from ctc_asr_b200.params import ModelConfig
from ctc_asr_b200 import synthetic
from oracle import model_ref, ref, torch_ref


def _small_cfg(cell, cudnn, layers=2):
    return ModelConfig(used_model="ds1", num_units_dense=24, num_units_rnn=16, num_layers_rnn=layers, rnn_cell=cell,
                       num_layers_dense=2, num_features=10, cudnn=cudnn, dense_dropout_rate=0.0)


def _batch(cfg, B=3, T=12, seed=0, ragged=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, cfg.num_features))
    sl = np.array([T, T - 3, T - 7][:B], np.int32) if ragged else np.full(B, T, np.int32)
    for b in range(B):
        x[b, sl[b]:] = 0
    lab, ll = synthetic.make_labels(rng, B, np.array([4, 3, 2][:B]), sl)
    return x, sl, lab, ll


@pytest.mark.parametrize("cell", ["rnn_tanh", "rnn_relu", "lstm", "gru"])
@pytest.mark.parametrize("cudnn", [False, True])
def test_whole_path_vs_torch_autograd(cell, cudnn):
    cfg = _small_cfg(cell, cudnn)
    params = synthetic.init_params(cfg, seed=1, dtype=np.float64)
    for k in params:                                    # non-zero biases exercise every term
        if k.endswith("bias"):
            params[k] = np.random.default_rng(5).standard_normal(params[k].shape) * 0.1
    x, sl, lab, ll = _batch(cfg)
    loss, grads, logits, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll)
    p = torch_ref.params_to_torch(params, torch.float64)
    tloss, tgrads, tlogits = torch_ref.train_step_grads(
        cfg, p, torch.tensor(x), torch.tensor(sl), torch.tensor(lab), torch.tensor(ll))
    np.testing.assert_allclose(logits, tlogits.numpy(), atol=1e-10)
    np.testing.assert_allclose(loss, float(tloss), rtol=1e-10)
    for k in params:
        np.testing.assert_allclose(grads[k], tgrads[k].numpy(), atol=1e-9, err_msg=k)


def _tf_to_torch_lstm(wx, wh, bias, H, forget_bias):
    """TF gate order (i, j, f, o) -> torch (i, f, g, o); forget_bias folded into b_ih."""
    def perm(m):                                        # m [..., 4H] in TF order
        i, j, f, o = np.split(m, 4, axis=-1)
        return np.concatenate([i, f, j, o], -1)
    b = perm(bias.copy())
    b[H:2 * H] += forget_bias
    return perm(wx).T.copy(), perm(wh).T.copy(), b


@pytest.mark.parametrize("use_len", [True, False])
def test_lstm_layer_vs_torch_nn_lstm(use_len):
    rng = np.random.default_rng(2)
    T, B, nin, H = 9, 4, 6, 5
    x = rng.standard_normal((T, B, nin))
    sl = np.array([9, 7, 4, 1], np.int32)
    wx = rng.standard_normal((nin, 8 * H)) * 0.4
    wh = rng.standard_normal((2, H, 4 * H)) * 0.4
    bias = rng.standard_normal(8 * H) * 0.2
    y, _, _ = ref.birnn_fwd(x, sl, wx, wh, bias, cell=2, use_len=use_len, forget_bias=1.0)
    lstm = torch.nn.LSTM(nin, H, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in enumerate(["", "_reverse"]):
            w_ih, w_hh, b = _tf_to_torch_lstm(wx[:, d * 4 * H:(d + 1) * 4 * H], wh[d],
                                              bias[d * 4 * H:(d + 1) * 4 * H], H, 1.0)
            getattr(lstm, "weight_ih_l0" + suf).copy_(torch.tensor(w_ih))
            getattr(lstm, "weight_hh_l0" + suf).copy_(torch.tensor(w_hh))
            getattr(lstm, "bias_ih_l0" + suf).copy_(torch.tensor(b))
            getattr(lstm, "bias_hh_l0" + suf).zero_()
    xt = torch.tensor(x)
    if use_len:
        packed = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(sl, dtype=torch.long))
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], total_length=T)
    else:
        out, _ = lstm(xt)
    np.testing.assert_allclose(y, out.detach().numpy(), atol=1e-12)


def test_tanh_layer_vs_torch_nn_rnn():
    rng = np.random.default_rng(4)
    T, B, nin, H = 8, 3, 5, 7
    x = rng.standard_normal((T, B, nin))
    sl = np.array([8, 5, 2], np.int32)
    wx = rng.standard_normal((nin, 2 * H)) * 0.5
    wh = rng.standard_normal((2, H, H)) * 0.5
    bias = rng.standard_normal(2 * H) * 0.2
    y, _, _ = ref.birnn_fwd(x, sl, wx, wh, bias, cell=0, use_len=True)
    rnn = torch.nn.RNN(nin, H, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in enumerate(["", "_reverse"]):
            getattr(rnn, "weight_ih_l0" + suf).copy_(torch.tensor(wx[:, d * H:(d + 1) * H].T.copy()))
            getattr(rnn, "weight_hh_l0" + suf).copy_(torch.tensor(wh[d].T.copy()))
            getattr(rnn, "bias_ih_l0" + suf).copy_(torch.tensor(bias[d * H:(d + 1) * H]))
            getattr(rnn, "bias_hh_l0" + suf).zero_()
    packed = torch.nn.utils.rnn.pack_padded_sequence(torch.tensor(x), torch.tensor(sl, dtype=torch.long))
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(rnn(packed)[0], total_length=T)
    np.testing.assert_allclose(y, out.detach().numpy(), atol=1e-12)
    assert not y[5:, 1].any() and not y[2:, 2].any()       # dynamic_rnn: zeros past the length


def test_dense_dropout_mask_consistency():
    """Forward keep-mask and backward mask come from the same counter hash."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((40, 7))
    w = rng.standard_normal((7, 11))
    b = rng.standard_normal(11) * 0.1
    y = ref.dense_fwd(x, w, b, act=1, cutoff=1.5, drop_rate=0.3, seed=42)
    y0 = ref.dense_fwd(x, w, b, act=1, cutoff=1.5, drop_rate=0.0)
    kept = y != 0
    assert 0.5 < kept[y0 > 0].mean() < 0.9
    np.testing.assert_allclose(y[kept], y0[kept] / 0.7, rtol=1e-6)
    dy = rng.standard_normal(y.shape)
    dx, dw, db = ref.dense_bwd(x, w, y, dy, act=1, cutoff=1.5, drop_rate=0.3, seed=42)
    dz = np.where(kept & (y0 > 0) & (y0 < 1.5), dy / 0.7, 0.0)
    np.testing.assert_allclose(dw, x.T @ dz, atol=1e-12)
    np.testing.assert_allclose(db, dz.sum(0), atol=1e-12)
    np.testing.assert_allclose(dx, dz @ w.T, atol=1e-12)


def test_adam_tf1_formula():
    rng = np.random.default_rng(1)
    p, g = rng.standard_normal(100), rng.standard_normal(100)
    m, v = np.zeros(100), np.zeros(100)
    p0 = p.copy()
    ref.adam(p, m, v, g, step=1, lr=1e-3)
    # first step of Adam moves every weight by ~lr against the sign of its gradient
    np.testing.assert_allclose(p0 - p, 1e-3 * np.sign(g), rtol=1e-2)
    tp = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1e-3, eps=1e-8)
    p2, m2, v2 = p0.copy(), np.zeros(100), np.zeros(100)
    for step in range(1, 4):
        tp.grad = torch.tensor(g * step)
        opt.step()
        ref.adam(p2, m2, v2, g * step, step=step, lr=1e-3)
    # TF1's epsilon placement differs from torch's by O(eps): agreement to ~1e-6 relative
    np.testing.assert_allclose(p2, tp.detach().numpy(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("use_len", [True, False])
def test_gru_layer_vs_torch_nn_gru(use_len):
    """cuDNN-formulation GRU (what the reference's CudnnGRU computes, asr/model.py:197) == torch.nn.GRU."""
    rng = np.random.default_rng(0)
    T, B, nin, H = 7, 3, 5, 4
    x = rng.standard_normal((T, B, nin))
    sl = np.array([7, 5, 2], np.int32)
    wx = rng.standard_normal((nin, 6 * H)) * 0.5
    wh = rng.standard_normal((2, H, 3 * H)) * 0.5
    bias = rng.standard_normal(8 * H) * 0.3
    y, g, q = ref.birnn_fwd(x, sl, wx, wh, bias, 3, use_len=use_len)
    gru = torch.nn.GRU(nin, H, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in enumerate(["", "_reverse"]):
            getattr(gru, "weight_ih_l0" + suf).copy_(torch.tensor(wx[:, d * 3 * H:(d + 1) * 3 * H].T.copy()))
            getattr(gru, "weight_hh_l0" + suf).copy_(torch.tensor(wh[d].T.copy()))
            getattr(gru, "bias_ih_l0" + suf).copy_(torch.tensor(bias[d * 3 * H:(d + 1) * 3 * H]))
            bhh = np.zeros(3 * H)
            bhh[2 * H:] = bias[6 * H + d * H:6 * H + (d + 1) * H]
            getattr(gru, "bias_hh_l0" + suf).copy_(torch.tensor(bhh))
    xt = torch.tensor(x, requires_grad=True)
    if use_len:
        packed = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(sl, dtype=torch.long))
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(gru(packed)[0], total_length=T)
    else:
        out, _ = gru(xt)
    np.testing.assert_allclose(y, out.detach().numpy(), atol=1e-12)
    dy = rng.standard_normal(y.shape)
    out.backward(torch.tensor(dy))
    dx, dwx, dwh, db = ref.birnn_bwd(x, sl, wx, wh, y, g, q, dy, 3, use_len=use_len)
    np.testing.assert_allclose(dx, xt.grad.numpy(), atol=1e-12)
    np.testing.assert_allclose(dwh[0], gru.weight_hh_l0.grad.numpy().T, atol=1e-12)
    np.testing.assert_allclose(dwx[:, 3 * H:], gru.weight_ih_l0_reverse.grad.numpy().T, atol=1e-12)
    np.testing.assert_allclose(db[6 * H:7 * H], gru.bias_hh_l0.grad.numpy()[2 * H:], atol=1e-12)


# ---- ds2 conv front-end (asr/util/tf_contrib.py:64-146) -------------------------------------------
@pytest.mark.parametrize("dims", [(13, 2, 10, 1, 4, 11, 5, 2, 2), (9, 3, 8, 3, 5, 3, 5, 1, 2), (7, 1, 7, 2, 3, 11, 21, 1, 2),
                                  (20, 2, 80, 1, 4, 11, 41, 2, 2)])
def test_conv2d_same_vs_torch_conv2d(dims):
    """The C conv layer (fwd + bwd) against torch F.conv2d + autograd with TF 'SAME' padding made
    explicit (odd pad unit after), incl. the reference's (11,41)/(2,2) and (11,21)/(1,2) geometries."""
    import torch.nn.functional as Fn
    T, B, F, C, N, kt, kf, st, sf = dims
    rng = np.random.default_rng(0)
    x = rng.standard_normal((T, B, F, C)); w = rng.standard_normal((kt, kf, C, N)) * 0.3; b = rng.standard_normal(N) * 0.1
    y = ref.conv2d_fwd(x, w, b, (st, sf), cutoff=1.0)
    assert 0.05 < (y >= 1.0).mean() < 0.95 and (y <= 0).any()          # both kinks are exercised
    dy = rng.standard_normal(y.shape)
    dx, dw, db = ref.conv2d_bwd(x, w, y, dy, (st, sf), cutoff=1.0)
    xt = torch.tensor(x).permute(1, 3, 0, 2).requires_grad_(True)       # [B,C,T,F]
    wt = torch.tensor(w).requires_grad_(True)
    bt = torch.tensor(b).requires_grad_(True)
    yt = torch.clamp(torch.relu(torch_ref._conv_same(xt, wt, bt, (st, sf))), max=1.0).permute(2, 0, 3, 1)
    assert tuple(yt.shape) == y.shape == (-(-T // st), B, -(-F // sf), N)
    np.testing.assert_allclose(y, yt.detach().numpy(), atol=1e-12)
    (yt * torch.tensor(dy)).sum().backward()
    np.testing.assert_allclose(dx, xt.grad.permute(2, 0, 3, 1).numpy(), atol=1e-12)
    np.testing.assert_allclose(dw, wt.grad.numpy(), atol=1e-11)
    np.testing.assert_allclose(db, bt.grad.numpy(), atol=1e-11)


def test_ds2_whole_path_vs_torch_autograd():
    cfg = ModelConfig(used_model="ds2", conv_filters=(3, 4, 64), num_units_dense=24, num_units_rnn=16, num_layers_rnn=2,
                      rnn_cell="lstm", num_features=12, cudnn=False, dense_dropout_rate=0.0)
    params = synthetic.init_params(cfg, seed=1, dtype=np.float64)
    assert params["conv/conv2d/kernel"].shape == (11, 41, 1, 3) and params["conv/conv2d_2/kernel"].shape == (11, 21, 4, 64)
    assert params["rnn/l0/wx"].shape[0] == 2 * 64                      # 12 features -> 6 -> 3 -> 2 bins x 64 filters
    for k in params:
        if k.endswith("bias"):
            params[k] = np.random.default_rng(5).standard_normal(params[k].shape) * 0.1
    x, sl, lab, ll = _batch(cfg, B=3, T=15)
    loss, grads, logits, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll)
    assert logits.shape == (8, 3, 29)                                  # ceil(15 / 2) frames reach the RNN and CTC
    p = torch_ref.params_to_torch(params, torch.float64)
    tloss, tgrads, tlogits = torch_ref.train_step_grads(
        cfg, p, torch.tensor(x), torch.tensor(sl), torch.tensor(lab), torch.tensor(ll))
    np.testing.assert_allclose(logits, tlogits.numpy(), atol=1e-10)
    np.testing.assert_allclose(loss, float(tloss), rtol=1e-10)
    for k in params:
        np.testing.assert_allclose(grads[k], tgrads[k].numpy(), atol=1e-9, err_msg=k)


def test_operand_rounding_mode_is_bf16_round_to_nearest_even():
    """oracle operand_rounding (the arithmetic compute='bf16' is compared against): both operands of every product
    rounded exactly like torch's float32 -> bfloat16 conversion, fp64 accumulation; off again afterwards."""
    import torch
    rng = np.random.default_rng(3)
    x, w, b = rng.standard_normal((9, 31)), rng.standard_normal((31, 7)), rng.standard_normal(7)
    bf = lambda a: torch.tensor(a, dtype=torch.float32).bfloat16().double().numpy()
    exact = ref.dense_fwd(x, w, b, act=0)
    with ref.operand_rounding(True):
        y = ref.dense_fwd(x, w, b, act=0)
        dx, dw, db = ref.dense_bwd(x, w, y, exact, act=0)
    assert np.abs(y - (bf(x) @ bf(w) + b)).max() < 1e-12
    assert np.abs(dw - bf(x).T @ bf(exact)).max() < 1e-12 and np.abs(dx - bf(exact) @ bf(w).T).max() < 1e-12
    assert np.abs(ref.dense_fwd(x, w, b, act=0) - exact).max() == 0
    assert 1e-4 < np.abs(y - exact).max() < 1e-1
    # a recurrent layer: h_{t-1} is rounded where it enters the product, not where it is stored
    T, B, nin, H = 5, 2, 6, 4
    xs = rng.standard_normal((T, B, nin)); sl = np.array([5, 3], np.int32)
    wx, wh, bias = rng.standard_normal((nin, 2 * H)) * 0.3, rng.standard_normal((2, H, H)) * 0.3, rng.standard_normal(2 * H) * 0.1
    with ref.operand_rounding(True):
        y, _, _ = ref.birnn_fwd(xs, sl, wx, wh, bias, 0, use_len=True)
    h = np.zeros(H)
    for t in range(5):                       # forward direction of utterance 0 by hand
        h = np.tanh(bf(xs[t, 0]) @ bf(wx[:, :H]) + bf(h) @ bf(wh[0]) + bias[:H])
        assert np.abs(y[t, 0, :H] - h).max() < 1e-12


def test_numpy_keep_mask_is_the_c_oracles():
    """ref.dropout (numpy) uses the same counter hash as oracle_dense_fwd (and the CUDA epilogues): identical keep-mask and
    scaling through an identity layer; about `rate` of the entries dropped; applying it twice to a gradient is what the
    backward pass of the RNN / conv dropout does."""
    rng = np.random.default_rng(0)
    M, N = 37, 24
    x = rng.standard_normal((M, N)) + 3.0
    y = ref.dense_fwd(x, np.eye(N), np.zeros(N), act=0, drop_rate=0.3, seed=12345)
    assert np.array_equal(y, ref.dropout(x, 0.3, 12345))
    assert 0.2 < (y == 0).mean() < 0.4
    assert ref.dropout(x, 0.0, 1) is x
    # the padded pitch enters the index: a [M, 8] tensor stored with pitch 64 is columns 0..7 of the [M, 64] mask
    wide = ref.dropout(np.ones((M, 64)), 0.5, 9)
    assert np.array_equal(ref.dropout(np.ones((M, 8)), 0.5, 9, pitch=64), wide[:, :8])


def test_whole_path_oracle_dropout_gradients_by_finite_differences():
    """Oracle with every RNN dropout on (fixed masks): directional finite difference of the loss against its gradients."""
    from types import SimpleNamespace
    cfg = SimpleNamespace(used_model="ds1", num_layers_dense=1, num_units_dense=8, num_layers_rnn=2, num_units_rnn=6,
                          rnn_cell="rnn_tanh", cudnn=False, dense_dropout_rate=0.2, rnn_dropout_rate=0.3, relu_cutoff=20.0,
                          forget_bias=1.0, num_classes=5, num_features=4, conv_filters=())
    rng = np.random.default_rng(3)
    shapes = {"dense/dense/kernel": (4, 8), "dense/dense/bias": (8,), "rnn/l0/wx": (8, 12), "rnn/l0/wh": (2, 6, 6),
              "rnn/l0/bias": (12,), "rnn/l1/wx": (12, 12), "rnn/l1/wh": (2, 6, 6), "rnn/l1/bias": (12,),
              "dense4/dense/kernel": (12, 8), "dense4/dense/bias": (8,), "logits/dense/kernel": (8, 5), "logits/dense/bias": (5,)}
    params = {k: rng.standard_normal(v) * 0.4 for k, v in shapes.items()}
    x = rng.standard_normal((2, 9, 4))
    sl, lab, ll = np.array([9, 7], np.int32), np.array([[1, 2], [3, 0]], np.int32), np.array([2, 1], np.int32)
    loss, grads, _, _ = model_ref.loss_and_grads(cfg, params, x, sl, lab, ll, training=True, seed=5)
    d = {k: rng.standard_normal(v.shape) for k, v in params.items()}
    eps = 1e-6
    lp = model_ref.loss_and_grads(cfg, {k: v + eps * d[k] for k, v in params.items()}, x, sl, lab, ll, training=True, seed=5)[0]
    lm = model_ref.loss_and_grads(cfg, {k: v - eps * d[k] for k, v in params.items()}, x, sl, lab, ll, training=True, seed=5)[0]
    fd = (lp - lm) / (2 * eps)
    an = sum(float((grads[k] * d[k]).sum()) for k in params)
    assert abs(fd - an) < 1e-6 * max(1.0, abs(an)), (fd, an)
