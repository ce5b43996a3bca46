"""GPU parity of the ds2 convolutional front-end (asr/util/tf_contrib.py:64-146,
asr/model.py:154-161): the conv layer through the C-ABI against the CPU oracle, the whole ds2 path
(conv x3 -> BiRNN -> dense -> CTC, loss + every gradient), and the reference's default geometry at
full size through size-independent properties.  Tolerance: 1e-3 relative, as everywhere."""
import numpy as np
import pytest
import torch

from ctc_asr_b200 import _lib, ops, synthetic
from ctc_asr_b200.params import ModelConfig, conv_plan, same_out
from oracle import ref

from test_gpu_parity import RTOL, _whole_path, dev, rel_err

pytestmark = pytest.mark.gpu


def _pad_kernel(w, N):
    kt, kf, C, filt = w.shape
    K = kt * kf * C
    Kp = (K + 7) // 8 * 8
    out = np.zeros((Kp, N), np.float32)
    out[:K, :filt] = w.reshape(K, filt)
    return out


@pytest.mark.parametrize("compute", ["fp32", "bf16x3"])
@pytest.mark.parametrize("dims", [
    # T, B, F, C, x_pitch, filters, kt, kf, st, sf
    (40, 4, 80, 1, 1, 32, 11, 41, 2, 2),        # the reference's first layer (on a short clip)
    (21, 3, 40, 32, 64, 32, 11, 21, 1, 2),      # second layer: 32 real channels in a 64-float pitch
    (20, 2, 20, 32, 64, 96, 11, 21, 1, 2),      # third layer: 96 filters = its own pitch
    (13, 2, 9, 3, 5, 4, 3, 5, 2, 3),            # odd everything: K = 45 -> Kp = 48, partial tiles
    (19, 5, 13, 32, 32, 64, 5, 7, 2, 3),        # implicit-GEMM path (C = 32) with strides in time and frequency, partial boxes
    (16, 3, 11, 64, 64, 128, 3, 5, 1, 1),       # two 32-channel chunks per tap, 128 filters (two B boxes)
])
def test_conv2d_layer_vs_oracle(dims, compute):
    T, B, F, C, xp, filt, kt, kf, st, sf = dims
    C_ID = _lib.COMPUTE_ID[compute]
    N = max(64, (filt + 7) // 8 * 8)
    rng = np.random.default_rng(T + F)
    x = rng.standard_normal((T, B, F, C)).astype(np.float32)
    w = (rng.standard_normal((kt, kf, C, filt)) * (1.5 / np.sqrt(kt * kf * C))).astype(np.float32)
    b = (rng.standard_normal(filt) * 0.1).astype(np.float32)
    xpad = np.full((T, B, F, xp), 7.0, np.float32)                     # pad channels must be ignored
    xpad[..., :C] = x
    bp = np.zeros(N, np.float32); bp[:filt] = b
    To, Fo = same_out(T, kt, st)[0], same_out(F, kf, sf)[0]
    dx_d, w_d = dev(xpad), dev(_pad_kernel(w, N))
    y_d = torch.full((To * B * Fo, N), float("nan"), device="cuda")
    ops.conv2d_fwd(dx_d, xp, w_d, dev(bp), y_d, T, B, F, C, kt, kf, st, sf, act=1, cutoff=1.0, compute=C_ID)
    y = y_d.cpu().numpy().reshape(To, B, Fo, N)
    want = ref.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (st, sf), cutoff=1.0)
    assert 0.02 < (want >= 1.0).mean() and (want <= 0).any()
    assert rel_err(y[..., :filt], want) < RTOL
    assert (y[..., filt:] == 0).all()                                  # pad channels are exact zeros
    # backward on the ORACLE's mask source: feed the GPU its own y (masks agree away from the kinks)
    dy = rng.standard_normal(want.shape).astype(np.float32)
    near_kink = (np.abs(want) < 1e-4) | (np.abs(want - 1.0) < 1e-4)
    dy[near_kink] = 0
    dyp = np.zeros((To, B, Fo, N), np.float32); dyp[..., :filt] = dy
    dy_d = dev(dyp)
    gx_d = torch.full((T * B * F, xp), float("nan"), device="cuda")
    gw_d = torch.full_like(w_d, float("nan"))
    gb_d = torch.full((N,), float("nan"), device="cuda")
    ops.conv2d_bwd(dx_d, xp, w_d, y_d, dy_d, gx_d, gw_d, gb_d, T, B, F, C, kt, kf, st, sf, act=1, cutoff=1.0, compute=C_ID)
    torch.cuda.synchronize()
    odx, odw, odb = ref.conv2d_bwd(x.astype(np.float64), w.astype(np.float64), want, dy.astype(np.float64), (st, sf), cutoff=1.0)
    gx = gx_d.cpu().numpy().reshape(T, B, F, xp)
    K = kt * kf * C
    gw = gw_d.cpu().numpy()
    assert rel_err(gx[..., :C], odx) < RTOL
    assert (gx[..., C:] == 0).all()
    assert rel_err(gw[:K, :filt].reshape(kt, kf, C, filt), odw) < RTOL
    assert (gw[K:] == 0).all() and (gw[:, filt:] == 0).all()           # the padding never learns
    assert rel_err(gb_d.cpu().numpy()[:filt], odb) < RTOL
    # dx = None (first layer of the model): dw, db unchanged
    dy_d2 = dev(dyp)
    gw2, gb2 = torch.empty_like(gw_d), torch.empty_like(gb_d)
    ops.conv2d_bwd(dx_d, xp, w_d, y_d, dy_d2, None, gw2, gb2, T, B, F, C, kt, kf, st, sf, act=1, cutoff=1.0, compute=C_ID)
    assert torch.equal(gw2, gw_d) and torch.equal(gb2, gb_d)           # deterministic


@pytest.mark.parametrize("compute", ["fp32", "bf16x3"])
def test_ds2_whole_path_small(compute):
    """conv x3 (the reference's kernel sizes and strides) -> 2 BiLSTM -> dense4 -> logits -> CTC, every
    gradient tensor against the fp64 oracle; CTC and the RNN see ceil(T/2) frames for every utterance."""
    cfg = ModelConfig(used_model="ds2", conv_filters=(8, 8, 64), num_units_dense=64, num_layers_rnn=2,
                      num_units_rnn=64 if compute == "bf16x3" else 32, rnn_cell="lstm", num_features=20,
                      cudnn=False, dense_dropout_rate=0.0, compute=compute)
    model = _whole_path(cfg, B=4, T=61, L=6, ragged=True)
    assert model._saved["T"] == 31


def test_implicit_conv_bf16_mode_and_dropout():
    """compute='bf16' (one product per operand pair) through the implicit-GEMM kernel against the oracle on operands
    rounded to bfloat16, with the conv dropout mask in its epilogue (index over the padded pitch)."""
    T, B, F, C, xp, filt, kt, kf, st, sf = 24, 4, 16, 32, 64, 32, 11, 21, 1, 2
    N = 64
    rng = np.random.default_rng(5)
    x = rng.standard_normal((T, B, F, C)).astype(np.float32)
    w = (rng.standard_normal((kt, kf, C, filt)) * (1.5 / np.sqrt(kt * kf * C))).astype(np.float32)
    b = (rng.standard_normal(filt) * 0.1).astype(np.float32)
    xpad = np.zeros((T, B, F, xp), np.float32); xpad[..., :C] = x
    bp = np.zeros(N, np.float32); bp[:filt] = b
    To, Fo = same_out(T, kt, st)[0], same_out(F, kf, sf)[0]
    y_d = torch.full((To * B * Fo, N), float("nan"), device="cuda")
    ops.conv2d_fwd(dev(xpad), xp, dev(_pad_kernel(w, N)), dev(bp), y_d, T, B, F, C, kt, kf, st, sf, act=1, cutoff=1.0,
                   drop_rate=0.2, seed=99, compute=_lib.COMPUTE_ID["bf16"])
    r = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float64).numpy()
    want = ref.conv2d_fwd(r(x), r(w), b.astype(np.float64), (st, sf), cutoff=1.0)
    want = ref.dropout(want, 0.2, 99, pitch=N)
    y = y_d.cpu().numpy().reshape(To, B, Fo, N)
    away = np.abs(np.abs(ref.conv2d_fwd(r(x), r(w), b.astype(np.float64), (st, sf), cutoff=1e9)) - 0.5) < 0.4999   # off both kinks
    assert np.abs(y[..., :filt] - want)[away].max() < 1e-4
    assert np.array_equal((y[..., :filt] == 0)[away], (want == 0)[away])
    assert (y[..., filt:] == 0).all()


def test_ds2_whole_path_with_conv_and_rnn_dropout():
    """Every dropout of the ds2 path on: tf.layers.dropout after each conv layer (asr/util/tf_contrib.py:135; the keep-mask
    index runs over the padded channel pitch), the DropoutWrapper masks of the RNN stack, dense4's dropout."""
    cfg = ModelConfig(used_model="ds2", conv_filters=(8, 8, 64), num_units_dense=64, num_layers_rnn=2, num_units_rnn=32,
                      rnn_cell="rnn_tanh", num_features=20, cudnn=False, dense_dropout_rate=0.1, conv_dropout_rate=0.15,
                      rnn_dropout_rate=0.2, compute="fp32", random_seed=77)
    _whole_path(cfg, B=4, T=61, L=6, ragged=True, training=True)


@pytest.mark.parametrize("compute", ["bf16x3"])
def test_ds2_whole_path_through_the_implicit_conv_kernels(compute):
    """32 filters in the first two layers, as in the reference: the second and third conv layer take the implicit-GEMM
    forward / weight-gradient / input-gradient kernels (conv_tc.cu) and the fused backward prologue, with conv, RNN and dense
    dropout on; every gradient tensor against the fp64 oracle.
    Data seed 23: no conv pre-activation lies within 1.7e-5 (relative to the layer's largest output) of a ReLU kink.  With
    seed 0 one first-layer pre-activation is 8e-8 away from 0, the two sides clip it differently and that single flipped
    mask element moves the first layer's kernel gradient by 1 % (tools/r2_diag_conv.py shows it stage by stage; the
    32-filter first layer has 4x the elements of the (8, 8, 64) tests)."""
    cfg = ModelConfig(used_model="ds2", conv_filters=(32, 32, 64), num_units_dense=64, num_layers_rnn=1, num_units_rnn=64,
                      rnn_cell="lstm", num_features=20, cudnn=True, dense_dropout_rate=0.1, conv_dropout_rate=0.1,
                      rnn_dropout_rate=0.0, compute=compute, random_seed=5)
    _whole_path(cfg, B=4, T=61, L=6, ragged=True, training=True, seed=23)


def test_ds2_model_shapes_follow_the_reference():
    cfg = ModelConfig(used_model="ds2", num_layers_rnn=1, num_units_rnn=64, num_units_dense=64, compute="fp32")
    plan = conv_plan(cfg, 999)
    assert [(d["To"], d["Fo"], d["filters"]) for d in plan] == [(500, 40, 32), (500, 20, 32), (500, 10, 96)]
    from ctc_asr_b200.model import CTCModel
    model = CTCModel(cfg, seed=1)
    assert tuple(model.p["conv/conv2d/kernel"].shape) == (11, 41, 1, 32)
    assert tuple(model.p["conv/conv2d_2/kernel"].shape) == (11, 21, 32, 96)
    assert model.p["rnn/l0/wx"].shape[0] == 960                        # 10 * filters[-1], tf_contrib.py:138
    k = synthetic.init_params(cfg, seed=1)["conv/conv2d_1/kernel"]
    assert np.array_equal(model.params_numpy()["conv/conv2d_1/kernel"], k)
    assert float(model.ps["conv/conv2d_1/kernel"].abs().sum()) == pytest.approx(float(np.abs(k).sum()), rel=1e-5)


def test_ds2_full_size_properties():
    """The reference's default front-end at the benchmarked batch (B=32 x 10 s): the step runs, the
    gradient is the derivative of the loss along a random direction, and a second identical step is
    bit-identical (no atomics anywhere on the path)."""
    from ctc_asr_b200.model import CTCModel
    cfg = ModelConfig(used_model="ds2", num_layers_dense=3, num_units_dense=2048, num_layers_rnn=2, num_units_rnn=2048,
                      rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, compute="bf16x3")
    model = CTCModel(cfg, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(32, 1000, 160, seed=0)
    batch = tuple(torch.from_numpy(a).cuda() for a in (x, sl, lab, ll))

    def loss_and_grad():
        logits, s2 = model.inference_fn(batch[0], batch[1], training=False)
        assert logits.shape == (500, 32, 29) and int(s2.min()) == 500
        loss = model.loss_fn(logits, s2, (batch[2], batch[3]))
        model.backward()
        return float(loss), model.grad_flat.clone()

    loss0, grad = loss_and_grad()
    assert np.isfinite(loss0) and bool(torch.isfinite(grad).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss1, grad1 = loss_and_grad()
    e1.record()
    torch.cuda.synchronize()
    print("ds2 full size: loss %.4f, fwd+CTC+bwd %.1f ms" % (loss0, e0.elapsed_time(e1)))
    assert loss1 == loss0 and torch.equal(grad, grad1)
    gen = torch.Generator(device="cuda").manual_seed(3)
    d = torch.randn(model.flat.shape, device="cuda", generator=gen)
    d *= grad.abs().mean() / (d.abs().mean() + 1e-30)
    d = grad + d
    d *= (model.flat != 0) | (grad != 0)             # stay inside the parameter set (not its zero padding)
    d /= d.norm()
    analytic = float((grad.double() * d.double()).sum())
    p0 = model.flat.clone()
    eps = 2e-4
    model.flat.copy_(p0 + eps * d); lp, _ = loss_and_grad()
    model.flat.copy_(p0 - eps * d); lm, _ = loss_and_grad()
    model.flat.copy_(p0)
    fd = (lp - lm) / (2 * eps)
    print("ds2 directional derivative: finite difference %.4f, analytic %.4f" % (fd, analytic))
    assert abs(fd - analytic) / abs(analytic) < 2e-2, (fd, analytic)


@pytest.mark.parametrize("compute,tol", [("bf16x3", 2e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("M,N,K,ta,tb", [
    (456, 64, 64000, True, False),      # conv weight gradient: 4 output tiles, long contraction -> split-K
    (7392, 96, 20000, True, False),     # 58 tiles x 5 slices
    (1000, 64, 456, False, False),      # conv forward: instruction N = 64 of the 256-column tile
    (1000, 96, 7392, False, False),
    (1000, 456, 64, False, True),       # conv dgrad: K = 64
    (264, 328, 9000, False, True),      # last N tile narrower (72 -> 80 columns), K-major B
])
def test_gemm_tcgen05_narrow_n_and_split_k(M, N, K, ta, tb, compute, tol):
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(dev(a), dev(b), ta=ta, tb=tb, compute=_lib.COMPUTE_ID[compute])
    c2 = ops.gemm(dev(a), dev(b), ta=ta, tb=tb, compute=_lib.COMPUTE_ID[compute])
    assert torch.equal(c, c2)                                          # slices are added in a fixed order
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    assert rel_err(c.cpu().numpy(), want) < tol
