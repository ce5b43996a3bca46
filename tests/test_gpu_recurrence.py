"""Persistent tcgen05 recurrences against the C oracle, through the C-ABI (ctcasr_birnn_fwd / _bwd).

One launch per layer and pass for every cell of the reference's menu (asr/params.py:48-50,
asr/util/tf_contrib.py:189): rnn_tanh / rnn_relu (rec_tc.cu, weights resident on chip), lstm / gru
(lstm_tc.cu).  Includes one full-width layer (H = 2048) per cell, at a sequence length the scalar oracle
finishes in seconds.
"""
import numpy as np
import pytest
import torch

from ctc_asr_b200 import _lib, ops
from oracle import ref

pytestmark = pytest.mark.gpu
BF16X3 = _lib.COMPUTE_BF16X3
RTOL = 1e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_err(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def _layer(cell, T, B, nin, H, use_len, seed, compute=BF16X3, wscale=None, tol=RTOL, ytol=1e-4):
    rng = np.random.default_rng(seed)
    cid = ref.CELL_IDS[cell]
    G = ref.NUM_GATES[cid]
    x = rng.standard_normal((T, B, nin)).astype(np.float32)
    sl = np.maximum(1, T - (3 * np.arange(B)) % T).astype(np.int32)
    wx = (rng.standard_normal((nin, 2 * G * H)) * 0.2).astype(np.float32)
    wscale = 1.0 / np.sqrt(H) if wscale is None else wscale
    wh = (rng.standard_normal((2, H, G * H)) * wscale).astype(np.float32)
    bias = (rng.standard_normal(2 * G * H + (2 * H if cell == "gru" else 0)) * 0.1).astype(np.float32)
    dy = rng.standard_normal((T, B, 2 * H)).astype(np.float32)
    rb, _ = ops.birnn_sizes(T, B, nin, H, cid)
    reserve = torch.empty(rb, dtype=torch.uint8, device="cuda")
    y = torch.empty((T, B, 2 * H), device="cuda")
    X, SL, WX, WH, BI = dev(x), dev(sl), dev(wx), dev(wh), dev(bias)
    lib = _lib.load()
    l0 = lib.ctcasr_launch_count()
    ops.birnn_fwd(X, SL, WX, WH, BI, y, reserve, cid, use_len, forget_bias=1.0, compute=compute)
    fwd_launches = lib.ctcasr_launch_count() - l0
    oy, og, oc = ref.birnn_fwd(x.astype(np.float64), sl, wx, wh, bias, cid, use_len=use_len)
    assert rel_err(y.cpu().numpy(), oy) < ytol
    dx = torch.empty((T, B, nin), device="cuda")
    dwx, dwh, db = torch.empty_like(WX), torch.empty_like(WH), torch.empty_like(BI)
    l0 = lib.ctcasr_launch_count()
    ops.birnn_bwd(X, SL, WX, WH, y, reserve, dev(dy), dx, dwx, dwh, db, cid, use_len, compute=compute)
    bwd_launches = lib.ctcasr_launch_count() - l0
    odx, odwx, odwh, odb = ref.birnn_bwd(x.astype(np.float64), sl, wx, wh, oy, og, oc, dy, cid, use_len=use_len)
    for got, want, name in [(dx, odx, "dx"), (dwx, odwx, "dwx"), (dwh, odwh, "dwh"), (db, odb, "dbias")]:
        assert rel_err(got.cpu().numpy(), want) < tol, name
    return fwd_launches, bwd_launches


@pytest.mark.parametrize("cell", ["rnn_tanh", "rnn_relu"])
@pytest.mark.parametrize("use_len", [True, False])
@pytest.mark.parametrize("T,B,nin,H", [(23, 5, 24, 256), (40, 32, 64, 512), (19, 70, 32, 256), (7, 1, 16, 768)])
def test_one_gate_persistent_layer_vs_oracle(cell, use_len, T, B, nin, H):
    """rec_tc.cu: 4-CTA clusters split K, weights resident in tensor / shared memory, one launch per pass.
    B = 70 runs as three 32-row batch slices; H = 768 keeps 3 k-blocks per CTA; ragged lengths."""
    fl, bl = _layer(cell, T, B, nin, H, use_len, seed=T * 11 + B)
    # no per-frame launches: input GEMM (+ operand splits) + weight pack + one recurrence launch per 32-row slice
    assert fl < 12 + 2 * ((B + 31) // 32) and bl < 40, (fl, bl)


@pytest.mark.parametrize("cell,B", [("rnn_tanh", 32), ("rnn_relu", 32), ("lstm", 32), ("gru", 32), ("lstm", 64), ("gru", 64), ("lstm", 70)])
def test_full_width_layer_vs_oracle(cell, B):
    """One H = 2048 layer per cell (the benchmarked width: 128 CTAs, every k-block path, L2 policies)
    against the fp64 oracle.  B = 64 / 70: the 64-row batch tile of the gated kernels (BASELINE cfg4's batch size; one
    weight stream per time step for all 64 rows), 70 = a full tile plus a 6-row one."""
    _layer(cell, T=6, B=B, nin=64, H=2048, use_len=True, seed=5)


@pytest.mark.parametrize("use_len", [True, False])
@pytest.mark.parametrize("T,B,nin,H", [(23, 5, 24, 256), (12, 32, 64, 512), (9, 40, 32, 256)])
def test_gru_persistent_layer_vs_oracle(use_len, T, B, nin, H):
    """GRU (cuDNN formulation, asr/model.py:197) on the persistent kernels of lstm_tc.cu: forward with 3 gate
    groups per CTA tile, backward with the 4-CTA clusters splitting K = 3H."""
    fl, bl = _layer("gru", T, B, nin, H, use_len, seed=T * 13 + B)
    assert fl < 14 + 2 * ((B + 31) // 32) and bl < 48, (fl, bl)


@pytest.mark.parametrize("cell,T,B,nin,H", [("lstm", 40, 32, 64, 128), ("lstm", 19, 70, 32, 256), ("lstm", 7, 3, 16, 192),
                                            ("gru", 12, 32, 64, 256), ("lstm", 6, 32, 64, 2048)])
def test_single_piece_bf16_recurrence(cell, T, B, nin, H):
    """compute = 'bf16' (BASELINE cfg3): operands rounded to bf16, ONE product per k-step, most of the weight
    slice resident in tensor memory.  Against the exact oracle the difference is the bf16 operand rounding
    (2^-9 per operand); the bf16-rounded-operand oracle comparison at 1e-3 is in test_gpu_bf16_oracle.py."""
    _layer(cell, T, B, nin, H, True, seed=T + B, compute=_lib.COMPUTE_BF16, tol=5e-2, ytol=2e-2)
