"""ctypes/numpy front-end of the CPU oracle (oracle/oracle.c) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (ctc_asr_b200) never does.  "parity unpinned": see
oracle/oracle_impl.h.

Every wrapper takes/returns numpy arrays; dtype float32 -> *_f32 symbols, float64 -> *_f64.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle.c with the committed Makefile (gcc only, seconds)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("oracle.c", "oracle_impl.h", "beam_search.h", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class operand_rounding:
    """with ref.operand_rounding(True): every matrix product of the oracle rounds both operands to bfloat16 (the
    arithmetic of the product path's compute='bf16' mode, BASELINE cfg3); exact products otherwise."""

    def __init__(self, on=True):
        self.on = int(bool(on))

    def __enter__(self):
        self.prev = ctypes.c_int.in_dll(lib(), "oracle_operand_rounding").value
        lib().oracle_set_operand_rounding(self.on)
        return self

    def __exit__(self, *exc):
        lib().oracle_set_operand_rounding(self.prev)
        return False


NUM_GATES = {0: 1, 1: 1, 2: 4, 3: 3}
CELL_IDS = {"rnn_tanh": 0, "rnn_relu": 1, "lstm": 2, "gru": 3}


def ctc_loss(logits, labels, label_len, seq_len, blank=None, want_grad=True):
    """logits [T,B,V]; labels [B,Lmax] int32 -> (loss[B], grad[T,B,V] or None, status[B])."""
    dtype = logits.dtype
    suf, _ = _suf(dtype)
    logits = _c(logits, dtype)
    T, B, V = logits.shape
    blank = V - 1 if blank is None else blank
    labels = _c(labels, np.int32).reshape(B, -1)
    label_len = _c(label_len, np.int32)
    seq_len = _c(seq_len, np.int32)
    loss = np.zeros(B, dtype)
    grad = np.zeros((T, B, V), dtype) if want_grad else None
    status = np.zeros(B, np.int32)
    getattr(lib(), "oracle_ctc_loss" + suf)(
        _p(logits), T, B, V, blank, _p(labels), labels.shape[1], _p(label_len), _p(seq_len),
        _p(loss), _p(grad), _p(status))
    return loss, grad, status


def greedy_decode(logits, seq_len, blank=None):
    dtype = logits.dtype
    suf, _ = _suf(dtype)
    logits = _c(logits, dtype)
    T, B, V = logits.shape
    blank = V - 1 if blank is None else blank
    ids = np.zeros((B, T), np.int32)
    n = np.zeros(B, np.int32)
    getattr(lib(), "oracle_greedy_decode" + suf)(_p(logits), T, B, V, blank,
                                                 _p(_c(seq_len, np.int32)), _p(ids), _p(n))
    return ids, n


def dense_fwd(x, w, b, act=1, cutoff=20.0, drop_rate=0.0, seed=0):
    dtype = x.dtype
    suf, cr = _suf(dtype)
    x, w = _c(x, dtype), _c(w, dtype)
    b = _c(b, dtype) if b is not None else None
    M, K = x.shape
    N = w.shape[1]
    y = np.zeros((M, N), dtype)
    getattr(lib(), "oracle_dense_fwd" + suf)(_p(x), _p(w), _p(b), _p(y), M, K, N, act, cr(cutoff),
                                             ctypes.c_float(drop_rate), ctypes.c_uint32(seed))
    return y


def dense_bwd(x, w, y, dy, act=1, cutoff=20.0, drop_rate=0.0, seed=0, want_dx=True):
    dtype = x.dtype
    suf, cr = _suf(dtype)
    x, w, y, dy = (_c(a, dtype) for a in (x, w, y, dy))
    M, K = x.shape
    N = w.shape[1]
    dx = np.zeros((M, K), dtype) if want_dx else None
    dw = np.zeros((K, N), dtype)
    db = np.zeros(N, dtype)
    getattr(lib(), "oracle_dense_bwd" + suf)(_p(x), _p(w), _p(y), _p(dy), _p(dx), _p(dw), _p(db),
                                             M, K, N, act, cr(cutoff), ctypes.c_float(drop_rate),
                                             ctypes.c_uint32(seed))
    return dx, dw, db


def birnn_fwd(x, seq_len, wx, wh, bias, cell, use_len=True, forget_bias=1.0):
    """x [T,B,in]; wx [in,2GH]; wh [2,H,GH]; bias [2GH] -> y [T,B,2H], gates [2,T,B,GH], c [2,T,B,H]."""
    dtype = x.dtype
    suf, cr = _suf(dtype)
    x, wx, wh, bias = (_c(a, dtype) for a in (x, wx, wh, bias))
    T, B, nin = x.shape
    H = wh.shape[1]
    G = NUM_GATES[cell]
    nb = 2 * G * H + (2 * H if cell == 3 else 0)       # GRU: + b_rn [2, H] after the input-side biases
    assert wx.shape == (nin, 2 * G * H) and wh.shape == (2, H, G * H) and bias.shape == (nb,)
    y = np.zeros((T, B, 2 * H), dtype)
    gates = np.zeros((2, T, B, G * H), dtype)
    cst = np.zeros((2, T, B, H), dtype)
    rc = getattr(lib(), "oracle_birnn_fwd" + suf)(
        _p(x), _p(_c(seq_len, np.int32)), _p(wx), _p(wh), _p(bias), _p(y), _p(gates), _p(cst),
        T, B, nin, H, cell, int(use_len), cr(forget_bias))
    assert rc == 0
    return y, gates, cst


def birnn_bwd(x, seq_len, wx, wh, y, gates, cst, dy, cell, use_len=True, want_dx=True):
    dtype = x.dtype
    suf, _ = _suf(dtype)
    x, wx, wh, y, gates, cst, dy = (_c(a, dtype) for a in (x, wx, wh, y, gates, cst, dy))
    T, B, nin = x.shape
    H = wh.shape[1]
    G = NUM_GATES[cell]
    dx = np.zeros((T, B, nin), dtype) if want_dx else None
    dwx = np.zeros((nin, 2 * G * H), dtype)
    dwh = np.zeros((2, H, G * H), dtype)
    db = np.zeros(2 * G * H + (2 * H if cell == 3 else 0), dtype)
    rc = getattr(lib(), "oracle_birnn_bwd" + suf)(
        _p(x), _p(_c(seq_len, np.int32)), _p(wx), _p(wh), _p(y), _p(gates), _p(cst), _p(dy),
        _p(dx), _p(dwx), _p(dwh), _p(db), T, B, nin, H, cell, int(use_len))
    assert rc == 0
    return dx, dwx, dwh, db


def conv_out_size(n, k, s):
    """TF 'SAME': (output size, pad_before)."""
    out = -(-n // s)
    return out, max((out - 1) * s + k - n, 0) // 2


def _hash32(x):
    """oracle_hash32 (oracle_impl.h) on a uint32 array."""
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d); x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b); x ^= x >> np.uint32(16)
    return x


def drop_keep(seed, idx, rate):
    """oracle_drop_keep (oracle_impl.h) for an array of flat indices: the keep-mask shared bit for bit with the CUDA path."""
    idx = np.asarray(idx, np.uint64)
    with np.errstate(over="ignore"):
        hi = (idx >> np.uint64(32)).astype(np.uint32) * np.uint32(0x9e3779b9)
        h = _hash32((idx & np.uint64(0xffffffff)).astype(np.uint32) ^ _hash32(np.uint32(seed & 0xffffffff) ^ hi))
    return ((h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)) >= np.float32(rate)


def dropout(x, rate, seed, pitch=None):
    """y = keep ? x / (1 - rate) : 0 over the flat index of x viewed as [rows, cols]; `pitch` >= cols: the row pitch the
    product path stores the tensor with (conv outputs are padded to the GEMM width), which enters the index.
    tf.layers.dropout / DropoutWrapper / the cuDNN RNNs' inter-layer dropout; applied to a gradient it is the backward."""
    if rate <= 0:
        return x
    x = np.asarray(x)
    cols = x.shape[-1]
    flat = x.reshape(-1, cols)
    pitch = cols if pitch is None else pitch
    idx = np.arange(flat.shape[0], dtype=np.uint64)[:, None] * np.uint64(pitch) + np.arange(cols, dtype=np.uint64)[None, :]
    inv_keep = x.dtype.type(1.0 / (1.0 - float(np.float32(rate))))
    return (np.where(drop_keep(seed, idx, rate), flat * inv_keep, x.dtype.type(0))).reshape(x.shape)


def conv2d_fwd(x, w, b, strides, act=1, cutoff=20.0):
    """x [T,B,F,C], w [kt,kf,C,N] (TF HWIO, height = time), 'SAME' -> y [To,B,Fo,N]."""
    dtype = x.dtype
    suf, cr = _suf(dtype)
    x, w = _c(x, dtype), _c(w, dtype)
    b = _c(b, dtype) if b is not None else None
    T, B, F, C = x.shape
    kt, kf, C2, N = w.shape
    assert C2 == C
    To, Fo = conv_out_size(T, kt, strides[0])[0], conv_out_size(F, kf, strides[1])[0]
    y = np.zeros((To, B, Fo, N), dtype)
    getattr(lib(), "oracle_conv2d_fwd" + suf)(_p(x), _p(w), _p(b), _p(y), T, B, F, C, kt, kf, strides[0], strides[1],
                                              N, act, cr(cutoff))
    return y


def conv2d_bwd(x, w, y, dy, strides, act=1, cutoff=20.0, want_dx=True):
    dtype = x.dtype
    suf, cr = _suf(dtype)
    x, w, y, dy = (_c(a, dtype) for a in (x, w, y, dy))
    T, B, F, C = x.shape
    kt, kf, _, N = w.shape
    dx = np.zeros_like(x) if want_dx else None
    dw = np.zeros_like(w)
    db = np.zeros(N, dtype)
    getattr(lib(), "oracle_conv2d_bwd" + suf)(_p(x), _p(w), _p(y), _p(dy), _p(dx), _p(dw), _p(db), T, B, F, C, kt, kf,
                                              strides[0], strides[1], N, act, cr(cutoff))
    return dx, dw, db


def adam(p, m, v, g, step, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8):
    """In-place TF1 Adam on flat arrays."""
    dtype = p.dtype
    suf, cr = _suf(dtype)
    assert all(a.flags.c_contiguous and a.dtype == dtype for a in (p, m, v, g))
    getattr(lib(), "oracle_adam" + suf)(_p(p), _p(m), _p(v), _p(g), ctypes.c_size_t(p.size), step,
                                        cr(lr), cr(b1), cr(b2), cr(eps))


def edit_distance(hyp, hyp_len, truth, truth_len, normalize=True):
    hyp, truth = _c(hyp, np.int32), _c(truth, np.int32)
    B = hyp.shape[0]
    out = np.zeros(B, np.float32)
    lib().oracle_edit_distance(_p(hyp), hyp.shape[1], _p(_c(hyp_len, np.int32)), _p(truth), truth.shape[1],
                               _p(_c(truth_len, np.int32)), B, int(normalize), _p(out))
    return out


def ctc_beam_search(logits, seq_len, beam_width=1024, merge_repeated=False, blank=None, reoffer_wipe=False):
    """tf.nn.ctc_beam_search_decoder(top_paths=1) as called at asr/model.py:292-296.
    logits [T,B,V] float32 -> (ids [B,T] -1 padded, lengths [B], log-probability of the best leaf [B]).
    reoffer_wipe=True reproduces the order-dependent artifact of TF's Step() (oracle/beam_search.h)."""
    logits = _c(logits, np.float32)
    T, B, V = logits.shape
    blank = V - 1 if blank is None else blank
    ids = np.zeros((B, max(T, 1)), np.int32)
    n = np.zeros(B, np.int32)
    lp = np.zeros(B, np.float32)
    rc = lib().oracle_ctc_beam_search(_p(logits), T, B, V, blank, _p(_c(seq_len, np.int32)), int(beam_width),
                                      int(merge_repeated), int(reoffer_wipe), _p(ids), _p(n), _p(lp))
    assert rc == 0
    return ids[:, :T], n, lp
