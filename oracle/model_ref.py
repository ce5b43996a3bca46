"""Whole-path oracle: the reference's inference_fn + loss_fn + backward, composed from oracle/ref.py.

TEST INFRASTRUCTURE ONLY ("parity unpinned", see oracle/oracle_impl.h).  Follows
asr/model.py:123-269: dense stack (asr/util/tf_contrib.py:50-61) -> stacked bidirectional RNN
(asr/model.py:169-216) -> dense4 (asr/model.py:219-226) -> logits, transposed to time-major
(asr/model.py:231-235) -> tf.nn.ctc_loss + reduce_mean (asr/model.py:259-267).

`cfg` is any object with the attribute names of ctc_asr_b200.params.ModelConfig; `params` is a
dict name -> numpy array with the names/shapes of ctc_asr_b200.params.param_specs.
"""
import numpy as np

from . import ref

_CELL = {"rnn_tanh": 0, "rnn_relu": 1, "lstm": 2, "gru": 3}
# defaults of conv_layers (asr/util/tf_contrib.py:66-67); height = time, width = features
CONV_KERNEL_SIZES = ((11, 41), (11, 21), (11, 21))
CONV_STRIDES = ((2, 2), (1, 2), (1, 2))


def _conv_names(cfg):
    return ["conv/conv2d" if i == 0 else "conv/conv2d_%d" % i for i in range(len(cfg.conv_filters))]


def _dense_names(cfg):
    return ["dense/dense" if i == 0 else "dense/dense_%d" % i for i in range(cfg.num_layers_dense)]


def forward(cfg, params, sequences, seq_length, training=False, seed=0, dtype=np.float64, round_operands=False):
    """sequences [B,T,F] batch-major (asr/model.py:129) -> logits [T,B,V] time-major + cache.
    round_operands: bf16 operand rounding in every matrix product but the 29-class logits layer (which the product
    path's compute='bf16' mode runs in exact fp32, its width being no tensor-core tile)."""
    with ref.operand_rounding(round_operands):
        return _forward(cfg, params, sequences, seq_length, training, seed, dtype)


def _forward(cfg, params, sequences, seq_length, training, seed, dtype):
    x = np.ascontiguousarray(np.transpose(np.asarray(sequences, dtype), (1, 0, 2)))   # [T,B,F]
    T, B, _ = x.shape
    seq_length = np.asarray(seq_length, np.int32)
    rate = cfg.dense_dropout_rate if training else 0.0
    cache = {"acts": [], "rnn": [], "conv": []}
    h = x.reshape(T * B, -1)
    if getattr(cfg, "used_model", "ds1") == "ds2":
        # asr/model.py:154-161: expand_dims(sequences, 3) -> conv_layers; asr/util/tf_contrib.py:123-144
        x4 = x.reshape(T, B, -1, 1)
        crate = getattr(cfg, "conv_dropout_rate", 0.0)      # on in every mode: conv_layers() is called without `training`
        for li, (name, strides) in enumerate(zip(_conv_names(cfg), CONV_STRIDES)):
            w, b = params[name + "/kernel"].astype(dtype), params[name + "/bias"].astype(dtype)
            y = ref.conv2d_fwd(x4, w, b, strides, act=1, cutoff=cfg.relu_cutoff)
            cache["conv"].append((x4, y))
            # tf.layers.dropout after the clipped ReLU (asr/util/tf_contrib.py:135); the index runs over the padded pitch
            x4 = ref.dropout(y, crate, seed + 200 + li, pitch=max(64, (y.shape[-1] + 7) // 8 * 8))
        T = x4.shape[0]
        seq_length = np.full(B, T, np.int32)     # tf.tile([shape(output)[1]], [batch]) (tf_contrib.py:144)
        h = x4.reshape(T * B, -1)                # [T', B, Fo * filters] (tf_contrib.py:138)
    else:
        for li, name in enumerate(_dense_names(cfg)):
            w, b = params[name + "/kernel"].astype(dtype), params[name + "/bias"].astype(dtype)
            y = ref.dense_fwd(h, w, b, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate, seed=seed + li)
            cache["acts"].append((h, y))
            h = y
    cell = _CELL[cfg.rnn_cell]
    use_len = not cfg.cudnn
    # DropoutWrapper(input_keep_prob, output_keep_prob) on the TF path (asr/util/tf_contrib.py:190-194), inter-layer
    # dropout on the cuDNN path (asr/model.py:201-206); rnn_dropout_rate if training else 0.0 (asr/model.py:167)
    rrate = getattr(cfg, "rnn_dropout_rate", 0.0) if training else 0.0
    L = cfg.num_layers_rnn
    for l in range(L):
        wx, wh, bias = (params["rnn/l%d/%s" % (l, k)].astype(dtype) for k in ("wx", "wh", "bias"))
        if not cfg.cudnn:
            h = ref.dropout(h, rrate, seed + 300 + l)
        xin = h.reshape(T, B, -1)
        y, gates, cst = ref.birnn_fwd(xin, seq_length, wx, wh, bias, cell, use_len=use_len,
                                      forget_bias=cfg.forget_bias)
        cache["rnn"].append((xin, y, gates, cst))
        h = y.reshape(T * B, -1)
        if not cfg.cudnn or l < L - 1:
            h = ref.dropout(h, rrate, seed + 400 + l)
    w, b = params["dense4/dense/kernel"].astype(dtype), params["dense4/dense/bias"].astype(dtype)
    y4 = ref.dense_fwd(h, w, b, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate, seed=seed + 100)
    cache["d4"] = (h, y4)
    w, b = params["logits/dense/kernel"].astype(dtype), params["logits/dense/bias"].astype(dtype)
    with ref.operand_rounding(False):
        logits = ref.dense_fwd(y4, w, b, act=0)
    cache["lg"] = (y4,)
    cache["meta"] = (T, B, rate, seed)
    cache["rrate"] = rrate
    cache["seq_length"] = seq_length
    return logits.reshape(T, B, -1), cache


def loss_and_grads(cfg, params, sequences, seq_length, labels, label_len, training=False, seed=0,
                   dtype=np.float64, round_operands=False):
    """Mean CTC loss (asr/model.py:267) and d loss / d params (what AdamOptimizer.minimize
    differentiates, asr/model.py:83).  Returns (loss, grads dict, logits, dlogits)."""
    with ref.operand_rounding(round_operands):
        return _loss_and_grads(cfg, params, sequences, seq_length, labels, label_len, training, seed, dtype)


def _loss_and_grads(cfg, params, sequences, seq_length, labels, label_len, training, seed, dtype):
    logits, cache = _forward(cfg, params, sequences, seq_length, training, seed, dtype)
    T, B, rate, seed = cache["meta"]
    seq_length = cache["seq_length"]             # ds2: the conv length for every utterance
    loss_b, g, status = ref.ctc_loss(logits, labels, label_len, seq_length, blank=cfg.num_classes - 1)
    assert (status == 0).all(), status
    loss = loss_b.mean()
    dlogits = g / B
    grads = {}
    dy = dlogits.reshape(T * B, -1)
    (y4,) = cache["lg"]
    w = params["logits/dense/kernel"].astype(dtype)
    with ref.operand_rounding(False):
        dy, grads["logits/dense/kernel"], grads["logits/dense/bias"] = ref.dense_bwd(
            y4, w, y4[:, :1], dy, act=0)
    h, y4 = cache["d4"]
    w = params["dense4/dense/kernel"].astype(dtype)
    dy, grads["dense4/dense/kernel"], grads["dense4/dense/bias"] = ref.dense_bwd(
        h, w, y4, dy, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate, seed=seed + 100)
    cell = _CELL[cfg.rnn_cell]
    use_len = not cfg.cudnn
    rrate, L = cache["rrate"], cfg.num_layers_rnn
    for l in reversed(range(L)):
        xin, y, gates, cst = cache["rnn"][l]
        wx, wh = (params["rnn/l%d/%s" % (l, k)].astype(dtype) for k in ("wx", "wh"))
        if not cfg.cudnn or l < L - 1:
            dy = ref.dropout(dy, rrate, seed + 400 + l)
        dx, dwx, dwh, db = ref.birnn_bwd(xin, seq_length, wx, wh, y, gates, cst,
                                         dy.reshape(T, B, -1), cell, use_len=use_len)
        grads["rnn/l%d/wx" % l], grads["rnn/l%d/wh" % l], grads["rnn/l%d/bias" % l] = dwx, dwh, db
        dy = dx.reshape(T * B, -1)
        if not cfg.cudnn:
            dy = ref.dropout(dy, rrate, seed + 300 + l)
    if getattr(cfg, "used_model", "ds1") == "ds2":
        names = _conv_names(cfg)
        crate = getattr(cfg, "conv_dropout_rate", 0.0)
        for li in reversed(range(len(names))):
            x4, y = cache["conv"][li]            # y: the layer's output before its dropout
            w = params[names[li] + "/kernel"].astype(dtype)
            dy = ref.dropout(dy.reshape(y.shape), crate, seed + 200 + li, pitch=max(64, (y.shape[-1] + 7) // 8 * 8))
            dy, grads[names[li] + "/kernel"], grads[names[li] + "/bias"] = ref.conv2d_bwd(
                x4, w, y, dy.reshape(y.shape), CONV_STRIDES[li], act=1, cutoff=cfg.relu_cutoff, want_dx=li > 0)
        return loss, grads, logits, dlogits
    names = _dense_names(cfg)
    for li in reversed(range(len(names))):
        h, y = cache["acts"][li]
        w = params[names[li] + "/kernel"].astype(dtype)
        dy, grads[names[li] + "/kernel"], grads[names[li] + "/bias"] = ref.dense_bwd(
            h, w, y, dy, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate, seed=seed + li, want_dx=li > 0)
    return loss, grads, logits, dlogits
