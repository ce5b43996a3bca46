"""torch-CPU restatement of the reference's TF graph for this path — TEST INFRASTRUCTURE ONLY.

Two uses: (1) an independent cross-check of the C oracle (different code, different maths
library, autograd instead of hand-written backward); (2) the "reference path timed beside it"
CPU baseline of bench.py (`--impl reference`, `cpu_baseline`), since TensorFlow 1.12 itself is
not installable offline.  "parity unpinned": see oracle/oracle_impl.h.

Mirrors, op for op, what the TF graph executes:
  tf.layers.dense -> relu -> tf.minimum(., relu_cutoff)           asr/util/tf_contrib.py:52-58
  stack_bidirectional_dynamic_rnn: per-timestep concat([x, h]) @ K + b -> gate math, `where`
  masking by sequence_length, reverse_sequence for the backward cell, concat(fw, bw) per layer
                                                                  asr/model.py:176-183
  dense4, logits, transpose to time-major                         asr/model.py:219-235
  tf.nn.ctc_loss(...) + tf.reduce_mean                            asr/model.py:259-267
"""
import torch
import torch.nn.functional as F


def params_to_torch(params, dtype=torch.float32, requires_grad=True):
    return {k: torch.tensor(v, dtype=dtype).requires_grad_(requires_grad) for k, v in params.items()}


def _dense_names(cfg):
    return ["dense/dense" if i == 0 else "dense/dense_%d" % i for i in range(cfg.num_layers_dense)]


def _reverse_sequence(x, lengths):
    """tf.reverse_sequence on [T,B,C] along time: only the first len_b frames of row b flip."""
    T, B = x.shape[0], x.shape[1]
    t = torch.arange(T).unsqueeze(1)                      # [T,1]
    ln = lengths.unsqueeze(0).to(torch.long)              # [1,B]
    idx = torch.where(t < ln, ln - 1 - t, t)              # [T,B]
    return torch.gather(x, 0, idx.unsqueeze(2).expand(-1, -1, x.shape[2]))


def _dynamic_rnn(cfg, x, lengths, kernel, bias):
    """tf.nn.dynamic_rnn(cell, x, sequence_length) on time-major x [T,B,in]; kernel [in+H, G*H]."""
    T, B, _ = x.shape
    H = cfg.num_units_rnn
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = []
    for t in range(T):
        if cfg.rnn_cell == "gru":                         # cuDNN GRU: r, z, n; bias = input-side [3H] + b_rn [H]
            nin = x.shape[2]
            px = x[t] @ kernel[:nin] + bias[:3 * H]
            rh = h @ kernel[nin:]
            r = torch.sigmoid(px[:, :H] + rh[:, :H])
            zg = torch.sigmoid(px[:, H:2 * H] + rh[:, H:2 * H])
            n = torch.tanh(px[:, 2 * H:] + r * (rh[:, 2 * H:] + bias[3 * H:]))
            h_new, c_new = (1 - zg) * n + zg * h, c
            z = None
        else:
            z = torch.cat([x[t], h], 1) @ kernel + bias
        if cfg.rnn_cell == "gru":
            pass
        elif cfg.rnn_cell == "lstm":
            i, j, f, o = z.split(H, 1)                    # TF LSTMCell gate order
            c_new = torch.sigmoid(f + cfg.forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
        else:
            h_new = torch.tanh(z) if cfg.rnn_cell == "rnn_tanh" else torch.relu(z)
            c_new = c
        if lengths is not None:
            live = (t < lengths).unsqueeze(1)
            outs.append(torch.where(live, h_new, torch.zeros_like(h_new)))
            h = torch.where(live, h_new, h)
            c = torch.where(live, c_new, c)
        else:
            outs.append(h_new)
            h, c = h_new, c_new
    return torch.stack(outs, 0)


CONV_STRIDES = ((2, 2), (1, 2), (1, 2))                  # asr/util/tf_contrib.py:67


def _conv_same(x, w, b, strides):
    """tf.layers.conv2d(padding='SAME') on NCHW x [B,C,T,F] with an HWIO kernel w [kt,kf,C,N]."""
    kt, kf = w.shape[0], w.shape[1]
    pads = []
    for n, k, s in ((x.shape[3], kf, strides[1]), (x.shape[2], kt, strides[0])):     # F.pad: last dim first
        out = -(-n // s)
        total = max((out - 1) * s + k - n, 0)
        pads += [total // 2, total - total // 2]
    return F.conv2d(F.pad(x, pads), w.permute(3, 2, 0, 1), b, stride=strides)


def conv_front_end(cfg, p, sequences):
    """asr/model.py:154-161 + asr/util/tf_contrib.py:123-144: [B,T,F] -> ([T',B,Fo*filters], T')."""
    x = sequences.unsqueeze(1)                            # [B,1,T,F]: height = time, width = features
    for i in range(len(cfg.conv_filters)):
        name = "conv/conv2d" if i == 0 else "conv/conv2d_%d" % i
        x = torch.clamp(torch.relu(_conv_same(x, p[name + "/kernel"], p[name + "/bias"], CONV_STRIDES[i])),
                        max=cfg.relu_cutoff)
    B, C, T, Fo = x.shape
    return x.permute(2, 0, 3, 1).reshape(T, B, Fo * C), T      # NHWC reshape [B,T,Fo*C], then time-major


def inference(cfg, p, sequences, seq_length):
    """sequences [B,T,F] -> logits [T,B,V] (dropout off; parity / eval mode)."""
    if getattr(cfg, "used_model", "ds1") == "ds2":
        x, T = conv_front_end(cfg, p, sequences)
        seq_length = torch.full_like(seq_length, T)
    else:
        x = sequences.transpose(0, 1)                     # time-major inside, like our layout
        for name in _dense_names(cfg):
            x = torch.clamp(torch.relu(x @ p[name + "/kernel"] + p[name + "/bias"]), max=cfg.relu_cutoff)
    lengths = None if cfg.cudnn else seq_length
    H, G = cfg.num_units_rnn, {"lstm": 4, "gru": 3}.get(cfg.rnn_cell, 1)
    T = x.shape[0]
    full = torch.full_like(seq_length, T)
    for l in range(cfg.num_layers_rnn):
        wx, wh, b = p["rnn/l%d/wx" % l], p["rnn/l%d/wh" % l], p["rnn/l%d/bias" % l]
        GH = G * H
        k_fw = torch.cat([wx[:, :GH], wh[0]], 0)
        k_bw = torch.cat([wx[:, GH:], wh[1]], 0)
        if cfg.rnn_cell == "gru":
            b_fw, b_bw = torch.cat([b[:GH], b[2 * GH:2 * GH + H]]), torch.cat([b[GH:2 * GH], b[2 * GH + H:]])
        else:
            b_fw, b_bw = b[:GH], b[GH:]
        fw = _dynamic_rnn(cfg, x, lengths, k_fw, b_fw)
        rl = seq_length if lengths is not None else full
        bw = _reverse_sequence(_dynamic_rnn(cfg, _reverse_sequence(x, rl), lengths, k_bw, b_bw), rl)
        x = torch.cat([fw, bw], 2)
    x = torch.clamp(torch.relu(x @ p["dense4/dense/kernel"] + p["dense4/dense/bias"]), max=cfg.relu_cutoff)
    return x @ p["logits/dense/kernel"] + p["logits/dense/bias"]


def loss(cfg, logits, seq_length, labels, label_len):
    """Mean over the batch of the per-utterance CTC negative log-likelihood."""
    per_utt = F.ctc_loss(F.log_softmax(logits, 2), labels.to(torch.long), seq_length.to(torch.long),
                         label_len.to(torch.long), blank=cfg.num_classes - 1, reduction="none",
                         zero_infinity=False)
    return per_utt.mean(), per_utt


def train_step_grads(cfg, p, sequences, seq_length, labels, label_len):
    logits = inference(cfg, p, sequences, seq_length)
    if getattr(cfg, "used_model", "ds1") == "ds2":
        seq_length = torch.full_like(seq_length, logits.shape[0])
    mean_loss, _ = loss(cfg, logits, seq_length, labels, label_len)
    for v in p.values():
        v.grad = None
    mean_loss.backward()
    return mean_loss.detach(), {k: v.grad for k, v in p.items()}, logits.detach()
