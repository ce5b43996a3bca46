"""Hyper-parameters of the hot path, with the reference's flag names and defaults.

Mirrors the subset of `tf.flags` definitions the hot path reads (asr/params.py:31-50, 66-100,
106, 128, 138-148); `asr/model.py` reads them as `FLAGS.<name>` inside `inference_fn`
(asr/model.py:149,167-172,200-207,220,224-225,232).  Here they are one frozen dataclass that is
handed to `CTCModel`; `FLAGS` below is the module-level default instance so reference-style code
(`from ctc_asr_b200.params import FLAGS`) keeps reading the same attribute names.
"""
import dataclasses
from dataclasses import dataclass

from . import labels as _labels

NP_FLOAT = "float32"        # asr/params.py:137-139  (TF_FLOAT = tf.float32)
NUM_FEATURES = 80           # asr/params.py:148
WIN_LENGTH = 0.025          # asr/params.py:146
WIN_STEP = 0.010            # asr/params.py:147
MIN_EXAMPLE_LENGTH = 0.7    # asr/params.py:142
MAX_EXAMPLE_LENGTH = 17.0   # asr/params.py:143

# asr/util/tf_contrib.py:66-67 (defaults of conv_layers; height = time, width = features)
CONV_KERNEL_SIZES = ((11, 41), (11, 21), (11, 21))
CONV_STRIDES = ((2, 2), (1, 2), (1, 2))

RNN_CELLS = ("rnn_relu", "rnn_tanh", "lstm", "gru")     # asr/model.py:194-199
NUM_GATES = {"rnn_relu": 1, "rnn_tanh": 1, "lstm": 4, "gru": 3}
CELL_ID = {"rnn_tanh": 0, "rnn_relu": 1, "lstm": 2, "gru": 3}   # ctcasr.h CTCASR_CELL_*


@dataclass(frozen=True)
class ModelConfig:
    # --- names and defaults of asr/params.py -------------------------------------------------
    used_model: str = "ds2"             # :31  'ds1' dense front-end | 'ds2' conv front-end
    num_units_dense: int = 2048         # :35
    relu_cutoff: float = 20.0           # :37
    conv_filters: tuple = (32, 32, 96)  # :40
    num_layers_rnn: int = 4             # :43
    num_units_rnn: int = 2048           # :45
    rnn_cell: str = "rnn_relu"          # :48
    batch_size: int = 16                # :53
    learning_rate: float = 1e-5         # :66
    adam_beta1: float = 0.9             # :77
    adam_beta2: float = 0.999           # :79
    adam_epsilon: float = 1e-8          # :81
    beam_width: int = 1024              # :85
    conv_dropout_rate: float = 0.0      # :89  after every conv layer (asr/util/tf_contrib.py:135)
    rnn_dropout_rate: float = 0.0       # :91  DropoutWrapper in/out (TF path) or between layers (cuDNN path)
    dense_dropout_rate: float = 0.1     # :93
    num_buckets: int = 96               # :97
    num_classes: int = _labels.num_classes()    # :100  (29, blank = 28)
    cudnn: bool = True                  # :106  True: RNN ignores seq_length (cuDNN path semantics)
    random_seed: int = 0                # :128
    # --- knobs the reference hard-codes ------------------------------------------------------
    num_layers_dense: int = 3           # asr/util/tf_contrib.py:35 (num_layers=3)
    num_features: int = NUM_FEATURES
    # forget-gate bias of the TF path's LSTMCell (tf.nn.rnn_cell.LSTMCell default 1.0).  Applies to cudnn=False
    # only: CudnnLSTM (asr/model.py:194-216, the reference's only LSTM) adds none, see `forget_bias`.
    lstm_forget_bias: float = 1.0
    # --- B200 execution choices (not in the reference) ---------------------------------------
    # 'fp32'  : SIMT FFMA kernels everywhere (exact-order fp32)
    # 'bf16x3': tcgen05 kind::f16 on bf16-split operands, 3 (6 for ReLU-kinked layers) products
    #           accumulated in fp32 TMEM: 1e-5 .. 3e-5 of fp64 at tensor-core speed (default)
    # 'tf32'  : tcgen05 kind::tf32 on the fp32 operands, 2^-11 operand rounding
    # 'bf16'  : BASELINE cfg3: GEMM operands rounded to bf16 (one product), fp32 accumulation, fp32 master weights
    #           and fp32 CTC; the LSTM recurrence stays bf16x3.  Fastest; gradients within ~2e-2 of fp32
    compute: str = "bf16x3"
    decoder: str = "beam_search"        # decode_fn: 'beam_search' (asr/model.py:292-296) | 'greedy'

    def __post_init__(self):
        if self.used_model not in ("ds1", "ds2"):
            raise ValueError('Unsupported model "{}" in flags.'.format(self.used_model))  # asr/model.py:163
        object.__setattr__(self, "conv_filters", tuple(int(f) for f in self.conv_filters))
        if self.used_model == "ds2":
            if len(self.conv_filters) != len(CONV_KERNEL_SIZES):     # asr/util/tf_contrib.py:118-120
                raise ValueError("conv_layers(): Arguments filters, kernel_size, and strides must contain "
                                 "the same number of elements.")
            # the last layer's output is the RNN input [T', B, Fo*filters] as it stands in memory, so its
            # filter count is also its channel pitch: a legal tcgen05 GEMM width (the reference's is 96)
            if self.conv_filters[-1] < 64 or self.conv_filters[-1] % 8:
                raise ValueError("conv_filters[-1] must be a multiple of 8 and >= 64")
        if self.rnn_cell not in RNN_CELLS:
            raise ValueError("rnn_cell must be one of {}".format(RNN_CELLS))
        for name in ("conv_dropout_rate", "rnn_dropout_rate", "dense_dropout_rate"):      # asr/params.py:89-93
            if not 0.0 <= getattr(self, name) < 1.0:
                raise ValueError("%s must be in [0, 1)" % name)
        if self.compute not in ("fp32", "tf32", "bf16x3", "bf16"):
            raise ValueError("compute must be 'fp32', 'bf16x3', 'tf32' or 'bf16'")

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    @property
    def forget_bias(self):
        """What is added to the forget-gate pre-activation: TF LSTMCell's forget_bias on the TF path (cudnn=False),
        nothing on the cuDNN path (cuDNN's LSTM has no such term; its biases are plain parameters)."""
        return 0.0 if self.cudnn else self.lstm_forget_bias

    @property
    def num_gates(self):
        return NUM_GATES[self.rnn_cell]

    @property
    def blank(self):
        return self.num_classes - 1     # asr/labels.py:6 — TF ctc_loss: blank = num_classes - 1


FLAGS = ModelConfig()


def same_out(n, k, s):
    """tf 'SAME' padding: (output size, pad_before)."""
    out = -(-n // s)
    return out, max((out - 1) * s + k - n, 0) // 2


def conv_plan(cfg: ModelConfig, T=None):
    """Per conv layer of the ds2 front-end: dict(kt, kf, st, sf, C, filters, N, K, Kp, F, Fo[, T, To]).
    N = channel pitch of the layer's output = GEMM width: max(64, roundup(filters, 8))."""
    plan, F, C = [], cfg.num_features, 1
    for filt, (kt, kf), (st, sf) in zip(cfg.conv_filters, CONV_KERNEL_SIZES, CONV_STRIDES):
        K = kt * kf * C
        d = dict(kt=kt, kf=kf, st=st, sf=sf, C=C, filters=filt, N=max(64, (filt + 7) // 8 * 8), K=K,
                 Kp=(K + 7) // 8 * 8, F=F, Fo=same_out(F, kf, sf)[0])
        if T is not None:
            d["T"], d["To"] = T, same_out(T, kt, st)[0]
            T = d["To"]
        plan.append(d)
        F, C = d["Fo"], filt
    return plan


def conv_out_frames(cfg: ModelConfig, T):
    """Frames the RNN sees for T input frames (asr/util/tf_contrib.py:141-144: every utterance's
    seq_length becomes this number)."""
    return conv_plan(cfg, T)[-1]["To"] if cfg.used_model == "ds2" else T


def param_specs(cfg: ModelConfig):
    """Ordered (name, shape, init) of every trainable tensor in the flat parameter buffer.

    Names follow the reference's variable scopes ('dense', 'rnn', 'dense4', 'logits';
    asr/util/tf_contrib.py:50, asr/model.py:166,219,231).  RNN layer l holds
      rnn/l{l}/wx   [in_l, 2*G*H]   input kernels of the fw | bw cell side by side
      rnn/l{l}/wh   [2, H, G*H]     recurrent kernels fw, bw
      rnn/l{l}/bias [2*G*H]
    i.e. TF's fused cell kernel [in+H, G*H] of direction d is vstack(wx[:, d*GH:(d+1)*GH], wh[d]).
    init: 'truncnorm' = truncated_normal(stddev=0.046875) (asr/model.py:146), 'glorot' = Glorot
    uniform over the fused [in+H, G*H] matrix (rnn_cell default / asr/model.py:209), 'zeros'.
    """
    specs = []
    D, H, G, V = cfg.num_units_dense, cfg.num_units_rnn, cfg.num_gates, cfg.num_classes
    nin = cfg.num_features
    if cfg.used_model == "ds2":
        # scope 'conv' (asr/model.py:155): tf.layers.conv2d names conv2d, conv2d_1, ...; HWIO kernels,
        # glorot_normal (asr/util/tf_contrib.py:68): truncated normal, var = 2 / (fan_in + fan_out)
        for i, d in enumerate(conv_plan(cfg)):
            scope = "conv/conv2d" if i == 0 else "conv/conv2d_%d" % i
            rf = d["kt"] * d["kf"]
            specs.append((scope + "/kernel", (d["kt"], d["kf"], d["C"], d["filters"]),
                          ("glorot_normal", rf * d["C"], rf * d["filters"])))
            specs.append((scope + "/bias", (d["filters"],), "zeros"))
            nin = d["Fo"] * d["filters"]
    else:
        for i in range(cfg.num_layers_dense):
            scope = "dense/dense" if i == 0 else "dense/dense_%d" % i
            specs.append((scope + "/kernel", (nin, D), "truncnorm"))
            specs.append((scope + "/bias", (D,), "zeros"))
            nin = D
    for l in range(cfg.num_layers_rnn):
        specs.append(("rnn/l%d/wx" % l, (nin, 2 * G * H), ("glorot", nin + H, G * H)))
        specs.append(("rnn/l%d/wh" % l, (2, H, G * H), ("glorot", nin + H, G * H)))
        # GRU: + b_rn [2, H], the recurrent bias of the candidate gate (cuDNN formulation)
        specs.append(("rnn/l%d/bias" % l, (2 * G * H + (2 * H if cfg.rnn_cell == "gru" else 0),), "zeros"))
        nin = 2 * H
    specs.append(("dense4/dense/kernel", (nin, D), "truncnorm"))
    specs.append(("dense4/dense/bias", (D,), "zeros"))
    specs.append(("logits/dense/kernel", (D, V), "truncnorm"))
    specs.append(("logits/dense/bias", (V,), "zeros"))
    return specs


def storage_shape(cfg: ModelConfig, name, shape):
    """Shape a tensor occupies in the flat buffer.  Conv kernels are stored as the zero-padded GEMM
    operand [Kp, N] the kernels read (include/ctcasr.h, ctcasr_conv2d_fwd), conv biases as [N]; the
    reference-shaped tensor is a strided view of that block.  Everything else is stored as is."""
    if name.startswith("conv/"):
        i = 0 if "_" not in name.split("/")[1] else int(name.split("/")[1].rsplit("_", 1)[1])
        d = conv_plan(cfg)[i]
        return (d["Kp"], d["N"]) if name.endswith("/kernel") else (d["N"],)
    return tuple(shape)


def param_offsets(cfg: ModelConfig, align=64):
    """name -> (offset, shape) in elements inside the flat buffer; every tensor starts on an
    `align`-element (256 B) boundary so TMA descriptors and float4 accesses are always legal."""
    out, off = {}, 0
    for name, shape, _ in param_specs(cfg):
        n = 1
        for s in storage_shape(cfg, name, shape):
            n *= s
        out[name] = (off, shape)
        off += (n + align - 1) // align * align
    return out, off


def flops_per_frame_fwd(cfg: ModelConfig):
    """GEMM FLOPs per input frame, forward (2 FLOP/MAC) — the figure of BASELINE.md §4."""
    F, D, H, G, V = cfg.num_features, cfg.num_units_dense, cfg.num_units_rnn, cfg.num_gates, cfg.num_classes
    if cfg.used_model == "ds2":
        # per INPUT frame: conv MACs per output position, output positions per input frame (time strides)
        f, per_in, nin = 0.0, 1.0, None
        for d in conv_plan(cfg):
            per_in /= d["st"]
            f += 2.0 * d["K"] * d["filters"] * d["Fo"] * per_in
            nin = d["Fo"] * d["filters"]
        rest = 0
        for _ in range(cfg.num_layers_rnn):
            rest += 2 * 2 * (nin + H) * G * H
            nin = 2 * H
        rest += 2 * nin * D + 2 * D * V
        return f + rest * per_in
    f = 2 * F * D + (cfg.num_layers_dense - 1) * 2 * D * D
    nin = D
    for _ in range(cfg.num_layers_rnn):
        f += 2 * 2 * (nin + H) * G * H
        nin = 2 * H
    f += 2 * nin * D + 2 * D * V
    return f


def gradient_buckets(cfg: ModelConfig):
    """[lo, hi) ranges of the flat gradient buffer in the order the backward pass finishes them: dense4 + logits, the
    RNN layers from the last to the first, the front-end.  Each range is contiguous (param_specs order) and 256-B
    aligned, and together they cover the buffer, so a data-parallel caller can all-reduce a bucket while the layers
    below it are still in backward."""
    offsets, num_flat = param_offsets(cfg)
    names = [n for n, _, _ in param_specs(cfg)]
    starts = [offsets[n][0] for n in names] + [num_flat]

    def span(pred):
        sel = [i for i, n in enumerate(names) if pred(n)]
        return (min(starts[i] for i in sel), max(starts[i + 1] for i in sel))
    buckets = [span(lambda n: n.startswith("dense4/") or n.startswith("logits/"))]
    for l in reversed(range(cfg.num_layers_rnn)):
        buckets.append(span(lambda n, l=l: n.startswith("rnn/l%d/" % l)))
    buckets.append(span(lambda n: n.startswith("dense/") or n.startswith("conv/")))
    return buckets
