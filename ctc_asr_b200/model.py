"""`CTCModel` — host-side mirror of the reference's hot path behind the same method names.

Reference: asr/model.py:23-345.  `inference_fn(sequences, seq_length, training=True)`,
`loss_fn(logits, seq_length, labels)` and `decode_fn(logits, seq_len, originals=None)` keep the
reference's names, positional order, defaults and tensor layouts (batch-major sequences in,
TIME-major logits out, asr/model.py:233-235).  The reference's versions are graph-building
@staticmethods whose variables live in TF scopes and whose backward pass is TF autodiff; here the
parameters live in one flat fp32 buffer owned by the instance and the backward pass is the
explicit `backward()` over the buffers the forward pass saved.

Everything is computed by libctcasr.so (hand-written sm_100a kernels); torch only owns memory.
"""
import math

import numpy as np
import torch

from . import _lib, labels as _labels, ops
from .params import CELL_ID, FLAGS, ModelConfig, conv_plan, gradient_buckets, param_offsets, param_specs, storage_shape


class CTCModel:
    def __init__(self, config: ModelConfig = FLAGS, device="cuda", seed=1, params=None):
        _lib.load()                                   # fail loudly if the CUDA library is missing
        self.cfg = config
        self.device = torch.device(device)
        self.compute = _lib.COMPUTE_ID[config.compute]
        self.offsets, self.num_flat = param_offsets(config)
        self.flat = torch.zeros(self.num_flat, dtype=torch.float32, device=self.device)
        self.grad_flat = torch.zeros_like(self.flat)
        self.adam_m = torch.zeros_like(self.flat)
        self.adam_v = torch.zeros_like(self.flat)
        self.global_step = 0
        self.p = {k: self._view(self.flat, k) for k in self.offsets}          # reference-shaped views
        self.g = {k: self._view(self.grad_flat, k) for k in self.offsets}
        self.ps = {k: self._view(self.flat, k, storage=True) for k in self.offsets}       # as the kernels read them
        self.gs = {k: self._view(self.grad_flat, k, storage=True) for k in self.offsets}
        if params is None:
            from .synthetic import init_params
            params = init_params(config, seed=seed)
        self.load_params(params)
        self._saved = None
        self._pending, self._host_check = None, None
        self._bufs = {}
        self.dropout_seed = int(config.random_seed)

    # ---- parameter plumbing -------------------------------------------------------------------
    def _view(self, flat, name, storage=False):
        off, shape = self.offsets[name]
        st = storage_shape(self.cfg, name, shape)
        n = int(np.prod(st))
        block = flat[off:off + n].view(*st)
        if storage or tuple(st) == tuple(shape):
            return block
        if len(shape) == 1:                              # conv bias [filters] inside [N]
            return block[:shape[0]]
        kt, kf, C, filt = shape                          # conv kernel HWIO inside the zero-padded [Kp, N] operand
        N = st[1]
        return block.as_strided((kt, kf, C, filt), (kf * C * N, C * N, N, 1))

    def load_params(self, params):
        for name, shape, _ in param_specs(self.cfg):
            a = np.asarray(params[name], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError("parameter %s has shape %s, expected %s" % (name, a.shape, shape))
            self.p[name].copy_(torch.from_numpy(a))

    def params_numpy(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.p.items()}

    def grads_numpy(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.g.items()}

    def gradient_buckets(self):
        """`params.gradient_buckets(config)`: [lo, hi) ranges of the flat gradient buffer in the order `backward()`
        finishes them."""
        return gradient_buckets(self.cfg)

    @property
    def num_params(self):
        return sum(int(np.prod(s)) for _, s, _ in param_specs(self.cfg))

    def _buf(self, tag, shape, dtype=torch.float32):
        """Grow-only activation buffers keyed by role: no allocator traffic in steady state."""
        n = int(np.prod(shape))
        b = self._bufs.get(tag)
        if b is None or b.numel() < n or b.dtype != dtype:
            b = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._bufs[tag] = b
        return b[:n].view(*shape)

    def _dense_names(self):
        return ["dense/dense" if i == 0 else "dense/dense_%d" % i for i in range(self.cfg.num_layers_dense)]

    def _conv_names(self):
        return ["conv/conv2d" if i == 0 else "conv/conv2d_%d" % i for i in range(len(self.cfg.conv_filters))]

    # ---- asr/model.py:123-236 -----------------------------------------------------------------
    def inference_fn(self, sequences, seq_length, training=True):
        """sequences [B,T,F] float32, seq_length [B] int32 -> (logits [T,B,V], seq_length).
        The returned logits are a view of a buffer that the next inference_fn call reuses."""
        cfg = self.cfg
        sequences = torch.as_tensor(sequences)
        if sequences.dim() != 3 or sequences.shape[2] != cfg.num_features:
            raise ValueError("sequences must be [batch_size, time, %d]" % cfg.num_features)
        B, T, F = sequences.shape
        sequences = torch.as_tensor(sequences).to(self.device, torch.float32).contiguous()
        seq_length = torch.as_tensor(seq_length).to(self.device, torch.int32).contiguous()
        if seq_length.numel() != B:
            raise ValueError("seq_length must be [batch_size]")
        rate = cfg.dense_dropout_rate if training else 0.0
        seed = self.dropout_seed + 1000003 * self.global_step
        saved = {"T": T, "B": B, "rate": rate, "seed": seed, "seq_length": seq_length, "dense": [], "rnn": []}

        x = ops.transpose01(sequences, out=self._buf("xT", (T, B, F))).view(T * B, F)
        h = x
        if cfg.used_model == "ds2":
            # conv front-end (asr/model.py:154-161, asr/util/tf_contrib.py:123-144); [T,B,F] is [T,B,F,1]
            saved["conv"] = []
            xin, pitch = x, 1
            for li, d in enumerate(conv_plan(cfg, T)):
                name = self._conv_names()[li]
                y = self._buf("conv%d" % li, (d["To"] * B * d["Fo"], d["N"]))
                # conv dropout is on in every mode: the reference calls conv_layers() without its `training` argument
                # (asr/model.py:161), whose default is True (asr/util/tf_contrib.py:70,135)
                ops.conv2d_fwd(xin, pitch, self.ps[name + "/kernel"], self.ps[name + "/bias"], y, d["T"], B, d["F"],
                               d["C"], d["kt"], d["kf"], d["st"], d["sf"], act=1, cutoff=cfg.relu_cutoff,
                               drop_rate=cfg.conv_dropout_rate, seed=seed + 200 + li, compute=self.compute)
                saved["conv"].append((xin, pitch, y, d))
                xin, pitch = y, d["N"]
            T = d["To"]                                   # every utterance is stretched to the conv length of
            seq_length = torch.full_like(seq_length, T)   # the longest one (asr/util/tf_contrib.py:141-144)
            saved["T"], saved["seq_length"] = T, seq_length
            h = xin.view(T * B, d["Fo"] * d["N"])         # [T', B, Fo*filters]: the reshape at tf_contrib.py:138
        else:
            for li, name in enumerate(self._dense_names()):
                y = ops.dense_fwd(h, self.p[name + "/kernel"], self.p[name + "/bias"], act=1, cutoff=cfg.relu_cutoff,
                                  drop_rate=rate, seed=seed + li, compute=self.compute,
                                  out=self._buf("dense%d" % li, (T * B, cfg.num_units_dense)))
                saved["dense"].append((h, y))
                h = y
        cell = CELL_ID[cfg.rnn_cell]
        use_len = not cfg.cudnn
        H = cfg.num_units_rnn
        # Dropout on the non-recurrent connections (asr/model.py:167): the TF path wraps every cell in
        # DropoutWrapper(input_keep_prob, output_keep_prob) (asr/util/tf_contrib.py:190-194) — the layer's input and its
        # output are dropped, the state that is fed back is not; the cuDNN RNNs drop between layers only
        # (asr/model.py:201-206).  The layer itself always sees / keeps un-dropped outputs: backward needs them.
        rnn_rate = cfg.rnn_dropout_rate if training else 0.0
        saved["rnn_rate"] = rnn_rate
        L = cfg.num_layers_rnn
        for l in range(L):
            nin = h.shape[1]
            if rnn_rate > 0.0 and not cfg.cudnn:
                h = ops.dropout(h, rnn_rate, seed + 300 + l, out=self._buf("rnn_xd%d" % l, (T * B, nin)))
            rb, _ = ops.birnn_sizes(T, B, nin, H, cell)
            reserve = self._buf("rnn_reserve%d" % l, (rb,), torch.uint8)
            y = self._buf("rnn_y%d" % l, (T, B, 2 * H))
            ops.birnn_fwd(h.view(T, B, nin), seq_length, self.p["rnn/l%d/wx" % l], self.p["rnn/l%d/wh" % l],
                          self.p["rnn/l%d/bias" % l], y, reserve, cell, use_len, cfg.forget_bias, self.compute)
            saved["rnn"].append((h, y, reserve))
            h = y.view(T * B, 2 * H)
            if rnn_rate > 0.0 and (not cfg.cudnn or l < L - 1):
                h = ops.dropout(h, rnn_rate, seed + 400 + l, out=self._buf("rnn_yd%d" % l, (T * B, 2 * H)))
        y4 = ops.dense_fwd(h, self.p["dense4/dense/kernel"], self.p["dense4/dense/bias"], act=1,
                           cutoff=cfg.relu_cutoff, drop_rate=rate, seed=seed + 100, compute=self.compute,
                           out=self._buf("dense4", (T * B, cfg.num_units_dense)))
        saved["d4"] = (h, y4)
        logits = ops.dense_fwd(y4, self.p["logits/dense/kernel"], self.p["logits/dense/bias"], act=0,
                               compute=self.compute, out=self._buf("logits", (T * B, cfg.num_classes)))
        logits = logits.view(T, B, cfg.num_classes)          # already time-major (asr/model.py:233)
        saved["logits"] = logits
        self._saved = saved
        return logits, seq_length

    # ---- asr/model.py:238-269 -----------------------------------------------------------------
    @staticmethod
    def _labels_to_padded(labels, device):
        """Accept what the reference passes (an int32 SparseTensor, asr/model.py:71) as a torch
        sparse COO tensor, or (padded [B,Lmax], lengths [B]), or a 0-padded dense [B,Lmax]."""
        if isinstance(labels, (tuple, list)):
            padded, lengths = (torch.as_tensor(a) for a in labels)      # numpy (the input pipeline) or torch
            return (padded.to(device, torch.int32).contiguous(), lengths.to(device, torch.int32).contiguous())
        labels = torch.as_tensor(labels)
        if labels.is_sparse:
            labels = labels.coalesce()
            idx, val = labels.indices(), labels.values()
            B, lmax = labels.shape
            padded = torch.zeros((B, max(lmax, 1)), dtype=torch.int32, device=idx.device)
            padded[idx[0], idx[1]] = val.to(torch.int32)
            lengths = torch.zeros(B, dtype=torch.int64, device=idx.device)
            lengths.scatter_reduce_(0, idx[0], idx[1] + 1, reduce="amax")
            return padded.to(device).contiguous(), lengths.to(device, torch.int32).contiguous()
        padded = labels.to(device, torch.int32).contiguous()
        lengths = (padded != _labels.PAD_ID).sum(1).to(torch.int32)   # dense_to_sparse(eos_token=0)
        return padded, lengths.contiguous()

    def loss_fn(self, logits, seq_length, labels, global_batch=None, defer_check=False):
        """Mean CTC loss over the batch (asr/model.py:259-267).  Also leaves d loss / d logits in
        the model for `backward()`.  `global_batch` (data parallel): number of utterances the mean
        runs over across all ranks; defaults to this batch.  defer_check: do not read the per-utterance
        status words back here (no host synchronisation); `train_step` checks them with the loss."""
        padded, lengths = self._labels_to_padded(labels, self.device)
        T, B, V = logits.shape
        seq_length = torch.as_tensor(seq_length).to(self.device, torch.int32).contiguous()
        gb = B if global_batch is None else global_batch
        per_utt, dlogits, status = ops.ctc_loss(logits, padded, lengths, seq_length, blank=self.cfg.blank,
                                                grad=True, grad_scale=1.0 / gb,
                                                out_grad=self._buf("dlogits", (T, B, V)))
        self.last_status = status
        if not defer_check:
            self._raise_on_status(status)                # the reference's op raises InvalidArgumentError
        if self._saved is not None and self._saved.get("logits") is not None \
                and self._saved["logits"].data_ptr() == logits.data_ptr():
            self._saved["dlogits"] = dlogits
        self.last_per_utterance_loss = per_utt
        return per_utt.sum() / gb if global_batch is not None else per_utt.mean()

    @staticmethod
    def _raise_on_status(status):
        st = status.cpu().tolist()
        bad = sum(1 for v in st if v != 0)
        if bad:
            raise ValueError("ctc_loss: %d utterance(s) rejected, status per utterance %s "
                             "(1: not enough time for target transition sequence, 2: label out of range, "
                             "3: sequence_length > max_time)" % (bad, st))

    # ---- asr/model.py:271-309 ------------------------------------------------------------------
    def decode_fn(self, logits, seq_len, originals=None, decoder=None):
        """-> (decoded ids as a list of int32 tensors, plaintexts, plaintext summary rows).
        decoder: 'beam_search' (the reference: tf.nn.ctc_beam_search_decoder, beam_width=FLAGS.beam_width,
        top_paths=1, merge_repeated=False, asr/model.py:292-296) or 'greedy' (the decoder the reference's
        comment at :290 mentions as the faster alternative); default `config.decoder`."""
        decoder = self.cfg.decoder if decoder is None else decoder
        seq_len = seq_len.to(self.device, torch.int32).contiguous()
        if decoder == "beam_search":
            ids, n, _ = ops.beam_search(logits, seq_len, beam_width=self.cfg.beam_width, merge_repeated=False,
                                        blank=self.cfg.blank)
        elif decoder == "greedy":
            ids, n = ops.greedy_decode(logits, seq_len, blank=self.cfg.blank)
        else:
            raise ValueError("decoder must be 'beam_search' or 'greedy'")
        ids_h, n_h = ids.cpu(), n.cpu().tolist()
        decoded = [ids_h[b, :n_h[b]] for b in range(len(n_h))]
        plaintext = [_labels.ids_to_text(d.tolist()) for d in decoded]
        if originals is None:
            summary = [[p] for p in plaintext]
        else:
            summary = [[p, o] for p, o in zip(plaintext, originals)]
        return decoded, plaintext, summary

    # ---- what AdamOptimizer.minimize differentiates (asr/model.py:79-83) ----------------------
    def backward(self, on_bucket=None):
        """d(mean loss)/d(parameters) into self.grad_flat (overwritten).  on_bucket(lo, hi) is called as soon as the
        kernels that produce grad_flat[lo:hi] are enqueued (the ranges of `gradient_buckets()`, in that order)."""
        buckets = self.gradient_buckets() if on_bucket is not None else None
        s = self._saved
        if s is None or "dlogits" not in s:
            raise RuntimeError("backward() needs inference_fn() followed by loss_fn() on its logits")
        cfg = self.cfg
        T, B, rate, seed = s["T"], s["B"], s["rate"], s["seed"]
        D, H = cfg.num_units_dense, cfg.num_units_rnn
        dy = s["dlogits"].view(T * B, cfg.num_classes)
        h, y4 = s["d4"]
        d4 = self._buf("g_d4", (T * B, D))
        ops.dense_bwd(y4, self.p["logits/dense/kernel"], None, dy, self.g["logits/dense/kernel"],
                      self.g["logits/dense/bias"], dx=d4, act=0, compute=self.compute)
        drnn = self._buf("g_rnn", (T * B, 2 * H))
        ops.dense_bwd(h, self.p["dense4/dense/kernel"], y4, d4, self.g["dense4/dense/kernel"],
                      self.g["dense4/dense/bias"], dx=drnn, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate,
                      seed=seed + 100, compute=self.compute)
        if on_bucket is not None:
            on_bucket(*buckets[0])
        cell = CELL_ID[cfg.rnn_cell]
        use_len = not cfg.cudnn
        dy = drnn
        rnn_rate, L = s.get("rnn_rate", 0.0), cfg.num_layers_rnn
        for l in reversed(range(L)):
            x, y, reserve = s["rnn"][l]
            nin = x.shape[1]
            if rnn_rate > 0.0 and (not cfg.cudnn or l < L - 1):
                ops.dropout(dy, rnn_rate, seed + 400 + l, out=dy)          # the layer's output dropout, on the gradient
            dx = self._buf("g_rnn_in%d" % (l % 2), (T * B, nin))
            ops.birnn_bwd(x.view(T, B, nin), s["seq_length"], self.p["rnn/l%d/wx" % l], self.p["rnn/l%d/wh" % l],
                          y, reserve, dy, dx, self.g["rnn/l%d/wx" % l], self.g["rnn/l%d/wh" % l],
                          self.g["rnn/l%d/bias" % l], cell, use_len, self.compute)
            if rnn_rate > 0.0 and not cfg.cudnn:
                ops.dropout(dx, rnn_rate, seed + 300 + l, out=dx)          # its input dropout
            dy = dx
            if on_bucket is not None:
                on_bucket(*buckets[cfg.num_layers_rnn - l])
        if cfg.used_model == "ds2":
            names = self._conv_names()
            for li in reversed(range(len(names))):
                xin, pitch, y, d = s["conv"][li]
                dx = self._buf("g_conv%d" % (li % 2), (d["T"] * B * d["F"], pitch)) if li > 0 else None
                ops.conv2d_bwd(xin, pitch, self.ps[names[li] + "/kernel"], y, dy, dx, self.gs[names[li] + "/kernel"],
                               self.gs[names[li] + "/bias"], d["T"], B, d["F"], d["C"], d["kt"], d["kf"], d["st"], d["sf"],
                               act=1, cutoff=cfg.relu_cutoff, drop_rate=cfg.conv_dropout_rate, seed=seed + 200 + li,
                               compute=self.compute)
                dy = dx
            if on_bucket is not None:
                on_bucket(*buckets[-1])
            return self.grad_flat
        names = self._dense_names()
        for li in reversed(range(len(names))):
            x, y = s["dense"][li]
            dx = self._buf("g_dense%d" % (li % 2), (T * B, x.shape[1])) if li > 0 else None
            ops.dense_bwd(x, self.p[names[li] + "/kernel"], y, dy, self.g[names[li] + "/kernel"],
                          self.g[names[li] + "/bias"], dx=dx, act=1, cutoff=cfg.relu_cutoff, drop_rate=rate,
                          seed=seed + li, compute=self.compute)
            dy = dx
        if on_bucket is not None:
            on_bucket(*buckets[-1])
        return self.grad_flat

    def apply_gradients(self, grad_scale=1.0):
        """One Adam step (TF1 formulation, asr/model.py:79-83, asr/params.py:66-82)."""
        cfg = self.cfg
        self.global_step += 1
        ops.adam(self.flat, self.adam_m, self.adam_v, self.grad_flat, self.global_step, cfg.learning_rate,
                 cfg.adam_beta1, cfg.adam_beta2, cfg.adam_epsilon, grad_scale)

    def train_step(self, sequences, seq_length, labels, global_batch=None, allreduce=None, overlap=False):
        """model_fn's TRAIN branch (asr/model.py:53-54, 74, 79-83) on one batch.  Returns the loss as a 0-d device
        tensor.  Nothing in the step waits for the GPU: the checks the reference makes on its values (CTCLoss's
        InvalidArgumentError, NanTensorHook(loss) at asr/model.py:368) are made on the previous step's results when the
        next step starts, or by `check_step()`.
        allreduce(tensor, async_op=False): sums a slice of the flat gradient over the data-parallel ranks in place and,
        with async_op=True, returns a handle with .wait() (torch.distributed semantics).  overlap=True issues one
        asynchronous all-reduce per `gradient_buckets()` range as soon as backward has produced it; the default is one
        all-reduce after backward: measured on 2 x B200 at cfg2 the overlapped form is no faster (151.4 vs 151.0 ms/step) —
        the collective's CTAs take SMs the persistent recurrence kernels of the next layer need all at once, which
        start late by about the collective's duration (profiles/r2_multi_gpu.md)."""
        self.check_step()
        logits, seq_length = self.inference_fn(sequences, seq_length, training=True)
        loss = self.loss_fn(logits, seq_length, labels, global_batch=global_batch, defer_check=True)
        if allreduce is None:
            self.backward()
        elif overlap:
            # data parallel: every bucket of the gradient is summed over the ranks (asynchronously, on the collective
            # library's stream) while the backward pass of the layers below it is still running
            works = []
            self.backward(on_bucket=lambda lo, hi: works.append(allreduce(self.grad_flat[lo:hi], async_op=True)))
            for w in works:
                if w is not None:
                    w.wait()
        else:
            self.backward()
            allreduce(self.grad_flat)
        self.apply_gradients()
        # loss + "any utterance rejected" travel to a pinned host slot behind the step's kernels
        if self._host_check is None:
            self._host_check = torch.empty(2, dtype=torch.float32).pin_memory()
        dev = torch.stack([loss.reshape(()), (self.last_status != 0).any().to(torch.float32)])
        self._host_check.copy_(dev, non_blocking=True)
        self._pending = (torch.cuda.Event(), self.last_status)
        self._pending[0].record()
        return loss

    def check_step(self):
        """Raise what the reference raises for the last `train_step` (waits for that step to finish)."""
        if self._pending is None:
            return
        ev, status = self._pending
        self._pending = None
        ev.synchronize()
        if self._host_check[1] != 0:
            self._raise_on_status(status)
        if not math.isfinite(float(self._host_check[0])):  # NanTensorHook(loss), asr/model.py:368
            raise FloatingPointError("Model diverged with loss = NaN/Inf")
