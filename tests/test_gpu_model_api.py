"""The reference-facing surface of `CTCModel` on the GPU: every label form `loss_fn` accepts
(asr/model.py:71 hands tf.nn.ctc_loss a SparseTensor made by dense_to_sparse from 0-padded rows), the deferred
checks of `train_step`, and one epoch of the CSV + WAV corpus format through input_fn_generator -> train_step."""
import os
import types
import wave

import numpy as np
import pytest
import torch

from ctc_asr_b200 import _lib, input_pipeline as ip, ops, synthetic
from ctc_asr_b200.model import CTCModel
from ctc_asr_b200.params import ModelConfig
from oracle import ref

pytestmark = pytest.mark.gpu

CFG = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=64, num_layers_rnn=1, num_units_rnn=64,
                  rnn_cell="lstm", cudnn=False, dense_dropout_rate=0.0, compute="fp32")


def _ragged_labels(B, lmax, seed=0):
    rng = np.random.default_rng(seed)
    ll = rng.integers(1, lmax + 1, B).astype(np.int32)
    ll[0] = lmax
    lab = np.zeros((B, lmax), np.int32)
    for b in range(B):
        lab[b, :ll[b]] = rng.integers(1, 28, ll[b])
    return lab, ll


def test_loss_fn_accepts_sparse_dense_and_tuple_labels():
    """The three label forms give the same loss and the same gradient, equal to the oracle's."""
    B, T, lmax = 6, 40, 9
    model = CTCModel(CFG, seed=1)
    x, sl, _, _ = synthetic.fixed_batch(B, T, 4, seed=0)
    lab, ll = _ragged_labels(B, lmax)
    logits, sl_out = model.inference_fn(torch.from_numpy(x), torch.from_numpy(sl), training=False)
    ol, og, _ = ref.ctc_loss(logits.cpu().numpy().astype(np.float64), lab, ll, sl)
    idx = np.array([(b, i) for b in range(B) for i in range(ll[b])]).T
    sparse = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(lab[idx[0], idx[1]]), (B, lmax))   # the SparseTensor of asr/model.py:71
    forms = {"tuple": (torch.from_numpy(lab), torch.from_numpy(ll)), "tuple_numpy": (lab, ll), "dense": torch.from_numpy(lab),
             "dense_numpy": lab, "sparse": sparse, "sparse_cuda": sparse.cuda()}
    for name, labels in forms.items():
        loss = model.loss_fn(logits, sl_out, labels)
        assert abs(float(loss) - ol.mean()) / ol.mean() < 1e-5, name
        g = model._saved["dlogits"].cpu().numpy() * B
        assert np.abs(g - og).max() / np.abs(og).max() < 1e-4, name
    # a row of zeros (an empty transcript) is a legal dense label row: length 0
    lab0 = lab.copy(); lab0[2] = 0
    loss0 = model.loss_fn(logits, sl_out, torch.from_numpy(lab0))
    ll0 = ll.copy(); ll0[2] = 0
    want = ref.ctc_loss(logits.cpu().numpy().astype(np.float64), lab0, ll0, sl)[0].mean()
    assert abs(float(loss0) - want) / want < 1e-5


def test_train_step_checks_are_deferred_but_not_lost():
    """train_step does not wait for the GPU; the InvalidArgumentError of tf.nn.ctc_loss (here: a label sequence
    longer than the utterance) surfaces at the next step or at check_step()."""
    B, T = 4, 12
    model = CTCModel(CFG, seed=1)
    x, sl, lab, ll = synthetic.fixed_batch(B, T, 4, seed=0)
    batch = (torch.from_numpy(x), torch.from_numpy(sl), (lab, ll))
    l0 = float(model.train_step(*batch))
    model.check_step()
    assert np.isfinite(l0)
    bad_lab, bad_ll = _ragged_labels(B, 2 * T)           # needs more frames than there are
    model.train_step(batch[0], batch[1], (bad_lab, bad_ll))
    with pytest.raises(ValueError, match="not enough time"):
        model.check_step()
    model.train_step(batch[0], batch[1], (bad_lab, bad_ll))
    with pytest.raises(ValueError):
        model.train_step(*batch)                          # ... or when the next step starts
    with pytest.raises(ValueError):
        model.loss_fn(*model.inference_fn(batch[0], batch[1]), (bad_lab, bad_ll))     # loss_fn on its own raises at once


def _write_corpus(root, n=26, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "corpus"))
    durations = np.sort(rng.uniform(0.7, 2.0, n))
    rows = ["path;label;length"]
    for i, d in enumerate(durations):
        t = np.arange(int(d * 16000)) / 16000.0
        pcm = (3000 * np.sin(2 * np.pi * (200 + 40 * i) * t) + 300 * rng.standard_normal(t.size)).astype("<i2")
        with wave.open(os.path.join(root, "corpus", "u%03d.wav" % i), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        text = "".join(rng.choice(list("abc de"), size=rng.integers(2, 7))).strip() or "a"
        rows.append("u%03d.wav;%s;%.4f" % (i, text, d))
    path = os.path.join(root, "train.csv")
    open(path, "w", encoding="utf-8").write("\n".join(rows) + "\n")
    return path


@pytest.mark.parametrize("target", ["train_bucket", "train_batch"])
def test_csv_wav_corpus_to_train_step(tmp_path, target):
    """asr/train.py's loop in miniature: CSV + WAV files -> input_fn_generator (GPU featuriser) -> train_step, for the
    bucketed and the SortaGrad (file order) epochs; the loss of a repeated epoch goes down."""
    csv_path = _write_corpus(str(tmp_path))
    flags = types.SimpleNamespace(train_csv=csv_path, dev_csv=csv_path, test_csv=csv_path, corpus_dir=str(tmp_path / "corpus"),
                                  batch_size=4, num_buckets=4, feature_type="mfcc", feature_normalization="local")
    cfg = CFG.replace(learning_rate=3e-4, compute="bf16x3", num_units_rnn=128)
    model = CTCModel(cfg, seed=1)
    input_fn = ip.input_fn_generator(target, flags, seed=1)
    epoch_loss = []
    for epoch in range(3):
        losses, seen = [], 0
        for features, label_encoded in input_fn():
            assert features["spectrogram"].is_cuda and features["spectrogram"].shape[2] == 80
            loss = model.train_step(features["spectrogram"], features["spectrogram_length"], label_encoded)
            losses.append(float(loss)); seen += len(features["label_plaintext"])
        model.check_step()
        assert seen == (25 if target == "train_bucket" else 24)      # last CSV row dropped; train_batch drops the remainder
        epoch_loss.append(float(np.mean(losses)))
    assert epoch_loss[-1] < epoch_loss[0], epoch_loss
    logits, sl = model.inference_fn(features["spectrogram"], features["spectrogram_length"], training=False)
    decoded, plaintext, _ = model.decode_fn(logits, sl, features["label_plaintext"], decoder="greedy")
    assert len(plaintext) == len(features["label_plaintext"])


def test_two_host_threads_with_their_own_contexts_and_streams():
    """ctcasr_handle_t: two host threads, each bound to its own context (own split-operand arena) and its own stream,
    run the tcgen05 dense layer at the same time; both get the bits the main thread gets alone."""
    import ctypes
    import threading
    lib = _lib.load()
    rng = np.random.default_rng(31)
    M, K, N = 4096, 512, 1024
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).cuda()
    ws = [torch.from_numpy((rng.standard_normal((K, N)) * 0.05).astype(np.float32)).cuda() for _ in range(2)]
    b = torch.zeros(N, device="cuda")
    C = _lib.COMPUTE_ID["bf16x3"]
    want = [ops.dense_fwd(x, w, b, act=0, compute=C).clone() for w in ws]
    torch.cuda.synchronize()
    got, errs = [None, None], []

    def worker(i):
        try:
            h = ctypes.c_void_p()
            assert lib.ctcasr_create(ctypes.byref(h)) == 0 and lib.ctcasr_use(h) == 0
            arena = torch.empty((64 << 20) + 1024, dtype=torch.uint8, device="cuda")
            base = (arena.data_ptr() + 1023) // 1024 * 1024
            assert lib.ctcasr_set_scratch(ctypes.c_void_p(base), 64 << 20) == 0
            stream = torch.cuda.Stream()
            y = torch.empty(M, N, device="cuda")
            with torch.cuda.stream(stream):
                for _ in range(20):
                    rc = lib.ctcasr_dense_fwd(_lib.ptr(x), _lib.ptr(ws[i]), _lib.ptr(b), _lib.ptr(y), M, K, N, 0, 20.0, 0.0, 0, C,
                                              ctypes.c_void_p(stream.cuda_stream))
                    assert rc == 0, lib.ctcasr_last_error()
            stream.synchronize()
            got[i] = y
            assert lib.ctcasr_use(None) == 0 and lib.ctcasr_destroy(h) == 0
        except Exception as e:  # noqa
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
