"""Host-side logic of the bucketing input pipeline (ctc_asr_b200/input_pipeline.py) against the reference's
rules (asr/input_functions.py:22-153, asr/util/csv_helper.py:9-38) on a small synthetic corpus in the
reference's on-disk format (';' CSV + 16 kHz mono WAV).  The featuriser is injected (numpy oracle), so no GPU."""
import os
import types
import wave

import numpy as np
import pytest

from ctc_asr_b200 import input_pipeline as ip
from oracle import features_ref


def _corpus(tmp_path, n=41, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(tmp_path / "corpus")
    durations = np.sort(rng.uniform(0.7, 3.0, n))            # the reference's CSVs are sorted by length
    rows = ["path;label;length"]
    for i, d in enumerate(durations):
        pcm = (rng.standard_normal(int(d * 16000)) * 2000).astype("<i2")
        with wave.open(str(tmp_path / "corpus" / ("u%03d.wav" % i)), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        text = "".join(rng.choice(list("abc de"), size=rng.integers(2, 9)))
        rows.append("u%03d.wav;%s;%.4f" % (i, text.strip() or "a", d))
    csv_path = tmp_path / "train.csv"
    csv_path.write_text("\n".join(rows) + "\n", encoding="utf-8")
    return str(csv_path), durations


def test_bucket_boundaries_follow_csv_helper(tmp_path):
    csv_path, durations = _corpus(tmp_path)
    lengths = [int(float("%.4f" % d) / 0.010) for d in durations]
    step = len(lengths) // 8
    want = sorted(set(lengths[i] for i in range(step, len(lengths), step)))
    assert ip.get_bucket_boundaries(csv_path, 8) == want
    assert ip.bucket_of(want[0] - 1, want) == 0 and ip.bucket_of(want[0], want) == 1      # [min, b0) | [b0, b1) ...
    assert ip.bucket_of(10 ** 9, want) == len(want)
    with pytest.raises(ValueError):
        ip.get_bucket_boundaries(csv_path, 1000)


def test_bucketed_epoch_covers_every_row_but_the_last_once(tmp_path):
    csv_path, durations = _corpus(tmp_path)
    batches = list(ip.plan_batches(csv_path, batch_size=4, use_buckets=True, num_buckets=8, seed=3))
    names = [r["path"] for b in batches for r in b]
    assert sorted(names) == ["u%03d.wav" % i for i in range(40)]                # u040 (last CSV row) is dropped, like the reference
    bounds = ip.get_bucket_boundaries(csv_path, 8)
    full = [b for b in batches if len(b) == 4]
    assert len(full) >= 5 and all(len(b) <= 4 for b in batches)
    for b in batches:                                                           # one bucket per batch
        assert len({ip.bucket_of(int(float(r["length"]) / 0.010), bounds) for r in b}) == 1
    again = list(ip.plan_batches(csv_path, 4, True, 8, seed=3))
    assert [[r["path"] for r in b] for b in again] == [[r["path"] for r in b] for b in batches]     # seeded order
    other = list(ip.plan_batches(csv_path, 4, True, 8, seed=4))
    assert [[r["path"] for r in b] for b in other] != [[r["path"] for r in b] for b in batches]
    # train_batch: file order, remainder dropped
    plain = list(ip.plan_batches(csv_path, 6, use_buckets=False))
    assert [r["path"] for b in plain for r in b] == ["u%03d.wav" % i for i in range(36)]


def test_input_fn_yields_padded_batches_in_the_reference_layout(tmp_path):
    csv_path, _ = _corpus(tmp_path)
    flags = types.SimpleNamespace(train_csv=csv_path, dev_csv=csv_path, test_csv=csv_path, corpus_dir=str(tmp_path / "corpus"),
                                  batch_size=4, num_buckets=8)

    def featurizer(clips):                                   # numpy oracle instead of the GPU kernel
        feats = [features_ref.load_sample(c, "mfcc", "local")[0] for c in clips]
        tmax = max(f.shape[0] for f in feats)
        out = np.zeros((len(feats), tmax, 80), np.float32)
        for b, f in enumerate(feats):
            out[b, :f.shape[0]] = f
        return out, np.array([f.shape[0] for f in feats], np.int32)

    def read_wav(path):
        with wave.open(path, "rb") as w:
            return w.getframerate(), np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)

    with pytest.raises(ValueError):
        ip.input_fn_generator("validation", flags)
    seen = 0
    for features, (labels, label_len) in ip.input_fn_generator("dev", flags, featurizer, read_wav, seed=1)():
        x, n = features["spectrogram"], features["spectrogram_length"]
        B = len(features["label_plaintext"])
        assert x.shape == (B, int(n.max()), 80) and labels.shape[0] == B
        for b, text in enumerate(features["label_plaintext"]):
            assert label_len[b] == len(text)
            assert "".join(" abcdefghijklmnopqrstuvwxyz"[i - 1] for i in labels[b, :label_len[b]]) == text
            assert (labels[b, label_len[b]:] == 0).all() and (x[b, n[b]:] == 0).all()
        # elements of one bucket: frame counts within the bucket's span (10 ms units of the CSV length ~ frames)
        assert int(n.max()) - int(n.min()) <= 60
        seen += B
    assert seen == 40


def test_frame_counts_and_real_length_bucketing(tmp_path):
    """Buckets key on the real spectrogram length (WAV header), halved when every second frame is dropped; the
    pure-Python frame formula equals the C-ABI's ctcasr_feature_frames."""
    assert [ip.num_frames(n) for n in (401, 560, 561, 16000, 160000, 272000)] == [2, 2, 3, 99, 999, 1699]
    assert ip.num_frames(16000, True) == 50 and ip.num_frames(160000, True) == 500
    try:
        from ctc_asr_b200 import _lib
        lib = _lib.load()
        for n in (401, 999, 16000, 123457, 272000):
            assert lib.ctcasr_feature_frames(n, 16000) == ip.num_frames(n)
    except _lib.CtcAsrError:
        pass
    csv_path, durations = _corpus(tmp_path)
    frames_of = ip.frames_of_wav(str(tmp_path / "corpus"))
    rows = ip._read_rows(csv_path)[1:-1]
    for r, d in zip(rows, durations):
        assert frames_of(r) == ip.num_frames(int(d * 16000))
        assert 0 <= int(float(r["length"]) / 0.010) - frames_of(r) <= 2          # the CSV length over-counts by 1-2 frames
    bounds = ip.get_bucket_boundaries(csv_path, 8)
    for b in ip.plan_batches(csv_path, 4, True, 8, seed=3, frames_of=frames_of):
        assert len({ip.bucket_of(frames_of(r), bounds) for r in b}) == 1
    half = ip.frames_of_wav(str(tmp_path / "corpus"), drop_every_second_frame=True)
    assert all(half(r) == (frames_of(r) + 1) // 2 for r in rows)
