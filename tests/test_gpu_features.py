"""GPU parity of the feature extraction (load_sample, asr/input_functions.py:156-349)
against the float64 numpy restatement of python_speech_features (oracle/features_ref.py).
Tolerance: 1e-3 of the largest feature magnitude (fp32 FFT against float64), frame counts exact."""
import os
import tempfile
import wave

import numpy as np
import pytest
import torch

from ctc_asr_b200 import features
from oracle import features_ref as fr

pytestmark = pytest.mark.gpu


def _speechlike(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    sig = 3000 * np.sin(2 * np.pi * (180 + 40 * seed) * t) + 1500 * np.sin(2 * np.pi * 1330 * t + 1.0) \
        + 700 * np.sin(2 * np.pi * 3100 * t * (1 + 0.1 * t)) + 300 * rng.standard_normal(n)
    sig *= 0.5 + 0.5 * np.sin(2 * np.pi * 3 * t) ** 2
    return np.clip(np.round(sig), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("feature_type", ["mfcc", "mel"])
@pytest.mark.parametrize("norm", ["none", "local", "local_scalar"])
@pytest.mark.parametrize("drop", [False, True])
def test_featurize_batch_vs_oracle(feature_type, norm, drop):
    lens = [16000, 401, 11237, 52000, 8000]                  # 1 s, the shortest legal clip, odd lengths, 3.25 s
    audio = [_speechlike(n, i) for i, n in enumerate(lens)]
    seq, frames = features.featurize(audio, feature_type, norm, drop_every_second_frame=drop)
    torch.cuda.synchronize()
    seq, frames = seq.cpu().numpy(), frames.cpu().numpy()
    tmax = max(fr.num_frames(n) for n in lens)
    assert seq.shape == (len(lens), (tmax + 1) // 2 if drop else tmax, 80) and seq.dtype == np.float32
    for b, a in enumerate(audio):
        want, n = fr.load_sample(a, feature_type, norm, drop_every_second_frame=drop)
        assert frames[b] == n
        got = seq[b, :n]
        if norm == "none" or n > 2:
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err < 1e-3, (b, err)
        assert (seq[b, n:] == 0).all()                       # padded_batch fill


def test_load_sample_reads_a_wav_file_like_the_reference():
    a = _speechlike(24000, 3)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "utt.wav")
        with wave.open(path, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(a.astype("<i2").tobytes())
        x, n = features.load_sample(path)                    # FLAGS defaults: 'mfcc', 'local'
        want, wn = fr.load_sample(a, "mfcc", "local")
        assert int(n) == wn == 149 and tuple(x.shape) == (149, 80)
        assert np.abs(x.cpu().numpy() - want).max() < 1e-3 * np.abs(want).max()
        with pytest.raises(ValueError):
            features.load_sample(os.path.join(d, "missing.wav"))
        with wave.open(path, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000); w.writeframes(a.astype("<i2").tobytes())
        with pytest.raises(RuntimeError):
            features.load_sample(path)                       # "Sampling rate is 8,000, expected 16,000."
    with pytest.raises(RuntimeError):
        features.featurize([a[:400]])                        # "Sample length 400 to short"
    with pytest.raises(ValueError):
        features.featurize([a], feature_type="spectrogram")


def test_features_feed_the_model_at_cfg2_size():
    """32 utterances of 10 s: int16 PCM in, [32, 999, 80] standardised MFCCs out, straight into inference_fn."""
    from ctc_asr_b200.model import CTCModel
    from ctc_asr_b200.params import ModelConfig
    audio = [_speechlike(160000, i) for i in range(32)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seq, frames = features.featurize(audio)
    e0.record()
    seq, frames = features.featurize(audio)
    e1.record()
    torch.cuda.synchronize()
    print("featurize 32 x 10 s (incl. pinned staging + H2D of the PCM): %.2f ms" % e0.elapsed_time(e1))
    assert tuple(seq.shape) == (32, 999, 80) and int(frames.min()) == 999
    assert bool(torch.isfinite(seq).all())
    assert float(seq.mean(1).abs().max()) < 1e-3 and float((seq.std(1, unbiased=False) - 1).abs().max()) < 1e-3
    cfg = ModelConfig(used_model="ds1", num_layers_dense=3, num_units_dense=128, num_layers_rnn=1, num_units_rnn=64, rnn_cell="lstm",
                      cudnn=False, compute="fp32")
    model = CTCModel(cfg, seed=1)
    logits, sl = model.inference_fn(seq, frames, training=False)
    assert tuple(logits.shape) == (999, 32, 29) and bool(torch.isfinite(logits).all())
