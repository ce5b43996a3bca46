"""Pins for the feature-extraction oracle (oracle/features_ref.py; psf is not installable offline, so
its published algorithm is checked piece by piece against numpy / scipy and analytic signals)."""
import numpy as np
import pytest
import scipy.fft

from oracle import features_ref as fr


def _speechlike(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    sig = 3000 * np.sin(2 * np.pi * 220 * t) + 1500 * np.sin(2 * np.pi * 1330 * t + 1.0) + 300 * rng.standard_normal(n)
    sig *= 0.5 + 0.5 * np.sin(2 * np.pi * 3 * t) ** 2
    return np.clip(np.round(sig), -32768, 32767).astype(np.int16)


def test_frame_count_and_shapes():
    assert fr.num_frames(400) == 1 and fr.num_frames(401) == 2 and fr.num_frames(16000) == 99
    assert fr.num_frames(160000) == 999                        # 10 s -> the 999 frames SURVEY.md §8 quotes
    a = _speechlike(16000)
    for ft in ("mel", "mfcc"):
        x, n = fr.load_sample(a, ft, "local")
        assert x.shape == (99, 80) and n == 99 and x.dtype == np.float32
        np.testing.assert_allclose(x.mean(0), 0, atol=1e-4)
        np.testing.assert_allclose(x.std(0), 1, rtol=1e-3)
    x, n = fr.load_sample(a, "mfcc", "none", drop_every_second_frame=True)
    assert x.shape == (50, 80) and n == 50
    with pytest.raises(RuntimeError):
        fr.load_sample(a[:400])


def test_filterbank_is_psf_shaped():
    fb = fr.get_filterbanks()
    b = fr.filterbank_bins()
    assert fb.shape == (80, 513) and b[0] == 4 and b[-1] == 512          # 64 Hz and Nyquist at nfft 1024
    assert (np.diff(b) >= 0).all()
    for j in (0, 17, 79):
        assert fb[j].max() == 1.0 and fb[j, b[j + 1]] == 1.0             # the falling edge overwrites the peak with 1
        assert (fb[j, :b[j]] == 0).all() and (fb[j, b[j + 2]:] == 0).all()


def test_dct_lifter_and_energy_slot():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((7, 80))
    np.testing.assert_allclose(fr._dct2_ortho(x), scipy.fft.dct(x, type=2, axis=1, norm="ortho"), atol=1e-12)
    a = _speechlike(8000)
    feat, energy = fr.fbank(a)
    m = fr.mfcc(a)
    np.testing.assert_allclose(m[:, 0], np.log(energy))                   # appendEnergy replaces c0
    want = scipy.fft.dct(np.log(feat), type=2, axis=1, norm="ortho")[:, :40] * (1 + 11 * np.sin(np.pi * np.arange(40) / 22))
    np.testing.assert_allclose(m[:, 1:], want[:, 1:], rtol=1e-9, atol=1e-11)


def test_parseval_and_pure_tone():
    a = _speechlike(4000)
    feat, energy = fr.fbank(a)
    sig = np.append(a[0], a[1:] - 0.97 * a[:-1].astype(np.float64))
    frame0 = sig[:400]
    # Parseval for the zero-padded 1024-point transform: sum_k |X_k|^2 over the full spectrum = N sum x^2;
    # the one-sided sum counts DC and Nyquist once and every other bin once of twice
    full = np.abs(np.fft.fft(frame0, 1024)) ** 2 / 1024
    assert abs(full.sum() - (frame0 ** 2).sum()) < 1e-6 * full.sum()
    assert abs(energy[0] - full[:513].sum()) < 1e-9 * energy[0]
    # a 1 kHz tone lands in the filter whose triangle covers bin 1000 / 16000 * 1025
    t = np.arange(8000) / 16000.0
    tone = np.round(8000 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16)
    lf = fr.logfbank(tone)
    b = fr.filterbank_bins()
    k = int(1000 * 1024 / 16000)
    hot = [j for j in range(80) if b[j] <= k < b[j + 2]]
    assert int(lf[10].argmax()) in hot


def test_delta_is_psf_regression():
    rng = np.random.default_rng(2)
    f = rng.standard_normal((9, 5))
    d = fr.delta(f, 2)
    t = 4
    np.testing.assert_allclose(d[t], (f[t + 1] - f[t - 1] + 2 * (f[t + 2] - f[t - 2])) / 10)
    np.testing.assert_allclose(d[0], (f[1] - f[0] + 2 * (f[2] - f[0])) / 10)        # edge padding
    ramp = np.arange(9.0)[:, None] * np.ones((1, 3))
    np.testing.assert_allclose(fr.delta(ramp, 2)[2:-2], 1.0)
