"""numpy (float64) restatement of the reference's feature extraction — TEST INFRASTRUCTURE ONLY.

The reference computes its network inputs in `load_sample` (asr/input_functions.py:156-262) with the
third-party package python_speech_features ("psf", requirements.txt; un-vendored, not installable
offline): `psf.mfcc(..., numcep=40, nfilt=80, nfft=1024, lowfreq=64, highfreq=8000, preemph=0.97,
ceplifter=22, appendEnergy=True)` + `psf.delta(mfcc, 2)` (asr/input_functions.py:264-294) or
`psf.logfbank(...)` (:297-318), then `__feature_normalization` (:321-349).  "parity unpinned": psf's
published algorithm (python_speech_features 0.6, base.py / sigproc.py [RECALL]) is restated below and
pinned in tests/test_oracle_features.py on numpy.fft / scipy.fft.dct identities and analytic signals.
"""
import math

import numpy as np

EPS = np.finfo(float).eps


def _round_half_up(x):
    return int(math.floor(x + 0.5))


def num_frames(n_samples, frame_len=400, frame_step=160):
    """sigproc.framesig: 1 frame if the signal fits, else 1 + ceil((n - len) / step)."""
    return 1 if n_samples <= frame_len else 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def filterbank_bins(nfilt=80, nfft=1024, samplerate=16000, lowfreq=64., highfreq=8000.):
    """psf.get_filterbanks: FFT-bin edges of the nfilt triangular filters (nfilt + 2 integers)."""
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate).astype(np.int64)


def get_filterbanks(nfilt=80, nfft=1024, samplerate=16000, lowfreq=64., highfreq=8000.):
    b = filterbank_bins(nfilt, nfft, samplerate, lowfreq, highfreq)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def fbank(signal, samplerate=16000, winlen=0.025, winstep=0.010, nfilt=80, nfft=1024, lowfreq=64., highfreq=8000.,
          preemph=0.97):
    """psf.fbank with the default rectangular window: (filterbank energies [T,nfilt], frame energy [T])."""
    signal = np.asarray(signal)
    sig = np.append(signal[0], signal[1:] - preemph * signal[:-1]).astype(np.float64)      # sigproc.preemphasis
    flen, fstep = _round_half_up(winlen * samplerate), _round_half_up(winstep * samplerate)
    nf = num_frames(len(sig), flen, fstep)
    padded = np.concatenate([sig, np.zeros((nf - 1) * fstep + flen - len(sig))])
    idx = np.arange(flen)[None, :] + (np.arange(nf) * fstep)[:, None]
    frames = padded[idx]
    pspec = 1.0 / nfft * np.square(np.absolute(np.fft.rfft(frames, nfft)))                 # sigproc.powspec
    energy = pspec.sum(1)
    energy = np.where(energy == 0, EPS, energy)
    feat = pspec @ get_filterbanks(nfilt, nfft, samplerate, lowfreq, highfreq).T
    feat = np.where(feat == 0, EPS, feat)
    return feat, energy


def logfbank(signal, **kw):
    return np.log(fbank(signal, **kw)[0])


def _dct2_ortho(x):
    """scipy.fftpack.dct(x, type=2, axis=1, norm='ortho') written out."""
    N = x.shape[1]
    n = np.arange(N)
    basis = np.cos(np.pi * np.outer(n, 2 * n + 1) / (2.0 * N))                             # [k, n]
    y = 2.0 * x @ basis.T
    y[:, 0] *= math.sqrt(1.0 / (4 * N))
    y[:, 1:] *= math.sqrt(1.0 / (2 * N))
    return y


def mfcc(signal, numcep=40, ceplifter=22, append_energy=True, **kw):
    feat, energy = fbank(signal, **kw)
    feat = _dct2_ortho(np.log(feat))[:, :numcep]
    if ceplifter > 0:
        feat = (1 + (ceplifter / 2.) * np.sin(np.pi * np.arange(numcep) / ceplifter)) * feat
    if append_energy:
        feat[:, 0] = np.log(energy)
    return feat


def delta(feat, N=2):
    """psf.delta: regression over +-N frames with edge padding."""
    den = 2 * sum(i ** 2 for i in range(1, N + 1))
    padded = np.pad(feat, ((N, N), (0, 0)), mode="edge")
    out = np.empty_like(feat)
    for t in range(len(feat)):
        out[t] = np.dot(np.arange(-N, N + 1), padded[t:t + 2 * N + 1]) / den
    return out


def normalize(features, method):
    """__feature_normalization (asr/input_functions.py:321-349)."""
    if method == "none":
        return features
    if method == "local":
        return (features - np.mean(features, axis=0)) / np.std(features, axis=0)
    if method == "local_scalar":
        return (features - np.mean(features)) / np.std(features)
    raise ValueError("Invalid normalization method.")


def load_sample(audio_data, feature_type="mfcc", feature_normalization="local", drop_every_second_frame=False,
                num_features=80, samplerate=16000):
    """load_sample (asr/input_functions.py:156-262) from the decoded int16 samples on: -> ([T, 80] float32, T)."""
    if len(audio_data) < 401:
        raise RuntimeError("Sample length {:,d} to short".format(len(audio_data)))
    kw = dict(samplerate=samplerate, nfilt=num_features, nfft=1024, lowfreq=64., highfreq=samplerate / 2.)
    if feature_type == "mfcc":
        m = mfcc(audio_data, numcep=num_features // 2, **kw)
        sample = np.concatenate([m, delta(m, 2)], axis=1)
    elif feature_type == "mel":
        sample = logfbank(audio_data, **kw)
    else:
        raise ValueError("Unsupported feature type")
    sample = sample.astype(np.float32)
    if drop_every_second_frame:
        sample = sample[::2, :]
    return normalize(sample, feature_normalization), np.int32(sample.shape[0])
