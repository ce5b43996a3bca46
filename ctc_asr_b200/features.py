"""Feature extraction on the GPU behind the reference's `load_sample` name and options.

Reference: asr/input_functions.py:156-349 (`load_sample`, `__mfcc`, `__mel`, `__feature_normalization`)
over python_speech_features.  Reading and decoding the WAV file is host I/O (stdlib `wave` here,
scipy.io.wavfile in the reference); everything from the int16 samples on runs in libctcasr.so
(csrc/featurize.cu).  `featurize` is the batched form the training loop wants: a list of utterances
in, the padded [B, T, 80] `sequences` tensor and the frame counts out — already on the device, so the
input pipeline's padded_batch + host-to-device copy of float features (10 MB per cfg2 batch) shrinks to
the int16 samples."""
import ctypes
import os
import wave

import numpy as np
import torch

from . import _lib
from .params import FLAGS, NUM_FEATURES

FEATURE_TYPES = {"mel": 0, "mfcc": 1}                       # asr/input_functions.py:193
NORMALIZATIONS = {"none": 0, "local": 1, "local_scalar": 2}  # :194
SAMPLING_RATE = 16000                                       # asr/params.py:102


_pinned = {}


def _pinned_i16(n):
    """Grow-only pinned staging buffer for the PCM (allocating pinned memory costs milliseconds per call)."""
    ev = _pinned.get("copied")
    if ev is not None:
        ev.synchronize()                                  # the previous call's H2D copy has read the buffer
    buf = _pinned.get("pcm")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 20), dtype=torch.int16).pin_memory()
        _pinned["pcm"] = buf
    return buf[:n]


def num_frames(num_samples, sampling_rate=SAMPLING_RATE):
    return _lib.load().ctcasr_feature_frames(int(num_samples), int(sampling_rate))


def featurize(audio, feature_type="mfcc", feature_normalization="local", drop_every_second_frame=False,
              sampling_rate=SAMPLING_RATE, num_features=NUM_FEATURES, device="cuda"):
    """audio: list of 1-D int16 arrays (decoded PCM).  -> (sequences [B, Tmax, num_features] float32 CUDA tensor,
    zero past each utterance's frames, seq_length [B] int32 CUDA tensor)."""
    if feature_type not in FEATURE_TYPES:
        raise ValueError("Requested feature type of {} isn't supported.".format(feature_type))
    if feature_normalization not in NORMALIZATIONS:
        raise ValueError("Requested feature normalization method {} is invalid.".format(feature_normalization))
    lib = _lib.load()
    audio = [np.ascontiguousarray(a) for a in audio]
    for a in audio:
        if a.dtype != np.int16 or a.ndim != 1:
            raise ValueError("audio must be 1-D int16 PCM (what wavfile.read returns for 16-bit files)")
        if len(a) < 401:
            raise RuntimeError("Sample length {:,d} to short".format(len(a)))
    B = len(audio)
    lens = np.array([len(a) for a in audio], np.int32)
    nmax = int(lens.max())
    host = _pinned_i16(B * nmax).view(B, nmax)
    for b, a in enumerate(audio):
        host[b, :len(a)] = torch.from_numpy(a)
        host[b, len(a):] = 0
    dev_audio = host.to(device, non_blocking=True)
    _pinned["copied"] = torch.cuda.Event()
    _pinned["copied"].record()
    tl = num_frames(nmax, sampling_rate)
    tmax = (tl + 1) // 2 if drop_every_second_frame else tl
    out = torch.empty((B, tmax, num_features), dtype=torch.float32, device=device)
    frames = torch.empty(B, dtype=torch.int32, device=device)
    wsb = lib.ctcasr_featurize_workspace_bytes(B, nmax, sampling_rate, num_features)
    from .ops import workspace, _stream
    ws = workspace(wsb, out.device, "features")
    _lib.check(lib.ctcasr_featurize(_lib.ptr(dev_audio), B, nmax, lens.ctypes.data_as(ctypes.c_void_p),
                                    FEATURE_TYPES[feature_type], NORMALIZATIONS[feature_normalization],
                                    int(bool(drop_every_second_frame)), int(sampling_rate), int(num_features),
                                    _lib.ptr(out), tmax, _lib.ptr(frames), _lib.ptr(ws), ws.numel(), _stream()), "featurize")
    return out, frames


def read_wav(file_path):
    """(sampling_rate, int16 samples) of a mono 16-bit PCM WAV file, like scipy.io.wavfile.read."""
    with wave.open(file_path, "rb") as w:
        if w.getsampwidth() != 2 or w.getnchannels() != 1:
            raise ValueError('"{}" is not mono 16-bit PCM.'.format(file_path))
        return w.getframerate(), np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)


def load_sample(file_path, feature_type=None, feature_normalization=None):
    """asr/input_functions.py:156: -> (features [time, num_features] float32 CUDA tensor, length int32 tensor)."""
    feature_type = feature_type if feature_type is not None else "mfcc"                      # FLAGS.feature_type, asr/params.py:55
    feature_normalization = feature_normalization if feature_normalization is not None else "local"   # :57
    if type(file_path) is not str:
        file_path = str(file_path, "utf-8")
    if not os.path.isfile(file_path):
        raise ValueError('"{}" does not exist.'.format(file_path))
    sampling_rate, audio_data = read_wav(file_path)
    if len(audio_data) < 401:
        raise RuntimeError("Sample length {:,d} to short: {}".format(len(audio_data), file_path))
    if not sampling_rate == SAMPLING_RATE:
        raise RuntimeError("Sampling rate is {:,d}, expected {:,d}.".format(sampling_rate, SAMPLING_RATE))
    seq, n = featurize([audio_data], feature_type, feature_normalization)
    return seq[0], n[0]
