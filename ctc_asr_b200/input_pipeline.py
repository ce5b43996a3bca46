"""Bucketing input pipeline over the reference's on-disk corpus format.

Reference: `input_fn_generator` / `__input_generator` (asr/input_functions.py:22-153) and
`get_bucket_boundaries` (asr/util/csv_helper.py:9-38).  A corpus is a `;`-separated CSV with the header
`path;label;length` (asr/params.py:151-155; `path` relative to `corpus_dir`, `length` in seconds) and 16 kHz
mono 16-bit WAV files.  The reference decodes + featurises every file on the host inside a tf.data Python
generator and lets `bucket_by_sequence_length` form the batches; here the host only reads the PCM and decides
the batches (the frame count follows from the sample count in the WAV header, `frames_of_wav`), and each batch is
featurised in one call on the GPU (`features.featurize`), so what crosses PCIe is int16 samples.

Behaviour kept from the reference:
  * rows = `list(reader)[1:-1]`: the header and the LAST data row are dropped (asr/input_functions.py:134);
  * shuffling of the rows when buckets are used (`random.shuffle`, :138); bucket boundaries = every
    (N // num_buckets)-th length of the CSV (in file order), in units of 10 ms, de-duplicated and sorted;
  * `bucket_by_sequence_length(element_length_func=spectrogram_length, bucket_batch_sizes=[batch_size]*,
    pad_to_bucket_boundary=False)`: an element of length l goes to bucket `bisect_right(boundaries, l)`, a bucket
    emits a batch as soon as it holds `batch_size` elements, padded to the longest element of the batch, and the
    partly filled buckets are emitted at the end of the epoch;
  * without buckets (`train_batch`): file order, `padded_batch(batch_size, drop_remainder=True)`;
  * labels: `ctoi` per character (asr/input_functions.py:151), 0-padded to the longest label of the batch.
The tf.data shuffle buffer (`dataset.shuffle(FLAGS.shuffle_buffer_size)`, a second, windowed shuffle of an
already shuffled list) is not reproduced: its only effect is a different random order.
"""
import bisect
import csv
import os
import random

import numpy as np

from . import labels as _labels
from .params import WIN_LENGTH, WIN_STEP

CSV_HEADER_PATH, CSV_HEADER_LABEL, CSV_HEADER_LENGTH = "path", "label", "length"        # asr/params.py:151-153
CSV_FIELDNAMES = [CSV_HEADER_PATH, CSV_HEADER_LABEL, CSV_HEADER_LENGTH]
CSV_DELIMITER = ";"

TARGETS = {              # asr/input_functions.py:40-57: target -> (csv flag, use_buckets)
    "train_bucket": ("train_csv", True), "train_batch": ("train_csv", False), "dev": ("dev_csv", True),
    "test": ("test_csv", True),
}


def _read_rows(csv_path):
    if not (os.path.exists(csv_path) and os.path.isfile(csv_path)):
        raise AssertionError(csv_path)
    with open(csv_path, "r", encoding="utf-8") as fh:
        return list(csv.DictReader(fh, delimiter=CSV_DELIMITER, fieldnames=CSV_FIELDNAMES))


def get_bucket_boundaries(csv_path, num_buckets):
    """asr/util/csv_helper.py:9-38, including its use of the CSV's own row order."""
    data = _read_rows(csv_path)[1:]
    lengths = [int(float(d[CSV_HEADER_LENGTH]) / WIN_STEP) for d in data]
    step = len(lengths) // num_buckets
    if step < 1:
        raise ValueError("fewer rows ({}) than buckets ({})".format(len(lengths), num_buckets))
    return sorted(set(lengths[i] for i in range(step, len(lengths), step)))


def bucket_of(length, boundaries):
    """bucket_by_sequence_length: bucket i holds boundaries[i-1] <= length < boundaries[i]."""
    return bisect.bisect_right(boundaries, length)


def num_frames(num_samples, drop_every_second_frame=False, sampling_rate=16000):
    """Frames python_speech_features makes of `num_samples` samples (25 ms windows, 10 ms steps, asr/params.py:146-147,
    asr/input_functions.py:273-276): 1 + ceil((n - 400) / 160); halved (rounding up) when every second frame is dropped
    (asr/input_functions.py:239-241).  Same numbers as ctcasr_feature_frames (tests/test_input_pipeline.py)."""
    win, step = int(round(WIN_LENGTH * sampling_rate)), int(round(WIN_STEP * sampling_rate))
    n = 1 if num_samples <= win else 1 + -(-(num_samples - win) // step)
    return (n + 1) // 2 if drop_every_second_frame else n


def frames_of_wav(corpus_dir, drop_every_second_frame=False):
    """row -> spectrogram_length of the row's clip, from the WAV header alone (no decoding): what the reference's
    bucket_by_sequence_length keys on (element_length_func = spectrogram_length, asr/input_functions.py:91-93)."""
    import wave

    def frames_of(row):
        with wave.open(os.path.join(corpus_dir, row[CSV_HEADER_PATH]), "rb") as w:
            return num_frames(w.getnframes(), drop_every_second_frame, w.getframerate())
    return frames_of


def pad_labels(rows):
    lmax = max(1, max(len(r) for r in rows))
    out = np.zeros((len(rows), lmax), np.int32)
    for b, r in enumerate(rows):
        out[b, :len(r)] = r
    return out, np.array([len(r) for r in rows], np.int32)


def plan_batches(csv_path, batch_size, use_buckets, num_buckets=96, seed=None, frames_of=None):
    """The batches of one epoch as lists of CSV rows — everything the reference decides before any arithmetic.
    frames_of(row) -> spectrogram length of the row's clip: `frames_of_wav(corpus_dir)` gives the real frame count the
    reference buckets on; the default derives it from the CSV's `length` column (how the boundaries themselves are
    computed, 1-2 frames longer than the real count), for planning without the audio files."""
    rows = _read_rows(csv_path)[1:-1]                      # header and final row (asr/input_functions.py:134)
    if frames_of is None:
        frames_of = lambda r: int(float(r[CSV_HEADER_LENGTH]) / WIN_STEP)
    if not use_buckets:
        for i in range(0, len(rows) - batch_size + 1, batch_size):      # drop_remainder=True
            yield rows[i:i + batch_size]
        return
    boundaries = get_bucket_boundaries(csv_path, num_buckets)
    random.Random(seed).shuffle(rows)
    buckets = [[] for _ in range(len(boundaries) + 1)]
    for r in rows:
        b = buckets[bucket_of(frames_of(r), boundaries)]
        b.append(r)
        if len(b) == batch_size:
            yield list(b)
            b.clear()
    for b in buckets:                                      # end of input: the partly filled buckets
        if b:
            yield list(b)


def input_fn_generator(target, flags, featurizer=None, read_wav=None, seed=None):
    """-> a function that yields one epoch of (features, label_encoded) like the reference's input_fn
    (asr/input_functions.py:22-124): features = {'spectrogram' [B,T,80], 'spectrogram_length' [B],
    'label_plaintext' [B]}, label_encoded = (0-padded int32 [B,Lmax], lengths [B]).
    `flags` carries train_csv / dev_csv / test_csv, corpus_dir, batch_size, num_buckets (asr/params.py names);
    featurizer(list of int16 clips) -> (sequences, seq_length) defaults to the GPU featuriser."""
    if target not in TARGETS:
        raise ValueError('Invalid target: "{}"'.format(target))
    csv_flag, use_buckets = TARGETS[target]
    csv_path = getattr(flags, csv_flag)
    if featurizer is None or read_wav is None:
        from . import features as _features
        featurizer = featurizer or (lambda clips: _features.featurize(
            clips, getattr(flags, "feature_type", "mfcc"), getattr(flags, "feature_normalization", "local"),
            getattr(flags, "features_drop_every_second_frame", False)))
        read_wav = read_wav or _features.read_wav

    frames_of = frames_of_wav(flags.corpus_dir, getattr(flags, "features_drop_every_second_frame", False))

    def input_fn():
        for rows in plan_batches(csv_path, flags.batch_size, use_buckets, getattr(flags, "num_buckets", 96), seed, frames_of):
            clips = []
            for r in rows:
                rate, pcm = read_wav(os.path.join(flags.corpus_dir, r[CSV_HEADER_PATH]))
                if rate != 16000:
                    raise RuntimeError("Sampling rate is {:,d}, expected {:,d}.".format(rate, 16000))
                clips.append(pcm)
            spectrogram, spectrogram_length = featurizer(clips)
            texts = [r[CSV_HEADER_LABEL] for r in rows]
            label_encoded = pad_labels([[_labels.ctoi(c) for c in t] for t in texts])
            yield {"spectrogram": spectrogram, "spectrogram_length": spectrogram_length, "label_plaintext": texts}, label_encoded

    return input_fn
