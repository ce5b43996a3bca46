"""Error metrics of the reference's evaluation path, same function names as asr/util/metrics.py:
`dense_to_text` (:9-47), `wer` (:52-75), `wer_batch` (:80-104), `levenshtein` (:115-141), plus
`edit_distance`, the GPU counterpart of `tf.edit_distance(decoded, labels)` (asr/model.py:338).
These are eval-time, host-side string metrics in the reference (run through tf.py_func); only the
label edit distance has a kernel (`ctcasr_edit_distance`)."""
import numpy as np

from . import labels as _labels
from .params import NP_FLOAT


def levenshtein(a, b):
    """Edit distance between two sequences (strings or lists of words); two-row dynamic programme
    over a numpy row, O(min(len)) memory."""
    if len(a) > len(b):
        a, b = b, a
    row = np.arange(len(a) + 1)
    for i, cb in enumerate(b, start=1):
        diag = row[:-1] + np.fromiter((ca != cb for ca in a), dtype=np.int64, count=len(a))
        new = np.minimum(row[1:] + 1, diag)                  # deletion / substitution
        out = np.empty_like(row)
        out[0] = i
        # insertions chain left to right: out[j] = min(new[j-1], out[j-1] + 1)
        for j in range(1, len(row)):
            out[j] = min(new[j - 1], out[j - 1] + 1)
        row = out
    return int(row[-1])


def wer(original, result):
    """Word error rate of one sentence pair: word-level edit distance / number of reference words."""
    ref_words, hyp_words = original.split(), result.split()
    return np.array(levenshtein(ref_words, hyp_words) / float(len(ref_words)), dtype=NP_FLOAT)


def wer_batch(originals, results):
    """-> (per-sample WER [batch], mean WER) like asr/util/metrics.py:80-104."""
    if len(originals) != len(results):
        raise AssertionError("batch sizes differ")
    rates = np.array([wer(o, r) for o, r in zip(originals, results)], dtype=NP_FLOAT)
    return rates, np.array(rates.mean() if len(rates) else 0.0, dtype=NP_FLOAT)


def dense_to_text(decoded, originals):
    """Integer label rows -> strings, and the [decoded; original] summary table ('n/a' when no
    originals are given), like asr/util/metrics.py:9-47."""
    texts = [_labels.ids_to_text(row) for row in decoded]
    if len(originals) > 0:
        origs = [o.decode("utf-8") if isinstance(o, bytes) else str(o) for o in originals]
    else:
        origs = ["n/a"] * len(texts)
    return np.array(texts, dtype=object), np.vstack([np.array(texts, dtype=object), np.array(origs, dtype=object)])


def edit_distance(hyp, hyp_len, truth, truth_len, normalize=True):
    """tf.edit_distance(decoded, labels) (asr/model.py:338) on padded int32 CUDA tensors
    hyp [B,Lh], truth [B,Lt] -> float32 [B] (distance / len(truth) when normalize)."""
    from . import ops
    return ops.edit_distance(hyp, hyp_len, truth, truth_len, normalize)
