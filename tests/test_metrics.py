"""Host-side error metrics with the reference's function names (asr/util/metrics.py) and the oracle's
edit distance; the GPU kernel is checked bit-exactly against the oracle in test_gpu_parity.py."""
import numpy as np

from ctc_asr_b200 import metrics
from oracle import ref


def test_levenshtein_known_values():
    assert metrics.levenshtein("kitten", "sitting") == 3
    assert metrics.levenshtein("", "abc") == 3
    assert metrics.levenshtein("abc", "abc") == 0
    assert metrics.levenshtein("the cat sat".split(), "the cat sat on the mat".split()) == 3


def test_wer_and_batch():
    assert abs(float(metrics.wer("the cat sat on the mat", "the cat sit on mat")) - 2 / 6) < 1e-6
    rates, mean = metrics.wer_batch(["a b c", "hello world"], ["a b c", "hello"])
    np.testing.assert_allclose(rates, [0.0, 0.5])
    assert abs(float(mean) - 0.25) < 1e-6


def test_dense_to_text_and_summary():
    dec = [[9, 6, 13, 13, 16], [1, 2, 0, 0]]          # 'hello', ' a' (0 = padding decodes to '')
    texts, summary = metrics.dense_to_text(dec, [b"hello", b"a"])
    assert texts.tolist() == ["hello", " a"] and summary.shape == (2, 2) and summary[1].tolist() == ["hello", "a"]
    _, summary = metrics.dense_to_text(dec, [])
    assert summary[1].tolist() == ["n/a", "n/a"]


def test_oracle_edit_distance_matches_python_levenshtein():
    rng = np.random.default_rng(0)
    B = 12
    hl, tl = rng.integers(0, 30, B).astype(np.int32), rng.integers(0, 30, B).astype(np.int32)
    tl[0] = 0; hl[0] = 0; tl[1] = 0
    hyp, truth = rng.integers(1, 6, (B, 30)).astype(np.int32), rng.integers(1, 6, (B, 30)).astype(np.int32)
    d = ref.edit_distance(hyp, hl, truth, tl, normalize=False)
    for b in range(B):
        assert d[b] == metrics.levenshtein(hyp[b, :hl[b]].tolist(), truth[b, :tl[b]].tolist())
    dn = ref.edit_distance(hyp, hl, truth, tl, normalize=True)
    assert dn[0] == 0.0 and (np.isinf(dn[1]) or hl[1] == 0)
