"""Model of the tcgen05 accumulator (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

Measured on the B200 (tools/accum_probe.py -> profiles/r2_accum_error.json): the error of one accumulation chain of the
bf16x3 GEMM grows linearly with the contraction length and is biased towards zero, -2.4e-8 of the accumulator per MMA.
That is what a tensor core does that forms the 16 products of an MMA exactly and adds their sum to the fp32 accumulator
with TRUNCATION (round towards zero) instead of round-to-nearest: the mean loss of a truncation is ~0.36 ulp = 2.1e-8 .. 4.3e-8
of the value.  This module restates that hypothesis in numpy so that the CPU test suite can check it against the committed
measurement, and evaluates the remedy the GEMM kernels use (csrc/gemm_tc.cu, "chained accumulation": chunks of the
chain in separate accumulators, added in fp32 round-to-nearest).

No reference file is followed here: the reference delegates every product to TensorFlow (asr/model.py:176-216 and
asr/util/tf_contrib.py:50-61 call tf.layers.dense / the RNN cells), whose CPU kernels accumulate in fp32 round-to-nearest.
"""
import numpy as np

UMMA_K = 16          # contraction elements of one bf16 MMA instruction


def to_bf16(x):
    """float32 -> bfloat16 (round to nearest even), returned as float32."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split2(x):
    hi = to_bf16(x)
    lo = to_bf16(np.asarray(x, np.float32) - hi)
    return hi, lo


def rz32(x64):
    """float64 -> float32 with round towards zero."""
    r = np.asarray(x64, np.float64).astype(np.float32)
    over = np.abs(r.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(r, np.float32(0)), r).astype(np.float32)


def chain_bf16x3(a, b, chunk=0, truncate=True):
    """sum_k a[k, :] * b[k, :] per column, the way gemm_tc.cu issues it: per 32-element k-block the three products
    hi*hi, hi*lo, lo*hi, each as two MMAs of 16 elements; every MMA adds its exact 16-term sum to the fp32 accumulator
    (truncating if `truncate`).  chunk > 0: a new accumulator every `chunk` elements, the chunks added in fp32 RNE."""
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    K = a.shape[0]
    ah, al = split2(a); bh, bl = split2(b)
    prods = [(ah, bh), (ah, bl), (al, bh)]
    total = None
    acc = np.zeros(a.shape[1], np.float32)
    since = 0
    for k0 in range(0, K, 32):
        for pa, pb in prods:
            for j in range(k0, min(K, k0 + 32), UMMA_K):
                s = (pa[j:j + UMMA_K].astype(np.float64) * pb[j:j + UMMA_K].astype(np.float64)).sum(0)
                t = acc.astype(np.float64) + s
                acc = rz32(t) if truncate else t.astype(np.float32)
        since += 32
        if chunk and since >= chunk and k0 + 32 < K:
            total = acc if total is None else (total + acc).astype(np.float32)
            acc = np.zeros_like(acc)
            since = 0
    return acc if total is None else (total + acc).astype(np.float32)


def chain_error(K, cols=2048, chunk=0, truncate=True, seed=0):
    """max |error| / max |exact| of `cols` independent chains of length K with N(0,1) terms (the probe's metric)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((K, cols)).astype(np.float32)
    b = rng.standard_normal((K, cols)).astype(np.float32)
    want = (a.astype(np.float64) * b.astype(np.float64)).sum(0)
    got = chain_bf16x3(a, b, chunk=chunk, truncate=truncate).astype(np.float64)
    return float(np.abs(got - want).max() / np.abs(want).max()), float(((got - want) * np.sign(want)).mean() / np.abs(want).mean())
