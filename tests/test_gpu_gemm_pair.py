"""The CTA-pair tcgen05 GEMM (`gemm_tc_pair_kernel`, cta_group::2: one 256 x 256 tile per two SMs) behind
ctcasr_dense_fwd / ctcasr_dense_bwd: all three operand orientations (forward x w, dgrad dz w^T, wgrad x^T dz) at a shape
that takes the pair kernel (N a multiple of 256, at least 74 pair tiles) with a ragged last row tile, against fp64
numpy products of the same operands, and against the single-CTA kernel (a narrow shape of the same data takes it).

Tolerances: compute='bf16x3' 3e-5 of max (fp32-level arithmetic); compute='bf16' 1e-5 against products of operands
rounded to bfloat16 the way the GPU rounds them (round to nearest even) — the arithmetic of BASELINE cfg3."""
import numpy as np
import pytest
import torch

from ctc_asr_b200 import _lib, ops

pytestmark = pytest.mark.gpu

BF16X3 = _lib.COMPUTE_ID["bf16x3"]
BF16 = _lib.COMPUTE_ID["bf16"]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda().contiguous()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


@pytest.mark.parametrize("compute", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M", [2000, 2304])
def test_pair_gemm_three_orientations(compute, M):
    rng = np.random.default_rng(21)
    K, N = 2560, 2816                       # forward: 8|9 x 11 pair tiles; dgrad 8|9 x 10; wgrad 10 x 11
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    cid = _lib.COMPUTE_ID[compute]
    before = ops.launch_count() if hasattr(ops, "launch_count") else None
    y = ops.dense_fwd(dev(x), dev(w), dev(b), act=0, compute=cid)
    dw, db, dx = torch.empty(K, N).cuda(), torch.empty(N).cuda(), torch.empty(M, K).cuda()
    dyd = dev(dy)
    ops.dense_bwd(dev(x), dev(w), y, dyd, dw, db, dx=dx, act=0, compute=cid)
    torch.cuda.synchronize()
    r = bf16_round if compute == "bf16" else (lambda a: np.asarray(a, np.float64))
    oy = r(x) @ r(w) + b.astype(np.float64)
    odx = r(dy) @ r(w).T
    odw = r(x).T @ r(dy)
    tol = 3e-5 if compute == "bf16x3" else 1e-5      # three bf16 products: 2^-16 per operand pair
    assert rel_err(y.cpu().numpy(), oy) < tol
    assert rel_err(dx.cpu().numpy(), odx) < tol
    assert rel_err(dw.cpu().numpy(), odw) < tol
    assert rel_err(db.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5
    assert before is None or ops.launch_count() > before


def test_pair_and_single_cta_kernels_agree_on_shared_columns():
    """The first 256 output columns computed inside a wide product (pair kernel) and as a narrow product of their own
    (N = 256: 8 pair tiles < 74, single-CTA kernel) accumulate the same k-blocks in the same order: equal to fp32
    rounding (bit-identical if the M = 256 and M = 128 instructions add in the same internal order; printed)."""
    rng = np.random.default_rng(22)
    M, K, N = 2048, 1024, 2816
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
    b = np.zeros(N, np.float32)
    for cid in (BF16X3, BF16):
        wide = ops.dense_fwd(dev(x), dev(w), dev(b), act=0, compute=cid)
        narrow = ops.dense_fwd(dev(x), dev(w[:, :256]), dev(b[:256]), act=0, compute=cid)
        torch.cuda.synchronize()
        print("compute %d: pair vs single-CTA kernel bit-identical: %s" % (cid, torch.equal(wide[:, :256], narrow)))
        assert rel_err(wide[:, :256].cpu().numpy(), narrow.cpu().numpy()) < 2e-6


@pytest.mark.parametrize("compute", ["bf16x3", "bf16"])
def test_narrow_linear_layer_on_the_tensor_cores(compute):
    """The 29-class logits layer over many rows (asr/model.py:229-232): operands widened to 64 columns, three tcgen05
    products in the fp32-level bf16x3 arithmetic in BOTH bf16 compute modes (the oracle of compute='bf16' keeps this layer
    exact), results narrowed back: forward, weight, bias and input gradients against fp64 numpy."""
    rng = np.random.default_rng(41)
    M, K, N = 4096, 256, 29
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) * 0.1).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    cid = _lib.COMPUTE_ID[compute]
    y = ops.dense_fwd(dev(x), dev(w), dev(b), act=0, compute=cid)
    dw, db, dx = torch.full((K, N), float("nan")).cuda(), torch.empty(N).cuda(), torch.full((M, K), float("nan")).cuda()
    ops.dense_bwd(dev(x), dev(w), y, dev(dy), dw, db, dx=dx, act=0, compute=cid)
    torch.cuda.synchronize()
    x64, w64, dy64 = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)
    assert rel_err(y.cpu().numpy(), x64 @ w64 + b) < 3e-5
    assert rel_err(dw.cpu().numpy(), x64.T @ dy64) < 3e-5
    assert rel_err(dx.cpu().numpy(), dy64 @ w64.T) < 3e-5
    assert rel_err(db.cpu().numpy(), dy64.sum(0)) < 1e-5
