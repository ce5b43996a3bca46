"""Accuracy of the tcgen05 GEMMs against fp64 as the contraction grows (run on the GPU box).

The weight-gradient products of the path contract over every frame of the batch (K = T x B = 32,000 at cfg2, 256,000 for the
single-GPU side of the 8-GPU correctness check).  The tensor core adds each MMA's products into the fp32 accumulator in
tensor memory; this probe measures how the error of one long accumulation chain grows with K, for random-sign terms (the
weight gradients) and for all-positive terms (the worst case of a biased accumulator), in every bf16 compute mode, against
torch fp64 on the same device (a checker, not a product path).  Output: one JSON object (profiles/r2_accum_error.json)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ctc_asr_b200 import _lib, ops  # noqa: E402


def err(c, want):
    d = (c.double() - want)
    return {"max_rel_to_max": float(d.abs().max() / want.abs().max()),
            "rms_rel_to_rms": float(d.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()),
            "signed_bias": float((d * want.sign()).mean() / want.abs().mean())}


def main():
    torch.manual_seed(0)
    M, N = 2048, 4096                        # 16 x 16 tiles of 128 x 256: no split over K, one accumulator per tile
    out = {"M": M, "N": N, "layout": "C[M,N] = A[K,M]^T B[K,N] (the weight-gradient orientation)", "cases": []}
    for K in (2048, 8000, 32000, 128000, 256000):
        for kind in ("random_sign", "positive"):
            a = torch.randn(K, M, device="cuda")
            b = torch.randn(K, N, device="cuda")
            if kind == "positive":
                a.abs_(); b.abs_()
            want = a.double().t() @ b.double()
            row = {"K": K, "terms": kind}
            for name in ("bf16x3", "bf16", "tf32"):
                c = ops.gemm(a, b, ta=True, compute=_lib.COMPUTE_ID[name])
                row[name] = err(c, want)
            # the same contraction as 8 partial sums added in fp32 (what a split over K would give)
            parts = [ops.gemm(a[i * K // 8:(i + 1) * K // 8], b[i * K // 8:(i + 1) * K // 8], ta=True,
                              compute=_lib.COMPUTE_ID["bf16x3"]) for i in range(8)]
            row["bf16x3_8_partial_sums"] = err(torch.stack(parts).sum(0), want)
            row["torch_fp32_matmul"] = err((a.t() @ b), want)
            out["cases"].append(row)
            print(json.dumps(row), flush=True)
            del a, b, want
    json.dump(out, open("gpurun_out/r2_accum_error.json", "w"), indent=1)


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    main()
