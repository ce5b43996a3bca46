"""Chained accumulation of the tcgen05 GEMMs (gemm_tc.cu / conv_tc.cu) with MANY chunks per output tile.

At the shipped chunk length (<= 8192 contraction elements) only full-size products are chunked.  Here the GEMM, CTA-pair
GEMM, split-K, conv-layer and whole-path parity tests are run again in a child process with CTCASR_GEMM_CHAIN=256 (the
library reads it once per process): every contraction longer than 384 is then summed in chunks of <= 256 — up to 250
chunks per tile, partial last chunks, chunks inside split-K slices, the implicit conv weight gradient's position groups —
and must still meet the same tolerances against the same oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_with_256_element_chunks():
    if os.environ.get("CTCASR_GEMM_CHAIN") == "256":
        pytest.skip("already inside the child process")
    env = dict(os.environ, CTCASR_GEMM_CHAIN="256")
    sel = ("gemm or conv2d_layer or dense or implicit or whole_path_bf16x3 or small_3d2r2d or contraction or pair or narrow")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", sel,
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_gpu_gemm_pair.py"),
           os.path.join(ROOT, "tests", "test_gpu_conv.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail
    print(r.stdout.splitlines()[-1])
