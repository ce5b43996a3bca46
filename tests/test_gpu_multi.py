"""Data-parallel correctness on real GPUs: needs a box with >= 2 GPUs (`gpurun --gpus 2`), skipped
otherwise.  Launches `bench.py --check` under torchrun with one rank per GPU over NCCL: the all-reduced gradient
(bucketed + overlapped, and the single all-reduce) equals the 1-GPU gradient of the concatenated global batch to 1e-5,
and after 5 training steps the parameters are bit-identical on every rank."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("cell,compute,units", [("lstm", "bf16x3", 512), ("rnn_relu", "bf16x3", 512), ("lstm", "bf16", 256)])
def test_all_reduced_gradient_equals_single_gpu_and_ranks_stay_identical(cell, compute, units):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--check", "--units", str(units),
           "--frames", "120", "--batch", "8", "--cell", cell, "--compute", compute]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    print(res)
    assert res["ok"] and res["params_bit_identical_after_5_steps"] and res["n_gpus"] == n
    for v in res["gradient_vs_single_gpu"].values():
        assert v["grad_rel_err"] < 1e-5 and v["loss_rel_err"] < 1e-5
