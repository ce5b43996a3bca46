"""NCCL all-reduce bus bandwidth at the gradient sizes of cfg2 / cfg3 (739 MB fp32, 370 MB), BASELINE.md section 3.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_busbw.py
busbw = 2 (N-1)/N * bytes / time (the nccl-tests convention).  Rank 0 prints one JSON line."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "sizes": {}}
    for nbytes in (739254388, 369627194, 64 << 20, 8 << 20):
        buf = torch.ones(nbytes // 4, dtype=torch.float32, device="cuda")
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            dist.all_reduce(buf)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = float(ms.item()) * 1e-3
        out["sizes"][str(nbytes)] = {"ms": t * 1e3, "algbw_gbs": nbytes / t / 1e9, "busbw_gbs": 2.0 * (world - 1) / world * nbytes / t / 1e9}
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
