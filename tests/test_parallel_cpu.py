"""World-size-2 gloo test (CPU) of the data-parallel host logic: equal shards, CTC gradient scaled by
1/global_batch, one all-reduce (sum) of the flat gradient == the single-process gradient of the
global mean loss (SURVEY.md §8e correctness test).  The per-shard arithmetic comes from the CPU oracle
(this is a test of the plumbing; the CUDA path is exercised by the gpu-marked tests and bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctc_asr_b200 import parallel, synthetic
from ctc_asr_b200.params import ModelConfig, gradient_buckets, param_offsets
from oracle import model_ref

CFG = ModelConfig(used_model="ds1", num_layers_dense=2, num_units_dense=16, num_layers_rnn=1, num_units_rnn=8, rnn_cell="lstm",
                  cudnn=False, dense_dropout_rate=0.0, num_features=6)


def _flat(cfg, grads):
    offs, n = param_offsets(cfg)
    flat = np.zeros(n)
    for k, (o, shape) in offs.items():
        flat[o:o + grads[k].size] = grads[k].ravel()
    return flat


def _worker(rank, world, port, x, sl, lab, ll, out, interleave=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = synthetic.init_params(CFG, seed=1, dtype=np.float64)
    xs, sls, labs, lls = parallel.shard_batch(x, sl, lab, ll, rank, world, interleave)
    gb = x.shape[0]
    loss, grads, _, _ = model_ref.loss_and_grads(CFG, params, xs, sls, labs, lls)
    # loss_and_grads averages over the SHARD; rescale to 1/global_batch like loss_fn(global_batch=gb)
    scale = xs.shape[0] / gb
    flat = torch.from_numpy(_flat(CFG, grads) * scale)
    parallel.allreduce_gradients(flat)
    mean_loss = parallel.allreduce_mean_loss(torch.tensor(loss * scale))
    if rank == 0:
        out["flat"], out["loss"] = flat.numpy().copy(), float(mean_loss)
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    x, sl, lab, ll = synthetic.fixed_batch(4, 14, 3, F=CFG.num_features, seed=5)
    sl[1], sl[3] = 11, 9
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, x, sl, lab, ll, out), nprocs=2, join=True)
    params = synthetic.init_params(CFG, seed=1, dtype=np.float64)
    loss, grads, _, _ = model_ref.loss_and_grads(CFG, params, x, sl, lab, ll)
    np.testing.assert_allclose(out["loss"], loss, rtol=1e-12)
    np.testing.assert_allclose(out["flat"], _flat(CFG, grads), atol=1e-12)


def test_interleaved_shards_of_a_bucketed_batch_balance_lengths_and_sum_to_the_same_gradient():
    """cfg4 across GPUs: a length-sorted (bucketed) batch cut into contiguous halves gives one rank all the long
    utterances; interleaved shards have the same length mix, and the all-reduced gradient is still exactly the
    single-process gradient of the global mean loss."""
    x, sl, lab, ll = synthetic.fixed_batch(6, 20, 3, F=CFG.num_features, seed=6)
    sl[:] = np.array([8, 10, 12, 15, 17, 20], np.int32)                 # sorted by duration, like a bucket
    for b in range(6):
        x[b, sl[b]:] = 0
    contiguous = [int(sl[parallel.shard_indices(6, r, 2)].sum()) for r in range(2)]
    interleaved = [int(sl[parallel.shard_indices(6, r, 2, interleave=True)].sum()) for r in range(2)]
    assert max(interleaved) - min(interleaved) < max(contiguous) - min(contiguous)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, x, sl, lab, ll, out, True), nprocs=2, join=True)
    params = synthetic.init_params(CFG, seed=1, dtype=np.float64)
    loss, grads, _, _ = model_ref.loss_and_grads(CFG, params, x, sl, lab, ll)
    np.testing.assert_allclose(out["loss"], loss, rtol=1e-12)
    np.testing.assert_allclose(out["flat"], _flat(CFG, grads), atol=1e-12)


def test_shard_bounds_reject_ragged_split():
    assert parallel.shard_bounds(256, 3, 8) == (96, 128)
    try:
        parallel.shard_bounds(30, 0, 4)
    except ValueError:
        return
    raise AssertionError("uneven shards must be rejected")


def _bucket_worker(rank, world, port, n, buckets, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(n, dtype=torch.float64) * (rank + 1)
    whole = g.clone()
    parallel.allreduce_gradients(whole)
    works = [parallel.allreduce_gradients(g[lo:hi], async_op=True) for lo, hi in buckets]       # views: reduced in place
    for w in works:
        w.wait()
    if rank == 0:
        out["same"] = bool(torch.equal(g, whole))
    dist.destroy_process_group()


def test_gradient_buckets_cover_the_buffer_and_reduce_like_one_allreduce():
    """The overlapped data-parallel step all-reduces `gradient_buckets()` one by one (asynchronously, in backward
    order): the buckets partition the flat buffer, for both front-ends, and the result is the single all-reduce."""
    for cfg in (CFG, CFG.replace(num_layers_rnn=3), ModelConfig(used_model="ds2", conv_filters=(8, 8, 64), num_units_dense=16,
                                                                num_layers_rnn=2, num_units_rnn=8, num_features=20)):
        _, n = param_offsets(cfg)
        buckets = gradient_buckets(cfg)
        assert len(buckets) == cfg.num_layers_rnn + 2
        assert sorted(buckets)[0][0] == 0 and sorted(buckets)[-1][1] == n
        srt = sorted(buckets)
        assert all(srt[i][1] == srt[i + 1][0] for i in range(len(srt) - 1))          # no gap, no overlap
        assert buckets[0][1] == n                                                     # dense4 + logits come first, at the end
        assert all(lo % 64 == 0 for lo, _ in buckets)                                 # 256-B aligned float offsets
    _, n = param_offsets(CFG)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, port, n, gradient_buckets(CFG), out), nprocs=2, join=True)
    assert out["same"]
    assert parallel.allreduce_gradients(torch.zeros(3), async_op=True) is None        # single process: nothing to wait for
