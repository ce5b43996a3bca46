"""Data parallelism for the hot path: one process per GPU, equal shards, ONE all-reduce (sum) of the
flat gradient buffer per step.  `overlap=True` issues it as one asynchronous all-reduce per bucket of
`CTCModel.gradient_buckets()` (dense4 + logits, each RNN layer, the front-end) as soon as the backward pass has produced
the bucket; measured on B200 this is no faster than the single all-reduce after backward (the default), see
CTCModel.train_step.  The reference is single-device
(`train_distribute=None`, asr/train.py:41); this is the only collective the path needs, so the
plumbing is `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) and nothing else.

Equal shard sizes make mean-of-shard-means equal the global mean; each rank scales its CTC gradient
by 1/global_batch (`CTCModel.loss_fn(..., global_batch=...)`) so the summed gradient is exactly the
gradient of the global mean loss (asr/model.py:267).
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world):
    """[lo, hi) of this rank's utterances.  The global batch must divide evenly (like the reference's
    drop_remainder=True batching, asr/input_functions.py:100-103)."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_indices(global_batch, rank, world, interleave=False):
    """Utterance indices of this rank's shard.  interleave=False: a contiguous block.  interleave=True:
    rank, rank + world, rank + 2 world, ... — when the caller has sorted the batch by duration (the synthetic cfg4
    batches are; the reference's bucketed batches hold similar lengths in arrival order) this gives every GPU the same
    count AND the same mix of lengths, so the ranks' step times stay balanced."""
    lo, hi = shard_bounds(global_batch, rank, world)
    if not interleave:
        return slice(lo, hi)
    return slice(rank, global_batch, world)


def shard_batch(sequences, seq_length, labels, label_length, rank, world, interleave=False):
    idx = shard_indices(sequences.shape[0], rank, world, interleave)
    return sequences[idx], seq_length[idx], labels[idx], label_length[idx]


def allreduce_gradients(flat_grad, group=None, async_op=False):
    """Sum the flat gradient (or one bucket of it) over all ranks, in place.  async_op=True returns the
    torch.distributed work handle (None when there is nothing to reduce)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else flat_grad
    return None if async_op else flat_grad


def allreduce_mean_loss(local_loss_sum_over_global_batch, group=None):
    """Each rank holds sum(per-utterance loss)/global_batch; the sum over ranks is the global mean."""
    t = local_loss_sum_over_global_batch.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0]


def train_step(model, sequences, seq_length, labels, label_length, group=None, interleave=False, overlap=False):
    """One data-parallel step of `CTCModel` on this rank's shard of the global batch."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # the partly filled batches a bucketed epoch ends with (asr/input_functions.py:90-98) need not divide by the
    # world size: the remainder utterances are dropped (every rank drops the same ones), a batch smaller than
    # the world size is skipped
    gb = sequences.shape[0] - sequences.shape[0] % world
    if gb == 0:
        return None
    sequences, seq_length, labels, label_length = sequences[:gb], seq_length[:gb], labels[:gb], label_length[:gb]
    x, sl, lab, ll = shard_batch(sequences, seq_length, labels, label_length, rank, world, interleave)
    loss = model.train_step(x, sl, (lab, ll), global_batch=gb, overlap=overlap,
                            allreduce=lambda g, async_op=False: allreduce_gradients(g, group, async_op))
    return allreduce_mean_loss(loss, group)
