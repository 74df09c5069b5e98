"""Batch sharding across the GPUs of one box.

The hot path is embarrassingly parallel over utterances: every rank runs front-end -> listener ->
speller on its own contiguous slice of the batch and there is NO data-path collective; the only
exchange is gathering the decoded ids (variable length) on the host side.  Works with any
``torch.distributed`` backend (nccl on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced slice [lo, hi) of ``n_items`` for ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(wave, n_samples, world_size=None, rank=None):
    """Slice a [B, N] batch (and its lengths) for this rank."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(wave.shape[0], world_size, rank)
    return wave[lo:hi], (n_samples[lo:hi] if n_samples is not None else None), (lo, hi)


def gather_ids(sample_ids, seq_len, n_global, pad_id=0):
    """All-gather per-rank decode results ([b_r, S_r] ids, [b_r] lengths) into [n_global, S_max] / [n_global]
    on every rank, in global utterance order.  Steps beyond a rank's own S_r are filled with ``pad_id``."""
    world = dist.get_world_size()
    dev = sample_ids.device
    shape = torch.tensor([sample_ids.shape[0], sample_ids.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    b_max = max(int(s[0]) for s in shapes)
    s_max = max(int(s[1]) for s in shapes)
    buf = torch.full((b_max, s_max), pad_id, dtype=torch.int32, device=dev)
    buf[:sample_ids.shape[0], :sample_ids.shape[1]] = sample_ids.to(torch.int32)
    lens = torch.zeros((b_max,), dtype=torch.int32, device=dev)
    lens[:seq_len.shape[0]] = seq_len.to(torch.int32)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    lbufs = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(bufs, buf)
    dist.all_gather(lbufs, lens)
    ids = torch.cat([bufs[r][:int(shapes[r][0])] for r in range(world)], dim=0)
    out_len = torch.cat([lbufs[r][:int(shapes[r][0])] for r in range(world)], dim=0)
    assert ids.shape[0] == n_global, (ids.shape, n_global)
    return ids, out_len


def allreduce_gradients(flat_grads):
    """The one exchange step of the training path: sum the flat fp32 gradient buffer over the
    ranks (NCCL over NVLink/NVSwitch on the GPU box).  ``train.apply_gradients`` has already clipped every tensor
    locally and divided by the world size, so the sum is the mean of the clipped shard gradients -- the order
    tf.tpu.CrossShardOptimizer gives the reference (clip at model_helper.py:416, cross-shard mean inside
    apply_gradients, model_helper.py:405-406,417).  All ranks then apply identical Adam updates to identical
    replicas of the parameters and optimiser state."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def broadcast_parameters(flat_params, src=0):
    """Make every replica start from rank ``src``'s parameters."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src=src)
    return flat_params
