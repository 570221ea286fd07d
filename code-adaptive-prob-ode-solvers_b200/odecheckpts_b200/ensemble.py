"""Ensemble sharding over the GPUs of one box (new relative to the reference, which has no
ensemble or multi-device code: SURVEY section 0.2, 8e).

Members are independent, so the solver loop needs no exchange step: member b goes to rank
b mod G (interleaved, so any smooth dependence of step count on the initial value spreads
evenly), every rank runs its persistent kernel on its shard, and ONE all-gather over
NCCL/NVLink collects checkpoint means, standard deviations and step statistics.
"""

import numpy as np


def shard_indices(B, rank, world_size):
    """Interleaved member assignment: rank r owns members r, r+G, r+2G, ..."""
    return np.arange(rank, B, world_size)


def shard_sizes(B, world_size):
    return [(B - r + world_size - 1) // world_size for r in range(world_size)]


def unshard_order(B, world_size):
    """Permutation that maps rank-major concatenated shards back to member order."""
    order = np.concatenate([shard_indices(B, r, world_size) for r in range(world_size)])
    inv = np.empty(B, dtype=np.int64)
    inv[order] = np.arange(B)
    return inv


def all_gather_results(local, B, group=None):
    """All-gather a dict of per-member tensors (leading axis = local members) and restore member
    order.  Works with NCCL (CUDA tensors) and gloo (CPU tensors).  Shards may differ in size by
    one member; they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    pad_to = max(sizes)
    inv = torch.as_tensor(unshard_order(B, world))
    out = {}
    for key, x in local.items():
        n_local = x.shape[0]
        if n_local < pad_to:
            pad = torch.zeros((pad_to - n_local,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            x = torch.cat([x, pad], 0)
        x = x.contiguous()
        gathered = torch.empty((world * pad_to,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(gathered, x, group=group)
        parts = [gathered[r * pad_to : r * pad_to + sizes[r]] for r in range(world)]
        cat = torch.cat(parts, 0)
        out[key] = cat[inv.to(cat.device)]
    return out
