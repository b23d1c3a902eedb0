"""Multi-GPU plumbing: one process per GPU, ensembles sharded by particle.

Test particles do not interact and the fields are closed-form (a few doubles of parameters, replicated),
so the path shards trivially: rank r owns members r, r+W, r+2W, ... of the ensemble
(round-robin, so every rank sees the same distribution of orbit lengths), runs the same kernels on its
shard with NO data-path collective, and at the end the final states are all-gathered and fixed-size
diagnostics (histograms, invariant statistics) all-reduced -- NCCL over NVLink on GPUs, gloo in the CPU
tests.  Nothing here launches kernels; it only moves results.
"""
import numpy as np


def shard_slice(n_total, world, rank):
    """Members owned by `rank`: a strided slice of the global index range."""
    return slice(rank, n_total, world)


def shard_sizes(n_total, world):
    return [len(range(r, n_total, world)) for r in range(world)]


def unshard(gathered, n_total, world):
    """Inverse of the round-robin sharding: `gathered[r]` holds rank r's rows (padded to the largest
    shard); returns the (n_total, ...) array in global member order."""
    sizes = shard_sizes(n_total, world)
    first = np.asarray(gathered[0])
    out = np.empty((n_total,) + first.shape[1:], dtype=first.dtype)
    for r in range(world):
        out[r::world] = np.asarray(gathered[r])[:sizes[r]]
    return out


def all_gather_final(local, n_total, group=None):
    """All-gather the per-rank final states (torch tensor (n_local, k), any device) and return the
    global (n_total, k) tensor in member order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_max = max(shard_sizes(n_total, world))
    k = local.shape[1]
    pad = torch.zeros((n_max, k), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    buf = torch.empty((world, n_max, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=group)
    sizes = shard_sizes(n_total, world)
    out = torch.empty((n_total, k), dtype=local.dtype, device=local.device)
    for r in range(world):
        out[r::world] = buf[r, :sizes[r]]
    return out


def all_reduce_histogram(values, bins, lo, hi, group=None):
    """Histogram of a per-particle diagnostic over the whole ensemble (sum-all-reduce of local counts)."""
    import torch
    import torch.distributed as dist
    h = torch.histc(values.to(torch.float64), bins=bins, min=lo, max=hi)
    dist.all_reduce(h, group=group)
    return h


def all_reduce_stats(values, group=None):
    """(count, sum, sum of squares, min, max) of a diagnostic over the whole ensemble."""
    import torch
    import torch.distributed as dist
    v = values.to(torch.float64)
    s = torch.stack([torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device), v.sum(), (v * v).sum()])
    mn = v.min().reshape(1); mx = v.max().reshape(1)
    dist.all_reduce(s, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return dict(count=float(s[0]), mean=float(s[1] / s[0]), var=float(s[2] / s[0] - (s[1] / s[0]) ** 2),
                min=float(mn[0]), max=float(mx[0]))
