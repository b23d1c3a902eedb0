"""Multi-GPU plumbing: one process per GPU, ensembles sharded by particle.

Test particles do not interact and the fields are closed-form (a few doubles of parameters, replicated),
so the path shards trivially: rank r owns members r, r+W, r+2W, ... of the ensemble
(round-robin, so every rank sees the same distribution of orbit lengths), runs the same kernels on its
shard with NO data-path collective, and at the end the final states are all-gathered and fixed-size
diagnostics (histograms, invariant statistics) all-reduced -- NCCL over NVLink on GPUs, gloo in the CPU
tests.  Nothing here launches kernels; it only moves results.
"""
import os

import numpy as np


def world_rank(group=None):
    """(world size, rank) of the torch.distributed job this process belongs to; (1, 0) outside one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    return 1, 0


def local_device():
    """The GPU of this rank: cuda:LOCAL_RANK under torchrun, cuda:0 otherwise."""
    return f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"


def _on_device_backend(group=None):
    import torch.distributed as dist
    return "nccl" in str(dist.get_backend(group))


def gather_rows(buf, rank, n_total, group=None, out=None):
    """All-gather of per-rank final-state rows.  buf: (world, n_max, k) tensor whose slot [rank] already holds this
    rank's rows (rapt_b200_final_diagnostics_dev packs them there, so the send buffer IS the receive slot: NCCL's
    in-place all-gather, no staging copy).  Returns the (n_total, k) tensor in global member order on every rank
    (one un-interleaving kernel, rapt_b200_unshard_dev; `out` is reused when given).
    With a host backend (gloo: the CPU tests and single-GPU boxes) the slot is staged through host memory."""
    import torch
    import torch.distributed as dist
    world = buf.shape[0]
    if world > 1:
        if _on_device_backend(group) or not buf.is_cuda:
            dist.all_gather_into_tensor(buf.view(-1), buf[rank].reshape(-1), group=group)
        else:
            h = torch.empty(buf.shape, dtype=buf.dtype)
            dist.all_gather_into_tensor(h.view(-1), buf[rank].reshape(-1).cpu(), group=group)
            buf.copy_(h)
    if out is None or tuple(out.shape) != (n_total, buf.shape[2]):
        out = torch.empty((n_total, buf.shape[2]), dtype=buf.dtype, device=buf.device)
    if buf.is_cuda:
        from . import engine
        engine.unshard_dev(buf, out, n_total)
    else:                                              # host tensors (gloo tests of the plumbing)
        sizes = shard_sizes(n_total, world)
        for r in range(world):
            out[r::world] = buf[r, :sizes[r]]
    return out


def reduce_diagnostics(hist, stats, group=None):
    """Sum-all-reduce of the fixed-size diagnostics (histogram counts int64, invariant sums float64), in place."""
    import torch.distributed as dist
    world, _ = world_rank(group)
    if world == 1:
        return hist, stats
    for t in (hist, stats):
        if _on_device_backend(group) or not t.is_cuda:
            dist.all_reduce(t, group=group)
        else:
            h = t.cpu(); dist.all_reduce(h, group=group); t.copy_(h)
    return hist, stats


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
