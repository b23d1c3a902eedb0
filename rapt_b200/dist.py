"""Multi-GPU plumbing: one process per GPU, ensembles sharded by particle.

Test particles do not interact and the fields are closed-form (a few doubles of parameters, replicated),
so the path shards trivially: rank r owns members r, r+W, r+2W, ... of the ensemble
(round-robin, so every rank sees the same distribution of orbit lengths; or runs of a 4096-member period
proportional to per-rank weights: ShardPlan), runs the same kernels on its shard with NO data-path collective, and at the end the final states are all-gathered and fixed-size
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


class ShardPlan:
    """Which member lives on which rank.  Shards are periodic: of every `period` consecutive members those at positions
    [off[r], off[r+1]) belong to rank r, in order.  weights None: period = world, one position each = round-robin (every
    GPU sees the same mix of orbit lengths).  With weights (one positive number per rank, e.g. the measured tracers per
    second of every GPU: `ens.shard(weights=...)`, `ens.reshard(...)`) the period is 4096 and the runs are proportional
    to the weights, so a slower GPU gets fewer tracers of the same mix; a small change of the weights moves only the
    members at the ends of the runs.  The last, partial period is cut in the same proportions."""

    PERIOD = 4096

    def __init__(self, n_total, world, weights=None):
        self.n_total, self.world = int(n_total), int(world)
        if weights is None or world == 1:
            self.period = self.world
            self.off = np.arange(self.world + 1, dtype=np.int32)
            self.uniform = True
        else:
            w = np.asarray(weights, dtype=np.float64).ravel()
            if len(w) != world or not np.all(np.isfinite(w)) or not np.all(w > 0):
                raise ValueError("ShardPlan: one positive weight per rank")
            self.period = max(min(self.PERIOD, self.n_total), self.world)
            exact = w / w.sum() * self.period
            cnt = np.maximum(np.floor(exact).astype(np.int64), 1)
            for k in np.argsort(-(exact - np.floor(exact)), kind="stable")[:max(self.period - int(cnt.sum()), 0)]:
                cnt[k] += 1
            while cnt.sum() > self.period:               # only when a floor was lifted to 1
                cnt[np.argmax(cnt)] -= 1
            self.off = np.concatenate(([0], np.cumsum(cnt))).astype(np.int32)
            self.uniform = False

    def _tail(self):
        """Cut of the last, partial period (n_total % period members): the same proportions, offsets scaled down --
        floor(off * rem / period), the arithmetic rapt_b200_unshard_dev repeats."""
        full, rem = divmod(self.n_total, self.period)
        return full, (self.off.astype(np.int64) * rem) // self.period

    def sizes(self):
        if self.uniform:
            return [len(range(r, self.n_total, self.world)) for r in range(self.world)]
        full, off2 = self._tail()
        return [int(full * (self.off[r + 1] - self.off[r]) + off2[r + 1] - off2[r]) for r in range(self.world)]

    def indices(self, rank):
        """Global member indices of `rank`'s shard, in shard order."""
        if self.uniform:
            return np.arange(rank, self.n_total, self.world, dtype=np.int64)
        full, off2 = self._tail()
        base = np.arange(self.off[rank], self.off[rank + 1], dtype=np.int64)
        idx = (np.arange(full, dtype=np.int64)[:, None] * self.period + base[None, :]).ravel()
        return np.concatenate([idx, full * self.period + np.arange(off2[rank], off2[rank + 1], dtype=np.int64)])

    def table(self):
        """(period, offsets) for rapt_b200_unshard_dev; offsets None = round-robin."""
        return (self.period, None) if self.uniform else (self.period, self.off)


def gather_rows(buf, rank, n_total, group=None, out=None, plan=None):
    """All-gather of per-rank final-state rows.  buf: (world, n_max, k) tensor whose slot [rank] already holds this
    rank's rows (rapt_b200_final_diagnostics_dev packs them there, so the send buffer IS the receive slot: NCCL's
    in-place all-gather, no staging copy).  Returns the (n_total, k) tensor in global member order on every rank
    (one un-interleaving kernel, rapt_b200_unshard_dev; `out` is reused when given).  `plan`: the ShardPlan the shards
    were cut with (default: round-robin).
    With a host backend (gloo: the CPU tests and single-GPU boxes) the slot is staged through host memory."""
    import torch
    import torch.distributed as dist
    world = buf.shape[0]
    plan = plan or ShardPlan(n_total, world)
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
        period, off = plan.table()
        engine.unshard_dev(buf, out, n_total, period, off)
    else:                                              # host tensors (gloo tests of the plumbing)
        sizes = plan.sizes()
        for r in range(world):
            out[torch.as_tensor(plan.indices(r))] = buf[r, :sizes[r]]
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
