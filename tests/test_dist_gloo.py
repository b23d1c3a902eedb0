"""CPU, world_size 2 (gloo): the multi-GPU plumbing -- round-robin sharding, all-gather of final
states back into member order, all-reduce of diagnostics.  Each rank advances its shard with the CPU
oracle standing in for the kernels (no GPU here); the reassembled result must equal the
single-process run bit for bit, because particles are independent."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- independent restatements of the collection (plain torch.distributed calls, round-robin shards) that the product's
# ---- gather_rows / reduce_diagnostics / ShardPlan are checked against
def rd_shard_sizes(n_total, world):
    return [len(range(r, n_total, world)) for r in range(world)]


def unshard(gathered, n_total, world):
    """Inverse of the round-robin sharding: `gathered[r]` holds rank r's rows (padded to the largest
    shard); returns the (n_total, ...) array in global member order."""
    sizes = rd_shard_sizes(n_total, world)
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
    n_max = max(rd_shard_sizes(n_total, world))
    k = local.shape[1]
    pad = torch.zeros((n_max, k), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    buf = torch.empty((world, n_max, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=group)
    sizes = rd_shard_sizes(n_total, world)
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



def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle as O
    from rapt_b200 import synth, dist as rd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
    sl = rd.shard_slice(n, world, rank)
    o = O.particle_advance(O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20), st[sl], ic["mass"][sl],
                           ic["charge"][sl], 0.05, store_every=0)
    local = torch.tensor(np.column_stack([o["state"], o["counters"][:, 1].astype(np.float64)]))
    full = all_gather_final(local, n)
    p = torch.tensor(np.linalg.norm(o["state"][:, 4:7], axis=1))
    h = all_reduce_histogram(torch.log10(p), 16, -21.0, -19.0)
    stats = all_reduce_stats(p)
    # the product's collection path (rapt_b200/ensemble.py:gather): this rank's rows already sit in its slot of the
    # gather buffer (on the GPU the packing kernel writes them there), in-place all-gather, reduce of hist + sums
    n_max = max(rd.shard_sizes(n, world))
    buf = torch.zeros((world, n_max, 8), dtype=torch.float64)
    buf[rank, :local.shape[0]] = local
    full2 = rd.gather_rows(buf, rank, n)
    hist = torch.histc(torch.log10(p), bins=16, min=-21.0, max=-19.0).to(torch.int64)
    sums = torch.tensor([float(len(p)), float(p.sum()), float((p * p).sum()), 0.0], dtype=torch.float64)
    rd.reduce_diagnostics(hist, sums)
    assert torch.equal(full2, full) and torch.equal(hist.to(torch.float64), h) and int(sums[0]) == n
    assert rd.world_rank() == (world, rank)
    # speed-weighted shards (dist.ShardPlan): another cut of the same ensemble, same gathered result
    plan = rd.ShardPlan(n, world, [1.0, 0.6])
    idx = plan.indices(rank)
    ow = O.particle_advance(O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20), st[idx], ic["mass"][idx],
                            ic["charge"][idx], 0.05, store_every=0)
    bufw = torch.zeros((world, max(plan.sizes()), 8), dtype=torch.float64)
    bufw[rank, :len(idx)] = torch.tensor(np.column_stack([ow["state"], ow["counters"][:, 1].astype(np.float64)]))
    assert len(idx) == plan.sizes()[rank] and plan.sizes()[0] > plan.sizes()[1]
    assert torch.equal(rd.gather_rows(bufw, rank, n, plan=plan), full)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), full=full.numpy(), hist=h.numpy(),
                 stats=np.array([stats["count"], stats["mean"], stats["min"], stats["max"]]))
    dist.destroy_process_group()


def test_sharded_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rapt_b200 import synth
    n, world = 101, 2          # odd: ragged shards (51 + 50)
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
    ref = O.particle_advance(O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20), st, ic["mass"], ic["charge"],
                             0.05, store_every=0)
    assert np.array_equal(got["full"][:, :7], ref["state"]), "N-rank sharded run must equal the 1-rank run bit for bit"
    assert np.array_equal(got["full"][:, 7], ref["counters"][:, 1])
    p = np.linalg.norm(ref["state"][:, 4:7], axis=1)
    assert got["hist"].sum() == n
    assert got["stats"][0] == n and got["stats"][2] == p.min() and got["stats"][3] == p.max()
    assert abs(got["stats"][1] / p.mean() - 1) < 1e-12


def test_shard_helpers():
    from rapt_b200 import dist as rd
    for n, w in ((10, 4), (7, 8), (100, 3), (0, 2)):
        sizes = rd.shard_sizes(n, w)
        assert sum(sizes) == n
        idx = np.arange(n)
        parts = []
        for r in range(w):
            a = idx[rd.shard_slice(n, w, r)]
            pad = np.full(max(sizes) if sizes else 0, -1); pad[:len(a)] = a
            parts.append(pad[:, None])
        if n:
            assert np.array_equal(unshard(parts, n, w)[:, 0], idx)


def test_shard_plan():
    """ShardPlan: every member on exactly one rank, sizes() = len(indices()), runs proportional to the weights, and the
    index arithmetic of the device un-interleave kernel (capi.cu:k_unshard) restated in numpy inverts it."""
    from rapt_b200 import dist as rd
    for n, w, weights in ((10, 4, None), (101, 2, [1, 1]), (101, 2, [1.0, 0.6]), (5000, 3, [1, 2, 3]), (20000, 8, 1 + 0.03 * np.arange(8)),
                          (4096, 8, np.ones(8)), (3, 4, [1, 1, 1, 1]), (0, 2, [1, 2]), (100000, 64, np.linspace(1, 2, 64))):
        plan = rd.ShardPlan(n, w, weights)
        sizes = plan.sizes()
        parts = [plan.indices(r) for r in range(w)]
        assert [len(p) for p in parts] == sizes and sum(sizes) == n
        assert np.array_equal(np.sort(np.concatenate(parts)) if n else np.zeros(0), np.arange(n))
        assert plan.off[0] == 0 and plan.off[-1] == plan.period and np.all(np.diff(plan.off) >= (0 if n < w else 1))
        if weights is not None and n >= 4096:
            wn = np.asarray(weights, float) / np.sum(weights)
            assert np.max(np.abs(np.array(sizes) / n - wn)) < 2.0 / 4096 + 2.0 / n
        # k_unshard: member m -> (rank r, row i of r's shard)
        m = np.arange(n)
        full, rem = divmod(n, plan.period)
        off2 = (plan.off.astype(np.int64) * rem) // plan.period if weights is not None else np.minimum(plan.off, rem)
        blk, j = m // plan.period, m % plan.period
        tail = blk == full
        r = np.where(tail, np.searchsorted(off2[1:], j, side="right"), np.searchsorted(plan.off[1:], j, side="right"))
        i = np.where(tail, full * (plan.off[r + 1] - plan.off[r]) + (j - off2[r]), blk * (plan.off[r + 1] - plan.off[r]) + (j - plan.off[r]))
        for rr in range(w):
            assert np.array_equal(parts[rr][i[r == rr]], m[r == rr])
    with pytest.raises(ValueError):
        rd.ShardPlan(10, 2, [1, 0])
    with pytest.raises(ValueError):
        rd.ShardPlan(10, 2, [1, 2, 3])


def test_ensemble_shard_and_reshard_host_logic(monkeypatch):
    """ParticleEnsemble.shard(weights) / reshard(): what each rank keeps (the device upload is stubbed out -- no GPU here).
    Every member array is cut with the same ShardPlan indices, the cuts of all ranks partition the ensemble, a second cut
    starts from the whole ensemble again, and reshard() without keep_full says so."""
    import rapt_b200 as R
    from rapt_b200 import synth, dist as rd, ensemble
    n, world = 10007, 3
    ic = synth.config2_protons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    monkeypatch.setattr(ensemble.ParticleEnsemble, "cuda", lambda self, device="cuda:0": self)
    full = R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole())
    seen = []
    for rank in range(world):
        monkeypatch.setattr(rd, "world_rank", lambda group=None, r=rank: (world, r))
        ens = R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole())
        ens.shard(weights=[1.0, 2.0, 1.5], keep_full=True)
        idx = rd.ShardPlan(n, world, [1.0, 2.0, 1.5]).indices(rank)
        assert ens.n == len(idx) and ens.n_total == n and (ens.world, ens.rank) == (world, rank)
        for name in ("state", "mass", "charge", "tcur", "counters", "status"):
            assert np.array_equal(getattr(ens, name), getattr(full, name)[idx]), name
        ens.reshard([1.0, 1.0, 1.0])                       # equal weights: runs of the period, not round-robin
        idx2 = rd.ShardPlan(n, world, [1, 1, 1]).indices(rank)
        assert ens.n == len(idx2) and np.array_equal(ens.state, full.state[idx2]) and not ens._plan.uniform
        seen.append(idx2)
        rr = R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole()).shard()
        assert rr._plan.uniform and np.array_equal(rr.state, full.state[rank::world])
        with pytest.raises(RuntimeError):
            rr.reshard([1, 1, 1])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n))
    # one process: nothing to cut, reshard is a no-op
    monkeypatch.setattr(rd, "world_rank", lambda group=None: (1, 0))
    one = R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole()).shard(weights=[1.0], keep_full=True)
    assert one.n == n and one.reshard([1.0]) is one
