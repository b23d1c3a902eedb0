"""CPU, world_size 2 (gloo): the multi-GPU plumbing -- round-robin sharding, all-gather of final
states back into member order, all-reduce of diagnostics.  Each rank advances its shard with the CPU
oracle standing in for the kernels (no GPU here); the reassembled result must equal the
single-process run bit for bit, because particles are independent."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    full = rd.all_gather_final(local, n)
    p = torch.tensor(np.linalg.norm(o["state"][:, 4:7], axis=1))
    h = rd.all_reduce_histogram(torch.log10(p), 16, -21.0, -19.0)
    stats = rd.all_reduce_stats(p)
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
            assert np.array_equal(rd.unshard(parts, n, w)[:, 0], idx)
