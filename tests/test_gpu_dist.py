"""GPU, world_size 2: the multi-GPU product path -- ParticleEnsemble / GuidingCenterEnsemble `.shard()`, `.advance()`,
`.gather()` under torch.distributed -- must reproduce the 1-rank device run bit for bit (tracers are independent, the
kernels are deterministic per tracer), and the all-reduced diagnostics must equal the ones of the unsharded run.
Two GPUs: NCCL, one rank per GPU.  One GPU (the driver's test box): both ranks on cuda:0 with the gloo backend (NCCL
refuses two ranks on one device); the kernels, the packing / histogram kernel and the sharding logic are the same."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(R, which, n):
    from rapt_b200 import synth
    if which == "particle":
        ic = synth.config2_protons(n)
        pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
        R.params["cyclotronresolution"] = 20
        return R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole()), 0.2
    ic = synth.config3_electrons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    R.params["GCtimestep"] = 0.1
    return R.GuidingCenterEnsemble(pos, ic["v"], pa=ic["pa"], mass=ic["mass"], charge=ic["charge"],
                                   field=R.fields.DoubleDipole()), 1.0


def _worker(rank, world, port, n, out_dir, two_gpus):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank if two_gpus else 0),
                      RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import rapt_b200 as R
    torch.cuda.set_device(rank if two_gpus else 0)
    dist.init_process_group("nccl" if two_gpus else "gloo", rank=rank, world_size=world)
    res = {}
    for which in ("particle", "gc"):
        ens, delta = _build(R, which, n)
        ens.shard().advance(delta)
        g = ens.gather(nbins=32)
        ens.pull()
        res[which + "_final"] = g["final"].cpu().numpy(); res[which + "_hist"] = g["hist"].cpu().numpy()
        res[which + "_stats"] = np.array([g["stats"]["ok"], g["stats"]["mean"], g["stats"]["var"], g["stats"]["outside"]])
        res[which + "_nlocal"] = np.array([ens.n, ens.n_total]); res[which + f"_steps{rank}"] = ens.counters[:, 1].sum()
        # speed-weighted shards (dist.ShardPlan), cut twice: shard(weights) then reshard(other weights)
        ens, delta = _build(R, which, n)
        ens.shard(weights=[1.0, 0.7], keep_full=True)
        nloc = ens.n
        ens.reshard([0.9, 1.0]).advance(delta)
        g = ens.gather(nbins=32)
        res[which + "_final_w"] = g["final"].cpu().numpy(); res[which + "_hist_w"] = g["hist"].cpu().numpy()
        res[which + "_nlocal_w"] = np.array([nloc, ens.n])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_run_equals_one_rank_run(tmp_path):
    import torch
    import torch.multiprocessing as mp
    import rapt_b200 as R
    from rapt_b200 import _lib
    n, world = 4097, 2                       # odd: ragged shards (2049 + 2048)
    two = torch.cuda.device_count() >= 2
    port = 29600 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path), two), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    _lib.init(0)
    old = dict(R.params)
    try:
        for which in ("particle", "gc"):
            ens, delta = _build(R, which, n)
            ens.cuda("cuda:0").advance(delta)
            g = ens.gather(nbins=32)           # world 1: the same kernel, no collective
            ens.pull()
            one = g["final"].cpu().numpy()
            assert np.array_equal(one, ens.state)
            assert tuple(r0[which + "_nlocal"]) == (2049, n) and tuple(r1[which + "_nlocal"]) == (2048, n)
            for r in (r0, r1):
                assert np.array_equal(r[which + "_final"], one), f"{which}: gathered 2-rank result != 1-rank run"
                assert np.array_equal(r[which + "_hist"], g["hist"].cpu().numpy())
                assert int(r[which + "_stats"][0]) == g["stats"]["ok"] == n
                assert abs(r[which + "_stats"][1] - g["stats"]["mean"]) <= 1e-12 * abs(g["stats"]["mean"])
            assert int(r0[which + "_steps0"]) + int(r1[which + "_steps1"]) == int(ens.counters[:, 1].sum())
            for r in (r0, r1):
                assert np.array_equal(r[which + "_final_w"], one), f"{which}: weighted shards, gathered result != 1-rank run"
                assert np.array_equal(r[which + "_hist_w"], g["hist"].cpu().numpy())
            assert r0[which + "_nlocal_w"][0] > r1[which + "_nlocal_w"][0] and r0[which + "_nlocal_w"][1] < r1[which + "_nlocal_w"][1]
            assert r0[which + "_nlocal_w"].sum() + r1[which + "_nlocal_w"].sum() == 2 * n
            # the histogram kernel against numpy on the same final states
            if which == "particle":
                m = ens.mass; p2 = np.sum(one[:, 4:7] ** 2, axis=1)
                gam = np.sqrt(1 + p2 / (m * R.c) ** 2)
                q = np.log10(np.where(gam - 1 < 1e-6, 0.5 * p2 / m, (gam - 1) * m * R.c ** 2) / R.e)
            else:
                q = np.linalg.norm(one[:, 1:4], axis=1) / R.Re
            ref, _ = np.histogram(q, bins=g["edges"])
            assert np.array_equal(ref, g["hist"].cpu().numpy()) and g["stats"]["outside"] == int(((q < g["edges"][0]) | (q > g["edges"][-1])).sum())
    finally:
        R.params.clear(); R.params.update(old)


def test_unshard_kernel_against_shard_plan():
    """rapt_b200_unshard_dev on its own: for round-robin and weighted plans (2 ... 64 ranks, ragged tails) the kernel puts
    every rank's rows back where ShardPlan.indices() took them from."""
    import torch
    from rapt_b200 import engine, _lib, dist as rd
    _lib.init(0)
    rng = np.random.default_rng(5)
    for n, w, weights in ((4097, 2, None), (4097, 2, [1.0, 0.7]), (100001, 8, 1 + 0.03 * np.arange(8)), (9000, 3, [1, 2, 3]),
                          (12345, 64, np.linspace(1, 2, 64)), (5, 8, np.ones(8)), (8192, 8, np.ones(8))):
        plan = rd.ShardPlan(n, w, weights)
        want = rng.standard_normal((n, 7))
        buf = np.zeros((w, max(max(plan.sizes()), 1), 7))
        for r in range(w):
            idx = plan.indices(r)
            buf[r, :len(idx)] = want[idx]
        out = torch.empty((n, 7), dtype=torch.float64, device="cuda:0")
        period, off = plan.table()
        engine.unshard_dev(torch.as_tensor(buf, device="cuda:0"), out, n, period, off)
        assert np.array_equal(out.cpu().numpy(), want), (n, w)
