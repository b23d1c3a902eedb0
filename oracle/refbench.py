"""Timing of the REFERENCE's own CPU implementation of the hot path (BASELINE.json north_star:
"a Python loop over Particle objects, plus a multiprocessing run across all host cores").

TEST / MEASUREMENT INFRASTRUCTURE ONLY: bench.py's `cpu_baseline` leg and `--impl reference` arm call this; nothing under
rapt_b200/ does.  The unmodified reference (rapt.Particle / rapt.GuidingCenter, scipy's dop853 / dopri5 driving the
Python right-hand side) is imported through oracle/refshim.py from /root/reference (build container) or from
oracle/_ref/ (the pip --target copy made by __graft_entry__.build(), which travels to the GPU box).

Particle-steps are scipy's own attempted-step counter iwork[17] summed over every r.integrate() call (refshim logs it).
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for _p in (_HERE, _ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_STATE = {}


def available():
    import refshim
    return refshim.reference_available()


def _setup(workload):
    """Per process: import the reference once, build the field and the seeded initial conditions."""
    if _STATE.get("workload") == workload:
        return _STATE
    import warnings
    warnings.filterwarnings("ignore")
    import refshim
    from rapt_b200 import synth          # initial conditions only (numpy; no device code is touched)
    rapt = refshim.load_reference()
    from rapt import fields as rf
    _STATE.update(workload=workload, rapt=rapt, refshim=refshim, synth=synth)
    if workload == "particle":
        _STATE.update(field=rf.EarthDipole(), par=dict(cyclotronresolution=20))
    elif workload == "gc":
        _STATE.update(field=rf.DoubleDipole(), par=dict(GCtimestep=0.1))
    elif workload == "belt":
        _STATE.update(field=rf.VarEarthDipole(0.1, 10), par=dict(GCtimestep=0.05))
    else:
        raise ValueError(workload)
    return _STATE


def _one(args):
    """advance() of one member of the seeded ensemble with the reference's class; returns (attempted steps, seconds)."""
    workload, n_total, i, delta = args
    S = _setup(workload)
    rapt, refshim = S["rapt"], S["refshim"]
    ic = S.get(("ic", n_total))
    if ic is None:
        gen = {"particle": S["synth"].config2_protons, "gc": S["synth"].config3_electrons, "belt": S["synth"].config5_belt}[workload]
        ic = S[("ic", n_total)] = gen(n_total)
    refshim.reset_params(rapt, **S["par"])
    refshim.SOLVER_LOG.clear()
    t = time.perf_counter()
    if workload == "particle":
        p = rapt.Particle(pos=(ic["x"][i], ic["y"][i], ic["z"][i]), vel=(ic["vx"][i], ic["vy"][i], ic["vz"][i]), t0=0,
                          mass=float(ic["mass"][i]), charge=float(ic["charge"][i]), field=S["field"])
    else:
        p = rapt.GuidingCenter(pos=(ic["x"][i], ic["y"][i], ic["z"][i]), v=float(ic["v"][i]), pa=float(ic["pa"][i]), t0=0,
                               mass=float(ic["mass"][i]), charge=float(ic["charge"][i]), field=S["field"])
    p.advance(delta)
    el = time.perf_counter() - t
    steps = sum(c[1] for c in refshim.SOLVER_LOG)
    refshim.SOLVER_LOG.clear()
    return steps, el


def time_loop(workload, n_sample, delta, n_total=None):
    """Plain Python loop over the first n_sample members on ONE core.  Returns (steps/s, steps, seconds)."""
    n_total = n_total or n_sample
    t = time.perf_counter()
    steps = sum(_one((workload, n_total, i, delta))[0] for i in range(n_sample))
    el = time.perf_counter() - t
    return steps / el, steps, el


_POOLS = {}


def _pool(procs):
    """One spawn-context pool per size, kept for the life of the process: the workers import the reference once, so the
    timed map measures advance() calls, not interpreter start-up.  spawn, not fork: the caller may hold a CUDA context."""
    import multiprocessing as mp
    if procs not in _POOLS:
        _POOLS[procs] = mp.get_context("spawn").Pool(procs)
    return _POOLS[procs]


def time_pool(workload, n_sample, delta, procs=None, n_total=None):
    """multiprocessing.Pool(procs) over the first n_sample members (one advance() per task).  Returns
    (steps/s, steps, wall seconds, procs)."""
    procs = procs or os.cpu_count() or 1
    n_total = n_total or n_sample
    pool = _pool(procs)
    pool.map(_one, [(workload, n_total, i % n_total, 1e-9) for i in range(procs)])     # workers up, reference imported
    t = time.perf_counter()
    res = pool.map(_one, [(workload, n_total, i, delta) for i in range(n_sample)], chunksize=1)
    el = time.perf_counter() - t
    steps = sum(r[0] for r in res)
    return steps / el, steps, el, procs


def close():
    for p in _POOLS.values():
        p.close(); p.join()
    _POOLS.clear()


if __name__ == "__main__":
    w = sys.argv[1] if len(sys.argv) > 1 else "particle"
    print("loop  :", time_loop(w, 4, 10.0, 64))
    print("pool  :", time_pool(w, 64, 10.0, None, 64))
    close()
