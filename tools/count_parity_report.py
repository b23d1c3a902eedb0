#!/usr/bin/env python
"""Parity report on solver counters (VERDICT r1 'weak #2'): for a seeded ensemble, how many tracers have
(nfcn, nstep, naccpt, nrejct) different from the oracle's (= scipy's), and how close the nearest accept/reject decision
of exactly those tracers was to err = 1 (oracle.errgap).  A differing tracer whose smallest |err - 1| is at the round-off
level of the flavour's error estimate is a legitimately flipped decision; anything else would be a defect.

    python tools/count_parity_report.py [--backend host|gpu] [--n 4096] [--delta 0.25] [--config 2|3|5]

backend host = the kernel source compiled for the CPU (tests/hostcheck), gpu = the CUDA library (needs a B200)."""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as O
from rapt_b200 import synth, engine, fields


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="host"); ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--delta", type=float, default=0.25); ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    if a.backend == "host":
        import hostkernel as K
    else:
        from rapt_b200 import _lib
        _lib.init(0)
    n = a.n
    out = {"config": a.config, "n": n, "delta": a.delta, "backend": a.backend}
    if a.config == 2:
        ic = synth.config2_protons(n)
        vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
        st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], engine.particle_momentum(vel, ic["mass"])])
        par = dict(cyclotronresolution=20)
        ref = O.errgap(O.particle_advance, n, O.make_field("EarthDipole"), O.make_params(**par), st, ic["mass"], ic["charge"],
                       a.delta, store_every=0, nthreads=a.threads)
        f = fields.EarthDipole()
        for arith in ("strict", "fast"):
            if a.backend == "host":
                o = K.particle_advance(f, st, ic["mass"], ic["charge"], a.delta, store_every=0, rkn=(arith == "fast"),
                                       nthreads=a.threads, arith=arith, **par)
            else:
                o = engine.particle_advance(f, st, ic["mass"], ic["charge"], a.delta, store_every=0, arith=arith, **par)
            out[arith] = summarise(o, ref)
            # which rows differ?  rerun the differing tracers one by one with every row stored (column 7 of a stored row
            # = cumulative attempted steps) and compare with the oracle's per-call counters
            differ = np.where(np.any(o["counters"] != ref["counters"], axis=1))[0][:64]
            rows_info = []
            for i in differ:
                mr = int(ref["nrows"][i]) + 8
                r1 = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**par), st[i], ic["mass"][i], ic["charge"][i],
                                        a.delta, store_every=1, max_rows=mr, want_percall=True)
                if a.backend == "host":
                    o1 = K.particle_advance(f, st[i], ic["mass"][i], ic["charge"][i], a.delta, store_every=1, max_rows=mr,
                                            rkn=(arith == "fast"), nthreads=1, arith=arith, **par)
                else:
                    o1 = engine.particle_advance(f, st[i], ic["mass"][i], ic["charge"][i], a.delta, store_every=1, max_rows=mr,
                                                 arith=arith, **par)
                k = int(o1["nstored"][0])
                mine = np.diff(np.concatenate(([0], o1["rows"][0, 1:k, 7]))).astype(int)
                theirs = r1["percall"][:, 1]; rej = r1["percall"][:, 3]
                j = np.where(mine != theirs[:len(mine)])[0]
                rows_info += [(int(theirs[q]), int(mine[q]), int(rej[q])) for q in j]
            if rows_info:
                ri = np.array(rows_info)
                out[arith]["differing_rows"] = {
                    "count": len(ri), "min_steps_in_row_reference": int(ri[:, 0].min()), "max_abs_step_difference": int(np.abs(ri[:, 0] - ri[:, 1]).max()),
                    "rows_with_a_rejected_step_in_reference": int((ri[:, 2] > 0).sum()),
                    "note": "a row normally takes 1 step; rows with >= 3 steps start from a tiny HINIT step (a coordinate or momentum "
                            "component next to zero makes its error scale atol + rtol |y| tiny); the error estimate of such steps is "
                            "round-off, so the growth factors differ between operation orders while every step is accepted"}
    else:
        gen, fname, fargs, dt = ((synth.config3_electrons, "DoubleDipole", (), 0.1) if a.config == 3 else
                                 (synth.config5_belt, "VarEarthDipole", (0.1, 10), 0.05))
        ic = gen(n)
        of = O.make_field(fname, *fargs); f = getattr(fields, fname)(*fargs)
        pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
        ppar, mu = O.gc_construct(of, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
        st = np.column_stack([ic["t0"], pos, ppar])
        ref = O.errgap(O.gc_advance, n, of, O.make_params(), st, mu, ic["v"], ic["mass"], ic["charge"], dt, a.delta,
                       store_every=0, nthreads=a.threads)
        for arith in ("strict", "fast"):
            if a.backend == "host":
                o = K.gc_advance(f, st, mu, ic["v"], ic["mass"], ic["charge"], dt, a.delta, store_every=0,
                                 nthreads=a.threads, arith=arith)
            else:
                o = engine.gc_advance(f, st, mu, ic["v"], ic["mass"], ic["charge"], dt, a.delta, store_every=0, arith=arith)
            out[arith] = summarise(o, ref)
    print(json.dumps(out))


def summarise(o, ref):
    differ = np.any(o["counters"] != ref["counters"], axis=1)
    gap = ref["errgap"]
    pos = np.linalg.norm(o["state"][:, 1:4] - ref["state"][:, 1:4], axis=1) / np.linalg.norm(ref["state"][:, 1:4], axis=1)
    r = {"tracers_with_different_counters": int(differ.sum()), "fraction": float(differ.mean()),
         "nstep_total": int(o["counters"][:, 1].sum()), "nstep_total_oracle": int(ref["counters"][:, 1].sum()),
         "max_pos_relerr": float(pos.max()),
         "errgap_median_all": float(np.median(gap)), "errgap_1st_percentile_all": float(np.quantile(gap, 0.01))}
    if differ.any():
        r["errgap_of_differing"] = {"max": float(gap[differ].max()), "median": float(np.median(gap[differ]))}
        r["max_pos_relerr_of_differing"] = float(pos[differ].max())
    return r


if __name__ == "__main__":
    main()
