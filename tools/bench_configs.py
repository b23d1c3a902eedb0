#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (3: guiding centre / DoubleDipole, 4: Adaptive
Speiser, 5: time-dependent dipole) on one GPU, device-resident, CUDA events.  The driver's headline line
is bench.py (config 2); this script records the rest under profiles/.

    python tools/bench_configs.py gc 1048576 10.0
    python tools/bench_configs.py belt 1048576 2.0
    python tools/bench_configs.py adaptive 65536 300
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from rapt_b200 import engine, synth, fields, _lib

what = sys.argv[1]; n = int(float(sys.argv[2])); delta = float(sys.argv[3])
arith = sys.argv[4] if len(sys.argv) > 4 else "fast"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
_lib.init(0)
dev = torch.device("cuda:0")
res = {"workload": what, "n": n, "delta": delta, "arith": arith}
fp64, _ = engine.fp64_peak()

if what in ("gc", "belt"):
    if what == "gc":
        ic = synth.config3_electrons(n); f = fields.DoubleDipole(); dtv = 0.1; F_B = 48
    else:
        ic = synth.config5_belt(n); f = fields.VarEarthDipole(0.1, 10); dtv = 0.05; F_B = 25 + 15
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = engine.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"], arith=arith)
    cols0 = [torch.tensor(a, device=dev) for a in (ic["t0"], ic["x"], ic["y"], ic["z"], ppar)]
    ex = {k: torch.tensor(v, device=dev) for k, v in dict(mu=mu, v=ic["v"], mass=ic["mass"], charge=ic["charge"], dt=np.full(n, dtv)).items()}
    out = engine.alloc_outputs(n, dev)
    ms = []
    for r in range(reps):
        cols = [c.clone() for c in cols0]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        engine.gc_advance_dev(f, cols, ex["mu"], ex["v"], ex["mass"], ex["charge"], ex["dt"], delta, out, arith=arith)
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    steps = int(out["counters"][:, 1].to(torch.int64).sum()); rows = int(out["nrows"].to(torch.int64).sum()) - n
    nrhs_B = 7 + (2 if what == "belt" else 0)
    f_rhs = nrhs_B * F_B + 176
    flops = steps * (6 * f_rhs + 4 * 64 + 20) + rows * (f_rhs + 40)
    t = min(ms) * 1e-3
    res.update(ms=min(ms), steps=steps, rows=rows, steps_per_s=steps / t, tflops_alg=flops / t / 1e12, fp64_peak=fp64,
               frac=flops / t / 1e12 / fp64, ok=int((out["status"] == 1).sum()), flop_per_step=flops / steps)
    # CPU oracle port on a sample
    import oracle as O
    ns = 2048
    of = O.make_field("DoubleDipole") if what == "gc" else O.make_field("VarEarthDipole", 0.1, 10)
    st = np.column_stack([ic["t0"][:ns], pos[:ns], ppar[:ns]])
    t0 = time.perf_counter()
    o = O.gc_advance(of, O.make_params(), st, mu[:ns], ic["v"][:ns], ic["mass"][:ns], ic["charge"][:ns], dtv, min(delta, 10.0),
                     store_every=0, nthreads=os.cpu_count())
    el = time.perf_counter() - t0
    res["cpu_port_steps_per_s"] = float(o["counters"][:, 1].sum() / el); res["cpu_cores"] = os.cpu_count()
elif what == "adaptive":
    ic = synth.config4_speiser(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    par = dict(solvertolerances=(1e-12, 1e-12), epss=0.02)
    ts = []
    for r in range(reps):
        t0 = time.perf_counter()
        o = engine.adaptive_advance(fields.Parabolic(), pos, vel, 0.0, 1.0, 1.0, delta, 1.0, store_every=0, max_rows=0, arith=arith, **par)
        ts.append(time.perf_counter() - t0)
    steps = int(o["counters"][:, 1].astype(np.int64).sum())
    res.update(wall_s=min(ts), steps=steps, steps_per_s=steps / min(ts), epochs=o["epochs"], nseg_hist=np.bincount(o["nseg"]).tolist(),
               ok=int((o["status"] == 1).sum()), final_mode_gc=int(o["mode"].sum()))
    import oracle as O
    ns = 64
    of = O.make_field("Parabolic"); op = O.make_params(GCtimestep=1, **par)
    t0 = time.perf_counter(); tot = 0
    for i in range(ns):
        nseg, rows, seglog, cnt = O.adaptive_c(of, op, pos[i], vel[i], 0.0, 1.0, 1.0, delta)
        tot += int(cnt[1])
    res["cpu_port_steps_per_s_1core"] = tot / (time.perf_counter() - t0)
print(json.dumps(res))
