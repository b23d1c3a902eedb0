"""Work-order study, step 1 (CPU, test infrastructure): the true per-tracer solver counters of the bench ensemble (config 2,
1,048,576 protons, advance(10 s)) from the C oracle, in chunks, to an .npz -- the ground truth the work-order predictor of
k_particle_dt (rapt_b200/csrc/kernels_tu.cu) is judged against (tools/work_order_study.py)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
from rapt_b200 import synth

if len(sys.argv) > 1 and sys.argv[1] in ("gc", "belt"):
    # guiding-centre ensembles (configs 3 and 5): python tools/work_order_counts.py gc|belt [n] [out.npz]
    name = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    out = sys.argv[3] if len(sys.argv) > 3 else f"/tmp/wo/{name}_counts.npz"
    gen, fld, dt = ((synth.config3_electrons, O.make_field("DoubleDipole"), 0.1) if name == "gc" else
                    (synth.config5_belt, O.make_field("VarEarthDipole", 0.1, 10), 0.05))
    ic = gen(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = O.gc_construct(fld, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
    st = np.column_stack([ic["t0"], pos, ppar])
    o = O.gc_advance(fld, O.make_params(), st, mu, ic["v"], ic["mass"], ic["charge"], dt, 10.0, store_every=0, nthreads=os.cpu_count())
    np.savez(out, counters=o["counters"], state=st, mu=mu, v=ic["v"], pa=ic["pa"], mass=ic["mass"], charge=ic["charge"])
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/wo/cfg2_counts.npz"
ic = synth.config2_protons(n)
vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
f, p = O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20)
cnt = np.zeros((n, 4), np.int32); nrows = np.zeros(n, np.int64)
t0 = time.time()
chunk = 65536
for k in range(0, n, chunk):
    sl = slice(k, min(k + chunk, n))
    o = O.particle_advance(f, p, st[sl], ic["mass"][sl], ic["charge"][sl], 10.0, store_every=0, nthreads=os.cpu_count())
    cnt[sl] = o["counters"]; nrows[sl] = o["nrows"]
    print(k, round(time.time() - t0, 1), flush=True)
np.savez_compressed(out, counters=cnt, nrows=nrows, state=st, mass=ic["mass"], charge=ic["charge"])
