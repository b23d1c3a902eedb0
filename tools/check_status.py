"""Which members of the config-2 ensemble end with a solver failure, and does the oracle agree?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from rapt_b200 import engine, synth, fields, _lib
import oracle as O
_lib.init(0)
n = 1 << 20; delta = 10.0
ic = synth.config2_protons(n)
vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]]); mom = engine.particle_momentum(vel, ic["mass"])
st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], mom])
o = engine.particle_advance(fields.EarthDipole(), st, ic["mass"], ic["charge"], delta, store_every=0, arith="fast", cyclotronresolution=20)
vals, cnts = np.unique(o["status"], return_counts=True)
print("status histogram:", dict(zip(vals.tolist(), cnts.tolist())))
bad = np.where(o["status"] != 1)[0]
print("n bad", len(bad))
if len(bad):
    r = np.sqrt(ic["x"] ** 2 + ic["y"] ** 2) / synth.Re
    rf = np.linalg.norm(o["state"][bad, 1:4], axis=1) / synth.Re
    print("initial L of bad:", np.percentile(r[bad], [0, 50, 100]), " KE MeV:", np.percentile(ic["ke_ev"][bad] / 1e6, [0, 50, 100]))
    print("final r/Re of bad:", np.percentile(rf, [0, 50, 100]), " t final:", np.percentile(o["state"][bad, 0], [0, 50, 100]))
    sel = bad[:24]
    ref = O.particle_advance(O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20), st[sel], ic["mass"][sel], ic["charge"][sel], delta, store_every=0, nthreads=8)
    print("oracle status for the same members:", ref["status"].tolist())
    print("gpu    status                     :", o["status"][sel].tolist())
    print("nrows equal:", np.array_equal(ref["nrows"], o["nrows"][sel]), " nstep equal:", (ref["counters"][:, 1] == o["counters"][sel, 1]).mean())
