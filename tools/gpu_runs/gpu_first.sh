set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_gpu_particle.py -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import numpy as np
from rapt_b200 import engine, synth, fields, _lib
_lib.init(0)
print("fp64 peak TF/s, MHz:", engine.fp64_peak())
print("fp64 peak TF/s, MHz:", engine.fp64_peak(1<<17))
for n, delta in [(1<<16, 1.0), (1<<18, 1.0), (1<<20, 1.0)]:
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]]); mom = engine.particle_momentum(vel, ic["mass"])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], mom])
    for arith in ("fast","strict"):
        if arith=="strict" and n > 1<<18: continue
        t=time.time(); o = engine.particle_advance(fields.EarthDipole(), st, ic["mass"], ic["charge"], delta, store_every=0, arith=arith, cyclotronresolution=20); el=time.time()-t
        steps = o["counters"][:,1].astype(np.int64).sum()
        print(f"n={n} delta={delta} {arith}: {el:.3f}s wall  steps={steps:.3e}  {steps/el:.3e} steps/s  rows={o['nrows'].astype(np.int64).sum():.3e} status_ok={np.all(o['status']==1)}")
PY
