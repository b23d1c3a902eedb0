#!/usr/bin/env python
"""BounceCenter.advance on an ensemble: wall time through the host-pointer C ABI, dopri5 steps,
right-hand sides (each = 5 field-line traces + quadratures) per second.  Usage: bench_bc.py [n] [delta_s] [arith]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapt_b200 as rb
from rapt_b200 import _lib, Re, m_el, e, c

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
delta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
arith = sys.argv[3] if len(sys.argv) > 3 else "fast"
_lib.init(0)
rng = np.random.default_rng(20261017)
L = rng.uniform(3, 7, n); phi = rng.uniform(0, 2 * np.pi, n)
pos = np.column_stack([L * Re * np.cos(phi), L * Re * np.sin(phi), rng.uniform(-0.05, 0.05, n) * Re])
ke = np.exp(rng.uniform(np.log(1e5), np.log(5e6), n)) * e
g = 1 + ke / (m_el * c * c)
v = c * np.sqrt(1 - 1 / g ** 2)
pa = np.radians(rng.uniform(72, 88, n))          # radians: the constructor takes cos(pa) as given (BounceCenter.py:114)
ens = rb.BounceCenterEnsemble(pos, v, 0.0, pa, m_el, -e, rb.fields.EarthDipole())
par = dict(rb.params)
ens.advance(0.05, params=par, arith=arith)        # warm-up (module load, pools)
st = ens.state.copy()
best = None
for _ in range(2):
    ens.state = st.copy()
    t0 = time.perf_counter()
    ens.advance(delta, params=par, arith=arith)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
cnt = ens.last_counters.sum(0)
print(json.dumps(dict(workload=f"{n} electrons / EarthDipole / BounceCenter.advance({delta} s) / BCtimestep 0.1 / {arith}",
                      wall_s=best, rows=int(ens.nrows.sum()), dopri5_steps=int(cnt[1]), rhs=int(cnt[0]),
                      field_line_traces=int(5 * cnt[0]), rhs_per_s=cnt[0] / best, steps_per_s=cnt[1] / best,
                      ok_fraction=float((ens.status == 1).mean()),
                      reference_rhs_per_s="12.5 (80 ms per right-hand side in the Python reference, oracle/gen_golden.py bc)")))
