python -m pytest tests/test_gpu_gc.py -m gpu -q -k "bounce_period_device and strict" 2>&1 | grep -E "^E|assert" | head -20
python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, helpers as H
from rapt_b200 import engine, synth, _lib
_lib.init(0)
d, par = H.load("e3_config3_first16")
n = int(d["n"]); ic = synth.config3_electrons(n); f = H.gpu_field("DoubleDipole", ())
pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
ppar, mu = engine.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
st = np.column_stack([ic["t0"], pos, ppar])
bp = engine.bounceperiod_device(f, st, mu, ic["mass"])
print("rel diff:", bp / d["bounceperiod"] - 1)
print("pa:", ic["pa"])
PY
