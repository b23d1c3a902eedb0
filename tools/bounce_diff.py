import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, helpers as H, scipy_legs
from rapt_b200 import engine, _lib
_lib.init(0)
for name in ("g2_gc_doubledipole", "gc_earthdipole", "gc_pa90_equatorial"):
    d, par = H.load(name); f = H.gpu_field(*H.GC_CASES[name])
    st0 = d["traj"][0]
    for arith in ("strict", "fast"):
        bs = engine.bounce_setup(f, st0, float(d["mu"]), float(d["mass"]), arith=arith)
        k = int(bs["npts"][0])
        bp = scipy_legs.bounceperiod(f, st0, float(d["mu"]), float(d["mass"]), arith=arith)[0]
        print(name, arith, "ds rel", bs["ds"][0]/float(d["bs_ds"])-1, "curve max rel", np.max(np.abs(bs["curve"][0,:k,:4]-d["bs_curve"]))/np.max(np.abs(d["bs_curve"])),
              "B rel", np.max(np.abs(bs["curve"][0,:k,4]/d["bs_B"]-1)), "period rel", bp/float(d["bs_period"])-1)
