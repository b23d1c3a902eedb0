python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "
import sys,time; sys.path.insert(0,'.')
import numpy as np
from rapt_b200 import engine, synth, fields, _lib
_lib.init(0)
n=1<<20
ic=synth.config3_electrons(n); f=fields.DoubleDipole()
pos=np.column_stack([ic['x'],ic['y'],ic['z']])
ppar,mu=engine.gc_construct(f,ic['t0'],pos,ic['v'],ic['pa'],ic['mass'])
st=np.column_stack([ic['t0'],pos,ppar])
t=time.time(); bp=engine.bounceperiod_device(f,st,mu,ic['mass'],arith='fast'); print('1M bounce periods on device: %.2f s, finite %.5f, median %.3f s'%(time.time()-t, np.isfinite(bp).mean(), np.nanmedian(bp)))
"
