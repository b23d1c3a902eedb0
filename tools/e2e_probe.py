import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from rapt_b200 import engine, synth, fields, _lib
from rapt_b200._lib import ptr, check
_lib.init(0)
n = 1 << 20
ic = synth.config2_protons(n)
vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]]); mom = engine.particle_momentum(vel, ic["mass"])
cols = [ic["t0"], ic["x"], ic["y"], ic["z"], mom[:, 0], mom[:, 1], mom[:, 2]]
pin = [torch.tensor(c).pin_memory() for c in cols]; pm = torch.tensor(ic["mass"]).pin_memory(); pq = torch.tensor(ic["charge"]).pin_memory()
hwork = [torch.empty_like(c).pin_memory() for c in pin]
hout = dict(nrows=np.zeros(n, np.int32), nstored=np.zeros(n, np.int32), counters=np.zeros((n, 4), np.int32), status=np.zeros(n, np.int32), tcur=np.zeros(n), dt=np.zeros(n))
f = fields.EarthDipole().device_descriptor(); p = engine.snapshot_params(None, False, cyclotronresolution=20)
lib = _lib.load()
for it in range(4):
    t0 = time.perf_counter()
    for w, p0 in zip(hwork, pin): w.copy_(p0)
    t1 = time.perf_counter()
    check(lib.rapt_b200_particle_advance(C.byref(f), C.byref(p), C.c_int64(n), *[ptr(w.numpy()) for w in hwork], ptr(pm.numpy()), ptr(pq.numpy()),
        C.c_double(10.0), C.c_int64(0), C.c_int64(0), None, ptr(hout["nrows"]), ptr(hout["nstored"]), ptr(hout["counters"]), ptr(hout["status"]), ptr(hout["tcur"]), ptr(hout["dt"])))
    t2 = time.perf_counter()
    print(f"iter {it}: host copy {1e3*(t1-t0):.1f} ms, C-ABI call {1e3*(t2-t1):.1f} ms", flush=True)
