"""Developer timing harness (not the contract bench): device-resident Particle.advance, CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rapt_b200 import engine, synth, fields, _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 20
delta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
arith = sys.argv[3] if len(sys.argv) > 3 else "fast"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
store_every = int(sys.argv[5]) if len(sys.argv) > 5 else 0
sort = int(sys.argv[6]) if len(sys.argv) > 6 else 1
_lib.init(0)
dev = torch.device("cuda:0")
ic = synth.config2_protons(n)
vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]]); mom = engine.particle_momentum(vel, ic["mass"])
st0 = [torch.tensor(a, device=dev) for a in (ic["t0"], ic["x"], ic["y"], ic["z"], mom[:, 0], mom[:, 1], mom[:, 2])]
mass = torch.tensor(ic["mass"], device=dev); charge = torch.tensor(ic["charge"], device=dev)
out = engine.alloc_outputs(n, dev)
max_rows = (int(sys.argv[7]) if len(sys.argv) > 7 else 32) if store_every else 0
rows = torch.empty((n, max_rows, 8), dtype=torch.float64, device=dev) if store_every else None
f = fields.EarthDipole()
for r in range(reps):
    cols = [c.clone() for c in st0]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    engine.particle_advance_dev(f, cols, mass, charge, delta, out, store_every=store_every, max_rows=max_rows, rows=rows,
                                arith=arith, cyclotronresolution=20, sort_by_work=sort)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = int(out["counters"][:, 1].to(torch.int64).sum()); nrow = int(out["nrows"].to(torch.int64).sum())
    if store_every:
        nst = int(out["nstored"].to(torch.int64).sum())
        print(f"   stored rows {nst:.4e} x 64 B = {nst*64/1e9:.2f} GB -> {nst*64/ms/1e6:.1f} GB/s written; buffer {n*max_rows*64/1e9:.1f} GB", flush=True)
    print(f"n={n} delta={delta} {arith} store_every={store_every} sort={sort}: {ms:.2f} ms  steps={steps:.4e} rows={nrow:.4e}  {steps/ms*1e3:.4e} steps/s "
          f" -> {steps/ms*1e3*1440/1e12:.2f} TFLOP/s alg", flush=True)
