"""Developer timing harness for the gridded-field path (fields.Grid): the config-2 proton
ensemble in an EarthDipole sampled on a Cartesian grid.  Device-resident state, CUDA events.
    python tools/bench_grid.py [n] [delta] [nfiles] [nx] [arith]
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rapt_b200 import engine, synth, fields, _lib, Re, B0

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 20
delta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
nfiles = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nx = int(sys.argv[4]) if len(sys.argv) > 4 else 129
arith = sys.argv[5] if len(sys.argv) > 5 else "fast"
nz = nx // 2 + 1


class DipoleGrid(fields.Grid):
    """EarthDipole sampled on [-9, 9] Re x [-9, 9] Re x [-4.5, 4.5] Re; "file" k is the field at t = 20 k s."""
    def parsefile(self, filename):
        k = int(filename)
        x = np.linspace(-9, 9, nx) * Re; y = np.linspace(-9, 9, nx) * Re; z = np.linspace(-4.5, 4.5, nz) * Re
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        r2 = np.maximum(X * X + Y * Y + Z * Z, (0.5 * Re) ** 2)
        s = -B0 * Re ** 3 * (1 + 0.01 * k) / (r2 * r2 * np.sqrt(r2))
        zero = np.zeros_like(X)
        return {"time": 20.0 * k, "x": x, "y": y, "z": z, "Bx": s * 3 * X * Z, "By": s * 3 * Y * Z,
                "Bz": s * (2 * Z * Z - X * X - Y * Y), "Ex": zero, "Ey": zero, "Ez": zero}


_lib.init(0)
dev = torch.device("cuda:0")
f = DipoleGrid([str(k) for k in range(nfiles)])
ic = synth.config2_protons(n)
vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]]); mom = engine.particle_momentum(vel, ic["mass"])
st0 = [torch.tensor(a, device=dev) for a in (ic["t0"], ic["x"], ic["y"], ic["z"], mom[:, 0], mom[:, 1], mom[:, 2])]
mass = torch.tensor(ic["mass"], device=dev); charge = torch.tensor(ic["charge"], device=dev)
out = engine.alloc_outputs(n, dev)
res = {}
for name, fld in (("grid", f), ("analytic", fields.EarthDipole())):
    for r in range(2):
        cols = [c.clone() for c in st0]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        engine.particle_advance_dev(fld, cols, mass, charge, delta, out, arith=arith, cyclotronresolution=20)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    steps = int(out["counters"][:, 1].to(torch.int64).sum()); nf = int(out["counters"][:, 0].to(torch.int64).sum())
    stat = out["status"].cpu().numpy()
    res[name] = dict(ms=ms, steps=steps, steps_per_s=steps / ms * 1e3, nfcn=nf, left_grid=int((stat == -6).sum()),
                     other_fail=int(((stat < 0) & (stat != -6)).sum()))
g = res["grid"]
table_mb = nfiles * nx * nx * nz * 32 / 1e6
corners = 8 if nfiles == 1 else 16
gather = g["nfcn"] * corners / 2 * 64            # one 64-byte segment per z-pair of cell vertices
print(json.dumps({"n": n, "delta": delta, "nfiles": nfiles, "grid": [nx, nx, nz], "arith": arith, "B_table_MB": table_mb,
                  "grid_run": g, "analytic_run": res["analytic"],
                  "gather_bytes": gather, "gather_GBps": gather / g["ms"] / 1e6,
                  "slowdown_vs_analytic": g["ms"] / res["analytic"]["ms"]}))
