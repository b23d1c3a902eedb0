"""Work-order study, step 2 (CPU, test infrastructure): candidate predictors of a tracer's solver steps against the oracle's
true counts (tools/work_order_counts.py), and what each order costs at the end of a launch: list scheduling of the tracers
in key order onto L lanes that each take `tau` per step (the persistent lanes of k_particle_rkn pulling from the sorted
queue), makespan against total work / L.

Predictors:
  rows        delta / dt                                   (round 1)
  sin         rows * max(1, 0.9 / sin(alpha))              (round 2, k_particle_dt until call 22)
  march       the parallel motion along the field line marched for delta (guiding-centre mirror force, leapfrog), steps
              accumulated as rows-per-time * g(B / B0), g(r) = max(1, c * r^p)
"""
import sys, heapq
import numpy as np

C = 299792458.0
B0E, RE = 3.07e-5, 6378137.0


def dipole(x, y, z):
    r2 = x * x + y * y + z * z
    w = -B0E * RE ** 3 / (r2 * r2 * np.sqrt(r2))
    return w * 3 * x * z, w * 3 * y * z, w * (2 * z * z - x * x - y * y)


def makespan(work, key, lanes):
    """List scheduling in ascending key order (longest-predicted first) onto `lanes` lanes; returns makespan / ideal."""
    order = np.argsort(key, kind="stable")
    w = work[order]
    n = len(w)
    if n <= lanes:
        return w.max() / (w.sum() / lanes)
    h = list(w[:lanes].astype(float))
    heapq.heapify(h)
    for v in w[lanes:]:
        heapq.heapreplace(h, h[0] + v)
    return max(h) / (w.sum() / lanes)


def march(st, mass, q, delta, cres, nsub=400, c=1.0, p=1.0, adaptive=None):
    """Predicted steps: leapfrog of (X, v_par) along the field line for `delta`; returns (pred_steps, rows, evals)."""
    x, y, z = st[:, 1].copy(), st[:, 2].copy(), st[:, 3].copy()
    px, py, pz = st[:, 4], st[:, 5], st[:, 6]
    p2 = px * px + py * py + pz * pz
    gm = np.sqrt(mass * mass + p2 / C ** 2)                 # gamma m
    v = np.sqrt(p2) / gm
    bx, by, bz = dipole(x, y, z)
    B0 = np.sqrt(bx * bx + by * by + bz * bz)
    dt_row = 2 * np.pi * gm / (np.abs(q) * B0) / cres
    vpar = (px * bx + py * by + pz * bz) / (gm * B0)
    Bm = B0 / np.maximum(1 - (vpar / v) ** 2, 1e-4)
    kacc = v * v / (2 * Bm)
    steps = np.zeros(len(x)); t = np.zeros(len(x))
    h = np.full(len(x), delta / nsub)
    eps = 1e-3 * np.sqrt(x * x + y * y + z * z)
    for it in range(nsub):
        bx, by, bz = dipole(x, y, z)
        B = np.sqrt(bx * bx + by * by + bz * bz)
        ux, uy, uz = bx / B, by / B, bz / B
        b1 = np.sqrt(sum(cc * cc for cc in dipole(x + eps * ux, y + eps * uy, z + eps * uz)))
        b2 = np.sqrt(sum(cc * cc for cc in dipole(x - eps * ux, y - eps * uy, z - eps * uz)))
        vpar = vpar - kacc * (b1 - b2) / (2 * eps) * h
        x += vpar * ux * h; y += vpar * uy * h; z += vpar * uz * h
        steps += h / dt_row * np.maximum(1.0, c * (B / B0) ** p)
    return steps, delta / dt_row


def study_gc(paths):
    """Guiding-centre keys (k_key_gc in capi.cu) against the oracle's counts: work_order_study.py gc <gc_counts.npz> ..."""
    for path in paths:
        d = np.load(path)
        ns = d["counters"][:, 1].astype(float); st = d["state"]; v = d["v"]
        n = len(ns); lanes = int(75776 * n / 1048576)
        r = np.linalg.norm(st[:, 1:4], axis=1)
        print(path, "steps per tracer: mean %.0f min %.0f max %.0f" % (ns.mean(), ns.min(), ns.max()),
              " steps / (delta v / r): 1 %% %.1f median %.1f 99 %% %.1f" % tuple(np.quantile(ns / (10 * v / r), [0.01, 0.5, 0.99])))
        for k, pr in {"member order": -np.arange(n).astype(float), "v/|r| (k_key_gc)": v / r, "true steps": ns}.items():
            print("  %-20s corr %.4f  makespan/ideal %.4f" % (k, np.corrcoef(pr, ns)[0, 1], makespan(ns, -pr, lanes)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gc":
        study_gc(sys.argv[2:] or ["/tmp/wo/gc_counts.npz", "/tmp/wo/belt_counts.npz"])
        sys.exit(0)
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "/tmp/wo/cfg2_counts.npz")
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else len(d["mass"])
    st, mass, q = d["state"][:nmax], d["mass"][:nmax], d["charge"][:nmax]
    nstep = d["counters"][:nmax, 1].astype(np.float64)
    n = len(nstep)
    lanes = int(round(75776 * n / 1048576))
    px, py, pz = st[:, 4], st[:, 5], st[:, 6]
    bx, by, bz = dipole(st[:, 1], st[:, 2], st[:, 3])
    B2 = bx * bx + by * by + bz * bz; p2 = px * px + py * py + pz * pz
    gm = np.sqrt(mass * mass + p2 / C ** 2)
    dt = 2 * np.pi * gm / (np.abs(q) * np.sqrt(B2)) / 20
    rows = 10.0 / dt
    sa = np.sqrt(np.maximum(1 - (px * bx + py * by + pz * bz) ** 2 / (p2 * B2), 1e-4))
    preds = {"exact": nstep, "rows": rows, "sin": rows / np.minimum(1.0, sa / 0.9)}
    for (c, p) in ((1.0, 1.0), (0.8, 1.0), (0.6, 1.0), (1.0, 0.5), (0.9, 0.5)):
        preds[f"march c={c} p={p}"] = march(st, mass, q, 10.0, 20, c=c, p=p)[0]
    print(f"n = {n}, lanes = {lanes}, steps: mean {nstep.mean():.0f} max {nstep.max():.0f}; ideal makespan {nstep.sum() / lanes:.0f} steps")
    for name, pr in preds.items():
        rel = pr / nstep
        top = nstep > np.quantile(nstep, 0.99)
        print(f"{name:22s} corr {np.corrcoef(pr, nstep)[0, 1]:.4f}  pred/true: median {np.median(rel):.3f} p1 {np.quantile(rel, 0.01):.3f} "
              f"p99 {np.quantile(rel, 0.99):.3f}; top 1 %: median {np.median(rel[top]):.3f} min {rel[top].min():.3f}   "
              f"makespan/ideal {makespan(nstep, -pr, lanes):.4f}")
