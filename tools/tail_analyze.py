#!/usr/bin/env python
"""Offline analysis of the per-tracer fetch / retire times written by tools/tail_profile.py (npz): step rate and busy lanes
over time, distribution of microseconds per step over tracers, who is still running at the end."""
import sys
import numpy as np

d = np.load(sys.argv[1])
t0 = d["fetched_us"].astype(float); t1 = d["retired_us"].astype(float); st = d["steps"].astype(float); rows = d["rows"].astype(float)
T = t1.max(); ups = (t1 - t0) / st
rate = st / np.maximum(t1 - t0, 1e-9)
ev_t = np.concatenate([t0, t1]); ev_r = np.concatenate([rate, -rate]); ev_n = np.concatenate([np.ones_like(t0), -np.ones_like(t0)])
o = np.argsort(ev_t); ev_t = ev_t[o]; cr = np.cumsum(ev_r[o]); cn = np.cumsum(ev_n[o])
print(f"kernel {T / 1e3:.1f} ms, {int(st.sum()):,} steps, {st.sum() / T * 1e6:.3e} steps/s overall")
for f in (0.1, 0.3, 0.5, 0.7, 0.8, 0.85, 0.9, 0.95, 0.99):
    i = np.searchsorted(ev_t, f * T) - 1
    print(f"  t = {f * T / 1e3:6.1f} ms ({f:.2f}): {int(cn[i]):6d} lanes busy, {cr[i] * 1e6:.3e} steps/s")
print("us per step over tracers: 1% / 10% / 50% / 90% / 99% / max =", np.round(np.quantile(ups, [.01, .1, .5, .9, .99, 1]), 2))
lw = st > 20000
print(f"{int(lw.sum())} tracers with > 20,000 steps: us per step 0 / 10 / 50 / 90 / 100 % =", np.round(np.quantile(ups[lw], [0, .1, .5, .9, 1]), 2),
      "; retired at (ms) 10 / 50 / 90 / 100 % =", np.round(np.quantile(t1[lw], [.1, .5, .9, 1]) / 1e3, 1))
late = t1 > 0.9 * T
print(f"{int(late.sum())} tracers retire in the last 10 %: steps mean {st[late].mean():.0f}, fetched in the first ms: {(t0[late] < 1e3).mean():.2f}, "
      f"us per step median {np.median(ups[late]):.2f}")
