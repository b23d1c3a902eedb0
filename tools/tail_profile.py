#!/usr/bin/env python
"""Where does the end of the particle kernel go?  Needs the profiling build of the library
(make -C rapt_b200/csrc B=build_trace OUT=../librapt_b200_trace.so EXTRA=-DRAPT_RKN_TRACE_TIMES=1; RAPT_B200_LIB=...):
every tracer then reports when it was fetched and when it was retired (global ns timer) instead of dt / tcur.
Prints a JSON summary: lanes busy over time, when the queue ran empty, who is still running at the end."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rapt_b200 as R
from rapt_b200 import synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
_lib.init(0)
ic = synth.config2_protons(n)
R.params["cyclotronresolution"] = 20
ens = R.ParticleEnsemble(np.column_stack([ic["x"], ic["y"], ic["z"]]), np.column_stack([ic["vx"], ic["vy"], ic["vz"]]), 0.0,
                         ic["mass"], ic["charge"], R.fields.EarthDipole()).cuda()
ens.advance(10.0); ens.pull()                      # warm-up (also the run that is analysed: deterministic schedule)
t1 = ens.tcur.copy(); t0 = ens.dt.copy(); steps = ens.last_counters[:, 1].astype(float); rows = ens.nrows - 1
start = t0.min(); t0 -= start; t1 -= start
T = t1.max()
out = dict(n=n, kernel_ms=T / 1e6, queue_empty_ms=t0.max() / 1e6, steps_total=float(steps.sum()))
# lanes busy over time (tracers in flight), sampled at 200 points
ts = np.linspace(0, T, 201)
busy = [(int(((t0 <= t) & (t1 > t)).sum())) for t in ts]
out["busy_lanes_at_fraction_of_kernel"] = {f"{q:.2f}": busy[int(q * 200)] for q in (0.05, 0.25, 0.5, 0.75, 0.85, 0.9, 0.93, 0.95, 0.97, 0.98, 0.99, 0.995)}
out["lane_time_integral_over_lanes_x_T"] = float((t1 - t0).sum() / (75776 * T))
late = t1 > 0.95 * T
out["retired_in_last_5pct"] = dict(count=int(late.sum()), steps_mean=float(steps[late].mean()), steps_max=float(steps[late].max()),
                                   fetched_at_ms_min=float(t0[late].min() / 1e6), fetched_at_ms_median=float(np.median(t0[late]) / 1e6),
                                   us_per_step_median=float(np.median((t1 - t0)[late] / steps[late]) / 1e3),
                                   steps_per_row_mean=float((steps[late] / np.maximum(rows[late], 1)).mean()))
last = np.argsort(-t1)[:16]
out["last_16"] = [dict(member=int(i), steps=int(steps[i]), rows=int(rows[i]), fetched_ms=round(t0[i] / 1e6, 2), retired_ms=round(t1[i] / 1e6, 2),
                       us_per_step=round((t1[i] - t0[i]) / steps[i] / 1e3, 2)) for i in last]
mid = (t0 > 0.3 * T) & (t1 < 0.7 * T)
out["us_per_step_mid_kernel_median"] = float(np.median((t1 - t0)[mid] / steps[mid]) / 1e3)
first = np.argsort(t0)[:75776]
out["first_wave"] = dict(steps_mean=float(steps[first].mean()), steps_max=float(steps[first].max()), retired_ms_max=float(t1[first].max() / 1e6),
                         retired_ms_median=float(np.median(t1[first]) / 1e6))
if len(sys.argv) > 2:      # raw per-tracer arrays for offline analysis
    np.savez_compressed(sys.argv[2], fetched_us=(t0 / 1e3).astype(np.float32), retired_us=(t1 / 1e3).astype(np.float32),
                        steps=steps.astype(np.int32), rows=rows.astype(np.int32))
print(json.dumps(out))
