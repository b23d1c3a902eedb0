#!/usr/bin/env python
"""Opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv`: executed warp instructions, thread instructions
and lanes per opcode class, plus the hottest low-occupancy regions (where lanes are lost)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
I = {k: i for i, k in enumerate(hdr)}
tot_w = tot_t = 0
ops = collections.defaultdict(lambda: [0, 0, 0])
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[I["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    base = op.split(".")[0]
    w = int(r[I["Instructions Executed"]]); t = int(r[I["Predicated-On Thread Instructions Executed"]])
    smp = int(r[I["# Samples"]])
    ops[base][0] += w; ops[base][1] += t; ops[base][2] += smp
    tot_w += w; tot_t += t
    recs.append((w, t, smp, src))
print(f"total warp inst {tot_w:.4g}  thread inst {tot_t:.4g}  lanes/inst {tot_t / tot_w:.2f}")
print(f"{'op':10s} {'warp inst':>12s} {'share':>7s} {'lanes':>6s} {'samples':>8s}")
for k, v in sorted(ops.items(), key=lambda x: -x[1][0])[:28]:
    print(f"{k:10s} {v[0]:12.4g} {100 * v[0] / tot_w:6.2f}% {v[1] / max(v[0], 1):6.2f} {v[2]:8d}")
fp64 = sum(v[0] for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe warp instructions: {fp64:.4g} = {100 * fp64 / tot_w:.1f}% of issued; DFMA share of FP64 {100 * ops['DFMA'][0] / fp64:.1f}%")
fl = 2 * ops["DFMA"][1] + ops["DMUL"][1] + ops["DADD"][1]
print(f"executed flop (2/DFMA, 1/DMUL, 1/DADD, predicated-on threads): {fl:.5g}")
# lane-loss histogram: warp instructions by active-lane bucket
b = collections.defaultdict(int)
for w, t, smp, src in recs:
    if w:
        b[min(int(t / w) // 4 * 4, 28)] += w
print("warp instructions by active lanes:", {f"{k}-{k + 3}": f"{100 * v / tot_w:.1f}%" for k, v in sorted(b.items())})
