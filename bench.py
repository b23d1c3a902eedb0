#!/usr/bin/env python
"""bench.py -- headline benchmark of the particle-advance hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload particle|gc|belt|adaptive] [--n-per-gpu M] [--scaling weak|strong]

Metric: particle-steps/s (fp64) = attempted Runge-Kutta steps of all particles per second of device time, whole job.
Workload at N = 1: BASELINE.json configs[1] -- 1 M protons in EarthDipole, full orbit, random energies 0.1-10 MeV and
pitch angles (rapt_b200/synth.py, seed 20260201), Particle.advance(10 s), cyclotronresolution 20.  One bench "step" =
one such advance of the whole ensemble from the same initial state.

N > 1 goes through the product's own multi-GPU path (rapt_b200/ensemble.py): every rank builds the ensemble,
`.shard()` keeps members r, r+N, ... on its GPU, `.advance()` runs the kernels with no data-path collective,
`.gather()` packs + bins the final states in one kernel and all-gathers / all-reduces them over NCCL -- all inside the
timed region.  `--scaling weak` (default): M tracers per GPU (N x M in total); `--scaling strong`: M tracers in total.

`value` is device-resident (inputs already in HBM, CUDA events on the launching stream, max over ranks); `e2e` is the
same metric through the host-buffer C ABI (pinned host inputs, H2D + D2H inside the timed region).  `roofline` is
against the FP64 DFMA peak measured in this run (MEASURED_PEAKS.json has no FP64 figure), with the algorithmic flop
count of DESIGN.md.  `per_rank` splits every rank's step into kernel time and collection time (pack + collectives +
waiting for the slowest rank).  `cpu_baseline` / `--impl reference` time the REFERENCE's own CPU path -- a Python
loop over rapt.Particle objects and a multiprocessing.Pool over all host cores (oracle/refbench.py, the unmodified
reference installed into oracle/_ref by build()) -- on a bounded sample of the same workload; the C/OpenMP oracle
port is reported beside it.  `extra.workloads` carries the guiding-centre (configs 3, 5) and adaptive (config 4)
lines measured in the same run at N = 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
DELTA = 10.0
PARAMS = dict(cyclotronresolution=20)
WORK_ORDER = {"value": 1}          # --work-order previous: 2
F_RHS = 38            # Lorentz 18 + EarthDipole 20 flop (E == 0 compiled out)
F_STAGE = 6 * 158 + 20
GC_DT = {"gc": 0.1, "belt": 0.05}          # params["GCtimestep"] of configs 3 and 5
ADAPTIVE = dict(delta=300.0, gc_dt=1.0, params={"solvertolerances": (1e-12, 1e-12), "epss": 0.02})   # config 4

WORKLOAD_NAME = {
    "particle": "config2: protons / EarthDipole / Particle.advance(10 s) / cyclotronresolution 20 / KE 0.1-10 MeV / seed 20260201",
    "gc": "config3: electrons / DoubleDipole / GuidingCenter.advance (TaoChanBrizard) / GCtimestep 0.1 / KE 50 keV-1 MeV / "
          "seed 20260301",
    "belt": "config5: electrons / VarEarthDipole(0.1, 10 s) / GuidingCenter.advance / GCtimestep 0.05 / seed 20260501",
    "adaptive": "config4: Adaptive Speiser orbits / Parabolic current sheet / advance(300) / tol 1e-12 / epss 0.02 / "
                "GCtimestep 1 / seed 20260401",
}


def algorithmic_flops(nstep, naccpt, ncalls, f_rhs=F_RHS):
    """DESIGN.md 'flop accounting': 11 RHS per attempted step + 1 per accepted step, stage sums + error
    norm + controller per attempted step, HINIT (1 extra RHS + norms) per solver call (= output row)."""
    return nstep * (11 * f_rhs + F_STAGE) + naccpt * f_rhs + ncalls * (f_rhs + 60)


def gc_flops(workload, nstep, ncalls):
    """6 RHS + stage sums per attempted step, HINIT per row; RHS = n_B field evaluations + 176.
    Belt (VarEarthDipole) counts what the fast kernel executes, not the reference's 9 evaluations with a sine
    each: 7 dipole evaluations scaled by one time factor (~25 flop) per RHS; db/dt is identically zero.
    Adaptive (Parabolic, F_B = 3): 7 evaluations."""
    f_b, n_b, f_t = {"gc": (48, 7, 0), "belt": (23, 7, 25), "adaptive": (3, 7, 0)}[workload]
    f_rhs = n_b * f_b + f_t + 176
    return nstep * (6 * f_rhs + 4 * 64 + 20) + ncalls * (f_rhs + 40)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("RAPT_BENCH_SMI_MS", "200"), "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w=float(np.median(pw)) if pw else None, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------
# ensembles (product API)
# ----------------------------------------------------------------------------------------------------------------
def gc_field(workload):
    from rapt_b200 import fields
    return fields.DoubleDipole() if workload == "gc" else fields.VarEarthDipole(0.1, 10)


def build_ensemble(workload, n_total, world, rank, weights=None, keep_full=False):
    """The seeded ensemble of `workload` as the product's ensemble object, sharded onto this rank's GPU
    (weights: rapt_b200/dist.py:ShardPlan; None = round-robin)."""
    import rapt_b200 as R
    from rapt_b200 import synth
    big = n_total > 20_000_000      # 100 M-tracer ensembles: every rank draws its own shard (seed + rank) instead of
    n_make = n_total // world if big else n_total                 # holding the whole ensemble in host memory
    if workload == "particle":
        ic = synth.config2_protons(n_make, **({"seed": 20260201 + 1000 * rank} if big else {}))
        ens = R.ParticleEnsemble(np.column_stack([ic["x"], ic["y"], ic["z"]]), np.column_stack([ic["vx"], ic["vy"], ic["vz"]]),
                                 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole())
    else:
        gen = synth.config3_electrons if workload == "gc" else synth.config5_belt
        seed = (20260301 if workload == "gc" else 20260501) + (1000 * rank if big else 0)
        ic = gen(n_make, seed=seed)
        ens = R.GuidingCenterEnsemble(np.column_stack([ic["x"], ic["y"], ic["z"]]), ic["v"], pa=ic["pa"], mass=ic["mass"],
                                      charge=ic["charge"], field=gc_field(workload))
    if big:
        ens.world, ens.rank, ens.n_total, ens._group = world, rank, ens.n * world, None
        from rapt_b200 import dist as rd
        ens._plan = rd.ShardPlan(ens.n_total, world)
        return ens.cuda(rd.local_device())
    return ens.shard(weights=weights, keep_full=keep_full)


def run_advance(ens, workload, delta, arith):
    if workload == "particle":
        ens.advance(delta, arith=arith, sort_by_work=WORK_ORDER["value"], **PARAMS)
    else:
        ens.advance(delta, dt=GC_DT[workload], arith=arith, sort_by_work=WORK_ORDER["value"])


# ----------------------------------------------------------------------------------------------------------------
# CPU baselines
# ----------------------------------------------------------------------------------------------------------------
def port_sample(workload, n_sample, delta, nthreads):
    """The C/OpenMP oracle port on n_sample tracers drawn by the same generator (same seed and distributions)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rapt_b200 import synth
    t = time.perf_counter()
    if workload == "particle":
        ic = synth.config2_protons(n_sample)
        vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
        st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
        t = time.perf_counter()
        o = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**PARAMS), st, ic["mass"], ic["charge"], delta,
                               store_every=0, nthreads=nthreads)
    else:
        ic = synth.config3_electrons(n_sample) if workload == "gc" else synth.config5_belt(n_sample)
        f = O.make_field("DoubleDipole") if workload == "gc" else O.make_field("VarEarthDipole", 0.1, 10)
        pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
        ppar, mu = O.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
        st = np.column_stack([ic["t0"], pos, ppar])
        t = time.perf_counter()
        o = O.gc_advance(f, O.make_params(), st, mu, ic["v"], ic["mass"], ic["charge"], GC_DT[workload], delta, store_every=0,
                         nthreads=nthreads)
    el = time.perf_counter() - t
    steps = int(o["counters"][:, 1].sum())
    return steps / el, steps, el


def reference_pool_sample(workload, n_sample, delta, cores):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refbench
    return refbench.time_pool(workload, n_sample, delta, cores, n_sample)


def cpu_baseline_record(workload, delta, port_n):
    """`cpu_baseline` of our arm: the unmodified reference (Python loop on 1 core + multiprocessing.Pool over all cores)
    on a bounded sample; the C/OpenMP port of it as a second figure.  Falls back to the port alone when the reference
    copy is absent (oracle/_ref is made by __graft_entry__.build())."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    pv, psteps, pel = port_sample(workload, port_n, delta, cores)
    port = {"value": pv, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": f"{port_n} tracers of the same generator and seed, advance({delta} s): {psteps} steps in {pel:.1f} s, C oracle port with OpenMP"}
    try:
        import refbench
        if not refbench.available():
            raise RuntimeError("oracle/_ref missing")
        n_pool = max(cores, 64)
        rv, rsteps, rel, procs = refbench.time_pool(workload, n_pool, delta, cores, n_pool)
        lv, lsteps, lel = refbench.time_loop(workload, 4, delta, n_pool)
        refbench.close()
        return {"value": rv, "unit": "particle-steps/s", "cores": procs, "kind": "reference",
                "sample": f"unmodified rapt classes, {n_pool} tracers of the same generator and seed, advance({delta} s) each: "
                          f"multiprocessing.Pool({procs}) {rsteps} steps in {rel:.1f} s; plain Python loop on one core over the "
                          f"first 4: {lsteps} steps in {lel:.1f} s",
                "python_loop_1core": {"value": lv, "unit": "particle-steps/s", "cores": 1},
                "port": port}
    except Exception as ex:          # no reference copy on this box
        port["reference_unavailable"] = str(ex)[:200]
        return port


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- the unmodified rapt
    classes under multiprocessing.Pool(os.cpu_count()), each bench step a bounded sample of the
    workload sized so that warmup + steps end within a few minutes.  Falls back to the C oracle port when the copy of
    the reference (oracle/_ref, made by build()) is not there."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    workload = args.workload if args.workload != "adaptive" else "particle"
    delta = args.delta
    import refbench
    use_ref = refbench.available()
    total = args.warmup + args.steps
    if use_ref:
        # ~1.3e3 (Particle) / ~0.5e3 (GuidingCenter) steps/s per core; ~2.9e3 / ~1e3 steps per tracer: ~2 core-seconds per
        # tracer either way.  Budget ~150 s of wall time for the whole run.
        n_sample = args.ref_sample or int(min(256, max(cores, (150.0 / total) * cores / 2.2)))
    else:
        n_sample = args.cpu_sample
    vals = []
    for i in range(total):
        if use_ref:
            v, steps, el, _ = refbench.time_pool(workload, n_sample, delta, cores, n_sample)
        else:
            v, steps, el = port_sample(workload, n_sample, delta, cores)
        if i >= args.warmup:
            vals.append((v, steps, el))
    if use_ref:
        refbench.close()
    steps = sum(s for _, s, _ in vals); el = sum(e for _, _, e in vals)
    value = steps / el
    kind = "reference" if use_ref else "port"
    how = (f"unmodified rapt classes, multiprocessing.Pool({cores})" if use_ref else f"C oracle port, OpenMP x{cores}")
    sample = (f"{n_sample} tracers of the {WORKLOAD_NAME[workload].split(':')[0]} generator (same seed and distributions), "
              f"advance({delta} s) each, per bench step; {how}")
    print(json.dumps({
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(len(vals), 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[workload], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------------------
def measured_traffic(workload, n, delta):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch, from the committed `ncu --set full`
    capture of this configuration (profiles/ncu_traffic.json, written from the .ncu-rep by tools/ncu_summary.py)."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for e in tab["captures"]:
            if e["workload"] == workload and e["n"] == n and e["delta"] == delta:
                return e["dram_bytes"], e["source"]
    except Exception:
        pass
    return None, None


def time_workload(args, workload, n_per_gpu, world, rank, dev, steps, warmup, want_e2e=True):
    """Device-resident timing of one workload through the ensemble API.  Returns the record pieces (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from rapt_b200 import engine, _lib
    n_total = n_per_gpu * world if args.scaling == "weak" else n_per_gpu
    rebalance = bool(args.rebalance) and world > 1 and n_total <= 20_000_000
    ens = build_ensemble(workload, n_total, world, rank, weights=np.ones(world) if rebalance else None, keep_full=rebalance)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    cur = {}

    def attach():
        cur["d"] = ens._dev
        cur["pristine"] = [c.clone() for c in ens._dev.cols]

    def one_step(ev=None):
        ens.load_state(cur["pristine"])
        flush.fill_(1)                                                   # L2 flush between steps
        if ev is not None:
            ev[0].record()
        run_advance(ens, workload, args.delta, args.arith)
        if ev is not None:
            ev[1].record()
        if world > 1:
            ens.gather(nbins=64, sync=False)                             # pack + histogram kernel, all-gather, all-reduce
        if ev is not None:
            ev[2].record()

    attach()
    shard_history = None
    if rebalance:
        # Time-weighted shards: equal shards of the same random ensemble take up to 5 % different kernel times (the end of a
        # launch hangs on the few longest orbits each shard happens to hold: the per-rank times repeat to 0.3 ms on another
        # box, profiles/r2_multi_gpu.md) and a step ends with the slowest rank.  Before the warm-up: time the kernel on every
        # rank, give every rank a share of each 4096-member period proportional to its measured tracers per second, and
        # repeat once (the second cut moves only the ends of the runs).  What an application does that advances an
        # ensemble in several calls: the first call's times set the shards of the next.  Not part of the timed region; the
        # calibration steps and the shard sizes are reported in per_rank.
        shard_history = []
        weights = np.ones(world)
        for _ in range(int(args.rebalance)):
            one_step()
            cevs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(2)]
            for cev in cevs:
                one_step(cev)
            torch.cuda.synchronize()
            mine = torch.tensor([min(cev[0].elapsed_time(cev[1]) for cev in cevs), float(ens.n)], dtype=torch.float64, device=dev)
            allm = torch.empty((world, 2), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allm.view(-1), mine)
            allm = allm.cpu().numpy()
            shard_history.append({"sizes": [int(v) for v in allm[:, 1]], "kernel_ms": [round(float(v), 3) for v in allm[:, 0]]})
            if not (np.all(np.isfinite(allm)) and np.all(allm > 0)):
                break                                     # no usable timing on some rank: keep the current cut
            speed = allm[:, 1] / allm[:, 0]
            weights = np.clip(speed / speed.mean(), 0.8, 1.25)     # one odd measurement must not empty a shard
            cur.clear()
            ens.reshard(weights)
            attach()
    n = ens.n
    d = ens._dev
    pristine = cur["pristine"]

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    t_wall = time.perf_counter()
    for k in range(steps):
        one_step(evs[k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms_kernel = sum(e[0].elapsed_time(e[1]) for e in evs)
    ms_collect = sum(e[1].elapsed_time(e[2]) for e in evs)
    ms = ms_kernel + ms_collect
    out = d.out
    cnt = out["counters"].to(torch.int64).sum(0)
    nrows = out["nrows"].to(torch.int64).sum()
    failed = torch.nonzero(out["status"] != 1).flatten()[:8].cpu().tolist()
    stats = torch.tensor([float(cnt[1]), float(cnt[2]), float(nrows) - n, float((out["status"] == 1).sum()), float(n)],
                         dtype=torch.float64, device=dev)
    times = torch.tensor([ms, ms_kernel, ms_collect, float(clocks.get("sm_mhz") or 0), float(clocks.get("power_w") or 0)],
                         dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # one more collection outside the timed region, with events between its parts
        dist.barrier()
        tm = ens.gather(nbins=64, profile=True)["timing_ms"]
        times = torch.cat([times, torch.tensor([tm["pack_hist"], tm["allgather_unshard"], tm["allreduce"]], dtype=torch.float64, device=dev)])
        allt = torch.empty((world, 8), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt.view(-1), times)
        dist.all_reduce(stats)
        ms = float(allt[:, 0].max())
        per_rank = {"kernel_ms_per_step": [round(float(v) / steps, 3) for v in allt[:, 1]],
                    "collect_ms_per_step": [round(float(v) / steps, 3) for v in allt[:, 2]],
                    "sm_mhz": [float(v) for v in allt[:, 3]], "power_w": [float(v) for v in allt[:, 4]],
                    "collect_split_ms_after_barrier": {"pack_hist_kernel": [round(float(v), 3) for v in allt[:, 5]],
                                                       "allgather_plus_unshard": [round(float(v), 3) for v in allt[:, 6]],
                                                       "allreduce": [round(float(v), 3) for v in allt[:, 7]]},
                    "note": "collect = pack/histogram kernel + NCCL all-gather + all-reduce, including the wait for the slowest rank"}
    nstep, naccpt, ncalls, nok, nall = (float(stats[i]) for i in range(5))
    ms_per_step = ms / steps
    rec = dict(value=nstep / (ms_per_step * 1e-3), ms_per_step=ms_per_step, nstep=nstep, naccpt=naccpt, ncalls=ncalls,
               failures=int(nall - nok), failed_members_rank0=[int(ens._plan.indices(rank)[int(i)]) for i in failed], clocks=clocks,
               launches=int(launches), wall=wall, per_rank=per_rank, n=n, n_total=n_total,
               kernel_ms_rank0=ms_kernel / steps,
               shards=("none (one GPU)" if world == 1 else "round-robin" if not rebalance else
                       f"time-weighted (rapt_b200/dist.py:ShardPlan; {int(args.rebalance)} calibration rounds of 3 steps each before the warm-up, outside the timed region)"))
    if per_rank is not None:
        per_rank["shard_sizes"] = ens._plan.sizes()
        if shard_history:
            per_rank["shard_calibration"] = shard_history

    # ---- end to end through the host-buffer C ABI (pinned inputs, H2D + D2H inside the timed region)
    if want_e2e:
        import ctypes as C
        from rapt_b200._lib import ptr, check
        lib = _lib.load()
        ncol = len(pristine)
        pin = [c.cpu().pin_memory() for c in pristine]
        hwork = [torch.empty_like(c).pin_memory() for c in pin]
        ex = {k: v.cpu().pin_memory() for k, v in d.extras.items()}
        hout = dict(nrows=np.zeros(n, np.int32), nstored=np.zeros(n, np.int32), counters=np.zeros((n, 4), np.int32),
                    status=np.zeros(n, np.int32), tcur=np.zeros(n), dt=np.zeros(n))
        f_desc = ens.field.device_descriptor()
        p_desc = engine.snapshot_params(None, False, arith=args.arith, **(PARAMS if workload == "particle" else {}))
        if workload != "particle":
            ex["dt"] = torch.full((n,), GC_DT[workload], dtype=torch.float64).pin_memory()

        def host_step():
            for w, p0 in zip(hwork, pin):
                w.copy_(p0)
            if workload == "particle":
                check(lib.rapt_b200_particle_advance(
                    C.byref(f_desc), C.byref(p_desc), C.c_int64(n), *[ptr(w.numpy()) for w in hwork], ptr(ex["mass"].numpy()),
                    ptr(ex["charge"].numpy()), C.c_double(args.delta), C.c_int64(0), C.c_int64(0), None, ptr(hout["nrows"]),
                    ptr(hout["nstored"]), ptr(hout["counters"]), ptr(hout["status"]), ptr(hout["tcur"]), ptr(hout["dt"])))
            else:
                check(lib.rapt_b200_gc_advance(
                    C.byref(f_desc), C.byref(p_desc), C.c_int(0), C.c_int64(n), *[ptr(w.numpy()) for w in hwork],
                    ptr(ex["mu"].numpy()), ptr(ex["v"].numpy()), ptr(ex["mass"].numpy()), ptr(ex["charge"].numpy()),
                    ptr(ex["dt"].numpy()), C.c_double(args.delta), C.c_int64(0), C.c_int64(0), None, ptr(hout["nrows"]),
                    ptr(hout["nstored"]), ptr(hout["counters"]), ptr(hout["status"]), ptr(hout["tcur"])))
        host_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            host_step()
            print(f"[e2e {workload}] host-buffer step {1e3 * (time.perf_counter() - ts):.1f} ms", file=sys.stderr)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        st2 = torch.tensor([float(hout["counters"][:, 1].astype(np.int64).sum())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(st2)
        n_in = ncol + len(ex)
        rec["e2e"] = {"value": float(st2[0]) * steps / float(tt[0]), "unit": "particle-steps/s",
                      "h2d_bytes_per_step": n * n_in * 8,
                      "d2h_bytes_per_step": n * (ncol * 8 + (2 if workload == "particle" else 1) * 8 + 4 * 4 + 3 * 4)}
    cur.clear()
    del ens, pristine, flush
    torch.cuda.empty_cache()
    return rec


def time_adaptive(args, n, dev, steps, warmup):
    """Config 4 through rapt_b200_adaptive_advance (host-pointer ABI: constructor arguments in, final states out; the
    epoch loop runs on the device).  Steps/s over the device time of the epoch loop, roofline from the executed mix."""
    import torch
    from rapt_b200 import engine, fields, synth, _lib
    ic = synth.config4_speiser(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    f = fields.Parabolic()
    recs = []
    l0 = 0
    for k in range(warmup + steps):
        if k == warmup:
            l0 = _lib.launch_count()
        t = time.perf_counter()
        r = engine.adaptive_advance(f, pos, vel, ic["t0"], ic["mass"], ic["charge"], ADAPTIVE["delta"], ADAPTIVE["gc_dt"],
                                    store_every=0, max_rows=0, arith=args.arith, **ADAPTIVE["params"])
        wall = time.perf_counter() - t
        if k >= warmup:
            recs.append((r["stats"], wall, r))
    launches = _lib.launch_count() - l0
    st = {k: float(np.mean([a[0][k] for a in recs])) for k in recs[0][0]}
    wall = float(np.mean([a[1] for a in recs]))
    r = recs[-1][2]
    nstep = st["steps_particle"] + st["steps_gc"]
    flops = (algorithmic_flops(st["steps_particle"], st["accepted_particle"], st["calls_particle"], f_rhs=18 + 3)
             + gc_flops("adaptive", st["steps_gc"], st["calls_gc"]))
    nseg = r["nseg"]
    return dict(st=st, wall=wall, nstep=nstep, flops=flops, launches=int(launches), ok=int((r["status"] == 1).sum()),
                nseg_max=int(nseg.max()), nseg_mean=float(nseg.mean()), n=n,
                nseg_hist={str(int(k)): int(v) for k, v in zip(*np.unique(np.minimum(nseg, 12), return_counts=True))})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="particle", choices=["particle", "gc", "belt", "adaptive"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU,
                    help="tracers per GPU (--scaling weak) or in total (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--rebalance", type=int, default=2,
                    help="N > 1: calibration rounds of time-weighted shards before the warm-up (0: round-robin shards)")
    ap.add_argument("--work-order", default="predicted", choices=["predicted", "previous"],
                    help="particle workload: longest-first order from the initial state (predicted steps) or from the "
                         "previous advance's step counts of the device-resident ensemble (sort_by_work = 2)")
    ap.add_argument("--delta", type=float, default=DELTA)
    ap.add_argument("--cpu-sample", type=int, default=8192, help="tracers of the C oracle-port sample")
    ap.add_argument("--ref-sample", type=int, default=0, help="tracers per step of the Python-reference sample (0: sized for ~150 s)")
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the gc / belt / adaptive sub-records (extra.workloads)")
    args = ap.parse_args()
    WORK_ORDER["value"] = 2 if args.work_order == "previous" else 1
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    fd_out = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from rapt_b200 import engine, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    _lib.init(local)
    fp64_peak, _ = engine.fp64_peak(1 << 15)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def roofline(flops_per_gpu_step, ms_per_step, nstep_per_gpu, traffic=(None, None), io_bytes=None):
        achieved = flops_per_gpu_step / (ms_per_step * 1e-3) / 1e12
        r = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
             "traffic": traffic[0], "traffic_source": traffic[1],
             "peak_source": "FP64 DFMA microbenchmark measured in this run (rapt_b200_fp64_peak); MEASURED_PEAKS.json has HBM and bf16 only",
             "algorithmic_flop_per_step": flops_per_gpu_step / max(nstep_per_gpu, 1)}
        if io_bytes:
            r["hbm"] = {"algorithmic_bytes": io_bytes, "achieved_GBps": io_bytes / (ms_per_step * 1e-3) / 1e9,
                        "peak_GBps": peaks.get("hbm_gbs"), "note": "state in/out only: compute-bound kernel"}
        return r

    res = None
    if args.workload == "adaptive":
        a = time_adaptive(args, args.n_per_gpu, dev, args.steps, args.warmup)
        ms = a["st"]["ms_epochs"]
        res = {"metric": "particle-steps/s", "value": a["nstep"] / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": 1,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": WORKLOAD_NAME["adaptive"], "particles_per_gpu": a["n"], "arith": args.arith,
                          "l2": "state (9 columns x n) is re-uploaded every step; epochs stream > L2 of state",
                          "epochs": a["st"]["epochs"], "tracers_ok": a["ok"], "segments_max": a["nseg_max"],
                          "segments_mean": a["nseg_mean"], "segments_hist": a["nseg_hist"],
                          "steps_particle_mode": a["st"]["steps_particle"], "steps_gc_mode": a["st"]["steps_gc"],
                          "kernel_ms": {k: a["st"][k] for k in ("ms_particle", "ms_gc", "ms_switch", "ms_epochs")},
                          "tracer_launches": [a["st"]["tracer_launches_particle"], a["st"]["tracer_launches_gc"]]},
               "gpu_launches": a["launches"], "wall_s_per_step_host_abi": a["wall"],
               "roofline": roofline(a["flops"], ms, a["nstep"]),
               "e2e": {"value": a["nstep"] / a["wall"], "unit": "particle-steps/s", "h2d_bytes_per_step": a["n"] * 9 * 8,
                       "d2h_bytes_per_step": a["n"] * (8 * 8 + 4 * 4 + 4 * 4)}}
    else:
        is_gc = args.workload != "particle"
        r = time_workload(args, args.workload, args.n_per_gpu, world, rank, dev, args.steps, args.warmup,
                          want_e2e=not args.no_e2e)
        if rank == 0:
            flops = (gc_flops(args.workload, r["nstep"], r["ncalls"]) if is_gc
                     else algorithmic_flops(r["nstep"], r["naccpt"], r["ncalls"])) * (r["n"] / r["n_total"])
            n = r["n"]                    # rank 0's shard (its kernel time and its share of the flops make the roofline)
            io_bytes = n * ((10 * 8 + 5 * 8 + 8 + 7 * 4) if is_gc else (9 * 8 + 7 * 8 + 2 * 8 + 7 * 4))
            # the roofline is the KERNEL's: rank 0's kernel time per step, its share of the flops
            res = {
                "metric": "particle-steps/s", "value": r["value"], "unit": "particle-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAME[args.workload], "particles_per_gpu": r["n_total"] // world, "particles_total": r["n_total"],
                           "delta_s": args.delta, "arith": args.arith, "shards": r["shards"], "work_order": args.work_order,
                           "l2": "512 MiB flush write between steps (inputs < L2)",
                           "particle_steps_per_bench_step": r["nstep"], "accepted": r["naccpt"], "output_rows": r["ncalls"],
                           # members whose row loop ended on scipy's nsteps = 500 limit, as the reference's does for the
                           # same member (tests/test_gpu_particle.py::test_config2_solver_failure_member)
                           "solver_failures": r["failures"], "failed_members_rank0": r["failed_members_rank0"],
                           "collective": ("ensemble.gather(): pack+histogram kernel, NCCL all_gather(final states, in place) + "
                                          "all_reduce(histogram, invariant sums)") if world > 1 else "none"},
                "clocks": r["clocks"],
                "gpu_launches": r["launches"],
                "wall_s_timed_region": r["wall"],
                "roofline": roofline(flops, r["kernel_ms_rank0"], r["nstep"] * (r["n"] / r["n_total"]),
                                     measured_traffic(args.workload, r["n_total"] // world, args.delta), io_bytes),
            }
            if r["per_rank"]:
                res["per_rank"] = r["per_rank"]
            if "e2e" in r:
                res["e2e"] = r["e2e"]
        # ---- the other named configurations, measured in the same run (N = 1 only): sub-records for the driver's file
        if world == 1 and not args.no_extra and args.workload == "particle" and args.n_per_gpu == N_PER_GPU:
            extra = {}
            for w in ("gc", "belt"):
                x = time_workload(args, w, N_PER_GPU, 1, 0, dev, 2, 3, want_e2e=False)
                fl = gc_flops(w, x["nstep"], x["ncalls"])
                extra[w] = {"workload": WORKLOAD_NAME[w], "particles": x["n"], "delta_s": args.delta, "value": x["value"],
                            "unit": "particle-steps/s", "ms_per_step": x["ms_per_step"], "steps": 2, "warmup": 3,
                            "gpu_launches": x["launches"], "solver_failures": x["failures"],
                            "roofline": roofline(fl, x["kernel_ms_rank0"], x["nstep"],
                                                 measured_traffic(w, x["n"], args.delta))}
            if args.work_order == "predicted":
                # the same ensemble in steady state of a multi-call run: every advance ordered by the previous advance's step
                # counts (what ParticleEnsemble.advance does by default on device-resident state; the headline above is the
                # cold call, ordered by the a-priori estimate)
                WORK_ORDER["value"] = 2
                x = time_workload(args, "particle", N_PER_GPU, 1, 0, dev, 3, 3, want_e2e=False)
                WORK_ORDER["value"] = 1
                fl = algorithmic_flops(x["nstep"], x["naccpt"], x["ncalls"])
                extra["particle_work_order_previous"] = {
                    "workload": WORKLOAD_NAME["particle"] + "; longest-first order from the previous advance's step counts "
                                "(sort_by_work = 2), not from the a-priori estimate", "particles": x["n"], "delta_s": args.delta,
                    "value": x["value"], "unit": "particle-steps/s", "ms_per_step": x["ms_per_step"], "steps": 3, "warmup": 3,
                    "gpu_launches": x["launches"], "roofline": roofline(fl, x["kernel_ms_rank0"], x["nstep"])}
                for w in ("gc", "belt"):                  # guiding centres: the cold call above is ordered by k_key_gc's estimate
                    WORK_ORDER["value"] = 2
                    try:
                        x = time_workload(args, w, N_PER_GPU, 1, 0, dev, 2, 3, want_e2e=False)
                    except Exception as ex:               # an optional sub-record must not cost the headline line
                        extra[w + "_work_order_previous"] = {"error": str(ex)[:300]}
                        continue
                    finally:
                        WORK_ORDER["value"] = 1
                    extra[w + "_work_order_previous"] = {
                        "workload": WORKLOAD_NAME[w] + "; longest-first order from the previous advance's step counts",
                        "particles": x["n"], "delta_s": args.delta, "value": x["value"], "unit": "particle-steps/s",
                        "ms_per_step": x["ms_per_step"], "steps": 2, "warmup": 3, "gpu_launches": x["launches"],
                        "roofline": roofline(gc_flops(w, x["nstep"], x["ncalls"]), x["kernel_ms_rank0"], x["nstep"])}
            a = time_adaptive(args, N_PER_GPU, dev, 1, 1)
            ms = a["st"]["ms_epochs"]
            extra["adaptive"] = {"workload": WORKLOAD_NAME["adaptive"], "particles": a["n"], "value": a["nstep"] / (ms * 1e-3),
                                 "unit": "particle-steps/s", "ms_per_step": ms, "steps": 1, "warmup": 1,
                                 "epochs": a["st"]["epochs"], "tracers_ok": a["ok"], "segments_max": a["nseg_max"],
                                 "kernel_ms": {k: a["st"][k] for k in ("ms_particle", "ms_gc", "ms_switch", "ms_epochs")},
                                 "e2e_value": a["nstep"] / a["wall"], "gpu_launches": a["launches"],
                                 "roofline": roofline(a["flops"], ms, a["nstep"])}
            res["extra"] = {"workloads": extra}

    if rank == 0 and res is not None:
        if not args.no_cpu_baseline and world == 1:
            w = args.workload if args.workload != "adaptive" else "particle"
            res["cpu_baseline"] = cpu_baseline_record(w, args.delta, args.cpu_sample if w == "particle" else min(args.cpu_sample, 2048))
        sys.stdout.flush()
        os.dup2(fd_out, 1)
        print(json.dumps(res), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
