#!/usr/bin/env python
"""bench.py -- headline benchmark of the particle-advance hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload particle|gc]

Metric: particle-steps/s (fp64) = attempted Runge-Kutta steps of all particles per second of device
time, whole job.  Workload at N = 1: BASELINE.json configs[1] -- 1 M protons in EarthDipole, full orbit,
random energies 0.1-10 MeV and pitch angles (rapt_b200/synth.py, seed 20260201), Particle.advance(10 s),
cyclotronresolution 20.  One bench "step" = one such advance of the whole ensemble from the same initial
state.  N > 1: weak scaling, 1 M protons per GPU (rank r takes members r, r+N, ... of the N-million
ensemble), no data-path collective; the final states are all-gathered and an energy histogram
all-reduced over NCCL inside the timed region.

`value` is device-resident (inputs already in HBM, CUDA events on the launching stream, max over
ranks); `e2e` is the same metric through the host-buffer C ABI (pinned host inputs, H2D + D2H inside
the timed region).  `roofline` is against the FP64 DFMA peak measured in this run (MEASURED_PEAKS.json
has no FP64 figure), with the algorithmic flop count of DESIGN.md.  `cpu_baseline` / `--impl reference`
time the CPU oracle port (oracle/, OpenMP over all host cores) on a bounded sample of the same ensemble.
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
F_RHS = 38            # Lorentz 18 + EarthDipole 20 flop (E == 0 compiled out)
F_STAGE = 6 * 158 + 20


def algorithmic_flops(nstep, naccpt, ncalls):
    """DESIGN.md 'flop accounting': 11 RHS per attempted step + 1 per accepted step, stage sums + error
    norm + controller per attempted step, HINIT (1 extra RHS + norms) per solver call (= output row)."""
    return nstep * (11 * F_RHS + F_STAGE) + naccpt * F_RHS + ncalls * (F_RHS + 60)


WORKLOAD_NAME = {
    "particle": "config2: 1M protons per GPU / EarthDipole / Particle.advance(10 s) / cyclotronresolution 20 / "
                "KE 0.1-10 MeV / seed 20260201",
    "gc": "config3: electrons / DoubleDipole / GuidingCenter.advance (TaoChanBrizard) / GCtimestep 0.1 / KE 50 keV-1 MeV / "
          "seed 20260301",
    "belt": "config5: electrons / VarEarthDipole(0.1, 10 s) / GuidingCenter.advance / GCtimestep 0.05 / seed 20260501",
}


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
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
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
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_ensemble(world, rank, n_per_gpu, workload):
    from rapt_b200 import synth, engine
    n_total = n_per_gpu * world
    if workload == "particle":
        ic = synth.config2_protons(n_total)
        sl = slice(rank, n_total, world)
        vel = np.column_stack([ic["vx"][sl], ic["vy"][sl], ic["vz"][sl]])
        mom = engine.particle_momentum(vel, ic["mass"][sl])
        cols = [ic["t0"][sl], ic["x"][sl], ic["y"][sl], ic["z"][sl], mom[:, 0], mom[:, 1], mom[:, 2]]
        return dict(cols=[np.ascontiguousarray(c) for c in cols], mass=np.ascontiguousarray(ic["mass"][sl]),
                    charge=np.ascontiguousarray(ic["charge"][sl]))
    if workload in ("gc", "belt"):
        gen = synth.config3_electrons if workload == "gc" else synth.config5_belt
        if n_total > 20_000_000:
            # 100 M-tracer ensembles: every rank draws its own shard (seed + rank) instead of slicing one
            # global draw, so that no process has to hold the whole ensemble in host memory
            ic = gen(n_per_gpu, seed=(20260301 if workload == "gc" else 20260501) + 1000 * rank)
            sl = slice(0, n_per_gpu)
        else:
            ic = gen(n_total)
            sl = slice(rank, n_total, world)
        field = gc_field(workload)
        pos = np.column_stack([ic["x"][sl], ic["y"][sl], ic["z"][sl]])
        ppar, mu = engine.gc_construct(field, ic["t0"][sl], pos, ic["v"][sl], ic["pa"][sl], ic["mass"][sl], arith="fast")
        cols = [ic["t0"][sl], pos[:, 0], pos[:, 1], pos[:, 2], ppar]
        return dict(cols=[np.ascontiguousarray(c) for c in cols], mass=np.ascontiguousarray(ic["mass"][sl]),
                    charge=np.ascontiguousarray(ic["charge"][sl]), mu=mu, v=np.ascontiguousarray(ic["v"][sl]),
                    dt=np.full(len(mu), GC_DT[workload]))
    raise ValueError(workload)


GC_DT = {"gc": 0.1, "belt": 0.05}          # params["GCtimestep"] of configs 3 and 5


def gc_field(workload):
    from rapt_b200 import fields
    return fields.DoubleDipole() if workload == "gc" else fields.VarEarthDipole(0.1, 10)


def gc_flops(workload, nstep, ncalls):
    """6 RHS + stage sums per attempted step, HINIT per row; RHS = n_B field evaluations + 176.
    Belt (VarEarthDipole) counts what the fast kernel executes, not the reference's 9 evaluations with a sine
    each: 7 dipole evaluations scaled by one time factor (~25 flop) per RHS; db/dt is identically zero."""
    f_b, n_b, f_t = (48, 7, 0) if workload == "gc" else (23, 7, 25)
    f_rhs = n_b * f_b + f_t + 176
    return nstep * (6 * f_rhs + 4 * 64 + 20) + ncalls * (f_rhs + 40)


def cpu_sample_gc(workload, n_sample, delta, nthreads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rapt_b200 import synth
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


def cpu_sample(n_sample, delta, nthreads):
    """The CPU oracle port on the first n_sample members of the same ensemble."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rapt_b200 import synth
    ic = synth.config2_protons(n_sample)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
    f, p = O.make_field("EarthDipole"), O.make_params(**PARAMS)
    t = time.perf_counter()
    o = O.particle_advance(f, p, st, ic["mass"], ic["charge"], delta, store_every=0, nthreads=nthreads)
    el = time.perf_counter() - t
    steps = int(o["counters"][:, 1].sum())
    return steps / el, steps, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python
    and cannot travel to the GPU box, so this is the C oracle port (bit-exact against the reference's
    own trajectories, tests/test_oracle_golden.py), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample
    vals = []
    for i in range(args.warmup + args.steps):
        v, steps, el = (cpu_sample(n_sample, DELTA, cores) if args.workload == "particle"
                        else cpu_sample_gc(args.workload, n_sample, DELTA, cores))
        if i >= args.warmup:
            vals.append((v, steps, el))
    steps = sum(s for _, s, _ in vals); el = sum(e for _, _, e in vals)
    value = steps / el
    sample = f"first {n_sample} tracers of the {WORKLOAD_NAME[args.workload].split(':')[0]} ensemble, advance({DELTA} s) each, OpenMP x{cores}"
    print(json.dumps({
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(len(vals), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[args.workload], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="particle", choices=["particle", "gc", "belt"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--delta", type=float, default=DELTA)
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    fd_out = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from rapt_b200 import engine, fields, _lib

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
    is_gc = args.workload != "particle"
    field = gc_field(args.workload) if is_gc else fields.EarthDipole()
    n = args.n_per_gpu
    ens = make_ensemble(world, rank, n, args.workload)
    ncol = len(ens["cols"])
    pristine = [torch.tensor(c, device=dev) for c in ens["cols"]]
    mass = torch.tensor(ens["mass"], device=dev); charge = torch.tensor(ens["charge"], device=dev)
    if is_gc:
        gmu = torch.tensor(ens["mu"], device=dev); gv = torch.tensor(ens["v"], device=dev); gdt = torch.tensor(ens["dt"], device=dev)
    out = engine.alloc_outputs(n, dev)
    work = [torch.empty_like(c) for c in pristine]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    gathered = torch.empty((world, n, ncol), dtype=torch.float64, device=dev) if world > 1 else None
    nbins = 64
    hist_edges = torch.logspace(4.5, 7.5, nbins + 1, dtype=torch.float64, device=dev)

    def one_step(timed_events=None):
        for w, p0 in zip(work, pristine):
            w.copy_(p0)
        flush.fill_(1)                                                   # L2 flush between steps
        if timed_events is not None:
            timed_events[0].record()
        if is_gc:
            engine.gc_advance_dev(field, work, gmu, gv, mass, charge, gdt, args.delta, out, arith=args.arith)
        else:
            engine.particle_advance_dev(field, work, mass, charge, args.delta, out, arith=args.arith, **PARAMS)
        if world > 1:
            fin = torch.stack(work, dim=1)
            dist.all_gather_into_tensor(gathered.view(-1), fin.view(-1))
            if is_gc:      # diagnostic: histogram of the radial distance (drift-shell occupation)
                h = torch.histc(torch.sqrt((fin[:, 1:4] ** 2).sum(1)) / 6378137.0, bins=nbins, min=0.0, max=16.0)
            else:          # diagnostic: kinetic-energy histogram
                p2 = (fin[:, 4:7] ** 2).sum(1)
                ke_ev = (torch.sqrt(1 + p2 / (mass * 299792458.0) ** 2) - 1) * mass * 299792458.0 ** 2 / 1.602176565e-19
                h = torch.histc(torch.log10(ke_ev), bins=nbins, min=4.5, max=7.5)
            dist.all_reduce(h)
        if timed_events is not None:
            timed_events[1].record()

    fp64_peak, _ = engine.fp64_peak(1 << 15)
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall = time.perf_counter()
    for k in range(args.steps):
        one_step(evs[k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    cnt = out["counters"].to(torch.int64).sum(0)
    nrows = out["nrows"].to(torch.int64).sum()
    stats = torch.tensor([float(cnt[1]), float(cnt[2]), float(nrows) - n, float((out["status"] == 1).sum()), ms],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
        ms = float(mx[4])
    nstep, naccpt, ncalls, nok = (float(stats[i]) for i in range(4))
    ms_per_step = ms / args.steps
    value = nstep / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned inputs, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e and not is_gc:
        pin = [torch.tensor(c).pin_memory() for c in ens["cols"]]
        pm = torch.tensor(ens["mass"]).pin_memory(); pq = torch.tensor(ens["charge"]).pin_memory()
        hwork = [torch.empty_like(c).pin_memory() for c in pin]
        hout = dict(nrows=np.zeros(n, np.int32), nstored=np.zeros(n, np.int32), counters=np.zeros((n, 4), np.int32),
                    status=np.zeros(n, np.int32), tcur=np.zeros(n), dt=np.zeros(n))
        import ctypes as C
        from rapt_b200._lib import ptr, check
        f_desc = field.device_descriptor(); p_desc = engine.snapshot_params(None, False, arith=args.arith, **PARAMS)
        lib = _lib.load()

        def host_step():
            for w, p0 in zip(hwork, pin):
                w.copy_(p0)
            check(lib.rapt_b200_particle_advance(
                C.byref(f_desc), C.byref(p_desc), C.c_int64(n), *[ptr(w.numpy()) for w in hwork], ptr(pm.numpy()), ptr(pq.numpy()),
                C.c_double(args.delta), C.c_int64(0), C.c_int64(0), None, ptr(hout["nrows"]), ptr(hout["nstored"]),
                ptr(hout["counters"]), ptr(hout["status"]), ptr(hout["tcur"]), ptr(hout["dt"])))
        host_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ts = time.perf_counter()
            host_step()
            print(f"[e2e] host-buffer step {1e3 * (time.perf_counter() - ts):.1f} ms", file=sys.stderr)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_steps = float(hout["counters"][:, 1].astype(np.int64).sum())
        st2 = torch.tensor([e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(st2)
        e2e = {"value": float(st2[0]) * args.steps / float(tt[0]), "unit": "particle-steps/s",
               "h2d_bytes_per_step": n * 9 * 8, "d2h_bytes_per_step": n * (7 * 8 + 2 * 8 + 4 * 4 + 3 * 4)}

    if not args.no_e2e and is_gc:
        # guiding-centre workloads: host-buffer C ABI (numpy in, numpy out; H2D + D2H inside the call)
        st_host = np.column_stack(ens["cols"])
        engine.gc_advance(field, st_host, ens["mu"], ens["v"], ens["mass"], ens["charge"], ens["dt"], args.delta,
                          store_every=0, arith=args.arith)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oh = engine.gc_advance(field, st_host, ens["mu"], ens["v"], ens["mass"], ens["charge"], ens["dt"], args.delta,
                                   store_every=0, arith=args.arith)
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        st2 = torch.tensor([float(oh["counters"][:, 1].astype(np.int64).sum())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(st2)
        e2e = {"value": float(st2[0]) * args.steps / float(tt[0]), "unit": "particle-steps/s",
               "h2d_bytes_per_step": n * 10 * 8, "d2h_bytes_per_step": n * (5 * 8 + 8 + 4 * 4 + 3 * 4)}

    if rank == 0:
        flops = gc_flops(args.workload, nstep, ncalls) if is_gc else algorithmic_flops(nstep, naccpt, ncalls)
        achieved = flops / world / (ms_per_step * 1e-3) / 1e12     # per GPU: the kernel's own rate
        io_bytes = n * ((10 * 8 + 5 * 8 + 8 + 7 * 4) if is_gc else (9 * 8 + 7 * 8 + 2 * 8 + 7 * 4))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        res = {
            "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[args.workload], "particles_per_gpu": n, "delta_s": args.delta, "arith": args.arith,
                       "l2": "512 MiB flush write between steps (inputs < L2)",
                       "particle_steps_per_bench_step": nstep, "accepted": naccpt, "output_rows": ncalls,
                       "solver_failures": int(n * world - nok),     # members whose row loop ended on scipy's nsteps=500
                                                                   # limit, exactly as the reference's does (checked vs the oracle)
                       "collective": "all_gather(final state) + all_reduce(KE histogram) over NCCL" if world > 1 else "none"},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "wall_s_timed_region": wall,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this configuration, one
                         # launch, from `ncu --set full` (profiles/r1_particle_v8_bench_kernel_ncu_summary.txt):
                         # 452 MB + 289 MB.  4x the algorithmic state I/O because tracers are fetched in
                         # longest-first order, i.e. as scattered 8-byte accesses (32-byte sectors); at 0.27 s
                         # per launch that is 3 GB/s and irrelevant to this compute-bound kernel.
                         "traffic": 740606720 if (n == N_PER_GPU and args.delta == DELTA and not is_gc) else None,
                         "peak_source": "FP64 DFMA microbenchmark measured in this run (rapt_b200_fp64_peak); "
                                        "MEASURED_PEAKS.json has HBM and bf16 only",
                         "algorithmic_flop_per_step": flops / nstep,
                         "hbm": {"algorithmic_bytes": io_bytes, "achieved_GBps": io_bytes / (ms_per_step * 1e-3) / 1e9,
                                 "peak_GBps": peaks.get("hbm_gbs"), "note": "state in/out only: compute-bound kernel"}},
        }
        if e2e is not None:
            res["e2e"] = e2e
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, steps, el = (cpu_sample_gc(args.workload, min(args.cpu_sample, 2048), args.delta, cores) if is_gc
                            else cpu_sample(args.cpu_sample, args.delta, cores))
            # the unmodified Python reference cannot travel to the GPU box (/root/reference is absent there); its rate was
            # measured in the build container (README Particle case 1.4-1.6e3, GuidingCenter 0.46e3
            # particle-steps/s on one core) and is quoted here for scale only -- the C port above is ~10^3 x faster per core
            py_ref = {"value": 0.46e3 if is_gc else 1.4e3, "unit": "particle-steps/s per core", "measured": "build container, SURVEY.md section 6; not re-timed in this run"}
            res["cpu_baseline"] = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port", "python_reference": py_ref,
                                   "sample": f"first {min(args.cpu_sample, 2048) if is_gc else args.cpu_sample} tracers of the same ensemble, advance({args.delta} s), "
                                             f"{steps} steps in {el:.1f} s, C oracle port with OpenMP"}
        sys.stdout.flush()
        os.dup2(fd_out, 1)
        print(json.dumps(res), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
