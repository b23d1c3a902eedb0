"""Ensemble entry points: millions of independent tracers per call (no reference counterpart; the
reference advances one Python object at a time).

Each class mirrors the constructor of its single-tracer namesake with array arguments and exposes
`advance(delta)`.  State can live on the host (numpy; every `advance` copies in and out through the
host-pointer C ABI) or on the device (`.cuda()`: torch CUDA tensors updated in place through the
device-pointer C ABI, no host traffic) -- the latter is what large runs and multi-GPU sharding use.
"""
import numpy as np

from . import c, params
from . import engine


def _arr(a, n=None):
    a = np.asarray(a, dtype=np.float64)
    if n is not None:
        a = np.broadcast_to(a, (n,))
    return np.ascontiguousarray(a).copy()


class _Resident:
    """Residency, sharding and result collection shared by ParticleEnsemble and GuidingCenterEnsemble.

    Multi-GPU: one process per GPU under torchrun.  Every rank builds the SAME ensemble and
    calls `.shard()`: rank r keeps members r, r+W, r+2W, ... (round-robin, so every GPU sees the same mix of orbit
    lengths; with `weights`, runs of a 4096-member period proportional to each GPU's speed: dist.ShardPlan) on
    cuda:LOCAL_RANK.  `advance()` then runs the same kernels on the shard with no data-path collective.
    `.gather()` packs the final states and bins the diagnostics in one kernel pass, all-gathers the rows (NCCL over
    NVLink, in place into this rank's slot) and sum-all-reduces the fixed-size histogram / invariant sums."""

    _MEMBER_ARRAYS = ()          # per-member host arrays a shard slices
    _DIAG = "ke"

    def shard(self, group=None, device=None, weights=None, keep_full=False):
        """Keep this rank's members and move them to its GPU.  weights: one positive number per rank (relative speed of
        its GPU); None = round-robin.  keep_full=True keeps the whole ensemble's host arrays so that `reshard()` can cut
        the shards again (e.g. after measuring every rank's kernel time: bench.py --rebalance)."""
        from . import dist as rd
        world, rank = rd.world_rank(group)
        full = getattr(self, "_full", None)
        if full is None:
            full = {name: getattr(self, name) for name in self._MEMBER_ARRAYS if getattr(self, name, None) is not None}
            self.n_total = self.n
        self._group, self.world, self.rank = group, world, rank
        self._plan = rd.ShardPlan(self.n_total, world, weights)
        if world > 1:
            idx = self._plan.indices(rank)
            sl = rd.shard_slice(self.n_total, world, rank) if self._plan.uniform else idx
            for name, a in full.items():
                setattr(self, name, np.ascontiguousarray(a[sl]))
            self.n = len(idx)
        self._full = full if (keep_full and world > 1) else None
        return self.cuda(device or rd.local_device())

    def reshard(self, weights):
        """Cut the shards again with new weights from the ensemble as it was when `.shard(keep_full=True)` was called
        (the device-resident results of the old shards are dropped)."""
        if getattr(self, "_full", None) is None:
            if getattr(self, "world", 1) == 1:
                return self
            raise RuntimeError("reshard() needs .shard(keep_full=True)")
        self._dev = None
        return self.shard(self._group, None, weights, keep_full=True)

    def load_state(self, cols):
        """Overwrite the device-resident state columns with the given CUDA tensors (e.g. to restart from the same
        initial conditions); stream-ordered copies, no host traffic."""
        for dst, src in zip(self._dev.cols, cols):
            dst.copy_(src)
        return self

    def _bind(self, device):
        """One process drives one GPU: bind the library and torch to the device the state goes to."""
        import torch
        self._host_counters = np.array(self.counters, dtype=np.int64)
        from . import _lib
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _lib.init(idx)
        torch.cuda.set_device(idx)
        return torch.device("cuda", idx)

    def _rows_buffer(self, store_every, max_rows):
        import torch
        d = self._dev
        if store_every > 0 and max_rows > 0:
            if d.rows is None or tuple(d.rows.shape) != (self.n, max_rows, 8):
                d.rows = torch.empty((self.n, max_rows, 8), dtype=torch.float64, device=d.device)
            d.rows_valid = True
            return d.rows
        d.rows_valid = False
        return None

    def pull(self):
        """Copy the device-resident results to the host attributes (state, status, tcur, counters, nrows and, when
        the last advance() stored rows, trajectory / nstored); the ensemble stays on the device."""
        import torch
        d = self._dev
        if d is None:
            return self
        torch.cuda.synchronize(d.device)
        self.state = np.column_stack([cc.cpu().numpy() for cc in d.cols])
        o = d.out
        self.status = o["status"].cpu().numpy(); self.tcur = o["tcur"].cpu().numpy()
        if "dt" in o and hasattr(self, "dt"):
            self.dt = o["dt"].cpu().numpy()
        self.last_counters = o["counters"].cpu().numpy().astype(np.int64)
        self.counters = self._host_counters + d.cum_counters.cpu().numpy()
        self.nrows = o["nrows"].cpu().numpy().astype(np.int64)
        if d.rows is not None and d.rows_valid:
            self.trajectory = d.rows.cpu().numpy(); self.nstored = o["nstored"].cpu().numpy()
        return self

    def cpu(self):
        if self._dev is not None:
            self.pull()
            self._dev = None
        return self

    def gather(self, nbins=64, lo=None, hi=None, kind=None, profile=False, sync=True):
        """Final states of the WHOLE ensemble in member order on every rank + all-reduced diagnostics.
        Returns dict(final=(n_total, ncol) CUDA tensor, hist=int64 [nbins], edges, stats=dict(ok, mean, var, outside)).
        kind 'ke': log10 kinetic energy [eV] (Particle.getke); 'r': radial distance [Re].
        profile=True adds timing_ms = device time of the pack+histogram kernel, the all-gather (+ un-interleave) and the
        all-reduce (CUDA events on the current stream).
        sync=False returns without touching the host: `stats` is then the raw CUDA tensor (ok count, sum q, sum q^2,
        outside) and nothing is synchronised, so a caller that advances again can queue its next launches behind the
        collectives (measured on 8 GPUs: the host round trip per step let the ranks drift apart by ~4 ms, which the
        slowest rank then paid inside the next all-gather: profiles/r2_multi_gpu.md)."""
        import torch
        from . import dist as rd
        d = self._dev
        if d is None:
            raise RuntimeError("gather() works on device-resident ensembles: call .shard() or .cuda() first")
        kind = kind or self._DIAG
        lo = (4.0 if kind == "ke" else 0.0) if lo is None else lo
        hi = (8.0 if kind == "ke" else 16.0) if hi is None else hi
        world, rank = getattr(self, "world", 1), getattr(self, "rank", 0)
        n_total = getattr(self, "n_total", self.n)
        plan = getattr(self, "_plan", None) or rd.ShardPlan(n_total, world)
        n_max = max(plan.sizes())
        ncol = len(d.cols)
        if d.gather_buf is None or tuple(d.gather_buf.shape) != (world, n_max, ncol):
            d.gather_buf = torch.zeros((world, n_max, ncol), dtype=torch.float64, device=d.device)
            d.final = torch.empty((n_total, ncol), dtype=torch.float64, device=d.device)
            d.stats = torch.zeros(4, dtype=torch.float64, device=d.device)
        if d.hist is None or d.hist.numel() != nbins:
            d.hist = torch.zeros(nbins, dtype=torch.int64, device=d.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if profile else None
        d.hist.zero_(); d.stats.zero_()
        if ev: ev[0].record()
        engine.final_diagnostics_dev(kind, d.cols, d.extras["mass"], d.out["status"], d.gather_buf[rank], nbins, lo, hi,
                                     d.hist, d.stats)
        if ev: ev[1].record()
        final = rd.gather_rows(d.gather_buf, rank, n_total, getattr(self, "_group", None), out=d.final, plan=plan)
        if ev: ev[2].record()
        rd.reduce_diagnostics(d.hist, d.stats, getattr(self, "_group", None))
        if ev: ev[3].record()
        if sync:
            st = d.stats.cpu().numpy()
            ok = max(st[0], 1.0)
            stats = dict(ok=int(st[0]), mean=st[1] / ok, var=st[2] / ok - (st[1] / ok) ** 2, outside=int(st[3]))
        else:
            stats = d.stats
        res = dict(final=final, hist=d.hist, edges=np.linspace(lo, hi, nbins + 1), kind=kind, stats=stats)
        if ev:
            ev[3].synchronize()
            res["timing_ms"] = dict(pack_hist=ev[0].elapsed_time(ev[1]), allgather_unshard=ev[1].elapsed_time(ev[2]),
                                    allreduce=ev[2].elapsed_time(ev[3]))
        return res


class _DeviceState:
    """Device-resident columns (torch CUDA float64 tensors) + output scratch."""

    def __init__(self, cols, extras, device):
        import torch
        self.device = torch.device(device)
        self.cols = [torch.as_tensor(np.ascontiguousarray(cc), device=self.device) for cc in cols]
        self.extras = {k: torch.as_tensor(np.ascontiguousarray(v), device=self.device) for k, v in extras.items()}
        self.out = engine.alloc_outputs(self.cols[0].numel(), self.device)
        self.rows = None; self.rows_valid = False
        self.pending_advances = 0
        # cumulative (nfcn, nstep, naccpt, nrejct) over the device-resident advance() calls, accumulated on the stream
        self.cum_counters = torch.zeros((self.cols[0].numel(), 4), dtype=torch.int64, device=self.device)
        self.gather_buf = None; self.final = None; self.hist = None; self.stats = None


class ParticleEnsemble(_Resident):
    """n full-orbit tracers: Particle(pos, vel, t0, mass, charge, field) with array arguments
    (reference constructor: rapt/Particle.py:59-109).

    pos, vel: (n,3); t0, mass, charge: scalars or (n,).  `state` is (n,7): t,x,y,z,px,py,pz.
    """

    def __init__(self, pos, vel, t0=0.0, mass=None, charge=None, field=None):
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        vel = np.asarray(vel, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        self.n = n
        self.mass = _arr(mass, n); self.charge = _arr(charge, n)
        self.field = field
        mom = engine.particle_momentum(vel, self.mass)               # Particle.py:106-107
        self.state = np.column_stack([_arr(t0, n), pos, mom])
        self.tcur = self.state[:, 0].copy()
        self.dt = np.zeros(n)
        self.counters = np.zeros((n, 4), dtype=np.int64)             # cumulative (nfcn, nstep, naccpt, nrejct)
        self.status = np.ones(n, dtype=np.int32)
        self.nrows = np.ones(n, dtype=np.int64)
        self.trajectory = None                                       # (n, max_rows, 8) of the last advance()
        self.nstored = None
        self.check_adiabaticity = False
        self._dev = None

    _MEMBER_ARRAYS = ("state", "mass", "charge", "tcur", "dt", "counters", "status", "nrows")
    _DIAG = "ke"

    # ---- residency
    def cuda(self, device="cuda:0"):
        self._dev = _DeviceState([self.state[:, i] for i in range(7)], dict(mass=self.mass, charge=self.charge),
                                 self._bind(device))
        return self

    # ---- the hot path
    def advance(self, delta, store_every=0, max_rows=0, **over):
        """Particle.advance(delta) for every member (rapt/Particle.py:230-309).
        store_every = k keeps every k-th output row (0: final state only) in `trajectory`."""
        if self._dev is not None:
            d = self._dev
            rows = self._rows_buffer(store_every, max_rows)
            # longest-first work order: from the second call on, the previous call's step counts of every member (still
            # in the device-resident counters) replace the a-priori estimate (rapt_params_t.sort_by_work = 2); measured on
            # config 2: 221.7 -> 210.8 ms per advance (profiles/r2_22_ab_work_order_previous.jsonl).  Scheduling only.
            over.setdefault("sort_by_work", 2)
            engine.particle_advance_dev(self.field, d.cols, d.extras["mass"], d.extras["charge"], float(delta), d.out,
                                        store_every=store_every, max_rows=max_rows, rows=rows,
                                        check_adiabaticity=self.check_adiabaticity, **over)
            d.cum_counters += d.out["counters"]    # stream-ordered, no synchronisation
            d.pending_advances += 1
            return self
        o = engine.particle_advance(self.field, self.state, self.mass, self.charge, float(delta), store_every=store_every,
                                    max_rows=max_rows, check_adiabaticity=self.check_adiabaticity, **over)
        self.state = o["state"]; self.tcur = o["tcur"]; self.dt = o["dt"]; self.status = o["status"]
        self.last_counters = o["counters"].astype(np.int64)
        self.counters += self.last_counters
        self.nrows = o["nrows"].astype(np.int64)
        self.trajectory = o["rows"]; self.nstored = o["nstored"]
        return self

    def member_trajectory(self, i):
        """(rows, 7) trajectory array of member i from the last advance(), as Particle.trajectory."""
        return self.trajectory[i, :self.nstored[i], :7]

    # ---- diagnostics (host)
    def getke(self):
        """Kinetic energy (J) of every member at its current state (Particle.getke, Particle.py:442-454)."""
        p2 = np.sum(self.state[:, 4:7] ** 2, axis=1)
        g = np.sqrt(1 + p2 / (self.mass * c) ** 2)
        return np.where(g - 1 < 1e-6, 0.5 * p2 / self.mass, (g - 1) * self.mass * c * c)


class GuidingCenterEnsemble(_Resident):
    """n guiding centres: GuidingCenter(pos, v, pa, ppar, t0, mass, charge, field) with array arguments
    (reference constructor: rapt/GuidingCenter.py:63-133).  `state` is (n,5): t,X,Y,Z,p_parallel."""

    def __init__(self, pos, v, pa=None, ppar=None, t0=0.0, mass=None, charge=None, field=None):
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        self.n = n
        self.v = _arr(v, n); self.mass = _arr(mass, n); self.charge = _arr(charge, n)
        self.field = field
        t0 = _arr(t0, n)
        if pa is not None:
            pp, mu = engine.gc_construct(field, t0, pos, self.v, _arr(pa, n), self.mass)
        else:
            # ppar given: mu from utils.magnetic_moment (GuidingCenter.py:129) via pitch angle of (ppar, v)
            g = 1 / np.sqrt(1 - (self.v / c) ** 2)
            pp = _arr(ppar, n)
            Bm = engine.field_ops(field, np.column_stack([t0, pos]), which=["magB"])["magB"]
            vpar = pp / (self.mass * g)
            mu = g ** 2 * self.mass * (self.v - vpar) * (self.v + vpar) / (2 * Bm)
        self.mu = mu
        self.state = np.column_stack([t0, pos, pp])
        self.tcur = t0.copy()
        self.counters = np.zeros((n, 4), dtype=np.int64)
        self.status = np.ones(n, dtype=np.int32)
        self.nrows = np.ones(n, dtype=np.int64)
        self.trajectory = None; self.nstored = None
        self.check_adiabaticity = False
        self._dev = None

    _MEMBER_ARRAYS = ("state", "v", "mass", "charge", "mu", "tcur", "counters", "status", "nrows")
    _DIAG = "r"

    def bounceperiod(self, method="quadpack"):
        """GuidingCenter.bounceperiod of every member (rapt/GuidingCenter.py:593-606), all on the device.
        "quadpack" (default): field-line trace, then scipy's spline, brentq and QUADPACK QAGS restated per thread: the
            reference's value to ~1e-9;
        "closed-form": mirror points and integral in closed form (within the error of the reference's QUADPACK call:
            1e-7 typical, 2e-5 worst seen)."""
        if method not in ("quadpack", "closed-form"):
            raise ValueError("method is 'quadpack' or 'closed-form' (the scipy cross-check lives in tests/scipy_legs.py)")
        return engine.bounceperiod_device(self.field, self.state, self.mu, self.mass, params["fieldlineresolution"],
                                          quadrature="closed" if method == "closed-form" else "quadpack")

    def _dt(self):
        if params["GCtimestep"] != 0:                                # GuidingCenter.py:443-446
            return np.full(self.n, float(params["GCtimestep"]))
        if self._dev is not None and self._dev.pending_advances:
            self.pull()     # the reference recomputes the bounce period from trajectory[-1] at every advance (:446)
            self._dev.pending_advances = 0
        return self.bounceperiod() / params["bounceresolution"]

    def cuda(self, device="cuda:0"):
        self._dev = _DeviceState([self.state[:, i] for i in range(5)],
                                 dict(mass=self.mass, charge=self.charge, mu=self.mu, v=self.v), self._bind(device))
        return self

    def advance(self, delta, eom="TaoChanBrizardEOM", store_every=0, max_rows=0, dt=None, **over):
        """GuidingCenter.advance(delta, eom) for every member (rapt/GuidingCenter.py:397-458)."""
        if dt is None and params["GCtimestep"] != 0:                 # GuidingCenter.py:443-444
            dt = float(params["GCtimestep"])
        uniform = float(dt) if np.isscalar(dt) else None
        if self._dev is not None:
            import torch
            d = self._dev
            if "dt" not in d.extras or d.extras["dt"].numel() != self.n:
                d.extras["dt"] = torch.empty(self.n, dtype=torch.float64, device=d.device)
                d.dt_uniform = None
            if uniform is None or getattr(d, "dt_uniform", None) != uniform:     # a repeated uniform step stays resident
                dt = self._dt() if dt is None else _arr(dt, self.n)
                d.extras["dt"].copy_(torch.as_tensor(dt), non_blocking=False)
                d.dt_uniform = uniform
            rows = self._rows_buffer(store_every, max_rows)
            over.setdefault("sort_by_work", 2)     # longest-first from the previous call's step counts (see ParticleEnsemble)
            engine.gc_advance_dev(self.field, d.cols, d.extras["mu"], d.extras["v"], d.extras["mass"], d.extras["charge"],
                                  d.extras["dt"], float(delta), d.out, eom=eom, store_every=store_every, max_rows=max_rows,
                                  rows=rows, check_adiabaticity=self.check_adiabaticity, **over)
            d.cum_counters += d.out["counters"]
            d.pending_advances += 1
            return self
        dt = self._dt() if dt is None else _arr(dt, self.n)
        o = engine.gc_advance(self.field, self.state, self.mu, self.v, self.mass, self.charge, dt, float(delta), eom=eom,
                              store_every=store_every, max_rows=max_rows, check_adiabaticity=self.check_adiabaticity, **over)
        self.state = o["state"]; self.tcur = o["tcur"]; self.status = o["status"]
        self.last_counters = o["counters"].astype(np.int64)
        self.counters += self.last_counters
        self.nrows = o["nrows"].astype(np.int64)
        self.trajectory = o["rows"]; self.nstored = o["nstored"]
        return self

    def member_trajectory(self, i):
        return self.trajectory[i, :self.nstored[i], :5]

    def getke(self):
        """Kinetic energy (J) at the current state (GuidingCenter.getke, GuidingCenter.py:574-591)."""
        Bm = engine.field_ops(self.field, self.state[:, :4], which=["magB"])["magB"]
        mc = self.mass * c
        g = np.sqrt(1 + 2 * self.mu * Bm / (mc * c) + (self.state[:, 4] / mc) ** 2)
        return np.where(g - 1 < 1e-6, self.mu * Bm + 0.5 * self.state[:, 4] ** 2 / self.mass, (g - 1) * mc * c)


class BounceCenterEnsemble:
    """n bounce centres: BounceCenter(pos, v, t0, pa, mass, charge, field) with array arguments
    (rapt/BounceCenter.py:74-115).  `state` is (n,4): t (row label), x, y, z.  As in the reference the magnetic moment
    uses cos(pa) of the pitch angle as given (BounceCenter.py:114)."""

    def __init__(self, pos, v, t0=0.0, pa=None, mass=None, charge=None, field=None):
        if not field.static:
            raise RuntimeError("BounceCenter does not work with nonstatic fields or electric fields.")
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        self.n = n
        self.v = _arr(v, n); self.mass = _arr(mass, n); self.charge = _arr(charge, n); self.pa = _arr(pa, n)
        self.field = field
        t0 = _arr(t0, n)
        g = 1 / np.sqrt(1 - (self.v / c) ** 2)
        Bmag = engine.field_ops(field, np.column_stack([t0, pos]), which=["magB"])["magB"]
        vpar = self.v * np.cos(self.pa)
        self.mu = g ** 2 * self.mass * (self.v - vpar) * (self.v + vpar) / (2 * Bmag)       # utils.py:214-216
        self.state = np.column_stack([t0, pos])
        self.tcur = t0.copy()
        self.counters = np.zeros((n, 4), dtype=np.int64)
        self.status = np.ones(n, dtype=np.int32)
        self.trajectory = None; self.nstored = None; self.nrows = None; self.dt = None

    def advance(self, delta, store_every=0, max_rows=0, dt=None, **over):
        """BounceCenter.advance(delta) for every member (rapt/BounceCenter.py:206-251)."""
        o = engine.bounce_center_advance(self.field, self.state, self.mu, self.v, self.mass, self.charge, float(delta), dt=dt,
                                         store_every=store_every, max_rows=max_rows, **over)
        self.state = o["state"]; self.tcur = o["state"][:, 0].copy(); self.status = o["status"]
        self.last_counters = o["counters"].astype(np.int64)
        self.counters += self.last_counters
        self.nrows = o["nrows"].astype(np.int64); self.dt = o["dt"]
        self.trajectory = o["rows"]; self.nstored = o["nstored"]
        return self

    def member_trajectory(self, i):
        return self.trajectory[i, :self.nstored[i], :4]


class AdaptiveEnsemble:
    """n Adaptive tracers: Adaptive(pos, vel, t0, mass, charge, field) with array arguments
    (rapt/Adaptive.py:70-104).  Mode switches run on the device: after every epoch a switch kernel
    applies the Particle<->GuidingCenter transforms and regroups tracers by mode with warp ballots and
    a shared-memory scan (rapt_b200/csrc/rapt_aux.cuh).  Needs params['GCtimestep'] != 0."""

    def __init__(self, pos, vel, t0=0.0, mass=None, charge=None, field=None):
        self.pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        self.vel = np.asarray(vel, dtype=np.float64).reshape(-1, 3)
        self.n = len(self.pos)
        self.t0 = _arr(t0, self.n); self.mass = _arr(mass, self.n); self.charge = _arr(charge, self.n)
        self.field = field
        self.result = None

    def advance(self, delta, store_every=1, max_rows=0, **over):
        """Adaptive.advance(delta) for every member (rapt/Adaptive.py:187-222), from the constructor state."""
        gc_dt = float(over.pop("GCtimestep", params["GCtimestep"]))
        if gc_dt == 0:
            raise ValueError("AdaptiveEnsemble needs params['GCtimestep'] != 0 (the bounce-period output step is a "
                             "per-tracer host quadrature; use Adaptive objects for that)")
        self.result = engine.adaptive_advance(self.field, self.pos, self.vel, self.t0, self.mass, self.charge,
                                              float(delta), gc_dt, store_every=store_every, max_rows=max_rows, **over)
        return self

    def segments(self, i):
        """List of (mode, rows) for member i; mode 0 = Particle rows (k,7), 1 = GuidingCenter rows (k,5)."""
        r = self.result
        rows = r["rows"][i, :r["nstored"][i]]
        tags = rows[:, 7].astype(np.int64)
        out = []
        for tag in np.unique(tags):
            seg = rows[tags == tag]
            mode = int(tag & 1)
            out.append((mode, seg[:, :7] if mode == 0 else seg[:, :5]))
        return out
