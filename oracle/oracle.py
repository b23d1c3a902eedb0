"""ctypes front-end of the CPU oracle (oracle/rapt_oracle.c) + the host-side pieces of the path
that the reference delegates to scipy (quadratic spline, brentq, QUADPACK in
rapt/flutils.py:308-314).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package rapt_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

c_light = 299792458
Re = 6378137
B0 = 3.07e-5

KIND = {"EarthDipole": 0, "DoubleDipole": 1, "UniformBz": 2, "UniformCrossedEB": 3, "VarEarthDipole": 4,
        "Parabolic": 5, "ChargedDipole": 100}
EOM = {"TaoChanBrizardEOM": 0, "BrizardChanEOM": 1, "NorthropTellerEOM": 2}


class OGrid(C.Structure):
    _fields_ = [("nt", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("t", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("B", C.c_void_p * 3), ("E", C.c_void_p * 3)]


class OField(C.Structure):
    _fields_ = [("kind", C.c_int), ("is_static", C.c_int), ("prm", C.c_double * 8),
                ("gradstep", C.c_double), ("tstep", C.c_double), ("grid", C.c_void_p)]


class OParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("cyclotronresolution", C.c_double),
                ("gctimestep", C.c_double), ("epss", C.c_double), ("epst", C.c_double),
                ("enforce_equatorial", C.c_int), ("dop853_reject_rule", C.c_int)]


def build(force=False):
    src = os.path.join(HERE, "rapt_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_fieldline_trace.restype = C.c_long
        _lib.oracle_adaptive_one.restype = C.c_long
    return _lib


def make_field(name, *args, gradstep=None, tstep=1e-3, static=None):
    """Field descriptor with the reference constructors' defaults (fields.py:294,336,373,408,449,501)."""
    f = OField()
    f.kind = KIND[name]
    prm = [0.0] * 8
    gs, st = 1e-6, True
    if name == "EarthDipole":
        b0 = args[0] if args else B0
        prm[0] = -3 * b0 * Re ** 3; gs = Re * 1e-6
    elif name == "DoubleDipole":
        b0 = args[0] if len(args) > 0 else B0
        dd = args[1] if len(args) > 1 else 20 * Re
        k = args[2] if len(args) > 2 else 1
        prm[0] = -b0 * Re ** 3; prm[1] = dd; prm[2] = k; gs = Re / 1000
    elif name == "UniformBz":
        prm[0] = args[0] if args else 1
    elif name == "UniformCrossedEB":
        ey = args[0] if len(args) > 0 else 1
        bz = args[1] if len(args) > 1 else 1
        prm[0] = bz; prm[1] = ey; st = False
    elif name == "VarEarthDipole":
        prm[0] = args[0] if len(args) > 0 else 0.1
        prm[1] = args[1] if len(args) > 1 else 10
        gs = Re / 1000; st = False
    elif name == "Parabolic":
        prm[0] = args[0] if len(args) > 0 else 10.0
        prm[1] = args[1] if len(args) > 1 else 1.0
        prm[2] = args[2] if len(args) > 2 else 0.2
    elif name == "ChargedDipole":
        prm[0] = args[0] if len(args) > 0 else 1
        prm[1] = args[1] if len(args) > 1 else 1
        prm[2] = 8.9875517873681764e9; st = False
    for i, v in enumerate(prm):
        f.prm[i] = float(v)
    f.gradstep = gs if gradstep is None else gradstep
    f.tstep = tstep
    f.is_static = int(st if static is None else static)
    return f


def make_grid_field(t, x, y, z, B, E, static=True, gradstep=1e-3 * Re, tstep=1e-3):
    """fields.Grid (fields.py:513-814) from parsed data: t (nt,), x, y, z node coordinates, B and E as three
    arrays each of shape (nt, nx, ny, nz) (or (nx, ny, nz) for a single time point).  Defaults follow the
    reference constructor (gradientstepsize 1e-3 Re, static left True, fields.py:577-579)."""
    f = OField()
    f.kind = 6; f.is_static = int(static); f.gradstep = gradstep; f.tstep = tstep
    g = OGrid()
    keep = []
    def arr(a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)); keep.append(a); return a.ctypes.data
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    g.nt, g.nx, g.ny, g.nz = len(t), len(x), len(y), len(z)
    g.t, g.x, g.y, g.z = arr(t), arr(x), arr(y), arr(z)
    for i in range(3):
        assert np.asarray(B[i]).size == g.nt * g.nx * g.ny * g.nz
        g.B[i] = arr(B[i]); g.E[i] = arr(E[i])
    keep.append(g)
    f.grid = C.addressof(g)
    f._keep = keep                      # the C side holds raw pointers into these
    return f


def make_params(**over):
    """Snapshot of rapt.params (rapt/__init__.py:21-34) with overrides by the reference's key names."""
    p = OParams()
    tol = over.get("solvertolerances", (1.49012e-8, 1.49012e-8))
    p.rtol, p.atol = float(tol[0]), float(tol[1])
    p.cyclotronresolution = float(over.get("cyclotronresolution", 10))
    p.gctimestep = float(over.get("GCtimestep", 0))
    p.epss = float(over.get("epss", 5e-2)); p.epst = float(over.get("epst", 5e-2))
    p.enforce_equatorial = int(bool(over.get("enforce equatorial", False)))
    p.dop853_reject_rule = int(over.get("dop853_reject_rule", 0))
    return p


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def field_ops(f, tpos):
    tpos = _d(tpos).reshape(-1, 4); n = len(tpos)
    o = dict(B=np.zeros((n, 3)), E=np.zeros((n, 3)), unitb=np.zeros((n, 3)), magB=np.zeros(n),
             gradB=np.zeros((n, 3)), jacobianB=np.zeros((n, 3, 3)), curlb=np.zeros((n, 3)), curvature=np.zeros(n),
             dBdt=np.zeros(n), dbdt=np.zeros((n, 3)), lengthscale=np.zeros(n), timescale=np.zeros(n))
    lib().oracle_field_ops(C.byref(f), C.c_long(n), _p(tpos), *[_p(o[k]) for k in
                           ("B", "E", "unitb", "magB", "gradB", "jacobianB", "curlb", "curvature", "dBdt", "dbdt",
                            "lengthscale", "timescale")])
    return o


def utils_ops(f, pos, vel, mass, charge):
    pos = _d(pos).reshape(-1, 3); vel = _d(vel).reshape(-1, 3); n = len(pos)
    o = dict(cycper=np.zeros(n), cycrad=np.zeros(n), gc_R=np.zeros((n, 3)), gc_vp=np.zeros(n), gc_v=np.zeros(n),
             mu=np.zeros(n), fp_pos=np.zeros((n, 3)), fp_vel=np.zeros((n, 3)), cycper2=np.zeros(n), cycrad2=np.zeros(n))
    lib().oracle_utils(C.byref(f), C.c_long(n), _p(pos), _p(vel), C.c_double(mass), C.c_double(charge),
                       *[_p(o[k]) for k in ("cycper", "cycrad", "gc_R", "gc_vp", "gc_v", "mu", "fp_pos", "fp_vel",
                                            "cycper2", "cycrad2")])
    return o


def getperp(v):
    v = _d(v); o = np.zeros(3)
    lib().oracle_getperp(_p(v), _p(o))
    return o


def test_solver(solver, which_ode, x0, xend, y0, rtol, atol, reject_rule=0):
    y = _d(y0).copy(); cnt = np.zeros(4, dtype=np.int64)
    idid = lib().oracle_test_solver(C.c_int(solver), C.c_int(which_ode), C.c_double(x0), C.c_double(xend), _p(y),
                                    C.c_double(rtol), C.c_double(atol), C.c_int(reject_rule), _p(cnt))
    return idid, y, cnt


def particle_momentum(vel, mass):
    """Particle.__init__ (Particle.py:106-107)."""
    vel = _d(vel)
    gamma = 1 / np.sqrt(1 - np.sum(vel * vel, axis=-1) / c_light ** 2)
    return (mass * gamma)[..., None] * vel if vel.ndim > 1 else mass * gamma * vel


def particle_advance(f, par, state, mass, charge, delta, check_adiab=False, store_every=1, max_rows=0,
                     want_percall=False, nthreads=1):
    """state: (n,7) rows (t,x,y,z,px,py,pz).  Returns dict (state updated copy)."""
    st = _d(state).reshape(-1, 7).copy(); n = len(st)
    cols = [np.ascontiguousarray(st[:, i]) for i in range(7)]
    mass = np.broadcast_to(_d(mass), (n,)).copy(); charge = np.broadcast_to(_d(charge), (n,)).copy()
    rows = np.zeros((n, max_rows, 8)) if (store_every > 0 and max_rows > 0) else None
    nrows = np.zeros(n, np.int64); nstored = np.zeros(n, np.int64); counters = np.zeros((n, 4), np.int64)
    status = np.zeros(n, np.int32); tcur = np.zeros(n); dt = np.zeros(n)
    max_calls = max_rows if want_percall else 0
    percall = np.zeros((max(max_calls, 1), 4), np.int64)
    lib().oracle_particle_advance(C.byref(f), C.byref(par), C.c_long(n), *[_p(c) for c in cols], _p(mass), _p(charge),
                                  C.c_double(delta), C.c_int(int(check_adiab)), C.c_long(store_every), C.c_long(max_rows),
                                  _p(rows), _p(nrows), _p(nstored), _p(counters), _p(status), _p(tcur), _p(dt),
                                  _p(percall) if want_percall else None, C.c_long(max_calls), C.c_int(nthreads))
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, dt=dt, percall=percall[:max(int(nrows[0]) - 1, 0)] if want_percall else None)


def errgap(fn, n, *a, **k):
    """Run oracle.particle_advance / gc_advance (fn) and also return, per tracer, the smallest |err - 1| over all its
    step attempts: how close its nearest accept/reject decision was (parity report, tests/test_gpu_properties.py)."""
    gap = np.full(n, 1e300)
    lib().oracle_set_errgap_out(_p(gap))
    try:
        o = fn(*a, **k)
    finally:
        lib().oracle_set_errgap_out(None)
    o["errgap"] = gap
    return o


def gc_construct(f, t0, pos, v, pa, mass):
    pos = _d(pos).reshape(-1, 3); n = len(pos)
    t0 = np.broadcast_to(_d(t0), (n,)).copy(); v = np.broadcast_to(_d(v), (n,)).copy()
    pa = np.broadcast_to(_d(pa), (n,)).copy(); mass = np.broadcast_to(_d(mass), (n,)).copy()
    ppar = np.zeros(n); mu = np.zeros(n)
    x, y, z = (np.ascontiguousarray(pos[:, i]) for i in range(3))
    lib().oracle_gc_construct(C.byref(f), C.c_long(n), _p(t0), _p(x), _p(y), _p(z), _p(v), _p(pa), _p(mass), _p(ppar), _p(mu))
    return ppar, mu


def gc_advance(f, par, state, mu, v, mass, charge, dt, delta, eom="TaoChanBrizardEOM", check_adiab=False,
               store_every=1, max_rows=0, want_percall=False, nthreads=1):
    st = _d(state).reshape(-1, 5).copy(); n = len(st)
    cols = [np.ascontiguousarray(st[:, i]) for i in range(5)]
    bc = lambda a: np.broadcast_to(_d(a), (n,)).copy()
    mu, v, mass, charge, dt = bc(mu), bc(v), bc(mass), bc(charge), bc(dt)
    rows = np.zeros((n, max_rows, 8)) if (store_every > 0 and max_rows > 0) else None
    nrows = np.zeros(n, np.int64); nstored = np.zeros(n, np.int64); counters = np.zeros((n, 4), np.int64)
    status = np.zeros(n, np.int32); tcur = np.zeros(n)
    max_calls = max_rows if want_percall else 0
    percall = np.zeros((max(max_calls, 1), 4), np.int64)
    lib().oracle_gc_advance(C.byref(f), C.byref(par), C.c_int(EOM[eom]), C.c_long(n), *[_p(c) for c in cols],
                            _p(mu), _p(v), _p(mass), _p(charge), _p(dt), C.c_double(delta), C.c_int(int(check_adiab)),
                            C.c_long(store_every), C.c_long(max_rows), _p(rows), _p(nrows), _p(nstored), _p(counters),
                            _p(status), _p(tcur), _p(percall) if want_percall else None, C.c_long(max_calls),
                            C.c_int(nthreads))
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, percall=percall[:max(int(nrows[0]) - 1, 0)] if want_percall else None)


# ---------------------------------------------------------------- bounce period (a13)
def fieldline_trace(f, tpos, Bm, fieldlineresolution=50):
    """Fieldline(tpos, field, Bmax=Bm).trace() (fieldline.py:13-105): returns curve (n,4), B (n,), ds."""
    tpos = _d(tpos); cap = 4096
    while True:
        curve = np.zeros((cap, 4)); B = np.zeros(cap); ds = C.c_double(0)
        n = lib().oracle_fieldline_trace(C.byref(f), _p(tpos), C.c_double(Bm), C.c_double(fieldlineresolution),
                                         _p(curve), _p(B), C.c_long(cap), C.byref(ds))
        if n <= cap:
            return curve[:n], B[:n], ds.value
        cap *= 4


def halfbouncepath_from_curve(s, b, Bm):
    """flutils.py:274-316 on a traced curve; the non-equatorial branch is scipy, as in the reference."""
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from scipy.integrate import quad
    n = len(b)
    inside = np.where(b <= Bm)[0]
    if len(inside) == 0:
        i1 = int((n - 3) / 2); i2 = int((n + 1) / 2)
    else:
        i1, i2 = inside[0] - 1, inside[-1] + 1
    b = np.delete(b, list(range(0, i1)) + list(range(i2 + 1, n)))
    s = np.delete(s, list(range(0, i1)) + list(range(i2 + 1, n)))
    n = len(b)
    if n == 3:
        s1, s2, s3 = s[0], s[1], s[2]
        B1, B2, B3 = b[0], b[1], b[2]
        s12 = s1 - s2; s23 = s2 - s3; s13 = s1 - s3
        B2s = 2 * (B1 * s23 - B2 * s13 + B3 * s12) / (s12 * s13 * s23)
        return np.pi * np.sqrt(2 * Bm / B2s)
    Bf = interp1d(s, b, kind='quadratic', assume_sorted=True)
    sm1 = brentq(lambda x: Bf(x) - Bm, s[0], s[1])
    sm2 = brentq(lambda x: Bf(x) - Bm, s[-2], s[-1])
    return quad(lambda x: 1 / np.sqrt(1 - Bf(x) / Bm), sm1, sm2, epsrel=1e-4)[0]


def gc_mirror(f, state, mu, mass):
    state = _d(state); Bm = C.c_double(0); v = C.c_double(0)
    lib().oracle_gc_mirror(C.byref(f), _p(state), C.c_double(mu), C.c_double(mass), C.byref(Bm), C.byref(v))
    return Bm.value, v.value


def bounceperiod(f, state, mu, mass, fieldlineresolution=50):
    """GuidingCenter.bounceperiod (GuidingCenter.py:593-606) -> flutils.bounceperiod (flutils.py:252)."""
    Bm, v = gc_mirror(f, state, mu, mass)
    curve, B, ds = fieldline_trace(f, state[:4], Bm, fieldlineresolution)
    return (2 / v) * halfbouncepath_from_curve(curve[:, 0], B, Bm)


# ---------------------------------------------------------------- Adaptive (a15-a17)
def switch_P2G(f, prow, mass, charge):
    prow = _d(prow); grow = np.zeros(5); mu = C.c_double(0); v = C.c_double(0)
    rc = lib().oracle_switch_P2G(C.byref(f), _p(prow), C.c_double(mass), C.c_double(charge), _p(grow), C.byref(mu), C.byref(v))
    return rc, grow, mu.value, v.value


def switch_G2P(f, grow, mu, mass, charge, t_eval=0.0):
    grow = _d(grow); prow = np.zeros(7)
    lib().oracle_switch_G2P(C.byref(f), _p(grow), C.c_double(mu), C.c_double(mass), C.c_double(charge), C.c_double(t_eval), _p(prow))
    return prow


def particle_isadiabatic(f, par, row, mass, charge):
    return bool(lib().oracle_particle_isadiabatic(C.byref(f), C.byref(par), _p(_d(row)), C.c_double(mass), C.c_double(charge)))


def gc_isadiabatic(f, par, row, mu, mass, charge):
    return bool(lib().oracle_gc_isadiabatic(C.byref(f), C.byref(par), _p(_d(row)), C.c_double(mu), C.c_double(mass), C.c_double(charge)))


def adaptive(f, par, pos, vel, t0, mass, charge, delta, bounceresolution=10, fieldlineresolution=50, max_rows=1 << 16):
    """Adaptive.__init__ + Adaptive.advance (Adaptive.py:96-104, 202-222) driven from Python so the
    bounce-period dt (GCtimestep == 0) can use the scipy quadrature.  Returns list of segments
    (mode, rows ndarray (k,8), mu) and counters."""
    vel = _d(vel)
    prow = np.concatenate(([t0], _d(pos), particle_momentum(vel, mass)))
    segs = []; counters = np.zeros(4, np.int64)
    if particle_isadiabatic(f, par, prow, mass, charge):
        rc, grow, mu, v = switch_P2G(f, prow, mass, charge)
        if rc:
            raise RuntimeError("guiding-centre iteration failed")
        mode, cur = 1, grow
    else:
        mode, cur, mu, v = 0, prow, 0.0, 0.0
    r0 = np.zeros(8); r0[:len(cur)] = cur
    if mode == 1:
        r0[5] = mu
    segs.append([mode, [r0], mu])
    t, tcur = 0.0, t0
    while t < delta:
        rem = delta - t
        if mode == 0:
            o = particle_advance(f, par, cur, mass, charge, rem, check_adiab=True, store_every=1, max_rows=max_rows)
            k = int(o["nstored"][0]); segs[-1][1].extend(o["rows"][0, 1:k]); cur = o["state"][0]; tcur = o["tcur"][0]
            counters += o["counters"][0]
            if o["status"][0] == 2:
                rc, grow, mu, v = switch_P2G(f, cur, mass, charge)
                if rc:
                    raise RuntimeError("guiding-centre iteration failed")
                mode, cur, tcur = 1, grow, grow[0]
                r0 = np.zeros(8); r0[:5] = grow; r0[5] = mu
                segs.append([1, [r0], mu])
            elif o["status"][0] < 0:
                break
        else:
            dt = par.gctimestep if par.gctimestep != 0 else bounceperiod(f, cur, mu, mass, fieldlineresolution) / bounceresolution
            o = gc_advance(f, par, cur, mu, v, mass, charge, dt, rem, check_adiab=True, store_every=1, max_rows=max_rows)
            k = int(o["nstored"][0]); segs[-1][1].extend(o["rows"][0, 1:k]); cur = o["state"][0]; tcur = o["tcur"][0]
            counters += o["counters"][0]
            if o["status"][0] == 3:
                prow = switch_G2P(f, cur, mu, mass, charge, 0.0)
                mode, cur, tcur = 0, prow, prow[0]
                r0 = np.zeros(8); r0[:7] = prow
                segs.append([0, [r0], 0.0])
            elif o["status"][0] < 0:
                break
        t = tcur
    return [(m, np.array(r), mu_) for m, r, mu_ in segs], counters


def adaptive_c(f, par, pos, vel, t0, mass, charge, delta, max_rows=1 << 16, max_segs=64):
    """Same, entirely in C (needs GCtimestep != 0); used for the CPU baseline of config 4."""
    rows = np.zeros((max_rows, 8)); seglog = np.zeros((max_segs, 3)); nrows = C.c_long(0); cnt = np.zeros(4, np.int64)
    nseg = lib().oracle_adaptive_one(C.byref(f), C.byref(par), _p(_d(pos)), _p(_d(vel)), C.c_double(t0), C.c_double(mass),
                                     C.c_double(charge), C.c_double(delta), _p(rows), C.c_long(max_rows), C.byref(nrows),
                                     _p(seglog), C.c_long(max_segs), _p(cnt))
    return nseg, rows[:nrows.value], seglog[:max(nseg, 0)], cnt


# ---------------------------------------------------------------- BounceCenter
def eye_from_curve(s, b, Bm):
    """flutils.eye (flutils.py:95-151) on a traced curve.  scipy's interp1d / brentq / quad as the reference calls
    them; the eqpa < 70 branch calls `simps`, which the reference never imports (NameError) -- evaluated with
    scipy.integrate.simpson(y, x=x), the function its import line (flutils.py:18) provides."""
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from scipy.integrate import quad, simpson
    s = np.array(s, dtype=float); b = np.array(b, dtype=float)
    n = len(b)
    Bmin = np.min(b)
    if Bmin > Bm:
        return 0
    if abs(Bmin - Bm) / Bm < 1e-12:
        return 0
    inside = np.where(b < Bm)[0]
    drop = list(range(0, inside[0] - 1)) + list(range(inside[-1] + 2, n))
    b = np.delete(b, drop); s = np.delete(s, drop)
    assert b[0] > Bm and b[1] < Bm and b[-2] < Bm and b[-1] > Bm
    eqpa = np.arcsin(np.sqrt(Bmin / Bm)) * 180 / np.pi
    if eqpa < 70:
        sm1 = (Bm - b[0]) * (s[1] - s[0]) / (b[1] - b[0]) + s[0]
        sm2 = (Bm - b[-2]) * (s[-1] - s[-2]) / (b[-1] - b[-2]) + s[-2]
        s[0], s[-1] = sm1, sm2
        b[0], b[-1] = Bm, Bm
        bi = np.sqrt(1 - b[1:-1] / Bm)
        I = simpson(bi, x=s[1:-1])
        d = s[-1] - s[-2]
        I += (2 / 3) * d * np.sqrt((Bm - b[-2]) / Bm)
        d = s[1] - s[0]
        I += (2 / 3) * d * np.sqrt((Bm - b[1]) / Bm)
        return I
    B = interp1d(s, b, kind='quadratic', assume_sorted=True)
    sm1 = brentq(lambda x: B(x) - Bm, s[0], s[1])
    if B(s[-2]) == Bm:
        sm2 = s[-2]
    else:
        sm2 = brentq(lambda x: B(x) - Bm, s[-2], s[-1])
    return quad(lambda x: np.sqrt(1 - B(x) / Bm), sm1, sm2, epsrel=1e-4)[0]


def eye(f, tpos, Bm, fieldlineresolution=50):
    curve, B, _ = fieldline_trace(f, tpos, Bm, fieldlineresolution)
    return eye_from_curve(curve[:, 0], B, Bm)


def halfbouncepath(f, tpos, Bm, fieldlineresolution=50):
    curve, B, _ = fieldline_trace(f, tpos, Bm, fieldlineresolution)
    return halfbouncepath_from_curve(curve[:, 0], B, Bm)


def gradI(f, tpos, Bm, d=0.03 * Re, fieldlineresolution=50):
    """flutils.gradI (flutils.py:183-229)."""
    tpos = np.array(tpos, dtype=float)
    x, y = tpos[1], tpos[2]
    r = np.array((x, y, 0)) / max((abs(x), abs(y)))
    b = field_ops(f, tpos)["unitb"][0]
    r = r - np.dot(r, b) * b
    r = r / np.sqrt(np.dot(r, r))
    v1 = np.zeros(4); v1[1:4] = r
    v2 = np.zeros(4); v2[1:4] = np.cross(b, r)
    I = lambda p: eye(f, p, Bm, fieldlineresolution)
    out = []
    for v in (v1, v2):
        I1 = I(tpos + d * v); I2 = I(tpos - d * v)
        if I1 == 0:
            out.append((I(tpos) - I2) / d)
        elif I2 == 0:
            out.append((I1 - I(tpos)) / d)
        else:
            out.append((I1 - I2) / (2 * d))
    return out[0] * v1[1:] + out[1] * v2[1:]


def bc_mirror_field(mu, v, mass):
    """BounceCenter.py:226-227"""
    gamma = 1.0 / np.sqrt(1 - (v / 299792458) ** 2)
    return mass * gamma ** 2 * v ** 2 / (2 * mu), gamma


def bc_deriv(f, t, Y, Bm, gamma, v, mass, charge, d=0.03 * Re, fieldlineresolution=50):
    """BounceCenter.advance.deriv (BounceCenter.py:232-243)."""
    tpos = np.concatenate(([t], Y))
    Bvec = field_ops(f, tpos)["B"][0]
    magBsq = np.dot(Bvec, Bvec)
    Sb = halfbouncepath(f, tpos, Bm, fieldlineresolution)
    gI = gradI(f, tpos, Bm, d, fieldlineresolution)
    return gamma * mass * v * v / (charge * Sb * magBsq) * np.cross(gI, Bvec)


def bounce_center_advance(f, row, mu, v, mass, charge, delta, bctimestep=0.1, tol=(1.49012e-8, 1.49012e-8),
                          d=0.03 * Re, fieldlineresolution=50):
    """BounceCenter.advance (BounceCenter.py:206-251) for one tracer from its last row (t, x, y, z): the reference's
    control flow, the C dopri5 of this oracle on the Python right-hand side above.  Returns (rows (k,4), counters, dt)."""
    row = np.array(row, dtype=float)
    Bm, gamma = bc_mirror_field(mu, v, mass)
    dt = bctimestep * (2 / v) * halfbouncepath(f, row[:4], Bm, fieldlineresolution)
    CB = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))

    def rhs(t, yp, dyp):
        Y = np.array([yp[0], yp[1], yp[2]])
        out = bc_deriv(f, t, Y, Bm, gamma, v, mass, charge, d, fieldlineresolution)
        dyp[0], dyp[1], dyp[2] = out[0], out[1], out[2]
    cb = CB(rhs)
    x = C.c_double(row[0]); y = row[1:4].copy(); cnt = np.zeros(4, dtype=np.int64)
    rows = []
    tcur = row[0]
    for t in np.arange(tcur, tcur + delta, dt):
        idid = lib().oracle_dopri5_callback(C.c_int(3), cb, C.byref(x), _p(y), C.c_double(x.value + dt),
                                            C.c_double(tol[0]), C.c_double(tol[1]), _p(cnt))
        if idid != 1:
            break
        rows.append(np.concatenate(([t], y)))
    return np.array(rows).reshape(-1, 4), cnt, dt
