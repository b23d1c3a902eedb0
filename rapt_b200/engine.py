"""Array-level front end of librapt_b200.so: numpy in, numpy out (host-pointer C ABI), or torch CUDA
tensors in place (device-pointer C ABI).  The reference-shaped classes (Particle, GuidingCenter,
Adaptive) and the ensemble classes are thin layers over these functions.
"""
import ctypes as C
import math
import numpy as np

from . import _lib
from ._lib import ParamsT, FieldT, EOM_KIND, check, ptr

_user_cache = {}


def snapshot_params(params=None, check_adiabaticity=False, **over):
    """By-value snapshot of rapt_b200.params (rapt/__init__.py:21-34) -> rapt_params_t."""
    from . import params as global_params
    src = dict(global_params if params is None else params)
    src.update(over)
    p = ParamsT()
    rtol, atol = src["solvertolerances"]
    p.rtol, p.atol = float(rtol), float(atol)
    p.cyclotronresolution = float(src["cyclotronresolution"])
    p.epss, p.epst = float(src["epss"]), float(src["epst"])
    p.enforce_equatorial = int(bool(src["enforce equatorial"]))
    p.check_adiabaticity = int(bool(check_adiabaticity))
    p.dop853_reject_rule = int(src.get("dop853_reject_rule", 0))
    arith = src.get("arith", "fast")
    p.arith = 1 if arith in ("strict", 1) else 0
    p.sort_by_work = int(src.get("sort_by_work", 1))
    return p


def _field_desc(field):
    return field if isinstance(field, FieldT) else field.device_descriptor()


def create_grid(t, x, y, z, B, E):
    """Upload the parsed data of a fields.Grid (rapt_b200_grid_create); returns the grid handle.
    t (nt,), x, y, z node coordinates, B and E: three arrays each of shape (nt, nx, ny, nz)."""
    t = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    shape = (len(t), len(x), len(y), len(z))
    comps = [np.ascontiguousarray(a, dtype=np.float64).reshape(shape) for a in list(B) + list(E)]
    gid = C.c_int32(-1)
    check(_lib.load().rapt_b200_grid_create(C.c_int64(shape[0]), C.c_int64(shape[1]), C.c_int64(shape[2]), C.c_int64(shape[3]),
                                            ptr(t), ptr(x), ptr(y), ptr(z), *[ptr(a) for a in comps], C.byref(gid)))
    return gid.value


def destroy_grid(grid_id):
    check(_lib.load().rapt_b200_grid_destroy(C.c_int32(grid_id)))


def compile_user_field(source, has_E):
    """NVRTC-compile a user field snippet (cached by source); returns the user_id handle."""
    key = (source, bool(has_E))
    if key not in _user_cache:
        lib = _lib.load()
        uid = C.c_int(0)
        log = C.create_string_buffer(1 << 16)
        rc = lib.rapt_b200_field_nvrtc(source.encode(), C.c_int(int(has_E)), C.byref(uid), log, C.c_int(len(log)))
        if rc != 0:
            raise _lib.RaptB200Error(f"NVRTC compilation of the field snippet failed ({rc}): "
                                     f"{lib.rapt_b200_last_error().decode()}\n{log.value.decode(errors='replace')}")
        _user_cache[key] = uid.value
    return _user_cache[key]


def _col(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))).copy()


def fp64_peak(iters=1 << 15):
    """Measured FP64 FMA peak of the current device in TFLOP/s (register-resident DFMA chains)."""
    tf = C.c_double(0); mhz = C.c_double(0)
    check(_lib.load().rapt_b200_fp64_peak(C.c_int(iters), C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


# ------------------------------------------------------------------------------------------------
def field_ops(field, tpos, arith="strict", which=None):
    """Batched _Field operators on the device (fields.py:43-280).  tpos: (n,4)."""
    f = _field_desc(field)
    tpos = np.ascontiguousarray(np.asarray(tpos, dtype=np.float64).reshape(-1, 4))
    n = len(tpos)
    shapes = dict(B=(n, 3), E=(n, 3), unitb=(n, 3), magB=(n,), gradB=(n, 3), jacobianB=(n, 3, 3), curlb=(n, 3),
                  curvature=(n,), dBdt=(n,), dbdt=(n, 3), lengthscale=(n,), timescale=(n,))
    names = list(shapes)
    which = names if which is None else list(which)
    out = {k: (np.zeros(shapes[k]) if k in which else None) for k in names}
    check(_lib.load().rapt_b200_field_ops(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int64(n),
                                          ptr(tpos), *[ptr(out[k]) for k in names]))
    return {k: v for k, v in out.items() if v is not None}


def particle_momentum(vel, mass):
    """Particle.__init__ (Particle.py:106-107): p = m gamma v."""
    from . import c
    vel = np.asarray(vel, dtype=np.float64)
    gamma = 1 / np.sqrt(1 - np.sum(vel * vel, axis=-1) / c ** 2)
    return (np.asarray(mass) * gamma)[..., None] * vel


def particle_advance(field, state, mass, charge, delta, store_every=1, max_rows=0, params=None,
                     check_adiabaticity=False, **over):
    """Particle.advance for n particles.  state: (n,7) rows (t,x,y,z,px,py,pz) [host numpy].
    Returns dict(state, rows, nrows, nstored, counters, status, tcur, dt)."""
    f = _field_desc(field)
    p = snapshot_params(params, check_adiabaticity, **over)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 7)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(7)]
    mass = _col(mass, n); charge = _col(charge, n)
    want = store_every > 0 and max_rows > 0
    rows = np.empty((n, max_rows, 8)) if want else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); tcur = np.zeros(n); dt = np.zeros(n)
    check(_lib.load().rapt_b200_particle_advance(
        C.byref(f), C.byref(p), C.c_int64(n), *[ptr(c_) for c_ in cols], ptr(mass), ptr(charge), C.c_double(delta),
        C.c_int64(store_every), C.c_int64(max_rows), ptr(rows), ptr(nrows), ptr(nstored), ptr(counters), ptr(status),
        ptr(tcur), ptr(dt)))
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, dt=dt)


def particle_dt(field, state, mass, charge, params=None, **over):
    """Output step Particle.advance will choose: cyclotron period / cyclotronresolution
    (Particle.py:282, utils.py:63-66), evaluated with the device field."""
    from . import c, params as gp
    st = np.asarray(state, dtype=np.float64).reshape(-1, 7)
    res = dict(gp if params is None else params)
    res.update(over)
    mass = _col(mass, len(st)); charge = _col(charge, len(st))
    mom = st[:, 4:7]
    gm = np.sqrt(mass ** 2 + np.sum(mom * mom, axis=1) / c ** 2)
    vel = mom / gm[:, None]
    gamma = 1.0 / np.sqrt(1 - np.sum(vel * vel, axis=1) / c ** 2)
    Bm = field_ops(field, st[:, :4], which=["magB"])["magB"]
    return 2 * np.pi * gamma * mass / Bm / np.abs(charge) / float(res["cyclotronresolution"])


def gc_construct(field, t0, pos, v, pa, mass, arith="strict"):
    """GuidingCenter.__init__ (GuidingCenter.py:123-133): (ppar, mu) from speed and pitch angle (deg)."""
    f = _field_desc(field)
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    t0, v, pa, mass = _col(t0, n), _col(v, n), _col(pa, n), _col(mass, n)
    x, y, z = (np.ascontiguousarray(pos[:, i]).copy() for i in range(3))
    ppar = np.zeros(n); mu = np.zeros(n)
    check(_lib.load().rapt_b200_gc_construct(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int64(n),
                                             ptr(t0), ptr(x), ptr(y), ptr(z), ptr(v), ptr(pa), ptr(mass), ptr(ppar), ptr(mu)))
    return ppar, mu


def gc_advance(field, state, mu, v, mass, charge, dt, delta, eom="TaoChanBrizardEOM", store_every=1, max_rows=0,
               params=None, check_adiabaticity=False, **over):
    """GuidingCenter.advance for n guiding centres.  state: (n,5) rows (t,X,Y,Z,ppar)."""
    f = _field_desc(field)
    p = snapshot_params(params, check_adiabaticity, **over)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 5)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(5)]
    mu, v, mass, charge, dt = _col(mu, n), _col(v, n), _col(mass, n), _col(charge, n), _col(dt, n)
    want = store_every > 0 and max_rows > 0
    rows = np.empty((n, max_rows, 8)) if want else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); tcur = np.zeros(n)
    check(_lib.load().rapt_b200_gc_advance(
        C.byref(f), C.byref(p), C.c_int(EOM_KIND[eom]), C.c_int64(n), *[ptr(c_) for c_ in cols],
        ptr(mu), ptr(v), ptr(mass), ptr(charge), ptr(dt), C.c_double(delta),
        C.c_int64(store_every), C.c_int64(max_rows), ptr(rows), ptr(nrows), ptr(nstored), ptr(counters), ptr(status),
        ptr(tcur)))
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, dt=dt)


def bounce_setup(field, state, mu, mass, fieldlineresolution=None, arith="strict", max_pts=256):
    """Device part of GuidingCenter.bounceperiod (GuidingCenter.py:593-606, fieldline.py:13-105):
    mirror field, speed, ds and the traced field line of every guiding centre.
    Returns dict(Bm, v, ds, npts, curve (n,max_pts,5): s,x,y,z,|B|)."""
    from . import params as gp
    f = _field_desc(field)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 5)
    n = len(st)
    flr = float(gp["fieldlineresolution"] if fieldlineresolution is None else fieldlineresolution)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(5)]
    mu, mass = _col(mu, n), _col(mass, n)
    while True:
        Bm = np.zeros(n); v = np.zeros(n); ds = np.zeros(n); npts = np.zeros(n, np.int32)
        curve = np.empty((n, max_pts, 5))
        check(_lib.load().rapt_b200_bounce_setup(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_double(flr),
                                                 C.c_int64(n), *[ptr(c_) for c_ in cols], ptr(mu), ptr(mass),
                                                 ptr(Bm), ptr(v), ptr(ds), ptr(npts), C.c_int64(max_pts), ptr(curve)))
        if npts.max(initial=0) <= max_pts:
            return dict(Bm=Bm, v=v, ds=ds, npts=npts, curve=curve)
        max_pts = int(2 ** math.ceil(math.log2(npts.max() + 1)))


def fieldline_trace(field, tpos, Bm, fieldlineresolution=None, arith="strict", max_pts=256):
    """Fieldline(tpos, field, Bmax=Bm).trace() (fieldline.py:13-105) on the device for one start point.
    Returns dict(curve (k,5): s,x,y,z,|B| ; ds)."""
    from . import params as gp
    f = _field_desc(field)
    flr = float(gp["fieldlineresolution"] if fieldlineresolution is None else fieldlineresolution)
    tpos = np.asarray(tpos, dtype=np.float64)
    cols = [np.array([tpos[i]]) for i in range(4)]
    while True:
        Bm_a = np.array([float(Bm)]); ds = np.zeros(1); npts = np.zeros(1, np.int32); curve = np.empty((1, max_pts, 5))
        check(_lib.load().rapt_b200_bounce_setup(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_double(flr),
                                                 C.c_int64(1), *[ptr(c_) for c_ in cols], None, None, None,
                                                 ptr(Bm_a), None, ptr(ds), ptr(npts), C.c_int64(max_pts), ptr(curve)))
        if npts[0] <= max_pts:
            return dict(curve=curve[0, :npts[0]].copy(), ds=float(ds[0]))
        max_pts *= 4


def fieldline_trace_many(field, tpos, Bm, fieldlineresolution=None, arith="strict", max_pts=256):
    """Fieldline(tpos[i], field, Bmax=Bm[i]).trace() for n start points in one device call.
    Returns a list of (k_i, 5) arrays s,x,y,z,|B| and the array of ds."""
    from . import params as gp
    f = _field_desc(field)
    flr = float(gp["fieldlineresolution"] if fieldlineresolution is None else fieldlineresolution)
    tpos = np.asarray(tpos, dtype=np.float64).reshape(-1, 4)
    n = len(tpos)
    cols = [np.ascontiguousarray(tpos[:, i]).copy() for i in range(4)]
    Bm_a = _col(Bm, n)
    while True:
        ds = np.zeros(n); npts = np.zeros(n, np.int32); curve = np.empty((n, max_pts, 5))
        check(_lib.load().rapt_b200_bounce_setup(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_double(flr),
                                                 C.c_int64(n), *[ptr(c_) for c_ in cols], None, None, None,
                                                 ptr(Bm_a), None, ptr(ds), ptr(npts), C.c_int64(max_pts), ptr(curve)))
        if npts.max(initial=0) <= max_pts:
            return [curve[i, :npts[i]].copy() for i in range(n)], ds
        max_pts = int(2 ** math.ceil(math.log2(npts.max() + 1)))


def eye(field, tpos, Bm, fieldlineresolution=None, arith="strict"):
    """flutils.eye (rapt/flutils.py:65-151) for n (t, x, y, z) start points, all on the device: field-line trace, then
    scipy's spline / brentq / QUADPACK route (or the Simpson branch below 70 degrees) per thread."""
    from . import params as gp
    f = _field_desc(field)
    tp = np.asarray(tpos, dtype=np.float64).reshape(-1, 4)
    n = len(tp)
    flr = float(gp["fieldlineresolution"] if fieldlineresolution is None else fieldlineresolution)
    cols = [np.ascontiguousarray(tp[:, i]).copy() for i in range(4)]
    Bm = _col(Bm, n)
    out = np.zeros(n); status = np.zeros(n, np.int32)
    check(_lib.load().rapt_b200_second_invariant(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int64(n),
                                                 *[ptr(c_) for c_ in cols], ptr(Bm), C.c_double(flr), ptr(out), ptr(status)))
    if np.any(status != 1):
        raise AssertionError("field-line trace does not bracket the mirror points")      # flutils.py:117
    return out


QUADRATURE = {"closed": 0, "quadpack": 1, 0: 0, 1: 1}


def bounceperiod_device(field, state, mu, mass, fieldlineresolution=None, arith="strict", quadrature="quadpack"):
    """GuidingCenter.bounceperiod for n guiding centres entirely on the device: field-line trace + scipy's
    quadratic spline rebuilt per thread, then
      quadrature="quadpack": brentq + QUADPACK QAGS restated per thread (flutils.py:308-314 as the reference runs
                             them; the reference's value to ~1e-9),
      quadrature="closed":   mirror points and integral in closed form (no quadrature error: differs from the
                             reference by its own QUADPACK error, 1e-7 typical, 2e-5 worst seen)."""
    from . import params as gp
    f = _field_desc(field)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 5)
    n = len(st)
    flr = float(gp["fieldlineresolution"] if fieldlineresolution is None else fieldlineresolution)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(5)]
    mu, mass = _col(mu, n), _col(mass, n)
    period = np.zeros(n)
    check(_lib.load().rapt_b200_bounce_period(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0),
                                              C.c_int(QUADRATURE[quadrature]), C.c_double(flr),
                                              C.c_int64(n), *[ptr(c_) for c_ in cols], ptr(mu), ptr(mass), ptr(period), None))
    return period


def bounce_center_advance(field, state, mu, v, mass, charge, delta, dt=None, store_every=1, max_rows=0, params=None,
                          arith="strict", quadrature="quadpack"):
    """BounceCenter.advance (rapt/BounceCenter.py:206-251) for n bounce centres; state (n,4) = t, x, y, z of the
    last rows.  dt None: BCtimestep * bounce period per tracer, computed on the device.  Returns a dict with the
    new last rows, stored rows (n, max_rows, 4), nrows, nstored, counters (n,4), status, dt."""
    from . import params as gp
    src = dict(gp if params is None else params)
    f = _field_desc(field)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 4)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(4)]
    mu, v, mass, charge = _col(mu, n), _col(v, n), _col(mass, n), _col(charge, n)
    dtin = None if dt is None else _col(dt, n)
    want = store_every > 0 and max_rows > 0
    rows = np.zeros((n, max_rows, 4)) if want else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); dt_out = np.zeros(n)
    rtol, atol = src["solvertolerances"]
    check(_lib.load().rapt_b200_bounce_center_advance(
        C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int(QUADRATURE[quadrature]), C.c_int64(n),
        *[ptr(c_) for c_ in cols], ptr(mu), ptr(v), ptr(mass), ptr(charge), ptr(dtin),
        C.c_double(float(src["BCtimestep"])), C.c_double(float(delta)), C.c_double(float(rtol)), C.c_double(float(atol)),
        C.c_double(float(src["fieldlineresolution"])), C.c_double(float(src["eyegradientstep"])),
        C.c_int64(store_every if want else 0), C.c_int64(max_rows if want else 0), ptr(rows),
        ptr(nrows), ptr(nstored), ptr(counters), ptr(status), ptr(dt_out)))
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters, status=status,
                dt=dt_out)


def bounce_center_terms(field, tpos, Bm, v=None, mass=None, charge=None, params=None, arith="strict", quadrature="quadpack"):
    """flutils.halfbouncepath / eye / gradI (rapt/flutils.py:65-316) and BounceCenter.advance's right-hand side at n
    points tpos (n,4) with mirror fields Bm: dict of Sb (n,), I (n,), gradI (n,3), deriv (n,3), status."""
    from . import params as gp
    src = dict(gp if params is None else params)
    f = _field_desc(field)
    tp = np.asarray(tpos, dtype=np.float64).reshape(-1, 4)
    n = len(tp)
    cols = [np.ascontiguousarray(tp[:, i]).copy() for i in range(4)]
    Bm = _col(Bm, n)
    opt = [None if a is None else _col(a, n) for a in (v, mass, charge)]
    out = np.zeros((n, 8)); status = np.zeros(n, np.int32)
    check(_lib.load().rapt_b200_bounce_center_terms(
        C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int(QUADRATURE[quadrature]), C.c_int64(n),
        *[ptr(c_) for c_ in cols], ptr(Bm), *[ptr(a) for a in opt],
        C.c_double(float(src["fieldlineresolution"])), C.c_double(float(src["eyegradientstep"])), ptr(out), ptr(status)))
    return dict(Sb=out[:, 0], I=out[:, 1], gradI=out[:, 2:5], deriv=out[:, 5:8], status=status)


def switch_p2g(field, prow, mass, charge, arith="strict"):
    """GuidingCenter.init(Particle) (GuidingCenter.py:168-186) for n rows (n,7) -> (grow (n,5), mu, v, status)."""
    f = _field_desc(field)
    prow = np.ascontiguousarray(np.asarray(prow, dtype=np.float64).reshape(-1, 7))
    n = len(prow)
    mass, charge = _col(mass, n), _col(charge, n)
    grow = np.zeros((n, 5)); mu = np.zeros(n); v = np.zeros(n); st = np.zeros(n, np.int32)
    check(_lib.load().rapt_b200_switch_p2g(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int64(n),
                                           ptr(prow), ptr(mass), ptr(charge), ptr(grow), ptr(mu), ptr(v), ptr(st)))
    return grow, mu, v, st


def switch_g2p(field, grow, mu, mass, charge, t_eval=0.0, arith="strict"):
    """Particle.init(GuidingCenter) (Particle.py:149-164) for n rows (n,5) -> prow (n,7)."""
    f = _field_desc(field)
    grow = np.ascontiguousarray(np.asarray(grow, dtype=np.float64).reshape(-1, 5))
    n = len(grow)
    mu, mass, charge = _col(mu, n), _col(mass, n), _col(charge, n)
    prow = np.zeros((n, 7))
    check(_lib.load().rapt_b200_switch_g2p(C.byref(f), C.c_int(1 if arith in ("strict", 1) else 0), C.c_int64(n),
                                           ptr(grow), ptr(mu), ptr(mass), ptr(charge), C.c_double(t_eval), ptr(prow)))
    return prow


def isadiabatic(field, mode, rows, mu, mass, charge, params=None, **over):
    """Particle.isadiabatic (mode 0, rows (n,7)) / GuidingCenter.isadiabatic (mode 1, rows (n,5))."""
    f = _field_desc(field)
    p = snapshot_params(params, **over)
    rows = np.ascontiguousarray(np.asarray(rows, dtype=np.float64))
    rows = rows.reshape(-1, rows.shape[-1])
    n = len(rows)
    mu, mass, charge = _col(mu, n), _col(mass, n), _col(charge, n)
    out = np.zeros(n, np.int32)
    check(_lib.load().rapt_b200_isadiabatic(C.byref(f), C.byref(p), C.c_int(mode), C.c_int64(n), ptr(rows),
                                            C.c_int64(rows.shape[1]), ptr(mu), ptr(mass), ptr(charge), ptr(out)))
    return out.astype(bool)


def adaptive_advance(field, pos, vel, t0, mass, charge, delta, gc_dt, store_every=1, max_rows=0, params=None, **over):
    """Adaptive.__init__ + Adaptive.advance (Adaptive.py:70-104, 187-222) for n tracers.
    Returns dict(rows, nstored, nseg, mode, final (n,8), counters, status, epochs)."""
    f = _field_desc(field)
    p = snapshot_params(params, True, **over)
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3); vel = np.asarray(vel, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    cols = [np.ascontiguousarray(pos[:, i]).copy() for i in range(3)] + [np.ascontiguousarray(vel[:, i]).copy() for i in range(3)]
    t0, mass, charge = _col(t0, n), _col(mass, n), _col(charge, n)
    rows = np.empty((n, max_rows, 8)) if max_rows > 0 else None
    nstored = np.zeros(n, np.int32); nseg = np.zeros(n, np.int32); mode = np.zeros(n, np.int32)
    fin = np.zeros((n, 8)); counters = np.zeros((n, 4), np.int32); status = np.zeros(n, np.int32)
    epochs = C.c_int32(0)
    check(_lib.load().rapt_b200_adaptive_advance(
        C.byref(f), C.byref(p), C.c_int64(n), *[ptr(c_) for c_ in cols], ptr(t0), ptr(mass), ptr(charge),
        C.c_double(gc_dt), C.c_double(delta), C.c_int64(store_every), C.c_int64(max_rows), ptr(rows),
        ptr(nstored), ptr(nseg), ptr(mode), ptr(fin), ptr(counters), ptr(status), C.byref(epochs)))
    st = np.zeros(15 + 4 * max(epochs.value, 0))
    check(_lib.load().rapt_b200_adaptive_last_stats(ptr(st), C.c_int(len(st))))
    keys = ("epochs", "launches_particle", "launches_gc", "tracer_launches_particle", "tracer_launches_gc", "steps_particle",
            "accepted_particle", "calls_particle", "steps_gc", "calls_gc", "ms_particle", "ms_gc", "ms_switch", "ms_epochs")
    return dict(rows=rows, nstored=nstored, nseg=nseg, mode=mode, final=fin, counters=counters, status=status,
                epochs=epochs.value, stats={k: float(v) for k, v in zip(keys, st)},
                per_epoch=st[15:].reshape(-1, 4))


# ------------------------------------------------------------------------------------------------
# device-resident variants: torch CUDA tensors in place, launched on torch's current stream
# ------------------------------------------------------------------------------------------------
def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def particle_advance_dev(field, cols, mass, charge, delta, out, store_every=0, max_rows=0, rows=None,
                         params=None, check_adiabaticity=False, **over):
    """Particle.advance on device-resident state.  cols: 7 float64 CUDA tensors (t,x,y,z,px,py,pz),
    updated in place; out: dict of CUDA tensors nrows,nstored,status (int32 n), counters (int32 n x 4),
    tcur, dt (float64 n).  No host<->device traffic, no synchronisation."""
    f = _field_desc(field)
    p = snapshot_params(params, check_adiabaticity, **over)
    n = cols[0].numel()
    check(_lib.load().rapt_b200_particle_advance_dev(
        C.byref(f), C.byref(p), C.c_int64(n), *[ptr(c_) for c_ in cols], ptr(mass), ptr(charge), C.c_double(delta),
        C.c_int64(store_every), C.c_int64(max_rows), ptr(rows), ptr(out["nrows"]), ptr(out["nstored"]),
        ptr(out["counters"]), ptr(out["status"]), ptr(out["tcur"]), ptr(out["dt"]), _stream_ptr()))


def gc_advance_dev(field, cols, mu, v, mass, charge, dt, delta, out, eom="TaoChanBrizardEOM", store_every=0,
                   max_rows=0, rows=None, params=None, check_adiabaticity=False, **over):
    """GuidingCenter.advance on device-resident state (5 CUDA tensors t,X,Y,Z,ppar)."""
    f = _field_desc(field)
    p = snapshot_params(params, check_adiabaticity, **over)
    n = cols[0].numel()
    check(_lib.load().rapt_b200_gc_advance_dev(
        C.byref(f), C.byref(p), C.c_int(EOM_KIND[eom]), C.c_int64(n), *[ptr(c_) for c_ in cols],
        ptr(mu), ptr(v), ptr(mass), ptr(charge), ptr(dt), C.c_double(delta),
        C.c_int64(store_every), C.c_int64(max_rows), ptr(rows), ptr(out["nrows"]), ptr(out["nstored"]),
        ptr(out["counters"]), ptr(out["status"]), ptr(out["tcur"]), _stream_ptr()))


def final_diagnostics_dev(kind, cols, mass, status, packed, nbins, lo, hi, hist, stats):
    """One kernel pass over device-resident final-state columns (rapt_b200_final_diagnostics_dev): packs them into
    `packed` ([n][ncol], may be this rank's slot of the all-gather buffer), accumulates the histogram of
    kind 0: log10 kinetic energy [eV] / kind 1: r / Re into `hist` (int64 [nbins]) and (ok count, sum q, sum q^2,
    out of range) into `stats` (float64 [4]).  All CUDA tensors; launched on torch's current stream."""
    n = cols[0].numel()
    arr = (C.c_void_p * len(cols))(*[c_.data_ptr() for c_ in cols])
    check(_lib.load().rapt_b200_final_diagnostics_dev(
        C.c_int({"ke": 0, "r": 1}.get(kind, kind)), C.c_int64(n), C.c_int(len(cols)), arr, ptr(mass), ptr(status), ptr(packed),
        C.c_int(nbins), C.c_double(lo), C.c_double(hi), ptr(hist), ptr(stats), _stream_ptr()))


def unshard_dev(gathered, out, n_total, period=0, offsets=None):
    """gathered (world, n_max, ncol) periodic shards -> out (n_total, ncol) in member order (rapt_b200_unshard_dev).
    offsets None: round-robin; else the ShardPlan's table (world + 1 ints from 0 to period)."""
    world, n_max, ncol = gathered.shape
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
    check(_lib.load().rapt_b200_unshard_dev(C.c_int(world), C.c_int(int(period)), ptr(off), C.c_int64(n_max), C.c_int(ncol),
                                            C.c_int64(n_total), ptr(gathered), ptr(out), _stream_ptr()))


def alloc_outputs(n, device):
    import torch
    return dict(nrows=torch.zeros(n, dtype=torch.int32, device=device),
                nstored=torch.zeros(n, dtype=torch.int32, device=device),
                counters=torch.zeros((n, 4), dtype=torch.int32, device=device),
                status=torch.zeros(n, dtype=torch.int32, device=device),
                tcur=torch.zeros(n, dtype=torch.float64, device=device),
                dt=torch.zeros(n, dtype=torch.float64, device=device))
