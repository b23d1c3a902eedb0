"""TEST INFRASTRUCTURE ONLY: ctypes access to tests/hostcheck/libkernelhost.so, the product's advance kernels
compiled for the host (tests/hostcheck/kernel_host.cpp).  Same argument conventions as rapt_b200.engine."""
import ctypes as C
import os
import numpy as np
from rapt_b200 import engine
from rapt_b200._lib import EOM_KIND, ptr

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")
_libs = {}


def lib(arith="fast"):
    """fast: libkernelhost.so (FMA contraction on); strict: libkernelhost_strict.so (the reference's operation order)."""
    if arith not in _libs:
        name = {"strict": "libkernelhost_strict.so", "fast-defer": "libkernelhost_defer.so"}.get(arith, "libkernelhost.so")
        _libs[arith] = C.CDLL(os.path.join(_DIR, name))
    return _libs[arith]


def _col(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))).copy()


def particle_advance(field, state, mass, charge, delta, store_every=1, max_rows=0, rkn=True, nthreads=4,
                     check_adiabaticity=False, arith="fast", **over):
    f = engine._field_desc(field)
    p = engine.snapshot_params(None, check_adiabaticity, arith=arith, **over)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 7)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(7)]
    mass = _col(mass, n); charge = _col(charge, n)
    want = store_every > 0 and max_rows > 0
    rows = np.zeros((n, max_rows, 8)) if want else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); tcur = np.zeros(n); dt = np.zeros(n)
    rc = lib(arith).hc_particle_advance(
        C.byref(f), C.byref(p), C.c_longlong(n), *[ptr(c_) for c_ in cols], ptr(mass), ptr(charge), C.c_double(delta),
        C.c_longlong(store_every), C.c_longlong(max_rows), ptr(rows), ptr(nrows), ptr(nstored), ptr(counters),
        ptr(status), ptr(tcur), ptr(dt), C.c_int(1 if rkn else 0), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, dt=dt)


def gc_advance(field, state, mu, v, mass, charge, dt, delta, eom="TaoChanBrizardEOM", store_every=1, max_rows=0,
               nthreads=4, check_adiabaticity=False, arith="fast", **over):
    f = engine._field_desc(field)
    p = engine.snapshot_params(None, check_adiabaticity, arith=arith, **over)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 5)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(5)]
    mu, v, mass, charge, dt = _col(mu, n), _col(v, n), _col(mass, n), _col(charge, n), _col(dt, n)
    want = store_every > 0 and max_rows > 0
    rows = np.zeros((n, max_rows, 8)) if want else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); tcur = np.zeros(n)
    rc = lib(arith).hc_gc_advance(
        C.byref(f), C.byref(p), C.c_int(EOM_KIND[eom]), C.c_longlong(n), *[ptr(c_) for c_ in cols],
        ptr(mu), ptr(v), ptr(mass), ptr(charge), ptr(dt), C.c_double(delta),
        C.c_longlong(store_every), C.c_longlong(max_rows), ptr(rows), ptr(nrows), ptr(nstored), ptr(counters),
        ptr(status), ptr(tcur), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters,
                status=status, tcur=tcur, dt=dt)


def adaptive_advance(field, pos, vel, t0, mass, charge, delta, gc_dt, store_every=1, max_rows=0, nthreads=4,
                     arith="fast", **over):
    """Same arguments and result dict as rapt_b200.engine.adaptive_advance."""
    f = engine._field_desc(field)
    p = engine.snapshot_params(None, True, arith=arith, **over)
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3); vel = np.asarray(vel, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    cols = [np.ascontiguousarray(pos[:, i]).copy() for i in range(3)] + [np.ascontiguousarray(vel[:, i]).copy() for i in range(3)]
    t0, mass, charge = _col(t0, n), _col(mass, n), _col(charge, n)
    rows = np.zeros((n, max_rows, 8)) if max_rows > 0 else None
    nstored = np.zeros(n, np.int32); nseg = np.zeros(n, np.int32); mode = np.zeros(n, np.int32)
    fin = np.zeros((n, 8)); counters = np.zeros((n, 4), np.int32); status = np.zeros(n, np.int32)
    epochs = C.c_int32(0)
    rc = lib(arith).hc_adaptive_advance(
        C.byref(f), C.byref(p), C.c_longlong(n), *[ptr(c_) for c_ in cols], ptr(t0), ptr(mass), ptr(charge),
        C.c_double(gc_dt), C.c_double(delta), C.c_longlong(store_every), C.c_longlong(max_rows), ptr(rows),
        ptr(nstored), ptr(nseg), ptr(mode), ptr(fin), ptr(counters), ptr(status), C.byref(epochs), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(rows=rows, nstored=nstored, nseg=nseg, mode=mode, final=fin, counters=counters, status=status,
                epochs=epochs.value)


def field_ops(field, tpos, arith="strict", which=None):
    """Same arguments and result dict as rapt_b200.engine.field_ops (k_field_ops on the host)."""
    f = engine._field_desc(field)
    tpos = np.ascontiguousarray(np.asarray(tpos, dtype=np.float64).reshape(-1, 4))
    n = len(tpos)
    shapes = dict(B=(n, 3), E=(n, 3), unitb=(n, 3), magB=(n,), gradB=(n, 3), jacobianB=(n, 3, 3), curlb=(n, 3),
                  curvature=(n,), dBdt=(n,), dbdt=(n, 3), lengthscale=(n,), timescale=(n,))
    names = list(shapes)
    which = names if which is None else list(which)
    out = {k: (np.zeros(shapes[k]) if k in which else None) for k in names}
    rc = lib(arith).hc_field_ops(C.byref(f), C.c_longlong(n), ptr(tpos), *[ptr(out[k]) for k in names])
    assert rc == 0, rc
    return {k: v for k, v in out.items() if v is not None}


def bounce(field, state, mu, mass, fieldlineresolution=50.0, arith="strict", max_pts=512, quadrature=1):
    """k_bounce_setup on the host: mirror field, speed, ds, the traced field line and the bounce period
    (quadrature 1: brentq + QAGS as the reference, 0: closed form) of every guiding centre."""
    f = engine._field_desc(field)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 5)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(5)]
    mu, mass = _col(mu, n), _col(mass, n)
    Bm = np.zeros(n); v = np.zeros(n); ds = np.zeros(n); npts = np.zeros(n, np.int32); period = np.zeros(n)
    curve = np.zeros((n, max_pts, 5))
    rc = lib(arith).hc_bounce(C.byref(f), C.c_int(quadrature), C.c_double(fieldlineresolution), C.c_longlong(n),
                              *[ptr(c_) for c_ in cols], ptr(mu), ptr(mass), ptr(Bm), ptr(v), ptr(ds), ptr(npts),
                              C.c_longlong(max_pts), ptr(curve), ptr(period))
    assert rc == 0, rc
    return dict(Bm=Bm, v=v, ds=ds, npts=npts, curve=curve, period=period)


def _bc_call(arith, f, op, n, cols, mu, Bm, v, mass, charge, dtin, src, delta, store_every, max_rows, rows, nrows, nstored,
             counters, status, dt_out, out, max_pts=1024):
    rtol, atol = src["solvertolerances"]
    rc = lib(arith).hc_bounce_center(
        C.byref(f), C.c_int(op), C.c_int(1), C.c_longlong(n), *[ptr(c_) for c_ in cols], ptr(mu), ptr(Bm), ptr(v), ptr(mass),
        ptr(charge), ptr(dtin), C.c_double(float(src["BCtimestep"])), C.c_double(float(delta)), C.c_double(float(rtol)),
        C.c_double(float(atol)), C.c_double(float(src["fieldlineresolution"])), C.c_double(float(src["eyegradientstep"])),
        C.c_longlong(store_every), C.c_longlong(max_rows), ptr(rows), ptr(nrows), ptr(nstored), ptr(counters), ptr(status),
        ptr(dt_out), ptr(out), C.c_longlong(max_pts))
    assert rc == 0, rc


def bounce_center_advance(field, state, mu, v, mass, charge, delta, store_every=1, max_rows=0, arith="strict"):
    """k_bounce_center op 0 on the host; arguments and result as rapt_b200.engine.bounce_center_advance."""
    from rapt_b200 import params as gp
    f = engine._field_desc(field)
    st = np.asarray(state, dtype=np.float64).reshape(-1, 4)
    n = len(st)
    cols = [np.ascontiguousarray(st[:, i]).copy() for i in range(4)]
    mu, v, mass, charge = _col(mu, n), _col(v, n), _col(mass, n), _col(charge, n)
    rows = np.zeros((n, max_rows, 4)) if max_rows > 0 else None
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    status = np.zeros(n, np.int32); dt_out = np.zeros(n)
    _bc_call(arith, f, 0, n, cols, mu, None, v, mass, charge, None, dict(gp), delta, store_every, max_rows, rows, nrows,
             nstored, counters, status, dt_out, None)
    return dict(state=np.column_stack(cols), rows=rows, nrows=nrows, nstored=nstored, counters=counters, status=status, dt=dt_out)


def bounce_center_terms(field, tpos, Bm, v, mass, charge, arith="strict"):
    """k_bounce_center op 1 on the host: S_b, I, gradI and the right-hand side at the given points."""
    from rapt_b200 import params as gp
    f = engine._field_desc(field)
    tp = np.asarray(tpos, dtype=np.float64).reshape(-1, 4)
    n = len(tp)
    cols = [np.ascontiguousarray(tp[:, i]).copy() for i in range(4)]
    Bm, v, mass, charge = _col(Bm, n), _col(v, n), _col(mass, n), _col(charge, n)
    out = np.zeros((n, 8)); status = np.zeros(n, np.int32)
    nrows = np.zeros(n, np.int32); nstored = np.zeros(n, np.int32); counters = np.zeros((n, 4), np.int32)
    _bc_call(arith, f, 1, n, cols, None, Bm, v, mass, charge, None, dict(gp), 0.0, 0, 0, None, nrows, nstored, counters,
             status, None, out)
    return dict(Sb=out[:, 0], I=out[:, 1], gradI=out[:, 2:5], deriv=out[:, 5:8], status=status)
