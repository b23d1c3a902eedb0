"""GPU: size-independent properties at the BASELINE.json ensemble sizes (1 M tracers), where the CPU
oracle cannot be run in full: invariants of the motion, independence from scheduling (work order,
permutation of the input, host vs device buffers), decimation consistency, failure statuses."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from rapt_b200 import engine, _lib
    _lib.init(0)
    return engine


def config2_state(eng, n):
    from rapt_b200 import synth
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    return np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], eng.particle_momentum(vel, ic["mass"])]), ic


def test_config2_full_size_invariants_and_scheduling(eng):
    """1,048,576 protons (config 2), advance(1 s): |p| conserved (static B), every tracer reaches t0+delta,
    and the result does not depend on the work order (longest-first sort on/off -> bit-identical)."""
    n = 1 << 20
    st, ic = config2_state(eng, n)
    f = H.gpu_field("EarthDipole", ())
    a = eng.particle_advance(f, st, ic["mass"], ic["charge"], 1.0, store_every=0, cyclotronresolution=20, sort_by_work=1)
    b = eng.particle_advance(f, st, ic["mass"], ic["charge"], 1.0, store_every=0, cyclotronresolution=20, sort_by_work=0)
    assert np.array_equal(a["state"], b["state"]) and np.array_equal(a["counters"], b["counters"])
    assert np.all(a["status"] == 1)
    assert np.all(a["state"][:, 0] >= 1.0) and np.all(a["state"][:, 0] < 1.0 + a["dt"] * (1 + 1e-12))
    assert np.array_equal(a["nrows"], 1 + np.ceil(np.round(1.0 / a["dt"], 9)).astype(np.int64)) or \
        np.max(np.abs(a["nrows"] - (1 + np.ceil(1.0 / a["dt"])))) <= 1
    p0 = np.linalg.norm(st[:, 4:7], axis=1); p1 = np.linalg.norm(a["state"][:, 4:7], axis=1)
    assert np.max(np.abs(p1 / p0 - 1)) < 5e-3          # momentum is not error-controlled (SI atol, quirk Q5)
    assert np.median(np.abs(p1 / p0 - 1)) < 1e-8
    # scipy's counter identity nfcn = 2*calls + 11*nstep + naccpt
    c = a["counters"].astype(np.int64)
    assert np.array_equal(c[:, 0], 2 * (a["nrows"].astype(np.int64) - 1) + 11 * c[:, 1] + c[:, 2])
    assert np.all(c[:, 1] >= c[:, 2]) and np.all(c[:, 2] >= a["nrows"] - 1)
    # first adiabatic invariant of a subsample: conserved to the guiding-centre approximation's accuracy
    sub = slice(0, 4096)
    _, mu0, _, s0 = eng.switch_p2g(f, st[sub], ic["mass"][sub], ic["charge"][sub])
    _, mu1, _, s1 = eng.switch_p2g(f, a["state"][sub], ic["mass"][sub], ic["charge"][sub])
    ok = (s0 == 0) & (s1 == 0)
    assert ok.mean() > 0.99
    assert np.median(np.abs(mu1[ok] / mu0[ok] - 1)) < 0.05


def test_permutation_invariance_and_fast_vs_strict(eng):
    """Per-particle results do not depend on which lane / which order a tracer is integrated in."""
    n = 50000
    st, ic = config2_state(eng, n)
    f = H.gpu_field("EarthDipole", ())
    a = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.3, store_every=0, cyclotronresolution=20)
    perm = np.random.default_rng(1).permutation(n)
    b = eng.particle_advance(f, st[perm], ic["mass"][perm], ic["charge"][perm], 0.3, store_every=0, cyclotronresolution=20)
    assert np.array_equal(a["state"][perm], b["state"]) and np.array_equal(a["counters"][perm], b["counters"])
    s = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.3, store_every=0, cyclotronresolution=20, arith="strict")
    assert H.vec_relerr(a["state"][:, 1:4], s["state"][:, 1:4]) < 1e-8
    assert H.vec_relerr(a["state"][:, 4:7], s["state"][:, 4:7]) < 1e-8
    assert (a["counters"][:, 1] == s["counters"][:, 1]).mean() > 0.995


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_decimation_and_row_cap(eng, arith):
    n = 300
    st, ic = config2_state(eng, n)
    f = H.gpu_field("EarthDipole", ())
    kw = dict(cyclotronresolution=20, arith=arith)
    full = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.05, store_every=1, max_rows=256, **kw)
    dec = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.05, store_every=3, max_rows=256, **kw)
    cap = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.05, store_every=1, max_rows=5, **kw)
    assert np.array_equal(full["state"], dec["state"]) and np.array_equal(full["state"], cap["state"])
    assert np.all(full["nstored"] == full["nrows"]) and np.all(full["nrows"] <= 256)
    for i in (0, 7, 299):
        k = int(full["nstored"][i]); kd = int(dec["nstored"][i])
        assert np.array_equal(dec["rows"][i, :kd, :7], full["rows"][i, :k:3, :7])
        assert np.array_equal(dec["rows"][i, :kd, 7], full["rows"][i, :k:3, 7])       # cumulative step counter
        assert np.array_equal(full["rows"][i, k - 1, :7], full["state"][i])
        assert np.all(np.diff(full["rows"][i, :k, 7]) >= 1)
    assert np.all(cap["nstored"] == np.minimum(cap["nrows"], 5)) and np.array_equal(cap["nrows"], full["nrows"])


def test_solver_failure_status(eng):
    """nsteps = 500 per output interval: an absurdly long output step fails like scipy (-2) and the row
    loop ends after appending the failed row (Particle.py:304-307)."""
    import oracle as O
    d, par = H.load("g1b_generic")
    f = H.gpu_field("EarthDipole", ())
    o = eng.particle_advance(f, d["traj"][0], float(d["mass"]), float(d["charge"]), 1e6, store_every=1, max_rows=4,
                             cyclotronresolution=1e-4, arith="strict")
    r = O.particle_advance(O.make_field("EarthDipole"), O.make_params(cyclotronresolution=1e-4), d["traj"][0], float(d["mass"]),
                           float(d["charge"]), 1e6, max_rows=4)
    assert o["status"][0] == -2 == r["status"][0]
    assert o["nrows"][0] == 2 == r["nrows"][0]
    assert o["counters"][0, 1] == r["counters"][0, 1] == 501


def test_gc_full_size_energy_conservation(eng):
    """262,144 electrons of config 3 (DoubleDipole, static): kinetic energy of the guiding centre is a
    constant of the Tao-Chan-Brizard equations; check it after 5 s of bounce + drift."""
    from rapt_b200 import synth, c
    n = 1 << 18
    ic = synth.config3_electrons(n)
    f = H.gpu_field("DoubleDipole", ())
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = eng.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
    st = np.column_stack([ic["t0"], pos, ppar])
    o = eng.gc_advance(f, st, mu, ic["v"], ic["mass"], ic["charge"], 0.1, 5.0, store_every=0)
    assert np.all(o["status"] == 1) and np.all((o["nrows"] == 51) | (o["nrows"] == 52))   # 0.1 accumulates in fp

    def gamma(state):
        Bm = eng.field_ops(f, state[:, :4], which=["magB"])["magB"]
        mc = ic["mass"] * c
        return np.sqrt(1 + 2 * mu * Bm / (mc * c) + (state[:, 4] / mc) ** 2)
    g0, g1 = gamma(st), gamma(o["state"])
    assert np.max(np.abs((g1 - 1) / (g0 - 1) - 1)) < 1e-4
    assert np.median(np.abs((g1 - 1) / (g0 - 1) - 1)) < 1e-6


def test_abi_argument_checks_and_degenerate_inputs(eng):
    """Bad arguments return an error code (never crash, never hang); degenerate tracers terminate with a
    failure status instead of looping forever."""
    import ctypes as C
    from rapt_b200 import _lib, fields
    from rapt_b200._lib import ptr
    lib = _lib.load()
    f = fields.EarthDipole().device_descriptor()
    p = eng.snapshot_params(None, False)
    one = np.ones(1); i1 = np.zeros(1, np.int32); i4 = np.zeros(4, np.int32)
    args = [C.byref(f), C.byref(p), C.c_int64(1)] + [ptr(one.copy()) for _ in range(7)] + [ptr(one), ptr(one), C.c_double(1.0),
            C.c_int64(0), C.c_int64(0), None, ptr(i1), ptr(i1.copy()), ptr(i4), ptr(i1.copy()), ptr(one.copy()), ptr(one.copy())]
    bad = list(args); bad[2] = C.c_int64(-1)
    assert lib.rapt_b200_particle_advance(*bad) == -3 and b"n < 0" in lib.rapt_b200_last_error()
    bad = list(args); bad[4] = None
    assert lib.rapt_b200_particle_advance(*bad) == -3
    f2 = fields.EarthDipole().device_descriptor(); f2.kind = 42
    bad = list(args); bad[0] = C.byref(f2)
    assert lib.rapt_b200_particle_advance(*bad) == -3 and b"unknown field kind" in lib.rapt_b200_last_error()
    with pytest.raises(_lib.RaptB200Error):
        eng.particle_advance(fields.EarthDipole(), np.ones((1, 7)), 1.0, 1.0, 1.0, cyclotronresolution=-5)
    with pytest.raises(_lib.RaptB200Error):
        eng.particle_advance(fields.EarthDipole(), np.ones((1, 7)), 1.0, 1.0, 1.0, solvertolerances=(0.0, 1e-8))
    # degenerate tracers: NaN position, zero field (particle at infinity), zero charge -> finite run time, failure status
    d, par = H.load("g1b_generic")
    st = np.tile(d["traj"][0], (4, 1))
    st[1, 1] = np.nan
    st[2, 1:4] = 1e200
    q = np.full(4, float(d["charge"])); q[3] = 0.0
    o = eng.particle_advance(fields.EarthDipole(), st, float(d["mass"]), q, 0.5, store_every=0, cyclotronresolution=20)
    assert o["status"][0] == 1
    assert np.all(o["status"][1:] < 0)
    # guiding centres with a non-positive output step terminate too
    g = eng.gc_advance(fields.DoubleDipole(), np.array([[0.0, 5e7, 1e6, 1e5, 1e-23]]), 1e-7, 1e8, 9.1e-31, -1.6e-19, 0.0, 1.0, store_every=0)
    assert g["status"][0] < 0


def test_launch_shapes_agree(eng):
    """A launch that fills the GPU runs one 512-thread block per SM, a smaller one 128-thread blocks spread over the SMs
    (kernels_tu.cu:block_threads) -- two launches of the same kernel.  131,072 tracers in one call must equal the same
    tracers in eight calls of 16,384, bit for bit: Particle (Nystrom kernel) and GuidingCenter, with stored rows."""
    from rapt_b200 import synth
    n, chunk = 131072, 16384
    st, ic = config2_state(eng, n)
    f = H.gpu_field("EarthDipole", ())
    kw = dict(store_every=5, max_rows=8, cyclotronresolution=20)
    big = eng.particle_advance(f, st, ic["mass"], ic["charge"], 0.05, **kw)
    for k in range(0, n, chunk):
        sl = slice(k, k + chunk)
        small = eng.particle_advance(f, st[sl], ic["mass"][sl], ic["charge"][sl], 0.05, **kw)
        for key in ("state", "counters", "status", "nrows", "nstored", "tcur", "dt"):
            assert np.array_equal(big[key][sl], small[key]), key
        m = int(small["nstored"].max())
        valid = np.arange(m)[None, :] < small["nstored"][:, None]           # rows past a tracer's nstored are not written
        assert m >= 2 and np.array_equal(big["rows"][sl, :m][valid], small["rows"][:, :m][valid])
    assert big["counters"][:, 1].sum() > 5 * n
    ic3 = synth.config3_electrons(n)
    f3 = H.gpu_field("DoubleDipole", ())
    pos = np.column_stack([ic3["x"], ic3["y"], ic3["z"]])
    ppar, mu = eng.gc_construct(f3, ic3["t0"], pos, ic3["v"], ic3["pa"], ic3["mass"])
    st3 = np.column_stack([ic3["t0"], pos, ppar])
    big = eng.gc_advance(f3, st3, mu, ic3["v"], ic3["mass"], ic3["charge"], 0.1, 0.5, store_every=2, max_rows=4)
    for k in range(0, n, 4 * chunk):
        sl = slice(k, k + 4 * chunk)
        small = eng.gc_advance(f3, st3[sl], mu[sl], ic3["v"][sl], ic3["mass"][sl], ic3["charge"][sl], 0.1, 0.5, store_every=2, max_rows=4)
        for key in ("state", "counters", "status", "nrows", "nstored", "tcur"):
            assert np.array_equal(big[key][sl], small[key]), key
        m = int(small["nstored"].max())
        valid = np.arange(m)[None, :] < small["nstored"][:, None]
        assert m >= 2 and np.array_equal(big["rows"][sl, :m][valid], small["rows"][:, :m][valid])


def test_work_order_from_previous_call(eng):
    """sort_by_work = 2 (device-resident ensembles): the previous call's step counts order the next call.  Scheduling only:
    two consecutive advances are bit-identical to the same two advances with the predicted order, counters included; the
    first call (no history in the counters buffer) runs with the predicted order."""
    import rapt_b200 as R
    from rapt_b200 import synth
    n = 65536
    ic = synth.config2_protons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    res = []
    for order in (1, 2):
        ens = R.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], R.fields.EarthDipole()).cuda("cuda:0")
        ens.advance(0.25, cyclotronresolution=20, sort_by_work=order).advance(0.25, cyclotronresolution=20, sort_by_work=order)
        ens.pull()
        res.append((ens.state.copy(), ens.counters.copy(), ens.last_counters.copy(), ens.status.copy()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert np.all(res[0][3] == 1) and res[0][2][:, 1].min() > 0
    # guiding centres: a-priori key (transit time) on the first call, the previous call's step counts afterwards
    ic = synth.config3_electrons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    res = []
    for order in (1, 2):
        ens = R.GuidingCenterEnsemble(pos, ic["v"], pa=ic["pa"], mass=ic["mass"], charge=ic["charge"],
                                      field=R.fields.DoubleDipole()).cuda("cuda:0")
        ens.advance(0.5, dt=0.1, sort_by_work=order).advance(0.5, dt=0.1, sort_by_work=order)
        ens.pull()
        res.append((ens.state.copy(), ens.counters.copy(), ens.last_counters.copy(), ens.status.copy()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert res[0][2][:, 1].min() > 0
