"""GPU parity: GuidingCenter.advance, GuidingCenter.__init__, the bounce-period set-up, the field
operators and the mode-switch transforms on the B200 vs golden vectors from the reference and vs
the CPU oracle."""
import numpy as np
import pytest

import helpers as H
import scipy_legs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from rapt_b200 import engine, _lib
    _lib.init(0)
    return engine


FIELDS = {"earthdipole": ("EarthDipole", ()), "doubledipole": ("DoubleDipole", ()), "uniformbz": ("UniformBz", (2e-4,)),
          "crossedeb": ("UniformCrossedEB", (2.0, 1e-4)), "vardipole": ("VarEarthDipole", (0.1, 10)),
          "parabolic": ("Parabolic", ())}


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("name", list(FIELDS))
def test_field_ops_vs_reference(eng, name, arith):
    """_Field operators (fields.py:76-280) at seeded points; golden values from the reference."""
    u = np.load(H.GOLDEN + "/units.npz")
    o = eng.field_ops(H.gpu_field(*FIELDS[name]), u[name + "_pts"], arith=arith)
    for k in ("B", "E", "unitb", "magB"):
        # pow() on the GPU is within 2 ulp of glibc's; rsqrt chain in fast mode a few ulp more
        assert H.relerr(o[k], u[f"{name}_{k}"], floor=1e-300) < (1e-14 if arith == "strict" else 1e-13), k
    # finite differences amplify ulp noise by |x|/d: compare against the vector norm
    d = H.gpu_field(*FIELDS[name]).gradientstepsize
    for k in ("gradB", "curlb", "jacobianB"):
        g = u[f"{name}_{k}"]; m = o[k]
        scale = np.max(np.abs(g)) + 1e-300
        noise = 1e-15 * (np.max(np.abs(u[name + "_pts"][:, 1:])) / d + 1) * 50
        if k == "curlb" and name == "parabolic":
            noise = 1e-9      # unit vectors of O(1) differenced over d = 1e-6
        assert np.max(np.abs(m - g)) / scale < max(noise, 1e-13), (k, np.max(np.abs(m - g)) / scale)
    # time derivatives: central difference over timederivstepsize = 1e-3 s; db/dt of a dipole whose
    # direction does not change is pure round-off (~1e-16/1e-3), so compare on that absolute scale
    g = u[f"{name}_dBdt"]; m = o["dBdt"]
    assert np.max(np.abs(m - g)) <= 1e-9 * np.max(np.abs(g)) + 1e-12 * np.max(np.abs(u[f"{name}_magB"])) / 1e-3
    g = u[f"{name}_dbdt"]; m = o["dbdt"]
    assert np.max(np.abs(m - g)) <= 1e-9 * np.max(np.abs(g)) + 1e-12 / 1e-3
    for k in ("lengthscale", "curvature"):
        g = u[f"{name}_{k}"]; m = o[k]
        fin = np.isfinite(g)
        assert np.array_equal(np.isfinite(m), fin)
        if fin.any():
            assert H.relerr(m[fin], g[fin]) < 1e-6, k


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("name", list(H.GC_CASES))
def test_gc_golden(eng, name, arith):
    d, par = H.load(name)
    fname, fargs = H.GC_CASES[name]
    f = H.gpu_field(fname, fargs)
    traj = d["traj"]
    mass, q, v = float(d["mass"]), float(d["charge"]), float(d["v"])
    ppar, mu = eng.gc_construct(f, traj[0, 0], d["pos"], v, float(d["pa"]), mass, arith=arith)
    assert H.relerr(mu, float(d["mu"])) < 1e-13 and abs(ppar[0] - traj[0, 4]) <= 1e-15 * abs(traj[0, 4])
    st0 = np.concatenate(([traj[0, 0]], d["pos"], ppar))
    if "bs_period" in d.files:
        # GuidingCenter.bounceperiod: device field-line trace + host quadrature (scipy, as the reference)
        bs = eng.bounce_setup(f, st0, mu, mass, arith=arith)
        k = int(bs["npts"][0])
        assert k == len(d["bs_curve"]), "field-line trace must return the same number of points"
        assert H.relerr(bs["Bm"][0], float(d["bs_Bm"])) < 1e-13
        assert H.relerr(bs["ds"][0], float(d["bs_ds"])) < 1e-6        # curvature is a finite difference
        assert np.max(np.abs(bs["curve"][0, :k, :4] - d["bs_curve"])) < 1e-6 * np.max(np.abs(d["bs_curve"]))
        bp = scipy_legs.bounceperiod(f, st0, mu, mass, arith=arith)[0]
        assert abs(bp / float(d["bs_period"]) - 1) < 1e-6
        dt = float(d["bs_period"]) / par.get("bounceresolution", 10)   # feed the reference's dt
    else:
        dt = par["GCtimestep"]
    eom = str(d["eom"]) if "eom" in d.files else "TaoChanBrizardEOM"
    gpar = {k: v_ for k, v_ in par.items() if k in ("solvertolerances", "enforce equatorial")}
    o = eng.gc_advance(f, st0, mu, v, mass, q, dt, float(d["delta"]), eom=eom, store_every=1, max_rows=len(traj) + 8,
                       arith=arith, **gpar)
    n = int(o["nstored"][0])
    assert o["status"][0] == 1
    assert o["nrows"][0] == len(traj) == n
    rows = o["rows"][0, :n]
    assert H.relerr(rows[:, 0], traj[:, 0]) < 1e-13
    assert H.vec_relerr(rows[:, 1:4], traj[:, 1:4]) < 1e-8
    pscale = max(np.max(np.abs(traj[:, 4])), 1e-3 * mass * v)      # p_par stays exactly 0 for pa = 90 in the reference
    assert np.max(np.abs(rows[:, 4] - traj[:, 4])) / pscale < 1e-7
    ref = d["counters"].sum(0)
    got = o["counters"][0]
    if name in ("gc_pa90_equatorial", "gc_equatorial_enforced", "g2_gc_doubledipole", "gc_earthdipole"):
        # starts with an exact-zero coordinate / zero p_par: sk = atol there, so HINIT and the first steps of
        # every row are round-off dominated; the strict flavour stays within a few steps, the
        # fast flavour (reciprocals, fused ops) within ~25 %: with p_par = z = 0 exactly the reference's HINIT
        # sees exact zeros where the GPU sees 1e-20-level noise over sk = atol, so the first step of each row
        # (then grown x10 per step) starts from a different round-off value -- 214 vs 181 steps over 59 rows
        # for gc_pa90_equatorial.  Trajectories agree to 1e-8 either way (asserted above).
        assert abs(int(got[1]) - int(ref[1])) <= max(3, (0.01 if arith == "strict" else 0.25) * ref[1]), (got.tolist(), ref.tolist())
    else:
        assert abs(int(got[1]) - int(ref[1])) <= (0 if arith == "strict" else 2), (got, ref)
    assert abs(o["tcur"][0] - float(d["tcur"])) <= 1e-12 * abs(float(d["tcur"]))
    assert np.all(rows[:, 5] == mu[0])


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("case", ["e3_config3_first16", "e5_config5_first16"])
def test_gc_ensembles_vs_reference(eng, case, arith):
    from rapt_b200 import synth
    d, par = H.load(case)
    n = int(d["n"])
    if case.startswith("e3"):
        ic = synth.config3_electrons(n); f = H.gpu_field("DoubleDipole", ())
    else:
        ic = synth.config5_belt(n); f = H.gpu_field("VarEarthDipole", (0.1, 10))
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = eng.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"], arith=arith)
    assert H.relerr(mu, d["mu"]) < 1e-12
    st0 = np.column_stack([ic["t0"], pos, ppar])
    o = eng.gc_advance(f, st0, mu, ic["v"], ic["mass"], ic["charge"], par["GCtimestep"], float(d["delta"]),
                       store_every=0, arith=arith)
    fin = d["final"]
    assert np.array_equal(o["nrows"], d["nrows"])
    assert H.vec_relerr(o["state"][:, 1:4], fin[:, 1:4]) < 1e-8
    assert np.max(np.abs(o["state"][:, 4] - fin[:, 4]) / np.max(np.abs(fin[:, 4]))) < 1e-7
    dn = np.abs(o["counters"][:, 1].astype(int) - d["totals"][:, 1].astype(int))
    assert dn.max() <= max(2, 0.005 * d["totals"][:, 1].max()), dn
    if case.startswith("e3"):
        bp = scipy_legs.bounceperiod(f, st0, mu, ic["mass"], arith=arith)
        assert np.max(np.abs(bp / d["bounceperiod"] - 1)) < 1e-5


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_gc_ensemble_vs_oracle(eng, arith):
    """2048 electrons of config 3 in DoubleDipole, 5 s, GCtimestep 0.1: CUDA vs CPU oracle."""
    import oracle as O
    from rapt_b200 import synth
    n = 2048
    ic = synth.config3_electrons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    of = O.make_field("DoubleDipole")
    ppar_o, mu_o = O.gc_construct(of, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
    f = H.gpu_field("DoubleDipole", ())
    ppar, mu = eng.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"], arith=arith)
    assert H.relerr(mu, mu_o) < 1e-12
    st0 = np.column_stack([ic["t0"], pos, ppar_o])
    ref = O.gc_advance(of, O.make_params(), st0, mu_o, ic["v"], ic["mass"], ic["charge"], 0.1, 5.0,
                       store_every=5, max_rows=16, nthreads=8)
    o = eng.gc_advance(f, st0, mu_o, ic["v"], ic["mass"], ic["charge"], 0.1, 5.0, store_every=5, max_rows=16, arith=arith)
    assert np.array_equal(o["nrows"], ref["nrows"]) and np.array_equal(o["nstored"], ref["nstored"])
    assert H.vec_relerr(o["state"][:, 1:4], ref["state"][:, 1:4]) < 1e-8
    same = (o["counters"][:, 1] == ref["counters"][:, 1]).mean()
    assert same > 0.97, same
    assert abs(int(o["counters"][:, 1].sum()) - int(ref["counters"][:, 1].sum())) <= 2e-3 * ref["counters"][:, 1].sum()
    k = int(o["nstored"][5])
    assert H.vec_relerr(o["rows"][5, :k, 1:4], ref["rows"][5, :k, 1:4]) < 1e-8


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_switch_transforms_and_predicates(eng, arith):
    """P->G (utils.guidingcenter iteration) and G->P (GCtoFP) vs golden values and the oracle."""
    import oracle as O
    from rapt_b200 import m_pr, e, c
    u = np.load(H.GOLDEN + "/units.npz")
    f = H.gpu_field("DoubleDipole", ()); of = O.make_field("DoubleDipole")
    pos, vel = u["utils_pos"], u["utils_vel"]
    mom = eng.particle_momentum(vel, np.full(len(pos), m_pr))
    prow = np.column_stack([np.zeros(len(pos)), pos, mom])
    grow, mu, v, st = eng.switch_p2g(f, prow, m_pr, e, arith=arith)
    assert np.all(st == 0)
    assert H.vec_relerr(grow[:, 1:4], u["utils_gc_R"]) < 1e-12
    assert H.relerr(v, u["utils_gc_v"]) < 1e-13
    assert H.relerr(mu, u["utils_mu"]) < 1e-9       # mu ~ (v - vpar)(v + vpar): cancellation near 0/180 deg
    gamma = 1 / np.sqrt(1 - (u["utils_gc_v"] / c) ** 2)
    assert np.max(np.abs(grow[:, 4] - m_pr * gamma * u["utils_gc_vp"])) < 1e-12 * np.max(np.abs(grow[:, 4]))
    back = eng.switch_g2p(f, grow, mu, m_pr, e, 0.0, arith=arith)
    for i in range(len(pos)):
        ref = O.switch_G2P(of, grow[i], mu[i], m_pr, e, 0.0)
        assert H.vec_relerr(back[i, 1:4], ref[1:4]) < 1e-12
        assert H.vec_relerr(back[i, 4:7], ref[4:7]) < 1e-9
    # predicates
    par = dict(epss=0.02)
    pa_gpu = eng.isadiabatic(f, 0, prow, 0.0, m_pr, e, arith=arith, **par)
    pa_ref = [O.particle_isadiabatic(of, O.make_params(**par), r, m_pr, e) for r in prow]
    assert list(pa_gpu) == pa_ref
    ga_gpu = eng.isadiabatic(f, 1, grow, mu, m_pr, e, arith=arith, **par)
    ga_ref = [O.gc_isadiabatic(of, O.make_params(**par), grow[i], mu[i], m_pr, e) for i in range(len(grow))]
    assert list(ga_gpu) == ga_ref


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_bounce_period_device_closed_form(eng, arith):
    """rapt_b200_bounce_period: trace + spline + closed-form quadrature on the device vs the reference's
    bounce periods (golden) -- agreement to QUADPACK's own error (epsrel 1e-4 requested, ~1e-7 delivered)."""
    from rapt_b200 import synth
    for name in ("g2_gc_doubledipole", "gc_earthdipole", "gc_pa90_equatorial"):
        d, par = H.load(name)
        f = H.gpu_field(*H.GC_CASES[name])
        traj = d["traj"]
        st0 = traj[0]
        bp = eng.bounceperiod_device(f, st0, float(d["mu"]), float(d["mass"]), arith=arith)[0]
        assert abs(bp / float(d["bs_period"]) - 1) < 2e-6, (name, bp, float(d["bs_period"]))
    d, par = H.load("e3_config3_first16")
    n = int(d["n"]); ic = synth.config3_electrons(n); f = H.gpu_field("DoubleDipole", ())
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = eng.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"], arith=arith)
    st = np.column_stack([ic["t0"], pos, ppar])
    bp = eng.bounceperiod_device(f, st, mu, ic["mass"], arith=arith)
    # the closed form is the exact integral of the spline; the difference is the error of the reference's
    # QUADPACK call, which asked for epsrel = 1e-4 (flutils.py:314): typically 1e-7, a few 1e-6 observed
    assert np.max(np.abs(bp / d["bounceperiod"] - 1)) < 1e-4
    assert np.median(np.abs(bp / d["bounceperiod"] - 1)) < 1e-6
    # 20,000 guiding centres in one call: finite, positive, and consistent with the host (scipy) leg on a subset
    ic = synth.config3_electrons(20000)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    ppar, mu = eng.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"], arith=arith)
    st = np.column_stack([ic["t0"], pos, ppar])
    bp = eng.bounceperiod_device(f, st, mu, ic["mass"], arith=arith)
    assert np.isfinite(bp).mean() > 0.999 and np.nanmin(bp) > 0
    host = scipy_legs.bounceperiod(f, st[:64], mu[:64], ic["mass"][:64], arith=arith)
    ok = np.isfinite(bp[:64])
    assert np.max(np.abs(bp[:64][ok] / host[ok] - 1)) < 1e-4
    assert np.median(np.abs(bp[:64][ok] / host[ok] - 1)) < 2e-6


@pytest.mark.parametrize("name", ["eye_pa80", "eye_pa45_simpson"])
def test_geteye_matches_reference(eng, name):
    """GuidingCenter.geteye (GuidingCenter.py:608-624): the second invariant along a trajectory, field lines
    traced on the device.  The golden values are the reference's own (see the fixture's note on `simps`)."""
    import rapt_b200 as R
    from rapt_b200 import GuidingCenter, params
    from rapt_b200.fields import EarthDipole
    d, par = H.load(name)
    old = params["GCtimestep"]
    params["GCtimestep"] = par["GCtimestep"]
    try:
        g = GuidingCenter(pos=tuple(d["pos"]), v=float(d["v"]), pa=float(d["pa"]), mass=float(d["mass"]),
                          charge=float(d["charge"]), field=EarthDipole())
        g.advance(float(d["delta"]))
    finally:
        params["GCtimestep"] = old
    assert g.trajectory.shape == d["traj"].shape
    out = g.geteye(step=int(d["step"]))
    assert out.shape == d["eye"].shape
    assert np.allclose(out[:, 0], d["eye"][:, 0], rtol=0, atol=1e-12)
    assert np.max(np.abs(out[:, 1] / d["eye"][:, 1] - 1)) < 1e-7
    # the same from the reference's own rows: isolates the trace + quadrature from the advance
    ref_rows = d["traj"][::int(d["step"])]
    val = eng.eye(EarthDipole(), ref_rows[:, :4], d["Bm"])
    assert np.max(np.abs(val / d["eye"][:, 1] - 1)) < 1e-8
    # device quadrature (spline / brentq / QAGS or Simpson per thread) against scipy's own on the same device traces
    host = scipy_legs.eye_host(EarthDipole(), ref_rows[:, :4], d["Bm"])
    assert np.max(np.abs(val / host - 1)) < 1e-10
