"""CPU: the oracle (oracle/rapt_oracle.c) against the golden vectors generated from the unmodified
reference (tests/golden/, oracle/gen_golden.py) and against scipy's `_dop` directly.  This is the
parity PIN of the oracle: trajectories and solver counters must agree bit for bit."""
import os

import numpy as np
import pytest

import helpers as H
import oracle as O


@pytest.mark.parametrize("name", list(H.PARTICLE_CASES) + ["p_chargeddipole"])
def test_particle_bit_exact(name):
    d, par = H.load(name)
    fname, fargs = H.PARTICLE_CASES.get(name, ("ChargedDipole", (1, 1e-6)))
    traj = d["traj"]
    o = O.particle_advance(O.make_field(fname, *fargs), O.make_params(**par), traj[0], float(d["mass"]),
                           float(d["charge"]), float(d["delta"]), max_rows=len(traj) + 10, want_percall=True)
    n = int(o["nstored"][0])
    assert n == len(traj)
    assert np.array_equal(o["rows"][0, :n, :7], traj), "trajectory must equal the reference's bit for bit"
    assert np.array_equal(o["percall"], d["counters"]), "per-row (nfcn,nstep,naccpt,nrejct) == scipy iwork[16:20]"
    assert o["tcur"][0] == float(d["tcur"])


def test_particle_second_call():
    d, par = H.load("g1b_second_call")
    f, p = O.make_field("EarthDipole"), O.make_params(**par)
    o1 = O.particle_advance(f, p, d["traj"][0], float(d["mass"]), float(d["charge"]), float(d["delta1"]), max_rows=2000)
    n1 = int(o1["nstored"][0])
    o2 = O.particle_advance(f, p, o1["state"][0], float(d["mass"]), float(d["charge"]), float(d["delta2"]), max_rows=2000)
    n2 = int(o2["nstored"][0])
    full = np.vstack([o1["rows"][0, :n1, :7], o2["rows"][0, 1:n2, :7]])
    assert np.array_equal(full, d["traj"])


@pytest.mark.parametrize("name", list(H.GC_CASES))
def test_gc_bit_exact(name):
    d, par = H.load(name)
    f = O.make_field(*((H.GC_CASES[name][0],) + tuple(H.GC_CASES[name][1])))
    p = O.make_params(**par)
    traj = d["traj"]; mass, q, v = float(d["mass"]), float(d["charge"]), float(d["v"])
    ppar, mu = O.gc_construct(f, traj[0, 0], d["pos"], v, float(d["pa"]), mass)
    assert mu[0] == float(d["mu"]) and ppar[0] == traj[0, 4]
    st0 = np.concatenate(([traj[0, 0]], d["pos"], ppar))
    if "bs_period" in d.files:
        Bm, vv = O.gc_mirror(f, st0, mu[0], mass)
        curve, B, ds = O.fieldline_trace(f, st0[:4], Bm)
        assert Bm == float(d["bs_Bm"]) and vv == float(d["bs_v"]) and ds == float(d["bs_ds"])
        assert np.array_equal(curve, d["bs_curve"]) and np.array_equal(B, d["bs_B"])
        bp = O.bounceperiod(f, st0, mu[0], mass)
        assert bp == float(d["bs_period"])
        dt = bp / par.get("bounceresolution", 10)
    else:
        dt = par["GCtimestep"]
    eom = str(d["eom"]) if "eom" in d.files else "TaoChanBrizardEOM"
    o = O.gc_advance(f, p, st0, mu, v, mass, q, dt, float(d["delta"]), eom=eom, max_rows=len(traj) + 10, want_percall=True)
    n = int(o["nstored"][0])
    assert n == len(traj)
    assert np.array_equal(o["rows"][0, :n, :5], traj)
    assert np.array_equal(o["percall"], d["counters"])
    assert o["tcur"][0] == float(d["tcur"])


@pytest.mark.parametrize("name", list(H.ADAPTIVE_CASES))
def test_adaptive_bit_exact(name):
    d, par = H.load(name)
    f = O.make_field(*((H.ADAPTIVE_CASES[name][0],) + tuple(H.ADAPTIVE_CASES[name][1])))
    segs, cnt = O.adaptive(f, O.make_params(**par), d["pos"], d["vel"], 0.0, float(d["mass"]), float(d["charge"]), float(d["delta"]))
    assert [s[0] for s in segs] == list(d["seg_mode"])
    assert [len(s[1]) for s in segs] == list(d["seg_nrows"])
    rows = np.vstack([s[1] for s in segs])
    assert np.array_equal(rows[:, :7], d["rows"][:, :7]), "every row of every segment, incl. the chaotic Speiser case"
    assert np.array_equal(cnt, d["counters"].sum(0))
    if par.get("GCtimestep", 0) != 0:
        nseg, rows2, seglog, cnt2 = O.adaptive_c(f, O.make_params(**par), d["pos"], d["vel"], 0.0, float(d["mass"]),
                                                 float(d["charge"]), float(d["delta"]))
        assert nseg == len(segs) and np.array_equal(rows2[:, :7], d["rows"][:, :7])


def test_notebook_known_answers():
    """Outputs stored in the reference's notebooks."""
    d, par = H.load("g3_speiser")
    out = str(d["stdout"]).splitlines()
    assert out[0] == "Switched to particle mode at time 168.0"
    assert out[1].startswith("Switched to guiding center mode at time 271.537802289")   # notebook prints 12 digits
    f = O.make_field("DoubleDipole")
    Re = O.Re
    o = O.field_ops(f, [[0, -7.8 * Re, 0, 0], [0, 10 * Re, 0, 0]])
    assert o["magB"][0] == 6.6121501357170836e-08      # BounceCenter example notebook :271
    assert o["magB"][1] == 6.1400000000000007e-08      # :291


def test_field_ops_and_utils_bit_exact():
    u = np.load(H.GOLDEN + "/units.npz")
    flds = {"earthdipole": ("EarthDipole",), "doubledipole": ("DoubleDipole",), "uniformbz": ("UniformBz", 2e-4),
            "crossedeb": ("UniformCrossedEB", 2.0, 1e-4), "vardipole": ("VarEarthDipole", 0.1, 10), "parabolic": ("Parabolic",)}
    for name, a in flds.items():
        o = O.field_ops(O.make_field(*a), u[name + "_pts"])
        for k in o:
            assert np.array_equal(o[k], u[f"{name}_{k}"], equal_nan=True), (name, k)
    m_pr, e = 1.672621777e-27, 1.602176565e-19
    o = O.utils_ops(O.make_field("DoubleDipole"), u["utils_pos"], u["utils_vel"], m_pr, e)
    for k in o:
        if k == "fp_vel":      # cos/sin/acos chain of numpy vs libm: 1-2 ulp
            assert H.relerr(o[k], u["utils_" + k]) < 1e-15
        else:
            assert np.array_equal(o[k], u["utils_" + k]), k
    for v, w in zip(u["getperp_in"], u["getperp_out"]):
        assert np.array_equal(O.getperp(v), w)


def test_ensembles_vs_reference():
    from rapt_b200 import synth
    d, par = H.load("e2_config2_first32")
    n = int(d["n"]); ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], O.particle_momentum(vel, ic["mass"])])
    o = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**par), st, ic["mass"], ic["charge"], float(d["delta"]),
                           store_every=0, nthreads=4)
    assert np.array_equal(o["state"], d["final"]) and np.array_equal(o["counters"], d["totals"])
    assert np.array_equal(o["nrows"], d["nrows"]) and np.array_equal(o["tcur"], d["tcur"])
    for case, fname, fargs, gen in (("e3_config3_first16", "DoubleDipole", (), synth.config3_electrons),
                                    ("e5_config5_first16", "VarEarthDipole", (0.1, 10), synth.config5_belt)):
        d, par = H.load(case)
        n = int(d["n"]); ic = gen(n); f = O.make_field(fname, *fargs)
        pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
        ppar, mu = O.gc_construct(f, ic["t0"], pos, ic["v"], ic["pa"], ic["mass"])
        assert np.array_equal(mu, d["mu"])
        st = np.column_stack([ic["t0"], pos, ppar])
        o = O.gc_advance(f, O.make_params(**par), st, mu, ic["v"], ic["mass"], ic["charge"], par["GCtimestep"], float(d["delta"]),
                         store_every=0, nthreads=4)
        assert np.array_equal(o["state"], d["final"]) and np.array_equal(o["counters"], d["totals"])
        if "bounceperiod" in d.files:
            bp = np.array([O.bounceperiod(f, st[i], mu[i], ic["mass"][i]) for i in range(4)])
            assert np.array_equal(bp, d["bounceperiod"][:4])


def test_solvers_vs_scipy_dop():
    """dop853 / dopri5 restatements against scipy.integrate.ode on test ODEs that force rejections:
    identical counters (nfcn, nstep, naccpt, nrejct) and results to round-off."""
    from scipy.integrate import ode

    def f0(t, y):
        return np.array([y[1], -y[0] * (1 + 5000 * np.exp(-((t - 1.5) / 0.02) ** 2) + 3000 * np.exp(-((t - 2.2) / 0.01) ** 2))])

    def f1(t, y):
        return np.array([y[1], -y[0] * (1 + 50 * np.sin(3 * t) ** 2)])
    for which, fn, xend in ((0, f0, 3.0), (1, f1, 4.0)):
        for solver, name, kw in ((853, "dop853", dict(beta=0.1)), (5, "dopri5", {})):
            r = ode(fn).set_integrator(name, rtol=1e-6, atol=1e-9, **kw)
            r.set_initial_value([1.0, 0.0], 0.0)
            y_ref = r.integrate(xend)
            cnt_ref = np.array(r._integrator.iwork[16:20], dtype=np.int64)
            idid, y, cnt = O.test_solver(solver, which, 0.0, xend, [1.0, 0.0], 1e-6, 1e-9)
            assert idid == 1
            assert np.array_equal(cnt, cnt_ref), (name, which, cnt, cnt_ref)
            assert np.max(np.abs(y - y_ref)) < 1e-6 * np.max(np.abs(y_ref))
    # Hairer's rejection rule differs from scipy 1.18.1's on this problem
    _, _, cnt_h = O.test_solver(853, 0, 0.0, 3.0, [1.0, 0.0], 1e-6, 1e-9, reject_rule=1)
    _, _, cnt_s = O.test_solver(853, 0, 0.0, 3.0, [1.0, 0.0], 1e-6, 1e-9, reject_rule=0)
    assert tuple(cnt_h) != tuple(cnt_s)


def test_edge_cases():
    f, p = O.make_field("EarthDipole"), O.make_params(cyclotronresolution=20)
    d, _ = H.load("g1b_generic")
    # zero duration: no rows
    o = O.particle_advance(f, p, d["traj"][0], float(d["mass"]), float(d["charge"]), 0.0, max_rows=4)
    assert o["nrows"][0] == 1 and np.array_equal(o["state"][0], d["traj"][0])
    # decimation: stored rows are every k-th row of the full trajectory
    full = O.particle_advance(f, p, d["traj"][0], float(d["mass"]), float(d["charge"]), 1.0, max_rows=400)
    dec = O.particle_advance(f, p, d["traj"][0], float(d["mass"]), float(d["charge"]), 1.0, store_every=5, max_rows=400)
    nf, nd = int(full["nstored"][0]), int(dec["nstored"][0])
    assert np.array_equal(dec["rows"][0, :nd, :7], full["rows"][0, :nf:5, :7])
    assert np.array_equal(dec["state"], full["state"])
    # row buffer too small: integration continues, only storage stops
    small = O.particle_advance(f, p, d["traj"][0], float(d["mass"]), float(d["charge"]), 1.0, max_rows=10)
    assert small["nstored"][0] == 10 and small["nrows"][0] == full["nrows"][0]
    assert np.array_equal(small["state"], full["state"])
    # nsteps = 500 per row: an absurd output step makes the solver fail like scipy (-2) and ends the loop
    pp = O.make_params(cyclotronresolution=1e-4)
    o = O.particle_advance(f, pp, d["traj"][0], float(d["mass"]), float(d["charge"]), 1e6, max_rows=4)
    assert o["status"][0] == -2 and o["nrows"][0] == 2


@pytest.mark.parametrize("name", ["eye_pa80", "eye_pa45_simpson"])
def test_oracle_full_fieldline_traces_of_geteye(name):
    """Fieldline(row[:4], field, Bmax=Bm).trace() as GuidingCenter.geteye calls it (GuidingCenter.py:620-622):
    the oracle's RKF45 trace reproduces the reference's s and |B| along every traced line bit for bit."""
    d, _ = H.load(name)
    f = O.make_field("EarthDipole")
    rows = d["traj"][::int(d["step"])]
    for i, row in enumerate(rows):
        curve, B, ds = O.fieldline_trace(f, row[:4], float(d["Bm"][i]))
        k = int(d["npts"][i])
        assert len(curve) == k
        assert np.array_equal(curve[:, 0], d["curves"][i, :k, 0]) and np.array_equal(B, d["curves"][i, :k, 1])


def test_grid_field_bit_exact():
    """fields.Grid (fields.py:513-814; scipy RegularGridInterpolator, linear, 4-D) on synthetic data files:
    field operators, a Particle and a GuidingCenter trajectory that cross the rolling window's update time, and
    the ValueError of a tracer leaving the grid -- all against the unmodified reference."""
    d, _ = H.load("grid_synthetic")
    G = H.synthetic_grid([str(s) for s in d["files"]])
    assert G["sha256"] == str(d["checksum"]), "synthetic grid must regenerate bit for bit"
    f = O.make_grid_field(G["t"], G["x"], G["y"], G["z"], G["B"], G["E"])
    assert f.gradstep == float(d["ops_gradstep"]) and bool(f.is_static) == bool(d["ops_static"])
    ops = O.field_ops(f, d["ops_pts"])
    for k in ("B", "E", "magB", "unitb", "gradB", "curlb", "lengthscale"):
        assert np.array_equal(ops[k], d["ops_" + k]), k
    import json
    par = json.loads(str(d["p_params"]))
    st0 = d["p_traj"][0]
    o = O.particle_advance(f, O.make_params(**par), st0, float(d["p_mass"]), float(d["p_charge"]), float(d["p_delta"]),
                           max_rows=len(d["p_traj"]) + 10, want_percall=True)
    n = int(o["nstored"][0])
    assert n == len(d["p_traj"]) and np.array_equal(o["rows"][0, :n, :7], d["p_traj"])
    assert np.array_equal(o["percall"], d["p_counters"]) and o["tcur"][0] == float(d["p_tcur"])
    par = json.loads(str(d["g_params"]))
    traj = d["g_traj"]; mass, q, v = float(d["g_mass"]), float(d["g_charge"]), float(d["g_v"])
    ppar, mu = O.gc_construct(f, traj[0, 0], d["g_pos"], v, float(d["g_pa"]), mass)
    assert mu[0] == float(d["g_mu"]) and ppar[0] == traj[0, 4]
    st0 = np.concatenate(([traj[0, 0]], d["g_pos"], ppar))
    o = O.gc_advance(f, O.make_params(**par), st0, mu, v, mass, q, par["GCtimestep"], float(d["g_delta"]),
                     max_rows=len(traj) + 10, want_percall=True)
    n = int(o["nstored"][0])
    assert n == len(traj) and np.array_equal(o["rows"][0, :n, :5], traj)
    assert np.array_equal(o["percall"], d["g_counters"]) and o["tcur"][0] == float(d["g_tcur"])
    # leaving the grid: the reference raises ValueError and keeps the rows appended so far
    assert bool(d["oob_raised"])
    from rapt_b200 import m_pr
    gm = np.sqrt(m_pr ** 2 / (1 - np.dot(d["oob_vel"], d["oob_vel"]) / 299792458.0 ** 2))
    st0 = np.concatenate(([0.0], d["oob_pos"], gm * d["oob_vel"]))
    o = O.particle_advance(f, O.make_params(cyclotronresolution=10), st0, m_pr, float(d["p_charge"]), 5.0, max_rows=64)
    assert int(o["status"][0]) == -6 and int(o["nstored"][0]) == len(d["oob_traj"])


# ---------------------------------------------------------------- BounceCenter
BC_CASES = ("bc_dipole_electron", "bc_dipole_proton", "bc_doubledipole_electron")


@pytest.mark.parametrize("name", BC_CASES)
def test_bounce_center_pieces_bit_exact(name):
    """flutils.halfbouncepath / eye / gradI at three rows of the reference's BounceCenter trajectory."""
    d = np.load(os.path.join(H.GOLDEN, name + ".npz"))
    f = O.make_field(str(d["field"]))
    Bm = float(d["Bm"])
    for r, Sb, I, gI in zip(d["pts"], d["Sb"], d["I"], d["gradI"]):
        assert O.halfbouncepath(f, r, Bm) == Sb
        assert O.eye(f, r, Bm) == I
        assert np.array_equal(O.gradI(f, r, Bm), gI)


def test_flutils_special_branches_bit_exact():
    """gradI's one-sided differences (a displaced line beyond the mirror field) and eye's Simpson branch."""
    d = np.load(os.path.join(H.GOLDEN, "bc_flutils.npz"))
    f = O.make_field("EarthDipole")
    for tp, Bm, Sb, I, gI in zip(d["tpos"], d["Bm"], d["Sb"], d["I"], d["gradI"]):
        assert O.halfbouncepath(f, tp, Bm) == Sb
        assert O.eye(f, tp, Bm) == I
        assert np.array_equal(O.gradI(f, tp, Bm), gI)


@pytest.mark.parametrize("name", ["bc_dipole_proton", "bc_doubledipole_electron"])
def test_bounce_center_advance_bit_exact(name):
    """BounceCenter.advance: every row and the dopri5 counters of every call (the electron/dipole fixture, 9 s of
    scipy quadrature, is left to the GPU test, which checks the device against it directly)."""
    d = np.load(os.path.join(H.GOLDEN, name + ".npz"))
    f = O.make_field(str(d["field"]))
    n1 = int(d["nrows_first_call"])
    rows, cnt, dt = O.bounce_center_advance(f, d["traj"][0], float(d["mu"]), float(d["v"]), float(d["mass"]),
                                            float(d["charge"]), float(d["delta"]))
    assert np.array_equal(rows, d["traj"][1:n1])
    assert np.array_equal(cnt, d["solver_log"][:n1 - 1].sum(0))
    assert rows[0, 0] == d["traj"][0, 0]            # the first computed row carries the START label (quirk kept)


def test_solver_failure_row_bit_exact():
    """scipy's nsteps = 500 limit inside advance(): the failed call's row is appended (Particle.py:304-307 label = row end
    time, GuidingCenter.py:452-456 label = time reached) and the loop ends -- fixtures p_fail_nmax / gc_fail_nmax."""
    d, par = H.load("p_fail_nmax")
    traj = d["traj"]
    o = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**par), traj[0], float(d["mass"]), float(d["charge"]),
                           float(d["delta"]), max_rows=8, want_percall=True)
    assert o["status"][0] == -2 and o["nrows"][0] == o["nstored"][0] == 2
    assert np.array_equal(o["rows"][0, :2, :7], traj) and np.array_equal(o["percall"], d["counters"])
    assert o["tcur"][0] == float(d["tcur"])
    d, par = H.load("gc_fail_nmax")
    traj = d["traj"]
    f = O.make_field("EarthDipole")
    ppar, mu = O.gc_construct(f, 0.0, d["pos"], float(d["v"]), float(d["pa"]), float(d["mass"]))
    o = O.gc_advance(f, O.make_params(), np.concatenate(([0.0], d["pos"], ppar)), mu, float(d["v"]), float(d["mass"]),
                     float(d["charge"]), par["GCtimestep"], float(d["delta"]), max_rows=8, want_percall=True)
    assert o["status"][0] == -2 and o["nrows"][0] == o["nstored"][0] == 2
    assert np.array_equal(o["rows"][0, :2, :5], traj) and np.array_equal(o["percall"], d["counters"])
    assert o["tcur"][0] == float(d["tcur"])
