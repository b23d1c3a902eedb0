"""GPU parity: Particle.advance on the B200 (through the C ABI) vs the golden vectors generated from
the reference and vs the CPU oracle on seeded ensembles.

Bars (BASELINE.json north_star): final position/momentum <= 1e-8 relative; step-accept counts equal
on non-chaotic cases.  Tolerances are written at each assert.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from rapt_b200 import engine, _lib
    _lib.init(0)
    return engine


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("name", list(H.PARTICLE_CASES))
def test_particle_golden(eng, name, arith):
    d, par = H.load(name)
    fname, fargs = H.PARTICLE_CASES[name]
    traj = d["traj"]
    o = eng.particle_advance(H.gpu_field(fname, fargs), traj[0], float(d["mass"]), float(d["charge"]), float(d["delta"]),
                             store_every=1, max_rows=len(traj) + 8, params=None, arith=arith,
                             **{k: v for k, v in par.items()})
    n = int(o["nstored"][0])
    assert o["status"][0] == 1
    assert o["nrows"][0] == len(traj) == n, "row count (1 + ceil(delta/dt)) must match the reference"
    rows = o["rows"][0, :n]
    # time labels are pure additions of dt: bit-exact or 1 ulp (dt itself carries the field evaluation)
    assert H.relerr(rows[:, 0], traj[:, 0]) < 1e-13
    chaotic = name in ("p_parabolic",)      # current-sheet crossings amplify round-off
    tol = 1e-8 if not chaotic else 1e-6
    # whole trajectory: position and momentum vectors, relative to their norms
    assert H.vec_relerr(rows[:, 1:4], traj[:, 1:4]) < tol
    assert H.vec_relerr(rows[:, 4:7], traj[:, 4:7]) < (tol if not chaotic else 1e-5)
    ref = d["counters"].sum(0)
    if name == "g1_readme" and arith == "fast":
        # zero-coordinate start: first row is round-off dominated; allow a few steps
        assert abs(int(o["counters"][0, 1]) - int(ref[1])) <= 4
    elif not chaotic:
        assert tuple(o["counters"][0]) == tuple(ref), "(nfcn, nstep, naccpt, nrejct) must equal scipy's"
        # cumulative step count stored with each row == cumulative sum of scipy's per-call nstep
        assert np.array_equal(rows[1:, 7].astype(np.int64), np.cumsum(d["counters"][:, 1]))
    assert abs(o["tcur"][0] - float(d["tcur"])) <= 1e-12 * abs(float(d["tcur"]))


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_particle_second_call(eng, arith):
    """advance() twice: dt is recomputed from the current state (Particle.py:282)."""
    d, par = H.load("g1b_second_call")
    traj = d["traj"]
    f = H.gpu_field("EarthDipole", ())
    o1 = eng.particle_advance(f, traj[0], float(d["mass"]), float(d["charge"]), float(d["delta1"]), max_rows=1000,
                              arith=arith, **par)
    n1 = int(o1["nstored"][0])
    o2 = eng.particle_advance(f, o1["state"][0], float(d["mass"]), float(d["charge"]), float(d["delta2"]), max_rows=1000,
                              arith=arith, **par)
    n2 = int(o2["nstored"][0])
    assert n1 + n2 - 1 == len(traj)
    full = np.vstack([o1["rows"][0, :n1, :7], o2["rows"][0, 1:n2, :7]])
    assert H.vec_relerr(full[:, 1:4], traj[:, 1:4]) < 1e-8
    assert H.vec_relerr(full[:, 4:7], traj[:, 4:7]) < 1e-8
    assert o1["dt"][0] != o2["dt"][0]


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_config2_first32_vs_reference(eng, arith):
    """First 32 protons of config 2 (seed 20260201), advance(1.0): golden from the reference."""
    from rapt_b200 import synth
    d, par = H.load("e2_config2_first32")
    n = int(d["n"])
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    mom = eng.particle_momentum(vel, ic["mass"])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], mom])
    o = eng.particle_advance(H.gpu_field("EarthDipole", ()), st, ic["mass"], ic["charge"], float(d["delta"]),
                             store_every=0, arith=arith, **par)
    fin = d["final"]
    assert np.array_equal(o["nrows"], d["nrows"])
    assert H.relerr(o["state"][:, 0], fin[:, 0]) < 1e-13
    assert H.vec_relerr(o["state"][:, 1:4], fin[:, 1:4]) < 1e-8
    assert H.vec_relerr(o["state"][:, 4:7], fin[:, 4:7]) < 1e-8
    assert np.array_equal(o["counters"], d["totals"]), "per-particle (nfcn,nstep,naccpt,nrejct) equal scipy's"


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_config2_first32_bench_horizon_vs_reference(eng, arith):
    """First 32 protons of config 2 at the BENCH horizon, advance(10 s): the headline workload itself against the
    reference (fixture e2_config2_first32_10s; 98 k attempted steps).  Bars: final state 1e-8, row counts equal,
    (nfcn, nstep, naccpt, nrejct) equal for every proton in the strict flavour; the fast flavour reports how many
    protons differ (an accept/reject decision whose err is within the flavour's round-off of 1.0)."""
    from rapt_b200 import synth
    d, par = H.load("e2_config2_first32_10s")
    n = int(d["n"])
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], eng.particle_momentum(vel, ic["mass"])])
    o = eng.particle_advance(H.gpu_field("EarthDipole", ()), st, ic["mass"], ic["charge"], float(d["delta"]),
                             store_every=0, arith=arith, **par)
    fin = d["final"]
    assert np.all(o["status"] == 1)
    assert np.array_equal(o["nrows"], d["nrows"])
    assert H.relerr(o["state"][:, 0], fin[:, 0]) < 1e-12
    assert H.vec_relerr(o["state"][:, 1:4], fin[:, 1:4]) < 1e-8
    assert H.vec_relerr(o["state"][:, 4:7], fin[:, 4:7]) < 1e-8
    same = np.all(o["counters"] == d["totals"], axis=1)
    print(f"[{arith}] config 2 x 10 s: {int(same.sum())}/{n} protons with scipy's exact counters; "
          f"nstep sum {int(o['counters'][:, 1].sum())} vs {int(d['totals'][:, 1].sum())}")
    if arith == "strict":
        assert same.all(), "per-proton (nfcn,nstep,naccpt,nrejct) equal scipy's over the whole 10 s"
    else:
        assert same.sum() >= n - 2
        assert abs(int(o["counters"][:, 1].sum()) - int(d["totals"][:, 1].sum())) <= 4
    assert np.allclose(o["tcur"], d["tcur"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_particle_325_gyroperiods_vs_reference(eng, arith):
    """north_star's short horizon, 'a few hundred gyroperiods': the g1b proton for 40 s = 324 gyroperiods, 6486 rows
    (fixture g1c_325_gyroperiods holds every 16th row, the last row and all per-call counters)."""
    d, par = H.load("g1c_325_gyroperiods")
    assert 300 < float(d["gyroperiods"]) < 350
    every = int(d["every"]); nrows = int(d["nrows"])
    st0 = d["traj_dec"][0]
    o = eng.particle_advance(H.gpu_field("EarthDipole", ()), st0, float(d["mass"]), float(d["charge"]), float(d["delta"]),
                             store_every=every, max_rows=len(d["traj_dec"]) + 8, arith=arith, **par)
    assert o["status"][0] == 1 and o["nrows"][0] == nrows
    k = int(o["nstored"][0])
    assert k == len(d["traj_dec"])
    rows = o["rows"][0, :k]
    assert H.relerr(rows[:, 0], d["traj_dec"][:, 0], floor=1e-3) < 1e-12
    assert H.vec_relerr(rows[:, 1:4], d["traj_dec"][:, 1:4]) < 1e-8
    assert H.vec_relerr(rows[:, 4:7], d["traj_dec"][:, 4:7]) < 1e-8
    assert H.vec_relerr(o["state"][0, 1:4], d["last"][1:4]) < 1e-8 and H.vec_relerr(o["state"][0, 4:7], d["last"][4:7]) < 1e-8
    ref = d["counters"].astype(np.int64)
    assert tuple(o["counters"][0]) == tuple(ref.sum(0)), "6499 attempted steps, all accepted, as scipy counts them"
    # cumulative attempted steps stored with each decimated row == scipy's running total at that row
    assert np.array_equal(rows[1:, 7].astype(np.int64), np.cumsum(ref[:, 1])[every - 1::every][:k - 1])
    # |p| (energy in a static B) drifts exactly as much as the reference's own integration lets it
    p0 = np.linalg.norm(st0[4:7]); dr = np.linalg.norm(d["last"][4:7]) / p0 - 1; dg = np.linalg.norm(o["state"][0, 4:7]) / p0 - 1
    assert abs(dg - dr) < 1e-10 and abs(dg) < 1e-5


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_config2_ensemble_vs_oracle(eng, arith):
    """4096 protons of config 2, advance(0.25 s): CUDA vs the CPU oracle on identical inputs, with
    decimated trajectory storage (store_every = 7)."""
    import oracle as O
    from rapt_b200 import synth
    n = 4096
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    mom = eng.particle_momentum(vel, ic["mass"])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], mom])
    par = dict(cyclotronresolution=20)
    ref = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**par), st, ic["mass"], ic["charge"], 0.25,
                             store_every=7, max_rows=64, nthreads=8)
    o = eng.particle_advance(H.gpu_field("EarthDipole", ()), st, ic["mass"], ic["charge"], 0.25,
                             store_every=7, max_rows=64, arith=arith, **par)
    assert np.array_equal(o["nrows"], ref["nrows"])
    assert np.array_equal(o["nstored"], ref["nstored"])
    assert np.all(o["status"] == 1)
    assert H.vec_relerr(o["state"][:, 1:4], ref["state"][:, 1:4]) < 1e-8
    assert H.vec_relerr(o["state"][:, 4:7], ref["state"][:, 4:7]) < 1e-8
    # attempted/accepted step counts: exact for (almost) every particle; report the fraction
    same = np.all(o["counters"] == ref["counters"], axis=1)
    assert same.mean() > (0.999 if arith == "strict" else 0.99), f"only {same.mean():.4f} of particles match counts"
    assert abs(int(o["counters"][:, 1].sum()) - int(ref["counters"][:, 1].sum())) <= 1e-4 * ref["counters"][:, 1].sum()
    # stored (decimated) rows
    for i in (0, 17, 4095):
        k = int(o["nstored"][i])
        assert H.vec_relerr(o["rows"][i, :k, 1:4], ref["rows"][i, :k, 1:4]) < 1e-8
        assert np.array_equal(o["rows"][i, :k, 0], ref["rows"][i, :k, 0]) or H.relerr(o["rows"][i, :k, 0], ref["rows"][i, :k, 0]) < 1e-13
    # physics: |p| is conserved in a static magnetic field -- to the same level as the reference
    # integrator conserves it (momentum is not error-controlled: atol is in SI units)
    p0 = np.linalg.norm(st[:, 4:7], axis=1); p1 = np.linalg.norm(o["state"][:, 4:7], axis=1)
    pr = np.linalg.norm(ref["state"][:, 4:7], axis=1)
    assert np.max(np.abs(p1 / p0 - 1)) < 1e-4
    assert np.max(np.abs(p1 / pr - 1)) < 1e-9


def test_empty_and_zero_delta(eng):
    f = H.gpu_field("EarthDipole", ())
    o = eng.particle_advance(f, np.zeros((0, 7)), np.zeros(0), np.zeros(0), 1.0)
    assert o["state"].shape == (0, 7)
    d, par = H.load("g1b_generic")
    o = eng.particle_advance(f, d["traj"][0], float(d["mass"]), float(d["charge"]), 0.0, max_rows=4, **par)
    assert o["nrows"][0] == 1 and o["nstored"][0] == 1
    assert np.array_equal(o["state"][0], d["traj"][0])


def test_fp64_peak(eng):
    tf, mhz = eng.fp64_peak()
    assert 5.0 < tf < 80.0, f"implausible FP64 peak {tf} TFLOP/s"


@pytest.mark.timeout(120)
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_solver_failure_row_and_no_hang(eng, arith):
    """Particle.advance / GuidingCenter.advance whose solver hits scipy's nsteps = 500: the reference warns, appends the
    failed call's row and returns (Particle.py:304-307, GuidingCenter.py:452-456; fixtures p_fail_nmax / gc_fail_nmax).
    The object-level row loop must end (it used to rerun the launch for ever)."""
    import warnings
    import rapt_b200 as R
    old = dict(R.params)
    try:
        d, par = H.load("p_fail_nmax")
        R.params.update(par); R.params["arith"] = arith
        p = R.Particle(pos=tuple(d["pos"]), vel=tuple(d["vel"]), t0=0, mass=float(d["mass"]), charge=float(d["charge"]),
                       field=R.fields.EarthDipole())
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            p.advance(float(d["delta"]))
        assert any("nsteps" in str(x.message) for x in w)
        traj = d["traj"]
        assert p.trajectory.shape == traj.shape == (2, 7)
        assert abs(p.trajectory[1, 0] / traj[1, 0] - 1) < 1e-13
        # rtol = 1e-15 is below the round-off of the error estimate itself: how far the 501 attempts get is not
        # reproducible across operation orders (1e-14 m in the reference, up to 1e-3 m here); the row, its label, the
        # counters and tcur are
        assert H.vec_relerr(p.trajectory[1, 1:4], traj[1, 1:4]) < 1e-8 and H.vec_relerr(p.trajectory[1, 4:7], traj[1, 4:7]) < 1e-8
        assert tuple(p.solver_counters) == tuple(d["counters"].sum(0))
        assert abs(p.tcur - float(d["tcur"])) < 1e-8 * float(d["tcur"])
        R.params.clear(); R.params.update(old)
        d, par = H.load("gc_fail_nmax")
        R.params.update(par); R.params["arith"] = arith
        g = R.GuidingCenter(pos=tuple(d["pos"]), v=float(d["v"]), pa=float(d["pa"]), mass=float(d["mass"]),
                            charge=float(d["charge"]), field=R.fields.EarthDipole())
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            g.advance(float(d["delta"]))
        assert any("nsteps" in str(x.message) for x in w)
        assert g.trajectory.shape == d["traj"].shape == (2, 5)
        assert int(g.solver_counters[1]) == 501
        # 501 attempts of a bounce motion asked for in 50 s output steps: where the budget runs out depends on every
        # accept/reject decision (18 of the reference's 501 attempts are rejected); on the device even the strict flavour
        # differs from the reference through CUDA's pow (0.97 % in the time reached), the host build of the same source is
        # bit-identical (tests/test_kernel_host.py::test_kernel_source_solver_failure_row)
        assert abs(g.trajectory[1, 0] / d["traj"][1, 0] - 1) < 0.05 and g.tcur == g.trajectory[1, 0]
    finally:
        R.params.clear(); R.params.update(old)


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_config2_solver_failure_member(eng, arith):
    """bench.py's `solver_failures: 1` on the headline run is member 408359 of the 1 M-proton ensemble.  The reference on
    that proton (fixture e2_config2_member408359): scipy warns 'dop853: larger nsteps is needed' in the call of row 369,
    appends that call's row and stops at 370 rows -- the kernel must end the same way."""
    from rapt_b200 import synth
    d, par = H.load("e2_config2_member408359")
    i = int(d["member"])
    ic = synth.config2_protons(int(d["n_total"]))
    assert np.array_equal([ic["x"][i], ic["y"][i], ic["z"][i]], d["pos"]) and bool(d["warned"])
    vel = np.array([[ic["vx"][i], ic["vy"][i], ic["vz"][i]]])
    st = np.concatenate(([0.0], d["pos"], eng.particle_momentum(vel, ic["mass"][i:i + 1])[0]))
    nrows = int(d["nrows"])
    o = eng.particle_advance(H.gpu_field("EarthDipole", ()), st, float(d["mass"]), float(d["charge"]), float(d["delta"]),
                             store_every=1, max_rows=nrows + 8, arith=arith, **par)
    assert o["status"][0] == -2, "the row loop ends on nsteps = 500, as the reference's does"
    assert o["nrows"][0] == o["nstored"][0] == nrows
    rows = o["rows"][0, nrows - 3:nrows]
    assert H.relerr(rows[:, 0], d["last_rows"][:, 0]) < 1e-12
    # this proton mirrors at 0.83 Re, inside the planet, where the dipole field and its gradient are so large that the step
    # size collapses: the last rows before the failure are already ill-conditioned (7.7e-7 between arithmetic flavours on
    # the host build, bit-identical there in the strict one); bar 1e-5
    assert np.linalg.norm(d["last_rows"][-1, 1:4]) < 0.85 * 6378137.0
    assert H.vec_relerr(rows[:, 1:4], d["last_rows"][:, 1:4]) < 1e-5
    assert abs(int(o["counters"][0, 1]) - int(d["totals"][1])) <= 0.01 * int(d["totals"][1])
