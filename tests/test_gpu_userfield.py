"""GPU parity: user-defined analytic field compiled with NVRTC (fields plugin interface)."""
import numpy as np
import pytest

import helpers as H
from userfield import make_charged_dipole

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rapt_b200
    from rapt_b200 import _lib
    _lib.init(0)
    return rapt_b200


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_user_field_particle_vs_reference(rb, arith):
    """Notebook 'Creating new fields', cells 10-19: proton in ChargedDipole(Q=1e-6), advance(4e-4)."""
    d, par = H.load("p_chargeddipole")
    f = make_charged_dipole()(Q=1e-6)
    traj = d["traj"]
    o = rb.engine.particle_advance(f, traj[0], float(d["mass"]), float(d["charge"]), float(d["delta"]), store_every=1,
                                   max_rows=len(traj) + 8, arith=arith)
    n = int(o["nstored"][0])
    assert o["status"][0] == 1 and n == len(traj)
    rows = o["rows"][0, :n]
    assert H.vec_relerr(rows[:, 1:4], traj[:, 1:4]) < 1e-8
    assert H.vec_relerr(rows[:, 4:7], traj[:, 4:7]) < 1e-8
    assert tuple(o["counters"][0]) == tuple(d["counters"].sum(0))
    # host-side API surface of the same plugin object
    p = rb.Particle(pos=d["pos"], vel=d["vel"], t0=0, mass=rb.m_pr, charge=rb.e, field=f)
    assert p.cycper() == pytest.approx(float(d["cycper"]), rel=1e-12)
    assert p.cycrad() == pytest.approx(float(d["cycrad"]), rel=1e-9)


def test_user_field_ops_and_gc(rb):
    """The NVRTC module also provides the field operators and the guiding-centre kernel."""
    import oracle as O
    f = make_charged_dipole()(B0=2.0, Q=1e-7)
    rng = np.random.default_rng(3)
    pts = np.column_stack([np.zeros(16), rng.uniform(2, 6, 16), rng.uniform(-3, 3, 16), rng.uniform(-2, 2, 16)])
    o = rb.engine.field_ops(f, pts, arith="strict")
    ref = O.field_ops(O.make_field("ChargedDipole", 2.0, 1e-7), pts)
    for k in ("B", "E", "magB", "unitb"):
        assert H.relerr(o[k], ref[k], floor=1e-300) < 1e-13, k
    for k in ("gradB", "curlb"):
        assert np.max(np.abs(o[k] - ref[k])) < 1e-6 * np.max(np.abs(ref[k])), k
    # a snippet with a syntax error is reported, not swallowed
    from rapt_b200 import fields, _lib

    class Broken(fields._Field):
        cuda_source = "__device__ void rapt_user_B(double t, double x, double y, double z, const double* prm, double* B) { B[0] = nonsense; }"
    with pytest.raises(_lib.RaptB200Error, match="nonsense"):
        rb.engine.field_ops(Broken(), pts)


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_user_field_guiding_centre_and_switches(rb, arith):
    """The NVRTC module carries every kernel family: guiding-centre advance, constructor, switch transforms
    and predicates with a user field, against the CPU oracle's restatement of the same field."""
    import oracle as O
    f = make_charged_dipole()(B0=3.0e-5 * 6378137.0 ** 3, Q=0.0)       # dipole moment of Earth's size, no charge
    f.static = True
    f.gradientstepsize = 6378137.0 / 1000
    of = O.make_field("ChargedDipole", 3.0e-5 * 6378137.0 ** 3, 0.0, gradstep=6378137.0 / 1000, static=True)
    Re = 6378137.0
    n = 64
    rng = np.random.default_rng(5)
    pos = np.column_stack([rng.uniform(3, 6, n) * Re, rng.uniform(-1, 1, n) * Re, rng.uniform(-0.3, 0.3, n) * Re])
    v = np.full(n, 1.5e8); pa = rng.uniform(35, 85, n); mass = np.full(n, rb.m_el); q = np.full(n, -rb.e)
    ppar, mu = rb.engine.gc_construct(f, 0.0, pos, v, pa, mass, arith=arith)
    ppar_o, mu_o = O.gc_construct(of, 0.0, pos, v, pa, mass)
    assert H.relerr(mu, mu_o) < 1e-12
    st = np.column_stack([np.zeros(n), pos, ppar_o])
    got = rb.engine.gc_advance(f, st, mu_o, v, mass, q, 0.05, 1.0, store_every=0, arith=arith)
    ref = O.gc_advance(of, O.make_params(), st, mu_o, v, mass, q, 0.05, 1.0, store_every=0, nthreads=4)
    assert np.array_equal(got["nrows"], ref["nrows"])
    assert H.vec_relerr(got["state"][:, 1:4], ref["state"][:, 1:4]) < 1e-8
    assert (got["counters"][:, 1] == ref["counters"][:, 1]).mean() > 0.9
    # particle <-> guiding centre with the user field
    vel = rng.normal(size=(n, 3)); vel *= (1.0e8 / np.linalg.norm(vel, axis=1))[:, None]
    prow = np.column_stack([np.zeros(n), pos, rb.engine.particle_momentum(vel, mass)])
    grow, mu2, v2, st2 = rb.engine.switch_p2g(f, prow, mass, q, arith=arith)
    for i in range(0, n, 7):
        rc, g_o, mu_r, v_r = O.switch_P2G(of, prow[i], mass[i], q[i])
        assert rc == st2[i]
        if rc == 0:
            assert H.vec_relerr(grow[i, 1:4], g_o[1:4]) < 1e-10 and abs(mu2[i] / mu_r - 1) < 1e-8
    ad = rb.engine.isadiabatic(f, 0, prow, 0.0, mass, q, arith=arith)
    ad_o = [O.particle_isadiabatic(of, O.make_params(), r, mass[0], q[0]) for r in prow]
    assert list(ad) == ad_o


def test_guiding_centre_ensemble_default_output_step(rb):
    """GuidingCenterEnsemble with params['GCtimestep'] == 0: dt = bounceperiod()/bounceresolution
    (GuidingCenter.py:443-446); above HOST_QUADRATURE_MAX members the bounce periods come from the
    device closed-form quadrature."""
    from rapt_b200 import synth
    n = 6000
    ic = synth.config3_electrons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]])
    g = rb.GuidingCenterEnsemble(pos, ic["v"], pa=ic["pa"], mass=ic["mass"], charge=ic["charge"], field=rb.fields.DoubleDipole())
    assert rb.params["GCtimestep"] == 0
    bp = g.bounceperiod()
    assert np.isfinite(bp).all() and bp.min() > 0.1 and bp.max() < 10
    sub = rb.GuidingCenterEnsemble(pos[:32], ic["v"][:32], pa=ic["pa"][:32], mass=ic["mass"][:32], charge=ic["charge"][:32],
                                   field=rb.fields.DoubleDipole())
    import scipy_legs
    ref = scipy_legs.bounceperiod(sub.field, sub.state, sub.mu, sub.mass, rb.params["fieldlineresolution"])
    assert np.max(np.abs(bp[:32] / ref - 1)) < 1e-4
    g.advance(1.0)
    assert np.all(g.status == 1)
    assert np.all(g.nrows == 1 + np.ceil(1.0 / (bp / rb.params["bounceresolution"]) - 1e-9))


def test_user_field_adaptive_epochs(rb):
    """Adaptive epochs (advance kernels + switch/compaction kernel) from the NVRTC module of a user field."""
    import oracle as O
    Re = 6378137.0
    M = 3.0e-5 * Re ** 3
    f = make_charged_dipole()(B0=M, Q=0.0); f.static = True; f.gradientstepsize = Re / 1000
    of = O.make_field("ChargedDipole", M, 0.0, gradstep=Re / 1000, static=True)
    n = 12
    rng = np.random.default_rng(11)
    pos = np.column_stack([rng.uniform(2.5, 4.5, n) * Re, rng.uniform(-0.5, 0.5, n) * Re, rng.uniform(-0.2, 0.2, n) * Re])
    spd = rb.utils.speedfromKE(3e4, rb.m_pr)
    d = rng.normal(size=(n, 3)); vel = d / np.linalg.norm(d, axis=1)[:, None] * spd
    par = dict(epss=0.05)
    o = rb.engine.adaptive_advance(f, pos, vel, 0.0, rb.m_pr, rb.e, 3.0, 0.25, store_every=1, max_rows=4096, arith="strict", **par)
    op = O.make_params(GCtimestep=0.25, **par)
    for i in range(n):
        nseg, rows, seglog, cnt = O.adaptive_c(of, op, pos[i], vel[i], 0.0, rb.m_pr, rb.e, 3.0)
        if nseg < 0:
            assert o["status"][i] == nseg
            continue
        assert o["status"][i] == 1 and o["nseg"][i] == nseg
        mine = o["rows"][i, :o["nstored"][i]]
        assert len(mine) == len(rows)
        assert np.max(np.abs(mine[:, 0] - rows[:, 0])) < 1e-9
        assert H.vec_relerr(mine[:, 1:4], rows[:, 1:4]) < 1e-7
