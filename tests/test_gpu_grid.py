"""GPU parity: fields.Grid (rapt/fields.py:513-814) -- gridded E/B, multilinear interpolation in (t, x, y, z)
on device-resident tables -- against the reference's own values on synthetic data files
(tests/golden/grid_synthetic.npz, oracle/gen_golden.py:case_grid) and against the CPU oracle on ensembles.
"""
import json
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from rapt_b200 import engine, _lib
    _lib.init(0)
    return engine


def synth_grid_class():
    from rapt_b200 import fields, synth

    class SynthGrid(fields.Grid):
        def parsefile(self, filename):
            return synth.dipole_grid_slice(int(filename))
    return SynthGrid


@pytest.fixture(scope="module")
def gold():
    d, _ = H.load("grid_synthetic")
    files = [str(s) for s in d["files"]]
    assert H.synthetic_grid(files)["sha256"] == str(d["checksum"])
    return d, files


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_grid_field_operators(eng, gold, arith):
    d, files = gold
    f = synth_grid_class()(files)
    assert f.gradientstepsize == float(d["ops_gradstep"]) and f.static == bool(d["ops_static"])
    ops = eng.field_ops(f, d["ops_pts"], which=["B", "E", "magB", "unitb", "gradB", "curlb", "lengthscale"], arith=arith)
    if arith == "strict":
        # same operations in the same order as scipy's RegularGridInterpolator: bit for bit
        assert np.array_equal(ops["B"], d["ops_B"]) and np.array_equal(ops["E"], d["ops_E"])
    else:
        assert H.vec_relerr(ops["B"], d["ops_B"]) < 1e-14 and H.relerr(ops["E"][:, 1], d["ops_E"][:, 1]) < 1e-14
    assert H.relerr(ops["magB"], d["ops_magB"]) < 1e-14
    assert H.vec_relerr(ops["unitb"], d["ops_unitb"]) < 1e-14
    # central differences over 2e-3 Re of a piecewise-linear field: differences of nearly equal numbers
    assert H.vec_relerr(ops["gradB"], d["ops_gradB"]) < 1e-9
    scale = np.max(np.abs(d["ops_curlb"]))
    assert np.max(np.abs(ops["curlb"] - d["ops_curlb"])) < 1e-9 * scale
    assert H.relerr(ops["lengthscale"], d["ops_lengthscale"]) < 1e-9


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_grid_particle_and_guiding_centre_trajectories(eng, gold, arith):
    """Particle.advance and GuidingCenter.advance in the gridded field; both cross t = 1.5 s, where the
    reference moves its three-point time window (fields.py:737-738)."""
    import rapt_b200 as R
    d, files = gold
    SynthGrid = synth_grid_class()
    old = dict(R.params)
    try:
        R.params.update(json.loads(str(d["p_params"])))
        R.params["arith"] = arith
        p = R.Particle(pos=tuple(d["p_pos"]), vel=tuple(d["p_vel"]), t0=0, mass=float(d["p_mass"]),
                       charge=float(d["p_charge"]), field=SynthGrid(files))
        p.advance(float(d["p_delta"]))
        traj = d["p_traj"]
        assert p.trajectory.shape == traj.shape
        assert H.relerr(p.trajectory[:, 0], traj[:, 0], floor=1e-3) < 1e-12
        assert H.vec_relerr(p.trajectory[:, 1:4], traj[:, 1:4]) < 1e-8
        assert H.vec_relerr(p.trajectory[:, 4:7], traj[:, 4:7]) < 1e-8
        if arith == "strict":
            assert tuple(p.solver_counters) == tuple(d["p_counters"].sum(0))
        else:       # the kinks of a piecewise-linear field put many error estimates next to 1
            assert abs(int(p.solver_counters[1]) - int(d["p_counters"][:, 1].sum())) <= 0.02 * d["p_counters"][:, 1].sum()
        R.params.clear(); R.params.update(old)
        R.params.update(json.loads(str(d["g_params"])))
        R.params["arith"] = arith
        g = R.GuidingCenter(pos=tuple(d["g_pos"]), v=float(d["g_v"]), pa=float(d["g_pa"]), mass=float(d["g_mass"]),
                            charge=float(d["g_charge"]), field=SynthGrid(files))
        assert H.relerr(g.mu, float(d["g_mu"])) < 1e-13
        g.advance(float(d["g_delta"]))
        traj = d["g_traj"]
        assert g.trajectory.shape == traj.shape
        assert np.allclose(g.trajectory[:, 0], traj[:, 0], rtol=0, atol=1e-12)
        # The guiding-centre right-hand side differentiates the field: on a multilinear interpolant grad|B|
        # and curl b are piecewise constant, i.e. the ODE is DISCONTINUOUS at every cell face.  The reference
        # itself rejects 965 of 2385 step attempts here and its result depends on the step sequence.  The
        # resolution of the comparison is therefore MEASURED, not chosen: the fixture holds the spread of the
        # reference's own result over eight reruns whose start position is moved by a few ulp
        # (oracle/gen_golden.py:case_grid -- positions 4.9e-5 ... 1.1e-4, p_par 2.3e-4 ... 5.2e-4 of its scale,
        # step counts +0.3 ... +3.3 %).  Gate: rows and times equal, and the CUDA result inside twice that band;
        # the right-hand side itself is compared bit for bit in test_grid_gc_rhs_probe below.
        band_pos, band_pp = float(np.max(d["g_band_pos"])), float(np.max(d["g_band_ppar"]))
        ref = d["g_counters"].sum(0)
        band_ns = float(np.max(np.abs(d["g_band_nstep"] - ref[1]))) / ref[1]
        assert 4e-5 < band_pos < 2e-4 and 2e-4 < band_pp < 1e-3 and 0.02 < band_ns < 0.05      # what was measured
        assert H.vec_relerr(g.trajectory[:, 1:4], traj[:, 1:4]) < 2 * band_pos
        pscale = np.max(np.abs(traj[:, 4]))
        assert np.max(np.abs(g.trajectory[:, 4] - traj[:, 4])) < 2 * band_pp * pscale
        assert abs(int(g.solver_counters[1]) - int(ref[1])) <= 2 * band_ns * ref[1]
    finally:
        R.params.clear(); R.params.update(old)


def test_grid_leaving_the_grid_raises_value_error(eng, gold):
    import rapt_b200 as R
    d, files = gold
    old = dict(R.params)
    try:
        R.params["cyclotronresolution"] = 10
        p = R.Particle(pos=tuple(d["oob_pos"]), vel=tuple(d["oob_vel"]), t0=0, mass=R.m_pr, charge=R.e,
                       field=synth_grid_class()(files))
        with pytest.raises(ValueError):
            p.advance(5.0)
        assert len(p.trajectory) == len(d["oob_traj"])          # rows before the failure are kept, none after
    finally:
        R.params.clear(); R.params.update(old)


@pytest.mark.parametrize("nfiles", [1, 4])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_grid_ensemble_vs_oracle(eng, arith, nfiles):
    """512 protons started inside the grid, 0.4 s: final states against the CPU oracle (which is pinned bit
    for bit against the reference on the same field); tracers that leave the grid report RAPT_ST_FIELD in both.
    nfiles = 1 is the time-independent (3-D) case."""
    import oracle as O
    from rapt_b200 import m_pr, e, Re
    from rapt_b200.utils import speedfromKE
    files = [str(k) for k in range(nfiles)]
    G = H.synthetic_grid(files)
    fo = O.make_grid_field(G["t"], G["x"], G["y"], G["z"], G["B"], G["E"])
    fg = synth_grid_class()(files)
    rng = np.random.default_rng(5)
    n = 512
    pos = np.column_stack([rng.uniform(4.0, 7.0, n), rng.uniform(-1.0, 1.0, n), rng.uniform(-1.5, 1.5, n)]) * Re
    v = np.array([speedfromKE(k, m_pr, 'ev') for k in 10 ** rng.uniform(5, 6.3, n)])
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    vel = dirs * v[:, None]
    mom = eng.particle_momentum(vel, np.full(n, m_pr))
    st0 = np.column_stack([np.zeros(n), pos, mom])
    og = eng.particle_advance(fg, st0, m_pr, e, 0.4, arith=arith, cyclotronresolution=10)
    oo = O.particle_advance(fo, O.make_params(cyclotronresolution=10), st0, m_pr, e, 0.4, store_every=0, nthreads=8)
    assert np.array_equal(og["status"], oo["status"])
    ok = og["status"] == 1
    assert ok.sum() > 0.9 * n
    # the interpolant is continuous but not differentiable at cell faces: an accept/reject decision that flips
    # there moves a tracer by ~1e-8, so the bar is 1e-8 for 99 % of the ensemble and 1e-7 for its maximum
    for a, b in ((1, 4), (4, 7)):
        err = np.linalg.norm(og["state"][ok, a:b] - oo["state"][ok, a:b], axis=1) / np.linalg.norm(oo["state"][ok, a:b], axis=1)
        assert np.quantile(err, 0.99) < 1e-8 and err.max() < 1e-7
    if arith == "strict":
        same = np.all(og["counters"][ok] == oo["counters"][ok], axis=1)
        assert same.mean() > 0.98


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_grid_gc_rhs_probe(eng, gold, arith):
    """One accepted 1e-6 s step from 256 random guiding centres: (y1 - y0) is the RK combination of the
    right-hand side alone.  Strict flavour: identical to the oracle (which equals the reference bit for bit
    on this field); fast flavour: 1e-10."""
    import oracle as O
    from rapt_b200 import Re
    d, files = gold
    G = H.synthetic_grid(files)
    fo = O.make_grid_field(G["t"], G["x"], G["y"], G["z"], G["B"], G["E"])
    fg = synth_grid_class()(files)
    mass, q, v = float(d["g_mass"]), float(d["g_charge"]), float(d["g_v"])
    rng = np.random.default_rng(3); m = 256
    pos = np.column_stack([rng.uniform(4, 7, m), rng.uniform(-1, 1, m), rng.uniform(-1.5, 1.5, m)]) * Re
    pa = rng.uniform(20, 160, m)
    pp, mu = O.gc_construct(fo, 0.3, pos, np.full(m, v), pa, mass)
    s0 = np.column_stack([np.full(m, 0.3), pos, pp])
    og = eng.gc_advance(fg, s0, mu, v, mass, q, 1e-6, 1e-6, arith=arith)
    oo = O.gc_advance(fo, O.make_params(), s0, mu, v, mass, q, 1e-6, 1e-6, store_every=0)
    dg, do = og["state"][:, 1:5] - s0[:, 1:5], oo["state"][:, 1:5] - s0[:, 1:5]
    if arith == "strict":
        assert np.array_equal(dg, do)
    else:
        assert H.vec_relerr(dg[:, :3], do[:, :3]) < 1e-10
        assert np.max(np.abs(dg[:, 3] - do[:, 3])) < 1e-10 * np.max(np.abs(do[:, 3]))
