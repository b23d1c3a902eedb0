"""GPU parity: Adaptive (particle <-> guiding-centre switching) and the reference-shaped classes
(Particle, GuidingCenter, Adaptive objects) against golden vectors from the reference."""
import io
import contextlib
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rapt_b200
    from rapt_b200 import _lib
    _lib.init(0)
    return rapt_b200


@pytest.fixture(autouse=True)
def reset_params(rb):
    saved = dict(rb.params)
    yield
    rb.params.clear(); rb.params.update(saved)


def split_segments(d):
    rows, out, k = d["rows"], [], 0
    for m, n in zip(d["seg_mode"], d["seg_nrows"]):
        out.append((int(m), rows[k:k + n])); k += n
    return out


@pytest.fixture(params=[1, 16], ids=["unsliced", "16slices"])
def slices(request, monkeypatch):
    """Epochs unsliced (default) and time-sliced: a sliced run must resume every call bit-identically."""
    monkeypatch.setenv("RAPT_B200_ADAPTIVE_SLICES", str(request.param))
    return request.param


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("name", ["g3_speiser", "e4_speiser_1", "e4_speiser_2", "e4_speiser_3", "e4_speiser_4", "e4_speiser_5"])
def test_adaptive_ensemble_kernel_vs_reference(rb, name, arith, slices):
    """Device epoch loop (advance kernels + switch/compaction kernel), one tracer per golden file."""
    d, par = H.load(name)
    f = H.gpu_field(*H.ADAPTIVE_CASES[name])
    ref = split_segments(d)
    nref = len(d["rows"])
    o = rb.engine.adaptive_advance(f, d["pos"], d["vel"], 0.0, float(d["mass"]), float(d["charge"]), float(d["delta"]),
                                   par["GCtimestep"], store_every=1, max_rows=nref + 64, arith=arith,
                                   solvertolerances=par["solvertolerances"], epss=par["epss"])
    assert o["status"][0] == 1
    assert o["nseg"][0] == len(ref), "number of mode switches must match"
    rows = o["rows"][0, :o["nstored"][0]]
    tags = rows[:, 7].astype(int)
    segs = [(int(t & 1), rows[tags == t]) for t in sorted(set(tags))]
    assert [m for m, _ in segs] == [m for m, _ in ref], "mode sequence"
    # first segments (before the current-sheet crossing amplifies round-off) must agree row for row
    m0, r0 = segs[0]; _, g0 = ref[0]
    assert len(r0) == len(g0)
    ncol = 7 if m0 == 0 else 5
    assert np.max(np.abs(r0[:, :ncol] - g0[:, :ncol]) / (np.abs(g0[:, :ncol]) + 1e-3)) < 1e-8
    # switch times: segment start times (gate on ~1e-10 for the chaotic Speiser case)
    t_sw = np.array([s[1][0, 0] for s in segs]); t_ref = np.array([s[1][0, 0] for s in ref])
    assert np.max(np.abs(t_sw - t_ref)) < 1e-7 * max(1.0, np.max(np.abs(t_ref)))
    # row counts per segment: equal, or off by a row or two where a switch sits on a round-off knife edge
    for (m, r), (mr, g) in zip(segs, ref):
        assert abs(len(r) - len(g)) <= 2
    # final state (chaotic: x1e6 amplification through the sheet crossing)
    fin = o["final"][0]; gl = ref[-1][1][-1]
    assert abs(fin[0] - gl[0]) < 1e-6
    assert np.linalg.norm(fin[1:4] - gl[1:4]) / np.linalg.norm(gl[1:4]) < 1e-5
    if name == "g3_speiser":
        # the notebook-stored answers (examples/Adaptive Example - Speiser orbits.ipynb:114,121)
        assert t_sw[1] == 168.0
        assert abs(t_sw[2] - 271.537802289) < 1e-7


def test_adaptive_ensemble_many(rb):
    """256 Speiser tracers of config 4 on the device vs the CPU oracle: segment structure and switch
    times; exercises the ballot/shared-memory regrouping with mixed modes in one warp."""
    import oracle as O
    from rapt_b200 import synth
    n = 256
    ic = synth.config4_speiser(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    par = dict(solvertolerances=(1e-12, 1e-12), epss=0.02)
    o = rb.engine.adaptive_advance(H.gpu_field("Parabolic", ()), pos, vel, 0.0, 1.0, 1.0, 200.0, 1.0, store_every=1,
                                   max_rows=2048, arith="strict", **par)
    assert np.all(o["status"] == 1)
    assert o["epochs"] >= 2
    of = O.make_field("Parabolic"); op = O.make_params(GCtimestep=1, **par)
    bad = 0
    for i in range(n):
        nseg, rows, seglog, cnt = O.adaptive_c(of, op, pos[i], vel[i], 0.0, 1.0, 1.0, 200.0)
        mine = o["rows"][i, :o["nstored"][i]]
        tags = mine[:, 7].astype(int)
        if o["nseg"][i] != nseg:
            bad += 1
            continue
        starts = np.array([mine[tags == t][0, 0] for t in sorted(set(tags))])
        ref_starts = np.array([rows[int(s[1]), 0] for s in seglog])
        if np.max(np.abs(starts - ref_starts)) > 1e-6:
            bad += 1
    assert bad <= 2, f"{bad} of {n} tracers differ in segment structure"


def test_particle_object_readme(rb):
    """README example through the reference-shaped class (README.md:33-52)."""
    from numpy import sin, cos, pi
    d, par = H.load("g1_readme")
    rb.params["cyclotronresolution"] = 20
    v = rb.utils.speedfromKE(1e6, rb.m_pr, 'ev'); pa = 30 * pi / 180
    p = rb.Particle(pos=(6 * rb.Re, 0, 0), vel=(0, -v * sin(pa), v * cos(pa)), t0=0, mass=rb.m_pr, charge=rb.e,
                    field=rb.fields.EarthDipole())
    p.advance(10)
    assert p.trajectory.shape == d["traj"].shape == (434, 7)
    assert H.vec_relerr(p.trajectory[:, 1:4], d["traj"][:, 1:4]) < 1e-8
    assert abs(p.tcur - float(d["tcur"])) < 1e-12
    assert p.gett()[-1] == pytest.approx(10.002375083910456, rel=1e-14)
    assert np.allclose(p.getke(), p.getke()[0], rtol=1e-5)
    # pickling round trip (save/load, Particle.py:311-343)
    import tempfile, os
    fn = os.path.join(tempfile.mkdtemp(), "p.pkl")
    p.save(fn); q = rb.Particle(); q.load(fn)
    assert np.array_equal(q.trajectory, p.trajectory)
    q.advance(0.5)
    assert len(q.trajectory) > len(p.trajectory)


def test_gc_object_notebook(rb):
    """GuidingCenter notebook cell 5-6 (bounce-period output step) through the class."""
    d, par = H.load("g2_gc_doubledipole")
    g = rb.GuidingCenter(pos=(0, -10 * rb.Re, 0), v=rb.utils.speedfromKE(1e5, rb.m_el), pa=80, mass=rb.m_el,
                         charge=-rb.e, field=rb.fields.DoubleDipole())
    assert g.mu == pytest.approx(float(d["mu"]), rel=1e-14)
    assert g.bounceperiod() == pytest.approx(float(d["bs_period"]), rel=1e-6)
    g.advance(20)
    # dt = bounceperiod()/10 carries the ~1e-7 FD noise of the curvature -> compare on the common rows loosely
    assert abs(len(g.trajectory) - len(d["traj"])) <= 1
    k = min(len(g.trajectory), len(d["traj"]))
    assert np.max(np.abs(g.gett()[:k] - d["traj"][:k, 0])) < 1e-4
    assert H.vec_relerr(g.trajectory[:k, 1:4], d["traj"][:k, 1:4]) < 1e-5
    assert np.all(np.isfinite(g.getke())) and np.all(g.getB() > 0)


def test_adaptive_object_speiser(rb):
    """Adaptive notebook (examples/Adaptive Example - Speiser orbits.ipynb cells 5-8): printed switch times."""
    d, par = H.load("g3_speiser")
    rb.params['solvertolerances'] = (1e-12, 1e-12); rb.params['epss'] = 0.02
    rb.params['Ptimestep'] = 0.1; rb.params["GCtimestep"] = 1
    rb.params["arith"] = "strict"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        pa = rb.Adaptive([5, -5, 0.9], [-0.1, 0.1, 0], 0, mass=1, charge=1, field=rb.fields.Parabolic())
        pa.advance(300)
    out = buf.getvalue().splitlines()
    assert out[0] == "Switched to particle mode at time 168.0"
    assert out[1].startswith("Switched to guiding center mode at time 271.5378022")
    assert [type(s).__name__ for s in pa.trajlist] == ["GuidingCenter", "Particle", "GuidingCenter"]
    assert len(pa.gett()) == pytest.approx(1990, abs=3)
    assert len(pa.getx()) == len(pa.gett()) == len(pa.getke())


def test_ensemble_classes_host_and_device(rb):
    """ParticleEnsemble on host buffers and device-resident give identical results."""
    import torch
    from rapt_b200 import synth
    n = 2000
    ic = synth.config2_protons(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    rb.params["cyclotronresolution"] = 20
    a = rb.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], rb.fields.EarthDipole()).advance(0.2, store_every=4, max_rows=32)
    b = rb.ParticleEnsemble(pos, vel, 0.0, ic["mass"], ic["charge"], rb.fields.EarthDipole()).cuda().advance(0.2)
    b.cpu()
    assert np.array_equal(a.state, b.state), "host-pointer and device-pointer paths must agree bit for bit"
    assert np.array_equal(a.last_counters, b.last_counters)
    assert a.member_trajectory(3).shape[1] == 7
    ic3 = synth.config3_electrons(500)
    pos3 = np.column_stack([ic3["x"], ic3["y"], ic3["z"]])
    rb.params["GCtimestep"] = 0.1
    g1 = rb.GuidingCenterEnsemble(pos3, ic3["v"], pa=ic3["pa"], mass=ic3["mass"], charge=ic3["charge"], field=rb.fields.DoubleDipole())
    g1.advance(2.0)
    g2 = rb.GuidingCenterEnsemble(pos3, ic3["v"], pa=ic3["pa"], mass=ic3["mass"], charge=ic3["charge"], field=rb.fields.DoubleDipole()).cuda()
    g2.advance(2.0).cpu()
    assert np.array_equal(g1.state, g2.state)
    ke0 = g1.getke()
    assert np.all(ke0 > 0)
