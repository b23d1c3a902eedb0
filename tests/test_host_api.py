"""CPU: the C-ABI library loads and exports every symbol include/rapt_b200.h declares, fails loudly
without a device, and the host-side mirror of the reference interface (fields, utils, params,
constructors, getters) reproduces the reference's values.  No compute calls on the GPU here."""
import ctypes
import os
import re
import numpy as np
import pytest

import helpers as H
import scipy_legs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rapt_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "rapt_b200.h")).read()
    names = sorted(set(re.findall(r"\b(rapt_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rapt_b200.h but not exported"
    assert b"sm_100a" in lib.rapt_b200_version()


def test_struct_layouts_match_header():
    from rapt_b200._lib import FieldT, ParamsT
    assert ctypes.sizeof(FieldT) == 4 * 4 + 16 * 8 + 2 * 8
    assert ctypes.sizeof(ParamsT) == 5 * 8 + 8 * 4


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry raises; with one, this test is skipped."""
    from rapt_b200 import _lib, engine, fields
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.RaptB200Error, match="no CUDA device"):
        engine.particle_advance(fields.EarthDipole(), np.zeros((1, 7)) + 1.0, 1.0, 1.0, 1.0)
    with pytest.raises(_lib.RaptB200Error):
        engine.field_ops(fields.EarthDipole(), np.ones((2, 4)))


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rapt_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "import oracle" not in src and "liboracle" not in src and "rapt_oracle" not in src, fn


def test_host_fields_match_reference():
    from rapt_b200 import fields
    u = np.load(H.GOLDEN + "/units.npz")
    mk = {"earthdipole": fields.EarthDipole(), "doubledipole": fields.DoubleDipole(), "uniformbz": fields.UniformBz(2e-4),
          "crossedeb": fields.UniformCrossedEB(2.0, 1e-4), "vardipole": fields.VarEarthDipole(0.1, 10),
          "parabolic": fields.Parabolic()}
    for name, f in mk.items():
        for i, tp in enumerate(u[name + "_pts"]):
            assert np.array_equal(f.B(tp), u[name + "_B"][i])
            assert np.array_equal(np.asarray(f.E(tp), float), u[name + "_E"][i])
            assert np.array_equal(f.gradB(tp), u[name + "_gradB"][i])
            assert np.array_equal(f.jacobianB(tp), u[name + "_jacobianB"][i])
            assert np.allclose(f.curlb(tp), u[name + "_curlb"][i], rtol=0, atol=1e-9 * (np.abs(u[name + "_curlb"][i]).max() + 1e-30) + 1e-300) \
                or np.array_equal(f.curlb(tp), u[name + "_curlb"][i])
            assert f.magB(tp) == u[name + "_magB"][i]
            assert f.curvature(tp) == u[name + "_curvature"][i] or np.isnan(u[name + "_curvature"][i])
    d = fields.EarthDipole().device_descriptor()
    assert d.kind == 0 and d.is_static == 1 and d.gradstep == 6378137 * 1e-6 and d.prm[0] == -3 * 3.07e-5 * 6378137 ** 3
    d = fields.UniformCrossedEB(2.0, 1e-4).device_descriptor()
    assert d.kind == 3 and d.is_static == 0 and (d.prm[0], d.prm[1]) == (1e-4, 2.0)
    with pytest.raises(NotImplementedError):
        class NoSnippet(fields._Field):
            pass
        NoSnippet().device_descriptor()


def test_host_utils_match_reference():
    from rapt_b200 import utils as ru, fields, m_pr, m_el, e
    u = np.load(H.GOLDEN + "/units.npz")
    f = fields.DoubleDipole()
    for i in range(len(u["utils_pos"])):
        pos, vel = u["utils_pos"][i], u["utils_vel"][i]
        assert ru.cyclotron_period(0, pos, vel, f, m_pr, e) == u["utils_cycper"][i]
        assert ru.cyclotron_radius(0, pos, vel, f, m_pr, e) == u["utils_cycrad"][i]
        R, vp, v = ru.guidingcenter(0, pos, vel, f, m_pr, e)
        assert np.array_equal(R, u["utils_gc_R"][i]) and vp == u["utils_gc_vp"][i] and v == u["utils_gc_v"][i]
        assert ru.magnetic_moment(0, R, vp, v, f, m_pr) == u["utils_mu"][i]
        pp, vv = ru.GCtoFP(0, R, vp, v, f, m_pr, e, 0)
        assert np.array_equal(pp, u["utils_fp_pos"][i]) and np.array_equal(vv, u["utils_fp_vel"][i])
    for v, w in zip(u["getperp_in"], u["getperp_out"]):
        assert np.array_equal(np.asarray(ru.getperp(v), float), w)
    for k, a, b in zip(u["speed_ke"], u["speed_pr"], u["speed_el"]):
        assert ru.speedfromKE(k, m_pr) == a and ru.speedfromKE(k, m_el) == b


def test_constructors_and_params_snapshot():
    import rapt_b200 as rb
    d, _ = H.load("g1_readme")
    p = rb.Particle(pos=d["pos"], vel=d["vel"], t0=0, mass=rb.m_pr, charge=rb.e, field=rb.fields.EarthDipole())
    assert np.array_equal(p.trajectory, d["traj"][:1])          # numpy arrays accepted (reference needs tuples on numpy 2)
    assert p.tcur == 0 and p.check_adiabaticity is False
    assert rb.Particle().trajectory.shape == (1, 7)
    d2, _ = H.load("g2_gc_doubledipole")
    g = rb.GuidingCenter(pos=tuple(d2["pos"]), v=float(d2["v"]), pa=80, mass=rb.m_el, charge=-rb.e, field=rb.fields.DoubleDipole())
    assert g.mu == float(d2["mu"]) and np.array_equal(g.trajectory, d2["traj"][:1])
    g90 = rb.GuidingCenter(pos=(-7.8 * rb.Re, 0, 0), v=1e6, pa=90, mass=rb.m_pr, charge=rb.e, field=rb.fields.DoubleDipole())
    assert g90.trajectory[0, 4] == 0.0                           # quirk Q9: pa == 90 -> p_par exactly 0
    assert set(["cyclotronresolution", "Ptimestep", "bounceresolution", "GCtimestep", "BCtimestep", "solvertolerances",
                "fieldlineresolution", "flsolver", "eyegradientstep", "epss", "epst", "enforce equatorial"]) <= set(rb.params)
    assert rb.params["solvertolerances"] == (1.49012e-8, 1.49012e-8) and rb.params["cyclotronresolution"] == 10
    s = rb.engine.snapshot_params(None, True, cyclotronresolution=20, **{"enforce equatorial": True})
    assert (s.cyclotronresolution, s.enforce_equatorial, s.check_adiabaticity, s.rtol) == (20.0, 1, 1, 1.49012e-8)
    assert rb.params["cyclotronresolution"] == 10, "snapshot overrides must not leak into the global params"
    with pytest.raises(RuntimeError):
        rb.BounceCenter(pos=(1, 0, 0), v=1.0, pa=30, mass=1, charge=1, field=rb.fields.UniformCrossedEB())
    b = rb.BounceCenter(pos=(4 * rb.Re, 0, 0), v=1e7, pa=0.5, mass=rb.m_pr, charge=rb.e, field=rb.fields.EarthDipole())
    assert b.trajectory.shape == (1, 4) and b.mu > 0
    b.isequatorial = True
    with pytest.raises(NotImplementedError):                   # the reference's branch cannot run (BounceCenter.py:235)
        b.advance(1.0)
    assert issubclass(rb.Adiabatic, Exception) and issubclass(rb.NonAdiabatic, Exception)


def test_synthetic_ensembles_are_deterministic():
    from rapt_b200 import synth
    a, b = synth.config2_protons(1000), synth.config2_protons(1000)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert np.all(np.abs(a["z"]) > 0) and np.all(a["x"] != 0) and np.all(a["y"] != 0)
    r = np.sqrt(a["x"] ** 2 + a["y"] ** 2) / synth.Re
    assert r.min() >= 2 and r.max() <= 6
    assert a["ke_ev"].min() >= 1e5 and a["ke_ev"].max() <= 1e7
    v = np.sqrt(a["vx"] ** 2 + a["vy"] ** 2 + a["vz"] ** 2)
    assert np.all(v < synth.c)
    c3 = synth.config3_electrons(1000)
    assert np.all(c3["x"] <= 8 * synth.Re + 1)
    c4 = synth.config4_speiser(8)
    assert (c4["x"][0], c4["y"][0], c4["z"][0], c4["vx"][0], c4["vy"][0]) == (5.0, -5.0, 0.9, -0.1, 0.1)


def test_halfbouncepath_host_leg_matches_reference():
    """The host leg of GuidingCenter.bounceperiod (flutils.py:274-316: trimming, equatorial 3-point formula,
    scipy spline/brentq/quad) on the reference's own traced curves gives the reference's value exactly."""
    from rapt_b200 import engine
    for name in ("g2_gc_doubledipole", "gc_earthdipole", "gc_pa90_equatorial"):
        d, _ = H.load(name)
        s, b, Bm = d["bs_curve"][:, 0], d["bs_B"], float(d["bs_Bm"])
        hp = scipy_legs.halfbouncepath_from_curve(s, b, Bm)
        assert hp == float(d["bs_halfpath"]), name
        assert (2 / float(d["bs_v"])) * hp == float(d["bs_period"])


def test_second_invariant_host_leg_matches_reference():
    """The host leg of GuidingCenter.geteye (flutils.py:65-151: trimming to one point beyond each mirror
    point, Simpson + closed-form end intervals below 70 degrees, spline/brentq/quad above) on the reference's
    own traced field lines gives the reference's I exactly.  eye_pa45_simpson was generated with the name
    `simps` (undefined in the reference, flutils.py:130) bound to scipy.integrate.simpson -- see its flag."""
    from rapt_b200 import engine
    for name, bound in (("eye_pa80", False), ("eye_pa45_simpson", True)):
        d, _ = H.load(name)
        assert bool(d["simps_name_bound"]) == bound
        for i in range(len(d["Bm"])):
            cv = d["curves"][i, :d["npts"][i]]
            assert scipy_legs.eye_from_curve(cv[:, 0], cv[:, 1], float(d["Bm"][i])) == d["eye"][i, 1], (name, i)
    # no mirror point on the line / equatorial particle -> 0 (flutils.py:106-109)
    s = np.linspace(0, 1, 9); b = 1 + (s - 0.5) ** 2
    assert scipy_legs.eye_from_curve(s, b, 0.9) == 0.0 and scipy_legs.eye_from_curve(s, b, 1.0) == 0.0


def test_grid_host_interpolation_matches_reference():
    """fields.Grid host side (Bgrid/Egrid and the derived operators of _Field) on the synthetic data files
    equals the reference's values bit for bit; outside the grid it raises ValueError like scipy."""
    from rapt_b200 import fields, synth, Re
    d, _ = H.load("grid_synthetic")
    files = [str(s) for s in d["files"]]

    class SynthGrid(fields.Grid):
        def parsefile(self, filename):
            return synth.dipole_grid_slice(int(filename))
    f = SynthGrid(files)
    assert f.gradientstepsize == float(d["ops_gradstep"]) and f.static == bool(d["ops_static"])
    for i, tp in enumerate(d["ops_pts"]):
        assert np.array_equal(f.B(tp), d["ops_B"][i]) and np.array_equal(f.E(tp), d["ops_E"][i])
        assert f.magB(tp) == d["ops_magB"][i] and np.array_equal(f.gradB(tp), d["ops_gradB"][i])
        assert np.array_equal(f.curlb(tp), d["ops_curlb"][i])
    with pytest.raises(ValueError):
        f.B([0.5, 9 * Re, 0, 0])
    with pytest.raises(ValueError):
        f.E([3.5, 5 * Re, 0, 0])          # time beyond the last data file
    one = SynthGrid(files[:1])            # single file: time-independent, any t
    assert np.array_equal(one.B([123.0, 5 * Re, 0.1 * Re, 0.2 * Re]), f.B([0.0, 5 * Re, 0.1 * Re, 0.2 * Re]))


def test_nystrom_tables_are_consistent_with_the_tableau():
    """rapt_particle_rkn.cuh integrates in Nystrom form with A.A, b.A, er.A, w.A (tools/gen_coeffs.py):
    check the generated constants against the DOP853 tableau and the order conditions they must inherit."""
    import re
    from fractions import Fraction as Fr
    from scipy.integrate._ivp import dop853_coefficients as dc
    co = open(os.path.join(ROOT, "rapt_b200", "csrc", "dop_coeffs.h")).read()
    val = {m.group(1): float(m.group(2)) for m in re.finditer(r"#define (D8N?_\w+) (\S+)", co)}
    A = dc.A[:12, :12]; B = dc.B[:12]; E5 = dc.E5[:12]; E3 = dc.E3[:12]
    AA = A @ A
    for i in range(2, 12):
        assert abs(val[f"D8N_RS{i+1}"] - A[i].sum()) < 1e-15
        assert abs(val[f"D8N_RS{i+1}"] - dc.C[i]) < 3e-15          # row-sum condition c_i = sum_j a_ij
        for l in range(i - 1):
            got = val.get(f"D8N_AA{i+1}_{l+1}", 0.0)
            assert abs(got - AA[i, l]) < 1e-14 * max(1.0, abs(AA[i, l])), (i, l)
    for name, w in (("BA", B), ("ERA", E5), ("WA", E3)):
        wa = w @ A
        for l in range(11):
            assert abs(val.get(f"D8N_{name}{l+1}", 0.0) - wa[l]) < 1e-14 * max(1.0, abs(wa[l])), (name, l)
    assert val["D8N_SB"] == 1.0 and abs(val["D8N_SER"]) < 1e-16 and abs(val["D8N_SW"]) < 1e-15
    # second-order condition of the position update: sum_l (b.A)_l = 1/2
    assert abs(sum(val.get(f"D8N_BA{l+1}", 0.0) for l in range(11)) - 0.5) < 1e-15
    # exact arithmetic agrees with what the generator rounded
    exact = float(sum(Fr(float(B[j])) * Fr(float(A[j, 5])) for j in range(12)))
    assert val["D8N_BA6"] == exact


def test_bounce_center_host_side():
    """BounceCenter (rapt/BounceCenter.py:74-115): constructor behaviour on the host; advance and the flutils
    integrals are device-only (no CPU fallback)."""
    import rapt_b200 as rb
    from rapt_b200 import _lib
    f = rb.fields.EarthDipole()
    d = np.load(os.path.join(ROOT, "tests", "golden", "bc_dipole_electron.npz"))
    b = rb.BounceCenter(pos=tuple(d["pos"]), v=float(d["v"]), t0=0, pa=float(d["pa"]), mass=float(d["mass"]),
                        charge=float(d["charge"]), field=f)
    assert b.trajectory.shape == (1, 4) and b.tcur == 0
    assert b.mu == pytest.approx(float(d["mu"]), rel=1e-14)          # cos() of the pitch angle as given (BounceCenter.py:114)
    assert b._mirror_field() == pytest.approx(float(d["Bm"]), rel=1e-14)
    with pytest.raises(RuntimeError, match="nonstatic"):
        rb.BounceCenter(pos=(1, 1, 1), v=1.0, pa=80, mass=1.0, charge=1.0, field=rb.fields.VarEarthDipole())
    with pytest.raises(RuntimeError, match="nonstatic"):
        rb.BounceCenterEnsemble(np.ones((2, 3)), 1.0, 0.0, 1.0, 1.0, 1.0, rb.fields.VarEarthDipole())
    # setpa: same energy, new pitch angle, data reset (BounceCenter.py:117-132 keeps the constructor's cos() quirk)
    b.setpa(30.0)
    assert b.trajectory.shape == (1, 4) and b.pa == 30.0 and b.mu != pytest.approx(float(d["mu"]), rel=1e-3)
    if _lib.device_count() == 0:
        with pytest.raises(_lib.RaptB200Error, match="no CUDA device"):
            b.advance(0.1)
        with pytest.raises(_lib.RaptB200Error, match="no CUDA device"):
            rb.eye(d["pts"][0], f, float(d["Bm"]))
        with pytest.raises(_lib.RaptB200Error, match="no CUDA device"):
            rb.engine.bounce_center_advance(f, d["traj"][0], float(d["mu"]), float(d["v"]), float(d["mass"]),
                                            float(d["charge"]), 0.1)


def test_user_field_module_compiles_with_nvrtc_without_a_device():
    """rapt_b200_field_nvrtc compiles the embedded kernel headers around a user snippet at registration time; NVRTC
    needs no device, so a header that nvcc accepts but NVRTC rejects shows up here and not first on the GPU box."""
    import userfield
    from rapt_b200 import engine
    CD = userfield.make_charged_dipole()
    uid = engine.compile_user_field(CD.cuda_source, True)
    assert uid >= 0
    with pytest.raises(Exception, match="NVRTC"):
        engine.compile_user_field("__device__ void rapt_user_B(double t) { this is not CUDA }", False)


def test_integration_doc_covers_every_entry_point():
    """INTEGRATION.md shows (or tabulates) the reference-side binding of every symbol the header declares."""
    hdr = open(os.path.join(ROOT, "include", "rapt_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(rapt_b200_[a-z0-9_]+)\s*\(", hdr))
    missing = sorted(s for s in declared if s not in doc)
    assert not missing, missing


def test_gpu_test_files_only_call_functions_that_exist():
    """Static check, runs without a GPU: every `eng.X` / `engine.X` / `scipy_legs.X` / `rd.X` attribute the -m gpu test
    files (and bench.py, smoke()) touch exists in the module it names -- a renamed or moved helper is caught here, not on
    the B200 box."""
    import ast
    import glob
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    targets = {"eng": "rapt_b200.engine", "engine": "rapt_b200.engine", "scipy_legs": "scipy_legs", "rd": "rapt_b200.dist",
               "synth": "rapt_b200.synth", "_lib": "rapt_b200._lib"}
    mods = {k: importlib.import_module(v) for k, v in targets.items()}
    files = glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"),
                                                                     os.path.join(ROOT, "tools", "count_parity_report.py"),
                                                                     os.path.join(ROOT, "tools", "tail_profile.py")]
    missing = []
    for fn in files:
        for node in ast.walk(ast.parse(open(fn).read())):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in mods:
                if not hasattr(mods[node.value.id], node.attr):
                    missing.append(f"{os.path.basename(fn)}:{node.lineno} {node.value.id}.{node.attr}")
    assert not missing, missing
