#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) here.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Run in the build container:

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py g1 e2      # selected cases

The reference cannot travel to the GPU box, so the vectors are committed as small fixtures.
Each fixture stores the inputs, rapt.params overrides, the reference trajectory (all rows) and
the per-solver-call counters (nfcn, nstep, naccpt, nrejct) = scipy iwork[16:20]
(scipy 1.18.1 `_dop`, numpy 2.3.5).
"""
import os, sys, time, json
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import warnings
warnings.filterwarnings("ignore", category=SyntaxWarning)
import refshim
from rapt_b200 import synth

rapt = refshim.load_reference()
from rapt import fields as rf, utils as ru
GOLD = os.path.join(ROOT, "tests", "golden")
Re, e, m_pr, m_el, c = rapt.Re, rapt.e, rapt.m_pr, rapt.m_el, rapt.c


def save(name, **arrs):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"  wrote {name}.npz ({os.path.getsize(path)/1024:.1f} KiB)")


def take_log():
    L = np.array(refshim.SOLVER_LOG, dtype=np.int64).reshape(-1, 4)
    refshim.SOLVER_LOG.clear()
    return L


def run_particle(pos, vel, mass, charge, field, delta, t0=0.0, **par):
    refshim.reset_params(rapt, **par)
    refshim.SOLVER_LOG.clear()
    p = rapt.Particle(pos=tuple(pos), vel=tuple(vel), t0=t0, mass=mass, charge=charge, field=field)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p.advance(delta)
    return p, take_log()


def run_gc(pos, v, pa, mass, charge, field, delta, eom="TaoChanBrizardEOM", t0=0.0, **par):
    refshim.reset_params(rapt, **par)
    refshim.SOLVER_LOG.clear()
    g = rapt.GuidingCenter(pos=tuple(pos), v=v, pa=pa, t0=t0, mass=mass, charge=charge, field=field)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.advance(delta, eom=eom)
    return g, take_log()


def parjson(**par):
    return np.array(json.dumps(par))


# ----------------------------------------------------------------------------- cases
def case_g1():
    """README.md:33-52."""
    v = ru.speedfromKE(1e6, m_pr, 'ev'); pa = 30 * np.pi / 180
    pos = (6 * Re, 0, 0); vel = (0, -v * np.sin(pa), v * np.cos(pa))
    p, L = run_particle(pos, vel, m_pr, e, rf.EarthDipole(), 10, cyclotronresolution=20)
    save("g1_readme", pos=np.array(pos, float), vel=np.array(vel), mass=m_pr, charge=e, delta=10.0,
         params=parjson(cyclotronresolution=20), traj=p.trajectory, counters=L, tcur=p.tcur)


def case_g1b():
    v = ru.speedfromKE(1e6, m_pr, 'ev')
    pos = (3.1 * Re, 2.3 * Re, 0.7 * Re)
    vel = (v * 0.31, -v * 0.52, v * np.sqrt(1 - 0.31 ** 2 - 0.52 ** 2))
    p, L = run_particle(pos, vel, m_pr, e, rf.EarthDipole(), 5, cyclotronresolution=20)
    save("g1b_generic", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=5.0,
         params=parjson(cyclotronresolution=20), traj=p.trajectory, counters=L, tcur=p.tcur)
    # second advance() call on the same object (dt recomputed from the current state)
    refshim.SOLVER_LOG.clear()
    p.advance(1.0)
    save("g1b_second_call", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e,
         delta1=5.0, delta2=1.0, params=parjson(cyclotronresolution=20), traj=p.trajectory,
         counters=take_log(), tcur=p.tcur)


def case_pfields():
    """Particle.advance in every other analytic field model (+ enforce equatorial)."""
    out = {}
    # VarEarthDipole: static=False -> gm recomputed per RHS (Particle.py:290-291)
    v = ru.speedfromKE(5e5, m_el, 'ev')
    pos = (3.7 * Re, -1.1 * Re, 0.4 * Re); vel = (0.3 * v, 0.5 * v, v * np.sqrt(1 - .09 - .25))
    p, L = run_particle(pos, vel, m_el, -e, rf.VarEarthDipole(amp=0.1, period=10), 0.02, t0=1.5,
                        cyclotronresolution=20)
    save("p_vardipole", pos=np.array(pos), vel=np.array(vel), mass=m_el, charge=-e, delta=0.02, t0=1.5,
         fieldprm=np.array([0.1, 10.0]), params=parjson(cyclotronresolution=20),
         traj=p.trajectory, counters=L, tcur=p.tcur)
    # UniformCrossedEB: E != 0, static False
    pos = (0.3, -0.2, 0.1); vel = (1.0e5, 2.0e4, 1.0e4)
    p, L = run_particle(pos, vel, m_pr, e, rf.UniformCrossedEB(Ey=2.0, Bz=1e-4), 0.01)
    save("p_crossedeb", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=0.01,
         fieldprm=np.array([2.0, 1e-4]), params=parjson(), traj=p.trajectory, counters=L, tcur=p.tcur)
    # UniformBz
    p, L = run_particle(pos, vel, m_pr, e, rf.UniformBz(Bz=2e-4), 0.005)
    save("p_uniformbz", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=0.005,
         fieldprm=np.array([2e-4]), params=parjson(), traj=p.trajectory, counters=L, tcur=p.tcur)
    # Parabolic: the notebook's reference Particle, tolerances 1e-12 (crosses |z| = 1: quirk Q6)
    pos = (5, -5, 0.9); vel = (-0.1, 0.1, 0)
    p, L = run_particle(pos, vel, 1, 1, rf.Parabolic(), 60, solvertolerances=(1e-12, 1e-12))
    save("p_parabolic", pos=np.array(pos, float), vel=np.array(vel, float), mass=1.0, charge=1.0, delta=60.0,
         fieldprm=np.array([10.0, 1.0, 0.2]), params=parjson(solvertolerances=(1e-12, 1e-12)),
         traj=p.trajectory, counters=L, tcur=p.tcur)
    # enforce equatorial
    v = ru.speedfromKE(2e6, m_pr, 'ev')
    pos = (4 * Re, 0.5 * Re, 0.0); vel = (0.6 * v, -0.8 * v, 0.0)
    kw = {"enforce equatorial": True, "cyclotronresolution": 15}
    p, L = run_particle(pos, vel, m_pr, e, rf.EarthDipole(), 3, **kw)
    save("p_equatorial", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=3.0,
         params=parjson(**kw), traj=p.trajectory, counters=L, tcur=p.tcur)
    # user-defined field from examples/Creating new fields.ipynb cell 10 (NVRTC path)
    class ChargedDipole(rf._Field):
        def __init__(self, B0=1, Q=1):
            rf._Field.__init__(self)
            self.B0 = B0; self.Q = Q; self._k = 8.9875517873681764e9; self.static = False
        def B(self, tpos):
            t, x, y, z = tpos
            return self.B0 * np.array([3*x*z, 3*y*z, (2*z*z - x*x - y*y)]) / pow(x*x+y*y+z*z, 5.0/2.0)
        def E(self, tpos):
            t, x, y, z = tpos
            return self._k*self.Q * np.array([x, y, z]) / pow(x*x+y*y+z*z, 3.0/2.0)
    refshim.reset_params(rapt)
    p = rapt.Particle([5, 0, 0], [0, 1, 0], t0=0, mass=m_pr, charge=e, field=ChargedDipole(Q=1e-6))
    p.setke(1)
    pos0 = p.trajectory[0, 1:4].copy(); mom0 = p.trajectory[0, 4:].copy()
    gm = np.sqrt(m_pr ** 2 + mom0 @ mom0 / c ** 2)
    refshim.SOLVER_LOG.clear()
    cp, cr = p.cycper(), p.cycrad()
    p.advance(4e-4)
    save("p_chargeddipole", pos=pos0, vel=mom0 / gm, mom=mom0, mass=m_pr, charge=e, delta=4e-4,
         fieldprm=np.array([1.0, 1e-6, 8.9875517873681764e9]), params=parjson(),
         traj=p.trajectory, counters=take_log(), tcur=p.tcur, cycper=cp, cycrad=cr)


def bounce_setup(g):
    """Intermediate values of GuidingCenter.bounceperiod (GuidingCenter.py:593-606)."""
    from rapt.fieldline import Fieldline
    tpos, ppar = g.trajectory[-1, 0:4], g.trajectory[-1, 4]
    Bmag = g.field.magB(tpos)
    gamma = np.sqrt(1 + 2*g.mu*Bmag/(g.mass*c*c) + (ppar/(g.mass*c))**2)
    if gamma - 1 < 1e-6:
        p = np.sqrt(2*g.mass*g.mu*Bmag + ppar**2); v = p/g.mass; Bm = (p**2)/(2*g.mass*g.mu)
    else:
        p = g.mass*c*np.sqrt((gamma+1)*(gamma-1)); Bm = p**2/((p-ppar)*(p+ppar))*Bmag; v = p/g.mass/gamma
    fl = Fieldline(tpos, g.field, Bmax=Bm)
    ds = fl.ds
    fl.trace()
    return dict(bs_gamma=gamma, bs_Bm=Bm, bs_v=v, bs_ds=ds, bs_curvature=g.field.curvature(tpos),
                bs_curve=fl.curve.copy(), bs_B=fl.getB(),
                bs_halfpath=rapt.flutils.halfbouncepath(tpos, g.field, Bm),
                bs_period=g.bounceperiod())


def case_g2():
    """GuidingCenter notebook cell 5, advance(20); + bounce-period set-up (G4)."""
    refshim.reset_params(rapt)
    f = rf.DoubleDipole()
    v = ru.speedfromKE(1e5, m_el)
    pos = (0, -10 * Re, 0)
    g0 = rapt.GuidingCenter(pos=pos, v=v, pa=80, mass=m_el, charge=-e, field=f)
    bs = bounce_setup(g0)
    g, L = run_gc(pos, v, 80, m_el, -e, f, 20)
    save("g2_gc_doubledipole", pos=np.array(pos, float), v=v, pa=80.0, mass=m_el, charge=-e, delta=20.0,
         params=parjson(), traj=g.trajectory, counters=L, tcur=g.tcur, mu=g.mu, **bs)


def case_gcfields():
    # EarthDipole, 1 MeV e-, L=5, pa 60 (bounce period 0.3642159344345973)
    refshim.reset_params(rapt)
    f = rf.EarthDipole()
    v = ru.speedfromKE(1e6, m_el)
    pos = (5 * Re, 0, 0)
    g0 = rapt.GuidingCenter(pos=pos, v=v, pa=60, mass=m_el, charge=-e, field=f)
    bs = bounce_setup(g0)
    g, L = run_gc(pos, v, 60, m_el, -e, f, 3)
    save("gc_earthdipole", pos=np.array(pos, float), v=v, pa=60.0, mass=m_el, charge=-e, delta=3.0,
         params=parjson(), traj=g.trajectory, counters=L, tcur=g.tcur, mu=g.mu, **bs)
    # generic IC, three EOMs, fixed GCtimestep
    pos = (4.2 * Re, -2.9 * Re, 0.6 * Re)
    for eom in ("TaoChanBrizardEOM", "BrizardChanEOM", "NorthropTellerEOM"):
        g, L = run_gc(pos, v, 55, m_el, -e, rf.DoubleDipole(), 8, eom=eom, GCtimestep=0.1)
        save("gc_eom_" + eom[:-3].lower(), pos=np.array(pos), v=v, pa=55.0, mass=m_el, charge=-e, delta=8.0,
             params=parjson(GCtimestep=0.1), eom=np.array(eom), traj=g.trajectory, counters=L,
             tcur=g.tcur, mu=g.mu)
    # pa = 90 (Q9: vpar = 0 exactly) with bounce-period dt -> equatorial special case
    refshim.reset_params(rapt)
    pos = (-7.8 * Re, 0, 0)
    vp = ru.speedfromKE(1e5, m_pr)
    g0 = rapt.GuidingCenter(pos=pos, v=vp, pa=90, mass=m_pr, charge=e, field=rf.DoubleDipole())
    bs = bounce_setup(g0)
    g, L = run_gc(pos, vp, 90, m_pr, e, rf.DoubleDipole(), 100, solvertolerances=(1e-6, 1e-6),
                  bounceresolution=20)
    save("gc_pa90_equatorial", pos=np.array(pos, float), v=vp, pa=90.0, mass=m_pr, charge=e, delta=100.0,
         params=parjson(solvertolerances=(1e-6, 1e-6), bounceresolution=20), traj=g.trajectory,
         counters=L, tcur=g.tcur, mu=g.mu, **bs)
    # VarEarthDipole (static False: dbdt path, 15 B evaluations per RHS)
    pos = (4.5 * Re, 1.2 * Re, -0.3 * Re)
    g, L = run_gc(pos, v, 50, m_el, -e, rf.VarEarthDipole(0.1, 10), 2, t0=0.7, GCtimestep=0.05)
    save("gc_vardipole", pos=np.array(pos), v=v, pa=50.0, mass=m_el, charge=-e, delta=2.0, t0=0.7,
         fieldprm=np.array([0.1, 10.0]), params=parjson(GCtimestep=0.05), traj=g.trajectory,
         counters=L, tcur=g.tcur, mu=g.mu)
    # UniformCrossedEB: pure E x B drift
    g, L = run_gc((0.3, -0.2, 0.1), 1.0e5, 70, m_pr, e, rf.UniformCrossedEB(Ey=2.0, Bz=1e-4), 0.05,
                  GCtimestep=0.005)
    save("gc_crossedeb", pos=np.array((0.3, -0.2, 0.1)), v=1.0e5, pa=70.0, mass=m_pr, charge=e, delta=0.05,
         fieldprm=np.array([2.0, 1e-4]), params=parjson(GCtimestep=0.005), traj=g.trajectory,
         counters=L, tcur=g.tcur, mu=g.mu)
    # enforce equatorial
    kw = {"enforce equatorial": True, "GCtimestep": 0.5}
    g, L = run_gc((6 * Re, 1 * Re, 0.0), v, 90, m_el, -e, rf.DoubleDipole(), 20, **kw)
    save("gc_equatorial_enforced", pos=np.array((6 * Re, 1 * Re, 0.0)), v=v, pa=90.0, mass=m_el, charge=-e,
         delta=20.0, params=parjson(**kw), traj=g.trajectory, counters=L, tcur=g.tcur, mu=g.mu)


def adaptive_dump(a):
    modes, nrows, rows = [], [], []
    for seg in a.trajlist:
        is_p = seg.trajectory.shape[1] == 7
        modes.append(0 if is_p else 1)
        nrows.append(seg.trajectory.shape[0])
        r = np.zeros((seg.trajectory.shape[0], 8))
        r[:, :seg.trajectory.shape[1]] = seg.trajectory
        if not is_p:
            r[:, 5] = seg.mu
        rows.append(r)
    return dict(seg_mode=np.array(modes), seg_nrows=np.array(nrows), rows=np.vstack(rows),
                seg_tcur=np.array([s.tcur for s in a.trajlist]))


def run_adaptive(pos, vel, mass, charge, field, delta, **par):
    import io, contextlib
    refshim.reset_params(rapt, **par)
    refshim.SOLVER_LOG.clear()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = rapt.Adaptive(tuple(pos), tuple(vel), 0, mass=mass, charge=charge, field=field)
        a.advance(delta)
    return a, take_log(), buf.getvalue()


SPEISER = dict(solvertolerances=(1e-12, 1e-12), epss=0.02, Ptimestep=0.1, GCtimestep=1)


def case_g3():
    a, L, txt = run_adaptive((5, -5, 0.9), (-0.1, 0.1, 0), 1, 1, rf.Parabolic(), 300, **SPEISER)
    print(txt)
    save("g3_speiser", pos=np.array((5, -5, 0.9)), vel=np.array((-0.1, 0.1, 0.0)), mass=1.0, charge=1.0,
         delta=300.0, params=parjson(**SPEISER), counters=L, stdout=np.array(txt), **adaptive_dump(a))


def case_e4():
    n = 6
    ic = synth.config4_speiser(n)
    for i in range(1, n):
        pos = (ic["x"][i], ic["y"][i], ic["z"][i]); vel = (ic["vx"][i], ic["vy"][i], ic["vz"][i])
        a, L, txt = run_adaptive(pos, vel, 1, 1, rf.Parabolic(), 150, **SPEISER)
        save(f"e4_speiser_{i}", index=i, pos=np.array(pos), vel=np.array(vel), mass=1.0, charge=1.0, delta=150.0,
             params=parjson(**SPEISER), counters=L, stdout=np.array(txt), **adaptive_dump(a))


def case_adaptive_dipole():
    """Adaptive in EarthDipole: adiabatic from the start -> begins as GuidingCenter
    (Adaptive.py:98-102), dt from the bounce period."""
    v = ru.speedfromKE(1e4, m_pr)
    pos = (4 * Re, 0.3 * Re, 0.2 * Re); vel = (0.2 * v, 0.5 * v, v * np.sqrt(1 - 0.04 - 0.25))
    a, L, txt = run_adaptive(pos, vel, m_pr, e, rf.EarthDipole(), 30)
    save("adaptive_dipole", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=30.0,
         params=parjson(), counters=L, stdout=np.array(txt), **adaptive_dump(a))


def case_e2():
    """First 32 protons of config 2, advance(1.0): rows, totals, final state."""
    n = 32
    ic = synth.config2_protons(n)
    fin, nrows, tot, tcur = [], [], [], []
    trajs = {}
    f = rf.EarthDipole()
    for i in range(n):
        pos = (ic["x"][i], ic["y"][i], ic["z"][i]); vel = (ic["vx"][i], ic["vy"][i], ic["vz"][i])
        p, L = run_particle(pos, vel, m_pr, e, f, 1.0, cyclotronresolution=20)
        fin.append(p.trajectory[-1]); nrows.append(p.trajectory.shape[0]); tot.append(L.sum(0)); tcur.append(p.tcur)
        if i < 4:
            trajs[f"traj{i}"] = p.trajectory; trajs[f"counters{i}"] = L
    save("e2_config2_first32", n=n, seed=20260201, delta=1.0, params=parjson(cyclotronresolution=20),
         final=np.array(fin), nrows=np.array(nrows), totals=np.array(tot), tcur=np.array(tcur), **trajs)


def case_e3():
    """First 16 electrons of config 3 (GC, DoubleDipole): GCtimestep=0.1 for 10 s, and the
    reference's bounce periods (dt parity)."""
    n = 16
    ic = synth.config3_electrons(n)
    f = rf.DoubleDipole()
    fin, nrows, tot, mu, bp = [], [], [], [], []
    for i in range(n):
        pos = (ic["x"][i], ic["y"][i], ic["z"][i])
        g, L = run_gc(pos, ic["v"][i], ic["pa"][i], m_el, -e, f, 10.0, GCtimestep=0.1)
        fin.append(g.trajectory[-1]); nrows.append(g.trajectory.shape[0]); tot.append(L.sum(0)); mu.append(g.mu)
        refshim.reset_params(rapt)
        g0 = rapt.GuidingCenter(pos=pos, v=ic["v"][i], pa=ic["pa"][i], mass=m_el, charge=-e, field=f)
        bp.append(g0.bounceperiod())
    save("e3_config3_first16", n=n, seed=20260301, delta=10.0, params=parjson(GCtimestep=0.1),
         final=np.array(fin), nrows=np.array(nrows), totals=np.array(tot), mu=np.array(mu),
         bounceperiod=np.array(bp))


def case_e5():
    n = 16
    ic = synth.config5_belt(n)
    f = rf.VarEarthDipole(0.1, 10)
    fin, nrows, tot, mu = [], [], [], []
    for i in range(n):
        pos = (ic["x"][i], ic["y"][i], ic["z"][i])
        g, L = run_gc(pos, ic["v"][i], ic["pa"][i], m_el, -e, f, 2.0, GCtimestep=0.05)
        fin.append(g.trajectory[-1]); nrows.append(g.trajectory.shape[0]); tot.append(L.sum(0)); mu.append(g.mu)
    save("e5_config5_first16", n=n, seed=20260501, delta=2.0, params=parjson(GCtimestep=0.05),
         final=np.array(fin), nrows=np.array(nrows), totals=np.array(tot), mu=np.array(mu))


def case_eye():
    """GuidingCenter.geteye (GuidingCenter.py:608-624 -> flutils.eye, flutils.py:65-151).
    pa 80: equatorial pitch angle >= 70 degrees, the spline/brentq/quad branch, UNMODIFIED reference.
    pa 45: the Simpson branch, which raises NameError in the unmodified reference (`simps`, flutils.py:130,
    is never imported; scipy.integrate.simpson is).  For that fixture only, the missing name is bound in the
    imported module's namespace to simpson(y, x=x) -- no reference file is changed -- and the fixture says so."""
    from rapt import flutils as rfu
    from rapt.fieldline import Fieldline
    from scipy.integrate import simpson
    f = rf.EarthDipole()
    v = ru.speedfromKE(1e6, m_el)
    for pa, name in ((80, "eye_pa80"), (45, "eye_pa45_simpson")):
        refshim.reset_params(rapt, GCtimestep=0.05)
        g = rapt.GuidingCenter(pos=(5 * Re, 0, 0), v=v, pa=pa, mass=m_el, charge=-e, field=f)
        g.advance(0.4)
        patched = False
        try:
            out = g.geteye(step=3)
        except NameError:
            rfu.simps = lambda y, x: simpson(y, x=x)
            patched = True
            out = g.geteye(step=3)
            del rfu.simps
        Bm = g.getBm()[::3]
        rows = g.trajectory[::3]
        curves = []
        for row, bm in zip(rows, Bm):
            fl = Fieldline(row[:4], f, Bmax=bm); fl.trace()
            curves.append(np.column_stack([fl.gets(), fl.getB()]))
        k = max(len(cv) for cv in curves)
        cur = np.full((len(curves), k, 2), np.nan)
        for i, cv in enumerate(curves):
            cur[i, :len(cv)] = cv
        save(name, pos=np.array((5 * Re, 0, 0)), v=v, pa=float(pa), mass=m_el, charge=-e, delta=0.4,
             params=parjson(GCtimestep=0.05), traj=g.trajectory, mu=g.mu, step=3, Bm=Bm, eye=out, curves=cur,
             npts=np.array([len(cv) for cv in curves]), simps_name_bound=patched)


def synth_grid_class():
    class SynthGrid(rf.Grid):
        """fields.Grid with the synthetic parser of rapt_b200/synth.py ("file name" = time index)."""
        def parsefile(self, filename):
            return synth.dipole_grid_slice(int(filename))
    return SynthGrid


def grid_checksum(files):
    import hashlib
    h = hashlib.sha256()
    for fn in files:
        g = synth.dipole_grid_slice(int(fn))
        for k in ("x", "y", "z", "Bx", "By", "Bz", "Ex", "Ey", "Ez"):
            h.update(np.ascontiguousarray(g[k]).tobytes())
    return np.array(h.hexdigest())


def case_grid():
    """fields.Grid (fields.py:513-814), UNMODIFIED reference, synthetic data files (4 time points, so the
    rolling three-point window is updated once during the advance, fields.py:697-705 and :737-738)."""
    SynthGrid = synth_grid_class()
    files = ["0", "1", "2", "3"]
    rng = np.random.default_rng(11)
    # field operators at points with ascending time (the window forgets earlier times once it moves)
    f = SynthGrid(files)
    npt = 10
    pts = np.column_stack([np.sort(rng.uniform(0, 2.9, npt)), rng.uniform(3.3, 7.7, npt) * Re,
                           rng.uniform(-1.7, 1.7, npt) * Re, rng.uniform(-2.7, 2.7, npt) * Re])
    pts[3, 1] = f.Bxt_interp.grid[1][5]          # exactly on a node plane
    rec = {k: [] for k in ("B", "E", "magB", "unitb", "gradB", "curlb", "lengthscale")}
    for tp in pts:
        rec["B"].append(f.B(tp)); rec["E"].append(f.E(tp)); rec["magB"].append(f.magB(tp))
        rec["unitb"].append(f.unitb(tp)); rec["gradB"].append(f.gradB(tp)); rec["curlb"].append(f.curlb(tp))
        rec["lengthscale"].append(f.lengthscale(tp))
    out = {"pts": pts, "gradstep": f.gradientstepsize, "static": f.static}
    out.update({k: np.array(v, dtype=float) for k, v in rec.items()})
    # Particle: 1 MeV proton, crosses t = 1.5 s (window update)
    v = ru.speedfromKE(1e6, m_pr); pa = 40 * np.pi / 180
    pos = (6 * Re, 0, 0); vel = (0, -v * np.sin(pa), v * np.cos(pa))
    p, L = run_particle(pos, vel, m_pr, e, SynthGrid(files), 2.6, cyclotronresolution=10)
    # GuidingCenter: 1 MeV electron, pa 60
    ve = ru.speedfromKE(1e6, m_el)
    g, Lg = run_gc(pos, ve, 60, m_el, -e, SynthGrid(files), 2.6, GCtimestep=0.05)
    # Sensitivity band of that guiding-centre run, measured with the reference itself: on a multilinear interpolant
    # grad|B| and curl b are piecewise constant, so the ODE is discontinuous at every cell face and the result depends on
    # the step sequence.  Eight reruns with the start position moved by a few ulp (relative 1e-15): the spread of the
    # reference's own trajectories / step counts is the resolution at which this case can be compared at all.
    prng = np.random.default_rng(12)
    band_pos, band_pp, band_ns = [], [], []
    for k in range(8):
        pk = tuple(np.array(pos) * (1 + 1e-15 * prng.choice([-3, -2, -1, 1, 2, 3], 3)) + np.array([0, 1e-9, 1e-9]) * prng.uniform(-1, 1, 3))
        gk, Lk = run_gc(pk, ve, 60, m_el, -e, SynthGrid(files), 2.6, GCtimestep=0.05)
        assert gk.trajectory.shape == g.trajectory.shape
        band_pos.append(np.max(np.linalg.norm(gk.trajectory[:, 1:4] - g.trajectory[:, 1:4], axis=1) / np.linalg.norm(g.trajectory[:, 1:4], axis=1)))
        band_pp.append(np.max(np.abs(gk.trajectory[:, 4] - g.trajectory[:, 4])) / np.max(np.abs(g.trajectory[:, 4])))
        band_ns.append(int(Lk[:, 1].sum()))
    print("  grid GC sensitivity band:", band_pos, band_pp, band_ns, int(Lg[:, 1].sum()))
    # leaving the grid: ValueError, rows before it are kept (Particle.py:304-307)
    refshim.reset_params(rapt, cyclotronresolution=10)
    pe = rapt.Particle(pos=(7.9 * Re, 0, 0), vel=(v * 0.6, 0, v * 0.8), t0=0, mass=m_pr, charge=e, field=SynthGrid(files))
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pe.advance(5.0)
        raised = False
    except ValueError:
        raised = True
    save("grid_synthetic", files=np.array(files), checksum=grid_checksum(files),
         p_pos=np.array(pos, float), p_vel=np.array(vel), p_mass=m_pr, p_charge=e, p_delta=2.6,
         p_params=parjson(cyclotronresolution=10), p_traj=p.trajectory, p_counters=L, p_tcur=p.tcur,
         g_pos=np.array(pos, float), g_v=ve, g_pa=60.0, g_mass=m_el, g_charge=-e, g_delta=2.6,
         g_params=parjson(GCtimestep=0.05), g_traj=g.trajectory, g_counters=Lg, g_tcur=g.tcur, g_mu=g.mu,
         g_band_pos=np.array(band_pos), g_band_ppar=np.array(band_pp), g_band_nstep=np.array(band_ns),
         oob_pos=np.array((7.9 * Re, 0, 0)), oob_vel=np.array((v * 0.6, 0, v * 0.8)), oob_raised=raised,
         oob_traj=pe.trajectory, **{"ops_" + k: v_ for k, v_ in out.items()})


def case_units():
    """Field operators and utils helpers at seeded points (fields.py:76-280, utils.py:29-433)."""
    rng = np.random.default_rng(7)
    flds = {
        "earthdipole": (rf.EarthDipole(), Re), "doubledipole": (rf.DoubleDipole(), Re),
        "uniformbz": (rf.UniformBz(2e-4), 1.0), "crossedeb": (rf.UniformCrossedEB(2.0, 1e-4), 1.0),
        "vardipole": (rf.VarEarthDipole(0.1, 10), Re), "parabolic": (rf.Parabolic(), 1.0),
    }
    out = {}
    for name, (f, scale) in flds.items():
        npt = 12
        if name == "parabolic":
            pts = np.column_stack([rng.uniform(0, 5, npt), rng.uniform(-6, 6, npt), rng.uniform(-6, 6, npt),
                                   rng.uniform(-1.6, 1.6, npt)])
        else:
            pts = np.column_stack([rng.uniform(0, 20, npt), rng.uniform(-7, 7, npt) * scale,
                                   rng.uniform(-7, 7, npt) * scale, rng.uniform(-3, 3, npt) * scale])
        rec = {k: [] for k in ("B", "E", "unitb", "magB", "gradB", "jacobianB", "curlb", "curvature",
                               "dBdt", "dbdt", "lengthscale", "timescale")}
        for tp in pts:
            rec["B"].append(f.B(tp)); rec["E"].append(np.asarray(f.E(tp), float)); rec["unitb"].append(f.unitb(tp))
            rec["magB"].append(f.magB(tp)); rec["gradB"].append(f.gradB(tp)); rec["jacobianB"].append(f.jacobianB(tp))
            rec["curlb"].append(f.curlb(tp)); rec["curvature"].append(f.curvature(tp))
            rec["dBdt"].append(float(f.dBdt(tp))); rec["dbdt"].append(np.zeros(3) + f.dbdt(tp))
            with np.errstate(divide="ignore"):
                rec["lengthscale"].append(f.lengthscale(tp))
                ts = f.timescale(tp)
            rec["timescale"].append(np.nan if ts is None else ts)
        out[name + "_pts"] = pts
        for k, v in rec.items():
            out[f"{name}_{k}"] = np.array(v, dtype=float)
    # utils helpers in DoubleDipole
    f = rf.DoubleDipole()
    npt = 12
    pos = np.column_stack([rng.uniform(-8, 6, npt), rng.uniform(-8, 8, npt), rng.uniform(-2, 2, npt)]) * Re
    spd = [ru.speedfromKE(k, m_pr) for k in 10 ** rng.uniform(4, 6.5, npt)]
    dirs = rng.normal(size=(npt, 3)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    vel = dirs * np.array(spd)[:, None]
    u = {k: [] for k in ("cycper", "cycrad", "gc_R", "gc_vp", "gc_v", "mu", "fp_pos", "fp_vel", "cycper2", "cycrad2")}
    for i in range(npt):
        t = 0.0
        u["cycper"].append(ru.cyclotron_period(t, pos[i], vel[i], f, m_pr, e))
        u["cycrad"].append(ru.cyclotron_radius(t, pos[i], vel[i], f, m_pr, e))
        R, vp, vv = ru.guidingcenter(t, pos[i], vel[i], f, m_pr, e)
        u["gc_R"].append(R); u["gc_vp"].append(vp); u["gc_v"].append(vv)
        u["mu"].append(ru.magnetic_moment(t, R, vp, vv, f, m_pr))
        pp, vv2 = ru.GCtoFP(t, R, vp, vv, f, m_pr, e, 0)
        u["fp_pos"].append(pp); u["fp_vel"].append(vv2)
        u["cycper2"].append(ru.cyclotron_period2(t, R, vv, f, m_pr, e))
        u["cycrad2"].append(ru.cyclotron_radius2(t, R, vp, vv, f, m_pr, e))
    out["utils_pos"] = pos; out["utils_vel"] = vel
    for k, v in u.items():
        out["utils_" + k] = np.array(v, dtype=float)
    out["getperp_in"] = np.array([[0, 1, 2.], [1, 0, 2.], [1, 2, 0.], [1, 2, 3.], [-2, 0.5, 1e-3]])
    out["getperp_out"] = np.array([np.asarray(ru.getperp(v), float) for v in out["getperp_in"]])
    kes = np.array([1.0, 1e3, 1e5, 1e6, 1e7, 5e8])
    out["speed_ke"] = kes
    out["speed_pr"] = np.array([ru.speedfromKE(k, m_pr) for k in kes])
    out["speed_el"] = np.array([ru.speedfromKE(k, m_el) for k in kes])
    # rkf.py on a small test ODE
    from rapt.rkf import rkf
    T, X = rkf(lambda x, t: np.array([x[1], -x[0] * (1 + 4 * np.sin(t) ** 2)]), 0.0, 6.0,
               np.array([1.0, 0.0]), 1e-6, 0.5, 1e-6)
    out["rkf_T"] = T; out["rkf_X"] = X
    save("units", **out)


def case_bc():
    """BounceCenter.advance (BounceCenter.py:206-251) and the flutils integrals behind it (flutils.py:65-316), from
    the UNMODIFIED reference.  Pitch angles are chosen so that the equatorial pitch angle is >= 70 degrees: below
    that the reference stops with NameError (`simps`, flutils.py:130).  NB the constructor feeds the pitch angle in
    degrees to a radian cosine (BounceCenter.py:114): pa = 80 means cos(80 rad) = -0.110, i.e. 96.3 degrees."""
    from rapt import flutils as rfu
    from scipy.integrate import simpson
    cases = [
        ("bc_dipole_electron", rf.EarthDipole(), "EarthDipole", (6 * Re, 0, 0.1 * Re), ru.speedfromKE(1e6, m_el), 80, m_el, -e, 0.3, 0.2),
        ("bc_dipole_proton", rf.EarthDipole(), "EarthDipole", (3 * Re, 2.5 * Re, -0.15 * Re), ru.speedfromKE(1e7, m_pr), 30, m_pr, e, 0.8, 0.0),
        ("bc_doubledipole_electron", rf.DoubleDipole(), "DoubleDipole", (-5 * Re, 4 * Re, 0.2 * Re), ru.speedfromKE(3e5, m_el), 80, m_el, -e, 0.35, 0.0),
    ]
    for name, f, fname, pos, v, pa, mass, charge, delta, delta2 in cases:
        refshim.reset_params(rapt)
        refshim.SOLVER_LOG.clear()
        b = rapt.BounceCenter(pos=pos, v=v, t0=0, pa=pa, mass=mass, charge=charge, field=f)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b.advance(delta)
            n1 = len(b.trajectory)
            if delta2:
                b.advance(delta2)            # second call: restarts from the last row LABEL (BounceCenter.py:247,250)
        log = take_log()
        gamma = 1.0 / np.sqrt(1 - (v / c) ** 2)
        Bm = mass * gamma ** 2 * v ** 2 / (2 * b.mu)
        # the pieces at the first, a middle and the last row
        pts = b.trajectory[[0, len(b.trajectory) // 2, -1]]
        Sb = np.array([rfu.halfbouncepath(r, f, Bm) for r in pts])
        I = np.array([rfu.eye(r, f, Bm) for r in pts])
        gI = np.array([rfu.gradI(r, f, Bm) for r in pts])
        save(name, field=fname, pos=np.array(pos), v=v, pa=float(pa), mass=mass, charge=charge, delta=delta, delta2=delta2,
             nrows_first_call=n1, traj=b.trajectory, mu=b.mu, Bm=Bm, solver_log=log, pts=pts, Sb=Sb, I=I, gradI=gI)
    # flutils on its own, including the cases BounceCenter cannot reach in the reference:
    #  * forward/backward differences in gradI when a displaced field line lies beyond the mirror field (flutils.py:212-215)
    #  * the Simpson branch (eqpa < 70) with the missing name bound to scipy's simpson (fixture says so)
    f = rf.EarthDipole()
    refshim.reset_params(rapt)
    rows = []
    for L, zz, eqpa in ((5.0, 0.0, 89.0), (5.0, 0.05, 88.0), (4.0, 0.3, 75.0), (6.5, -0.2, 80.0)):
        tpos = np.array([0.0, L * Re * np.cos(0.3), L * Re * np.sin(0.3), zz * Re])
        Bm = f.magB(np.array([0.0, L * Re * np.cos(0.3), L * Re * np.sin(0.3), 0.0])) / np.sin(np.radians(eqpa)) ** 2
        rows.append((tpos, Bm, rfu.halfbouncepath(tpos, f, Bm), rfu.eye(tpos, f, Bm), rfu.gradI(tpos, f, Bm), 0))
    rfu.simps = lambda y, x: simpson(y, x=x)
    for L, zz, eqpa in ((5.0, 0.2, 45.0), (3.0, -0.1, 60.0)):
        tpos = np.array([0.0, L * Re * np.cos(1.3), L * Re * np.sin(1.3), zz * Re])
        Bm = f.magB(np.array([0.0, L * Re * np.cos(1.3), L * Re * np.sin(1.3), 0.0])) / np.sin(np.radians(eqpa)) ** 2
        rows.append((tpos, Bm, rfu.halfbouncepath(tpos, f, Bm), rfu.eye(tpos, f, Bm), rfu.gradI(tpos, f, Bm), 1))
    del rfu.simps
    save("bc_flutils", field="EarthDipole", tpos=np.array([r[0] for r in rows]), Bm=np.array([r[1] for r in rows]),
         Sb=np.array([r[2] for r in rows]), I=np.array([r[3] for r in rows]), gradI=np.array([r[4] for r in rows]),
         simps_name_bound=np.array([r[5] for r in rows]))


def case_fail():
    """Solver failure inside advance(): scipy's nsteps=500 limit ends the `while r.successful()` loop AFTER the row of the
    failed call has been appended (Particle.py:304-307 -- label = the row's end time, values = the state reached,
    tcur = time reached + dt; GuidingCenter.py:452-456 -- label = tcur = the time reached)."""
    v = ru.speedfromKE(1e6, m_pr); pa = 30 * np.pi / 180
    pos = (6 * Re, 0, 0); vel = (0, -v * np.sin(pa), v * np.cos(pa))
    par = dict(cyclotronresolution=20, solvertolerances=(1e-15, 1e-30))
    p, L = run_particle(pos, vel, m_pr, e, rf.EarthDipole(), 10, **par)
    assert p.trajectory.shape[0] == 2
    save("p_fail_nmax", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=10.0, params=parjson(**par),
         traj=p.trajectory, counters=L, tcur=p.tcur)
    ve = ru.speedfromKE(1e6, m_el)
    par = dict(GCtimestep=50.0)
    g, Lg = run_gc(pos, ve, 30, m_el, -e, rf.EarthDipole(), 100, **par)
    assert g.trajectory.shape[0] == 2
    save("gc_fail_nmax", pos=np.array(pos), v=ve, pa=30.0, mass=m_el, charge=-e, delta=100.0, params=parjson(**par),
         traj=g.trajectory, counters=Lg, tcur=g.tcur, mu=g.mu)


def case_e2long():
    """First 32 protons of config 2 at the BENCH horizon, advance(10.0) (BASELINE.json configs[1]): final state, row
    count, (nfcn, nstep, naccpt, nrejct) totals per proton.  ~3e3 steps per proton-second."""
    n = 32
    ic = synth.config2_protons(n)
    fin, nrows, tot, tcur = [], [], [], []
    f = rf.EarthDipole()
    for i in range(n):
        pos = (ic["x"][i], ic["y"][i], ic["z"][i]); vel = (ic["vx"][i], ic["vy"][i], ic["vz"][i])
        p, L = run_particle(pos, vel, m_pr, e, f, 10.0, cyclotronresolution=20)
        fin.append(p.trajectory[-1]); nrows.append(p.trajectory.shape[0]); tot.append(L.sum(0)); tcur.append(p.tcur)
    save("e2_config2_first32_10s", n=n, seed=20260201, delta=10.0, params=parjson(cyclotronresolution=20),
         final=np.array(fin), nrows=np.array(nrows), totals=np.array(tot), tcur=np.array(tcur))


def case_e2fail():
    """The one member of the 1 M-proton headline ensemble (BASELINE.json configs[1]) whose row loop ends on scipy's
    nsteps = 500 limit: member 408359 of synth.config2_protons(1 << 20) (bench.py reports it as solver_failures: 1,
    failed_members_rank0: [408359]).  The reference on the same proton: rows until the failed call, its appended row,
    the totals."""
    i = 408359
    ic = synth.config2_protons(1 << 20)
    pos = (ic["x"][i], ic["y"][i], ic["z"][i]); vel = (ic["vx"][i], ic["vy"][i], ic["vz"][i])
    refshim.reset_params(rapt, cyclotronresolution=20)
    refshim.SOLVER_LOG.clear()
    p = rapt.Particle(pos=pos, vel=vel, t0=0, mass=m_pr, charge=e, field=rf.EarthDipole())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p.advance(10.0)
    L = take_log()
    msgs = [str(x.message) for x in w]
    print("  rows", p.trajectory.shape, "tcur", p.tcur, "last call", L[-1], "warnings", msgs[-1:] )
    save("e2_config2_member408359", member=i, n_total=1 << 20, seed=20260201, delta=10.0, params=parjson(cyclotronresolution=20),
         pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, nrows=p.trajectory.shape[0], last_rows=p.trajectory[-3:],
         totals=L.sum(0), last_call=L[-1], tcur=p.tcur, warned=np.array(any("nsteps" in m for m in msgs)), ke_ev=ic["ke_ev"][i])


def case_g1long():
    """A few hundred gyroperiods (north_star's short horizon): the g1b proton for 40 s = ~325 gyroperiods, 6.5e3 rows.
    Stored: every 16th row + the last, per-call counters, totals."""
    v = ru.speedfromKE(1e6, m_pr, 'ev')
    pos = (3.1 * Re, 2.3 * Re, 0.7 * Re)
    vel = (v * 0.31, -v * 0.52, v * np.sqrt(1 - 0.31 ** 2 - 0.52 ** 2))
    p, L = run_particle(pos, vel, m_pr, e, rf.EarthDipole(), 40, cyclotronresolution=20)
    tr = p.trajectory
    save("g1c_325_gyroperiods", pos=np.array(pos), vel=np.array(vel), mass=m_pr, charge=e, delta=40.0,
         params=parjson(cyclotronresolution=20), nrows=tr.shape[0], every=16, traj_dec=tr[::16], last=tr[-1],
         counters=L.astype(np.int16), tcur=p.tcur,
         gyroperiods=40.0 / ru.cyclotron_period(0.0, np.array(pos), np.array(vel), rf.EarthDipole(), m_pr, e))


def case_getters():
    """N2 getters over stored trajectories: Particle.guidingcenter / mu (Particle.py:463-482) and GuidingCenter.getB /
    getgamma / getBm / getke (GuidingCenter.py:486-591), UNMODIFIED reference, on the g1b proton (1 s, DoubleDipole and
    EarthDipole) and the g2-type electron."""
    v = ru.speedfromKE(1e6, m_pr, 'ev')
    pos = (3.1 * Re, 2.3 * Re, 0.7 * Re)
    vel = (v * 0.31, -v * 0.52, v * np.sqrt(1 - 0.31 ** 2 - 0.52 ** 2))
    out = {}
    for tag, f in (("ed", rf.EarthDipole()), ("dd", rf.DoubleDipole())):
        p, _ = run_particle(pos, vel, m_pr, e, f, 1.0, cyclotronresolution=20)
        out[f"p_{tag}_traj"] = p.trajectory
        out[f"p_{tag}_gc"] = p.guidingcenter()
        out[f"p_{tag}_mu"] = p.mu()
        out[f"p_{tag}_cycrad"] = p.cycrad(); out[f"p_{tag}_cycper"] = p.cycper()
    ve = ru.speedfromKE(1e6, m_el)
    for tag, f, ke in (("dd", rf.DoubleDipole(), 1e6), ("ed_nr", rf.EarthDipole(), 0.3)):     # relativistic / gamma-1 < 1e-6
        vv = ru.speedfromKE(ke, m_el)
        g, _ = run_gc((6 * Re, 0.4 * Re, 0.3 * Re), vv, 40, m_el, -e, f, 3.0, GCtimestep=0.05)
        out[f"g_{tag}_traj"] = g.trajectory; out[f"g_{tag}_mu"] = g.mu; out[f"g_{tag}_v"] = vv
        out[f"g_{tag}_B"] = g.getB(); out[f"g_{tag}_gamma"] = g.getgamma(); out[f"g_{tag}_Bm"] = g.getBm()
        out[f"g_{tag}_ke"] = g.getke(); out[f"g_{tag}_cycrad"] = g.cycrad()
    save("getters", pos=np.array(pos), vel=np.array(vel), mass_p=m_pr, mass_e=m_el, charge=e, **out)


CASES = {
    "e2long": case_e2long, "e2fail": case_e2fail, "g1long": case_g1long, "getters": case_getters,
    "fail": case_fail,
    "g1": case_g1, "g1b": case_g1b, "pfields": case_pfields, "g2": case_g2, "gcfields": case_gcfields,
    "g3": case_g3, "e4": case_e4, "adip": case_adaptive_dipole, "e2": case_e2, "e3": case_e3,
    "e5": case_e5, "units": case_units, "eye": case_eye, "grid": case_grid, "bc": case_bc,
}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    sel = sys.argv[1:] or list(CASES)
    for k in sel:
        t = time.time()
        print(f"[{k}]")
        CASES[k]()
        print(f"  {time.time()-t:.1f} s")
    with open(os.path.join(GOLD, "VERSIONS.txt"), "w") as fh:
        import scipy
        fh.write(f"numpy {np.__version__}\nscipy {scipy.__version__}\npython {sys.version.split()[0]}\n"
                 "reference mkozturk/rapt at /root/reference (unmodified; shims in oracle/refshim.py)\n")
