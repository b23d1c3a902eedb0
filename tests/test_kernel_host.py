"""CPU run of the product's own advance kernels.

tests/hostcheck/kernel_host.cpp compiles rapt_b200/csrc/rapt_particle.cuh, rapt_particle_rkn.cuh and rapt_gc.cuh (fast
arithmetic flavour, the one bench.py times) for the host through a small shim; here those builds are compared with the
golden vectors generated from the reference and with the CPU oracle, with the same bars as the `-m gpu` parity tests
(tests/test_gpu_particle.py, tests/test_gpu_gc.py): positions / momenta <= 1e-8 relative, solver counters equal on the
non-chaotic cases.  This is a check of the kernel SOURCE (step control, HINIT, FSAL reuse, Nystrom-form stage sums, row
bookkeeping, work queue) that runs without a GPU; it is not a product path -- rapt_b200 never loads that library.
"""
import numpy as np
import pytest

import helpers as H
import hostkernel as K


@pytest.mark.parametrize("rkn", [True, False], ids=["k_particle_rkn", "k_particle_dop853"])
@pytest.mark.parametrize("name", list(H.PARTICLE_CASES))
def test_particle_kernel_source_vs_golden(name, rkn):
    d, par = H.load(name)
    fname, fargs = H.PARTICLE_CASES[name]
    field = H.gpu_field(fname, fargs)
    if rkn and (not field.static or par.get("enforce equatorial")):
        pytest.skip("the Nystrom-form kernel is only launched for static fields without the equatorial constraint")
    traj = d["traj"]
    o = K.particle_advance(field, traj[0], float(d["mass"]), float(d["charge"]), float(d["delta"]),
                           store_every=1, max_rows=len(traj) + 8, rkn=rkn, nthreads=1, **par)
    n = int(o["nstored"][0])
    assert o["status"][0] == 1
    assert o["nrows"][0] == len(traj) == n
    rows = o["rows"][0, :n]
    assert H.relerr(rows[:, 0], traj[:, 0]) < 1e-13
    chaotic = name in ("p_parabolic",)
    tol = 1e-8 if not chaotic else 1e-6
    assert H.vec_relerr(rows[:, 1:4], traj[:, 1:4]) < tol
    assert H.vec_relerr(rows[:, 4:7], traj[:, 4:7]) < (tol if not chaotic else 1e-5)
    ref = d["counters"].sum(0)
    if name == "g1_readme":
        assert abs(int(o["counters"][0, 1]) - int(ref[1])) <= 4      # zero-coordinate start, as in the GPU test
    elif not chaotic:
        assert tuple(o["counters"][0]) == tuple(ref), "(nfcn, nstep, naccpt, nrejct) must equal scipy's"
        assert np.array_equal(rows[1:, 7].astype(np.int64), np.cumsum(d["counters"][:, 1]))
    assert abs(o["tcur"][0] - float(d["tcur"])) <= 1e-12 * abs(float(d["tcur"]))


@pytest.mark.parametrize("rkn", [True, False], ids=["k_particle_rkn", "k_particle_dop853"])
def test_particle_kernel_source_config2_vs_oracle(rkn):
    """2048 protons of config 2, advance(0.25 s), decimated rows: kernel source on the host vs the oracle."""
    import oracle as O
    from rapt_b200 import engine, synth
    n = 2048
    ic = synth.config2_protons(n)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], engine.particle_momentum(vel, ic["mass"])])
    par = dict(cyclotronresolution=20)
    ref = O.particle_advance(O.make_field("EarthDipole"), O.make_params(**par), st, ic["mass"], ic["charge"], 0.25,
                             store_every=7, max_rows=64, nthreads=8)
    o = K.particle_advance(H.gpu_field("EarthDipole", ()), st, ic["mass"], ic["charge"], 0.25,
                           store_every=7, max_rows=64, rkn=rkn, nthreads=8, **par)
    assert np.array_equal(o["nrows"], ref["nrows"])
    assert np.array_equal(o["nstored"], ref["nstored"])
    assert np.all(o["status"] == 1)
    assert H.vec_relerr(o["state"][:, 1:4], ref["state"][:, 1:4]) < 1e-8
    assert H.vec_relerr(o["state"][:, 4:7], ref["state"][:, 4:7]) < 1e-8
    same = np.all(o["counters"] == ref["counters"], axis=1)
    assert same.mean() > 0.99, f"only {same.mean():.4f} of particles match counts"
    assert abs(int(o["counters"][:, 1].sum()) - int(ref["counters"][:, 1].sum())) <= 1e-4 * ref["counters"][:, 1].sum()
    for i in (0, 17, n - 1):
        k = int(o["nstored"][i])
        assert H.vec_relerr(o["rows"][i, :k, 1:4], ref["rows"][i, :k, 1:4]) < 1e-8
        assert H.relerr(o["rows"][i, :k, 0], ref["rows"][i, :k, 0]) < 1e-13


def test_particle_kernel_source_work_queue_is_order_independent():
    """One lane or eight lanes pulling from the queue: same bits (tracers do not interact)."""
    from rapt_b200 import engine, synth
    ic = synth.config2_protons(64)
    vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    st = np.column_stack([ic["t0"], ic["x"], ic["y"], ic["z"], engine.particle_momentum(vel, ic["mass"])])
    f = H.gpu_field("EarthDipole", ())
    a = K.particle_advance(f, st, ic["mass"], ic["charge"], 0.1, store_every=3, max_rows=32, nthreads=1, cyclotronresolution=20)
    b = K.particle_advance(f, st, ic["mass"], ic["charge"], 0.1, store_every=3, max_rows=32, nthreads=8, cyclotronresolution=20)
    for k in ("state", "nrows", "nstored", "counters", "status", "tcur", "dt"):
        assert np.array_equal(a[k], b[k]), k
    for i in range(64):
        assert np.array_equal(a["rows"][i, :a["nstored"][i]], b["rows"][i, :b["nstored"][i]])


@pytest.mark.parametrize("name", list(H.GC_CASES))
def test_gc_kernel_source_vs_golden(name):
    d, par = H.load(name)
    fname, fargs = H.GC_CASES[name]
    traj = d["traj"]
    mass, q, v = float(d["mass"]), float(d["charge"]), float(d["v"])
    st0 = traj[0, :5]
    # the reference's own output step (bounce period / bounceresolution, or GCtimestep)
    dt = float(d["bs_period"]) / par.get("bounceresolution", 10) if "bs_period" in d.files else par["GCtimestep"]
    eom = str(d["eom"]) if "eom" in d.files else "TaoChanBrizardEOM"
    gpar = {k: v_ for k, v_ in par.items() if k in ("solvertolerances", "enforce equatorial")}
    o = K.gc_advance(H.gpu_field(fname, fargs), st0, float(d["mu"]), v, mass, q, dt, float(d["delta"]), eom=eom,
                     store_every=1, max_rows=len(traj) + 8, nthreads=1, **gpar)
    n = int(o["nstored"][0])
    assert o["status"][0] == 1
    assert o["nrows"][0] == len(traj) == n
    rows = o["rows"][0, :n]
    assert H.relerr(rows[:, 0], traj[:, 0]) < 1e-13
    assert H.vec_relerr(rows[:, 1:4], traj[:, 1:4]) < 1e-8
    pscale = max(np.max(np.abs(traj[:, 4])), 1e-3 * mass * v)      # p_par stays exactly 0 for pa = 90 in the reference
    assert np.max(np.abs(rows[:, 4] - traj[:, 4])) / pscale < 1e-7
    ref = d["counters"].sum(0)
    if name not in ("gc_pa90_equatorial", "gc_equatorial_enforced", "g2_gc_doubledipole", "gc_earthdipole"):
        # (the four excluded cases start from an exact-zero coordinate / zero p_par: round-off dominated first steps,
        # tests/test_gpu_gc.py gives them the same allowance)
        assert abs(int(o["counters"][0, 1]) - int(ref[1])) <= max(2, 0.01 * ref[1])


def test_gc_kernel_source_config3_vs_oracle():
    import oracle as O
    from rapt_b200 import synth
    n = 256
    ic3 = synth.config3_electrons(n)
    pos = np.column_stack([ic3["x"], ic3["y"], ic3["z"]])
    of = O.make_field("DoubleDipole")
    con = O.gc_construct(of, ic3["t0"], pos, ic3["v"], ic3["pa"], ic3["mass"])
    ppar, mu = con
    st3 = np.column_stack([ic3["t0"], pos, ppar])
    r = O.gc_advance(of, O.make_params(), st3, mu, ic3["v"], ic3["mass"], ic3["charge"], 0.1, 2.0, store_every=0, nthreads=8)
    g = K.gc_advance(H.gpu_field("DoubleDipole", ()), st3, mu, ic3["v"], ic3["mass"], ic3["charge"], 0.1, 2.0,
                     store_every=0, nthreads=8)
    assert np.array_equal(g["nrows"], r["nrows"])
    assert H.vec_relerr(g["state"][:, 1:4], r["state"][:, 1:4]) < 1e-8
    assert np.max(np.abs(g["state"][:, 4] - r["state"][:, 4])) < 1e-8 * np.max(np.abs(r["state"][:, 4]))
    same = np.all(g["counters"] == r["counters"], axis=1)
    assert same.mean() > 0.98, f"only {same.mean():.4f} of guiding centres match counts"
