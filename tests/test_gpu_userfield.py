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
