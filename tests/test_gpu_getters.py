"""GPU parity of the getter / diagnostic kernels over stored trajectories against values the
UNMODIFIED reference returned for the same trajectories (tests/golden/getters.npz, oracle/gen_golden.py:case_getters):
Particle.guidingcenter / mu / cycrad / cycper (rapt/Particle.py:463-494) and GuidingCenter.getB / getgamma / getBm /
getke / cycrad (rapt/GuidingCenter.py:486-591).  The objects are given the reference's own trajectory, so the getters
are compared on identical rows; bar 1e-12 relative."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rapt_b200 as R
    from rapt_b200 import _lib
    _lib.init(0)
    old = dict(R.params)
    yield R
    R.params.clear(); R.params.update(old)


@pytest.fixture(scope="module")
def gold():
    return np.load(H.GOLDEN + "/getters.npz")


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("tag,fname", [("ed", "EarthDipole"), ("dd", "DoubleDipole")])
def test_particle_guidingcenter_and_mu_vs_reference(rb, gold, tag, fname, arith):
    rb.params["arith"] = arith
    p = rb.Particle(pos=tuple(gold["pos"]), vel=tuple(gold["vel"]), t0=0, mass=float(gold["mass_p"]),
                    charge=float(gold["charge"]), field=getattr(rb.fields, fname)())
    p.trajectory = gold[f"p_{tag}_traj"].copy()
    gc = p.guidingcenter()
    ref = gold[f"p_{tag}_gc"]
    assert gc.shape == ref.shape
    assert H.vec_relerr(gc[:, :3], ref[:, :3]) < 1e-12, "guiding-centre position (utils.guidingcenter fixed point)"
    # parallel speed changes sign along the bounce: relative to the particle speed
    assert np.max(np.abs(gc[:, 3] - ref[:, 3])) < 1e-10 * np.max(ref[:, 4])
    assert H.relerr(gc[:, 4], ref[:, 4]) < 1e-12
    assert H.relerr(p.mu(), gold[f"p_{tag}_mu"]) < 1e-10      # mu ~ v_perp^2 = (v - v_par)(v + v_par): cancellation near the equator
    assert p.cycrad() == pytest.approx(float(gold[f"p_{tag}_cycrad"]), rel=1e-12)
    assert p.cycper() == pytest.approx(float(gold[f"p_{tag}_cycper"]), rel=1e-12)
    assert H.relerr(p.getB(), [np.linalg.norm(b) for b in rb.engine.field_ops(p.field, p.trajectory[:, :4], which=["B"])["B"]]) < 1e-14


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("tag,fname", [("dd", "DoubleDipole"), ("ed_nr", "EarthDipole")])
def test_guiding_centre_getters_vs_reference(rb, gold, tag, fname, arith):
    """tag ed_nr is a 0.3 eV electron: gamma - 1 < 1e-6, the reference's non-relativistic branches."""
    rb.params["arith"] = arith
    traj = gold[f"g_{tag}_traj"]
    g = rb.GuidingCenter(pos=tuple(traj[0, 1:4]), v=float(gold[f"g_{tag}_v"]), pa=40, t0=0, mass=float(gold["mass_e"]),
                         charge=-float(gold["charge"]), field=getattr(rb.fields, fname)())
    assert g.mu == pytest.approx(float(gold[f"g_{tag}_mu"]), rel=1e-13)
    g.mu = float(gold[f"g_{tag}_mu"])
    g.trajectory = traj.copy()
    assert H.relerr(g.getB(), gold[f"g_{tag}_B"]) < 1e-13
    assert H.relerr(g.getgamma(), gold[f"g_{tag}_gamma"]) < 1e-13
    assert H.relerr(g.getBm(), gold[f"g_{tag}_Bm"]) < 1e-10       # 1 - (p_par/mc)^2/((g-1)(g+1)): cancellation at the mirror
    assert H.relerr(g.getke(), gold[f"g_{tag}_ke"]) < 1e-9 if tag == "dd" else H.relerr(g.getke(), gold[f"g_{tag}_ke"]) < 1e-12
    assert g.cycrad() == pytest.approx(float(gold[f"g_{tag}_cycrad"]), rel=1e-11)
    assert (tag == "ed_nr") == bool(np.all(gold[f"g_{tag}_gamma"] - 1 < 1e-6))
