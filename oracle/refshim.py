"""Import the UNMODIFIED reference (mkozturk/rapt, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py to produce tests/golden/*.npz and by
oracle/refbench.py, the timing of the reference's own CPU path that bench.py reports as its baseline
(`cpu_baseline`, `--impl reference`).  /root/reference does not exist on the GPU box; there the copy that
`__graft_entry__.build()` pip-installed into oracle/_ref/ is imported.  No test reads either at run time.

Two non-invasive shims (no reference file is modified), see SURVEY.md §8(c):
  1. scipy.misc.derivative is gone in scipy >= 1.12 but rapt/flutils.py:19 imports it at
     package import -> inject a stub (only criticalpoints, flutils.py:61, calls it).
  2. numpy >= 2 raises on `ndarray == []` (Particle.py:105, GuidingCenter.py:123) -> wrap the
     two __init__s so ndarray pos/vel arguments arrive as tuples.
It also instruments scipy.integrate._ode.dopri5.run (shared by dop853) to log the solver
counters (nfcn, nstep, naccpt, nrejct = iwork[16:20]) of every solver call.
"""
import sys, types
import numpy as np

import os
_HERE = os.path.dirname(os.path.abspath(__file__))
# the source tree in the build container; on the GPU box only the copy that __graft_entry__.build() installed with pip
# into oracle/_ref/ (git-ignored build output, travels with the snapshot) exists
REFERENCE_PATH = "/root/reference" if os.path.isdir("/root/reference/rapt") else os.path.join(_HERE, "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_PATH, "rapt"))
SOLVER_LOG = []          # one (nfcn, nstep, naccpt, nrejct) tuple per r.integrate() call


def load_reference():
    import scipy.misc
    if not hasattr(scipy.misc, "derivative"):
        def _derivative(*a, **k):
            raise NotImplementedError("scipy.misc.derivative stub (off the hot path)")
        scipy.misc.derivative = _derivative
    if REFERENCE_PATH not in sys.path:
        sys.path.insert(0, REFERENCE_PATH)
    import rapt

    def _tup(a):
        return tuple(a.tolist()) if isinstance(a, np.ndarray) else a

    if not getattr(rapt.Particle, "_shimmed", False):
        p_init = rapt.Particle.__init__

        def particle_init(self, pos=[], vel=[], t0=0, mass=None, charge=None, field=None):
            p_init(self, _tup(pos), _tup(vel), t0, mass, charge, field)
        rapt.Particle.__init__ = particle_init
        rapt.Particle._shimmed = True

        g_init = rapt.GuidingCenter.__init__

        def gc_init(self, pos=[], v=0, pa=None, ppar=None, t0=0, mass=None, charge=None, field=None):
            g_init(self, _tup(pos), v, pa, ppar, t0, mass, charge, field)
        rapt.GuidingCenter.__init__ = gc_init

        from scipy.integrate import _ode
        run0 = _ode.dopri5.run

        def run(self, *a, **k):
            out = run0(self, *a, **k)
            SOLVER_LOG.append(tuple(int(v) for v in self.iwork[16:20]))
            return out
        _ode.dopri5.run = run
    return rapt


def default_params():
    return {
        "cyclotronresolution": 10, "Ptimestep": 0, "bounceresolution": 10, "GCtimestep": 0,
        "BCtimestep": 0.1, "solvertolerances": (1.49012e-8, 1.49012e-8),
        "fieldlineresolution": 50, "flsolver": "rkf", "eyegradientstep": 0.03 * 6378137,
        "epss": 5e-2, "epst": 5e-2, "enforce equatorial": False,
    }


def reset_params(rapt, **over):
    rapt.params.update(default_params())
    rapt.params.update(over)
