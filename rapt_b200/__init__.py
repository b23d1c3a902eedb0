"""rapt_b200 -- a B200-native (sm_100a) engine for RAPT's particle-advance hot path.

Drop-in for the `.advance()` integrations of mkozturk/rapt's `Particle`, `GuidingCenter` and
`Adaptive` tracers: same constructors, `params`, getters and field-plugin interface
(reference: rapt/__init__.py:1-42), with the integration itself running as hand-written CUDA kernels
in librapt_b200.so (C ABI: include/rapt_b200.h).  Ensemble entry points (`rapt_b200.ensemble`) advance
millions of independent tracers per call.  There is no CPU fallback.
"""
# Constants (rapt/__init__.py:5-10)
e = 1.602176565e-19      # Elementary charge (Coulomb)
m_pr = 1.672621777e-27   # Proton mass (kg)
m_el = 9.10938291e-31    # Electron mass (kg)
c = 299792458            # speed of light (m/s)
B0 = 3.07e-5             # Earth field strength at magnetic equator (Tesla)
re = Re = 6378137        # Earth radius (meter)


# Mode-switch signalling (rapt/__init__.py:14-17)
class Adiabatic(Exception):
    pass


class NonAdiabatic(Exception):
    pass


# Parameters and defaults (rapt/__init__.py:21-34); read at every advance() call
params = {
    "cyclotronresolution": 10,
    "Ptimestep": 0,
    "bounceresolution": 10,
    "GCtimestep": 0,
    "BCtimestep": 0.1,
    "solvertolerances": (1.49012e-8, 1.49012e-8),
    "fieldlineresolution": 50,
    "flsolver": "rkf",
    "eyegradientstep": 0.03 * Re,
    "epss": 5e-2,
    "epst": 5e-2,
    "enforce equatorial": False,
    # engine-only knobs (no reference counterpart)
    "arith": "fast",            # "fast" (FMA, reciprocal multiplies) or "strict" (mirrors CPU operation order)
    "dop853_reject_rule": 0,    # 0: scipy 1.18.1 `_dop`; 1: Hairer's Fortran (scipy 1.3.1)
}

from . import utils, fields, engine          # noqa: E402
from .Particle import Particle               # noqa: E402
from .GuidingCenter import GuidingCenter     # noqa: E402
from .Adaptive import Adaptive               # noqa: E402
from .BounceCenter import BounceCenter       # noqa: E402
from .fieldline import Fieldline             # noqa: E402
from .ensemble import ParticleEnsemble, GuidingCenterEnsemble, AdaptiveEnsemble, BounceCenterEnsemble   # noqa: E402
from . import flutils                        # noqa: E402
from .flutils import eye, gradI, halfbouncepath   # noqa: E402  (rapt/__init__.py:42)
