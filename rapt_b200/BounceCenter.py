"""BounceCenter: API surface only (rapt/BounceCenter.py:74-316).

The bounce-averaged drift tracer is outside the B200 hot path: its
right-hand side is host-side quadrature and root finding over five field-line traces per evaluation,
the reference's implementation depends on the broken `flutils.eye` (`simps` undefined,
flutils.py:130), and its author deprecates it (BounceCenter.py:59-70).  The constructor, `setpa`,
`save`/`load` and the getters keep the reference's behaviour; `advance` raises.
"""
import pickle
import numpy as np

from . import utils as ru


class BounceCenter:
    def __init__(self, pos=[], v=None, t0=0, pa=None, mass=None, charge=None, field=None):
        # rapt/BounceCenter.py:74-115
        self.pos = pos
        self.v = v
        self.t0 = t0
        self.tcur = t0
        self.mass = mass
        self.charge = charge
        self.field = field
        self.isequatorial = False
        if not field.static:
            raise RuntimeError("BounceCenter does not work with nonstatic fields or electric fields.")
        self.pa = pa
        if not ((hasattr(pos, "__len__") and len(pos) == 0) or v is None or pa is None):
            self.trajectory = np.reshape(np.concatenate(([t0], pos)), (1, 4))
            # as the reference: cos() of the pitch angle as given (BounceCenter.py:114)
            self.mu = ru.magnetic_moment(t0, pos, self.v * np.cos(self.pa), self.v, field, mass)
            assert self.mu > 0

    def setpa(self, pa):
        """Reinitialise with a new pitch angle in degrees (rapt/BounceCenter.py:117-132)."""
        self.__init__(self.pos, self.v, self.t0, pa, self.mass, self.charge, self.field)

    def advance(self, delta):
        raise NotImplementedError("BounceCenter.advance is outside the B200 hot path (SURVEY.md §8f N4); use "
                                  "GuidingCenter.advance for bounce + drift motion")

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    def load(self, filename):
        with open(filename, "rb") as f:
            p = pickle.load(f)
        for k in p.__dict__.keys():
            self.__dict__[k] = p.__dict__[k]

    def getr(self):
        return np.sqrt(self.getx() ** 2 + self.gety() ** 2 + self.getz() ** 2)

    def gettheta(self):
        return np.arctan2(self.gety(), self.getx())

    def getphi(self):
        return np.arccos(self.getz() / self.getr())

    def getB(self):
        return np.array([self.field.magB(row) for row in self.trajectory])


for _i, _name in enumerate(("gett", "getx", "gety", "getz")):
    setattr(BounceCenter, _name, (lambda i: lambda self: self.trajectory[:, i])(_i))
