"""BounceCenter (rapt/BounceCenter.py:17-316): the bounce-averaged drift tracer, SURVEY.md §8f N4.

`advance` runs on the device (rapt_b200/csrc/rapt_bc.cuh): scipy's dopri5 per output row on
dR/dt = gamma m v^2/(q S_b B^2) gradI x B, every right-hand side tracing five field lines and running the
reference's spline / brentq / QUADPACK route on them, one thread per tracer.  Reference behaviour kept: the
constructor takes cos() of the pitch angle as given (BounceCenter.py:114 -- degrees fed to a radian cosine), rows are
labelled with the START time of their step (BounceCenter.py:248-249), only static fields are accepted.
Two reference defects are not reproduced: equatorial pitch angles below 70 degrees stop with NameError in
flutils.eye (`simps`, flutils.py:130) -- here that branch runs with scipy.integrate.simpson's rule, the function the
reference imports; and `isequatorial = True` indexes a 3-vector as a 4-vector (BounceCenter.py:235) -- not offered.
For many tracers use rapt_b200.ensemble.BounceCenterEnsemble.
"""
import pickle
import warnings

import numpy as np

from . import utils as ru
from . import engine, params


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
        """Advance the bounce centre by `delta` seconds (rapt/BounceCenter.py:206-251) on the device."""
        if self.isequatorial:
            raise NotImplementedError("BounceCenter.isequatorial: the reference's branch cannot run (BounceCenter.py:235)")
        last = self.trajectory[-1, :4]
        # rows needed: len(np.arange(tcur, tcur+delta, dt)); dt comes from the device, so size the buffer in two passes
        probe = engine.bounce_center_terms(self.field, last, self._mirror_field(), params=params)
        if probe["status"][0] != 1 or not np.isfinite(probe["Sb"][0]):
            raise RuntimeError("BounceCenter.advance: the field line through the current position could not be traced "
                               "between its mirror points")
        dt = params["BCtimestep"] * (2 / self.v) * probe["Sb"][0]
        nrows = len(np.arange(self.tcur, self.tcur + delta, dt))
        res = engine.bounce_center_advance(self.field, last, self.mu, self.v, self.mass, self.charge, delta,
                                           store_every=1, max_rows=max(nrows, 1), params=params)
        k = int(res["nstored"][0])
        if k:
            self.trajectory = np.vstack((self.trajectory, res["rows"][0, :k]))
        st = int(res["status"][0])
        if st != 1:
            why = {-2: "dopri5: larger nsteps is needed", -3: "dopri5: step size becomes too small",
                   -7: "a field line could not be traced between its mirror points"}.get(st, "error")
            warnings.warn(f"BounceCenter.advance stopped after {k} of {nrows} rows (status {st}: {why})", UserWarning)
        self.tcur = self.trajectory[-1, 0]

    def _mirror_field(self):
        # BounceCenter.py:226-227
        from . import c
        gamma = 1.0 / np.sqrt(1 - (self.v / c) ** 2)
        return self.mass * gamma ** 2 * self.v ** 2 / (2 * self.mu)

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
