"""Fieldline: magnetic field-line tracing with the reference's interface (rapt/fieldline.py:12-133),
restricted to what the hot path uses: trace both ways from a point until |B| > Bmax with the RKF45
scheme of rapt/rkf.py, on the device (rapt_b200/csrc/rapt_aux.cuh, k_bounce_setup)."""
import numpy as np

from . import params
from . import engine


class Fieldline:
    def __init__(self, tpos, field, ds=0, stopcond=None, Bmin=None, Bmax=None):
        if stopcond is not None or Bmin is not None or ds != 0 or Bmax is None:
            raise NotImplementedError("the device tracer implements the Bmax-terminated trace with the automatic "
                                      "step that flutils.halfbouncepath uses (fieldline.py:13-35)")
        self.time = tpos[0]
        self.initpt = np.concatenate(([0], tpos[1:]))
        self.curve = np.zeros((1, 4))
        self.curve[0, :] = self.initpt
        self.field = field
        self.Bmin, self.Bmax = Bmin, Bmax
        self.solver = params["flsolver"]
        self.ds = None
        self._B = None

    def trace(self):
        """Fill `curve` (rows s,x,y,z) (rapt/fieldline.py:37-105)."""
        tpos = np.concatenate(([self.time], self.initpt[1:]))
        o = engine.fieldline_trace(self.field, tpos, self.Bmax, params["fieldlineresolution"])
        self.curve = o["curve"][:, :4].copy()
        self._B = o["curve"][:, 4].copy()
        self.ds = o["ds"]

    def reset(self):
        self.curve = np.zeros((1, 4))

    def gets(self):
        return self.curve[:, 0]

    def getx(self):
        return self.curve[:, 1]

    def gety(self):
        return self.curve[:, 2]

    def getz(self):
        return self.curve[:, 3]

    def getB(self):
        if self._B is not None and len(self._B) == len(self.curve):
            return self._B
        tp = np.column_stack([np.full(len(self.curve), self.time), self.curve[:, 1:]])
        return engine.field_ops(self.field, tp, which=["magB"])["magB"]

    def getr(self):
        return np.sqrt(self.getx() ** 2 + self.gety() ** 2 + self.getz() ** 2)

    def gettheta(self):
        return np.arctan2(self.gety(), self.getx())

    def getphi(self):
        return np.arccos(self.getz() / self.getr())
