"""Adaptive: automatic switching between Particle and GuidingCenter (rapt/Adaptive.py:18-274).

A single Adaptive object keeps the reference's control flow verbatim -- a list of Particle /
GuidingCenter segments, switched on the `Adiabatic` / `NonAdiabatic` exceptions raised by their
(GPU) `advance()` -- so `trajlist`, the getters and the printed switch messages behave as in the
reference.  Ensembles use `rapt_b200.AdaptiveEnsemble`, where the switching runs on the device.
"""
import pickle
import numpy as np

from . import Adiabatic, NonAdiabatic
from .Particle import Particle
from .GuidingCenter import GuidingCenter


class Adaptive:
    def __init__(self, pos=None, vel=None, t0=0, mass=None, charge=None, field=None):
        self.pos = np.array(pos)
        self.vel = np.array(vel)
        self.tcur = t0
        self.mass = mass
        self.charge = charge
        self.field = field
        self.p = Particle(pos, vel, t0, mass, charge, field)
        self.p.check_adiabaticity = True
        self._choose_initial()

    def _choose_initial(self):
        # Adaptive.py:98-104
        if self.p.isadiabatic():
            g = GuidingCenter()
            g.check_adiabaticity = True
            g.init(self.p)
            g.check_adiabaticity = True
            self.trajlist = [g]
        else:
            self.trajlist = [self.p]

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    def load(self, filename):
        with open(filename, "rb") as f:
            p = pickle.load(f)
        for k in p.__dict__.keys():
            self.__dict__[k] = p.__dict__[k]

    def setke(self, ke, unit="ev"):
        self.p.setke(ke, unit)
        self.p.check_adiabaticity = True
        self._choose_initial()

    def setpa(self, pa):
        self.p.setpa(pa)
        self.p.check_adiabaticity = True
        self._choose_initial()

    def advance(self, delta):
        """Advance for `delta` seconds, switching modes as needed (rapt/Adaptive.py:187-222).
        As in the reference, the loop compares the tracer's absolute `tcur` with `delta` (:205,222)."""
        t = 0
        current = self.trajlist[-1]
        assert current.check_adiabaticity is True
        while t < delta:
            try:
                current.advance(delta - t)
            except NonAdiabatic:
                p = Particle()
                p.init(current)
                p.check_adiabaticity = True
                self.trajlist.append(p)
                current = self.trajlist[-1]
                print("Switched to particle mode at time", current.tcur, flush=True)
            except Adiabatic:
                g = GuidingCenter()
                g.init(current)
                g.check_adiabaticity = True
                self.trajlist.append(g)
                current = self.trajlist[-1]
                print("Switched to guiding center mode at time", current.tcur, flush=True)
            t = current.tcur
        self.tcur = t

    def _cat(self, name):
        res = np.array([])
        for p in self.trajlist:
            res = np.concatenate((res, getattr(p, name)()))
        return res


def _concatenating(name):
    def get(self):
        return self._cat(name)
    get.__doc__ = f"`{name}()` of every segment in `trajlist`, concatenated (rapt/Adaptive.py:224-274)."
    return get


for _name in ("gett", "getx", "gety", "getz", "getr", "getphi", "gettheta", "getke"):
    setattr(Adaptive, _name, _concatenating(_name))
