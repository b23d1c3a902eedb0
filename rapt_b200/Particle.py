"""Particle: full-orbit tracer with the reference's interface (rapt/Particle.py:18-494).

`advance()` runs on the GPU: the relativistic Newton-Lorentz equation integrated by the per-thread
DOP853 kernel (rapt_b200/csrc/rapt_particle.cuh) through the C ABI; a single Particle is an ensemble
of size 1.  Everything else (constructor, getters, setke/setpa, save/load) is host-side Python, as in
the reference.
"""
import pickle
import numpy as np

from . import c, params, Adiabatic
from . import utils as ru
from . import engine


def _empty(a):
    return a is None or (hasattr(a, "__len__") and len(a) == 0)


class Particle:
    """A classical relativistic charged particle in given E and B fields (rapt/Particle.py:18-58).

    Parameters: pos (m), vel (m/s), t0 (s), mass (kg), charge (C), field (rapt_b200.fields object).
    Attributes: tcur, trajectory (n x 7: t,x,y,z,px,py,pz), check_adiabaticity.
    """

    def __init__(self, pos=[], vel=[], t0=0, mass=None, charge=None, field=None):
        self.pos = np.array(pos, dtype=float)
        self.vel = np.array(vel, dtype=float)
        self.t0 = t0
        self.tcur = t0
        self.mass = mass
        self.charge = charge
        self.field = field
        self.trajectory = np.zeros((1, 7))
        self.check_adiabaticity = False
        self.solver_counters = np.zeros(4, dtype=np.int64)   # (nfcn, nstep, naccpt, nrejct) of the last advance()
        if not (_empty(pos) or _empty(vel) or self.mass is None):    # Particle.py:105-109
            g = 1 / np.sqrt(1 - np.dot(self.vel, self.vel) / c ** 2)
            mom = self.mass * g * self.vel
            self.trajectory = np.reshape(np.concatenate(([self.tcur], self.pos, mom)), (1, 7))

    def init(self, p, gyrophase=0):
        """Initialise from the last state of another Particle or GuidingCenter (rapt/Particle.py:111-166)."""
        from .GuidingCenter import GuidingCenter
        if isinstance(p, Particle):
            mom = p.trajectory[-1, 4:]
            gm = np.sqrt(p.mass ** 2 + np.dot(mom, mom) / c ** 2)
            self.__init__(pos=p.trajectory[-1, 1:4], vel=p.trajectory[-1, 4:] / gm, t0=p.trajectory[-1, 0],
                          mass=p.mass, charge=p.charge, field=p.field)
        elif isinstance(p, GuidingCenter):
            # field evaluated at the NEW object's tcur, as the reference does (Particle.py:157)
            B = p.field.magB(p.trajectory[-1, :4])
            gammasq = 1 + 2 * p.mu * B / (p.mass * c * c) + (p.trajectory[-1, 4] / p.mass / c) ** 2
            if np.sqrt(gammasq) - 1 < 1e-6:
                v = np.sqrt(2 * p.mu * B / p.mass + (p.trajectory[-1, 4] / p.mass) ** 2)
            else:
                v = c * np.sqrt(1 - 1 / gammasq)
            vpar = p.trajectory[-1, 4] / p.mass / np.sqrt(gammasq)
            pos, vel = ru.GCtoFP(self.tcur, p.trajectory[-1, 1:4], vpar, v, p.field, p.mass, p.charge, gyrophase)
            self.__init__(pos=pos, vel=vel, t0=p.trajectory[-1, 0], mass=p.mass, charge=p.charge, field=p.field)
        else:
            raise ValueError("Particle or GuidingCenter objects required.")

    def setke(self, ke, unit="ev"):
        """Rescale the velocity to the given kinetic energy; reinitialises (rapt/Particle.py:168-187)."""
        assert ke > 0
        s = ru.speedfromKE(ke, self.mass, unit)
        mom = self.trajectory[-1, 4:]
        gm = np.sqrt(self.mass ** 2 + np.dot(mom, mom) / c ** 2)
        v = mom / gm
        v = v * (s / np.sqrt(np.dot(v, v)))
        self.__init__(self.pos, v, self.t0, self.mass, self.charge, self.field)

    def setpa(self, pa):
        """Reinitialise with pitch angle `pa` degrees at constant speed (rapt/Particle.py:189-228)."""
        tpos = self.trajectory[-1, 0:4]
        mom = self.trajectory[-1, 4:]
        gm = np.sqrt(self.mass ** 2 + np.dot(mom, mom) / c ** 2)
        v = mom / gm
        s = np.sqrt(np.dot(v, v))
        b = self.field.unitb(tpos)
        spar = np.dot(v, b)
        if abs(spar - s) < 1e-12:
            p = ru.getperp(b)
        else:
            vperp = v - spar * b
            p = vperp / np.sqrt(np.dot(vperp, vperp))
        w = s * np.sin(pa * np.pi / 180) * p + s * np.cos(pa * np.pi / 180) * b
        self.__init__(self.pos, w, self.t0, self.mass, self.charge, self.field)

    def advance(self, delta):
        """Advance position and momentum for `delta` seconds (rapt/Particle.py:230-309) on the GPU.

        Output rows are one cyclotron period / params['cyclotronresolution'] apart; every row is a fresh
        DOP853 call with rtol, atol = params['solvertolerances'].  Raises `Adiabatic` after the row at
        which the motion became adiabatic if `check_adiabaticity` is set.  May be called repeatedly.
        """
        last = self.trajectory[-1]
        dt = float(engine.particle_dt(self.field, last, self.mass, self.charge)[0])
        max_rows = max(int(np.ceil(delta / dt)) + 8, 8) if delta > 0 and np.isfinite(dt) and dt > 0 else 8
        asked = -1
        while True:
            o = engine.particle_advance(self.field, last, self.mass, self.charge, float(delta), store_every=1,
                                        max_rows=max_rows, check_adiabaticity=self.check_adiabaticity)
            n = int(o["nstored"][0])
            if o["nrows"][0] <= n or o["nrows"][0] == asked:
                break                              # all rows stored (or a rerun that asks for the same size again)
            asked = int(o["nrows"][0])
            max_rows = asked + 8                   # buffer was too small (dt estimate off): rerun, deterministic
        self.trajectory = np.vstack((self.trajectory, o["rows"][0, 1:n, :7]))
        self.solver_counters = o["counters"][0].astype(np.int64)
        if n > 1:
            self.tcur = float(o["tcur"][0])
        status = int(o["status"][0])
        if status == -6:       # RAPT_ST_FIELD: scipy's ValueError inside Grid.Bgrid/Egrid; the rows so far are kept
            raise ValueError("One of the requested xi is out of bounds: the tracer left the grid of the field")
        if status < 0:
            import warnings
            warnings.warn({-2: "dop853: larger nsteps is needed", -3: "dop853: step size becomes too small"}.get(
                status, f"dop853: solver status {status}"), stacklevel=2)
        if self.check_adiabaticity and status == 2:
            raise Adiabatic

    def save(self, filename):
        """Pickle the object (rapt/Particle.py:311-324)."""
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    def load(self, filename):
        """Replace this object's data with a pickled one (rapt/Particle.py:326-343)."""
        with open(filename, "rb") as f:
            p = pickle.load(f)
        for k in p.__dict__.keys():
            self.__dict__[k] = p.__dict__[k]

    def isadiabatic(self):
        """rho_c / L < epss [and tau_c / T < epst if the field is not static] at the last row
        (rapt/Particle.py:345-384); evaluated by the same device code the advance kernel uses."""
        return bool(engine.isadiabatic(self.field, 0, self.trajectory[-1], 0.0, self.mass, self.charge)[0])

    # ---- getters (rapt/Particle.py:386-461): gett, getx, gety, getz, getpx, getpy, getpz are column views,
    # attached below from _COLUMNS
    def getp(self):
        return np.sqrt(self.getpx() ** 2 + self.getpy() ** 2 + self.getpz() ** 2)

    def getgamma(self):
        psq = self.trajectory[:, 4] ** 2 + self.trajectory[:, 5] ** 2 + self.trajectory[:, 6] ** 2
        return np.sqrt(1 + psq / (self.mass * c) ** 2)

    def getvx(self):
        return self.getpx() / self.getgamma() / self.mass

    def getvy(self):
        return self.getpy() / self.getgamma() / self.mass

    def getvz(self):
        return self.getpz() / self.getgamma() / self.mass

    def getv(self):
        """Particle speed.  (The reference's formula, Particle.py:431, takes sqrt(1-gamma^2) and yields NaN;
        this returns c*sqrt(1-1/gamma^2).)"""
        g = self.getgamma()
        return c * np.sqrt(1 - 1 / g ** 2)

    def getr(self):
        return np.sqrt(self.getx() ** 2 + self.gety() ** 2 + self.getz() ** 2)

    def gettheta(self):
        return np.arctan2(self.gety(), self.getx())

    def getphi(self):
        return np.arccos(self.getz() / self.getr())

    def getke(self):
        g = self.getgamma()
        ke_nr = 0.5 * (self.trajectory[:, 4] ** 2 + self.trajectory[:, 5] ** 2 + self.trajectory[:, 6] ** 2) / self.mass
        return np.where(g - 1 < 1e-6, ke_nr, (g - 1) * self.mass * c * c)

    def getB(self):
        """|B| along the trajectory (the reference's version forgets to return, Particle.py:456-461)."""
        return engine.field_ops(self.field, self.trajectory[:, :4], which=["magB"])["magB"]

    def guidingcenter(self):
        """Guiding-centre position, parallel speed and speed for every row (rapt/Particle.py:463-472)."""
        grow, mu, v, st = engine.switch_p2g(self.field, self.trajectory, self.mass, self.charge)
        g = 1 / np.sqrt(1 - (v / c) ** 2)
        return np.column_stack([grow[:, 1:4], grow[:, 4] / (self.mass * g), v])

    def mu(self):
        """First adiabatic invariant for every row (rapt/Particle.py:473-482)."""
        grow, mu, v, st = engine.switch_p2g(self.field, self.trajectory, self.mass, self.charge)
        return mu

    def cycrad(self):
        t, r, mom = self.trajectory[-1, 0], self.trajectory[-1, 1:4], self.trajectory[-1, 4:]
        gm = np.sqrt(self.mass ** 2 + np.dot(mom, mom) / c ** 2)
        return ru.cyclotron_radius(t, r, mom / gm, self.field, self.mass, self.charge)

    def cycper(self):
        t, r, mom = self.trajectory[-1, 0], self.trajectory[-1, 1:4], self.trajectory[-1, 4:]
        gm = np.sqrt(self.mass ** 2 + np.dot(mom, mom) / c ** 2)
        return ru.cyclotron_period(t, r, mom / gm, self.field, self.mass, self.charge)


def _column_getter(index, what):
    def get(self):
        return self.trajectory[:, index]
    get.__doc__ = f"1-d array of {what} along the trajectory."
    return get


_COLUMNS = (("gett", "time values"), ("getx", "the x coordinate"), ("gety", "the y coordinate"), ("getz", "the z coordinate"),
            ("getpx", "the x component of the momentum"), ("getpy", "the y component of the momentum"),
            ("getpz", "the z component of the momentum"))
for _i, (_name, _what) in enumerate(_COLUMNS):
    setattr(Particle, _name, _column_getter(_i, _what))
