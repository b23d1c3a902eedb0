"""GuidingCenter: guiding-centre tracer with the reference's interface (rapt/GuidingCenter.py:18-624).

`advance()` runs on the GPU (per-thread DOPRI5, three selectable equations of motion on the
finite-difference field operators: rapt_b200/csrc/rapt_gc.cuh).  The output step is
params['GCtimestep'] or bounceperiod()/params['bounceresolution']; the bounce period (field-line trace, then
scipy's quadratic spline, brentq and QUADPACK QAGS restated per thread) is computed on the device as well.
"""
import pickle
import numpy as np

from . import c, params, NonAdiabatic
from . import utils as ru
from . import engine


def _empty(a):
    return a is None or (hasattr(a, "__len__") and len(a) == 0)


class GuidingCenter:
    """Guiding centre of a charged particle; mu is a constant of the motion (rapt/GuidingCenter.py:18-133).

    Parameters: pos (m), v (speed, m/s), pa (pitch angle, degrees) or ppar (kg m/s), t0, mass, charge, field.
    Attributes: tcur, trajectory (n x 5: t,x,y,z,p_parallel), mu, check_adiabaticity.
    """

    def __init__(self, pos=[], v=0, pa=None, ppar=None, t0=0, mass=None, charge=None, field=None):
        self.pos = pos
        self.v = v
        self.t0 = t0
        self.tcur = t0
        self.mass = mass
        self.charge = charge
        self.field = field
        self.trajectory = np.zeros((1, 5))
        self.check_adiabaticity = False
        self.solver_counters = np.zeros(4, dtype=np.int64)
        if not (_empty(pos) or v == 0):                       # GuidingCenter.py:123-133
            g = 1 / np.sqrt(1 - (v / c) ** 2)
            if pa is not None:
                vpar = 0 if pa == 90 else v * np.cos(pa * np.pi / 180)
                ppar = g * mass * vpar
            self.mu = ru.magnetic_moment(self.tcur, self.pos, ppar / (mass * g), self.v, self.field, self.mass)
            self.trajectory[0, 0] = t0
            self.trajectory[0, 1:4] = pos[:]
            self.trajectory[0, 4] = ppar

    def init(self, p):
        """Initialise from the last state of a Particle or GuidingCenter (rapt/GuidingCenter.py:135-188)."""
        from .Particle import Particle
        if isinstance(p, GuidingCenter):
            B = p.field.magB(p.trajectory[-1, :4])
            g = np.sqrt(1 + 2 * p.mu * B / (p.mass * c * c) + (p.trajectory[-1, 4] / p.mass / c) ** 2)
            if g - 1 < 1e-6:
                v = np.sqrt(2 * p.mu * B / p.mass + (p.trajectory[-1, 4] / p.mass) ** 2)
            else:
                v = c * np.sqrt(1 - 1 / g ** 2)
            self.__init__(pos=p.trajectory[-1, 1:4], v=v, ppar=p.trajectory[-1, 4], t0=p.trajectory[-1, 0],
                          mass=p.mass, charge=p.charge, field=p.field)
            self.check_adiabaticity = p.check_adiabaticity
        elif isinstance(p, Particle):
            mom = p.trajectory[-1, 4:]
            gm = np.sqrt(p.mass ** 2 + np.dot(mom, mom) / c ** 2)
            vel = mom / gm
            res = ru.guidingcenter(p.trajectory[-1, 0], p.trajectory[-1, 1:4], vel, p.field, p.mass, p.charge)
            if res is None:
                raise TypeError("guiding-centre iteration did not converge (utils.guidingcenter returned None)")
            pos, vp, v = res
            g = 1 / np.sqrt(1 - (v / c) ** 2)
            self.__init__(pos=pos, v=v, ppar=p.mass * g * vp, t0=p.trajectory[-1, 0],
                          mass=p.mass, charge=p.charge, field=p.field)
            self.check_adiabaticity = p.check_adiabaticity
        else:
            raise ValueError("Particle or GuidingCenter objects required.")

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    def load(self, filename):
        with open(filename, "rb") as f:
            p = pickle.load(f)
        for k in p.__dict__.keys():
            self.__dict__[k] = p.__dict__[k]

    def setke(self, ke, unit="ev"):
        """New speed for kinetic energy `ke`, same pitch angle; reinitialises (rapt/GuidingCenter.py:225-255).
        (The reference passes the pitch angle in radians where degrees are expected, :248,254; here it is
        converted to degrees.)"""
        mc = self.mass * c
        t, x, y, z, ppar = self.trajectory[-1]
        B = self.field.magB(self.trajectory[-1, :4])
        gammasq_minus_1 = 2 * self.mu * B / (mc * c) + (ppar / mc) ** 2
        if np.sqrt(gammasq_minus_1 + 1) - 1 < 1e-6:
            ptot = np.sqrt(2 * self.mass * self.mu * B + ppar ** 2)
        else:
            ptot = np.sqrt(gammasq_minus_1) * mc
        pa_old = np.arccos(ppar / ptot) * 180 / np.pi
        v_new = ru.speedfromKE(ke, self.mass, unit)
        self.__init__(pos=[x, y, z], v=v_new, pa=pa_old, t0=t, mass=self.mass, charge=self.charge, field=self.field)

    def setpa(self, pa):
        """Reinitialise with pitch angle `pa` degrees at constant energy (rapt/GuidingCenter.py:257-285)."""
        mc = self.mass * c
        t, x, y, z, ppar = self.trajectory[-1]
        B = self.field.magB(self.trajectory[-1, :4])
        gammasq = 1 + 2 * self.mu * B / (mc * c) + (ppar / mc) ** 2
        if np.sqrt(gammasq) - 1 < 1e-6:
            v = np.sqrt(2 * self.mass * self.mu * B + ppar ** 2) / self.mass
        else:
            v = c * np.sqrt(1 - 1 / gammasq)
        self.__init__(pos=[x, y, z], v=v, pa=pa, t0=t, mass=self.mass, charge=self.charge, field=self.field)

    def isadiabatic(self):
        """rho_c / L < epss [and tau_c / T < epst] at the last row (rapt/GuidingCenter.py:287-327), device code."""
        return bool(engine.isadiabatic(self.field, 1, self.trajectory[-1], self.mu, self.mass, self.charge)[0])

    def advance(self, delta, eom="TaoChanBrizardEOM"):
        """Advance position and parallel momentum for `delta` seconds (rapt/GuidingCenter.py:397-458) on the GPU.

        eom in {'TaoChanBrizardEOM', 'BrizardChanEOM', 'NorthropTellerEOM'}.  Raises `NonAdiabatic` after
        the row at which the motion stopped being adiabatic if `check_adiabaticity` is set."""
        if params["GCtimestep"] != 0:
            dt = params["GCtimestep"]
        else:
            dt = self.bounceperiod() / params["bounceresolution"]
        last = self.trajectory[-1]
        max_rows = max(int(np.ceil(delta / dt)) + 8, 8) if delta > 0 and np.isfinite(dt) and dt > 0 else 8
        asked = -1
        while True:
            o = engine.gc_advance(self.field, last, self.mu, self.v, self.mass, self.charge, dt, float(delta), eom=eom,
                                  store_every=1, max_rows=max_rows, check_adiabaticity=self.check_adiabaticity)
            n = int(o["nstored"][0])
            if o["nrows"][0] <= n or o["nrows"][0] == asked:
                break
            asked = int(o["nrows"][0])
            max_rows = asked + 8
        self.trajectory = np.vstack((self.trajectory, o["rows"][0, 1:n, :5]))
        self.solver_counters = o["counters"][0].astype(np.int64)
        if n > 1:
            self.tcur = float(o["tcur"][0])
        status = int(o["status"][0])
        if status == -6:       # RAPT_ST_FIELD: scipy's ValueError inside Grid.Bgrid/Egrid; the rows so far are kept
            raise ValueError("One of the requested xi is out of bounds: the tracer left the grid of the field")
        if status < 0:
            import warnings
            warnings.warn({-2: "dopri5: larger nsteps is needed", -3: "dopri5: step size becomes too small"}.get(
                status, f"dopri5: solver status {status}"), stacklevel=2)
        if self.check_adiabaticity and status == 3:
            raise NonAdiabatic

    # ---- getters (rapt/GuidingCenter.py:460-591): gett, getx, gety, getz, getpp are column views (attached below)
    def getr(self):
        return np.sqrt(self.getx() ** 2 + self.gety() ** 2 + self.getz() ** 2)

    def gettheta(self):
        return np.arctan2(self.gety(), self.getx())

    def getphi(self):
        return np.arccos(self.getz() / self.getr())

    def getB(self):
        return engine.field_ops(self.field, self.trajectory[:, :4], which=["magB"])["magB"]

    def getgamma(self):
        mc = self.mass * c
        return np.sqrt(1 + 2 * self.mu * self.getB() / (mc * c) + (self.trajectory[:, 4] / mc) ** 2)

    def getp(self):
        """Total momentum.  (The reference's relativistic branch lacks a square root, GuidingCenter.py:507.)"""
        mc = self.mass * c
        g = self.getgamma(); B = self.getB(); pp = self.trajectory[:, 4]
        return np.where(g - 1 < 1e-6, np.sqrt(2 * self.mass * self.mu * B + pp ** 2), mc * np.sqrt((g - 1) * (g + 1)))

    def getv(self):
        """Particle speed along the trajectory (the reference calls a non-existent self.gamma, :513)."""
        return self.getp() / self.getgamma() / self.mass

    def cycrad(self):
        t, r, pp = self.trajectory[-1, 0], self.trajectory[-1, 1:4], self.trajectory[-1, 4]
        Bmag = self.field.magB(self.trajectory[-1, :4])
        g = np.sqrt(1 + 2 * self.mu * Bmag / (self.mass * c * c) + (pp / self.mass / c) ** 2)
        if g - 1 < 1e-6:
            vp = pp / self.mass
            v = np.sqrt(2 * self.mu * Bmag / self.mass + vp ** 2)
        else:
            vp = pp / self.mass / g
            v = c * np.sqrt(1 - 1 / g ** 2)
        return ru.cyclotron_radius2(t, r, vp, v, self.field, self.mass, self.charge)

    def cycper(self):
        # keeps the reference's pp**2 (not (pp/mc)**2) in gamma, GuidingCenter.py:535
        t, r, pp = self.trajectory[-1, 0], self.trajectory[-1, 1:4], self.trajectory[-1, 4]
        Bmag = self.field.magB(self.trajectory[-1, :4])
        g = np.sqrt(1 + 2 * self.mu * Bmag / (self.mass * c * c) + pp ** 2)
        if g - 1 < 1e-6:
            vp = pp / self.mass
            v = np.sqrt(2 * self.mu * Bmag / self.mass + vp ** 2)
        else:
            v = c * np.sqrt(1 - 1 / g ** 2)
        return ru.cyclotron_period2(t, r, v, self.field, self.mass, self.charge)

    def getBm(self):
        mc = self.mass * c
        g = self.getgamma(); B = self.getB(); pp = self.trajectory[:, 4]
        with np.errstate(divide="ignore", invalid="ignore"):
            nr = B + 0.5 * pp ** 2 / (self.mu * self.mass)
            rel = B / (1 - (pp / mc) ** 2 / ((g - 1) * (g + 1)))
        return np.where(g - 1 < 1e-6, nr, rel)

    def getke(self):
        g = self.getgamma(); B = self.getB(); pp = self.trajectory[:, 4]
        return np.where(g - 1 < 1e-6, self.mu * B + 0.5 * pp ** 2 / self.mass, (g - 1) * self.mass * c * c)

    def bounceperiod(self):
        """Bounce period at the current position (rapt/GuidingCenter.py:593-606), all on the device: field-line
        trace (RKF45, rapt_aux.cuh), then scipy's spline / brentq / QUADPACK route restated per thread (rapt_quad.cuh)."""
        return float(engine.bounceperiod_device(self.field, self.trajectory[-1], self.mu, self.mass,
                                                params["fieldlineresolution"], quadrature="quadpack")[0])

    def geteye(self, step=1):
        """(time, second invariant I) for every `step`-th row (rapt/GuidingCenter.py:608-624, flutils.py:65-151):
        every row's field line is traced and integrated on the device in one call."""
        rows = self.trajectory[::step]
        res = engine.eye(self.field, rows[:, :4], self.getBm()[::step])
        return np.column_stack([rows[:, 0], res])


def _column_getter(index, what):
    def get(self):
        return self.trajectory[:, index]
    get.__doc__ = f"1-d array of {what} along the trajectory."
    return get


for _i, (_name, _what) in enumerate((("gett", "time values"), ("getx", "the x coordinate"), ("gety", "the y coordinate"),
                                     ("getz", "the z coordinate"), ("getpp", "the parallel momentum"))):
    setattr(GuidingCenter, _name, _column_getter(_i, _what))
