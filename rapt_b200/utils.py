"""Host-side physics helpers with the reference's names and argument meaning (rapt/utils.py:15-433).

These are API surface for user scripts (single values, numpy).  The `.advance()` hot path does not
call them: cyclotron period, magnetic moment, the particle<->guiding-centre transforms etc. run as
device code (rapt_b200/csrc/rapt_aux.cuh).  The analytic dipole diagnostics of rapt/utils.py:435-607
are off the hot path and not provided (out of scope).
"""
import numpy as np
from . import c, e


def gamma(mass=0, **kwargs):
    """Relativistic factor from velocity=... or momentum=... (+mass) (rapt/utils.py:15-27)."""
    if "velocity" in kwargs:
        v = np.asarray(kwargs["velocity"], dtype=float)
        return 1 / np.sqrt(1 - np.dot(v, v) / c ** 2)
    if "momentum" in kwargs:
        if mass == 0:
            raise ValueError("Particle mass not given.")
        p = np.asarray(kwargs["momentum"], dtype=float)
        return np.sqrt(1 + np.dot(p, p) / (mass * c) ** 2)
    raise ValueError("Either velocity of momentum vectors should be given.")


def _tpos(t, pos):
    return np.concatenate([[t], np.asarray(pos, dtype=float)])


def cyclotron_period(t, pos, vel, field, mass, charge):
    """2 pi gamma m / (|q| B) from the velocity vector (rapt/utils.py:29-66)."""
    vel = np.array(vel, dtype=float)
    g = 1.0 / np.sqrt(1 - np.dot(vel, vel) / c ** 2)
    B = field.magB(_tpos(t, pos))
    return 2 * np.pi * g * mass / B / abs(charge)


def cyclotron_period2(t, pos, speed, field, mass, charge):
    """Same from the total speed (rapt/utils.py:68-104)."""
    g = 1.0 / np.sqrt(1 - (speed / c) ** 2)
    B = field.magB(_tpos(t, pos))
    return 2 * np.pi * g * mass / B / abs(charge)


def cyclotron_radius(t, pos, vel, field, mass, charge):
    """gamma m v_perp / (|q| B) from the velocity vector (rapt/utils.py:106-145)."""
    vel = np.asarray(vel, dtype=float)
    vsq = np.dot(vel, vel)
    g = 1.0 / np.sqrt(1 - vsq / c ** 2)
    B = field.B(_tpos(t, pos))
    Bmag = np.sqrt(np.dot(B, B))
    vpar = np.dot(vel, B) / Bmag
    vperp = np.sqrt(vsq - vpar ** 2)
    return g * mass * vperp / (abs(charge) * Bmag)


def cyclotron_radius2(t, pos, vpar, v, field, mass, charge):
    """Same from parallel and total speed (rapt/utils.py:147-187)."""
    g = 1.0 / np.sqrt(1 - (v / c) ** 2)
    B = field.B(_tpos(t, pos))
    Bmag = np.sqrt(np.dot(B, B))
    vperp = np.sqrt((v - vpar) * (v + vpar))
    return g * mass * vperp / (abs(charge) * Bmag)


def magnetic_moment(t, pos, vpar, v, field, mass):
    """gamma^2 m v_perp^2 / (2B) (rapt/utils.py:189-216)."""
    g = 1.0 / np.sqrt(1 - (v / c) ** 2)
    Bmag = field.magB(_tpos(t, pos))
    return g ** 2 * mass * (v - vpar) * (v + vpar) / (2 * Bmag)


def speedfromKE(KE, mass, unit="ev"):
    """Speed for a relativistic kinetic energy (rapt/utils.py:218-249)."""
    mc2 = mass * c ** 2
    if unit.lower() == "ev":
        KE = KE * e
    if KE / mc2 < 1e-6:
        return np.sqrt(2 * KE / mass)
    return c * np.sqrt(1 - (mc2 / (mc2 + KE)) ** 2)


def guidingcenter(t, r, v, field, mass, charge, tol=1e-3, maxiter=20, debug=False):
    """Guiding centre R = r - rho(R) by fixed-point iteration (rapt/utils.py:251-326).
    Returns (R, v_parallel, speed); prints and returns None if not converged, like the reference."""
    r = np.asarray(r, dtype=float); v = np.asarray(v, dtype=float)

    def gyrovector(rr):
        vsq = np.dot(v, v)
        g = 1 / np.sqrt(1 - vsq / c ** 2)
        B = field.B(_tpos(t, rr))
        return g * mass / (charge * np.dot(B, B)) * np.cross(B, v)

    def norm(a):
        return np.sqrt(np.dot(a, a))

    old = r - gyrovector(r)
    hist = [old]
    it = 1
    while it <= maxiter:
        gc = r - gyrovector(old)
        hist.append(gc)
        if norm(gc - old) / norm(gc) < tol:
            if debug:
                return hist
            B = field.B(_tpos(t, gc))
            return gc, np.dot(v, B) / norm(B), norm(v)
        old = gc
        it += 1
    print("Could not reach the specified tolerance after ", it, " iterations.")


def getperp(v):
    """A unit vector perpendicular to v (rapt/utils.py:328-374)."""
    assert len(v) == 3
    if v[0] == 0.0 and v[1] == 0.0 and v[2] == 0.0:
        raise ValueError('Zero vector')
    if v[0] == 0:
        return np.array([1, 0, 0])
    if v[1] == 0:
        return np.array([0, 1, 0])
    if v[2] == 0:
        return np.array([0, 0, 1])
    cc = -1.0 * (v[0] + v[1]) / v[2]
    return np.array([1, 1, cc]) / np.sqrt(2 + cc ** 2)


def GCtoFP(t, R, vp, speed, field, mass, charge, gyrophase=0):
    """Particle position/velocity for a guiding-centre state (rapt/utils.py:376-433)."""
    R = np.asarray(R, dtype=float)
    B = field.B(_tpos(t, R))
    b = B / np.sqrt(np.dot(B, B))
    pa = np.arccos(vp / speed)
    rc = cyclotron_radius2(t, R, vp, speed, field, mass, charge)
    u = getperp(B)
    u = u / np.sqrt(np.dot(u, u))
    w = np.cross(b, u)
    s = np.sign(charge)
    pos = R + rc * (np.cos(gyrophase) * u + np.sin(gyrophase) * w)
    vel = speed * (np.cos(pa) * b + s * np.sin(pa) * np.sin(gyrophase) * u - s * np.sin(pa) * np.cos(gyrophase) * w)
    return pos, vel
