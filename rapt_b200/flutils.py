"""rapt.flutils' field-line integrals (rapt/flutils.py:65-316; exported by rapt/__init__.py:42) on the device.

Each call traces the field line(s) through the given point(s) with the reference's RKF45 and runs the reference's
own route over the curve -- scipy's quadratic interpolating spline, brentq, QUADPACK QAGS -- restated per thread
(rapt_b200/csrc/rapt_quad.cuh, rapt_bc.cuh).  `tpos` may be one 4-vector (returns a scalar / 3-vector, as the
reference) or an (n, 4) array (returns arrays).  `usedipole=True` (closed-form dipole helpers, utils.py) is not offered.
"""
import numpy as np

from . import engine


def _call(tpos, field, Bm):
    tp = np.asarray(tpos, dtype=float)
    res = engine.bounce_center_terms(field, tp.reshape(-1, 4), Bm)
    return tp.ndim == 1, res


def halfbouncepath(tpos, field, Bm):
    """Half-bounce path length S_b (rapt/flutils.py:254-316)."""
    one, r = _call(tpos, field, Bm)
    return float(r["Sb"][0]) if one else r["Sb"]


def bounceperiod(tpos, field, Bm, v):
    """Bounce period (2/v) S_b (rapt/flutils.py:232-252)."""
    return (2 / v) * halfbouncepath(tpos, field, Bm)


def eye(tpos, field, Bm, usedipole=False):
    """Second invariant I (rapt/flutils.py:65-151).  Equatorial pitch angles below 70 degrees use
    scipy.integrate.simpson's rule where the reference calls the undefined `simps` (flutils.py:130)."""
    if usedipole:
        raise NotImplementedError("usedipole=True is not offered on the device")
    one, r = _call(tpos, field, Bm)
    return float(r["I"][0]) if one else r["I"]


def gradI(tpos, field, Bm, usedipole=False):
    """Gradient of the second invariant (rapt/flutils.py:153-229), step params["eyegradientstep"]."""
    if usedipole:
        raise NotImplementedError("usedipole=True is not offered on the device")
    one, r = _call(tpos, field, Bm)
    return r["gradI"][0] if one else r["gradI"]
