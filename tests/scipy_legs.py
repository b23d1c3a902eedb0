"""TEST INFRASTRUCTURE: scipy's own interp1d / brentq / quad / simpson applied to field lines traced on the device -- the
way the reference evaluates flutils.halfbouncepath / eye (flutils.py:65-316).  The product computes these integrals per
thread in rapt_quad.cuh; these host legs exist only to cross-check that restatement (tests/test_gpu_gc.py,
tests/test_gpu_userfield.py, tests/test_host_api.py, tools/bounce_diff.py).  Nothing under rapt_b200/ imports this file.
"""
import numpy as np

from rapt_b200 import engine
from rapt_b200.engine import _col, bounce_setup, fieldline_trace_many


def eye_from_curve(s, b, Bm):
    """Second invariant I = integral of sqrt(1 - B(s)/Bm) ds between the mirror points of a traced field line
    (flutils.eye, flutils.py:65-151).  Equatorial pitch angle below 70 degrees: Simpson's rule over the interior
    points plus the closed-form end intervals -- the reference calls an undefined name `simps` there
    (flutils.py:130, NameError); scipy.integrate.simpson(y, x=x) is what it imports and means.  Otherwise
    scipy's quadratic spline, brentq and QUADPACK exactly as the reference (third-party there too)."""
    s = np.asarray(s, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    bmin = b.min()
    if bmin > Bm or abs(bmin - Bm) / Bm < 1e-12:          # no mirror points on the line / equatorial
        return 0.0
    inside = np.flatnonzero(b < Bm)
    lo, hi = inside[0] - 1, inside[-1] + 1                # keep exactly one point beyond each mirror point
    if lo < 0 or hi >= len(b):
        raise AssertionError("field-line trace does not bracket the mirror points")
    s = s[lo:hi + 1].copy(); b = b[lo:hi + 1].copy()
    eqpa = np.arcsin(np.sqrt(bmin / Bm)) * 180 / np.pi
    if eqpa < 70:
        s[0] = (Bm - b[0]) * (s[1] - s[0]) / (b[1] - b[0]) + s[0]               # mirror points by linear interpolation
        s[-1] = (Bm - b[-2]) * (s[-1] - s[-2]) / (b[-1] - b[-2]) + s[-2]
        from scipy.integrate import simpson
        val = simpson(np.sqrt(1 - b[1:-1] / Bm), x=s[1:-1])
        val += (2 / 3) * (s[-1] - s[-2]) * np.sqrt((Bm - b[-2]) / Bm)          # sqrt-type end intervals
        val += (2 / 3) * (s[1] - s[0]) * np.sqrt((Bm - b[1]) / Bm)
        return float(val)
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from scipy.integrate import quad
    Bf = interp1d(s, b, kind='quadratic', assume_sorted=True)
    sm1 = brentq(lambda x: Bf(x) - Bm, s[0], s[1])
    sm2 = s[-2] if Bf(s[-2]) == Bm else brentq(lambda x: Bf(x) - Bm, s[-2], s[-1])
    return float(quad(lambda x: np.sqrt(1 - Bf(x) / Bm), sm1, sm2, epsrel=1e-4)[0])


def eye_host(field, tpos, Bm, fieldlineresolution=None, arith="strict"):
    """Cross-check of `eye`: device traces + scipy's own interp1d / brentq / quad on the host."""
    curves, _ = fieldline_trace_many(field, tpos, Bm, fieldlineresolution, arith)
    Bm = _col(Bm, len(curves))
    return np.array([eye_from_curve(cv[:, 0], cv[:, 4], Bm[i]) for i, cv in enumerate(curves)])


def halfbouncepath_from_curve(s, b, Bm):
    """flutils.halfbouncepath (flutils.py:274-316) on a traced curve.  The non-equatorial branch uses
    scipy's quadratic spline / brentq / QUADPACK exactly as the reference does (third-party there too)."""
    n = len(b)
    inside = np.where(b <= Bm)[0]
    if len(inside) == 0:
        i1 = int((n - 3) / 2); i2 = int((n + 1) / 2)
    else:
        i1, i2 = inside[0] - 1, inside[-1] + 1
    keep = [i for i in range(n) if i1 <= i <= i2]
    b = np.asarray(b)[keep]; s = np.asarray(s)[keep]
    n = len(b)
    if n == 3:
        s12, s23, s13 = s[0] - s[1], s[1] - s[2], s[0] - s[2]
        B2s = 2 * (b[0] * s23 - b[1] * s13 + b[2] * s12) / (s12 * s13 * s23)
        return np.pi * np.sqrt(2 * Bm / B2s)
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from scipy.integrate import quad
    Bf = interp1d(s, b, kind='quadratic', assume_sorted=True)
    sm1 = brentq(lambda x: Bf(x) - Bm, s[0], s[1])
    sm2 = brentq(lambda x: Bf(x) - Bm, s[-2], s[-1])
    return quad(lambda x: 1 / np.sqrt(1 - Bf(x) / Bm), sm1, sm2, epsrel=1e-4)[0]


def bounceperiod(field, state, mu, mass, fieldlineresolution=None, arith="strict"):
    """Cross-check of `bounceperiod_device`: device field-line traces + scipy's own quadrature on the host."""
    bs = bounce_setup(field, state, mu, mass, fieldlineresolution, arith)
    n = len(bs["Bm"])
    out = np.zeros(n)
    for i in range(n):
        k = bs["npts"][i]
        cv = bs["curve"][i, :k]
        out[i] = (2 / bs["v"][i]) * halfbouncepath_from_curve(cv[:, 0], cv[:, 4], bs["Bm"][i])
    return out
