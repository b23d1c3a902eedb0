"""rapt_quad.cuh (spline / brentq / QUADPACK QAGS / Simpson, the numerics behind flutils.halfbouncepath and
flutils.eye, flutils.py:65-151,254-316) compiled for the host and pinned against scipy itself -- the library the
reference calls.  CPU only; the CUDA build of the same header is checked in tests/test_gpu_bc.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.integrate import quad, simpson
from scipy.interpolate import interp1d
from scipy.optimize import brentq

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    d = os.path.join(HERE, "hostcheck")
    subprocess.check_call(["make", "-C", d, "-s"])
    lib = C.CDLL(os.path.join(d, "libquadhost.so"))
    lib.hc_brentq.restype = C.c_double
    lib.hc_halfbounce.restype = C.c_double
    lib.hc_eye.restype = C.c_double
    lib.hc_simpson.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


FNS = {
    0: lambda x, p: np.sqrt(x),
    1: lambda x, p: 1 / np.sqrt(x),
    2: lambda x, p: np.log(x) / np.sqrt(x),
    3: lambda x, p: 1 / np.sqrt(abs(x - p)),
    4: lambda x, p: np.cos(p * x) * np.exp(-x),
    5: lambda x, p: 1 / (1 + p * x * x),
    6: lambda x, p: x ** p,
    7: lambda x, p: np.sqrt(abs(np.sin(p * x))),
}

QAGS_CASES = [(0, 0, 0, 1), (0, 0, 0, 3.7), (1, 0, 0, 1), (1, 0, 0, 2.5), (2, 0, 0, 1), (3, 0.3, 0, 1), (3, 1 / 3, 0, 1),
              (4, 10, 0, 5), (4, 50, 0, 3), (5, 100, -1, 1), (5, 1e4, -1, 2), (6, -0.9, 0, 1), (6, -0.5, 0, 1), (6, 0.1, 0, 1),
              (6, 2.5, 0, 2), (7, 3, 0, 4), (7, 9, 0.1, 5)]


@pytest.mark.parametrize("fid,p,a,b", QAGS_CASES)
@pytest.mark.parametrize("eps", [(1.49e-8, 1e-4), (1.49e-8, 1.49e-8), (0.0, 1e-10)])
def test_qags_matches_scipy(hc, fid, p, a, b, eps):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = quad(FNS[fid], a, b, args=(p,), epsabs=eps[0], epsrel=eps[1], full_output=1)
    out = np.zeros(5)
    hc.hc_qags(fid, C.c_double(p), C.c_double(a), C.c_double(b), C.c_double(eps[0]), C.c_double(eps[1]), _p(out))
    info = ref[2]
    ier = 0 if len(ref) == 3 else {"maximum number": 1, "roundoff": 2, "bad integrand": 3, "not converge": 4,
                                    "divergent": 5}.get(next((k for k in ("maximum number", "roundoff", "bad integrand",
                                                                         "not converge", "divergent") if k in ref[3]), ""), -1)
    assert int(out[2]) == info["neval"], (out, ref[0], ref[1], info["neval"])
    assert int(out[4]) == info["last"]
    assert int(out[3]) == ier
    assert out[0] == pytest.approx(ref[0], rel=1e-14, abs=1e-300)
    # the error estimate is a difference of two rules: where it is at round-off level it moves with the last bit
    # of libm's cos/exp/pow against numpy's
    assert abs(out[1] - ref[1]) <= 1e-6 * ref[1] + 1e-12 * abs(ref[0])


@pytest.mark.parametrize("fid,p,a,b,fn", [(8, 2.0, 0, 3, lambda x: x ** 3 - 2.0), (8, 1e-9, -1, 1, lambda x: x ** 3 - 1e-9),
                                           (9, 1.0, 0, 2, lambda x: np.cos(x) - x), (9, 30.0, 0, 1, lambda x: np.cos(x) - 30 * x)])
def test_brentq_matches_scipy(hc, fid, p, a, b, fn):
    root, res = brentq(fn, a, b, full_output=True)
    calls = C.c_int(0)
    got = hc.hc_brentq(fid, C.c_double(p), C.c_double(a), C.c_double(b), C.byref(calls))
    assert calls.value == res.function_calls
    assert got == pytest.approx(root, rel=1e-15, abs=1e-300)


def _curve(rng, n, pa_eq):
    """B(s) along a dipole-like field line, irregular spacing, overshooting Bm = Bmin/sin^2(pa_eq) at both ends."""
    Bmin = 1.4e-7
    Bm = Bmin / np.sin(np.radians(pa_eq)) ** 2
    smax = 1.0
    f = lambda s: Bmin * (1 + 3.2 * s ** 2 + 1.1 * s ** 4 + 0.2 * s ** 3)
    while f(smax) < 1.3 * Bm or f(-smax) < 1.3 * Bm:
        smax *= 1.1
    # spacing as a trace produces it: a fixed chunk length, some chunks split by the RKF45 step control
    h = rng.uniform(0.3, 1.7, n); h[rng.uniform(size=n) < 0.15] *= 0.2
    s = np.cumsum(h); s = (s - s[0]) / (s[-1] - s[0]) * 2 * smax - smax
    s = s * 3e7
    b = f(s / 3e7)
    inside = np.where(b <= Bm)[0]
    assert inside[0] >= 1 and inside[-1] <= n - 2
    return s, b, Bm


def test_spline_matches_interp1d(hc):
    rng = np.random.default_rng(5)
    for n in (4, 5, 9, 40, 200):
        s, b, _ = _curve(rng, max(n, 30), 80)
        s, b = s[:n], b[:n]
        B = interp1d(s, b, kind="quadratic", assume_sorted=True)
        x = np.concatenate([rng.uniform(s[0], s[-1], 200), s])
        out = np.zeros(len(x)); coef = np.zeros(n)
        hc.hc_spline(_p(s), _p(b), n, _p(x), len(x), _p(out), _p(coef))
        assert np.max(np.abs(out - B(x)) / np.abs(B(x))) < 5e-15


@pytest.mark.parametrize("pa_eq", [72, 80, 85, 88])
@pytest.mark.parametrize("n", [61, 150, 400])
def test_halfbounce_and_eye_match_the_reference_route(hc, pa_eq, n):
    """flutils.halfbouncepath / eye's scipy route (interp1d + brentq + quad epsrel 1e-4) on synthetic curves."""
    rng = np.random.default_rng(100 * pa_eq + n)
    s, b, Bm = _curve(rng, n, pa_eq)
    # halfbouncepath trimming (<=), flutils.py:276-290
    inside = np.where(b <= Bm)[0]
    i1, i2 = inside[0] - 1, inside[-1] + 1
    ss, bb = s[i1:i2 + 1], b[i1:i2 + 1]
    if len(bb) > 3:
        B = interp1d(ss, bb, kind="quadratic", assume_sorted=True)
        sm1 = brentq(lambda x: B(x) - Bm, ss[0], ss[1]); sm2 = brentq(lambda x: B(x) - Bm, ss[-2], ss[-1])
        ref = quad(lambda x: 1 / np.sqrt(1 - B(x) / Bm), sm1, sm2, epsrel=1e-4)[0]
        got = hc.hc_halfbounce(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), 1)
        # 1/sqrt(1 - B/Bm) is evaluated within ~1e-10 of the mirror points, where 1 - B/Bm has lost most of its
        # digits: the reference's own value is defined to ~1e-9 only
        assert got == pytest.approx(ref, rel=1e-8)
        closed = hc.hc_halfbounce(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), 0)
        assert closed == pytest.approx(ref, rel=1e-4)
        refI = quad(lambda x: np.sqrt(1 - B(x) / Bm), sm1, sm2, epsrel=1e-4)[0]
        err = C.c_int(0)
        gotI = hc.hc_eye(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), C.byref(err))
        assert err.value == 0
        assert gotI == pytest.approx(refI, rel=2e-12)


@pytest.mark.parametrize("N", [3, 4, 5, 8, 17, 18, 64, 257, 258, 700])
def test_simpson_matches_scipy(hc, N):
    rng = np.random.default_rng(N)
    x = np.cumsum(rng.uniform(0.5, 1.5, N)); y = np.sqrt(rng.uniform(0, 1, N))
    got = hc.hc_simpson(_p(y), _p(x), N)
    ref = float(simpson(y, x=x))
    assert got == pytest.approx(ref, rel=4e-16)


@pytest.mark.parametrize("pa_eq", [30, 45, 65])
def test_eye_simpson_branch(hc, pa_eq):
    """eqpa < 70: flutils.py:119-137 with `simps` read as scipy.integrate.simpson(y, x=x)."""
    rng = np.random.default_rng(pa_eq)
    s, b, Bm = _curve(rng, 90, pa_eq)
    inside = np.where(b < Bm)[0]
    n = len(b)
    keep = np.delete(np.arange(n), list(range(0, inside[0] - 1)) + list(range(inside[-1] + 2, n)))
    ss, bb = s[keep].copy(), b[keep].copy()
    sm1 = (Bm - bb[0]) * (ss[1] - ss[0]) / (bb[1] - bb[0]) + ss[0]
    sm2 = (Bm - bb[-2]) * (ss[-1] - ss[-2]) / (bb[-1] - bb[-2]) + ss[-2]
    ss[0], ss[-1] = sm1, sm2
    bb[0], bb[-1] = Bm, Bm
    ref = simpson(np.sqrt(1 - bb[1:-1] / Bm), x=ss[1:-1])
    ref += (2 / 3) * (ss[-1] - ss[-2]) * np.sqrt((Bm - bb[-2]) / Bm)
    ref += (2 / 3) * (ss[1] - ss[0]) * np.sqrt((Bm - bb[1]) / Bm)
    err = C.c_int(0)
    got = hc.hc_eye(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), C.byref(err))
    assert err.value == 0
    assert got == pytest.approx(ref, rel=1e-15)
