"""GPU: the DEVICE build of rapt_quad.cuh (tests/hostcheck/quad_dev.cu) against scipy itself -- the same cases as
tests/test_quad_host.py, executed by the code the kernels run."""
import ctypes as C
import os
import subprocess
import warnings

import numpy as np
import pytest
from scipy.integrate import quad
from scipy.optimize import brentq

import helpers as H
from test_quad_host import FNS, QAGS_CASES, _curve, _p

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dv():
    d = os.path.join(HERE, "hostcheck")
    if not os.path.exists(os.path.join(d, "libquaddev.so")):
        subprocess.check_call(["make", "-C", d, "-s", "libquaddev.so"])
    lib = C.CDLL(os.path.join(d, "libquaddev.so"))
    return lib


@pytest.mark.parametrize("fid,p,a,b", QAGS_CASES)
@pytest.mark.parametrize("eps", [(1.49e-8, 1e-4), (1.49e-8, 1.49e-8), (0.0, 1e-10)])
def test_device_qags_matches_scipy(dv, fid, p, a, b, eps):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = quad(FNS[fid], a, b, args=(p,), epsabs=eps[0], epsrel=eps[1], full_output=1)
    out = np.zeros(5)
    assert dv.dv_qags(fid, C.c_double(p), C.c_double(a), C.c_double(b), C.c_double(eps[0]), C.c_double(eps[1]), _p(out)) == 0
    info = ref[2]
    # device libm (pow, log, cos, exp: <= 2 ulp) against numpy's: the subdivision path is the same unless an error
    # estimate sits at round-off level, where one more or fewer bisection changes nothing at the requested accuracy
    if int(out[2]) == info["neval"]:
        assert out[0] == pytest.approx(ref[0], rel=1e-13, abs=1e-300)
    else:
        assert out[0] == pytest.approx(ref[0], rel=max(10 * eps[1], 1e-12), abs=10 * eps[0] + 1e-300)
        assert abs(int(out[2]) - info["neval"]) <= 42 * 2


@pytest.mark.parametrize("fid,p,a,b,fn", [(8, 2.0, 0, 3, lambda x: x ** 3 - 2.0), (8, 1e-9, -1, 1, lambda x: x ** 3 - 1e-9),
                                           (9, 1.0, 0, 2, lambda x: np.cos(x) - x), (9, 30.0, 0, 1, lambda x: np.cos(x) - 30 * x)])
def test_device_brentq_matches_scipy(dv, fid, p, a, b, fn):
    root, res = brentq(fn, a, b, full_output=True)
    out = np.zeros(2)
    assert dv.dv_brentq(fid, C.c_double(p), C.c_double(a), C.c_double(b), _p(out)) == 0
    assert abs(int(out[1]) - res.function_calls) <= 1
    assert out[0] == pytest.approx(root, rel=1e-14, abs=1e-300)


@pytest.mark.parametrize("pa_eq", [72, 80, 85, 88])
@pytest.mark.parametrize("n", [61, 150, 400])
def test_device_curve_integrals_match_host_build(dv, pa_eq, n):
    """halfbouncepath / eye on synthetic curves: device build vs host build of the same header (the host build is
    pinned against scipy in tests/test_quad_host.py)."""
    hc = C.CDLL(os.path.join(HERE, "hostcheck", "libquadhost.so"))
    hc.hc_halfbounce.restype = C.c_double; hc.hc_eye.restype = C.c_double
    rng = np.random.default_rng(100 * pa_eq + n)
    s, b, Bm = _curve(rng, n, pa_eq)
    out = np.zeros(2); err = C.c_int(0)
    for what, ref, tol in ((0, hc.hc_halfbounce(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), 1), 1e-8),
                           (1, hc.hc_halfbounce(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), 0), 1e-9),
                           (2, hc.hc_eye(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), C.byref(err)), 1e-9)):
        # near-equatorial curves (88 degrees): 1 - B/Bm <= 1e-3 everywhere, so the integrands carry 1e-13 of round-off and
        # the closed form differences two nearly equal primitives; device libm (asin, log) is within 2 ulp of glibc's
        assert dv.dv_curve(_p(s), _p(b), C.c_longlong(n), C.c_double(Bm), what, _p(out)) == 0
        if np.isnan(ref):            # fewer than two points inside the mirror points: the reference's assert fails on both builds
            assert np.isnan(out[0]) and (what != 2 or (out[1] == 1 and err.value == 1))
            continue
        if abs(out[0] / ref - 1) >= tol:
            inside = np.where(b <= Bm)[0] if what < 2 else np.where(b < Bm)[0]
            i1, m = int(inside[0] - 1), int(inside[-1] + 1 - (inside[0] - 1) + 1)
            d_ = np.zeros(13 + 1200); h_ = np.zeros(13 + 1200)
            dv.dv_parts(_p(s), _p(b), C.c_longlong(n), C.c_longlong(i1), m, C.c_double(Bm), 1 if what == 0 else 0, _p(d_))
            dv.hs_parts(_p(s), _p(b), C.c_longlong(n), C.c_longlong(i1), m, C.c_double(Bm), 1 if what == 0 else 0, _p(h_))
            print("what", what, "dev", out[0], "host", ref, "\n device parts", d_[:13], "\n host parts  ", h_[:13])
            np.set_printoptions(linewidth=250, precision=10)
            for k in range(14):
                print(" it", k + 2, "D", d_[13 + 12 * k: 25 + 12 * k]); print(" it", k + 2, "H", h_[13 + 12 * k: 25 + 12 * k])
                print(" ex", k + 2, "D", d_[613 + 12 * k: 625 + 12 * k]); print(" ex", k + 2, "H", h_[613 + 12 * k: 625 + 12 * k])
        assert out[0] == pytest.approx(ref, rel=tol), what
