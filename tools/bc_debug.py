#!/usr/bin/env python
"""One-off probe (GPU box): where do the device's I / S_b differ from the reference's -- in the trace or in the
quadrature?  Device curve vs the oracle's curve point by point; device I/S_b vs the host build of the SAME header
(tests/hostcheck) on the device's curve and on the oracle's curve."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O
from rapt_b200 import engine, fields, _lib

_lib.init(0)
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostcheck"), "-s"])
hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libquadhost.so"))
hc.hc_halfbounce.restype = C.c_double; hc.hc_eye.restype = C.c_double
P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))

d = np.load(os.path.join(ROOT, "tests", "golden", "bc_flutils.npz"))
f = fields.EarthDipole(); of = O.make_field("EarthDipole")
dev = engine.bounce_center_terms(f, d["tpos"], d["Bm"], arith="strict")
for i, (tp, Bm) in enumerate(zip(d["tpos"], d["Bm"])):
    cvd = engine.fieldline_trace(f, tp, Bm, arith="strict")
    cd = cvd["curve"]
    co, Bo, dso = O.fieldline_trace(of, tp, Bm)
    print(f"--- point {i}: npts dev {len(cd)} oracle {len(co)}  ds rel {cvd['ds']/dso-1:.3e}")
    if len(cd) == len(co):
        print("   curve max abs diff (s,x,y,z):", np.max(np.abs(cd[:, :4] - co), axis=0), " B rel:", np.max(np.abs(cd[:, 4] / Bo - 1)))
    err = C.c_int(0)
    for label, s, b in (("dev curve", np.ascontiguousarray(cd[:, 0]), np.ascontiguousarray(cd[:, 4])),
                        ("oracle curve", np.ascontiguousarray(co[:, 0]), np.ascontiguousarray(Bo))):
        Sq = hc.hc_halfbounce(P(s), P(b), C.c_longlong(len(s)), C.c_double(Bm), 1)
        Sc = hc.hc_halfbounce(P(s), P(b), C.c_longlong(len(s)), C.c_double(Bm), 0)
        I = hc.hc_eye(P(s), P(b), C.c_longlong(len(s)), C.c_double(Bm), C.byref(err))
        print(f"   host header on {label}: Sb_quadpack/ref-1 {Sq/d['Sb'][i]-1:.3e}  Sb_closed/ref-1 {Sc/d['Sb'][i]-1:.3e}  I/ref-1 {I/d['I'][i]-1:.3e}")
    print(f"   device kernel:            Sb/ref-1 {dev['Sb'][i]/d['Sb'][i]-1:.3e}  I/ref-1 {dev['I'][i]/d['I'][i]-1:.3e}")
