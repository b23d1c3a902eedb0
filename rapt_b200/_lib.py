"""ctypes binding of librapt_b200.so (include/rapt_b200.h).

The library is the product: if it is missing, or no CUDA device is present, every compute call
raises -- there is no CPU fallback.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAPT_B200_LIB", os.path.join(_HERE, "librapt_b200.so"))   # override: A/B builds

RAPT_OK = 0
FIELD_KIND = {"EarthDipole": 0, "DoubleDipole": 1, "UniformBz": 2, "UniformCrossedEB": 3,
              "VarEarthDipole": 4, "Parabolic": 5, "Grid": 6, "User": 100}
EOM_KIND = {"TaoChanBrizardEOM": 0, "BrizardChanEOM": 1, "NorthropTellerEOM": 2}
ST_OK, ST_ADIABATIC, ST_NONADIABATIC = 1, 2, 3


class FieldT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("is_static", C.c_int32), ("user_id", C.c_int32), ("nprm", C.c_int32),
                ("prm", C.c_double * 16), ("gradstep", C.c_double), ("tstep", C.c_double)]


class ParamsT(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("cyclotronresolution", C.c_double),
                ("epss", C.c_double), ("epst", C.c_double),
                ("enforce_equatorial", C.c_int32), ("check_adiabaticity", C.c_int32),
                ("dop853_reject_rule", C.c_int32), ("arith", C.c_int32), ("sort_by_work", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class RaptB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load librapt_b200.so (no device needed for loading)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RaptB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C rapt_b200/csrc`. rapt_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.rapt_b200_last_error.restype = C.c_char_p
        _lib.rapt_b200_version.restype = C.c_char_p
        _lib.rapt_b200_launch_count.restype = C.c_int64
    return _lib


def check(rc):
    if rc != RAPT_OK:
        raise RaptB200Error(f"librapt_b200 error {rc}: {load().rapt_b200_last_error().decode()}")


def device_count():
    return load().rapt_b200_device_count()


def init(device=0):
    check(load().rapt_b200_init(C.c_int(device)))


def launch_count():
    return int(load().rapt_b200_launch_count())


def ptr(a):
    """void* of a numpy array / torch tensor / int address / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))
