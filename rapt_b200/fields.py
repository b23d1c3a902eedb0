"""Field-model plugin interface (reference: rapt/fields.py).

A field object provides `B(tpos)` and `E(tpos)` (4-vector (t,x,y,z) in, 3-vector out, SI, Cartesian)
plus the attributes `static`, `gradientstepsize`, `timederivstepsize` and the derived operators
`unitb, magB, gradB, jacobianB, curvature, curlb, dBdt, dbdt, lengthscale, timescale`
(rapt/fields.py:38-41, 76-280).  The host-side methods below keep that interface for user code;
the `.advance()` hot path never calls them -- it runs the same formulas as inlined device functions
(rapt_b200/csrc/rapt_fields.cuh), selected through `device_descriptor()`.

User-defined analytic fields: subclass `_Field`, override `B`/`E` for host-side use, and give the
class a `cuda_source` string defining

    __device__ void rapt_user_B(double t, double x, double y, double z, const double* prm, double* B);
    __device__ void rapt_user_E(...same...);      // optional; set `cuda_has_E = True`

plus `cuda_params()` returning the parameter vector `prm`.  The snippet is JIT-compiled with NVRTC into
the same kernel templates as the built-ins.  A subclass without a snippet cannot be advanced (there is
no CPU fallback) and raises.
"""
import numpy as np
from . import Re, B0
from ._lib import FieldT, FIELD_KIND


class _Field:
    """Superclass for fields (rapt/fields.py:5-280)."""

    # curl by central differences: rows pick +-components of the six shifted unit vectors (fields.py:33-36)
    _M1 = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, -1, 0, 0, 1, 0],
                    [0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0],
                    [0, 1, 0, 0, -1, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]])

    cuda_source = None      # user snippet (see module docstring)
    cuda_has_E = False
    _device_kind = None     # built-in kind name

    def __init__(self):
        self.gradientstepsize = 1e-6
        self.timederivstepsize = 1e-3
        self.static = True

    # ---- device side
    def cuda_params(self):
        return []

    def device_descriptor(self):
        """POD snapshot (rapt_field_t) of this field for the C ABI; taken at every advance() call."""
        f = FieldT()
        if self._device_kind is not None:
            f.kind = FIELD_KIND[self._device_kind]
        elif self.cuda_source is not None:
            from . import engine
            f.kind = FIELD_KIND["User"]
            f.user_id = engine.compile_user_field(self.cuda_source, bool(self.cuda_has_E))
        else:
            raise NotImplementedError(
                f"{type(self).__name__} has no device implementation: built-in analytic fields or a "
                "`cuda_source` snippet are required (rapt_b200 has no CPU fallback)")
        prm = [float(v) for v in self.cuda_params()]
        if len(prm) > 16:
            raise ValueError("at most 16 field parameters")
        f.nprm = len(prm)
        for i, v in enumerate(prm):
            f.prm[i] = v
        f.is_static = int(bool(self.static))
        f.gradstep = float(self.gradientstepsize)
        f.tstep = float(self.timederivstepsize)
        return f

    # ---- host side (API surface; same formulas as the reference)
    def B(self, tpos):
        return np.zeros(3)

    def E(self, tpos):
        return np.zeros(3)

    def unitb(self, tpos):
        Bvec = self.B(tpos)
        return Bvec / np.sqrt(np.dot(Bvec, Bvec))

    def magB(self, tpos):
        Bvec = self.B(tpos)
        return np.sqrt(np.dot(Bvec, Bvec))

    def _shift(self, tpos, axis, d):
        q = np.array(tpos, dtype=float)
        q[axis] = q[axis] + d
        return q

    def gradB(self, tpos):
        d = self.gradientstepsize
        return np.array([(self.magB(self._shift(tpos, i, d)) - self.magB(self._shift(tpos, i, -d))) / (2 * d)
                         for i in (1, 2, 3)])

    def jacobianB(self, tpos):
        d = self.gradientstepsize
        J = np.zeros((3, 3))
        for j in (1, 2, 3):
            J[:, j - 1] = (self.B(self._shift(tpos, j, d)) - self.B(self._shift(tpos, j, -d))) / (2 * d)
        return J

    def curvature(self, tpos):
        # reference quirk (fields.py:173): np.dot(gB, |B|) is element-wise, not a projection
        Bvec = self.B(tpos)
        Bm = np.sqrt(np.dot(Bvec, Bvec))
        gB = self.gradB(tpos)
        gBperp = gB - (gB * Bm / Bm ** 2) * Bvec
        return np.sqrt(np.dot(gBperp, gBperp)) / Bm

    def curlb(self, tpos):
        d = self.gradientstepsize
        beta = np.concatenate([self.unitb(self._shift(tpos, j, s * d)) for j in (1, 2, 3) for s in (1, -1)])
        return np.dot(self._M1, beta) / (2 * d)

    def dBdt(self, tpos):
        if self.static:
            return 0
        d = self.timederivstepsize
        return (self.magB(self._shift(tpos, 0, d)) - self.magB(self._shift(tpos, 0, -d))) / d / 2

    def dbdt(self, tpos):
        if self.static:
            return 0
        d = self.timederivstepsize
        return (self.unitb(self._shift(tpos, 0, d)) - self.unitb(self._shift(tpos, 0, -d))) / d / 2

    def lengthscale(self, tpos):
        with np.errstate(divide="ignore"):
            return self.magB(tpos) / np.max(abs(self.jacobianB(tpos)))

    def timescale(self, tpos):
        if self.static:
            return None
        with np.errstate(divide="ignore"):
            return self.magB(tpos) / abs(self.dBdt(tpos))


class EarthDipole(_Field):
    """Earth's static dipole, zero tilt (rapt/fields.py:282-317)."""
    _device_kind = "EarthDipole"

    def __init__(self, B0=B0):
        _Field.__init__(self)
        self.gradientstepsize = Re * 1e-6
        self._coeff = -3 * B0 * Re ** 3

    def cuda_params(self):
        return [self._coeff]

    def B(self, tpos):
        t, x, y, z = tpos
        r2 = x * x + y * y + z * z
        return self._coeff / pow(r2, 2.5) * np.array([x * z, y * z, (z * z - r2 / 3)])


class DoubleDipole(_Field):
    """Two parallel Earth dipoles, the image at x = distance (rapt/fields.py:319-362)."""
    _device_kind = "DoubleDipole"

    def __init__(self, B0=B0, distance=20 * Re, imagestrength=1):
        _Field.__init__(self)
        self.gradientstepsize = Re / 1000
        self._dd = distance
        assert imagestrength >= 1
        self._k = imagestrength
        self._coeff = -B0 * Re ** 3

    def cuda_params(self):
        return [self._coeff, self._dd, self._k]

    def B(self, tpos):
        t, x, y, z = tpos
        B1 = np.array([3 * x * z, 3 * y * z, (2 * z * z - x * x - y * y)]) / pow(x * x + y * y + z * z, 5.0 / 2.0)
        x = x - self._dd
        B2 = self._k * np.array([3 * x * z, 3 * y * z, (2 * z * z - x * x - y * y)]) / pow(x * x + y * y + z * z, 5.0 / 2.0)
        return self._coeff * (B1 + B2)


class UniformBz(_Field):
    """Uniform static field B = (0,0,Bz) (rapt/fields.py:364-390)."""
    _device_kind = "UniformBz"

    def __init__(self, Bz=1):
        _Field.__init__(self)
        self.Bz = Bz

    def cuda_params(self):
        return [self.Bz]

    def B(self, tpos):
        return np.array((0, 0, self.Bz))


class UniformCrossedEB(UniformBz):
    """E = (0,Ey,0), B = (0,0,Bz); static = False (rapt/fields.py:392-427)."""
    _device_kind = "UniformCrossedEB"

    def __init__(self, Ey=1, Bz=1):
        UniformBz.__init__(self)
        self.static = False
        self.Ey = Ey
        self.Bz = Bz

    def cuda_params(self):
        return [self.Bz, self.Ey]

    def E(self, tpos):
        return np.array((0, self.Ey, 0))


class VarEarthDipole(_Field):
    """Earth dipole whose moment oscillates sinusoidally; induced E ignored (rapt/fields.py:429-470)."""
    _device_kind = "VarEarthDipole"

    def __init__(self, amp=0.1, period=10):
        _Field.__init__(self)
        self.gradientstepsize = Re / 1000
        self.static = False
        self._amp = amp
        self._period = period

    def cuda_params(self):
        return [self._amp, self._period]

    def B(self, tpos):
        t, x, y, z = tpos
        return -B0 * Re ** 3 * (1 + self._amp * np.sin(2 * np.pi * t / self._period)) * \
            np.array([3 * x * z, 3 * y * z, (2 * z * z - x * x - y * y)]) / pow(x * x + y * y + z * z, 5.0 / 2.0)


class Parabolic(_Field):
    """Parabolic current-sheet model (rapt/fields.py:472-511).  As in the reference, |z| > 1 uses the
    module-level Earth B0 for Bx (not self.B0)."""
    _device_kind = "Parabolic"

    def __init__(self, B0=10.0, Bn=1.0, d=0.2):
        _Field.__init__(self)
        self.B0 = B0
        self.Bn = Bn
        self.d = d

    def cuda_params(self):
        return [self.B0, self.Bn, self.d]

    def B(self, tpos):
        z = tpos[3]
        if abs(z) <= 1.0:
            return np.array([self.B0 * z / self.d, 0, self.Bn])
        return np.array([np.sign(z) * B0, 0, self.Bn])


class Grid(_Field):
    """Gridded-data fields (rapt/fields.py:513-814) are outside the device hot path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("fields.Grid is not part of the B200 hot path (analytic fields and NVRTC "
                                  "snippets only); see DESIGN.md 'out of scope'")
