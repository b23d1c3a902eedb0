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
    """Fields sampled on a Cartesian (rectilinear) grid (rapt/fields.py:513-814).

    Not for direct use: subclass and override `parsefile(filename)`, which must return a dictionary with
    "time" (float), "x", "y", "z" (1-D node coordinates, uniform spacing not required) and the 3-D arrays
    "Bx", "By", "Bz", "Ex", "Ey", "Ez" (SI units) -- the reference's contract (fields.py:553-562).

    `Grid(filelist)` takes the data files in time order.  Two or more files: linear interpolation in
    (t, x, y, z); one file: time-independent, interpolation in (x, y, z).  Differences from the reference,
    both consequences of running on a B200:

    * every file is parsed at construction and all time points are uploaded to HBM
      (`rapt_b200_grid_create`).  The reference keeps a rolling window of three files on the host to save
      memory (fields.py:697-705) and therefore "forgets" earlier times; the interpolated values are the same
      (linear interpolation between the two bracketing time points) but a second tracer started at an
      earlier time works here.
    * the single-file case works (the reference sets `self.time_indep` but tests `self._time_indep`,
      fields.py:593 vs :733, and then calls a 3-D interpolator with four coordinates).

    Outside the grid `B`/`E` raise ValueError like scipy's RegularGridInterpolator does in the reference;
    a tracer that leaves the grid during `advance` keeps the rows computed so far and raises ValueError.
    `static` stays True as in the reference (fields.py:41 is never overridden by Grid): set it to False on the
    instance when dB/dt or E matter for gamma.  Overriding `B`/`E` in a subclass (e.g. to add a dipole) has
    no device counterpart; such a subclass cannot be advanced.
    """
    _device_kind = "Grid"

    def __init__(self, filelist):
        assert len(filelist) > 0
        _Field.__init__(self)
        self.gradientstepsize = 1e-3 * Re                     # fields.py:578
        self.files = []                                        # nothing left to load (reference: remaining file names)
        grids = [self.parsefile(fn) for fn in filelist]
        g0 = grids[0]
        self._t = np.array([float(g["time"]) for g in grids])
        self._x, self._y, self._z = (np.ascontiguousarray(g0[k], dtype=np.float64) for k in ("x", "y", "z"))
        shape = (len(self._x), len(self._y), len(self._z))
        for g in grids:
            for k in ("Bx", "By", "Bz", "Ex", "Ey", "Ez"):
                if np.shape(g[k]) != shape:
                    raise ValueError(f"{k} has shape {np.shape(g[k])}, expected {shape}: the grid must be the same in all files")
        self._B = [np.ascontiguousarray(np.stack([g[k] for g in grids]), dtype=np.float64) for k in ("Bx", "By", "Bz")]
        self._E = [np.ascontiguousarray(np.stack([g[k] for g in grids]), dtype=np.float64) for k in ("Ex", "Ey", "Ez")]
        self._time_indep = len(grids) == 1
        self.t0 = self._t[0]
        if len(grids) > 1:
            self.t1 = self._t[1]
        if len(grids) > 2:
            self.t2 = self._t[2]
        self._grid_id = None

    def parsefile(self, filename):
        """Parse one data file (one time point); override in the subclass (fields.py:596-622)."""
        return dict()

    # ---- device side
    def device_descriptor(self):
        if type(self).B is not Grid.B or type(self).E is not Grid.E:
            raise NotImplementedError(f"{type(self).__name__} overrides B/E of fields.Grid; only the interpolated grid "
                                      "field has a device implementation (rapt_b200 has no CPU fallback)")
        if self._grid_id is None:
            from . import engine
            self._grid_id = engine.create_grid(self._t, self._x, self._y, self._z, self._B, self._E)
        f = FieldT()
        f.kind = FIELD_KIND["Grid"]
        f.user_id = self._grid_id
        f.nprm = 0
        f.is_static = int(bool(self.static))
        f.gradstep = float(self.gradientstepsize)
        f.tstep = float(self.timederivstepsize)
        return f

    def __del__(self):
        gid = getattr(self, "_grid_id", None)
        if gid is not None:
            try:
                from . import engine
                engine.destroy_grid(gid)
            except Exception:
                pass

    # ---- host side: scipy's linear RegularGridInterpolator restated for one point (the reference builds six
    # of them, fields.py:643-694); used by the host-side API only, never by advance()
    @staticmethod
    def _interval(g, v, axis):
        if not (g[0] <= v <= g[-1]):
            raise ValueError(f"One of the requested xi is out of bounds in dimension {axis}")
        i = int(np.searchsorted(g, v, side="right")) - 1
        return min(max(i, 0), len(g) - 2)

    def _interp(self, comps, tpos):
        axes = ([] if self._time_indep else [self._t]) + [self._x, self._y, self._z]
        vals = list(tpos[1:]) if self._time_indep else list(tpos)
        idx, w = [], []
        for d, (g, v) in enumerate(zip(axes, vals)):
            i = self._interval(g, float(v), d)
            idx.append(i); w.append((float(v) - g[i]) / (g[i + 1] - g[i]))
        nd = len(axes)
        out = np.zeros(3)
        for c in range(3):
            arr = comps[c][0] if self._time_indep else comps[c]
            value = 0.0
            for corner in range(1 << nd):
                weight = 1.0; ii = []
                for d in range(nd):
                    up = (corner >> (nd - 1 - d)) & 1
                    weight = weight * (w[d] if up else 1 - w[d])
                    ii.append(idx[d] + up)
                value = value + arr[tuple(ii)] * weight
            out[c] = value
        return out

    def Bgrid(self, tpos):
        """Interpolated magnetic field vector (fields.py:707-741)."""
        return self._interp(self._B, np.asarray(tpos, dtype=float))

    def Egrid(self, tpos):
        """Interpolated electric field vector (fields.py:743-772)."""
        return self._interp(self._E, np.asarray(tpos, dtype=float))

    def B(self, tpos):
        return self.Bgrid(tpos)

    def E(self, tpos):
        return self.Egrid(tpos)
