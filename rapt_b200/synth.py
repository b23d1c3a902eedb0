"""Synthetic ensembles of the named BASELINE.json shapes.

Pure numpy, no device code: these generators only make *initial conditions*.  They are shared
by bench.py, the parity tests and oracle/gen_golden.py so that the golden fixtures, the oracle
and the CUDA path all see bit-identical inputs (`numpy.random.default_rng(seed)`).

Every generator returns a dict of float64 arrays in SI units with t0 = 0.
"""
import numpy as np

# constants as in the reference's rapt/__init__.py:5-10
e = 1.602176565e-19
m_pr = 1.672621777e-27
m_el = 9.10938291e-31
c = 299792458
B0 = 3.07e-5
Re = 6378137


def speed_from_ke(ke_ev, mass):
    """Vectorised form of rapt/utils.py:218-249 (speedfromKE, unit='ev')."""
    ke = np.asarray(ke_ev, dtype=np.float64) * e
    mc2 = mass * c ** 2
    nonrel = np.sqrt(2 * ke / mass)
    rel = c * np.sqrt(1 - (mc2 / (mc2 + ke)) ** 2)
    return np.where(ke / mc2 < 1e-6, nonrel, rel)


def _dipole_b(x, y, z):
    """Unit vector of the zero-tilt Earth dipole (direction only)."""
    r2 = x * x + y * y + z * z
    bx, by, bz = 3 * x * z, 3 * y * z, 2 * z * z - x * x - y * y
    # field = -B0 Re^3 (bx,by,bz)/r^5 : direction is minus the bracket
    n = np.sqrt(bx * bx + by * by + bz * bz)
    return -bx / n, -by / n, -bz / n


def _perp_basis(bx, by, bz):
    """Two unit vectors perpendicular to b (deterministic)."""
    # u = b x zhat (or b x xhat where b ~ zhat), w = b x u
    ux, uy, uz = by, -bx, np.zeros_like(bx)
    n = np.sqrt(ux * ux + uy * uy + uz * uz)
    small = n < 1e-6
    ux = np.where(small, 0.0, ux); uy = np.where(small, bz, uy); uz = np.where(small, -by, uz)
    n = np.sqrt(ux * ux + uy * uy + uz * uz)
    ux, uy, uz = ux / n, uy / n, uz / n
    wx = by * uz - bz * uy
    wy = bz * ux - bx * uz
    wz = bx * uy - by * ux
    return (ux, uy, uz), (wx, wy, wz)


def config2_protons(n, seed=20260201):
    """Config 2: protons, EarthDipole, full orbit.  L~U[2,6], KE log-uniform 0.1-10 MeV,
    pitch angle U[20,160] deg, gyrophase U[0,2pi), z jitter so no coordinate is exactly 0."""
    rng = np.random.default_rng(seed)
    L = rng.uniform(2.0, 6.0, n)
    az = rng.uniform(0.0, 2 * np.pi, n)
    zj = rng.uniform(-0.05, 0.05, n) * Re
    ke = 10 ** rng.uniform(np.log10(0.1e6), np.log10(10e6), n)
    pa = np.deg2rad(rng.uniform(20.0, 160.0, n))
    ph = rng.uniform(0.0, 2 * np.pi, n)
    x, y, z = L * Re * np.cos(az), L * Re * np.sin(az), zj
    v = speed_from_ke(ke, m_pr)
    bx, by, bz = _dipole_b(x, y, z)
    (ux, uy, uz), (wx, wy, wz) = _perp_basis(bx, by, bz)
    cp, sp = np.cos(pa), np.sin(pa)
    vx = v * (cp * bx + sp * (np.cos(ph) * ux + np.sin(ph) * wx))
    vy = v * (cp * by + sp * (np.cos(ph) * uy + np.sin(ph) * wy))
    vz = v * (cp * bz + sp * (np.cos(ph) * uz + np.sin(ph) * wz))
    return dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, t0=np.zeros(n),
                mass=np.full(n, m_pr), charge=np.full(n, e), ke_ev=ke)


def config3_electrons(n, seed=20260301):
    """Config 3: electrons, DoubleDipole, guiding centre.  r~U[6,10] Re, x <= 8 Re,
    KE log-uniform 50 keV-1 MeV, pa U[30,90] deg."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(6.0, 10.0, n) * Re
    az = rng.uniform(0.0, 2 * np.pi, n)
    # keep x <= 8 Re (inside the x = 10 Re mirror plane): reflect offending azimuths
    x = r * np.cos(az)
    az = np.where(x > 8 * Re, np.pi - az, az)
    x, y = r * np.cos(az), r * np.sin(az)
    z = rng.uniform(-0.05, 0.05, n) * Re
    ke = 10 ** rng.uniform(np.log10(50e3), np.log10(1e6), n)
    pa = rng.uniform(30.0, 90.0, n)
    v = speed_from_ke(ke, m_el)
    return dict(x=x, y=y, z=z, v=v, pa=pa, t0=np.zeros(n),
                mass=np.full(n, m_el), charge=np.full(n, -e), ke_ev=ke)


def config4_speiser(n, seed=20260401):
    """Config 4: Adaptive Speiser orbits in Parabolic(); member 0 is the notebook IC
    (examples/Adaptive Example - Speiser orbits.ipynb cell 7)."""
    rng = np.random.default_rng(seed)
    jit = rng.uniform(-0.05, 0.05, (n, 3))
    spd = 0.1 * np.sqrt(2.0) * rng.uniform(0.8, 1.2, n)
    rot = np.deg2rad(rng.uniform(-10.0, 10.0, n))
    x = 5.0 + jit[:, 0]; y = -5.0 + jit[:, 1]; z = 0.9 + jit[:, 2]
    d0x, d0y = -1.0 / np.sqrt(2.0), 1.0 / np.sqrt(2.0)
    vx = spd * (np.cos(rot) * d0x - np.sin(rot) * d0y)
    vy = spd * (np.sin(rot) * d0x + np.cos(rot) * d0y)
    vz = np.zeros(n)
    if n > 0:
        x[0], y[0], z[0] = 5.0, -5.0, 0.9
        vx[0], vy[0], vz[0] = -0.1, 0.1, 0.0
    return dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, t0=np.zeros(n),
                mass=np.ones(n), charge=np.ones(n))


def config5_belt(n, seed=20260501):
    """Config 5: radiation-belt electrons, VarEarthDipole(amp 0.1, period 10), guiding centre.
    L U[3,7], KE log-uniform 0.1-5 MeV, pa U[20,90] deg."""
    rng = np.random.default_rng(seed)
    L = rng.uniform(3.0, 7.0, n)
    az = rng.uniform(0.0, 2 * np.pi, n)
    z = rng.uniform(-0.05, 0.05, n) * Re
    ke = 10 ** rng.uniform(np.log10(0.1e6), np.log10(5e6), n)
    pa = rng.uniform(20.0, 90.0, n)
    x, y = L * Re * np.cos(az), L * Re * np.sin(az)
    v = speed_from_ke(ke, m_el)
    return dict(x=x, y=y, z=z, v=v, pa=pa, t0=np.zeros(n),
                mass=np.full(n, m_el), charge=np.full(n, -e), ke_ev=ke)


def dipole_grid_slice(k, nx=21, ny=17, nz=25):
    """One synthetic "data file" for fields.Grid, in the dictionary form Grid.parsefile must return
    (rapt/fields.py:553-562): a dipole scaled by (1 + 0.02 k) plus a dawn-dusk electric field, sampled on
    x in [3, 8] Re, y in [-2, 2] Re, z in [-3, 3] Re at time k seconds.  Only IEEE basic operations and
    sqrt, so every machine regenerates the same bits (the golden fixtures store a checksum)."""
    from . import Re, B0
    x = np.linspace(3.0, 8.0, nx) * Re
    y = np.linspace(-2.0, 2.0, ny) * Re
    z = np.linspace(-3.0, 3.0, nz) * Re
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    r2 = X * X + Y * Y + Z * Z
    r5 = r2 * r2 * np.sqrt(r2)
    s = -B0 * (Re * Re * Re) * (1.0 + 0.02 * k)
    zero = np.zeros_like(X)
    return {"time": float(k), "x": x, "y": y, "z": z,
            "Bx": s * (3.0 * X * Z) / r5, "By": s * (3.0 * Y * Z) / r5, "Bz": s * (2.0 * Z * Z - X * X - Y * Y) / r5,
            "Ex": zero.copy(), "Ey": 1e-4 * (1.0 + 0.1 * k) * (X / Re) / 6.0, "Ez": zero.copy()}
