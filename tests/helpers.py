"""Shared helpers for the parity tests."""
import json
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# golden name -> (field class name, constructor args) for both the oracle and rapt_b200.fields
PARTICLE_CASES = {
    "g1_readme": ("EarthDipole", ()),
    "g1b_generic": ("EarthDipole", ()),
    "p_equatorial": ("EarthDipole", ()),
    "p_vardipole": ("VarEarthDipole", (0.1, 10)),
    "p_crossedeb": ("UniformCrossedEB", (2.0, 1e-4)),
    "p_uniformbz": ("UniformBz", (2e-4,)),
    "p_parabolic": ("Parabolic", ()),
}
GC_CASES = {
    "g2_gc_doubledipole": ("DoubleDipole", ()),
    "gc_earthdipole": ("EarthDipole", ()),
    "gc_eom_taochanbrizard": ("DoubleDipole", ()),
    "gc_eom_brizardchan": ("DoubleDipole", ()),
    "gc_eom_northropteller": ("DoubleDipole", ()),
    "gc_pa90_equatorial": ("DoubleDipole", ()),
    "gc_vardipole": ("VarEarthDipole", (0.1, 10)),
    "gc_crossedeb": ("UniformCrossedEB", (2.0, 1e-4)),
    "gc_equatorial_enforced": ("DoubleDipole", ()),
}
ADAPTIVE_CASES = {
    "g3_speiser": ("Parabolic", ()),
    "e4_speiser_1": ("Parabolic", ()), "e4_speiser_2": ("Parabolic", ()), "e4_speiser_3": ("Parabolic", ()),
    "e4_speiser_4": ("Parabolic", ()), "e4_speiser_5": ("Parabolic", ()),
    "adaptive_dipole": ("EarthDipole", ()),
}


def load(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = json.loads(str(d["params"])) if "params" in d.files else {}
    if "solvertolerances" in par:
        par["solvertolerances"] = tuple(par["solvertolerances"])
    return d, par


def relerr(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over all elements."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(b), floor if floor > 0 else 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def vec_relerr(a, b):
    """|a-b| / |b| for 3-vectors along the last axis (max over leading axes)."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


def gpu_field(name, args):
    from rapt_b200 import fields
    return getattr(fields, name)(*args)


def oracle_field(name, args):
    import oracle as O
    return O.make_field(name, *args)


def synthetic_grid(files=("0", "1", "2", "3")):
    """The parsed synthetic data files of the Grid fixtures (rapt_b200/synth.py:dipole_grid_slice), stacked
    the way Grid._set_interpolator stacks them: t (nt,), x, y, z, B and E as 3 arrays (nt, nx, ny, nz);
    plus the sha256 the golden generator recorded for the same arrays."""
    import hashlib
    from rapt_b200 import synth
    gs = [synth.dipole_grid_slice(int(fn)) for fn in files]
    h = hashlib.sha256()
    for g in gs:
        for k in ("x", "y", "z", "Bx", "By", "Bz", "Ex", "Ey", "Ez"):
            h.update(np.ascontiguousarray(g[k]).tobytes())
    t = np.array([g["time"] for g in gs])
    B = [np.stack([g[k] for g in gs]) for k in ("Bx", "By", "Bz")]
    E = [np.stack([g[k] for g in gs]) for k in ("Ex", "Ey", "Ez")]
    return dict(t=t, x=gs[0]["x"], y=gs[0]["y"], z=gs[0]["z"], B=B, E=E, sha256=h.hexdigest(), slices=gs)
