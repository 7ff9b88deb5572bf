"""Shared helpers of the test-suite (golden loader, error norms, seeded states)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
_GEN_PREFIXES = ("sph", "ldw")     # general-grid fixtures (GenOracle / the gen path of the library)
GOLDEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith(_GEN_PREFIXES))
GEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz") if p.stem.startswith(_GEN_PREFIXES))

# north_star tolerances
TOL_STEP = 1e-12     # relative, per step
TOL_RUN = 1e-9       # relative L1 after a full run


def load_golden(name):
    z = np.load(GOLDEN / (name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("recon", "rk", "solver", "ref_config"):
        g[k] = str(g[k])
    g["dims"] = int(g["dims"])
    g["bcs"] = tuple(str(b) for b in g["bcs"])
    for k in ("gamma", "cfl", "cfl_max_var", "first_dt", "tstop"):
        g[k] = float(g[k])
    g["nx"] = tuple(int(x) for x in g["nx"])
    # optional keys of the later fixtures (tracers, body force, non-unit domains)
    g["xbeg"] = tuple(float(x) for x in g["xbeg"]) if "xbeg" in g else (0.0, 0.0, 0.0)
    g["xend"] = tuple(float(x) for x in g["xend"]) if "xend" in g else (1.0, 1.0, 1.0)
    g["ntracer"] = int(g["ntracer"]) if "ntracer" in g else 0
    g["body_force"] = str(g["body_force"]) if "body_force" in g else "none"
    g["grav"] = float(g["grav"]) if "grav" in g else 0.0
    g["limiter"] = str(g["limiter"]) if "limiter" in g else "DEFAULT"
    return g


BODY_FORCE = dict(none=0, vector=1, potential=2)


def gen_kwargs_from_golden(g):
    """Constructor keywords shared by GenOracle and Hydro for a general-grid fixture."""
    grid = []
    for row in g["gridspec"]:
        if row[3] == 1.0:
            grid.append((float(row[0]), int(row[1]), float(row[2]), "r", float(row[4])))
        else:
            grid.append((float(row[0]), int(row[1]), float(row[2])))
    return dict(dimensions=g["dims"], grid=grid, geometry=str(g["geometry"]), gamma=g["gamma"],
                reconstruction=g["recon"], time_stepping=g["rk"], solver=g["solver"], bcs=g["bcs"],
                ntracer=g["ntracer"], limiter=g["limiter"], body_force=BODY_FORCE[g["body_force"]],
                char_limiting=bool(int(g["char_limiting"])), shock_flattening=bool(int(g["shock_flattening"])),
                entropy_switch=bool(int(g["entropy_switch"])))


def set_point_mass_gravity(obj, gm):
    """BodyForceVector of oracle/problems/sph/init.c: g = (-GM/x1^2, 0, 0) at the zone centres."""
    x1 = obj.x(0)
    obj.set_body_force_vector(0, (-gm / (x1 * x1)).reshape(1, 1, -1))
    obj.set_body_force_vector(1, np.zeros((1, 1, 1)))
    obj.set_body_force_vector(2, np.zeros((1, 1, 1)))


def kwargs_from_golden(g):
    return dict(dimensions=g["dims"], nx=g["nx"], gamma=g["gamma"], reconstruction=g["recon"],
                time_stepping=g["rk"], solver=g["solver"], bcs=g["bcs"], xbeg=g["xbeg"], xend=g["xend"],
                ntracer=g["ntracer"], limiter=g["limiter"], body_force=BODY_FORCE[g["body_force"]])


def set_gravity_x2(obj, kind, grav):
    """Hand a constant gravity GRAV along x2 to a Hydro or an Oracle, the way the shim does it for
    oracle/problems/rt/init.c: BodyForceVector = (0, GRAV, 0), BodyForcePotential = -GRAV*x2,
    evaluated on the object's own grid (cell centres x2[j], upper faces x2p[j])."""
    if kind == "none":
        return
    ny = obj.tot[1]
    if kind == "vector":
        for comp, val in enumerate((0.0, grav, 0.0)):
            obj.set_body_force_vector(comp, np.full((1, 1, 1), val))
        return
    dx2 = (obj_xend(obj, 1) - obj_xbeg(obj, 1)) / obj.nx[1] if obj.dimensions > 1 else 1.0
    j = np.arange(ny) - obj.beg[1]
    if obj.dimensions > 1:
        # Src/set_grid.c: xl = xbeg + j dx, xr = xl + dx, x = 0.5 (xl + xr)
        xl = obj_xbeg(obj, 1) + j * dx2
        xr = xl + dx2
    else:   # inactive direction: one zone spanning the ini range
        xl = np.array([obj_xbeg(obj, 1)]); xr = np.array([obj_xend(obj, 1)])
    x2 = 0.5 * (xl + xr)
    phic = (-grav * x2).reshape(1, -1, 1)
    phif = (-grav * xr).reshape(1, -1, 1)
    obj.set_body_force_potential(0, phic)
    obj.set_body_force_potential(1, phic)   # Phi(x1p, x2, x3) = Phi(x2)
    obj.set_body_force_potential(2, phif)
    obj.set_body_force_potential(3, phic)


def obj_xbeg(obj, d):
    return obj.xbeg[d]


def obj_xend(obj, d):
    return obj.xend[d]


def rel_err(a, b):
    """max over variables of max|a-b| / max|b| (velocities share one scale)."""
    a = np.asarray(a); b = np.asarray(b)
    worst = 0.0
    vscale = max(np.abs(b[1:4]).max(), 1e-300)
    for nv in range(b.shape[0]):
        scale = vscale if 1 <= nv <= 3 else max(np.abs(b[nv]).max(), 1e-300)
        worst = max(worst, np.abs(a[nv] - b[nv]).max() / scale)
    return worst


def rel_l1(a, b):
    worst = 0.0
    vscale = max(np.abs(b[1:4]).sum(), 1e-300)
    for nv in range(b.shape[0]):
        scale = vscale if 1 <= nv <= 3 else max(np.abs(b[nv]).sum(), 1e-300)
        worst = max(worst, np.abs(a[nv] - b[nv]).sum() / scale)
    return worst


def random_state(shape_int, seed, smooth=True):
    """Seeded primitive state [5][nz][ny][nx]: smooth waves plus a few jumps (shocks/contacts)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape_int
    z, y, x = np.meshgrid(np.linspace(0, 1, nz, endpoint=False), np.linspace(0, 1, ny, endpoint=False),
                          np.linspace(0, 1, nx, endpoint=False), indexing="ij")
    v = np.empty((5, nz, ny, nx))
    ph = rng.uniform(0, 2 * np.pi, size=(5, 3))
    k = rng.integers(1, 4, size=(5, 3))
    def wave(n):
        return (np.sin(2 * np.pi * k[n, 0] * x + ph[n, 0]) * np.cos(2 * np.pi * k[n, 1] * y * (ny > 1) + ph[n, 1])
                * np.cos(2 * np.pi * k[n, 2] * z * (nz > 1) + ph[n, 2]))
    v[0] = 1.0 + 0.4 * wave(0)
    v[1] = 0.8 * wave(1)
    v[2] = 0.8 * wave(2)
    v[3] = 0.8 * wave(3)
    v[4] = 1.0 + 0.5 * wave(4)
    if not smooth:
        # jumps: a dense over-pressured box and a rarefied slab
        sx = slice(nx // 4, max(nx // 4 + 1, nx // 2)); sy = slice(ny // 4, max(ny // 4 + 1, ny // 2))
        sz = slice(nz // 4, max(nz // 4 + 1, nz // 2))
        v[0][sz, sy, sx] *= 4.0
        v[4][sz, sy, sx] *= 20.0
        v[0][..., (3 * nx) // 4:] *= 0.125
        v[4][..., (3 * nx) // 4:] *= 0.1
        v[1:4] += rng.normal(0, 0.05, size=(3, nz, ny, nx))
    return v


def hydro_kwargs_from_gen(kw):
    """GenOracle keywords -> pluto_sirocco_b200.Hydro keywords (grid arrays from the restated
    set_grid.c of oracle/pluto_grid.py, exactly what the shim passes from the reference's Grid)."""
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    from pluto_grid import make_grid
    kw = dict(kw)
    grid = kw.pop("grid")
    nd = kw["dimensions"]
    ng = kw.get("nghost", 2)
    arrays = [make_grid(grid[d], ng if d < nd else 0) for d in range(3)]
    kw["nx"] = tuple(int(grid[d][1]) if d < nd else 1 for d in range(3))
    kw["xbeg"] = tuple(float(grid[d][0]) for d in range(3))
    kw["xend"] = tuple(float(grid[d][2]) for d in range(3))
    kw["grid_arrays"] = arrays
    return kw
