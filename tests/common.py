"""Shared helpers of the test-suite (golden loader, error norms, seeded states)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
_GEN_PREFIXES = ("sph", "ldw", "cart")     # general-grid fixtures (GenOracle / the gen path of the library)
# cylindrical / polar / Roe / two-shock / ONED / general-grid PPM fixtures pin the ORACLE only (the CUDA path refuses
# these options); the iso* fixtures (EOS ISOTHERMAL) run on the CUDA path too: ISO_CASES
_CURV_PREFIXES = ("cyl", "pol", "iso", "roe", "twoshock", "oned", "ppmg", "pot", "ausm", "ring")
ISO_CASES = ["iso2d_hll", "iso2d_hllc", "iso2d_flat_hllc", "iso3d_tvdlf", "iso_sph2d_flat_hll"]
# CYLINDRICAL / POLAR geometry and BODY_FORCE POTENTIAL on curvilinear grids: on the CUDA path as well
CURV_GPU_CASES = ["cyl2d_axis_hllc", "cyl2d_flat_tvdlf", "cyl2d_grav_hll", "pol2d_hllc", "pol3d_hll", "pot_sph2d_hllc",
                  "pot_sph3d_both_hll",
                  # Roe_Solver (both equations of state) and TwoShock_Solver
                  "roe_cart2d", "roe_iso2d", "roe_sph2d_flat", "pot_pol2d_roe", "twoshock_sph2d_flat", "twoshock_sph3d", "ausm_sph2d",
                  # SHOCK_FLATTENING ONED (States/flatten.c)
                  "oned_iso2d_hll", "oned_sph2d_hllc", "oned_sph2d_char_roe",
                  # RECONSTRUCTION PARABOLIC + RK3 with the general-grid weights of States/ppm_coeffs.c
                  "ppmg_cyl2d_flat_stretched", "ppmg_cyl2d_uniform", "ppmg_iso2d_char", "ppmg_kh3d_stretched", "ppmg_pol2d",
                  "ppmg_sph2d_char_flat", "ppmg_sph2d_stretched", "ppmg_sph2d_uniform", "ppmg_sph3d",
                  # RING_AVERAGE (Src/ring_average.c) with POLARAXIS boundaries
                  "ring_pol2d_mp5", "ring_pol2d_vl", "ring_pol3d_mp5", "ring_sph3d_mp5"]
CURV_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz") if p.stem.startswith(_CURV_PREFIXES))
GOLDEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz")
                      if not p.stem.startswith(_GEN_PREFIXES + _CURV_PREFIXES))
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


BODY_FORCE = dict(none=0, vector=1, potential=2, both=3)


def _entr_code(g):
    """fixtures made before SELECTIVE existed stored 0/1 for NO/ALWAYS"""
    v = int(g["entropy_switch"])
    return v if "entr_codes" in g or v != 1 else 2


def gen_kwargs_from_golden(g):
    """Constructor keywords shared by GenOracle and Hydro for a general-grid fixture."""
    grid = []
    for row in g["gridspec"]:
        if row[3] == 1.0:
            grid.append((float(row[0]), int(row[1]), float(row[2]), "r", float(row[4])))
        else:
            grid.append((float(row[0]), int(row[1]), float(row[2])))
    flat = int(g["shock_flattening"])      # 0 NO, 1 MULTID, 2 ONED
    extra = {}
    if "eos" in g and str(g["eos"]) == "ISOTHERMAL":     # only the iso* fixtures carry these keys
        extra = dict(eos="ISOTHERMAL", iso_sound_speed=float(g["iso_cs"]))
    if "ring_average" in g:     # RING_AVERAGE fixtures
        extra.update(ring_average=int(g["ring_average"]), ring_average_rec=int(g["ring_rec"]))
    return dict(**extra, dimensions=g["dims"], grid=grid, geometry=str(g["geometry"]), gamma=g["gamma"],
                reconstruction=g["recon"], time_stepping=g["rk"], solver=g["solver"], bcs=g["bcs"],
                ntracer=g["ntracer"], limiter=g["limiter"], body_force=BODY_FORCE[g["body_force"]],
                char_limiting=bool(int(g["char_limiting"])), shock_flattening={0: False, 1: True, 2: "ONED"}[flat],
                entropy_switch={0: False, 1: "SELECTIVE", 2: "ALWAYS"}[_entr_code(g)],
                nghost=max({0: 2, 1: 3, 2: 4}[flat], 3 if g["recon"] == "PARABOLIC" else 2,
                           3 if ("ring_rec" in g and int(g["ring_rec"]) > 2) else 2))     # GetNghost(), Src/get_nghost.c:37-57


def set_point_mass_gravity(obj, gm):
    """BodyForceVector of oracle/problems/sph/init.c: g = (-GM/x1^2, 0, 0) at the zone centres."""
    bf = getattr(obj, "cfg", getattr(obj, "c", None)).body_force
    if not bf:
        return
    x1 = obj.x(0)
    if bf & 1:
        obj.set_body_force_vector(0, (-gm / (x1 * x1)).reshape(1, 1, -1))
        obj.set_body_force_vector(1, np.zeros((1, 1, 1)))
        obj.set_body_force_vector(2, np.zeros((1, 1, 1)))
    if bf & 2:     # BodyForcePotential = -GM/x1 at the centres, the x1 upper faces, and (x1 centre) the x2 / x3 faces
        obj.set_body_force_potential(0, (-gm / x1).reshape(1, 1, -1))
        obj.set_body_force_potential(1, (-gm / obj.xr[0]).reshape(1, 1, -1))
        obj.set_body_force_potential(2, (-gm / x1).reshape(1, 1, -1))
        obj.set_body_force_potential(3, (-gm / x1).reshape(1, 1, -1))


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
    kw["grid_uniform"] = tuple(len(grid[d]) <= 3 or grid[d][3] == "u" for d in range(3))   # grid->uniform[d]
    return kw


# ---------------------------------------------------------------------------------------------
#  Line-driven wind (Test_Problems/LineDrivenWind/cv_idl): parameters and synthetic sirocco tables
# ---------------------------------------------------------------------------------------------
MSUN = 1.98840987e33
LDW_PARAMS = dict(MU=0.6, RHO_0=1e-9, R_0=8.31e8, RHO_ALPHA=0.0, CENT_MASS=0.6 * MSUN,
                  DISK_MDOT=3.14e-8 * MSUN / (365.25 * 24 * 3600), T_ISO=40000.0, L_star=9.05e34, f_x=0.1,
                  f_uv=0.9, T_x=160000.0, KRAD=0.59, ALPHARAD=-0.6, DFLOOR=1e-24, GAMMA=5.0 / 3.0)
LDW_UNITS = dict(density=1e-14, length=1e9, velocity=1e9)
LDW_BCS = ("userdef", "outflow", "userdef", "reflective", "outflow", "outflow")
LDW_NANGLES = 36


def ldw_flux_tables(x1, x2, nangles=LDW_NANGLES, roundtrip=True):
    """Synthetic directional UV fluxes in the layout of flux_{r,t,p}_UV[iangle][k][j][i]
    (line_connect.c:97-110), for ALL zones of the 1-D coordinate arrays given (the reference only
    fills the interior; ghost entries are never read).  Bin a points along (sin th_a, cos th_a) in
    the (x, z) plane with th_a = (a + 1/2) 2 pi / nangles (line_connect.c:563,654); the flux a disc
    + central source would send through it is modelled as F0(r) w_a(theta), zero in the bins
    pointing back at the source (so that the "no flux" branch, dvds = -999, is exercised)."""
    r = np.asarray(x1)[None, None, :] * LDW_UNITS["length"]
    th = np.asarray(x2)[None, :, None]
    a = (np.arange(nangles) + 0.5) * (2.0 * np.pi) / nangles
    da = a[:, None, None] - th
    w = np.clip(np.cos(da), 0.0, None) ** 3 * (1.0 + 0.3 * np.sin(3.0 * a[:, None, None]))
    f0 = LDW_PARAMS["f_uv"] * LDW_PARAMS["L_star"] / (4.0 * np.pi * r * r) / 6.0
    fr = f0 * w * np.cos(da)
    ft = f0 * w * np.sin(da)
    shape = (nangles, 1, th.shape[1], r.shape[2])
    # round-trip through the text format of the flux files so both sides hold the same doubles
    if not roundtrip:
        return fr.reshape(shape), ft.reshape(shape), np.zeros(shape)
    rt = np.vectorize(lambda q: float("%.17e" % q))
    return rt(fr).reshape(shape), rt(ft).reshape(shape), np.zeros(shape)


def write_ldw_flux_files(wd, x1, x2, ng, fr, ft, fp):
    """directional_flux_{r,theta,phi}.dat in the format read by read_sirocco_fluxes()
    (line_connect.c:84-165): two header lines (the 2nd ends with the number of angular bins), then
    `i j inwind r[cm] theta[rad] f0 ... f{n-1}` per interior zone."""
    nang = fr.shape[0]
    for name, tab in (("r", fr), ("theta", ft), ("phi", fp)):
        with open(Path(wd) / ("directional_flux_%s.dat" % name), "w") as f:
            f.write("# synthetic directional fluxes\n# NANGLES %d\n" % nang)
            for j in range(ng, len(x2) - ng):
                for i in range(ng, len(x1) - ng):
                    vals = " ".join("%.17e" % tab[a, 0, j, i] for a in range(nang))
                    f.write("%d %d 0 %.17e %.17e %s\n" % (i - ng, j - ng, x1[i] * LDW_UNITS["length"], x2[j], vals))


def ldw_setup(obj, x1, x2, fit=False):
    """Hand the cv_idl problem to a GenOracle or a Hydro: gravity of the central mass
    (BodyForceVector, init.c:372-386), units, parameters and the synthetic flux tables."""
    gm_code = 6.6726e-8 * LDW_PARAMS["CENT_MASS"] / (LDW_UNITS["length"] * LDW_UNITS["velocity"] ** 2)
    obj.set_body_force_vector(0, (-1.0 * gm_code / (x1 * x1)).reshape(1, 1, -1))
    obj.set_body_force_vector(1, np.zeros((1, 1, 1)))
    obj.set_body_force_vector(2, np.zeros((1, 1, 1)))
    fr, ft, fp = ldw_flux_tables(x1, x2)
    if fit:
        t, M, lt, lM = ldw_mfit_tables(x1, x2)
        obj.set_ldw(params=dict(LDW_PARAMS, KRAD=999.0, ALPHARAD=999.0), units=LDW_UNITS, flux_r=fr, flux_t=ft,
                    flux_p=fp, t_fit=lt, m_fit=lM)
    else:
        obj.set_ldw(params=LDW_PARAMS, units=LDW_UNITS, flux_r=fr, flux_t=ft, flux_p=fp)


LDW_MPOINTS = 12


def ldw_mfit_tables(x1, x2, mpoints=LDW_MPOINTS):
    """Synthetic force-multiplier fit M(t) per zone for the KRAD = ALPHARAD = 999 mode of LineForce()
    (M_UV_data.dat, line_connect.c:185-256): t on a decade grid, M = k(zone) t^-0.6 capped at 2000.
    Returns (t, M[mpoints][1][ny][nx]) as written to the file (text round trip) and their log10 taken
    with libm (math.log10), which is what the reference stores in t_fit / M_UV_fit."""
    import math
    t = np.array([float("%.17e" % (10.0 ** e)) for e in np.linspace(-9.0, 2.0, mpoints)])
    kz = 0.3 + 0.5 * (np.sin(3.0 * np.asarray(x1))[None, None, :] ** 2) * (0.5 + 0.5 * np.cos(2.0 * np.asarray(x2))[None, :, None])
    M = np.minimum(kz[None] * t[:, None, None, None] ** -0.6, 2000.0)
    M = np.vectorize(lambda q: float("%.17e" % q))(M)
    lt = np.array([math.log10(v) for v in t])
    lM = np.vectorize(math.log10)(M)
    return t, M, lt, lM


def write_ldw_mfit_file(wd, x1, x2, ng, t, M):
    with open(Path(wd) / "M_UV_data.dat", "w") as f:
        f.write("# %d\n" % len(t))
        f.write("t " + " ".join("%.17e" % v for v in t) + "\n")
        for j in range(ng, len(x2) - ng):
            for i in range(ng, len(x1) - ng):
                vals = " ".join("%.17e" % M[m, 0, j, i] for m in range(len(t)))
                f.write("%d %d %.17e %.17e %s\n" % (i - ng, j - ng, x1[i] * LDW_UNITS["length"], x2[j], vals))
