#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference executables in oracle/_ref
(built by oracle/build_ref.py from /root/reference).  Run in the build container only;
the fixtures are committed so that the GPU box (no /root/reference) can check against them.

Each fixture holds, for one configuration: the run parameters, the list of (nstep, t, dt)
records of restart.out, and selected per-step dumps of d->Vc (interior zones)."""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import build_ref  # noqa: E402
import refrun     # noqa: E402

SOD_BCS = ("outflow", "outflow", "periodic", "periodic", "outflow", "outflow")
SEDOV_BCS = ("reflective", "outflow") * 3
SEDOV_PAR = dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4)

CASES = {
    # name: (ref config, dims, N, recon, rk, solver, bcs, maxsteps, params, gamma, cfl, first_dt, tstop, keep)
    "sod_plm_hllc": ("sod", 1, 400, "LINEAR", "RK2", "hllc", SOD_BCS, 60, {"SCRH": 0}, 1.4, 0.8, 1e-4, 0.2),
    "sod_plm_hll": ("sod", 1, 400, "LINEAR", "RK2", "hll", SOD_BCS, 30, {"SCRH": 0}, 1.4, 0.8, 1e-4, 0.2),
    "sod_plm_tvdlf": ("sod", 1, 400, "LINEAR", "RK2", "tvdlf", SOD_BCS, 30, {"SCRH": 0}, 1.4, 0.8, 1e-4, 0.2),
    "sod_ppm_hllc": ("sod_ppm", 1, 400, "PARABOLIC", "RK3", "hllc", SOD_BCS, 30, {"SCRH": 0}, 1.4, 0.8, 1e-4, 0.2),
    "sedov2d_plm_hllc": ("sedov2d", 2, 32, "LINEAR", "RK2", "hllc", SEDOV_BCS, 16, SEDOV_PAR, 1.4, 0.3, 1e-9, 0.5),
    "sedov3d_plm_hllc": ("sedov3d", 3, 16, "LINEAR", "RK2", "hllc", SEDOV_BCS, 12, SEDOV_PAR, 1.4, 0.3, 1e-9, 0.5),
    "sedov3d_plm_hll": ("sedov3d", 3, 16, "LINEAR", "RK2", "hll", SEDOV_BCS, 8, SEDOV_PAR, 1.4, 0.3, 1e-9, 0.5),
    "sedov3d_ppm_hllc": ("sedov3d_ppm", 3, 16, "PARABOLIC", "RK3", "hllc", SEDOV_BCS, 10, SEDOV_PAR, 1.4, 0.3, 1e-9, 0.5),
    "sedov3d_ppm_tvdlf": ("sedov3d_ppm", 3, 16, "PARABOLIC", "RK3", "tvdlf", SEDOV_BCS, 6, SEDOV_PAR, 1.4, 0.3, 1e-9, 0.5),
}


# cases with tracers / body forces / non-unit domains (dict form)
RT_BCS = ("periodic", "periodic", "reflective", "reflective", "periodic", "periodic")
RT_PAR = dict(ETA=2.0, GRAV=-0.1)
KH_BCS = ("periodic", "periodic", "outflow", "outflow", "periodic", "periodic")
CASES2 = {
    "rt3d_vec_hllc": dict(cfg="rt3d_vec", dims=3, nx=(12, 24, 10), xbeg=(-0.5, -1.0, -0.5), xend=(0.5, 1.0, 0.5),
                          solver="hllc", bcs=RT_BCS, maxsteps=8, params=RT_PAR, gamma=5. / 3., cfl=0.4,
                          first_dt=1e-3, tstop=5.0, ntracer=1, body_force="vector", limiter="DEFAULT"),
    "rt3d_pot_hll": dict(cfg="rt3d_pot", dims=3, nx=(12, 24, 10), xbeg=(-0.5, -1.0, -0.5), xend=(0.5, 1.0, 0.5),
                         solver="hll", bcs=RT_BCS, maxsteps=8, params=RT_PAR, gamma=5. / 3., cfl=0.4,
                         first_dt=1e-3, tstop=5.0, ntracer=1, body_force="potential", limiter="DEFAULT"),
    "rt2d_vec_hllc": dict(cfg="rt2d_vec", dims=2, nx=(16, 48, 1), xbeg=(-0.5, -1.5, -0.5), xend=(0.5, 1.5, 0.5),
                          solver="hllc", bcs=RT_BCS, maxsteps=12, params=RT_PAR, gamma=5. / 3., cfl=0.4,
                          first_dt=1e-3, tstop=5.0, ntracer=1, body_force="vector", limiter="DEFAULT"),
    "rt2d_pot_mc_hllc": dict(cfg="rt2d_pot", dims=2, nx=(16, 48, 1), xbeg=(-0.5, -1.5, -0.5), xend=(0.5, 1.5, 0.5),
                             solver="hllc", bcs=RT_BCS, maxsteps=12, params=RT_PAR, gamma=5. / 3., cfl=0.4,
                             first_dt=1e-3, tstop=5.0, ntracer=1, body_force="potential", limiter="MC_LIM"),
    "rt1d_vec_tvdlf": dict(cfg="rt1d_vec", dims=1, nx=(64, 1, 1), xbeg=(-0.5, -1.0, -0.5), xend=(0.5, 1.0, 0.5),
                           solver="tvdlf", bcs=RT_BCS, maxsteps=10, params=RT_PAR, gamma=5. / 3., cfl=0.4,
                           first_dt=1e-3, tstop=5.0, ntracer=1, body_force="vector", limiter="DEFAULT"),
    "kh3d_hllc": dict(cfg="kh3d", dims=3, nx=(16, 20, 8), xbeg=(0.0, -0.5, 0.0), xend=(1.0, 0.5, 0.5),
                      solver="hllc", bcs=KH_BCS, maxsteps=8, params=dict(A_KH=0.05, DRHO=1.0, MACH=0.8),
                      gamma=1.4, cfl=0.4, first_dt=1e-4, tstop=5.0, ntracer=1, body_force="none",
                      limiter="DEFAULT"),
}


def make_case2(out, name, c):
    build_ref.build(c["cfg"])
    nd, nx = c["dims"], c["nx"]
    nvar = 5 + c["ntracer"]
    with tempfile.TemporaryDirectory() as wd:
        r = refrun.run(c["cfg"], wd, shape=(nx[2], nx[1], nx[0]), nvar=nvar, maxsteps=c["maxsteps"],
                       grid=[(c["xbeg"][d], nx[d], c["xend"][d]) for d in range(3)], cfl=c["cfl"],
                       tstop=c["tstop"], first_dt=c["first_dt"], solver=c["solver"], bcs=c["bcs"],
                       dbl=(-1.0, 1), params=c["params"])
    nd_ = len(r["data"]) - 1
    steps = np.array(r["steps"][:nd_], dtype=np.float64)
    data = np.stack(r["data"][:nd_])
    np.savez_compressed(out / (name + ".npz"), data=data, steps=steps, nx=np.array(nx), dims=nd,
                        recon="LINEAR", rk="RK2", solver=c["solver"], bcs=np.array(c["bcs"]),
                        gamma=c["gamma"], cfl=c["cfl"], cfl_max_var=1.1, first_dt=c["first_dt"],
                        tstop=c["tstop"], ref_config=c["cfg"], xbeg=np.array(c["xbeg"]),
                        xend=np.array(c["xend"]), ntracer=c["ntracer"], body_force=c["body_force"],
                        grav=c["params"].get("GRAV", 0.0), limiter=c["limiter"])
    print(name, data.shape, "%.1f kB" % ((out / (name + ".npz")).stat().st_size / 1e3))


# general-grid cases (spherical geometry, stretched grids, characteristic limiting, MULTID
# flattening, entropy switch): user files oracle/problems/sph
import pluto_grid  # noqa: E402

SPH_PAR = dict(GM=1.0, RBLOB=2.0, TBLOB=1.0, PBLOB=5.0)
HALF_PI = 1.5707963267948966
SPH_GRID2 = [(1.0, 40, 4.0, "r", 1.03), (0.2, 28, HALF_PI, "r", 0.97), (0.0, 1, 1.0)]
SPH_BCS = ("outflow", "outflow", "axisymmetric", "reflective", "periodic", "periodic")
CASES3 = {
    "sph2d_hll": dict(cfg="sph2d", dims=2, grid=SPH_GRID2, solver="hll", bcs=SPH_BCS, maxsteps=10),
    "sph2d_hllc": dict(cfg="sph2d", dims=2, grid=[(1.0, 40, 4.0), (0.2, 28, HALF_PI), (0.0, 1, 1.0)],
                       solver="hllc", bcs=("reflective", "outflow", "reflective", "eqtsymmetric", "periodic", "periodic"),
                       maxsteps=10),
    "sph2d_char_hll": dict(cfg="sph2d_char", dims=2, grid=SPH_GRID2, solver="hll", bcs=SPH_BCS, maxsteps=10,
                           char_limiting=True, limiter="VANLEER_LIM"),
    "sph2d_flat_hllc": dict(cfg="sph2d_flat", dims=2, grid=SPH_GRID2, solver="hllc", bcs=SPH_BCS, maxsteps=10,
                            char_limiting=True, limiter="VANLEER_LIM", shock_flattening=True),
    "sph2d_entr_hll": dict(cfg="sph2d_entr", dims=2, grid=SPH_GRID2, solver="hll", bcs=SPH_BCS, maxsteps=10,
                           char_limiting=True, limiter="VANLEER_LIM", shock_flattening=True, entropy_switch=True),
    "sph2d_sel_hllc": dict(cfg="sph2d_sel", dims=2, grid=SPH_GRID2, solver="hllc", bcs=SPH_BCS, maxsteps=12,
                           entropy_switch="SELECTIVE"),
    # Cartesian, stretched grids: the FAST path with a non-uniform grid->dx (UNIFORM_CARTESIAN_GRID limiters)
    "cart3d_stretched_hllc": dict(cfg="sedov3d", dims=3, geometry="CARTESIAN", ntracer=0, body_force="none",
                                  grid=[(0.0, 20, 1.0, "r", 1.06), (0.0, 16, 1.0, "r", 0.95), (0.0, 12, 1.0)],
                                  solver="hllc", bcs=("reflective", "outflow") * 3, maxsteps=10,
                                  params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4), gamma=1.4, cfl=0.3, first_dt=1e-9, tstop=0.5),
    "sph1d_tvdlf": dict(cfg="sph1d", dims=1, grid=[(1.0, 64, 4.0, "r", 1.02), (1.0, 1, 1.2), (0.0, 1, 1.0)],
                        solver="tvdlf", bcs=("reflective", "outflow") * 3, maxsteps=10),
    "sph3d_hllc": dict(cfg="sph3d", dims=3, grid=[(1.0, 20, 3.0, "r", 1.04), (0.3, 14, HALF_PI), (0.0, 10, 1.0)],
                       solver="hllc", bcs=("outflow", "outflow", "reflective", "reflective", "periodic", "periodic"),
                       maxsteps=6),
}


def make_case3(out, name, c):
    build_ref.build(c["cfg"])
    nd = c["dims"]
    grid = c["grid"]
    nx = [int(grid[d][1]) if d < nd else 1 for d in range(3)]
    ntr = c.get("ntracer", 1)
    iso = c.get("eos", "IDEAL") == "ISOTHERMAL"
    nvar = (4 if iso else 5) + ntr      # ENTR is not written: Boundary() recomputes it (ComputeEntropy)
    with tempfile.TemporaryDirectory() as wd:
        r = refrun.run(c["cfg"], wd, shape=(nx[2], nx[1], nx[0]), nvar=nvar, maxsteps=c["maxsteps"],
                       grid=[pluto_grid.ini_string(g) for g in grid], cfl=c.get("cfl", 0.4), tstop=c.get("tstop", 10.0),
                       first_dt=c.get("first_dt", 1e-5), solver=c["solver"], bcs=c["bcs"], dbl=(-1.0, 1),
                       params=c.get("params", SPH_PAR), timeout=120)
    nd_ = len(r["data"]) - 1
    steps = np.array(r["steps"][:nd_], dtype=np.float64)
    data = np.stack(r["data"][:nd_])
    gridarr = np.array([[g[0], g[1], g[2], 1.0 if (len(g) > 3 and g[3] == "r") else 0.0,
                         g[4] if len(g) > 4 else 1.0] for g in grid], dtype=np.float64)
    np.savez_compressed(out / (name + ".npz"), data=data, steps=steps, nx=np.array(nx), dims=nd,
                        recon=c.get("recon", "LINEAR"), rk=c.get("rk", "RK2"), solver=c["solver"], bcs=np.array(c["bcs"]),
                        gamma=c.get("gamma", 5. / 3.), cfl=c.get("cfl", 0.4), cfl_max_var=1.1, first_dt=c.get("first_dt", 1e-5),
                        tstop=c.get("tstop", 10.0),
                        ref_config=c["cfg"], gridspec=gridarr, geometry=c.get("geometry", "SPHERICAL"), ntracer=ntr,
                        body_force=c.get("body_force", "vector"), gm=c.get("params", SPH_PAR).get("GM", 0.0), limiter=c.get("limiter", "DEFAULT"),
                        char_limiting=int(c.get("char_limiting", False)),
                        shock_flattening=2 if c.get("shock_flattening") == "ONED" else int(bool(c.get("shock_flattening", False))),
                        entropy_switch={False: 0, True: 2, "SELECTIVE": 1, "ALWAYS": 2}[c.get("entropy_switch", False)],
                        entr_codes=1, **(dict(eos="ISOTHERMAL", iso_cs=c["params"]["CS_ISO"]) if iso else {}),
                        **(dict(ring_average=c["ring_average"], ring_rec=c.get("ring_rec", 5)) if c.get("ring_average") else {}))
    print(name, data.shape, "%.1f kB" % ((out / (name + ".npz")).stat().st_size / 1e3))


# cylindrical / polar geometry (user files oracle/problems/cyl).  ORACLE fixtures: the CUDA path does not
# build these geometries yet, so the names carry their own prefixes (tests/common.py: CURV_CASES)
TWO_PI = 6.283185307179586
CYL_PAR = dict(GM=1.0, RBLOB=1.6, ZBLOB=0.6, PBLOB=4.0)
CASES5 = {
    # axis at r = 0: AXISYMMETRIC flips v_r and v_phi, ghost zones have r < 0 (|x1| in dV and the areas)
    "cyl2d_axis_hllc": dict(cfg="cyl2d_nobf", dims=2, geometry="CYLINDRICAL", body_force="none",
                            grid=[(0.0, 40, 2.4), (-0.6, 32, 1.4, "r", 1.02), (0.0, 1, 1.0)], solver="hllc",
                            bcs=("axisymmetric", "outflow", "outflow", "reflective", "periodic", "periodic"),
                            params=CYL_PAR, maxsteps=10),
    "cyl2d_grav_hll": dict(cfg="cyl2d", dims=2, geometry="CYLINDRICAL",
                           grid=[(0.8, 36, 3.0, "r", 1.03), (0.0, 28, 1.5), (0.0, 1, 1.0)], solver="hll",
                           bcs=("outflow", "outflow", "eqtsymmetric", "outflow", "periodic", "periodic"),
                           params=CYL_PAR, maxsteps=10),
    "cyl2d_flat_tvdlf": dict(cfg="cyl2d_flat", dims=2, geometry="CYLINDRICAL", char_limiting=True,
                             shock_flattening=True, limiter="VANLEER_LIM",
                             grid=[(0.8, 36, 3.0, "r", 1.03), (0.0, 28, 1.5), (0.0, 1, 1.0)], solver="tvdlf",
                             bcs=("reflective", "outflow", "eqtsymmetric", "outflow", "periodic", "periodic"),
                             params=CYL_PAR, maxsteps=10),
    "pol2d_hllc": dict(cfg="pol2d", dims=2, geometry="POLAR",
                       grid=[(0.8, 32, 3.0, "r", 1.03), (0.0, 40, TWO_PI), (0.0, 1, 1.0)], solver="hllc",
                       bcs=("reflective", "outflow", "periodic", "periodic", "periodic", "periodic"),
                       params=CYL_PAR, maxsteps=10),
    "pol3d_hll": dict(cfg="pol3d", dims=3, geometry="POLAR",
                      grid=[(0.8, 18, 2.6), (0.0, 20, TWO_PI), (0.0, 10, 1.2, "r", 1.05)], solver="hll",
                      bcs=("outflow", "outflow", "periodic", "periodic", "reflective", "outflow"),
                      params=CYL_PAR, maxsteps=6),
}


# EOS ISOTHERMAL (user files oracle/problems/iso): ORACLE fixtures like the cyl / pol ones
ISO_PAR = dict(CS_ISO=0.7, GM=1.0)
CASES6 = {
    "iso2d_hllc": dict(cfg="iso2d", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                       grid=[(0.0, 40, 1.0), (0.0, 32, 1.0, "r", 1.02), (0.0, 1, 1.0)], solver="hllc",
                       bcs=("outflow", "reflective", "periodic", "periodic", "periodic", "periodic"),
                       params=ISO_PAR, maxsteps=10, first_dt=1e-4),
    "iso2d_hll": dict(cfg="iso2d", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                      grid=[(0.0, 40, 1.0), (0.0, 32, 1.0), (0.0, 1, 1.0)], solver="hll",
                      bcs=("reflective", "outflow", "outflow", "reflective", "periodic", "periodic"),
                      params=ISO_PAR, maxsteps=8, first_dt=1e-4),
    "iso2d_flat_hllc": dict(cfg="iso2d_flat", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                            char_limiting=True, shock_flattening=True, limiter="VANLEER_LIM",
                            grid=[(0.0, 40, 1.0), (0.0, 32, 1.0, "r", 1.02), (0.0, 1, 1.0)], solver="hllc",
                            bcs=("outflow", "reflective", "periodic", "periodic", "periodic", "periodic"),
                            params=ISO_PAR, maxsteps=10, first_dt=1e-4),
    "iso3d_tvdlf": dict(cfg="iso3d", dims=3, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                        grid=[(0.0, 20, 1.0), (0.0, 16, 1.0), (0.0, 12, 1.0)], solver="tvdlf",
                        bcs=("outflow", "outflow", "periodic", "periodic", "reflective", "outflow"),
                        params=ISO_PAR, maxsteps=6, first_dt=1e-4),
    # Roe solver (HD/roe.c), ideal and isothermal; "roe" prefix: ORACLE fixtures
    "roe_cart2d": dict(cfg="sedov2d", dims=2, geometry="CARTESIAN", body_force="none", ntracer=0, gamma=1.4,
                       grid=[(0.0, 32, 1.0), (0.0, 32, 1.0), (0.0, 1, 1.0)], solver="roe",
                       bcs=("reflective", "outflow", "reflective", "outflow", "outflow", "outflow"),
                       params=SEDOV_PAR, maxsteps=14, first_dt=1e-9, cfl=0.3),
    "roe_sph2d_flat": dict(cfg="sph2d_flat", dims=2, grid=SPH_GRID2, solver="roe", bcs=SPH_BCS, maxsteps=10,
                           char_limiting=True, shock_flattening=True, limiter="VANLEER_LIM"),
    "roe_iso2d": dict(cfg="iso2d", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                      grid=[(0.0, 40, 1.0), (0.0, 32, 1.0, "r", 1.02), (0.0, 1, 1.0)], solver="roe",
                      bcs=("outflow", "reflective", "periodic", "periodic", "periodic", "periodic"),
                      params=ISO_PAR, maxsteps=10, first_dt=1e-4),
    # TwoShock_Solver (HD/two_shock.c); "twoshock" prefix: ORACLE fixtures
    "twoshock_sph2d_flat": dict(cfg="sph2d_flat", dims=2, grid=SPH_GRID2, solver="two_shock", bcs=SPH_BCS, maxsteps=10,
                                char_limiting=True, shock_flattening=True, limiter="VANLEER_LIM"),
    "twoshock_sph3d": dict(cfg="sph3d", dims=3, grid=[(1.0, 20, 3.0, "r", 1.04), (0.3, 14, HALF_PI), (0.0, 10, 1.0)],
                           solver="two_shock", bcs=("outflow", "outflow", "reflective", "reflective", "periodic", "periodic"),
                           maxsteps=6),
    # SHOCK_FLATTENING ONED (States/flatten.c, NGHOST 4); "oned" prefix: ORACLE fixtures
    "oned_sph2d_hllc": dict(cfg="sph2d_oned", dims=2, grid=SPH_GRID2, solver="hllc", bcs=SPH_BCS, maxsteps=10,
                            shock_flattening="ONED"),
    "oned_sph2d_char_roe": dict(cfg="sph2d_char_oned", dims=2, grid=SPH_GRID2, solver="roe", bcs=SPH_BCS, maxsteps=10,
                                shock_flattening="ONED", char_limiting=True, limiter="MC_LIM"),
    "oned_iso2d_hll": dict(cfg="iso2d_oned", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                           grid=[(0.0, 40, 1.0), (0.0, 32, 1.0, "r", 1.02), (0.0, 1, 1.0)], solver="hll",
                           bcs=("outflow", "reflective", "periodic", "periodic", "periodic", "periodic"),
                           params=ISO_PAR, maxsteps=10, first_dt=1e-4, shock_flattening="ONED"),
    # RECONSTRUCTION PARABOLIC + RK3 on general grids (ppm_coeffs.c weights); "ppmg" prefix: ORACLE fixtures
    "ppmg_kh3d_stretched": dict(cfg="kh3d_ppm", dims=3, geometry="CARTESIAN", body_force="none", recon="PARABOLIC", rk="RK3",
                                grid=[(0.0, 16, 1.0, "r", 1.04), (-0.5, 20, 0.5, "r", 0.97), (0.0, 8, 0.5)],
                                solver="hllc", bcs=KH_BCS, params=dict(A_KH=0.05, DRHO=1.0, MACH=0.8), gamma=1.4,
                                maxsteps=6, first_dt=1e-4),
    "ppmg_cyl2d_uniform": dict(cfg="cyl2d_ppm", dims=2, geometry="CYLINDRICAL", recon="PARABOLIC", rk="RK3",
                               grid=[(0.8, 36, 3.0), (0.0, 28, 1.5), (0.0, 1, 1.0)], solver="hllc",
                               bcs=("outflow", "outflow", "eqtsymmetric", "outflow", "periodic", "periodic"),
                               params=CYL_PAR, maxsteps=8),
    "ppmg_cyl2d_flat_stretched": dict(cfg="cyl2d_ppm_flat", dims=2, geometry="CYLINDRICAL", recon="PARABOLIC", rk="RK3",
                                      shock_flattening=True,
                                      grid=[(0.8, 36, 3.0, "r", 1.03), (0.0, 28, 1.5, "r", 1.02), (0.0, 1, 1.0)], solver="hll",
                                      bcs=("reflective", "outflow", "eqtsymmetric", "outflow", "periodic", "periodic"),
                                      params=CYL_PAR, maxsteps=8),
    "ppmg_pol2d": dict(cfg="pol2d_ppm", dims=2, geometry="POLAR", recon="PARABOLIC", rk="RK3",
                       grid=[(0.8, 32, 3.0, "r", 1.03), (0.0, 40, TWO_PI), (0.0, 1, 1.0)], solver="roe",
                       bcs=("reflective", "outflow", "periodic", "periodic", "periodic", "periodic"),
                       params=CYL_PAR, maxsteps=8),
    "ppmg_sph2d_uniform": dict(cfg="sph2d_ppm", dims=2, recon="PARABOLIC", rk="RK3",
                               grid=[(1.0, 40, 4.0), (0.2, 28, HALF_PI), (0.0, 1, 1.0)], solver="hllc", bcs=SPH_BCS, maxsteps=8),
    "ppmg_sph2d_stretched": dict(cfg="sph2d_ppm", dims=2, recon="PARABOLIC", rk="RK3", grid=SPH_GRID2, solver="hll",
                                 bcs=SPH_BCS, maxsteps=8),
    "ppmg_sph3d": dict(cfg="sph3d_ppm", dims=3, recon="PARABOLIC", rk="RK3",
                       grid=[(1.0, 20, 3.0, "r", 1.04), (0.3, 14, HALF_PI), (0.0, 10, 1.0)], solver="hllc",
                       bcs=("outflow", "outflow", "reflective", "reflective", "periodic", "periodic"), maxsteps=5),
    # BODY_FORCE POTENTIAL on curvilinear grids; "pot" prefix: ORACLE fixtures
    "pot_sph2d_hllc": dict(cfg="sph2d_pot", dims=2, grid=SPH_GRID2, solver="hllc", bcs=SPH_BCS, maxsteps=8,
                           body_force="potential"),
    "pot_sph3d_both_hll": dict(cfg="sph3d_pot", dims=3, grid=[(1.0, 20, 3.0, "r", 1.04), (0.3, 14, HALF_PI), (0.0, 10, 1.0)],
                               solver="hll", bcs=("outflow", "outflow", "reflective", "reflective", "periodic", "periodic"),
                               maxsteps=5, body_force="both"),
    "pot_pol2d_roe": dict(cfg="pol2d_pot", dims=2, geometry="POLAR", body_force="potential",
                          grid=[(0.8, 32, 3.0, "r", 1.03), (0.0, 40, TWO_PI), (0.0, 1, 1.0)], solver="roe",
                          bcs=("reflective", "outflow", "periodic", "periodic", "periodic", "periodic"),
                          params=CYL_PAR, maxsteps=8),
    "ausm_sph2d": dict(cfg="sph2d", dims=2, grid=SPH_GRID2, solver="ausm+", bcs=SPH_BCS, maxsteps=8),
    "ppmg_sph2d_char_flat": dict(cfg="sph2d_ppm_char", dims=2, recon="PARABOLIC", rk="RK3", grid=SPH_GRID2, solver="hllc",
                                 bcs=SPH_BCS, maxsteps=8, char_limiting=True, shock_flattening=True),
    "ppmg_iso2d_char": dict(cfg="iso2d_ppm_char", dims=2, geometry="CARTESIAN", eos="ISOTHERMAL", body_force="none",
                            recon="PARABOLIC", rk="RK3", char_limiting=True,
                            grid=[(0.0, 40, 1.0), (0.0, 32, 1.0, "r", 1.02), (0.0, 1, 1.0)], solver="hll",
                            bcs=("outflow", "reflective", "periodic", "periodic", "periodic", "periodic"),
                            params=ISO_PAR, maxsteps=8, first_dt=1e-4),
    "iso_sph2d_flat_hll": dict(cfg="iso_sph2d", dims=2, geometry="SPHERICAL", eos="ISOTHERMAL",
                               char_limiting=True, shock_flattening=True, limiter="VANLEER_LIM",
                               grid=SPH_GRID2, solver="hll", bcs=SPH_BCS, params=ISO_PAR, maxsteps=10),
}


# the line-driven disc wind (UNMODIFIED user files of the reference, cv_idl) with synthetic
# sirocco flux tables (tests/common.py: ldw_flux_tables)
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402

LDW_GRID = [(0.87, 48, 8.7, "r", 1.05), (0.0, 36, HALF_PI, "r", 0.95), (0.0, 1, 1.0)]
CASES4 = {
    "ldw_nocool_hll": dict(cfg="ldw_nocool", grid=LDW_GRID, solver="hll", maxsteps=10, cooling=False),
    "ldw_nocool_hllc": dict(cfg="ldw_nocool", grid=LDW_GRID, solver="hllc", maxsteps=8, cooling=False),
    # with the BLONDIN source step (Strang alternation of Src/main.c:479-485); the prefactor tables
    # of a first (non-restart) run are never assigned by the reference (Src/initialize.c:505-520)
    "ldw_cool_hll": dict(cfg="ldw", grid=LDW_GRID, solver="hll", maxsteps=9, cooling=True),
    # force multiplier from the per-zone M(t) fit file instead of k t^alpha (KRAD = ALPHARAD = 999)
    "ldw_nocool_fit_hll": dict(cfg="ldw_nocool", grid=LDW_GRID, solver="hll", maxsteps=8, cooling=False, fit=True),
    # EOS ISOTHERMAL twin (cv_iso user files, unmodified): ORACLE fixture ("iso" prefix, tests/common.py)
    "iso_ldw_hll": dict(cfg="ldw_iso", grid=LDW_GRID, solver="hll", maxsteps=8, cooling=False, iso=True),
}


# RING_AVERAGE (Src/ring_average.c) with POLARAXIS boundaries; "ring" prefix
CASES7 = {
    "ring_pol2d_mp5": dict(cfg="pol2d_ring", dims=2, geometry="POLAR", body_force="none", ring_average=8,
                           grid=[(0.0, 24, 2.4), (0.0, 32, TWO_PI), (0.0, 1, 1.0)], solver="hllc",
                           bcs=("polaraxis", "outflow", "periodic", "periodic", "periodic", "periodic"),
                           params=CYL_PAR, maxsteps=8),
    "ring_pol2d_vl": dict(cfg="pol2d_ring_vl", dims=2, geometry="POLAR", body_force="none", ring_average=8, ring_rec=2,
                          grid=[(0.0, 24, 2.4, "r", 1.03), (0.0, 32, TWO_PI), (0.0, 1, 1.0)], solver="hll",
                          bcs=("polaraxis", "outflow", "periodic", "periodic", "periodic", "periodic"),
                          params=CYL_PAR, maxsteps=8),
    "ring_pol3d_mp5": dict(cfg="pol3d_ring", dims=3, geometry="POLAR", body_force="none", ring_average=4,
                           grid=[(0.0, 14, 2.1), (0.0, 16, TWO_PI), (0.0, 8, 1.2)], solver="hll",
                           bcs=("polaraxis", "outflow", "periodic", "periodic", "reflective", "outflow"),
                           params=CYL_PAR, maxsteps=6),
    "ring_sph3d_mp5": dict(cfg="sph3d_ring", dims=3, ring_average=4,
                           grid=[(1.0, 14, 3.0, "r", 1.04), (0.0, 12, HALF_PI), (0.0, 16, TWO_PI)], solver="hllc",
                           bcs=("outflow", "outflow", "polaraxis", "eqtsymmetric", "periodic", "periodic"), maxsteps=6),
}


def make_case4(out, name, c):
    build_ref.build(c["cfg"])
    grid = c["grid"]
    ng = 3
    nx = [int(grid[0][1]), int(grid[1][1]), 1]
    xl1, xr1, _ = pluto_grid.make_grid(grid[0], ng)
    xl2, xr2, _ = pluto_grid.make_grid(grid[1], ng)
    x1, x2 = 0.5 * (xl1 + xr1), 0.5 * (xl2 + xr2)
    fr, ft, fp = common.ldw_flux_tables(x1, x2)
    with tempfile.TemporaryDirectory() as wd:
        common.write_ldw_flux_files(wd, x1, x2, ng, fr, ft, fp)
        params = common.LDW_PARAMS
        if c.get("fit"):
            t, M, lt, lM = common.ldw_mfit_tables(x1, x2)
            common.write_ldw_mfit_file(wd, x1, x2, ng, t, M)
            params = dict(params, KRAD=999.0, ALPHARAD=999.0)
        iso = bool(c.get("iso"))
        r = refrun.run(c["cfg"], wd, shape=(1, nx[1], nx[0]), nvar=5 if iso else 6, maxsteps=c["maxsteps"],
                       grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-4,
                       solver=c["solver"], bcs=common.LDW_BCS, dbl=(-1.0, 1), params=params, timeout=250)
    nd_ = len(r["data"]) - 1
    steps = np.array(r["steps"][:nd_], dtype=np.float64)
    data = np.stack(r["data"][:nd_])
    gridarr = np.array([[g[0], g[1], g[2], 1.0 if (len(g) > 3 and g[3] == "r") else 0.0,
                         g[4] if len(g) > 4 else 1.0] for g in grid], dtype=np.float64)
    np.savez_compressed(out / (name + ".npz"), data=data, steps=steps, nx=np.array(nx), dims=2,
                        recon="LINEAR", rk="RK2", solver=c["solver"], bcs=np.array(common.LDW_BCS),
                        gamma=5. / 3., cfl=0.4, cfl_max_var=1.1, first_dt=1e-4, tstop=1.0,
                        ref_config=c["cfg"], gridspec=gridarr, geometry="SPHERICAL", ntracer=1,
                        body_force="vector", limiter="VANLEER_LIM", char_limiting=1, shock_flattening=1,
                        entropy_switch=0 if iso else 2, entr_codes=1, cooling=int(c["cooling"]),
                        fit=int(bool(c.get("fit"))),
                        **(dict(eos="ISOTHERMAL",      # init.c:61-63 of cv_iso: g_isoSoundSpeed
                                iso_cs=np.sqrt(8.3144598e7 * common.LDW_PARAMS["T_ISO"] / 0.6) / common.LDW_UNITS["velocity"])
                           if iso else {}))
    print(name, data.shape, "%.1f kB" % ((out / (name + ".npz")).stat().st_size / 1e3))


def make_long_sedov(out):
    """runs/long_sedov3d_64.npz: the Sedov problem (Test_Problems/HD/Sedov conf 05) at 64^3, free running to t = 0.5 with
    the unmodified reference (1273 steps, ~5 core-minutes), reduced to what a GPU-box test compares against."""
    build_ref.build("sedov3d")
    with tempfile.TemporaryDirectory() as wd:
        r = refrun.run("sedov3d", wd, shape=(64, 64, 64), grid=[(0, 64, 1)] * 3, bcs=("reflective", "outflow") * 3,
                       params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4), cfl=0.3, tstop=0.5, first_dt=1e-9,
                       solver="hllc", dbl=(1000.0, -1), timeout=3600)
    a, st = r["data"][-1], r["steps"][-1]
    (out / "runs").mkdir(exist_ok=True)
    np.savez_compressed(out / "runs" / "long_sedov3d_64.npz", nstep=st[0], t=st[1], dt=st[2], sub=a[:, ::4, ::4, ::4].copy(),
                        plane=a[:, 0].copy(), sums=a.reshape(5, -1).sum(axis=1), sumsq=(a.reshape(5, -1) ** 2).sum(axis=1),
                        mass=a[0].sum(), energy=(0.5 * a[0] * (a[1] ** 2 + a[2] ** 2 + a[3] ** 2) + a[4] / 0.4).sum())


def main():
    out = Path(__file__).resolve().parent
    only = set(sys.argv[1:])
    if "long_sedov3d_64" in only:
        make_long_sedov(out)
        return
    for name, c in CASES2.items():
        if not only or name in only:
            make_case2(out, name, c)
    for name, c in CASES3.items():
        if not only or name in only:
            make_case3(out, name, c)
    for name, c in CASES5.items():
        if not only or name in only:
            make_case3(out, name, c)
    for name, c in CASES6.items():
        if not only or name in only:
            make_case3(out, name, c)
    for name, c in CASES7.items():
        if not only or name in only:
            make_case3(out, name, c)
    for name, c in CASES4.items():
        if not only or name in only:
            make_case4(out, name, c)
    for name, (cfg, nd, N, recon, rk, solver, bcs, maxsteps, params, gamma, cfl, first_dt, tstop) in CASES.items():
        if only and name not in only:
            continue
        build_ref.build(cfg)
        nx = [N if d < nd else 1 for d in range(3)]
        with tempfile.TemporaryDirectory() as wd:
            r = refrun.run(cfg, wd, shape=(nx[2], nx[1], nx[0]), maxsteps=maxsteps,
                           grid=[(0, nx[0], 1), (0, nx[1], 1), (0, nx[2], 1)], cfl=cfl, tstop=tstop,
                           first_dt=first_dt, solver=solver, bcs=bcs, dbl=(-1.0, 1), params=params)
        nd_ = len(r["data"]) - 1          # last file = state after the final (non-dumped) steps
        steps = np.array(r["steps"][:nd_], dtype=np.float64)   # (nstep, t, dt-of-next-step)
        data = np.stack(r["data"][:nd_])
        np.savez_compressed(out / (name + ".npz"), data=data, steps=steps, nx=np.array(nx),
                            dims=nd, recon=recon, rk=rk, solver=solver, bcs=np.array(bcs),
                            gamma=gamma, cfl=cfl, cfl_max_var=1.1, first_dt=first_dt, tstop=tstop,
                            ref_config=cfg)
        print(name, data.shape, "%.1f kB" % ((out / (name + ".npz")).stat().st_size / 1e3))


if __name__ == "__main__":
    main()
