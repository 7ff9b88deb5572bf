"""GPU: the real drop-in.  The reference's own executable (main loop, pluto.ini parser, grid,
init.c, output - all unmodified objects) linked with pluto_shim.c + libplutob200.so in place of
rk_step.o (integration/build_dropin.py) must write the same data.NNNN.dbl / restart.out as the
stock reference executable on the same pluto.ini."""
from pathlib import Path

import numpy as np
import pytest

import refrun
from common import TOL_RUN, TOL_STEP, rel_err, rel_l1

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

SOD_BCS = ("outflow", "outflow", "periodic", "periodic", "outflow", "outflow")
SEDOV = dict(bcs=("reflective", "outflow") * 3, params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4), cfl=0.3,
             tstop=0.5, first_dt=1e-9)
RT = dict(bcs=("periodic", "periodic", "reflective", "reflective", "periodic", "periodic"),
          params=dict(ETA=2.0, GRAV=-0.1), cfl=0.4, tstop=5.0, first_dt=1e-3)
CASES = {
    "sod": dict(shape=(1, 1, 400), grid=[(0, 400, 1), (0, 1, 1), (0, 1, 1)], cfl=0.8, tstop=0.2, first_dt=1e-4,
                bcs=SOD_BCS, params={"SCRH": 0}, maxsteps=80),
    "sedov3d": dict(shape=(24, 24, 24), grid=[(0, 24, 1)] * 3, maxsteps=15, **SEDOV),
    "sedov3d_ppm": dict(shape=(24, 24, 24), grid=[(0, 24, 1)] * 3, maxsteps=10, **SEDOV),
    # tracer + BODY_FORCE: the shim evaluates the user's BodyForceVector / BodyForcePotential
    "rt3d_vec": dict(shape=(10, 24, 12), nvar=6, grid=[(-0.5, 12, 0.5), (-1.0, 24, 1.0), (-0.5, 10, 0.5)],
                     maxsteps=10, **RT),
    # arbitrary, time-dependent UserDefBoundary(): the shim falls back to the reference's Boundary() per stage
    "jet2d": dict(shape=(1, 64, 48), nvar=6, grid=[(-3.0, 48, 3.0), (0.0, 64, 8.0), (0.0, 1, 1.0)], maxsteps=30,
                  bcs=("outflow", "outflow", "userdef", "outflow", "outflow", "outflow"),
                  params=dict(ETA=10.0, MACH=5.0), cfl=0.4, tstop=5.0, first_dt=1e-4),
    "jet2d_ppm": dict(shape=(1, 64, 48), nvar=6, grid=[(-3.0, 48, 3.0), (0.0, 64, 8.0), (0.0, 1, 1.0)], maxsteps=20,
                      bcs=("outflow", "outflow", "userdef", "outflow", "outflow", "outflow"),
                      params=dict(ETA=10.0, MACH=5.0), cfl=0.4, tstop=5.0, first_dt=1e-4),
    "rt2d_pot": dict(shape=(1, 48, 16), nvar=6, grid=[(-0.5, 16, 0.5), (-1.5, 48, 1.5), (-0.5, 1, 0.5)],
                     maxsteps=12, **RT),
}


@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_line_driven_wind_matches_reference_executable(cuda_lib, tmp_path, resident):
    """C4 end to end: the reference's own driver + cv_idl user files + BLONDIN cooling, with
    AdvanceStep replaced by libplutob200, on synthetic sirocco flux files read by the reference's
    own reader.  Host-buffer mode keeps SplitSource() on the reference's CPU code; resident mode
    runs BlondinCooling on the device too (--wrap=SplitSource), where a zone sitting on the Brent
    solver's 1 K stopping threshold may end one iteration apart (see test_gpu_gen.py)."""
    import pluto_grid
    from common import LDW_BCS, LDW_PARAMS, ldw_flux_tables, write_ldw_flux_files
    cfg = "ldw"
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    grid = [(0.87, 48, 8.7, "r", 1.05), (0.0, 36, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    xl1, xr1, _ = pluto_grid.make_grid(grid[0], 3)
    xl2, xr2, _ = pluto_grid.make_grid(grid[1], 3)
    x1, x2 = 0.5 * (xl1 + xr1), 0.5 * (xl2 + xr2)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    out = {}
    for tag, ex in (("ref", None), ("b200", exe)):
        wd = tmp_path / tag
        wd.mkdir()
        write_ldw_flux_files(wd, x1, x2, 3, fr, ft, fp)
        out[tag] = refrun.run(cfg, wd, shape=(1, 36, 48), nvar=6, maxsteps=12, timeout=250, exe=ex,
                              env={"PB200_RESIDENT": resident},
                              grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-4,
                              solver="hll", bcs=LDW_BCS, dbl=(-1.0, 1), params=LDW_PARAMS)
    ref, got = out["ref"], out["b200"]
    assert "runs on the GPU" in got["log"]
    assert len(got["data"]) == len(ref["data"]) >= 10
    for (n1, t1, d1), (n2, t2, d2) in zip(ref["steps"], got["steps"]):
        assert n1 == n2 and abs(t1 - t2) <= 1e-11 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-10 * d1
    assert np.array_equal(ref["data"][0], got["data"][0])
    if resident == "0":
        assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
        assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN
    else:
        assert rel_err(got["data"][1], ref["data"][1]) <= 2e-4      # pressure: Brent tolerance 1 K
        assert rel_l1(got["data"][-1], ref["data"][-1]) <= 1e-6


@pytest.mark.parametrize("cfg", list(CASES))
@pytest.mark.parametrize("solver", ["hllc", "hll"])
@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_executable_matches_reference_executable(cuda_lib, tmp_path, cfg, solver, resident):
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    kw = dict(CASES[cfg])
    ref = refrun.run(cfg, tmp_path / "ref", solver=solver, dbl=(-1.0, 1), **kw)
    got = refrun.run(cfg, tmp_path / "b200", solver=solver, dbl=(-1.0, 1), exe=exe,
                     env={"PB200_RESIDENT": resident}, **kw)
    assert "runs on the GPU" in got["log"]
    assert len(got["data"]) == len(ref["data"]) and len(ref["data"]) >= 8
    # identical step numbering, time and dt sequence (NextTimeStep stays the reference's code)
    for (n1, t1, d1), (n2, t2, d2) in zip(ref["steps"], got["steps"]):
        assert n1 == n2
        assert abs(t1 - t2) <= 1e-11 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-10 * d1
    # the very first step starts from bit-identical data: per-step tolerance
    assert np.array_equal(ref["data"][0], got["data"][0])
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    # end of the run: 1e-9 relative L1
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN


@pytest.mark.parametrize("cfg,kw", [
    ("sod", dict(shape=(1, 1, 400), grid=[(0, 400, 1), (0, 1, 1), (0, 1, 1)], cfl=0.8, tstop=0.2, first_dt=1e-4,
                 bcs=SOD_BCS, params={"SCRH": 0})),
    ("sedov3d", dict(shape=(24, 24, 24), grid=[(0, 24, 1)] * 3, bcs=("reflective", "outflow") * 3,
                     params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4), cfl=0.3, tstop=0.05, first_dt=1e-9)),
])
def test_dropin_full_test_problem_run(cuda_lib, tmp_path, cfg, kw):
    """The whole test problem to its tstop (C1: Sod to t = 0.2; Sedov 24^3 to t = 0.05), free running:
    same number of steps and <= 1e-9 relative L1 at the end (north-star tolerance for a full run)."""
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    ref = refrun.run(cfg, tmp_path / "ref", solver="hllc", dbl=(10.0, -1), **kw)       # dumps: t = 0 and the end
    got = refrun.run(cfg, tmp_path / "b200", solver="hllc", dbl=(10.0, -1), exe=exe, env={"PB200_RESIDENT": "1"}, **kw)
    assert len(ref["data"]) == len(got["data"]) == 2
    assert ref["steps"][-1][0] == got["steps"][-1][0] > 100                 # same step count, a real run
    assert abs(ref["steps"][-1][1] - got["steps"][-1][1]) <= 1e-12 * ref["steps"][-1][1]
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN
