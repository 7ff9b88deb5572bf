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
TUNNEL = dict(bcs=("userdef", "outflow", "reflective", "reflective", "outflow", "outflow"), params=dict(MACH=3.0),
              cfl=0.4, tstop=4.0, first_dt=1e-4)
KH = dict(bcs=("periodic", "periodic", "reflective", "reflective", "periodic", "periodic"),
          params=dict(A_KH=0.05, DRHO=1.0, MACH=0.8), cfl=0.4, tstop=5.0, first_dt=1e-4)
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
    # INTERNAL_BOUNDARY YES: FLAG_INTERNAL_BOUNDARY zones set by UserDefBoundary(side 0) -> InternalBoundaryReset()
    # (Src/int_bound_reset.c:17): the mask travels to the device after every host Boundary() call
    "tunnel2d": dict(shape=(1, 32, 96), grid=[(0.0, 96, 3.0), (0.0, 32, 1.0), (0.0, 1, 1.0)], maxsteps=40, **TUNNEL),
    "tunnel3d_ppm": dict(shape=(6, 20, 60), grid=[(0.0, 60, 3.0), (0.0, 20, 1.0), (0.0, 6, 0.3)], maxsteps=20, **TUNNEL),
    "kh3d": dict(shape=(8, 20, 16), nvar=6, grid=[(0.0, 16, 1.0), (-0.5, 20, 0.5), (0.0, 8, 0.5)], maxsteps=10, **KH),
}


@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_line_driven_wind_matches_reference_executable(cuda_lib, tmp_path, resident):
    """C4 end to end: the reference's own driver + cv_idl user files + BLONDIN cooling, with
    AdvanceStep replaced by libplutob200, on synthetic sirocco flux files read by the reference's
    own reader.  Host-buffer mode keeps SplitSource() on the reference's CPU code; resident mode
    runs BlondinCooling on the device too (--wrap=SplitSource)."""
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
    # both modes within the north-star tolerances: resident mode runs BlondinCooling on the device with the
    # reference's arithmetic (csrc/pb200_cool.cu + glibc_math.cuh), host-buffer mode keeps the reference's SplitSource()
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN


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


def test_internal_boundary_zones_are_frozen_in_the_dropin(cuda_lib, tmp_path):
    """The tunnel problem flags a strip WITHOUT resetting it: with rhs = 0 its state must not move at all
    (both executables), while its neighbours do."""
    cfg = "tunnel2d"
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    kw = dict(CASES[cfg], maxsteps=25)
    got = refrun.run(cfg, tmp_path / "b200", solver="hllc", dbl=(-1.0, 1), exe=exe, **kw)
    x = (np.arange(96) + 0.5) * 3.0 / 96
    y = (np.arange(32) + 0.5) / 32
    strip = np.ix_([0], np.where((y > 0.2) & (y <= 0.3))[0], np.where((x >= 1.8) & (x <= 2.1))[0])
    a, b = got["data"][0], got["data"][-1]
    for nv in range(5):
        assert np.abs(b[nv][strip] - a[nv][strip]).max() <= 1e-13 * max(np.abs(a[nv]).max(), 1.0)   # cons(prim(U)) round-off only
    assert np.abs(b[0] - a[0]).max() > 1e-3                        # the flow around it evolves


@pytest.mark.parametrize("cfg,ranks,devices", [
    ("sedov3d", 2, "0,0"), ("sedov3d", 3, "0,0,0"), ("sedov3d_ppm", 2, "0,0"), ("sedov3d_ppm", 3, "0,0,0"),
    ("rt3d_vec", 2, "0,0"), ("kh3d", 2, "0,0"),
    ("sedov3d", 2, "0,1"), ("sedov3d_ppm", 2, "0,1"), ("rt3d_vec", 2, "0,1"), ("kh3d", 2, "0,1")])
@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_multi_gpu_matches_reference_executable(cuda_lib, tmp_path, cfg, resident, ranks, devices):
    """PB200_NGPUS: the unmodified serial C driver on N slabs (one host thread, pb200_multi_*: NCCL exchange or,
    for ranks sharing a device, device copies) against the stock reference executable.  Replaces the MPI build
    (Src/Parallel, boundary.c:139-158, main.c:288,547)."""
    import torch
    if len(set(devices.split(","))) > torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % len(set(devices.split(","))))
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    kw = dict(CASES[cfg])
    ref = refrun.run(cfg, tmp_path / "ref", solver="hllc", dbl=(-1.0, 1), **kw)
    got = refrun.run(cfg, tmp_path / "b200", solver="hllc", dbl=(-1.0, 1), exe=exe,
                     env={"PB200_RESIDENT": resident, "PB200_NGPUS": str(ranks), "PB200_DEVICES": devices}, **kw)
    assert "%d GPUs" % ranks in got["log"]
    assert len(got["data"]) == len(ref["data"]) >= 8
    for (n1, t1, d1), (n2, t2, d2) in zip(ref["steps"], got["steps"]):
        assert n1 == n2 and abs(t1 - t2) <= 1e-11 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-10 * d1
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN


def test_dropin_line_driven_wind_with_host_boundaries(cuda_lib, tmp_path):
    """PB200_LDW_CVIDL_BC=0: the device copies of cv_idl's UserDefBoundary() are NOT used; every stage the
    reference's own Boundary() -> UserDefBoundary() runs on the host copy (what a modified init.c gets)."""
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
    for tag, ex, env in (("ref", None, {}), ("b200", exe, {"PB200_LDW_CVIDL_BC": "0"}), ("probe", exe, {})):
        wd = tmp_path / tag
        wd.mkdir()
        write_ldw_flux_files(wd, x1, x2, 3, fr, ft, fp)
        out[tag] = refrun.run(cfg, wd, shape=(1, 36, 48), nvar=6, maxsteps=8, timeout=250, exe=ex, env=env,
                              grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-4,
                              solver="hll", bcs=LDW_BCS, dbl=(-1.0, 1), params=LDW_PARAMS)
    ref, got = out["ref"], out["b200"]
    assert "boundaries by the host's Boundary()" in got["log"]
    assert "verified against the device version" in out["probe"]["log"]      # default: probe, then device hooks
    assert len(got["data"]) == len(ref["data"]) >= 8
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN
    assert rel_l1(out["probe"]["data"][-1], ref["data"][-1]) <= TOL_RUN


SEDOV64 = dict(shape=(64, 64, 64), grid=[(0, 64, 1)] * 3, bcs=("reflective", "outflow") * 3,
               params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4), cfl=0.3, tstop=0.5, first_dt=1e-9)


def test_dropin_sedov_64_to_tstop_vs_reference_fixture(cuda_lib, tmp_path):
    """SURVEY 8d / north star: the Sedov test problem at 64^3 to t = 0.5 (1273 steps, the blast fills the box),
    free running through the drop-in executable.  The unmodified reference needs ~5 core-minutes for this run, so
    its end state was reduced to a fixture in the container (tests/golden/runs/long_sedov3d_64.npz: every 4th zone
    per direction, the k = 0 plane, per-variable sums; written from oracle/_ref/sedov3d by the recipe in
    tests/golden/make_golden.py::make_long_sedov) instead of re-running it on the GPU box."""
    cfg = "sedov3d"
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists():
        pytest.skip("drop-in executable was not built in the container (needs /root/reference)")
    g = np.load(ROOT / "tests" / "golden" / "runs" / "long_sedov3d_64.npz")
    got = refrun.run(cfg, tmp_path / "b200", solver="hllc", dbl=(1000.0, -1), exe=exe, env={"PB200_RESIDENT": "1"},
                     timeout=900, **SEDOV64)
    n, t, dt = got["steps"][-1]
    assert n == int(g["nstep"]) and abs(t - float(g["t"])) <= 1e-12 * t
    a = got["data"][-1]
    assert rel_l1(a[:, ::4, ::4, ::4], g["sub"]) <= TOL_RUN
    assert rel_l1(a[:, 0], g["plane"]) <= TOL_RUN
    assert np.all(np.abs(a.reshape(5, -1).sum(axis=1) - g["sums"]) <= 1e-10 * np.abs(g["sums"]).max())
    mass = a[0].sum()
    energy = (0.5 * a[0] * (a[1] ** 2 + a[2] ** 2 + a[3] ** 2) + a[4] / 0.4).sum()
    assert abs(mass - float(g["mass"])) <= 1e-12 * mass and abs(energy - float(g["energy"])) <= 1e-11 * energy


LONG = [
    # 3-D RT and KH, free running against the reference executable on the box.  (cfg, ini, tstop, strict):
    # strict = the north-star bound 1e-9 rel. L1; otherwise the run is long enough for the flow to amplify
    # round-off differences (measured: x100 per unit time for this RT set-up) and the bound is 10x the distance
    # between the reference and ITSELF compiled with FMA contraction (oracle/build_ref.py: *_fma), never < 1e-9.
    ("rt3d_vec", dict(shape=(16, 64, 24), nvar=6, grid=[(-0.5, 24, 0.5), (-1.0, 64, 1.0), (-0.5, 16, 0.5)], **RT), 1.0, True),
    ("rt3d_vec", dict(shape=(16, 64, 24), nvar=6, grid=[(-0.5, 24, 0.5), (-1.0, 64, 1.0), (-0.5, 16, 0.5)], **RT), 4.0, False),
    ("kh3d", dict(shape=(12, 48, 48), nvar=6, grid=[(0.0, 48, 1.0), (-0.5, 48, 0.5), (0.0, 12, 0.25)], **KH), 1.0, True),
    ("kh3d", dict(shape=(12, 48, 48), nvar=6, grid=[(0.0, 48, 1.0), (-0.5, 48, 0.5), (0.0, 12, 0.25)], **KH), 3.0, False),
]


@pytest.mark.parametrize("cfg,kw,tstop,strict", LONG)
def test_dropin_long_runs_to_tstop(cuda_lib, tmp_path, cfg, kw, tstop, strict):
    """Full-run tolerance on the 3-D instability problems: the drop-in (resident mode) and the stock executable
    run freely to tstop; same step count, same time, conserved totals equal to round-off, and the end states
    within 1e-9 relative L1 - or, for runs long enough to amplify round-off, within 10x of what separates two
    roundings of the reference itself."""
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    kw = dict(kw, tstop=tstop)
    ref = refrun.run(cfg, tmp_path / "ref", solver="hllc", dbl=(1000.0, -1), timeout=900, **kw)
    got = refrun.run(cfg, tmp_path / "b200", solver="hllc", dbl=(1000.0, -1), exe=exe, env={"PB200_RESIDENT": "1"},
                     timeout=900, **kw)
    assert len(ref["data"]) == len(got["data"]) == 2
    assert ref["steps"][-1][0] == got["steps"][-1][0] >= (80 if strict else 250)
    assert abs(ref["steps"][-1][1] - got["steps"][-1][1]) <= 1e-12 * ref["steps"][-1][1]
    bound = TOL_RUN
    if not strict:
        if not refrun.have_ref(cfg + "_fma"):
            pytest.skip("the FMA build of the reference is missing")
        fma = refrun.run(cfg + "_fma", tmp_path / "fma", solver="hllc", dbl=(1000.0, -1), timeout=900, **kw)
        assert fma["steps"][-1][0] == ref["steps"][-1][0]
        bound = max(TOL_RUN, 10.0 * rel_l1(fma["data"][-1], ref["data"][-1]))
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= bound
    a, b = ref["data"][-1], got["data"][-1]
    g1 = 1.0 / (5.0 / 3.0 - 1.0) if cfg.startswith("rt") else 1.0 / 0.4
    for tot in ("mass",) + (() if cfg.startswith("rt") else ("energy",)):   # gravity does work: no energy invariant for RT
        qa = a[0].sum() if tot == "mass" else (0.5 * a[0] * (a[1] ** 2 + a[2] ** 2 + a[3] ** 2) + g1 * a[4]).sum()
        qb = b[0].sum() if tot == "mass" else (0.5 * b[0] * (b[1] ** 2 + b[2] ** 2 + b[3] ** 2) + g1 * b[4]).sum()
        assert abs(qa - qb) <= 1e-11 * abs(qa)


def test_dropin_isothermal_line_driven_wind_matches_reference_executable(cuda_lib, tmp_path):
    """The fork's second wind problem, Test_Problems/LineDrivenWind/cv_iso (EOS ISOTHERMAL, COOLING NO, unmodified user
    files): the reference's driver with AdvanceStep on the GPU against the stock executable."""
    import pluto_grid
    from common import LDW_BCS, LDW_PARAMS, ldw_flux_tables, write_ldw_flux_files
    cfg = "ldw_iso"
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
        out[tag] = refrun.run(cfg, wd, shape=(1, 36, 48), nvar=5, maxsteps=10, timeout=250, exe=ex,
                              grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-4,
                              solver="hll", bcs=LDW_BCS, dbl=(-1.0, 1), params=LDW_PARAMS)
    ref, got = out["ref"], out["b200"]
    assert "runs on the GPU" in got["log"]
    assert len(got["data"]) == len(ref["data"]) >= 8
    for (n1, t1, d1), (n2, t2, d2) in zip(ref["steps"], got["steps"]):
        assert n1 == n2 and abs(t1 - t2) <= 1e-11 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-10 * d1
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN


HALF_PI, TWO_PI = 1.5707963267948966, 6.283185307179586
SPH_PAR = dict(GM=1.0, RBLOB=2.0, TBLOB=1.0, PBLOB=5.0)
CYL_PAR = dict(GM=1.0, RBLOB=1.6, ZBLOB=0.6, PBLOB=4.0)
GENERAL_OPTION_CASES = {
    # RECONSTRUCTION PARABOLIC + RK3 on a stretched Cartesian grid: the shim hands grid->uniform over and the
    # context leaves the marching kernels for the general path (PPM_FindWeights weights); Roe solver
    "kh3d_ppm": dict(nvar=6, grid=[(0.0, 16, 1.0, "r", 1.04), (-0.5, 20, 0.5, "r", 0.97), (0.0, 8, 0.5)], solver="roe",
                     bcs=("periodic", "periodic", "outflow", "outflow", "periodic", "periodic"),
                     params=dict(A_KH=0.05, DRHO=1.0, MACH=0.8), first_dt=1e-4, maxsteps=8),
    # POLAR geometry + PPM: the reference's own Grid arrays through pb200_set_geometry, iMPHI = VX2
    "pol2d_ppm": dict(nvar=6, grid=[(0.8, 32, 3.0, "r", 1.03), (0.0, 40, TWO_PI), (0.0, 1, 1.0)], solver="hllc",
                      bcs=("reflective", "outflow", "periodic", "periodic", "periodic", "periodic"),
                      params=CYL_PAR, first_dt=1e-5, maxsteps=10),
    # CYLINDRICAL + PPM + MULTID flattening with the AUSM+ solver
    "cyl2d_ppm_flat": dict(nvar=6, grid=[(0.8, 36, 3.0, "r", 1.03), (0.0, 28, 1.5, "r", 1.02), (0.0, 1, 1.0)], solver="ausm+",
                           bcs=("reflective", "outflow", "eqtsymmetric", "outflow", "periodic", "periodic"),
                           params=CYL_PAR, first_dt=1e-5, maxsteps=10),
    # SPHERICAL + CHAR_LIMITING + SHOCK_FLATTENING ONED (4 ghost zones) with the two-shock solver
    "sph2d_char_oned": dict(nvar=6, grid=[(1.0, 40, 4.0, "r", 1.03), (0.2, 28, HALF_PI, "r", 0.97), (0.0, 1, 1.0)],
                            solver="two_shock", bcs=("outflow", "outflow", "axisymmetric", "reflective", "periodic", "periodic"),
                            params=SPH_PAR, first_dt=1e-5, maxsteps=10),
    # RING_AVERAGE 8 (MP5 on the reduced grid) in POLAR geometry from r = 0 with the polaraxis boundary
    "pol2d_ring": dict(nvar=6, grid=[(0.0, 24, 2.4), (0.0, 32, TWO_PI), (0.0, 1, 1.0)], solver="hllc",
                       bcs=("polaraxis", "outflow", "periodic", "periodic", "periodic", "periodic"),
                       params=CYL_PAR, first_dt=1e-5, maxsteps=10),
    # RING_AVERAGE 4 in SPHERICAL geometry from theta = 0 (non-axisymmetric state)
    "sph3d_ring": dict(nvar=6, grid=[(1.0, 14, 3.0, "r", 1.04), (0.0, 12, HALF_PI), (0.0, 16, TWO_PI)], solver="hll",
                       bcs=("outflow", "outflow", "polaraxis", "eqtsymmetric", "periodic", "periodic"),
                       params=SPH_PAR, first_dt=1e-5, maxsteps=8),
}


@pytest.mark.parametrize("cfg", list(GENERAL_OPTION_CASES))
@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_general_path_options_match_reference_executable(cuda_lib, tmp_path, cfg, resident):
    """The options the general path gained last (CYLINDRICAL / POLAR, PPM on stretched and curvilinear grids, ONED
    flattening, Roe / two-shock / AUSM+) through the reference's own driver: pluto.ini parser, SetGrid / SetGeometry,
    the user's Init / BodyForceVector, Boundary and NextTimeStep are the reference's objects; AdvanceStep is the shim."""
    import pluto_grid
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    kw = dict(GENERAL_OPTION_CASES[cfg])
    grid = kw.pop("grid")
    nx = [int(g[1]) for g in grid]
    kw.update(shape=(nx[2], nx[1], nx[0]), grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=10.0, dbl=(-1.0, 1),
              timeout=200)
    ref = refrun.run(cfg, tmp_path / "ref", **kw)
    got = refrun.run(cfg, tmp_path / "b200", exe=exe, env={"PB200_RESIDENT": resident}, **kw)
    assert "runs on the GPU" in got["log"]
    assert len(got["data"]) == len(ref["data"]) >= 7
    for (n1, t1, d1), (n2, t2, d2) in zip(ref["steps"], got["steps"]):
        assert n1 == n2 and abs(t1 - t2) <= 1e-11 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-10 * d1
    assert np.array_equal(ref["data"][0], got["data"][0])
    assert rel_err(got["data"][1], ref["data"][1]) <= TOL_STEP
    assert rel_l1(got["data"][-1], ref["data"][-1]) <= TOL_RUN


def _coupling_cycles(cfg, wd, exe, env, ncycles=3, steps_per_cycle=5):
    """What Test_Problems/LineDrivenWind/cv_idl/pluto_sirocco_dir_iso.py:100-140 does around the hydro code, with
    synthetic sirocco output: cycle 0 runs ./pluto with the k-alpha force multiplier and the initial flux files; every
    later cycle gets new directional_flux_*.dat, M_UV_data.dat (KRAD = ALPHARAD = 999), py_heatcool.dat and
    prefactors.dat and runs ./pluto -restart from the last dbl file (Src/main.c:142-201 re-reads all tables)."""
    import pluto_grid
    from common import (LDW_BCS, LDW_PARAMS, LDW_UNITS, ldw_flux_tables, ldw_mfit_tables, write_ldw_flux_files,
                        write_ldw_mfit_file)
    from test_sirocco_tables import _write_heatcool
    grid = [(0.87, 48, 8.7, "r", 1.05), (0.0, 36, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    ng = 3
    xl1, xr1, _ = pluto_grid.make_grid(grid[0], ng)
    xl2, xr2, _ = pluto_grid.make_grid(grid[1], ng)
    x1, x2 = 0.5 * (xl1 + xr1), 0.5 * (xl2 + xr2)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    wd.mkdir()
    out = []
    for cyc in range(ncycles):
        scale = 1.0 + 0.15 * cyc                      # the radiation field changes from cycle to cycle
        write_ldw_flux_files(wd, x1, x2, ng, fr * scale, ft * scale, fp)
        params = dict(LDW_PARAMS)
        if cyc > 0:
            rng = np.random.default_rng(100 + cyc)
            t, M, lt, lM = ldw_mfit_tables(x1, x2)
            write_ldw_mfit_file(wd, x1, x2, ng, t, M * scale)
            _write_heatcool(wd / "py_heatcool.dat", x1, x2, ng, rng)
            rows = []
            for j in range(ng, len(x2) - ng):
                for i in range(ng, len(x1) - ng):
                    rows.append("%d %.17e %d %.17e %s" % (i - ng, x1[i] * LDW_UNITS["length"], j - ng, x2[j],
                                                          " ".join("%.17e" % q for q in rng.uniform(0.5, 2.0, 7))))
            (wd / "prefactors.dat").write_text("# header\n" + "\n".join(rows) + "\n")
            params.update(KRAD=999.0, ALPHARAD=999.0)
        r = refrun.run(cfg, wd, shape=(1, 36, 48), nvar=6, maxsteps=steps_per_cycle * (cyc + 1), timeout=250, exe=exe,
                       env=env, keep=cyc > 0, extra_args=("-restart", str(len(out[-1]["data"]) - 1)) if cyc > 0 else (),
                       grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-4,
                       solver="hll", bcs=LDW_BCS, dbl=(-1.0, 1), params=params)
        out.append(r)
    return out


@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_coupling_cycles_with_restart(cuda_lib, tmp_path, resident):
    """three cycles of the pluto <-> sirocco loop on the drop-in executable - restart from its own
    data.NNNN.dbl / restart.out, new flux, force-multiplier, heating / cooling and prefactor files every cycle (read by
    the library's table readers, PB200_FAST_TABLES=1) - against the stock executable put through the same cycles."""
    cfg = "ldw"
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    if not exe.exists() or not refrun.have_ref(cfg):
        pytest.skip("drop-in / reference executables were not built in the container (needs /root/reference)")
    ref = _coupling_cycles(cfg, tmp_path / "ref", None, {})
    got = _coupling_cycles(cfg, tmp_path / "b200", exe, {"PB200_RESIDENT": resident, "PB200_FAST_TABLES": "1"})
    for cyc, (r, g) in enumerate(zip(ref, got)):
        assert "runs on the GPU" in g["log"]
        if cyc > 0:
            assert "libplutob200 table readers" in g["log"] and "Read in 1728 py_heatcool entries" in r["log"]
        assert len(g["data"]) == len(r["data"]) and len(r["steps"]) == len(g["steps"])
        for (n1, t1, d1), (n2, t2, d2) in zip(r["steps"], g["steps"]):
            assert n1 == n2 and abs(t1 - t2) <= 1e-10 * max(t1, 1e-30) and abs(d1 - d2) <= 1e-9 * d1
        assert rel_l1(g["data"][-1], r["data"][-1]) <= TOL_RUN, (cyc, rel_l1(g["data"][-1], r["data"][-1]))
    assert ref[-1]["steps"][-1][0] >= 15 and len(ref[-1]["data"]) > len(ref[0]["data"])
