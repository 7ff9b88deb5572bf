"""GPU: the general-grid path of the library (spherical geometry, stretched grids, characteristic
limiting, MULTID flattening, entropy switch, tracer, BODY_FORCE VECTOR) through the C ABI against
the golden dumps of the compiled reference and against the general oracle on seeded states."""
import numpy as np
import pytest

from common import (GEN_CASES, TOL_STEP, gen_kwargs_from_golden, hydro_kwargs_from_gen, ldw_setup, load_golden,
                    rel_err, set_point_mass_gravity)
from gen_oracle import GenOracle

pytestmark = pytest.mark.gpu
SPH_CASES = [c for c in GEN_CASES if c.startswith(("sph", "cart"))]      # cart*: stretched Cartesian grids on the FAST path


@pytest.fixture(scope="module")
def Hydro(cuda_lib):
    from pluto_sirocco_b200 import Hydro as H
    return H


@pytest.mark.parametrize("name", SPH_CASES)
def test_gen_per_step_vs_reference_dumps(Hydro, name):
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    set_point_mass_gravity(h, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]
    for n in range(len(data) - 1):
        v = data[n]
        if h.nvar > nfile:     # ENTR is not dumped: Boundary() recomputes it
            v = np.concatenate([v, np.ones((h.nvar - nfile,) + v.shape[1:])])
        h.set_interior(v)
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior()[:nfile], data[n + 1])
        assert e <= TOL_STEP, (name, n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], (name, n)
    h.close()


@pytest.mark.parametrize("name", SPH_CASES)
def test_gen_free_run_vs_reference_dumps(Hydro, name):
    from common import TOL_RUN, rel_l1
    from pluto_sirocco_b200 import Runtime, Simulation
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    set_point_mass_gravity(h, float(g["gm"]))
    nfile = g["data"].shape[1]
    v = g["data"][0]
    if h.nvar > nfile:
        v = np.concatenate([v, np.ones((h.nvar - nfile,) + v.shape[1:])])
    h.set_interior(v)
    sim = Simulation(h, Runtime(cfl=g["cfl"], cfl_max_var=g["cfl_max_var"], tstop=g["tstop"], first_dt=g["first_dt"]))
    sim.run(maxsteps=len(g["data"]) - 1)
    assert abs(sim.g_time - g["steps"][-1, 1]) <= 1e-11 * g["steps"][-1, 1]
    assert rel_l1(h.get_interior()[:nfile], g["data"][-1]) <= TOL_RUN
    h.close()


@pytest.mark.parametrize("geometry", ["SPHERICAL", "CARTESIAN"])
@pytest.mark.parametrize("limiter,char,flat,entr", [("DEFAULT", False, False, False), ("DEFAULT", True, True, False),
                                                    ("MC_LIM", True, False, True), ("OSPRE_LIM", False, True, True),
                                                    ("VANALBADA_LIM", True, True, True), ("UMIST_LIM", False, False, False),
                                                    ("MINMOD_LIM", True, False, False),
                                                    ("DEFAULT", False, False, "SELECTIVE"), ("VANLEER_LIM", True, True, "SELECTIVE")])
@pytest.mark.parametrize("rk,solver", [("RK2", "hllc"), ("RK3", "hll"), ("EULER", "tvdlf")])
def test_gen_options_vs_oracle(Hydro, geometry, limiter, char, flat, entr, rk, solver):
    """Every limiter / limiting mode / flattening / entropy / solver / RK mix on a stretched 2-D grid
    with shocks (so that FlagShock flags zones), tracers and a space-dependent body force."""
    grid = [(0.8, 37, 3.1, "r", 1.02), (0.35, 29, 1.4, "r", 0.985), (0.0, 1, 1.0)]
    kw = dict(dimensions=2, grid=grid, geometry=geometry, gamma=1.4, time_stepping=rk, solver=solver, limiter=limiter,
              bcs=("reflective", "outflow", "outflow", "reflective", "periodic", "periodic"), ntracer=2, body_force=1,
              char_limiting=char, shock_flattening=flat, entropy_switch=entr)
    o = GenOracle(**kw)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    rng = np.random.default_rng(5)
    for comp in range(3):
        tab = rng.uniform(-1.0, 1.0, size=o.shape[1:])
        o.set_body_force_vector(comp, tab); h.set_body_force_vector(comp, tab)
    from common import random_state
    v5 = random_state((1, 29, 37), seed=3, smooth=False)
    tr = np.stack([np.clip(np.sin(5 * v5[0]), 0, 1), np.clip(np.cos(3 * v5[4]), 0, 1)])
    v = np.concatenate([v5, tr] + ([np.ones((1, 1, 29, 37))] if entr else []))
    vc = o.embed(v)
    h.set_interior(v)
    dt = 1e-4
    for n in range(3):
        inv, mach, nf = o.advance_step(vc, dt)
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), vc[o.interior()])
        assert e <= TOL_STEP, (n, e)
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        assert abs(info.maxMach - mach) <= 1e-11 * mach
        h.set_interior(vc[o.interior()])
    h.close(); o.close()


@pytest.mark.parametrize("name", __import__("common").ISO_CASES)
def test_isothermal_per_step_vs_reference_dumps(Hydro, name):
    """EOS ISOTHERMAL on the CUDA path (Src/EOS/Isothermal, NFLX = 4: p = cs^2 rho in the fluxes and in FlagShock,
    the isothermal HLLC star state hllc.c:137-150, the isothermal eigenvectors eigenv.c:175-196) against the dumps
    of the compiled reference (user files oracle/problems/iso): Cartesian 2-D / 3-D, spherical 2-D with gravity,
    characteristic limiting and MULTID flattening, HLL / HLLC / TVDLF."""
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    assert kw["eos"] == "ISOTHERMAL"
    h = Hydro(**hydro_kwargs_from_gen(kw))
    set_point_mass_gravity(h, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    assert h.nvar == data.shape[1] == 4 + g["ntracer"]
    for n in range(len(data) - 1):
        h.set_interior(data[n])
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), data[n + 1])
        assert e <= TOL_STEP, (name, n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], (name, n)
    h.close()


@pytest.mark.parametrize("name", __import__("common").CURV_GPU_CASES)
def test_cylindrical_polar_potential_per_step_vs_reference_dumps(Hydro, name):
    """GEOMETRY CYLINDRICAL (r, z) and POLAR (r, phi[, z]) and BODY_FORCE POTENTIAL (and VECTOR + POTENTIAL) on
    spherical grids, on the CUDA path against the dumps of the compiled reference (user files oracle/problems/cyl,
    sph): volumes / areas / centroids of set_geometry.c, the |r| weighting of the angular-momentum flux
    (rhs.c:535-538, :268) with iMPHI = VX2 for POLAR, the centrifugal source on (vp + vm)/2 (rhs_source.c:201-227),
    r dphi in the polar C_dt, the AXISYMMETRIC flip of iVPHI, Phi at the faces in the energy flux (rhs.c:171-179) and
    the potential gradient / work terms (rhs_source.c:274-279,378-383,442-447).  roe_* / twoshock_*: Roe_Solver
    (HD/roe.c, both equations of state) and TwoShock_Solver (HD/two_shock.c); oned_*: SHOCK_FLATTENING ONED
    (States/flatten.c, 4 ghost zones); ppmg_*: RECONSTRUCTION PARABOLIC + RK3 with the weights of PPM_CoefficientsSet
    (States/ppm_coeffs.c: LU solve on stretched grids, closed forms on uniform radial grids, Gauss moments of
    sin(theta)), with and without CHAR_LIMITING / MULTID flattening (States/ppm_states.c); ring_*: RING_AVERAGE
    (Src/ring_average.c) with the polaraxis boundary in POLAR and SPHERICAL geometry."""
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    set_point_mass_gravity(h, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    for n in range(len(data) - 1):
        h.set_interior(data[n])
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), data[n + 1])
        assert e <= TOL_STEP, (name, n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], (name, n)
    h.close()


def test_isothermal_line_driven_wind_vs_reference_dumps(Hydro):
    """Test_Problems/LineDrivenWind/cv_iso, the fork's isothermal wind problem (unmodified user files in the
    reference run that made the fixture): line force with T = T_ISO (line_connect.c:851-855), floors and user
    boundaries without their pressure parts (init.c "#if EOS != ISOTHERMAL"), NVAR = 5 (tracer at index 4)."""
    g = load_golden("iso_ldw_hll")
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(h, h.x(0), h.x(1))
    data, steps = g["data"], g["steps"]
    assert h.nvar == data.shape[1] == 5
    for n in range(len(data) - 1):
        h.set_interior(data[n])
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), data[n + 1])
        assert e <= TOL_STEP, (n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], n
    h.close()


LDW_CASES = [c for c in GEN_CASES if c.startswith("ldw_nocool")]


def _with_entr(v, nvar):
    return v if v.shape[0] == nvar else np.concatenate([v, np.ones((nvar - v.shape[0],) + v.shape[1:])])


@pytest.mark.parametrize("name", LDW_CASES)
def test_ldw_per_step_vs_reference_dumps(Hydro, name):
    """C4: the reference's line-driven disc wind (cv_idl, unmodified user files) with synthetic
    sirocco flux tables: per-step parity against the dumps of the reference executable."""
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(h, h.x(0), h.x(1), fit="fit" in name)      # ..._fit_...: M(t) fit tables instead of k t^alpha
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]
    for n in range(len(data) - 1):
        h.set_interior(_with_entr(data[n], h.nvar))
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior()[:nfile], data[n + 1])
        assert e <= TOL_STEP, (name, n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], (name, n)
    h.close()


def test_ldw_bench_grid_1024x512_vs_oracle(Hydro):
    """C4 at the size BASELINE.json names: the line-driven wind on the 1024 x 512 r-theta grid of bench.py (same grid
    ratios, same synthetic 36-angle tables handed over directly, same initial state), 4 free-running RK2 steps of the
    library against the general oracle: per-step tolerance on every variable and on the dt the step returns (each step
    starts from the oracle's state, like the per-step tests on the reference dumps)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench
    from common import LDW_BCS, LDW_PARAMS, LDW_UNITS, ldw_flux_tables
    n1, n2 = 1024, 512
    grid = [(0.87, n1, 8.7, "r", 1.005), (0.0, n2, float(np.radians(90.0)), "r", 0.995), (0.0, 1, 1.0)]
    kw = dict(dimensions=2, grid=grid, geometry="SPHERICAL", gamma=5. / 3., time_stepping="RK2", solver="hll",
              limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1, char_limiting=True,
              shock_flattening=True, entropy_switch=True, nghost=3)
    o = GenOracle(**kw)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    x1, x2 = o.x(0), o.x(1)
    gm_code = 6.6726e-8 * LDW_PARAMS["CENT_MASS"] / (LDW_UNITS["length"] * LDW_UNITS["velocity"] ** 2)
    fr, ft, fp = ldw_flux_tables(x1, x2, roundtrip=False)
    for obj in (o, h):
        obj.set_body_force_vector(0, (-gm_code / (x1 * x1)).reshape(1, 1, -1))
        obj.set_body_force_vector(1, np.zeros((1, 1, 1)))
        obj.set_body_force_vector(2, np.zeros((1, 1, 1)))
        obj.set_ldw(params=LDW_PARAMS, units=LDW_UNITS, flux_r=fr, flux_t=ft, flux_p=fp)
    v = bench.ldw_state(x1[3:-3], x2[3:-3], LDW_PARAMS, LDW_UNITS)
    assert v.shape == (7, 1, n2, n1)
    vc = o.embed(v)
    h.set_interior(v)
    dt = dto = 1e-4
    for n in range(4):
        inv, mach, nf = o.advance_step(vc, dto)
        info = h.advance_step(dt)
        got, ref = h.get_interior(), vc[o.interior()]
        # 3.1 M values per step: all of them within 1e-11 and all but a handful within the 1e-12 contract.  The
        # stragglers (measured: one zone per step, 3e-12 of the velocity scale) sit on the disc / wind interface, where
        # the density jumps by ten decades and the round-off of the dense neighbour's momentum flux lands on a zone
        # of 1e-10 the density (same effect, same remark: test_ldw_floors_and_boundaries_vs_oracle)
        assert rel_err(got[:6], ref[:6]) <= 1e-11, (n, rel_err(got[:6], ref[:6]))
        nbad = 0
        for nv in range(6):
            scale = np.abs(ref[1:4]).max() if 1 <= nv <= 3 else np.abs(ref[nv]).max()
            nbad += int((np.abs(got[nv] - ref[nv]) > TOL_STEP * scale).sum())
        assert nbad <= 4, (n, nbad)
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        dt = dto = min(0.4 / inv, 1.1 * dto)
        h.set_interior(ref)
    h.close(); o.close()


@pytest.mark.parametrize("alpha", [-0.6, -0.7, -0.45])
def test_ldw_force_multiplier_exponents_vs_oracle(Hydro, alpha):
    """M(t) = k t^alpha of LineForce() (line_connect.c:858-870) for the exponent cv_idl ships (-0.6: the per-bin
    dvds^0.6 goes through pow_three_fifths(), a Newton fifth root) and for others (exp / log): three steps from a
    developed wind state against the oracle, which calls libm's pow() like the reference."""
    from common import LDW_BCS, LDW_PARAMS, LDW_UNITS, ldw_flux_tables
    grid = [(0.87, 48, 8.7, "r", 1.05), (0.0, 36, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    kw = dict(dimensions=2, grid=grid, geometry="SPHERICAL", gamma=5. / 3., time_stepping="RK2", solver="hll",
              limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1, char_limiting=True,
              shock_flattening=True, entropy_switch=True, nghost=3)
    o = GenOracle(**kw)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    params = dict(LDW_PARAMS, ALPHARAD=alpha)
    gm_code = 6.6726e-8 * params["CENT_MASS"] / (LDW_UNITS["length"] * LDW_UNITS["velocity"] ** 2)
    for obj in (o, h):
        x1, x2 = obj.x(0), obj.x(1)
        fr, ft, fp = ldw_flux_tables(x1, x2)
        obj.set_body_force_vector(0, (-gm_code / (x1 * x1)).reshape(1, 1, -1))
        obj.set_body_force_vector(1, np.zeros((1, 1, 1)))
        obj.set_body_force_vector(2, np.zeros((1, 1, 1)))
        obj.set_ldw(params=params, units=LDW_UNITS, flux_r=fr, flux_t=ft, flux_p=fp)
    g = load_golden("ldw_nocool_hll")
    v = np.zeros((7, 1, 36, 48))
    v[:6] = g["data"][5]
    v[6] = 1.0
    vc = o.embed(v); h.set_interior(v)
    dt = float(g["steps"][5, 2])
    for n in range(3):
        inv, mach, nf = o.advance_step(vc, dt)
        info = h.advance_step(dt)
        got, ref = h.get_interior(), vc[o.interior()]
        assert rel_err(got[:6], ref[:6]) <= TOL_STEP, (alpha, n, rel_err(got[:6], ref[:6]))
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        h.set_interior(ref)
    h.close(); o.close()


def test_ldw_floors_and_boundaries_vs_oracle(Hydro):
    """User boundaries of cv_idl on a state that triggers the density / pressure floors (also
    inside stage 2, where Uc is re-derived), the mid-plane reset and the hybrid X2_BEG fill."""
    from common import LDW_BCS
    grid = [(0.87, 40, 8.7, "r", 1.05), (0.0, 30, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    kw = dict(dimensions=2, grid=grid, geometry="SPHERICAL", gamma=5. / 3., time_stepping="RK2", solver="hll",
              limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1, char_limiting=True,
              shock_flattening=True, entropy_switch=True, nghost=3)
    o = GenOracle(**kw)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(o, o.x(0), o.x(1)); ldw_setup(h, h.x(0), h.x(1))
    g = load_golden("ldw_nocool_hll")
    rng = np.random.default_rng(11)
    v = np.zeros((7, 1, 30, 40))
    base = g["data"][3][:, :, :30, :40]
    v[:6] = base
    # holes and cold spots in the tenuous wind region only: a floored zone (rho = 1e-10) next to
    # disc material (rho ~ 1e5) would turn the round-off of the dense neighbour's flux into an
    # O(1e-7) velocity error, which says nothing about the boundary code under test
    wind = base[0:1] < 1e-8
    holes = wind[0] & (rng.random(size=(1, 30, 40)) < 0.25)   # zones below the density floor
    v[0][holes] *= 0.2
    cold = wind[0] & (rng.random(size=(1, 30, 40)) < 0.15)    # zones below the pressure floor
    v[4][cold] *= 1e-4
    assert holes.sum() > 20 and cold.sum() > 10
    v[1] += 1e-3 * rng.normal(size=(1, 30, 40)); v[2] += 1e-3 * rng.normal(size=(1, 30, 40))
    v[6] = 1.0
    vc = o.embed(v); h.set_interior(v)
    # Boundary() alone
    o.boundary(vc); h.boundary()
    assert rel_err(h.download(), vc) <= 1e-13
    dt = 2e-5
    for n in range(3):
        inv, mach, nf = o.advance_step(vc, dt)
        info = h.advance_step(dt)
        got, ref = h.get_interior(), vc[o.interior()]
        e = rel_err(got, ref)
        d = np.abs(got - ref)
        w = np.unravel_index(np.argmax(d / np.abs(ref).max(axis=(1, 2, 3), keepdims=True)), d.shape)
        assert e <= TOL_STEP, (n, e, w, got[w], ref[w], ref[:, w[1], w[2], w[3]])
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        h.set_interior(vc[o.interior()])
    h.close(); o.close()


def test_ldw_with_blondin_cooling_vs_reference_dumps(Hydro):
    """C4 with the BLONDIN source step in the Strang order of Src/main.c:479-485 (even steps:
    AdvanceStep then SplitSource, odd steps the reverse) against the reference's dumps.
    BlondinCooling() stops its Brent iteration at |dT| <= 1 K: transcendental functions that
    differ in the last ulp from glibc's can end it one iteration earlier or later in a zone that
    sits on that threshold, which moves T_f by up to the solver's own tolerance (1 K of >= 1e4 K).
    So the bulk of the zones must agree to the per-step tolerance and none may be off by more
    than the reference's own convergence criterion."""
    g = load_golden("ldw_cool_hll")
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(h, h.x(0), h.x(1))
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]
    bad = tot = 0
    for n in range(len(data) - 1):
        h.set_interior(_with_entr(data[n], h.nvar))
        dt, t = steps[n, 2], steps[n, 1]
        if n % 2 == 0:
            h.advance_step(dt); h.split_source(dt, t)
        else:
            h.split_source(dt, t); h.advance_step(dt)
        got, ref = h.get_interior()[:nfile], data[n + 1]
        # every step within the per-step contract, pressure included: the cooling solve runs the reference's
        # arithmetic (csrc/pb200_cool.cu: no FMA contraction, IEEE sqrt / division, glibc's exp / pow / log10 bit for
        # bit), so the Brent iterates are the reference's and the solve is well conditioned (test_glibc_math.py)
        for nv in (0, 1, 2, 3, 4, 5):
            scale = np.abs(ref[1:4]).max() if 1 <= nv <= 3 else np.abs(ref[nv]).max()
            assert np.abs(got[nv] - ref[nv]).max() <= TOL_STEP * scale, (n, nv, np.abs(got[nv] - ref[nv]).max() / scale)
        relp = np.abs(got[4] - ref[4]) / ref[4]
        assert relp.max() <= 1e-11, (n, relp.max())       # zone by zone (pressure spans 8 decades)
    h.close()


def test_blondin_cooling_vs_oracle_with_tables(Hydro):
    """BlondinCooling with non-trivial prefactor tables, both ionisation-parameter branches
    (g_time <= 3: analytic; > 3: sirocco_xi / sirocco_t_r tables), temperatures from below the 1e4 K
    cut-off to 1e8 K so that net heating, net cooling and the equilibrium re-solve all occur."""
    from common import LDW_BCS
    grid = [(0.87, 40, 8.7, "r", 1.05), (0.0, 30, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    kw = dict(dimensions=2, grid=grid, geometry="SPHERICAL", gamma=5. / 3., time_stepping="RK2", solver="hll",
              limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1, char_limiting=True,
              shock_flattening=True, entropy_switch=True, nghost=3)
    o = GenOracle(**kw)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(o, o.x(0), o.x(1)); ldw_setup(h, h.x(0), h.x(1))
    rng = np.random.default_rng(3)
    full = o.shape[1:]
    tabs = [rng.uniform(0.3, 3.0, size=full) for _ in range(5)]
    tabs += [10.0 ** rng.uniform(-2, 5, size=full), rng.uniform(5e4, 5e5, size=full)]
    h.set_cooling_tables(tabs)
    KELVIN_MU = (1e9 ** 2) * 1.66053886e-24 / 1.3806505e-16 * 0.6
    v = np.zeros((7, 1, 30, 40))
    v[0] = 10.0 ** rng.uniform(-2, 4, size=(1, 30, 40))
    T = 10.0 ** rng.uniform(3.5, 8.0, size=(1, 30, 40))
    v[4] = v[0] * T / KELVIN_MU
    v[5] = 0.5; v[6] = 1.0
    for g_time, dt in ((1.0, 1e-2), (5.0, 1.0), (5.0, 30.0)):
        vc = o.embed(v); h.set_interior(v)
        o.blondin_cooling(vc, dt, g_time, tabs)
        h.split_source(dt, g_time)
        got, ref = h.get_interior(), vc[o.interior()]
        assert np.array_equal(got[[0, 1, 2, 3, 5]], ref[[0, 1, 2, 3, 5]])
        relp = np.abs(got[4] - ref[4]) / ref[4]
        changed = np.abs(ref[4] - v[4]) / v[4]
        assert (changed > 1e-3).mean() > 0.2, "test state did not exercise the cooling"
        # same input bits, same arithmetic (csrc/pb200_cool.cu: -fmad=false, IEEE sqrt and division, glibc's exp /
        # pow / log10 restated bit for bit in csrc/glibc_math.cuh): the same sequence of Brent iterates
        assert np.array_equal(got[4], ref[4]), (relp.max(), int((relp > 0).sum()))
    h.close(); o.close()


def test_ldw_cooling_main_loop_reproduces_reference_dt_sequence(Hydro):
    """Simulation(cooling=True) mirrors Integrate / NextTimeStep with COOLING != NO (Strang order, dt
    renewed every second step, the double division of invDt_hyp by DIMENSIONS across a pair of steps,
    Src/main.c:326-330,406-415,479-485, update_stage.c:391-392): free-running from the first dump it
    must reproduce the reference's (t, dt) records."""
    from pluto_sirocco_b200 import Runtime, Simulation
    g = load_golden("ldw_cool_hll")
    kw = gen_kwargs_from_golden(g)
    h = Hydro(**hydro_kwargs_from_gen(kw))
    ldw_setup(h, h.x(0), h.x(1))
    h.set_interior(_with_entr(g["data"][0], h.nvar))
    sim = Simulation(h, Runtime(cfl=g["cfl"], cfl_max_var=g["cfl_max_var"], tstop=g["tstop"], first_dt=g["first_dt"]),
                     cooling=True)
    nsteps = len(g["data"]) - 1
    sim.run(maxsteps=nsteps)
    for (n, t, dt), ref in zip(sim.history, g["steps"]):
        assert n == int(ref[0]) and abs(t - ref[1]) <= 1e-9 * max(ref[1], 1e-30) and abs(dt - ref[2]) <= 1e-7 * ref[2], (n, t, dt, ref)
    got, ref = h.get_interior()[:6], g["data"][nsteps]
    relp = np.abs(got[4] - ref[4]) / ref[4]
    assert relp.max() <= 5e-4 and np.abs(got[0] - ref[0]).max() <= 1e-6 * np.abs(ref[0]).max()
    h.close()


@pytest.mark.parametrize("name,ref_cfg", [("ring_pol2d_vl", "pol2d_ring_vl"), ("oned_sph2d_char_roe", "sph2d_char_oned")])
def test_hydro_from_definitions_and_pluto_ini_reproduces_the_reference_dumps(Hydro, tmp_path, name, ref_cfg):
    """The unchanged user surface end to end in Python: Definitions.parse(definitions.h) + Runtime.parse(pluto.ini) ->
    Hydro.from_files() -> the per-step dumps of the reference executable that was compiled from that definitions.h."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import build_ref
    import pluto_grid
    import refrun
    from pluto_sirocco_b200.hydro import Definitions, Runtime
    g = load_golden(name)
    cfg = build_ref.CONFIGS[ref_cfg]
    d = Definitions.parse(build_ref.patch_definitions((build_ref.HERE / "problems" / cfg["local"] / "definitions.h").read_text(),
                                                      cfg["overrides"]))
    grid = [(float(r[0]), int(r[1]), float(r[2]), "r", float(r[4])) if r[3] == 1.0 else (float(r[0]), int(r[1]), float(r[2]))
            for r in g["gridspec"]]
    refrun.write_ini(tmp_path / "pluto.ini", grid=[pluto_grid.ini_string(s) for s in grid], cfl=float(g["cfl"]),
                     tstop=float(g["tstop"]), first_dt=float(g["first_dt"]), solver=str(g["solver"]),
                     bcs=tuple(str(b) for b in g["bcs"]))
    h = Hydro.from_files(d, Runtime.parse(tmp_path / "pluto.ini"), gamma=float(g["gamma"]))
    set_point_mass_gravity(h, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    for n in range(3):
        h.set_interior(data[n])
        info = h.advance_step(steps[n, 2])
        assert rel_err(h.get_interior(), data[n + 1]) <= TOL_STEP, (name, n)
    h.close()
