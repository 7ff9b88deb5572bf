"""CPU: the general-grid oracle (oracle/gen_oracle.c) against golden dumps of the compiled
reference on spherical / stretched grids, with characteristic limiting, MULTID flattening and the
entropy switch (user files oracle/problems/sph)."""
import numpy as np
import pytest

from common import (CURV_CASES, GEN_CASES, gen_kwargs_from_golden, ldw_setup, load_golden, rel_err,
                    set_point_mass_gravity)
from gen_oracle import GenOracle

SPH_CASES = [c for c in GEN_CASES if c.startswith(("sph", "cart"))]


@pytest.mark.parametrize("name", SPH_CASES)
def test_gen_oracle_per_step_matches_reference_dumps(name):
    g = load_golden(name)
    o = GenOracle(**gen_kwargs_from_golden(g))
    set_point_mass_gravity(o, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]            # ENTR is not part of the dumps
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt = steps[n, 2]
        inv, mach, nf = o.advance_step(vc, dt)
        got = vc[o.interior()][:nfile]
        assert np.array_equal(got, data[n + 1]), (name, n, rel_err(got, data[n + 1]))
        dtn = o.next_time_step(inv, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert dtn == steps[n + 1, 2], (name, n)
    o.close()


@pytest.mark.parametrize("name", [c for c in CURV_CASES if "ldw" not in c])
def test_gen_oracle_cylindrical_polar_isothermal_match_reference_dumps(name):
    """GEOMETRY CYLINDRICAL (r, z) and POLAR (r, phi[, z]) of the oracle against the compiled reference
    (user files oracle/problems/cyl): volumes / areas / centroids of set_geometry.c, the |r| weighting of
    the angular-momentum flux (rhs.c:535-538, :268), the centrifugal source on (vp + vm)/2
    (rhs_source.c:201-227), r dphi in the polar C_dt, the AXISYMMETRIC flip of iVPHI on the axis with
    r < 0 ghost zones, with and without characteristic limiting + MULTID flattening; and EOS ISOTHERMAL
    (user files oracle/problems/iso; the equation of state of the fork's LineDrivenWind/cv_iso problem):
    NFLX = 4 state vector, p = cs^2 rho in the fluxes and in FlagShock, the isothermal HLLC star state
    (hllc.c:137-150) and eigenvectors (eigenv.c:175-196), Cartesian 2-D / 3-D and spherical with gravity.
    The roe_* fixtures add Roe_Solver (HD/roe.c: Roe average, entropy fix, HLL inside strong shocks and
    flagged zones) for both equations of state, the twoshock_* ones TwoShock_Solver (HD/two_shock.c),
    the oned_* ones SHOCK_FLATTENING ONED (States/flatten.c, 4 ghost zones), the ppmg_* ones RECONSTRUCTION
    PARABOLIC + RK3 with the general-grid weights of States/ppm_coeffs.c (PPM_FindWeights through the LU
    solve on stretched grids, the closed forms on uniform cylindrical / spherical radial grids, the 5-point
    Gauss moments of sin(theta) for the meridional direction, PPM_Q6_Coeffs), the pot_* ones BODY_FORCE
    POTENTIAL (and VECTOR + POTENTIAL) on spherical and polar grids, the ring_* ones RING_AVERAGE (Src/ring_average.c:
    RingAverageCons around every stage, RingAverageReconstruct with MP5 / van Leer on the reduced grid, the chunked C_dt)
    with the polaraxis boundary (boundary.c:770-840) in POLAR 2-D / 3-D and SPHERICAL 3-D.
    These fixtures pin the oracle; the CUDA path runs every one of them too (tests/test_gpu_gen.py: ISO_CASES,
    CURV_GPU_CASES)."""
    g = load_golden(name)
    o = GenOracle(**gen_kwargs_from_golden(g))
    set_point_mass_gravity(o, float(g["gm"]))
    data, steps = g["data"], g["steps"]
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt = steps[n, 2]
        inv, mach, nf = o.advance_step(vc, dt)
        got = vc[o.interior()]
        assert np.array_equal(got, data[n + 1]), (name, n, rel_err(got, data[n + 1]))
        dtn = o.next_time_step(inv, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert dtn == steps[n + 1, 2], (name, n)
    o.close()


def test_gen_oracle_isothermal_ldw_matches_reference_dumps():
    """Test_Problems/LineDrivenWind/cv_iso (unmodified user files, EOS ISOTHERMAL): line force with
    T = T_ISO (line_connect.c:851-855), floors and user boundaries without their pressure parts
    (init.c #if EOS != ISOTHERMAL blocks), g_isoSoundSpeed of init.c:61-63."""
    g = load_golden("iso_ldw_hll")
    o = GenOracle(**gen_kwargs_from_golden(g))
    ldw_setup(o, o.x(0), o.x(1))
    data, steps = g["data"], g["steps"]
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt = steps[n, 2]
        inv, mach, nf = o.advance_step(vc, dt)
        got = vc[o.interior()]
        assert np.array_equal(got, data[n + 1]), (n, rel_err(got, data[n + 1]))
        dtn = o.next_time_step(inv, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert dtn == steps[n + 1, 2], n
    o.close()


LDW_CASES = [c for c in GEN_CASES if c.startswith("ldw_nocool")]


@pytest.mark.parametrize("name", LDW_CASES)
def test_gen_oracle_ldw_per_step_matches_reference_dumps(name):
    """The line-driven wind problem of the reference (cv_idl user files, unmodified) with synthetic
    sirocco flux tables: VGradCalc + LineForce, user boundaries and floors, entropy switch, MULTID
    flattening, characteristic limiting, tracer, spherical stretched grid."""
    g = load_golden(name)
    o = GenOracle(**gen_kwargs_from_golden(g))
    ldw_setup(o, o.x(0), o.x(1), fit="fit" in name)
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt = steps[n, 2]
        inv, mach, nf = o.advance_step(vc, dt)
        got = vc[o.interior()][:nfile]
        assert np.array_equal(got, data[n + 1]), (name, n, rel_err(got, data[n + 1]))
        dtn = o.next_time_step(inv, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert dtn == steps[n + 1, 2], (name, n)
    o.close()


def test_gen_oracle_ldw_with_blondin_cooling_matches_reference_dumps():
    """COOLING BLONDIN in the Strang order of Src/main.c:479-485 (even steps: AdvanceStep then
    SplitSource, odd steps the reverse; dt renewed every second step), prefactors at their defaults
    (1.0, line_connect.c:383-393), analytic ionisation parameter (g_time <= 3): bit-exact."""
    g = load_golden("ldw_cool_hll")
    o = GenOracle(**gen_kwargs_from_golden(g))
    ldw_setup(o, o.x(0), o.x(1))
    data, steps = g["data"], g["steps"]
    nfile = data.shape[1]
    tabs = [np.ones((1, 1, 1))] * 5 + [np.zeros((1, 1, 1))] * 2
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt, t = steps[n, 2], steps[n, 1]
        if n % 2 == 0:
            o.advance_step(vc, dt); o.blondin_cooling(vc, dt, t, tabs)
        else:
            o.blondin_cooling(vc, dt, t, tabs); o.advance_step(vc, dt)
        assert np.array_equal(vc[o.interior()][:nfile], data[n + 1]), n
        if n % 2 == 1:      # NextTimeStep every second step (main.c:326-330)
            assert steps[n + 1, 2] != steps[n, 2]
        else:
            assert steps[n + 1, 2] == steps[n, 2]
    o.close()
