"""CPU: the general-grid oracle (oracle/gen_oracle.c) against golden dumps of the compiled
reference on spherical / stretched grids, with characteristic limiting, MULTID flattening and the
entropy switch (user files oracle/problems/sph)."""
import numpy as np
import pytest

from common import GEN_CASES, gen_kwargs_from_golden, ldw_setup, load_golden, rel_err, set_point_mass_gravity
from gen_oracle import GenOracle

SPH_CASES = [c for c in GEN_CASES if c.startswith("sph")]


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


LDW_CASES = [c for c in GEN_CASES if c.startswith("ldw")]


@pytest.mark.parametrize("name", LDW_CASES)
def test_gen_oracle_ldw_per_step_matches_reference_dumps(name):
    """The line-driven wind problem of the reference (cv_idl user files, unmodified) with synthetic
    sirocco flux tables: VGradCalc + LineForce, user boundaries and floors, entropy switch, MULTID
    flattening, characteristic limiting, tracer, spherical stretched grid."""
    g = load_golden(name)
    o = GenOracle(**gen_kwargs_from_golden(g))
    ldw_setup(o, o.x(0), o.x(1))
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
