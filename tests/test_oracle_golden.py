"""CPU: the oracle (C restatement) against the golden dumps of the compiled reference."""
import numpy as np
import pytest

from common import GOLDEN_CASES, kwargs_from_golden, load_golden, rel_err, set_gravity_x2
from oracle import Oracle


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_per_step_matches_reference_dumps(name):
    g = load_golden(name)
    o = Oracle(**kwargs_from_golden(g))
    set_gravity_x2(o, g["body_force"], g["grav"])
    data, steps = g["data"], g["steps"]
    for n in range(len(data) - 1):
        vc = o.embed(data[n])
        dt = steps[n, 2]
        inv, mach, nf = o.advance_step(vc, dt)
        # the restatement keeps the reference's operation order: bit-exact
        assert np.array_equal(vc[o.interior()], data[n + 1]), (name, n, rel_err(vc[o.interior()], data[n + 1]))
        dtn = o.next_time_step(inv, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert dtn == steps[n + 1, 2], (name, n)
        assert steps[n + 1, 1] == steps[n, 1] + dt


@pytest.mark.parametrize("name", ["sod_plm_hllc", "sedov3d_plm_hllc", "sedov3d_ppm_hllc"])
def test_oracle_integrate_reproduces_reference_run(name):
    g = load_golden(name)
    o = Oracle(**kwargs_from_golden(g))
    data, steps = g["data"], g["steps"]
    vc = o.embed(data[0])
    n, t, dt = o.integrate(vc, len(data) - 1, t=0.0, dt=g["first_dt"], tstop=g["tstop"], cfl=g["cfl"],
                           cfl_max_var=g["cfl_max_var"], first_dt=g["first_dt"])
    assert n == len(data) - 1
    assert t == steps[-1, 1] and dt == steps[-1, 2]
    assert np.array_equal(vc[o.interior()], data[-1])


def test_oracle_boundary_fills():
    o = Oracle(dimensions=2, nx=(6, 5, 1), bcs=("reflective", "outflow", "periodic", "periodic", "outflow", "outflow"))
    rng = np.random.default_rng(0)
    vc = o.embed(rng.uniform(0.5, 1.5, size=(5, 1, 5, 6)))
    o.boundary(vc)
    ng = 2
    # reflective x1-beg: mirror with v_x1 flipped (boundary.c:503,701)
    assert np.array_equal(vc[0, 0, ng:-ng, ng - 1], vc[0, 0, ng:-ng, ng])
    assert np.array_equal(vc[1, 0, ng:-ng, ng - 2], -vc[1, 0, ng:-ng, ng + 1])
    # outflow x1-end copies IEND
    assert np.array_equal(vc[4, 0, ng:-ng, -1], vc[4, 0, ng:-ng, -ng - 1])
    # periodic x2 wraps, including the already-filled x1 ghosts (side order)
    assert np.array_equal(vc[:, 0, 0, :], vc[:, 0, 5, :])
    assert np.array_equal(vc[:, 0, -1, :], vc[:, 0, ng + 1, :])
