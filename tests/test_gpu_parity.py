"""GPU: the CUDA path (through the C ABI) against the oracle, the golden dumps of the
compiled reference, and - when oracle/_ref travelled to the box - the reference executable
itself.  Tolerances are the north star's: <=1e-12 relative per step, <=1e-9 relative L1
after a run."""
import itertools

import numpy as np
import pytest

from common import (GOLDEN_CASES, TOL_RUN, TOL_STEP, kwargs_from_golden, load_golden, random_state,
                    rel_err, rel_l1, set_gravity_x2)
from oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Hydro(cuda_lib):
    from pluto_sirocco_b200 import Hydro as H
    return H


# ---------------------------------------------------------------------------- golden ---
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_per_step_vs_reference_dumps(Hydro, name):
    g = load_golden(name)
    h = Hydro(**kwargs_from_golden(g))
    set_gravity_x2(h, g["body_force"], g["grav"])
    data, steps = g["data"], g["steps"]
    worst = 0.0
    for n in range(len(data) - 1):
        h.set_interior(data[n])
        dt = steps[n, 2]
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), data[n + 1])
        worst = max(worst, e)
        assert e <= TOL_STEP, (name, n, e)
        dtn = h.next_time_step(info.invDt_hyp, g["cfl"], g["cfl_max_var"], dt, g["first_dt"])
        assert abs(dtn - steps[n + 1, 2]) <= TOL_STEP * steps[n + 1, 2], (name, n)
    h.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_full_run_vs_reference_dumps(Hydro, name):
    """Free-running (own dt sequence) from the initial dump to the last dump."""
    from pluto_sirocco_b200 import Runtime, Simulation
    g = load_golden(name)
    h = Hydro(**kwargs_from_golden(g))
    set_gravity_x2(h, g["body_force"], g["grav"])
    h.set_interior(g["data"][0])
    rt = Runtime(cfl=g["cfl"], cfl_max_var=g["cfl_max_var"], tstop=g["tstop"], first_dt=g["first_dt"])
    sim = Simulation(h, rt)
    nsteps = len(g["data"]) - 1
    sim.run(maxsteps=nsteps)
    assert abs(sim.g_time - g["steps"][-1, 1]) <= 1e-11 * g["steps"][-1, 1]
    assert rel_l1(h.get_interior(), g["data"][-1]) <= TOL_RUN
    # the C-side loop gives the same answer as the Python-side loop
    h2 = Hydro(**kwargs_from_golden(g))
    set_gravity_x2(h2, g["body_force"], g["grav"])
    h2.set_interior(g["data"][0])
    n, t, dt = h2.integrate(nsteps, t=0.0, dt=g["first_dt"], tstop=g["tstop"], cfl=g["cfl"],
                            cfl_max_var=g["cfl_max_var"], first_dt=g["first_dt"])
    assert n == nsteps and t == sim.g_time and dt == sim.g_dt
    assert np.array_equal(h2.get_interior(), h.get_interior())
    h.close(); h2.close()


# ---------------------------------------------------------------------------- oracle ---
CONFIGS = list(itertools.product(("LINEAR", "PARABOLIC"), ("hllc", "hll", "tvdlf")))
BCSETS = {
    "periodic": ("periodic",) * 6,
    "mixed": ("reflective", "outflow", "outflow", "reflective", "periodic", "periodic"),
}


@pytest.mark.parametrize("recon,solver", CONFIGS)
@pytest.mark.parametrize("dims,nx", [(1, (257, 1, 1)), (2, (67, 45, 1)), (3, (37, 19, 23))])
@pytest.mark.parametrize("bcname", list(BCSETS))
def test_steps_vs_oracle_seeded(Hydro, recon, solver, dims, nx, bcname):
    """Ragged sizes, shocks and contacts, every solver x reconstruction x boundary mix."""
    rk = "RK3" if recon == "PARABOLIC" else "RK2"
    kw = dict(dimensions=dims, nx=nx, gamma=1.4, reconstruction=recon, time_stepping=rk, solver=solver,
              bcs=BCSETS[bcname])
    h, o = Hydro(**kw), Oracle(**kw)
    v = random_state((nx[2], nx[1], nx[0]), seed=dims * 100 + len(solver), smooth=False)
    vc = o.embed(v)
    h.set_interior(v)
    dt = 2e-4
    for n in range(4):
        inv, mach, nf = o.advance_step(vc, dt)
        info = h.advance_step(dt)
        got = h.get_interior()
        e = rel_err(got, vc[o.interior()])
        assert e <= TOL_STEP, (n, e)
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        assert abs(info.maxMach - mach) <= 1e-11 * mach
        assert abs(int(info.c2p_failures) - nf) <= 4
        # per-step test: restart the device from the oracle state so errors do not accumulate
        h.set_interior(vc[o.interior()])
        dt = min(o.next_time_step(inv, 0.3, 1.1, dt, 1e-6), 1.1 * dt)
    h.close()


def _state_with_tracers(shape_int, ntr, seed):
    v5 = random_state(shape_int, seed=seed, smooth=False)
    rng = np.random.default_rng(seed + 1)
    tr = [np.clip(0.5 + 0.5 * np.sin(7.0 * v5[0] + k) + rng.normal(0, 0.05, size=v5[0].shape), 0, 1)
          for k in range(ntr)]
    return np.concatenate([v5, np.stack(tr)]) if ntr else v5


@pytest.mark.parametrize("recon,rk", [("LINEAR", "RK2"), ("PARABOLIC", "RK3")])
@pytest.mark.parametrize("dims,nx", [(1, (131, 1, 1)), (2, (53, 37, 1)), (3, (29, 17, 21))])
@pytest.mark.parametrize("ntr,bf", [(1, 0), (2, 1), (0, 2), (1, 3)])
def test_tracers_and_body_force_vs_oracle(Hydro, recon, rk, dims, nx, ntr, bf):
    """NTRACER 1-2 (AdvectFlux, adv_flux.c:61-72) and BODY_FORCE VECTOR / POTENTIAL / both
    (rhs_source.c:253-281, rhs.c:524-526) with tables that vary in all three directions."""
    kw = dict(dimensions=dims, nx=nx, gamma=1.4, reconstruction=recon, time_stepping=rk, solver="hllc",
              bcs=("periodic", "periodic", "reflective", "outflow", "periodic", "periodic"), ntracer=ntr,
              body_force=bf)
    h, o = Hydro(**kw), Oracle(**kw)
    rng = np.random.default_rng(17 * dims + bf)
    full = (o.tot[2], o.tot[1], o.tot[0])
    shapes = [full, (1, full[1], 1), (1, 1, 1), (full[0], 1, full[2])]   # 3-D, x2 only, constant, x1-x3
    if bf & 1:
        for comp in range(3):
            tab = rng.uniform(-2.0, 2.0, size=shapes[comp])
            h.set_body_force_vector(comp, tab); o.set_body_force_vector(comp, tab)
    if bf & 2:
        for where in range(4):
            tab = rng.uniform(-0.5, 0.5, size=shapes[(where + 1) % 4])
            h.set_body_force_potential(where, tab); o.set_body_force_potential(where, tab)
    v = _state_with_tracers((nx[2], nx[1], nx[0]), ntr, seed=dims + 10 * ntr)
    vc = o.embed(v)
    h.set_interior(v)
    dt = 2e-4
    for n in range(3):
        inv, mach, nf = o.advance_step(vc, dt)
        info = h.advance_step(dt)
        e = rel_err(h.get_interior(), vc[o.interior()])
        assert e <= TOL_STEP, (n, e)
        assert abs(info.invDt_hyp - inv) <= TOL_STEP * inv
        h.set_interior(vc[o.interior()])
    h.close()


def test_body_force_tables_are_required(Hydro):
    from pluto_sirocco_b200._lib import PB200Error
    h = Hydro(dimensions=2, nx=(16, 16, 1), body_force=1)
    h.set_interior(np.concatenate([np.ones((1, 1, 16, 16)), np.zeros((3, 1, 16, 16)), np.ones((1, 1, 16, 16))]))
    with pytest.raises(PB200Error):
        h.advance_step(1e-3)          # tables not handed over yet
    with pytest.raises(PB200Error):
        h.set_body_force_potential(0, np.zeros((1, 1, 1)))   # cfg.body_force has no POTENTIAL part
    h.close()


@pytest.mark.parametrize("limiter", ["MINMOD_LIM", "VANLEER_LIM", "MC_LIM", "VANALBADA_LIM", "OSPRE_LIM",
                                     "UMIST_LIM", "FLAT_LIM"])
def test_limiters_vs_oracle(Hydro, limiter):
    kw = dict(dimensions=2, nx=(48, 40, 1), gamma=5. / 3., limiter=limiter, bcs=("periodic",) * 6)
    h, o = Hydro(**kw), Oracle(**kw)
    v = random_state((1, 40, 48), seed=7, smooth=False)
    vc = o.embed(v); h.set_interior(v)
    for n in range(3):
        o.advance_step(vc, 3e-4); h.advance_step(3e-4)
    assert rel_err(h.get_interior(), vc[o.interior()]) <= 3 * TOL_STEP
    h.close()


@pytest.mark.parametrize("rk", ["EULER", "RK2", "RK3"])
@pytest.mark.parametrize("recon", ["FLAT", "LINEAR"])
def test_time_stepping_variants(Hydro, rk, recon):
    kw = dict(dimensions=3, nx=(20, 12, 9), gamma=1.4, reconstruction=recon, time_stepping=rk, nghost=2,
              bcs=("outflow",) * 6)
    h, o = Hydro(**kw), Oracle(**kw)
    v = random_state((9, 12, 20), seed=3, smooth=False)
    vc = o.embed(v); h.set_interior(v)
    for n in range(3):
        o.advance_step(vc, 2e-4); h.advance_step(2e-4)
    assert rel_err(h.get_interior(), vc[o.interior()]) <= 3 * TOL_STEP
    h.close()


def test_minimum_sizes(Hydro):
    """Smallest legal blocks: nx == nghost in every active direction."""
    for recon, ng in (("LINEAR", 2), ("PARABOLIC", 3)):
        kw = dict(dimensions=3, nx=(ng, ng, ng), gamma=1.4, reconstruction=recon,
                  time_stepping="RK2", bcs=("periodic",) * 6)
        h, o = Hydro(**kw), Oracle(**kw)
        v = random_state((ng, ng, ng), seed=11)
        vc = o.embed(v); h.set_interior(v)
        o.advance_step(vc, 1e-3); h.advance_step(1e-3)
        assert rel_err(h.get_interior(), vc[o.interior()]) <= TOL_STEP
        h.close()


def test_floors_match_oracle(Hydro):
    """Cons->prim floors (negative pressure) are applied like Src/HD/mappers.c:206-218."""
    kw = dict(dimensions=1, nx=(64, 1, 1), gamma=1.4, bcs=("outflow",) * 6)
    h, o = Hydro(**kw), Oracle(**kw)
    v = np.zeros((5, 1, 1, 64)); v[0] = 1.0; v[4] = 1e-2
    v[1, ..., :32] = -6.0; v[1, ..., 32:] = 6.0        # Mach-50 rarefaction: near-vacuum, E - kin < 0
    v[1, ..., 20:24] = 3.0
    vc = o.embed(v); h.set_interior(v)
    nf_tot = 0
    for n in range(6):
        inv, mach, nf = o.advance_step(vc, 2.4e-3)
        assert np.isfinite(vc).all()     # (the reference aborts on NaN: CheckNaN, update_stage.c:227)
        info = h.advance_step(2.4e-3)
        # zones whose pressure is 0 +- rounding may fall on either side of the p<0 test, so
        # the COUNT may differ by a few; the floored states must still agree
        assert (info.c2p_failures > 0) == (nf > 0) or abs(int(info.c2p_failures) - nf) <= 4
        assert rel_err(h.get_interior(), vc[o.interior()]) <= TOL_STEP
        nf_tot += nf
        h.set_interior(vc[o.interior()])
    assert nf_tot > 0, "test state did not trigger the floors"
    h.close()


# -------------------------------------------------------------------------- behaviour ---
def test_host_call_equals_resident_call(Hydro):
    kw = dict(dimensions=2, nx=(40, 24, 1), gamma=1.4, bcs=("periodic",) * 6)
    h1, h2 = Hydro(**kw), Hydro(**kw)
    v = random_state((1, 24, 40), seed=5)
    h1.set_interior(v)
    vc = h2.new_vc(); vc[:] = 1.0; vc[1:4] = 0.0; vc[h2.interior()] = v
    for n in range(3):
        h1.advance_step(1e-3)
        h2.advance_step_host(vc, 1e-3)
    assert np.array_equal(h1.get_interior(), vc[h2.interior()])
    h1.close(); h2.close()


@pytest.mark.parametrize("recon,rk,bcs,ntr,bf", [
    ("LINEAR", "RK2", ("reflective", "outflow") * 3, 0, 0),
    ("PARABOLIC", "RK2", ("periodic", "periodic", "outflow", "reflective", "outflow", "reflective"), 1, 1),
    ("PARABOLIC", "RK3", ("outflow", "reflective", "periodic", "periodic", "reflective", "outflow"), 1, 0),
    ("LINEAR", "EULER", ("outflow",) * 6, 0, 0),
])
def test_pipelined_host_call_equals_resident_call_3d(Hydro, monkeypatch, recon, rk, bcs, ntr, bf):
    """pb200_advance_step_host() pipelines upload / all RK stages / download over slabs of x3 planes
    (3-D, stage q running q-1 slabs behind stage 1): identical results to the resident call, ragged
    last slab included."""
    import torch
    monkeypatch.setenv("PB200_HOST_PIPELINE", "8")
    nx = (40, 24, 45)                      # 45 planes: 5 slabs of 8, the last one 13 planes thick
    kw = dict(dimensions=3, nx=nx, gamma=1.4, reconstruction=recon, time_stepping=rk, bcs=bcs, ntracer=ntr,
              body_force=bf)
    h1, h2 = Hydro(**kw), Hydro(**kw)
    if bf:
        for comp, val in enumerate((0.0, -0.3, 0.1)):
            h1.set_body_force_vector(comp, np.full((1, 1, 1), val)); h2.set_body_force_vector(comp, np.full((1, 1, 1), val))
    v = _state_with_tracers((nx[2], nx[1], nx[0]), ntr, seed=9)
    h1.set_interior(v)
    pin = torch.empty(h2.shape, dtype=torch.float64, pin_memory=True)
    vc = pin.numpy()
    vc[:] = 1.0; vc[1:4] = 0.0; vc[h2.interior()] = v
    for n in range(3):
        i1 = h1.advance_step(2e-4)
        i2 = h2.advance_step_host(vc, 2e-4)
        assert abs(i1.invDt_hyp - i2.invDt_hyp) <= 1e-14 * i1.invDt_hyp and abs(i1.maxMach - i2.maxMach) <= 1e-14 * i1.maxMach
        assert i2.launches > i1.launches          # the slab-wise schedule really ran
    if recon == "LINEAR":
        assert np.array_equal(h1.get_interior(), vc[h2.interior()])
    else:   # PPM: a zone may sit at the other parity of the 2x-unrolled march (see test_slab_nccl_gpu.py)
        assert rel_err(vc[h2.interior()], h1.get_interior()) <= 1e-14
    h1.close(); h2.close()


@pytest.mark.parametrize("recon,rk", [("LINEAR", "RK2"), ("PARABOLIC", "RK3")])
def test_deep_halo_host_steps_equal_undecomposed(Hydro, monkeypatch, recon, rk):
    """Host-buffer steps of a slab-decomposed grid without any exchange inside the step (bench.py e2e at N > 1): two
    blocks, each with its own planes plus nghost x nstages planes of the neighbour on the cut face, go through the
    slab-pipelined pb200_advance_step_host(); only the own planes come back and count in invDt_hyp / maxMach
    (pb200_set_owned_planes); the host halos are refreshed from the neighbour's own planes between steps.  Same
    states, invDt_hyp and maxMach as the undecomposed grid."""
    import torch
    monkeypatch.setenv("PB200_HOST_PIPELINE", "6")
    nx, nz = (32, 24), 40
    bcs = ("reflective", "outflow") * 3
    kw = dict(dimensions=3, gamma=1.4, reconstruction=recon, time_stepping=rk, bcs=bcs)
    full = Hydro(nx=nx + (nz,), **kw)
    ng, E = full.nghost, full.nghost * full.nstages()
    v = random_state((nz, nx[1], nx[0]), seed=21, smooth=False)
    full.set_interior(v)
    half = nz // 2
    dz = 1.0 / nz
    blocks = []
    for r, (k0, k1, lo, hi) in enumerate([(0, half, 0, E), (half, nz, E, 0)]):
        nloc = (k1 - k0) + lo + hi
        h = Hydro(nx=nx + (nloc,), xbeg=(0., 0., (k0 - lo) * dz), xend=(1., 1., (k1 + hi) * dz), dx=(1. / nx[0], 1. / nx[1], dz), **kw)
        h.set_owned_planes(lo, lo + (k1 - k0))
        pin = torch.empty(h.shape, dtype=torch.float64, pin_memory=True)
        vc = pin.numpy()
        vc[:] = 1.0; vc[1:4] = 0.0
        vc[h.interior()] = v[:, k0 - lo:k1 + hi]
        blocks.append(dict(h=h, vc=vc, pin=pin, k0=k0, k1=k1, lo=lo, hi=hi))
    dt = 2e-4
    for n in range(3):
        i0 = full.advance_step(dt)
        infos = [b["h"].advance_step_host(b["vc"], dt) for b in blocks]
        assert abs(max(i.invDt_hyp for i in infos) - i0.invDt_hyp) <= 1e-14 * i0.invDt_hyp
        assert abs(max(i.maxMach for i in infos) - i0.maxMach) <= 1e-14 * i0.maxMach
        a, b = blocks
        own_a = a["vc"][:, ng:ng + half]                       # block 0 owns planes [0, half)
        own_b = b["vc"][:, ng + E:ng + E + (nz - half)]        # block 1 owns [half, nz)
        a["vc"][:, ng + half:ng + half + E] = own_b[:, :E]      # refresh the host halos from the neighbour's own planes
        b["vc"][:, ng:ng + E] = own_a[:, half - E:half]
        dt = min(0.3 / i0.invDt_hyp, 1.1 * dt)
    ref = full.get_interior()
    got = np.concatenate([blocks[0]["vc"][blocks[0]["h"].interior()][:, :half],
                          blocks[1]["vc"][blocks[1]["h"].interior()][:, E:]], axis=1)
    assert rel_err(got, ref) <= 1e-14
    full.close()
    for b in blocks:
        b["h"].close()


@pytest.mark.parametrize("recon,rk", [("LINEAR", "RK2"), ("PARABOLIC", "RK3")])
@pytest.mark.parametrize("bcs", [("reflective", "outflow", "outflow", "reflective", "periodic", "periodic"),
                                 ("periodic", "periodic", "reflective", "reflective", "outflow", "reflective"),
                                 ("outflow",) * 6])
def test_fused_boundaries_equal_boundary_kernels(Hydro, monkeypatch, recon, rk, bcs):
    """The fused x1+x2 kernel maps the x1 ghost zones at load time (outflow / mirror / periodic source
    zone, v_x1 flipped for mirrors) instead of reading ghosts materialised by bc_fill: same interior
    states, two launches fewer per stage."""
    nx = (37, 21, 26)
    kw = dict(dimensions=3, nx=nx, gamma=1.4, reconstruction=recon, time_stepping=rk, bcs=bcs, ntracer=1)
    v = _state_with_tracers((nx[2], nx[1], nx[0]), 1, seed=21)
    out, launches = [], []
    for fuse in ("1", "0"):
        monkeypatch.setenv("PB200_FUSE_BC", fuse)
        h = Hydro(**kw)
        h.set_interior(v)
        n = 0
        for _ in range(4):
            info = h.advance_step(3e-4)
            n += info.launches
        out.append(h.get_interior()); launches.append(n)
        h.close()
    assert launches[0] < launches[1]                 # the fused run skipped bc_fill launches
    assert np.array_equal(out[0], out[1])


def test_errors(Hydro):
    from pluto_sirocco_b200._lib import ENAN, PB200Error
    with pytest.raises(ValueError):
        Hydro(dimensions=1, nx=(32, 1, 1), solver="hlld")         # SetSolver (HD/set_solver.c): not an HD solver
    h = Hydro(dimensions=1, nx=(32, 1, 1))
    v = np.ones((5, 1, 1, 32)); v[1:4] = 0
    h.set_interior(v)
    with pytest.raises(PB200Error):
        h.advance_step(-1.0)
    v[0, 0, 0, 10] = np.nan
    h.set_interior(v)
    with pytest.raises(PB200Error) as ei:
        h.advance_step(1e-3)                                       # CheckNaN -> QUIT_PLUTO
    assert ei.value.code == ENAN
    h.close()


def test_ppm_on_stretched_grid_takes_the_general_path(Hydro):
    """The reference derives grid-dependent PPM weights on non-uniform grids (ppm_coeffs.c:124-136); the marching
    kernels only carry the uniform ones, so a stretched direction routes the context to the general path (which
    holds PPM_FindWeights' weights; parity: the ppmg_* fixtures of tests/test_gpu_gen.py).  Here: a constant state
    stays constant (the weights sum to one) and the same grid flagged uniform agrees with the marching kernels."""
    from pluto_sirocco_b200 import make_grid
    arrays = [make_grid((0.0, 32, 1.0, "r", 1.05), 3), make_grid((0.0, 1, 1.0), 0), make_grid((0.0, 1, 1.0), 0)]
    h = Hydro(dimensions=1, nx=(32, 1, 1), reconstruction="PARABOLIC", time_stepping="RK3", grid_arrays=arrays)
    v = np.ones((5, 1, 1, 32)); v[1:4] = 0
    h.set_interior(v)
    h.advance_step(1e-3)
    assert rel_err(h.get_interior(), v) <= 1e-14
    h.close()
    rng = np.random.default_rng(5)
    v = 1.0 + 0.1 * rng.random((5, 1, 1, 32)); v[2:4] = 0
    uni = [make_grid((0.0, 32, 1.0), 3), make_grid((0.0, 1, 1.0), 0), make_grid((0.0, 1, 1.0), 0)]
    out = []
    for flag in ((1, 1, 1), (0, 1, 1)):      # (0, ..): PPM_FindWeights on the uniform grid, general path
        h = Hydro(dimensions=1, nx=(32, 1, 1), reconstruction="PARABOLIC", time_stepping="RK3", grid_arrays=uni,
                  grid_uniform=flag, bcs=("periodic",) * 6)
        h.set_interior(v)
        for _ in range(3):
            h.advance_step(2e-3)
        out.append(h.get_interior())
        h.close()
    assert rel_err(out[1], out[0]) <= 1e-12


def test_uniform_state_is_a_fixed_point(Hydro):
    h = Hydro(dimensions=3, nx=(32, 16, 8), gamma=1.4, bcs=("periodic",) * 6)
    v = np.zeros((5, 8, 16, 32)); v[0] = 2.0; v[1] = 0.3; v[2] = -0.2; v[3] = 0.1; v[4] = 1.5
    h.set_interior(v)
    for n in range(5):
        h.advance_step(1e-2)
    assert rel_err(h.get_interior(), v) <= 1e-14
    h.close()


def test_conservation_periodic_large(Hydro):
    """Size-independent property at a large grid: mass, momentum and energy are conserved to
    round-off on a periodic box (flux form), 256^3 zones."""
    N = 256
    h = Hydro(dimensions=3, nx=(N, N, N), gamma=1.4, bcs=("periodic",) * 6)
    v = random_state((N, N, N), seed=1, smooth=False)
    h.set_interior(v)

    def totals(p):
        rho = p[0]; e = 0.5 * rho * (p[1] ** 2 + p[2] ** 2 + p[3] ** 2) + p[4] / 0.4
        return np.array([rho.sum(), (rho * p[1]).sum(), (rho * p[2]).sum(), (rho * p[3]).sum(), e.sum()])

    t0 = totals(v)
    dt = 1e-4
    for n in range(3):
        info = h.advance_step(dt)
    assert info.c2p_failures == 0
    t1 = totals(h.get_interior())
    scale = np.array([t0[0], np.abs(v[0]).sum(), np.abs(v[0]).sum(), np.abs(v[0]).sum(), t0[4]])
    assert np.all(np.abs(t1 - t0) / scale < 1e-12), (t1 - t0) / scale
    h.close()


def test_sedov_octant_symmetry_large(Hydro):
    """Sedov in the reflective octant is invariant under any permutation of the axes
    (x<->y<->z with the velocity components permuted), 192^3 zones, PLM+HLLC+RK2."""
    N = 192
    bcs = ("reflective", "outflow") * 3
    h = Hydro(dimensions=3, nx=(N, N, N), gamma=1.4, bcs=bcs)
    x = (np.arange(N) + 0.5) / N
    z3, y3, x3 = np.meshgrid(x, x, x, indexing="ij")
    r = np.sqrt(x3 ** 2 + y3 ** 2 + z3 ** 2)
    dr = 3.5 / N
    v = np.zeros((5, N, N, N)); v[0] = 1.0
    v[4] = np.where(r <= dr, 0.4 * 1.0 / (4.0 / 3.0 * np.pi * dr ** 3), 1e-5)
    h.set_interior(v)
    n, t, dt = h.integrate(12, t=0.0, dt=1e-9, tstop=0.5, cfl=0.3, cfl_max_var=1.1, first_dt=1e-9)
    assert n == 12
    p = h.get_interior()
    # swap x <-> z: arrays transpose (k,j,i)->(i,j,k) and vx1 <-> vx3
    q = p.transpose(0, 3, 2, 1)[[0, 3, 2, 1, 4]]
    assert rel_err(q, p) <= 1e-11
    # swap x <-> y
    q = p.transpose(0, 1, 3, 2)[[0, 2, 1, 3, 4]]
    assert rel_err(q, p) <= 1e-11
    h.close()


def test_against_reference_executable_when_present(Hydro, tmp_path):
    """If oracle/_ref travelled to this box, run the UNMODIFIED reference here and compare."""
    import refrun
    if not refrun.have_ref("sedov3d"):
        pytest.skip("oracle/_ref/sedov3d/pluto not present on this box")
    N = 32
    bcs = ("reflective", "outflow") * 3
    r = refrun.run("sedov3d", tmp_path, shape=(N, N, N), maxsteps=10,
                   grid=[(0, N, 1)] * 3, cfl=0.3, tstop=0.5, first_dt=1e-9, solver="hllc", bcs=bcs,
                   dbl=(-1.0, 1), params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4))
    h = Hydro(dimensions=3, nx=(N, N, N), gamma=1.4, bcs=bcs)
    for n in range(len(r["data"]) - 2):
        h.set_interior(r["data"][n])
        h.advance_step(r["steps"][n][2])
        assert rel_err(h.get_interior(), r["data"][n + 1]) <= TOL_STEP
    h.close()
