"""CPU: parsing of the unchanged user surface (definitions.h, pluto.ini)."""
from pluto_sirocco_b200.hydro import Definitions, Runtime

DEFS = """#define  PHYSICS                        HD
#define  DIMENSIONS                     3
#define  GEOMETRY                       CARTESIAN
#define  BODY_FORCE                     NO
#define  COOLING                        NO
#define  RECONSTRUCTION                 PARABOLIC
#define  TIME_STEPPING                  RK3
#define  NTRACER                        0
#define  PARTICLES                      NO
#define  USER_DEF_PARAMETERS            3

/* -- physics dependent declarations -- */

#define  EOS                            IDEAL
#define  ENTROPY_SWITCH                 NO

/* -- user-defined parameters (labels) -- */

#define  ENRG0                          0
#define  DNST0                          1
#define  GAMMA                          2

/* [Beg] user-defined constants (do not change this line) */

#define  INITIAL_SMOOTHING              YES

/* [End] user-defined constants (do not change this line) */
"""

INI = """[Grid]

X1-grid    1   0.0    48    u    1.0
X2-grid    1   0.0    32    u    2.0
X3-grid    1  -1.0    16    u    1.0

[Time]

CFL              0.3
CFL_max_var      1.1
tstop            0.5
first_dt         1.e-9

[Solver]

Solver         hll

[Boundary]

X1-beg        reflective
X1-end        outflow
X2-beg        periodic
X2-end        periodic
X3-beg        reflective
X3-end        outflow

[Parameters]

ENRG0                       1.0
DNST0                       1.0
GAMMA                       1.4
"""


def test_definitions_parse():
    d = Definitions.parse(DEFS)
    assert d.DIMENSIONS == 3 and d.RECONSTRUCTION == "PARABOLIC" and d.TIME_STEPPING == "RK3"
    assert d.user_params == ["ENRG0", "DNST0", "GAMMA"]
    assert d.extra["INITIAL_SMOOTHING"] == "YES"
    assert d.nghost() == 3
    d.check_supported()


def test_runtime_parse(tmp_path):
    p = tmp_path / "pluto.ini"
    p.write_text(INI)
    rt = Runtime.parse(p)
    assert rt.npoint == [48, 32, 16] and rt.xbeg == [0.0, 0.0, -1.0] and rt.xend == [1.0, 2.0, 1.0]
    assert rt.cfl == 0.3 and rt.first_dt == 1e-9 and rt.solver == "hll"
    assert rt.left_bound == ["reflective", "periodic", "reflective"]
    assert rt.right_bound == ["outflow", "periodic", "outflow"]
    assert rt.params["GAMMA"] == 1.4


def test_make_grid_matches_the_grid_restated_for_the_oracle_and_grid_out(tmp_path):
    """pluto_sirocco_b200.make_grid (what a Python caller hands to pb200_set_grid) and
    oracle/pluto_grid.make_grid (the restatement the oracle was pinned with) are independent
    transcriptions of Src/set_grid.c:395-450,100-127 and must agree bit for bit; both agree with
    the 12 digits the reference prints into grid.out when its executable is available."""
    import numpy as np
    import pluto_grid
    from pluto_sirocco_b200 import make_grid
    specs = [(0.0, 37, 1.0), (-0.5, 16, 0.5, "u"), (0.87, 48, 8.7, "r", 1.05), (0.0, 36, 1.5707963267948966, "r", 0.95)]
    for spec in specs:
        for ng in (0, 2, 3):
            a = make_grid(spec, ng)
            b = pluto_grid.make_grid(spec, ng)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
            xl, xr, dx = a
            assert np.all(xr > xl) and np.allclose(xr - xl, dx, rtol=1e-13)
            assert abs(xl[ng] - spec[0]) < 1e-15 and abs(xr[len(xl) - ng - 1] - spec[2]) < 1e-12 * max(1.0, abs(spec[2]))
    import refrun
    if refrun.have_ref("sph2d"):
        grid = [specs[2], specs[3], (0.0, 1, 1.0)]
        refrun.run("sph2d", tmp_path, shape=(1, 36, 48), nvar=6, maxsteps=0, timeout=60,
                   grid=[pluto_grid.ini_string(g) for g in grid], cfl=0.4, tstop=1.0, first_dt=1e-5, solver="hll",
                   bcs=("outflow",) * 6, dbl=(-1.0, 1), params=dict(GM=1.0, RBLOB=2.0, TBLOB=1.0, PBLOB=5.0))
        rows = [l.split() for l in (tmp_path / "grid.out").read_text().splitlines() if l and not l.startswith("#")]
        n1 = int(rows[0][0])
        ref1 = np.array([[float(r[1]), float(r[2])] for r in rows[1:1 + n1]])
        xl, xr, _ = make_grid(specs[2], 0)
        assert np.allclose(ref1[:, 0], xl, rtol=1e-11, atol=1e-12) and np.allclose(ref1[:, 1], xr, rtol=1e-11, atol=1e-12)


import numpy as np
import pytest

# fixture -> reference build configuration of oracle/build_ref.py whose definitions.h it was generated with
FILES_CASES = {"sph2d_flat_hllc": "sph2d_flat", "pol3d_hll": "pol3d", "iso2d_flat_hllc": "iso2d_flat", "oned_sph2d_char_roe": "sph2d_char_oned",
               "ppmg_kh3d_stretched": "kh3d_ppm", "ring_pol2d_vl": "pol2d_ring_vl", "ring_sph3d_mp5": "sph3d_ring",
               "pot_sph3d_both_hll": "sph3d_pot", "sph2d_sel_hllc": "sph2d_sel"}


@pytest.mark.parametrize("name", list(FILES_CASES))
def test_kwargs_from_definitions_and_pluto_ini(name, tmp_path):
    """Hydro.kwargs_from_files(): the definitions.h the reference executable of a fixture was compiled with and the
    pluto.ini it ran on give the same library configuration (options, ghost zones, grid arrays, grid->uniform) as the
    fixture's own record - the host-side mirror of what csrc/pluto_shim.c reads from the macros, Runtime and Grid."""
    import build_ref
    import pluto_grid
    import refrun
    from common import gen_kwargs_from_golden, hydro_kwargs_from_gen, load_golden
    from pluto_sirocco_b200.hydro import Hydro
    g = load_golden(name)
    cfg = build_ref.CONFIGS[FILES_CASES[name]]
    defs_text = build_ref.patch_definitions((build_ref.HERE / "problems" / cfg["local"] / "definitions.h").read_text(),
                                            cfg["overrides"])
    d = Definitions.parse(defs_text)
    want = hydro_kwargs_from_gen(gen_kwargs_from_golden(g))
    grid = []
    for row in g["gridspec"]:
        grid.append((float(row[0]), int(row[1]), float(row[2]), "r", float(row[4])) if row[3] == 1.0
                    else (float(row[0]), int(row[1]), float(row[2])))
    refrun.write_ini(tmp_path / "pluto.ini", grid=[pluto_grid.ini_string(s) for s in grid], cfl=float(g["cfl"]),
                     tstop=float(g["tstop"]), first_dt=float(g["first_dt"]), solver=str(g["solver"]), bcs=tuple(str(b) for b in g["bcs"]))
    rt = Runtime.parse(tmp_path / "pluto.ini")
    got = Hydro.kwargs_from_files(d, rt, gamma=float(g["gamma"]), iso_sound_speed=float(g["iso_cs"]) if "iso_cs" in g else 0.0)
    nd = int(g["dims"])
    assert got["dimensions"] == nd and got["nx"][:nd] == tuple(want["nx"])[:nd] and got["nghost"] == want["nghost"]
    for key in ("geometry", "reconstruction", "time_stepping", "solver", "limiter", "ntracer", "body_force"):
        assert got[key] == want[key], (key, got[key], want[key])
    assert bool(got["char_limiting"]) == bool(want["char_limiting"])
    assert got["shock_flattening"] == want["shock_flattening"]
    assert (got["entropy_switch"] or False) == (want["entropy_switch"] or False)
    assert got["eos"] == want.get("eos", "IDEAL")
    assert got["ring_average"] == want.get("ring_average", 0)
    if got["ring_average"]:
        assert got["ring_average_rec"] == want["ring_average_rec"]
    assert tuple(got["bcs"])[:2 * nd] == tuple(str(b) for b in g["bcs"])[:2 * nd]
    for dd in range(nd):
        for a, b in zip(got["grid_arrays"][dd], want["grid_arrays"][dd]):
            assert np.array_equal(a, b), (name, dd)
        assert got["grid_uniform"][dd] == want["grid_uniform"][dd]
