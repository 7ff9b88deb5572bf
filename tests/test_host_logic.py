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
