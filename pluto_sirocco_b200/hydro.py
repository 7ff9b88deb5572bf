"""Host-side mirror of the reference interface for the hot path.

Names follow the reference: `Runtime` (Src/structs.h:434, filled from pluto.ini by
Src/runtime_setup.c:67-507), `Definitions` (the problem's definitions.h), `Grid`
(Src/structs.h:133), `Data.Vc` (Src/structs.h:525), `AdvanceStep` (rk_step.c:29),
`NextTimeStep` / `Integrate` / the main loop (Src/main.c:215-337,383,521).

Everything numerical happens in libplutob200.so (CUDA, sm_100a) through the C ABI of
include/pluto_b200.h; this module only parses the reference's input files, owns the host
copy of Vc and sequences calls.  It has no CPU implementation of the update.
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import _lib as L

NVAR_HD = 5
RHO, VX1, VX2, VX3, PRS = 0, 1, 2, 3, 4

_SOLVERS = {"tvdlf": L.TVDLF, "hll": L.HLL, "hllc": L.HLLC, "roe": 4, "two_shock": 5, "ausm+": 6}
_RECON = {"FLAT": L.FLAT, "LINEAR": L.LINEAR, "PARABOLIC": L.PARABOLIC}
_TSTEP = {"EULER": L.EULER, "RK2": L.RK2, "RK3": L.RK3}


# --------------------------------------------------------------------------------------
#  definitions.h / pluto.ini  (the unchanged user surface)
# --------------------------------------------------------------------------------------
@dataclass
class Definitions:
    """Compile-time options of a problem (definitions.h)."""
    PHYSICS: str = "HD"
    DIMENSIONS: int = 1
    GEOMETRY: str = "CARTESIAN"
    BODY_FORCE: str = "NO"
    COOLING: str = "NO"
    RECONSTRUCTION: str = "LINEAR"
    TIME_STEPPING: str = "RK2"
    NTRACER: int = 0
    EOS: str = "IDEAL"
    ENTROPY_SWITCH: str = "NO"
    LIMITER: str = "DEFAULT"
    CHAR_LIMITING: str = "NO"
    SHOCK_FLATTENING: str = "NO"
    RING_AVERAGE: str = "NO"          # chunk size at the axis (an integer > 1) or NO, Src/pluto.h:479
    RING_AVERAGE_REC: str = ""        # 1 / 2 / 5; empty: the reference's default (5 when RING_AVERAGE is on)
    user_params: list = field(default_factory=list)   # labels, in index order
    extra: dict = field(default_factory=dict)

    @classmethod
    def parse(cls, path_or_text) -> "Definitions":
        text = Path(path_or_text).read_text() if "\n" not in str(path_or_text) else str(path_or_text)
        d = cls()
        in_user = False
        for line in text.splitlines():
            if "user-defined parameters (labels)" in line:
                in_user = True
                continue
            if "[Beg] user-defined constants" in line:
                in_user = False
            m = re.match(r"\s*#define\s+(\w+)\s+(\S+)", line)
            if not m:
                continue
            k, v = m.group(1), m.group(2)
            if in_user:
                d.user_params.append(k)
            elif k in ("DIMENSIONS", "NTRACER"):
                setattr(d, k, int(v))
            elif hasattr(d, k) and k not in ("user_params", "extra"):
                setattr(d, k, v)
            else:
                d.extra[k] = v
        return d

    def ring_average(self):
        """(RING_AVERAGE, RING_AVERAGE_REC) as integers; (0, 1) when off (Src/pluto.h:479-489)."""
        ra = 0 if self.RING_AVERAGE in ("NO", "0", "1") else int(self.RING_AVERAGE)
        rec = int(self.RING_AVERAGE_REC) if self.RING_AVERAGE_REC else (5 if ra > 1 else 1)
        return ra, rec

    def nghost(self) -> int:
        """GetNghost(), Src/get_nghost.c:19-60."""
        n = 2
        if self.RECONSTRUCTION == "PARABOLIC" or self.ring_average()[1] > 2:
            n = 3
        if self.SHOCK_FLATTENING == "ONED":
            n = max(4, n)
        elif self.SHOCK_FLATTENING == "MULTID":
            n = max(3, n)
        return n

    def body_force_bits(self) -> int:
        """BODY_FORCE as the bit mask of Src/pluto.h:76-77 (VECTOR 1, POTENTIAL 2)."""
        b = self.BODY_FORCE.upper()
        return (1 if "VECTOR" in b else 0) | (2 if "POTENTIAL" in b else 0)

    def check_supported(self):
        if self.PHYSICS != "HD":
            raise NotImplementedError("PHYSICS %s: only HD is on the hot path" % self.PHYSICS)
        if self.EOS not in ("IDEAL", "ISOTHERMAL"):
            raise NotImplementedError("EOS %s" % self.EOS)
        if self.GEOMETRY not in ("CARTESIAN", "CYLINDRICAL", "POLAR", "SPHERICAL"):
            raise NotImplementedError("GEOMETRY %s" % self.GEOMETRY)
        if self.RECONSTRUCTION not in _RECON:
            raise NotImplementedError("RECONSTRUCTION %s" % self.RECONSTRUCTION)
        if self.TIME_STEPPING not in _TSTEP:
            raise NotImplementedError("TIME_STEPPING %s" % self.TIME_STEPPING)
        if self.SHOCK_FLATTENING not in ("NO", "MULTID", "ONED"):
            raise NotImplementedError("SHOCK_FLATTENING %s" % self.SHOCK_FLATTENING)
        if self.ENTROPY_SWITCH not in ("NO", "SELECTIVE", "ALWAYS"):
            raise NotImplementedError("ENTROPY_SWITCH %s" % self.ENTROPY_SWITCH)
        if self.COOLING != "NO":
            raise NotImplementedError("COOLING %s: the source step goes through the drop-in shim (csrc/pluto_shim.c)" % self.COOLING)


@dataclass
class Runtime:
    """Run-time options (pluto.ini), Src/structs.h:434 / Src/runtime_setup.c."""
    npoint: list = field(default_factory=lambda: [1, 1, 1])
    xbeg: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    xend: list = field(default_factory=lambda: [1.0, 1.0, 1.0])
    cfl: float = 0.4
    cfl_max_var: float = 1.1
    tstop: float = 1.0
    first_dt: float = 1.e-4
    solver: str = "hllc"
    left_bound: list = field(default_factory=lambda: ["outflow"] * 3)
    right_bound: list = field(default_factory=lambda: ["outflow"] * 3)
    params: dict = field(default_factory=dict)
    grid: list = field(default_factory=lambda: [None, None, None])   # make_grid() specs of the [Grid] lines

    @classmethod
    def parse(cls, path) -> "Runtime":
        rt = cls()
        section = None
        for raw in Path(path).read_text().splitlines():
            line = raw.split("#")[0].strip()
            if not line:
                continue
            m = re.match(r"\[(.+)\]", line)
            if m:
                section = m.group(1).strip()
                continue
            tok = line.split()
            key = tok[0]
            if section == "Grid" and re.match(r"X[123]-grid", key):
                d = int(key[1]) - 1
                npatch = int(tok[1])
                if npatch != 1 or tok[4] not in ("u", "uniform", "r"):
                    raise NotImplementedError("only single 'u' (uniform) or 'r' (ratio, this fork) patches are parsed here")
                rt.xbeg[d] = float(tok[2])
                rt.npoint[d] = int(tok[3])
                if tok[4] == "r":       # X1-grid 1 xL n r xR ratio (set_grid.c, the fork's ratio grid)
                    rt.xend[d] = float(tok[5])
                    rt.grid[d] = (rt.xbeg[d], rt.npoint[d], rt.xend[d], "r", float(tok[6]))
                else:
                    rt.xend[d] = float(tok[5])
                    rt.grid[d] = (rt.xbeg[d], rt.npoint[d], rt.xend[d])
            elif section == "Time":
                if key == "CFL": rt.cfl = float(tok[1])
                elif key == "CFL_max_var": rt.cfl_max_var = float(tok[1])
                elif key == "tstop": rt.tstop = float(tok[1])
                elif key == "first_dt": rt.first_dt = float(tok[1])
            elif section == "Solver" and key == "Solver":
                rt.solver = tok[1]
            elif section == "Boundary":
                m = re.match(r"X([123])-(beg|end)", key)
                if m:
                    d = int(m.group(1)) - 1
                    (rt.left_bound if m.group(2) == "beg" else rt.right_bound)[d] = tok[1]
            elif section == "Parameters":
                rt.params[key] = float(tok[1])
        return rt


def make_grid(spec, nghost):
    """One direction of the reference's grid for a single-patch [Grid] line of pluto.ini:
    spec = (xL, n, xR) or (xL, n, xR, 'u') uniform, (xL, n, xR, 'r', ratio) the fork's ratio grid.
    Mirrors MakeGrid() (Src/set_grid.c:395-450) and the ghost extension (Src/set_grid.c:100-127) so
    that a Python caller can hand pb200_set_grid() the same xl / xr / dx the reference's Grid holds."""
    import math
    xL, n, xR = float(spec[0]), int(spec[1]), float(spec[2])
    kind = spec[3] if len(spec) > 3 else "u"
    b = nghost
    xl = np.zeros(n + 2 * b); xr = np.zeros(n + 2 * b); dx = np.zeros(n + 2 * b)
    if kind == "u":
        for i in range(n):
            dx[b + i] = (xR - xL) / float(n)
            xl[b + i] = xL + float(i) * dx[b + i]
            xr[b + i] = xl[b + i] + dx[b + i]
    elif kind == "r":
        ratio = float(spec[4])
        xl[b] = xL
        dx[b] = (xR - xL) * (ratio - 1.0) / (math.pow(ratio, n) - 1.0)
        xr[b] = xl[b] + dx[b]
        for i in range(1, n):
            dx[b + i] = dx[b + i - 1] * ratio
            xl[b + i] = xl[b + i - 1] + dx[b + i - 1]
            xr[b + i] = xl[b + i] + dx[b + i]
    else:
        raise NotImplementedError("grid type %r" % kind)
    e = b + n - 1
    for i in range(b):
        dx[i] = dx[b]; xl[i] = xl[b] - (b - i) * dx[b]; xr[i] = xl[i] + dx[b]
        dx[e + i + 1] = dx[e]; xl[e + i + 1] = xl[e] + (i + 1) * dx[e]; xr[e + i + 1] = xl[e] + (i + 2) * dx[e]
    return xl, xr, dx


# --------------------------------------------------------------------------------------
#  Hydro: one block of the grid resident on one B200
# --------------------------------------------------------------------------------------
class MultiHydro:
    """Front end of pb200_multi_* (csrc/pb200_multi.cu): one host thread, N GPUs, slab split along the
    outermost active direction, NCCL exchange inside the library.  Arrays are GLOBAL d->Vc arrays.
    devices may repeat an ordinal (ranks sharing a GPU: the slab logic on a single-GPU box)."""

    def __init__(self, ngpus, *, dimensions, nx, devices=None, xbeg=(0., 0., 0.), xend=(1., 1., 1.), gamma=5. / 3.,
                 reconstruction="LINEAR", time_stepping="RK2", solver="hllc", limiter="DEFAULT",
                 bcs=("outflow",) * 6, ntracer=0, nghost=None, small_density=1e-12, small_pressure=1e-12,
                 body_force=0):
        lib = L.load()
        cfg = L.Config()
        lib.pb200_config_default(C.byref(cfg))
        cfg.dimensions = dimensions
        for d in range(3):
            cfg.nx[d] = int(nx[d]) if d < dimensions else 1
            cfg.xbeg[d] = float(xbeg[d])
            cfg.xend[d] = float(xend[d])
        cfg.reconstruction = _RECON[reconstruction]
        cfg.nghost = nghost if nghost is not None else (3 if reconstruction == "PARABOLIC" else 2)
        cfg.ntracer = ntracer
        cfg.limiter = L.LIMITERS[limiter]
        cfg.time_stepping = _TSTEP[time_stepping]
        cfg.solver = _SOLVERS[solver]
        for s in range(6):
            cfg.bc[s] = L.BC[bcs[s]] if isinstance(bcs[s], str) else int(bcs[s])
        cfg.gamma, cfg.small_density, cfg.small_pressure = gamma, small_density, small_pressure
        cfg.body_force = int(body_force)
        self.cfg, self._lib, self.ngpus = cfg, lib, int(ngpus)
        devs = None if devices is None else (C.c_int * self.ngpus)(*devices)
        h = C.c_void_p()
        L.check(lib.pb200_multi_create(C.byref(cfg), self.ngpus, devs, C.byref(h)))
        self._h = h
        ng = cfg.nghost
        self.nghost, self.dimensions = ng, dimensions
        self.tot = tuple(cfg.nx[d] + 2 * ng if d < dimensions else 1 for d in range(3))
        self.nvar = 5 + ntracer
        self.shape = (self.nvar, self.tot[2], self.tot[1], self.tot[0])
        self.beg = tuple(ng if d < dimensions else 0 for d in range(3))
        self.nx = tuple(cfg.nx[d] for d in range(3))
        self.last = L.StepInfo()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    interior = lambda self: Hydro.interior(self)
    _bf_table = lambda self, tab: Hydro._bf_table(self, tab)

    def set_body_force_vector(self, comp, tab):
        a, si, sj, sk = self._bf_table(tab)
        L.check(self._lib.pb200_multi_set_body_force_vector(self._h, int(comp), a.ctypes.data_as(C.c_void_p), a.size, si, sj, sk))

    def set_body_force_potential(self, where, tab):
        a, si, sj, sk = self._bf_table(tab)
        L.check(self._lib.pb200_multi_set_body_force_potential(self._h, int(where), a.ctypes.data_as(C.c_void_p), a.size, si, sj, sk))

    def upload(self, vc):
        vc = np.ascontiguousarray(vc, dtype=np.float64)
        assert vc.shape == self.shape
        L.check(self._lib.pb200_multi_upload_vc(self._h, vc.ctypes.data_as(C.c_void_p)))

    def download(self, out=None):
        if out is None:
            out = np.zeros(self.shape, dtype=np.float64)
        L.check(self._lib.pb200_multi_download_vc(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_interior(self, v_int):
        vc = np.ones(self.shape)
        vc[1:4] = 0.0
        vc[self.interior()] = v_int
        self.upload(vc)

    def get_interior(self):
        return self.download()[self.interior()].copy()

    def advance_step(self, dt):
        L.check(self._lib.pb200_multi_advance_step(self._h, float(dt), C.byref(self.last)))
        return self.last

    def advance_step_host(self, vc, dt):
        assert vc.flags["C_CONTIGUOUS"] and vc.shape == self.shape and vc.dtype == np.float64
        L.check(self._lib.pb200_multi_advance_step_host(self._h, vc.ctypes.data_as(C.c_void_p), float(dt), C.byref(self.last)))
        return self.last

    def integrate(self, nsteps, *, t, dt, tstop, cfl, cfl_max_var, first_dt):
        tt, dd = C.c_double(t), C.c_double(dt)
        n = L.check(self._lib.pb200_multi_integrate(self._h, int(nsteps), float(tstop), float(cfl), float(cfl_max_var),
                                                    float(first_dt), C.byref(tt), C.byref(dd), C.byref(self.last)))
        return n, tt.value, dd.value


class Hydro:
    """Device-resident d->Vc plus the AdvanceStep family of calls."""

    def __init__(self, *, dimensions, nx, xbeg=(0., 0., 0.), xend=(1., 1., 1.), gamma=5. / 3.,
                 reconstruction="LINEAR", time_stepping="RK2", solver="hllc", limiter="DEFAULT",
                 bcs=("outflow",) * 6, ntracer=0, nghost=None, device=0,
                 small_density=1e-12, small_pressure=1e-12, dx=None, body_force=0,
                 geometry="CARTESIAN", grid_arrays=None, grid_uniform=None, char_limiting=False, shock_flattening=False,
                 entropy_switch=False, eos="IDEAL", iso_sound_speed=0.0, ring_average=0, ring_average_rec=None):
        lib = L.load()
        cfg = L.Config()
        lib.pb200_config_default(C.byref(cfg))
        cfg.dimensions = dimensions
        for d in range(3):
            cfg.nx[d] = int(nx[d]) if d < dimensions else 1
            cfg.xbeg[d] = float(xbeg[d])
            cfg.xend[d] = float(xend[d])
        cfg.reconstruction = _RECON[reconstruction]
        cfg.nghost = nghost if nghost is not None else (3 if reconstruction == "PARABOLIC" else 2)
        cfg.ntracer = ntracer
        cfg.limiter = L.LIMITERS[limiter]
        cfg.time_stepping = _TSTEP[time_stepping]
        if solver not in _SOLVERS:
            # SetSolver(): "is not available" -> QUIT_PLUTO (Src/HD/set_solver.c:52-56)
            raise ValueError("! SetSolver: '%s' is not available on the B200 path" % solver)
        cfg.solver = _SOLVERS[solver]
        for s in range(6):
            b = bcs[s]
            cfg.bc[s] = L.BC[b] if isinstance(b, str) else int(b)
        cfg.gamma = gamma
        cfg.small_density = small_density
        cfg.small_pressure = small_pressure
        cfg.device = device
        cfg.body_force = int(body_force)
        cfg.geometry = {"CARTESIAN": 1, "CYLINDRICAL": 2, "POLAR": 3, "SPHERICAL": 4}[geometry]
        cfg.char_limiting = int(bool(char_limiting))
        cfg.shock_flattening = {False: 0, None: 0, True: 1, "NO": 0, "MULTID": 1, "ONED": 2}[shock_flattening]
        cfg.entropy_switch = {False: 0, None: 0, True: 2, "NO": 0, "SELECTIVE": 1, "ALWAYS": 2}[entropy_switch]
        cfg.eos = {"IDEAL": 0, "ISOTHERMAL": 1}[eos]
        cfg.iso_sound_speed = float(iso_sound_speed)
        cfg.ring_average = int(ring_average)                    # RING_AVERAGE, RING_AVERAGE_REC (pluto.h:479-489)
        cfg.ring_average_rec = int(ring_average_rec or 0)
        self.cfg = cfg
        self._lib = lib
        h = C.c_void_p()
        L.check(lib.pb200_create(C.byref(cfg), C.byref(h)))
        self._h = h
        tot = (C.c_int * 3)()
        nv = C.c_int()
        L.check(lib.pb200_shape(h, C.byref(tot), C.byref(nv)))
        self.tot = tuple(tot)                      # NX1_TOT, NX2_TOT, NX3_TOT
        self.nvar = nv.value
        self.nghost = cfg.nghost
        self.dimensions = dimensions
        self.shape = (self.nvar, tot[2], tot[1], tot[0])   # Vc[nv][k][j][i]
        self.beg = tuple(cfg.nghost if d < dimensions else 0 for d in range(3))
        self.nx = tuple(cfg.nx[d] for d in range(3))
        self.xbeg = tuple(cfg.xbeg)
        self.xend = tuple(cfg.xend)
        self.last = L.StepInfo()
        self._grid = None
        if grid_arrays is not None:   # grid->xl, xr, dx of every active direction (np_tot entries)
            self._grid = []
            for d in range(3):
                xl, xr, dxa = (np.ascontiguousarray(a, dtype=np.float64) for a in grid_arrays[d])
                assert xl.size == self.tot[d], (d, xl.size, self.tot[d])
                self._grid.append((xl, xr, dxa))
                L.check(lib.pb200_set_grid(h, d, xl.ctypes.data_as(C.c_void_p), xr.ctypes.data_as(C.c_void_p),
                                           dxa.ctypes.data_as(C.c_void_p)))
        if grid_uniform is not None:   # grid->uniform[d]: one uniform patch in pluto.ini (PPM weights, ppm_coeffs.c:88-136)
            u = (C.c_int * 3)(*[int(bool(grid_uniform[d])) if d < len(grid_uniform) else 1 for d in range(3)])
            L.check(lib.pb200_set_grid_uniform(h, u))
        if dx is not None:      # block of a larger uniform grid: impose the global grid->dx
            for d in range(dimensions):
                n = self.tot[d]
                xl = cfg.xbeg[d] + (np.arange(n) - self.beg[d]) * dx[d]
                xr = xl + dx[d]
                dxa = np.full(n, dx[d])
                L.check(lib.pb200_set_grid(h, d, xl.ctypes.data_as(C.c_void_p),
                                           xr.ctypes.data_as(C.c_void_p), dxa.ctypes.data_as(C.c_void_p)))

    @staticmethod
    def kwargs_from_files(definitions: Definitions, runtime: Runtime, gamma=5. / 3., iso_sound_speed=0.0, device=0):
        """Constructor keywords for the problem a definitions.h + pluto.ini pair describes: the compile-time options
        the shim reads as macros (csrc/pluto_shim.c: shim_fill_config) and the run-time ones it reads from Runtime /
        Grid.  g_gamma and g_isoSoundSpeed are set by the user's Init() in the reference: pass them in."""
        definitions.check_supported()
        nd = definitions.DIMENSIONS
        bcs = []
        for d in range(3):
            bcs += [runtime.left_bound[d], runtime.right_bound[d]]
        ng = definitions.nghost()
        ra, rec = definitions.ring_average()
        kw = dict(dimensions=nd, nx=tuple(int(runtime.npoint[d]) if d < nd else 1 for d in range(3)),
                  xbeg=tuple(runtime.xbeg), xend=tuple(runtime.xend), gamma=gamma,
                  reconstruction=definitions.RECONSTRUCTION, time_stepping=definitions.TIME_STEPPING,
                  solver=runtime.solver, limiter=definitions.LIMITER, bcs=tuple(bcs), ntracer=definitions.NTRACER,
                  nghost=ng, device=device, geometry=definitions.GEOMETRY, body_force=definitions.body_force_bits(),
                  char_limiting=definitions.CHAR_LIMITING == "YES",
                  shock_flattening={"NO": False, "MULTID": True, "ONED": "ONED"}[definitions.SHOCK_FLATTENING],
                  entropy_switch={"NO": False, "SELECTIVE": "SELECTIVE", "ALWAYS": "ALWAYS"}[definitions.ENTROPY_SWITCH],
                  eos=definitions.EOS, iso_sound_speed=iso_sound_speed, ring_average=ra, ring_average_rec=rec)
        specs = [runtime.grid[d] if runtime.grid[d] is not None else (runtime.xbeg[d], runtime.npoint[d], runtime.xend[d])
                 for d in range(3)]
        kw["grid_arrays"] = [make_grid(specs[d], ng if d < nd else 0) if d < nd else make_grid((specs[d][0], 1, specs[d][2]), 0)
                             for d in range(3)]
        kw["grid_uniform"] = tuple(len(specs[d]) <= 3 for d in range(3))      # grid->uniform[d], set_grid.c:67-72
        return kw

    @classmethod
    def from_files(cls, definitions: Definitions, runtime: Runtime, gamma=5. / 3., iso_sound_speed=0.0, device=0):
        return cls(**cls.kwargs_from_files(definitions, runtime, gamma=gamma, iso_sound_speed=iso_sound_speed, device=device))

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry helpers (host) -----------------------------------------------------------
    def interior(self):
        """slices selecting the interior (DOM) zones of Vc[nv][k][j][i]."""
        sl = [slice(None)]
        for d in (2, 1, 0):
            b = self.beg[d]
            sl.append(slice(b, b + self.nx[d]))
        return tuple(sl)

    def x(self, d):
        """grid->x[d] (cell centres incl. ghosts)."""
        if self._grid is not None:
            return 0.5 * (self._grid[d][0] + self._grid[d][1])
        return self.cell_centers(d)

    @property
    def xr(self):
        """grid->xr[d] (upper zone faces incl. ghosts) of the three directions."""
        if self._grid is not None:
            return [self._grid[d][1] for d in range(3)]
        out = []
        for d in range(3):
            dx = (self.cfg.xend[d] - self.cfg.xbeg[d]) / self.nx[d]
            out.append(self.cfg.xbeg[d] + (np.arange(self.tot[d]) - self.beg[d] + 1) * dx)
        return out

    def cell_centers(self, d):
        n = self.tot[d]
        dx = (self.cfg.xend[d] - self.cfg.xbeg[d]) / self.nx[d]
        return self.cfg.xbeg[d] + (np.arange(n) - self.beg[d] + 0.5) * dx

    def new_vc(self):
        return np.zeros(self.shape, dtype=np.float64)

    # -- BODY_FORCE tables ---------------------------------------------------------------------
    def _bf_table(self, tab):
        """tab: array broadcastable to [NX3_TOT][NX2_TOT][NX1_TOT]; axes of length 1 get stride 0."""
        a = np.asarray(tab, dtype=np.float64)
        while a.ndim < 3:
            a = a[None]
        full = (self.tot[2], self.tot[1], self.tot[0])
        for ax in range(3):
            if a.shape[ax] not in (1, full[ax]):
                raise ValueError("body-force table axis %d has length %d, expected 1 or %d" % (ax, a.shape[ax], full[ax]))
        a = np.ascontiguousarray(a)
        st = [a.strides[ax] // 8 if a.shape[ax] > 1 else 0 for ax in range(3)]   # k, j, i
        return a, st[2], st[1], st[0]

    def set_body_force_vector(self, comp, tab):
        """g[comp] of BodyForceVector at zone centres (rhs_source.c:256)."""
        a, si, sj, sk = self._bf_table(tab)
        L.check(self._lib.pb200_set_body_force_vector(self._h, int(comp), a.ctypes.data_as(C.c_void_p),
                                                      a.size, si, sj, sk))

    def set_body_force_potential(self, where, tab):
        """BodyForcePotential at zone centres (where=0) or at the upper x1/x2/x3 faces (1,2,3)."""
        a, si, sj, sk = self._bf_table(tab)
        L.check(self._lib.pb200_set_body_force_potential(self._h, int(where), a.ctypes.data_as(C.c_void_p),
                                                         a.size, si, sj, sk))

    # -- line-driven wind ----------------------------------------------------------------------
    def set_ldw(self, *, params, units, flux_r, flux_t, flux_p=None, userdef_bc=True, t_fit=None, m_fit=None):
        """LINE_DRIVEN_WIND SIROCCO_MODE: g_inputParam[] of the cv_idl problem (dict by label), UNIT_*
        (dict), directional fluxes [nangles][k][j][i] incl. ghosts (read_sirocco_fluxes())."""
        lc = L.LdwConfig()
        fr = np.ascontiguousarray(flux_r, dtype=np.float64)
        ft = np.ascontiguousarray(flux_t, dtype=np.float64)
        assert fr.shape[1:] == self.shape[1:] and ft.shape == fr.shape
        lc.nangles = fr.shape[0]
        lc.userdef_bc = int(userdef_bc)
        lc.unit_length, lc.unit_velocity, lc.unit_density = units["length"], units["velocity"], units["density"]
        lc.mu, lc.krad, lc.alpharad = params["MU"], params["KRAD"], params["ALPHARAD"]
        lc.dfloor, lc.rho_0, lc.rho_alpha = params["DFLOOR"], params["RHO_0"], params["RHO_ALPHA"]
        lc.cent_mass, lc.disk_mdot = params["CENT_MASS"], params["DISK_MDOT"]
        lc.lx, lc.tx = params["L_star"] * params["f_x"], params["T_x"]
        lc.t_iso = params.get("T_ISO", 0.0)
        L.check(self._lib.pb200_ldw_enable(self._h, C.byref(lc)))
        fp = None if flux_p is None else np.ascontiguousarray(flux_p, dtype=np.float64)
        L.check(self._lib.pb200_ldw_set_fluxes(self._h, fr.ctypes.data_as(C.c_void_p), ft.ctypes.data_as(C.c_void_p),
                                               None if fp is None else fp.ctypes.data_as(C.c_void_p)))
        if t_fit is not None:     # KRAD = ALPHARAD = 999: log10(t) [MPOINTS], log10(M) [MPOINTS][k][j][i]
            tf = np.ascontiguousarray(t_fit, dtype=np.float64)
            mf = np.ascontiguousarray(m_fit, dtype=np.float64)
            assert mf.shape == (tf.size,) + self.shape[1:]
            L.check(self._lib.pb200_ldw_set_mfit(self._h, tf.size, tf.ctypes.data_as(C.c_void_p),
                                                 mf.ctypes.data_as(C.c_void_p)))

    def set_cooling_tables(self, tabs):
        """Data->comp_h_pre, comp_c_pre, xray_h_pre, line_c_pre, brem_c_pre, sirocco_xi, sirocco_t_r
        (None entries keep the reference's defaults)."""
        keep = [None if t is None else np.ascontiguousarray(np.broadcast_to(t, self.shape[1:]), dtype=np.float64)
                for t in tabs]
        ptrs = (C.c_void_p * 7)(*[None if a is None else a.ctypes.data for a in keep])
        L.check(self._lib.pb200_cooling_set_tables(self._h, C.byref(ptrs)))

    def set_internal_boundary_mask(self, mask):
        """FLAG_INTERNAL_BOUNDARY zones (bool/uint8 [NX3_TOT][NX2_TOT][NX1_TOT], None: none): rhs = 0 in every
        sweep, InternalBoundaryReset() (Src/int_bound_reset.c:17)."""
        if mask is None:
            L.check(self._lib.pb200_set_internal_boundary_mask(self._h, None))
            return
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert m.shape == self.shape[1:]
        L.check(self._lib.pb200_set_internal_boundary_mask(self._h, m.ctypes.data_as(C.c_void_p)))

    def split_source(self, dt, g_time):
        """SplitSource(d, dt, Dts, grid) for COOLING BLONDIN (Src/split_source.c:53)."""
        L.check(self._lib.pb200_split_source(self._h, float(dt), float(g_time)))

    # -- data movement -----------------------------------------------------------------------
    def upload(self, vc: np.ndarray):
        vc = np.ascontiguousarray(vc, dtype=np.float64)
        assert vc.shape == self.shape, (vc.shape, self.shape)
        L.check(self._lib.pb200_upload_vc(self._h, vc.ctypes.data_as(C.c_void_p)))

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        assert out.flags["C_CONTIGUOUS"] and out.shape == self.shape
        L.check(self._lib.pb200_download_vc(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_interior(self, v_int: np.ndarray):
        """Load interior zones [nv][nz][ny][nx] (e.g. a data.NNNN.dbl dump); ghosts are
        filled by the first Boundary() of the step."""
        vc = self.new_vc()
        vc[:, :, :, :] = 1.0   # harmless positive filler for corner ghosts never touched by BCs
        vc[1:4] = 0.0
        vc[self.interior()] = v_int
        self.upload(vc)

    def get_interior(self) -> np.ndarray:
        return self.download()[self.interior()].copy()

    # -- the reference calls -------------------------------------------------------------------
    def boundary(self):
        """Boundary(d, 0, grid)."""
        L.check(self._lib.pb200_boundary(self._h))

    def advance_step(self, dt: float) -> L.StepInfo:
        """AdvanceStep(d, Dts, grid) with g_dt = dt on the device-resident state."""
        L.check(self._lib.pb200_advance_step(self._h, float(dt), C.byref(self.last)))
        return self.last

    def advance_step_host(self, vc: np.ndarray, dt: float) -> L.StepInfo:
        """AdvanceStep on a HOST d->Vc (H2D + step + D2H inside the call)."""
        assert vc.flags["C_CONTIGUOUS"] and vc.shape == self.shape and vc.dtype == np.float64
        L.check(self._lib.pb200_advance_step_host(self._h, vc.ctypes.data_as(C.c_void_p),
                                                  float(dt), C.byref(self.last)))
        return self.last

    def set_owned_planes(self, k0, k1):
        """Deep-halo block of a slab decomposition: only the x3 planes [k0, k1) are downloaded by
        advance_step_host() and counted in invDt_hyp / maxMach (pb200_set_owned_planes)."""
        L.check(self._lib.pb200_set_owned_planes(self._h, int(k0), int(k1)))

    def next_time_step(self, invDt_hyp, cfl, cfl_max_var, g_dt, first_dt) -> float:
        """NextTimeStep(), Src/main.c:521."""
        r = self._lib.pb200_next_time_step(invDt_hyp, cfl, cfl_max_var, g_dt, first_dt)
        if r < 0:
            raise RuntimeError("! NextTimeStep(): dt is too small. Cannot continue.")
        return r

    def integrate(self, nsteps, *, t, dt, tstop, cfl, cfl_max_var, first_dt):
        """nsteps iterations of the main loop (Src/main.c:215-337) without leaving C."""
        tt, dd = C.c_double(t), C.c_double(dt)
        n = L.check(self._lib.pb200_integrate(self._h, int(nsteps), float(tstop), float(cfl),
                                              float(cfl_max_var), float(first_dt), C.byref(tt),
                                              C.byref(dd), C.byref(self.last)))
        return n, tt.value, dd.value

    # -- stage-level access (halo exchange between stages) ---------------------------------
    def step_begin(self, dt):
        L.check(self._lib.pb200_step_begin(self._h, float(dt)))

    def stage(self, s):
        L.check(self._lib.pb200_stage(self._h, int(s)))

    def stage_boundary(self, s):
        L.check(self._lib.pb200_stage_boundary(self._h, int(s)))

    def stage_begin(self, s):
        L.check(self._lib.pb200_stage_begin(self._h, int(s)))

    def stage_finish(self, s):
        L.check(self._lib.pb200_stage_finish(self._h, int(s)))

    def step_end(self) -> L.StepInfo:
        L.check(self._lib.pb200_step_end(self._h, C.byref(self.last)))
        return self.last

    def nstages(self) -> int:
        return self._lib.pb200_nstages(self._h)

    def stage_array_ptr(self, s) -> int:
        return self._lib.pb200_stage_array(self._h, int(s))

    def device_vc_ptr(self) -> int:
        return self._lib.pb200_device_vc(self._h)

    def stream_ptr(self) -> int:
        return self._lib.pb200_stream(self._h)

    def set_profiling(self, on: bool):
        L.check(self._lib.pb200_set_profiling(self._h, int(on)))

    def kernel_times(self):
        """[(ms, dir, stage)] of the sweep launches of the last step (profiling on)."""
        ms = (C.c_float * 16)(); di = (C.c_int * 16)(); st = (C.c_int * 16)()
        n = L.check(self._lib.pb200_kernel_times(self._h, 16, ms, di, st))
        return [(ms[k], di[k], st[k]) for k in range(n)]

    def halo_layout(self, d):
        v = [C.c_long() for _ in range(6)]
        L.check(self._lib.pb200_halo_layout(self._h, d, *[C.byref(x) for x in v]))
        keys = ("lo_ghost", "lo_edge", "hi_edge", "hi_ghost", "count", "var_stride")
        return {k: x.value for k, x in zip(keys, v)}


# --------------------------------------------------------------------------------------
#  Simulation: the reference's main() time loop around Hydro
# --------------------------------------------------------------------------------------
class Simulation:
    """main() loop of Src/main.c:215-337: clip dt at tstop, Integrate, g_time += g_dt,
    g_dt = NextTimeStep, g_stepNumber++.  With cooling=True (COOLING BLONDIN) Integrate alternates
    AdvanceStep / SplitSource (even steps: hydro then source, odd steps: source then hydro,
    Src/main.c:479-485), the time-step accumulators are reset only every second step
    (Src/main.c:406-415) and NextTimeStep runs every second step (Src/main.c:326-330)."""

    def __init__(self, hydro: Hydro, runtime: Runtime, cooling: bool = False):
        self.h = hydro
        self.rt = runtime
        self.cooling = cooling
        self._invDt = 0.0
        self.g_time = 0.0
        self.g_dt = runtime.first_dt
        self.g_stepNumber = 0
        self.g_maxMach = 0.0
        self.history = []     # (nstep, t, dt) before each step, like restart.out records

    def step(self) -> bool:
        """One pass of the loop body; returns True if this was the last step."""
        rt = self.rt
        last = False
        if (self.g_time + self.g_dt) >= rt.tstop * (1.0 - 1.e-8):
            self.g_dt = rt.tstop - self.g_time
            last = True
        self.history.append((self.g_stepNumber, self.g_time, self.g_dt))
        if not self.cooling:
            info = self.h.advance_step(self.g_dt)
            self.g_maxMach = info.maxMach
            self.g_time += self.g_dt
            self.g_dt = self.h.next_time_step(info.invDt_hyp, rt.cfl, rt.cfl_max_var, self.g_dt,
                                              rt.first_dt)
            self.g_stepNumber += 1
            return last
        n = self.g_stepNumber
        if (n - 1) % 2 == 1:               # Dts->invDt_hyp = 0 only every second step
            self._invDt = 0.0
        if n % 2 == 0:
            info = self.h.advance_step(self.g_dt)
            self.h.split_source(self.g_dt, self.g_time)
        else:
            self.h.split_source(self.g_dt, self.g_time)
            info = self.h.advance_step(self.g_dt)
        self.g_maxMach = info.maxMach
        # UpdateStage divides the running maximum by DIMENSIONS on EVERY call (update_stage.c:391-392),
        # so the first step's contribution of a pair is divided twice
        nd = float(self.h.dimensions) if self.h.dimensions > 1 else 1.0
        self._invDt = max(self._invDt / nd, info.invDt_hyp) if self._invDt > 0.0 else info.invDt_hyp
        self.g_time += self.g_dt
        if n % 2 == 1:
            self.g_dt = self.h.next_time_step(self._invDt, rt.cfl, rt.cfl_max_var, self.g_dt, rt.first_dt)
        self.g_stepNumber += 1
        return last

    def run(self, maxsteps=None):
        n = 0
        while True:
            last = self.step()
            n += 1
            if last or (maxsteps is not None and n >= maxsteps):
                break
        return n
