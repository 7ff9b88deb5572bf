"""ctypes wrapper of oracle/hd_oracle.c (the CPU restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"

RECON = dict(FLAT=1, LINEAR=2, PARABOLIC=3)
RK = dict(EULER=1, RK2=2, RK3=3)
SOLVER = dict(tvdlf=1, hll=2, hllc=3)
LIMITER = dict(DEFAULT=0, FLAT_LIM=1, MINMOD_LIM=2, VANLEER_LIM=3, MC_LIM=4, VANALBADA_LIM=5,
               OSPRE_LIM=6, UMIST_LIM=7)
BCS = dict(outflow=1, reflective=2, periodic=5, neighbour=100, userdef=8)


class Cfg(C.Structure):
    _fields_ = [("ndim", C.c_int), ("nx", C.c_int * 3), ("ng", C.c_int), ("nvar", C.c_int),
                ("recon", C.c_int), ("limiter", C.c_int), ("rk", C.c_int), ("solver", C.c_int),
                ("bc", C.c_int * 6), ("gamma", C.c_double), ("small_dn", C.c_double),
                ("small_pr", C.c_double), ("xbeg", C.c_double * 3), ("xend", C.c_double * 3), ("dx", C.c_double * 3),
                ("body_force", C.c_int), ("bf_g", C.c_void_p * 3), ("bf_phi", C.c_void_p * 4)]


def build(force=False):
    if LIB.exists() and not force and LIB.stat().st_mtime >= max((HERE / f).stat().st_mtime
                                                                 for f in ("hd_oracle.c", "gen_oracle.c", "tables_oracle.c")):
        return LIB
    r = subprocess.run(["make", "-C", str(HERE), "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_advance_step.restype = C.c_int
        _lib.orc_advance_step.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_double,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_step_begin.restype = C.c_void_p
        _lib.orc_step_begin.argtypes = [C.POINTER(Cfg), C.c_void_p]
        _lib.orc_stage.restype = None
        _lib.orc_stage.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_step_end.restype = C.c_int
        _lib.orc_step_end.argtypes = [C.c_void_p]
        _lib.orc_boundary.restype = None
        _lib.orc_boundary.argtypes = [C.POINTER(Cfg), C.c_void_p]
        _lib.orc_next_time_step.restype = C.c_double
        _lib.orc_next_time_step.argtypes = [C.c_double] * 5
        _lib.orc_integrate.restype = C.c_int
        _lib.orc_integrate.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return _lib


class Oracle:
    """Same constructor keywords as pluto_sirocco_b200.Hydro so tests can build both alike."""

    def __init__(self, *, dimensions, nx, xbeg=(0., 0., 0.), xend=(1., 1., 1.), gamma=5. / 3.,
                 reconstruction="LINEAR", time_stepping="RK2", solver="hllc", limiter="DEFAULT",
                 bcs=("outflow",) * 6, ntracer=0, nghost=None, small_density=1e-12,
                 small_pressure=1e-12, dx=None, body_force=0, **_):
        c = Cfg()
        c.ndim = dimensions
        for d in range(3):
            c.nx[d] = int(nx[d]) if d < dimensions else 1
            c.xbeg[d] = xbeg[d]
            c.xend[d] = xend[d]
        c.ng = nghost if nghost is not None else (3 if reconstruction == "PARABOLIC" else 2)
        c.nvar = 5 + ntracer
        c.recon = RECON[reconstruction]
        c.limiter = LIMITER[limiter]
        c.rk = RK[time_stepping]
        c.solver = SOLVER[solver]
        for s in range(6):
            c.bc[s] = BCS[bcs[s]] if isinstance(bcs[s], str) else int(bcs[s])
        c.gamma = gamma
        for d in range(3):
            c.dx[d] = dx[d] if dx is not None else 0.0
        c.small_dn = small_density
        c.small_pr = small_pressure
        self.c = c
        self.dimensions = dimensions
        self.nghost = c.ng
        self.nx = tuple(c.nx)
        self.beg = tuple(c.ng if d < dimensions else 0 for d in range(3))
        self.tot = tuple(self.nx[d] + 2 * self.beg[d] for d in range(3))
        self.nvar = c.nvar
        self.xbeg = tuple(xbeg)
        self.xend = tuple(xend)
        self.shape = (self.nvar, self.tot[2], self.tot[1], self.tot[0])
        c.body_force = body_force
        self._bf = {}

    # BODY_FORCE tables: full [k][j][i] arrays incl. ghosts (broadcast from whatever shape)
    def set_body_force_vector(self, comp, tab):
        a = np.ascontiguousarray(np.broadcast_to(tab, self.shape[1:]), dtype=np.float64)
        self._bf[("g", comp)] = a
        self.c.bf_g[comp] = a.ctypes.data

    def set_body_force_potential(self, where, tab):
        a = np.ascontiguousarray(np.broadcast_to(tab, self.shape[1:]), dtype=np.float64)
        self._bf[("phi", where)] = a
        self.c.bf_phi[where] = a.ctypes.data

    def interior(self):
        sl = [slice(None)]
        for d in (2, 1, 0):
            sl.append(slice(self.beg[d], self.beg[d] + self.nx[d]))
        return tuple(sl)

    def embed(self, v_int):
        vc = np.ones(self.shape)
        vc[1:4] = 0.0
        vc[self.interior()] = v_int
        return vc

    def boundary(self, vc):
        lib().orc_boundary(C.byref(self.c), vc.ctypes.data_as(C.c_void_p))

    def advance_step(self, vc, dt):
        """In place on vc; returns (invDt_hyp, maxMach, nfail)."""
        assert vc.flags["C_CONTIGUOUS"] and vc.shape == self.shape and vc.dtype == np.float64
        inv, mach = C.c_double(0.0), C.c_double(0.0)
        nf = lib().orc_advance_step(C.byref(self.c), vc.ctypes.data_as(C.c_void_p), float(dt),
                                    C.byref(inv), C.byref(mach))
        return inv.value, mach.value, nf

    @staticmethod
    def next_time_step(invDt_hyp, cfl, cfl_max_var, g_dt, first_dt):
        return lib().orc_next_time_step(invDt_hyp, cfl, cfl_max_var, g_dt, first_dt)

    def integrate(self, vc, nsteps, *, t, dt, tstop, cfl, cfl_max_var, first_dt):
        tt, dd, mm = C.c_double(t), C.c_double(dt), C.c_double(0.0)
        n = lib().orc_integrate(C.byref(self.c), vc.ctypes.data_as(C.c_void_p), int(nsteps), tstop,
                                cfl, cfl_max_var, first_dt, C.byref(tt), C.byref(dd), C.byref(mm))
        return n, tt.value, dd.value
