"""ctypes wrapper of oracle/gen_oracle.c (general-grid CPU restatement).  TEST INFRASTRUCTURE
ONLY: imported by tests/, never by the product."""
from __future__ import annotations

import ctypes as C

import numpy as np

import oracle as _o
from pluto_grid import make_grid

GEOMETRY = dict(CARTESIAN=1, CYLINDRICAL=2, POLAR=3, SPHERICAL=4)
BCS = dict(outflow=1, reflective=2, axisymmetric=3, eqtsymmetric=4, periodic=5, userdef=8, polaraxis=9, neighbour=100)


class GenCfg(C.Structure):
    _fields_ = [("ndim", C.c_int), ("nx", C.c_int * 3), ("ng", C.c_int), ("ntracer", C.c_int),
                ("entropy", C.c_int), ("geometry", C.c_int), ("limiter", C.c_int),
                ("char_limiting", C.c_int), ("flattening", C.c_int), ("rk", C.c_int), ("solver", C.c_int),
                ("bc", C.c_int * 6), ("gamma", C.c_double), ("small_dn", C.c_double), ("small_pr", C.c_double),
                ("xl", C.c_void_p * 3), ("xr", C.c_void_p * 3), ("body_force", C.c_int),
                ("bf_g", C.c_void_p * 3),
                ("ldw", C.c_int), ("ldw_bc", C.c_int), ("nangles", C.c_int),
                ("flux_r", C.c_void_p), ("flux_t", C.c_void_p), ("flux_p", C.c_void_p),
                ("unit_length", C.c_double), ("unit_velocity", C.c_double), ("unit_density", C.c_double),
                ("mu", C.c_double), ("krad", C.c_double), ("alpharad", C.c_double), ("t_iso", C.c_double),
                ("dfloor", C.c_double), ("rho0", C.c_double), ("rho_alpha", C.c_double),
                ("cent_mass", C.c_double), ("disk_mdot", C.c_double),
                ("cooling", C.c_int), ("cool_tab", C.c_void_p * 8), ("lx", C.c_double), ("tx", C.c_double),
                ("mpoints", C.c_int), ("t_fit", C.c_void_p), ("m_fit", C.c_void_p),
                ("iso", C.c_int), ("iso_cs", C.c_double), ("flatten_oned", C.c_int),
                ("ppm", C.c_int), ("uniform", C.c_int * 3), ("bf_phi", C.c_void_p * 4),
                ("ring_average", C.c_int), ("ring_rec", C.c_int)]


_bound = False


def lib():
    global _bound
    L = _o.lib()
    if not _bound:
        L.gen_create.restype = C.c_void_p
        L.gen_create.argtypes = [C.POINTER(GenCfg), C.c_void_p, C.c_void_p, C.c_void_p]
        L.gen_destroy.argtypes = [C.c_void_p]
        L.gen_nvar.argtypes = [C.c_void_p]
        L.gen_boundary.argtypes = [C.c_void_p, C.c_void_p]
        L.gen_get_geometry.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gen_get_ppm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gen_blondin_cooling.restype = None
        L.gen_blondin_cooling.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.gen_advance_step.restype = C.c_int
        L.gen_advance_step.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
        _bound = True
    return L


class GenOracle:
    """Keyword-compatible with pluto_sirocco_b200.Hydro (general-grid options included)."""

    def __init__(self, *, dimensions, grid, geometry="CARTESIAN", gamma=5. / 3., reconstruction="LINEAR",
                 time_stepping="RK2", solver="hllc", limiter="DEFAULT", bcs=("outflow",) * 6, ntracer=0,
                 nghost=2, small_density=1e-12, small_pressure=1e-12, body_force=0, char_limiting=False,
                 shock_flattening=False, entropy_switch=False, ldw=None, eos="IDEAL",
                 iso_sound_speed=1.0, ring_average=0, ring_average_rec=None, **_):
        assert reconstruction in ("LINEAR", "PARABOLIC")
        c = GenCfg()
        c.ndim = dimensions
        self._keep = []
        self.xl, self.xr, self.dx = [], [], []
        for d in range(3):
            ng = nghost if d < dimensions else 0
            xl, xr, dx = make_grid(grid[d], ng)
            self.xl.append(xl); self.xr.append(xr); self.dx.append(dx)
            c.nx[d] = int(grid[d][1]) if d < dimensions else 1
            c.xl[d] = xl.ctypes.data
            c.xr[d] = xr.ctypes.data
        c.ng = nghost
        c.ppm = int(reconstruction == "PARABOLIC")
        for d in range(3):      # grid->uniform[d]: a single uniform patch (set_grid.c:67-72)
            c.uniform[d] = int(len(grid[d]) <= 3 or grid[d][3] == "u")
        c.ntracer = ntracer
        c.entropy = {False: 0, None: 0, True: 2, "NO": 0, "SELECTIVE": 1, "ALWAYS": 2}[entropy_switch]
        c.geometry = GEOMETRY[geometry]
        c.limiter = _o.LIMITER[limiter]
        c.char_limiting = int(bool(char_limiting))
        c.flattening = int(bool(shock_flattening) and shock_flattening != "ONED")     # MULTID
        c.flatten_oned = int(shock_flattening == "ONED")
        c.rk = _o.RK[time_stepping]
        c.solver = dict(_o.SOLVER, roe=4, two_shock=5, **{"ausm+": 6})[solver]     # Roe, TwoShock, AUSM+: general-grid oracle only
        for s in range(6):
            c.bc[s] = BCS[bcs[s]] if isinstance(bcs[s], str) else int(bcs[s])
        c.gamma = gamma
        c.small_dn = small_density
        c.small_pr = small_pressure
        c.body_force = body_force
        c.iso = int(eos == "ISOTHERMAL")     # NFLX = 4: (rho, v1, v2, v3), tracers from index 4
        c.iso_cs = float(iso_sound_speed)
        c.ring_average = int(ring_average)     # RING_AVERAGE (pluto.h:479-489): REC defaults to 5 when on
        c.ring_rec = int(ring_average_rec) if ring_average_rec else (5 if c.ring_average > 1 else 1)
        assert not (c.iso and c.entropy), "ENTROPY_SWITCH needs an energy equation"
        self.c = c
        self.dimensions = dimensions
        self.nghost = nghost
        self.nx = tuple(c.nx)
        self.beg = tuple(nghost if d < dimensions else 0 for d in range(3))
        self.tot = tuple(self.nx[d] + 2 * self.beg[d] for d in range(3))
        self.nvar = (4 if c.iso else 5) + ntracer + (1 if c.entropy else 0)
        self.shape = (self.nvar, self.tot[2], self.tot[1], self.tot[0])
        self._h = None
        self._bf = {}
        if ldw is not None:
            self.set_ldw(**ldw)

    def x(self, d):
        return 0.5 * (self.xl[d] + self.xr[d])

    def set_body_force_vector(self, comp, tab):
        a = np.ascontiguousarray(np.broadcast_to(tab, self.shape[1:]), dtype=np.float64)
        self._bf[comp] = a
        self.c.bf_g[comp] = a.ctypes.data

    def set_body_force_potential(self, where, tab):
        """where: 0 zone centres, 1..3 the x1 / x2 / x3 upper faces (same convention as Hydro)."""
        a = np.ascontiguousarray(np.broadcast_to(tab, self.shape[1:]), dtype=np.float64)
        self._bf[10 + where] = a
        self.c.bf_phi[where] = a.ctypes.data

    def set_ldw(self, *, params, units, flux_r, flux_t, flux_p, userdef_bc=True, t_fit=None, m_fit=None):
        """LINE_DRIVEN_WIND SIROCCO_MODE: g_inputParam[] of cv_idl (dict by label), UNIT_* (dict),
        directional fluxes [nangles][k][j][i] incl. ghosts."""
        c = self.c
        c.ldw = 1
        c.ldw_bc = int(userdef_bc)
        self._flux = [np.ascontiguousarray(a, dtype=np.float64) for a in (flux_r, flux_t, flux_p)]
        assert self._flux[0].shape[1:] == self.shape[1:], (self._flux[0].shape, self.shape)
        c.nangles = self._flux[0].shape[0]
        c.flux_r, c.flux_t, c.flux_p = (a.ctypes.data for a in self._flux)
        c.unit_length, c.unit_velocity, c.unit_density = units["length"], units["velocity"], units["density"]
        c.mu, c.krad, c.alpharad, c.t_iso = params["MU"], params["KRAD"], params["ALPHARAD"], params["T_ISO"]
        c.dfloor, c.rho0, c.rho_alpha = params["DFLOOR"], params["RHO_0"], params["RHO_ALPHA"]
        c.cent_mass, c.disk_mdot = params["CENT_MASS"], params["DISK_MDOT"]
        c.lx, c.tx = params["L_star"] * params["f_x"], params["T_x"]
        if t_fit is not None:      # force-multiplier fit: t_fit = log10(t) [MPOINTS], m_fit = log10(M) [MPOINTS][k][j][i]
            self._tfit = np.ascontiguousarray(t_fit, dtype=np.float64)
            self._mfit = np.ascontiguousarray(m_fit, dtype=np.float64)
            assert self._mfit.shape == (self._tfit.size,) + self.shape[1:]
            c.mpoints, c.t_fit, c.m_fit = self._tfit.size, self._tfit.ctypes.data, self._mfit.ctypes.data

    def _handle(self):
        if self._h is None:
            self._h = lib().gen_create(C.byref(self.c), self.dx[0].ctypes.data, self.dx[1].ctypes.data,
                                       self.dx[2].ctypes.data)
        return self._h

    def close(self):
        if self._h is not None:
            lib().gen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def interior(self):
        sl = [slice(None)]
        for d in (2, 1, 0):
            sl.append(slice(self.beg[d], self.beg[d] + self.nx[d]))
        return tuple(sl)

    def embed(self, v_int):
        vc = np.ones(self.shape)
        vc[1:4] = 0.0
        vc[self.interior()][:v_int.shape[0]] = v_int
        return vc

    def boundary(self, vc):
        lib().gen_boundary(self._handle(), vc.ctypes.data)

    def geometry(self, which):
        out = np.zeros(self.shape[1:])
        lib().gen_get_geometry(self._handle(), which, out.ctypes.data)
        return out

    def ppm_coefficients(self, d):
        """(w[tot][4], h+[tot], h-[tot]) of PPM_CoefficientsSet for direction d."""
        n = self.shape[3 - d]
        w, hp, hm = np.zeros((n, 4)), np.zeros(n), np.zeros(n)
        lib().gen_get_ppm(self._handle(), d, w.ctypes.data, hp.ctypes.data, hm.ctypes.data)
        return w, hp, hm

    def advance_step(self, vc, dt):
        assert vc.flags["C_CONTIGUOUS"] and vc.shape == self.shape and vc.dtype == np.float64
        inv, mach = C.c_double(0.0), C.c_double(0.0)
        nf = lib().gen_advance_step(self._handle(), vc.ctypes.data, float(dt), C.byref(inv), C.byref(mach))
        return inv.value, mach.value, nf

    def blondin_cooling(self, vc, dt, g_time, tabs):
        """SplitSource() -> BlondinCooling(d->Vc, d, dt): tabs = [comp_h_pre, comp_c_pre, xray_h_pre,
        line_c_pre, brem_c_pre, sirocco_xi, sirocco_t_r], each [k][j][i] incl. ghosts."""
        arrs = [np.ascontiguousarray(np.broadcast_to(t, self.shape[1:]), dtype=np.float64) for t in tabs]
        ptrs = (C.c_void_p * 7)(*[a.ctypes.data for a in arrs])
        lib().gen_blondin_cooling(self._handle(), vc.ctypes.data, float(dt), float(g_time), ptrs)

    next_time_step = staticmethod(_o.Oracle.next_time_step)
