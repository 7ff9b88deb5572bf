"""SIROCCO table files -> arrays, through the readers of libplutob200.so (csrc/sirocco_tables.c,
include/pluto_b200_tables.h): the host-side counterpart of read_sirocco_fluxes()
(Src/LineDriven/line_connect.c:43-262) for callers that drive the library from Python.

    fr, ft, fp = read_flux_files(rundir, x1, x2, nghost, unit_length)
    hydro.set_ldw(params=..., units=..., flux_r=fr, flux_t=ft, flux_p=fp)

Arrays come back in the library's layout [table][k = 1][j][i] including ghost zones (zero where no
row of the file matches a zone, i.e. in the ghost zones)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _lib as L


class TableGrid(C.Structure):
    """pb200_table_grid"""
    _fields_ = [("nx1_tot", C.c_int), ("nx2_tot", C.c_int), ("ibeg", C.c_int), ("iend", C.c_int),
                ("jbeg", C.c_int), ("jend", C.c_int), ("x1", C.c_void_p), ("x2", C.c_void_p),
                ("unit_length", C.c_double)]


class TableError(RuntimeError):
    pass


_ERR = {-1: "no such file", -2: "bad header", -3: "truncated or malformed row", -4: "number of angular bins differs"}


def _grid(x1, x2, nghost, unit_length):
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    g = TableGrid(len(x1), len(x2), nghost, len(x1) - nghost - 1, nghost, len(x2) - nghost - 1,
                  x1.ctypes.data, x2.ctypes.data, float(unit_length))
    return g, x1, x2     # keep the arrays alive with the struct


def _check(n, path):
    if n < 0:
        raise TableError("%s: %s" % (path, _ERR.get(int(n), "error %d" % n)))
    return int(n)


def read_flux_files(directory, x1, x2, nghost, unit_length):
    """directional_flux_{r,theta,phi}.dat of `directory` -> (flux_r, flux_t, flux_p), each
    [nangles][1][len(x2)][len(x1)]; flux_p is None when its file is absent (the reference prints
    "No flux file" and carries on, line_connect.c:80-83).  x1, x2: zone centres incl. ghosts."""
    lib = L.load()
    g, x1, x2 = _grid(x1, x2, nghost, unit_length)
    d = Path(directory)
    nang = _check(lib.pb200_flux_file_nangles(str(d / "directional_flux_r.dat").encode()), d / "directional_flux_r.dat")
    out = []
    for name in ("r", "theta", "phi"):
        path = d / ("directional_flux_%s.dat" % name)
        if name == "phi" and not path.exists():
            out.append(None)
            continue
        a = np.zeros((nang, 1, g.nx2_tot, g.nx1_tot))
        _check(lib.pb200_read_flux_file(str(path).encode(), C.byref(g), nang, a.ctypes.data), path)
        out.append(a)
    return tuple(out)


def read_mfit_file(directory, x1, x2, nghost, unit_length):
    """M_UV_data.dat -> (t_fit = log10 t [mpoints], m_fit = log10 M [mpoints][1][j][i]): what
    Hydro.set_ldw(t_fit=..., m_fit=...) takes for KRAD = ALPHARAD = 999 (line_connect.c:185-256)."""
    lib = L.load()
    g, x1, x2 = _grid(x1, x2, nghost, unit_length)
    path = Path(directory) / "M_UV_data.dat"
    mp = C.c_int(0)
    _check(lib.pb200_read_mfit_file(str(path).encode(), C.byref(g), C.byref(mp), None, None), path)
    t = np.zeros(mp.value)
    m = np.zeros((mp.value, 1, g.nx2_tot, g.nx1_tot))
    _check(lib.pb200_read_mfit_file(str(path).encode(), C.byref(g), C.byref(mp), t.ctypes.data, m.ctypes.data), path)
    return t, m
