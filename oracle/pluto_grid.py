"""Restatement of the reference's 1-D grid generation (test infrastructure, see oracle/build_ref.py):
Src/set_grid.c:395-450 MakeGrid() for uniform ('u') and ratio ('r', this fork's addition) patches and
Src/set_grid.c:100-138 ghost-zone extension.  Returns exactly the doubles grid->xl, grid->xr and
grid->dx hold in the reference (same operations in the same order, libm pow)."""
from __future__ import annotations

import math

import numpy as np


def make_grid(spec, nghost):
    """spec = (xL, n, xR) | (xL, n, xR, 'u') uniform, or (xL, n, xR, 'r', ratio); single patch.
    nghost = 0 for an inactive dimension.  Returns xl, xr, dx (np_tot entries each)."""
    xL, n, xR = float(spec[0]), int(spec[1]), float(spec[2])
    kind = spec[3] if len(spec) > 3 else "u"
    xl = np.zeros(n + 2 * nghost)
    xr = np.zeros(n + 2 * nghost)
    dx = np.zeros(n + 2 * nghost)
    b = nghost
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
    for i in range(nghost):
        dx[i] = dx[b]
        xl[i] = xl[b] - (nghost - i) * dx[b]
        xr[i] = xl[i] + dx[b]
        dx[e + i + 1] = dx[e]
        xl[e + i + 1] = xl[e] + (i + 1) * dx[e]
        xr[e + i + 1] = xl[e] + (i + 2) * dx[e]
    return xl, xr, dx


def ini_string(spec):
    """the X?-grid line of pluto.ini for a spec accepted by make_grid."""
    xL, n, xR = spec[0], spec[1], spec[2]
    kind = spec[3] if len(spec) > 3 else "u"
    if kind == "u":
        return "1  %r  %d  u  %r" % (float(xL), int(n), float(xR))
    return "1  %r  %d  r  %r  %r" % (float(xL), int(n), float(xR), float(spec[4]))
