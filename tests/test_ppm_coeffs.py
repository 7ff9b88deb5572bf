"""PPM_CoefficientsSet (States/ppm_coeffs.c:60-290) as the library evaluates it on the host, against the oracle's
(which the ppmg_* fixtures pin bit-exact to the compiled reference).  No GPU: pb200_ppm_coefficients is host code."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
from common import CURV_CASES, gen_kwargs_from_golden, hydro_kwargs_from_gen, load_golden

GEO = {"CARTESIAN": 1, "CYLINDRICAL": 2, "POLAR": 3, "SPHERICAL": 4}
PPMG = [c for c in CURV_CASES if c.startswith("ppmg")]


@pytest.mark.parametrize("name", PPMG)
def test_library_ppm_coefficients_equal_the_oracles(name):
    """Interface weights (closed forms on uniform Cartesian / radial grids, PPM_FindWeights' LU solve on stretched
    grids and always in theta, with the 5-point Gauss moments of sin(theta)) and h+ / h- (PPM_Q6_Coeffs): the same
    doubles.  Entry 0 of h+/h- in theta reads thp[-1] in the reference and is never used."""
    from gen_oracle import GenOracle
    from pluto_sirocco_b200 import _lib as L
    lib = L.load()
    g = load_golden(name)
    kw = gen_kwargs_from_golden(g)
    o = GenOracle(**kw)
    hk = hydro_kwargs_from_gen(kw)
    for d in range(int(g["dims"])):
        w0, hp0, hm0 = o.ppm_coefficients(d)
        xl, xr, dx = (np.ascontiguousarray(a, dtype=np.float64) for a in hk["grid_arrays"][d])
        n = xl.size
        w, hp, hm = np.zeros((n, 4)), np.zeros(n), np.zeros(n)
        L.check(lib.pb200_ppm_coefficients(GEO[str(g["geometry"])], d, n, xl.ctypes.data, xr.ctypes.data, dx.ctypes.data,
                                           int(hk["grid_uniform"][d]), w.ctypes.data, hp.ctypes.data, hm.ctypes.data))
        assert np.array_equal(w[1:n - 2], w0[1:n - 2]), (name, d)
        assert np.array_equal(hp[1:], hp0[1:]) and np.array_equal(hm[1:], hm0[1:]), (name, d)
        assert np.allclose(w[1:n - 2].sum(axis=1), 1.0, atol=1e-13)
    o.close()


def test_ppm_coefficients_rejects_bad_arguments():
    from pluto_sirocco_b200 import _lib as L
    lib = L.load()
    x = np.linspace(0.0, 1.0, 9)
    w, hp, hm = np.zeros((8, 4)), np.zeros(8), np.zeros(8)
    assert lib.pb200_ppm_coefficients(9, 0, 8, x[:-1].ctypes.data, x[1:].ctypes.data, None, 1, w.ctypes.data, hp.ctypes.data,
                                      hm.ctypes.data) == L.EINVAL
    xl, xr = np.ascontiguousarray(x[:-1]), np.ascontiguousarray(x[1:])
    L.check(lib.pb200_ppm_coefficients(1, 0, 8, xl.ctypes.data, xr.ctypes.data, None, 1, w.ctypes.data, hp.ctypes.data,
                                       hm.ctypes.data))
    assert np.array_equal(w[1:6], np.tile([-1.0 / 12.0, 7.0 / 12.0, 7.0 / 12.0, -1.0 / 12.0], (5, 1)))
    assert np.array_equal(hp, np.full(8, 3.0)) and np.array_equal(hm, np.full(8, 3.0))
