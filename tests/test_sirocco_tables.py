"""CPU: the bisection-based SIROCCO table readers of the product (pluto_sirocco_b200/csrc/
sirocco_tables.c, include/pluto_b200_tables.h) against the restatement of the reference's
O(rows x zones) readers (oracle/tables_oracle.c <- Src/LineDriven/line_connect.c:43-262): identical
arrays and zone counts for the files of the line-driven-wind problem, for shuffled rows, rows that
match no zone (ghost-zone rows, foreign coordinates), coordinates perturbed inside / outside the
reference's 1e-6 tolerance, duplicated rows, and truncated files."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as _o              # noqa: E402
import pluto_grid                # noqa: E402
from common import (LDW_UNITS, ldw_flux_tables, ldw_mfit_tables, write_ldw_flux_files,   # noqa: E402
                    write_ldw_mfit_file)


class TGrid(C.Structure):
    _fields_ = [("nx1_tot", C.c_int), ("nx2_tot", C.c_int), ("ibeg", C.c_int), ("iend", C.c_int),
                ("jbeg", C.c_int), ("jend", C.c_int), ("x1", C.c_void_p), ("x2", C.c_void_p),
                ("unit_length", C.c_double)]


@pytest.fixture(scope="module")
def readers(tmp_path_factory):
    """(libplutob200.so - the readers are plain host C inside the product library - , oracle library)."""
    sys.path.insert(0, str(ROOT))
    from pluto_sirocco_b200.build import build_library
    P = C.CDLL(str(build_library()))
    P.pb200_read_flux_file.restype = C.c_long
    P.pb200_read_flux_file.argtypes = [C.c_char_p, C.POINTER(TGrid), C.c_int, C.c_void_p]
    P.pb200_read_mfit_file.restype = C.c_long
    P.pb200_read_mfit_file.argtypes = [C.c_char_p, C.POINTER(TGrid), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    P.pb200_flux_file_nangles.argtypes = [C.c_char_p]
    O = _o.lib()
    O.ref_read_flux_file.restype = C.c_long
    O.ref_read_flux_file.argtypes = [C.c_char_p, C.POINTER(TGrid), C.POINTER(C.c_int), C.c_void_p]
    O.ref_read_mfit_file.restype = C.c_long
    O.ref_read_mfit_file.argtypes = [C.c_char_p, C.POINTER(TGrid), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    for L_, pre in ((P, "pb200_"), (O, "ref_")):
        f = getattr(L_, pre + "read_heatcool_file")
        f.restype, f.argtypes = C.c_long, [C.c_char_p, C.POINTER(TGrid), C.c_void_p, C.c_void_p]
        f = getattr(L_, pre + "read_prefactors_file")
        f.restype, f.argtypes = C.c_long, [C.c_char_p, C.POINTER(TGrid), C.c_void_p]
    return P, O


def make_grid(n1=24, n2=18, ng=3):
    xl1, xr1, _ = pluto_grid.make_grid((0.87, n1, 8.7, "r", 1.05), ng)
    xl2, xr2, _ = pluto_grid.make_grid((0.0, n2, 1.5707963267948966, "r", 0.95), ng)
    x1, x2 = np.ascontiguousarray(0.5 * (xl1 + xr1)), np.ascontiguousarray(0.5 * (xl2 + xr2))
    g = TGrid(len(x1), len(x2), ng, ng + n1 - 1, ng, ng + n2 - 1, x1.ctypes.data, x2.ctypes.data, LDW_UNITS["length"])
    return g, x1, x2, ng


def read_both(readers, path, g, nang, shape):
    P, O = readers
    a = np.full(shape, -7.0)
    b = np.full(shape, -7.0)
    n_ref = C.c_int(0)
    na = P.pb200_read_flux_file(str(path).encode(), C.byref(g), nang, a.ctypes.data)
    nb = O.ref_read_flux_file(str(path).encode(), C.byref(g), C.byref(n_ref), b.ctypes.data)
    return na, a, nb, b, n_ref.value


def test_flux_files_of_the_wind_problem(readers, tmp_path):
    g, x1, x2, ng = make_grid()
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    P, _ = readers
    for name, tab in (("r", fr), ("theta", ft), ("phi", fp)):
        path = tmp_path / ("directional_flux_%s.dat" % name)
        assert P.pb200_flux_file_nangles(str(path).encode()) == fr.shape[0]
        na, a, nb, b, nang = read_both(readers, path, g, fr.shape[0], tab[:, 0].shape)
        assert na == nb == (g.iend - g.ibeg + 1) * (g.jend - g.jbeg + 1) and nang == fr.shape[0]
        assert np.array_equal(a, b)
        inner = (slice(None), slice(g.jbeg, g.jend + 1), slice(g.ibeg, g.iend + 1))
        assert np.array_equal(a[inner], tab[:, 0][inner])          # what the file was written from
        assert np.all(a[:, :g.jbeg] == -7.0) and np.all(a[:, :, :g.ibeg] == -7.0)   # ghosts untouched


def _rows(path):
    lines = Path(path).read_text().splitlines()
    return lines[:2], lines[2:]


def test_shuffled_foreign_perturbed_and_duplicated_rows(readers, tmp_path):
    g, x1, x2, ng = make_grid(20, 14)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    head, rows = _rows(tmp_path / "directional_flux_r.dat")
    rng = np.random.default_rng(5)
    rng.shuffle(rows)
    nang = fr.shape[0]
    vals = " ".join("%.17e" % v for v in rng.uniform(1.0, 2.0, nang))
    UL = LDW_UNITS["length"]
    extra = [
        "0 0 0 %.17e %.17e %s" % (x1[0] * UL, x2[0], vals),                    # a ghost zone: DOM_LOOP never sees it
        "7 7 0 %.17e %.17e %s" % (123.0 * UL, 0.5, vals),                      # matches nothing
        "3 4 0 %.17e %.17e %s" % (x1[ng + 3] * UL * (1 + 4e-7), x2[ng + 4] * (1 - 4e-7), vals),   # inside the tolerance
        "3 4 0 %.17e %.17e %s" % (x1[ng + 5] * UL * (1 + 3e-6), x2[ng + 4], vals),                # outside
        "9 9 0 %.17e %.17e %s" % (x1[ng + 9] * UL, x2[ng + 9], vals),          # duplicate of an existing zone: last row wins
        "0 0 0 %.17e %.17e %s" % (x1[ng] * UL * (1 - 9.9e-7), x2[ng] * (1 + 9.9e-7), vals),       # first zone, edge of the tolerance
        "0 0 0 %.17e %.17e %s" % (x1[g.iend] * UL * (1 + 9.9e-7), x2[g.jend] * (1 - 9.9e-7), vals),   # last zone
    ]
    rows = rows[:50] + extra[:4] + rows[50:] + extra[4:]
    path = tmp_path / "mixed.dat"
    path.write_text("\n".join(head + rows) + "\n")
    na, a, nb, b, _ = read_both(readers, path, g, nang, fr[:, 0].shape)
    assert na == nb and np.array_equal(a, b)
    assert na == len(rows) - 3          # the ghost row, the foreign row and the out-of-tolerance row match no zone


def test_truncated_file_and_wrong_bin_count(readers, tmp_path):
    g, x1, x2, ng = make_grid(10, 8)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    P, O = readers
    path = tmp_path / "directional_flux_theta.dat"
    text = path.read_text()
    (tmp_path / "cut.dat").write_text(text[: len(text) // 2].rsplit(" ", 3)[0])
    a = np.zeros(ft[:, 0].shape)
    n = C.c_int(0)
    assert P.pb200_read_flux_file(str(tmp_path / "cut.dat").encode(), C.byref(g), ft.shape[0], a.ctypes.data) == -3
    assert O.ref_read_flux_file(str(tmp_path / "cut.dat").encode(), C.byref(g), C.byref(n), a.ctypes.data) == -3
    assert P.pb200_read_flux_file(str(path).encode(), C.byref(g), ft.shape[0] + 1, a.ctypes.data) == -4
    assert P.pb200_read_flux_file(str(tmp_path / "absent.dat").encode(), C.byref(g), ft.shape[0], a.ctypes.data) == -1


def test_force_multiplier_fit_file(readers, tmp_path):
    g, x1, x2, ng = make_grid(16, 12)
    t, M, lt, lM = ldw_mfit_tables(x1, x2)
    write_ldw_mfit_file(tmp_path, x1, x2, ng, t, M)
    P, O = readers
    path = str(tmp_path / "M_UV_data.dat").encode()
    mp = C.c_int(0)
    assert P.pb200_read_mfit_file(path, C.byref(g), C.byref(mp), None, None) == 0 and mp.value == len(t)
    shape = (len(t),) + M[:, 0].shape[1:]
    ta, tb = np.zeros(len(t)), np.zeros(len(t))
    a, b = np.full(shape, -7.0), np.full(shape, -7.0)
    mb = C.c_int(0)
    na = P.pb200_read_mfit_file(path, C.byref(g), C.byref(mp), ta.ctypes.data, a.ctypes.data)
    nb = O.ref_read_mfit_file(path, C.byref(g), C.byref(mb), tb.ctypes.data, b.ctypes.data)
    assert na == nb == (g.iend - g.ibeg + 1) * (g.jend - g.jbeg + 1) and mp.value == mb.value
    assert np.array_equal(ta, tb) and np.array_equal(a, b)
    assert np.array_equal(ta, lt)       # log10(t), what Hydro.set_ldw(t_fit=...) is handed in the parity tests
    inner = (slice(None), slice(g.jbeg, g.jend + 1), slice(g.ibeg, g.iend + 1))
    assert np.array_equal(a[inner], lM[:, 0][inner])


def test_non_monotonic_axis_falls_back_to_a_scan(readers, tmp_path):
    """A theta axis that is not ascending (never produced by set_grid.c, but the reference's reader
    does not care): same result through the linear-scan fallback."""
    g, x1, x2, ng = make_grid(12, 10)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    x2r = np.ascontiguousarray(x2[::-1])
    g2 = TGrid(g.nx1_tot, g.nx2_tot, g.ibeg, g.iend, g.jbeg, g.jend, g.x1, x2r.ctypes.data, g.unit_length)
    na, a, nb, b, _ = read_both(readers, tmp_path / "directional_flux_r.dat", g2, fr.shape[0], fr[:, 0].shape)
    assert na == nb > 0 and np.array_equal(a, b)


def test_bisection_is_faster_than_the_zone_scan(readers, tmp_path):
    """96 x 64 zones, 36 bins: the reference-style reader tests 6144 zones per row; report both times."""
    g, x1, x2, ng = make_grid(96, 64)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    P, O = readers
    path = str(tmp_path / "directional_flux_r.dat").encode()
    a, b = np.zeros(fr[:, 0].shape), np.zeros(fr[:, 0].shape)
    n = C.c_int(0)
    t0 = time.perf_counter(); P.pb200_read_flux_file(path, C.byref(g), fr.shape[0], a.ctypes.data); t1 = time.perf_counter()
    O.ref_read_flux_file(path, C.byref(g), C.byref(n), b.ctypes.data); t2 = time.perf_counter()
    assert np.array_equal(a, b)
    print("index-free bisection reader %.3f s, zone-scan reader %.3f s" % (t1 - t0, t2 - t1))
    assert (t1 - t0) < (t2 - t1)


def test_dropin_reads_the_tables_through_the_wrapped_reader(tmp_path):
    """The drop-in executable (reference driver + pluto_shim.c, --wrap=read_sirocco_fluxes): with
    PB200_FAST_TABLES=1 main() reads the three flux files through the library's readers and reports
    the same zone counts as the reference's reader.  Without a GPU the run then stops - loudly - at
    the first AdvanceStep() (no CPU fallback); with one it just takes its two steps."""
    import refrun
    from common import LDW_BCS, LDW_PARAMS
    exe = ROOT / "integration" / "_build" / "ldw" / "pluto_b200"
    if not exe.exists():
        pytest.skip("drop-in executable not built (needs /root/reference at build time)")
    grid = [(0.87, 48, 8.7, "r", 1.05), (0.0, 36, 1.5707963267948966, "r", 0.95), (0.0, 1, 1.0)]
    xl1, xr1, _ = pluto_grid.make_grid(grid[0], 3)
    xl2, xr2, _ = pluto_grid.make_grid(grid[1], 3)
    x1, x2 = 0.5 * (xl1 + xr1), 0.5 * (xl2 + xr2)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    logs = {}
    for fast in ("0", "1"):
        wd = tmp_path / fast
        wd.mkdir()
        write_ldw_flux_files(wd, x1, x2, 3, fr, ft, fp)
        try:
            out = refrun.run("ldw", wd, shape=(1, 36, 48), nvar=6, maxsteps=2, timeout=200, exe=exe,
                             env={"PB200_FAST_TABLES": fast}, grid=[pluto_grid.ini_string(g) for g in grid],
                             cfl=0.4, tstop=1.0, first_dt=1e-4, solver="hll", bcs=LDW_BCS, dbl=(-1.0, 1),
                             params=LDW_PARAMS)
            logs[fast] = out["log"]
        except Exception as e:          # no GPU: refrun raises with the log of the failed run
            logs[fast] = str(e)
            assert "no CUDA device" in logs[fast], logs[fast][-2000:]
    assert "libplutob200 table readers" in logs["1"] and "libplutob200 table readers" not in logs["0"]
    for fast in ("0", "1"):
        assert logs[fast].count("Read 36 fluxes for 1728 cells") == 3, logs[fast][-2000:]


def _write_heatcool(path, x1, x2, ng, rng, extra=()):
    """py_heatcool.dat as read by read_sirocco_heatcool() (line_connect.c:334-341): a header line, then
    i j rcen thetacen vol t_e t_r xi ne heat_xray heat_comp heat_lines heat_ff cool_comp cool_lines cool_ff rho n_h"""
    UL = LDW_UNITS["length"]
    rows = []
    for j in range(ng, len(x2) - ng):
        for i in range(ng, len(x1) - ng):
            v = rng.uniform(0.1, 10.0, 14)
            v[2] = 10.0 ** rng.uniform(2.0, 6.0)       # t_r: some below the 1e3 floor
            v[3] = 10.0 ** rng.uniform(-1.0, 4.0)      # xi: some below the floor of 1
            rows.append("%d %d %.17e %.17e %s" % (i - ng, j - ng, x1[i] * UL, x2[j], " ".join("%.17e" % q for q in v)))
    rng.shuffle(rows)
    Path(path).write_text("# header\n" + "\n".join(list(rows) + list(extra)) + "\n")


def test_heatcool_and_prefactor_files(readers, tmp_path):
    g, x1, x2, ng = make_grid(22, 16)
    P, O = readers
    rng = np.random.default_rng(11)
    UL = LDW_UNITS["length"]
    junk = " ".join("%.17e" % q for q in rng.uniform(1.0, 2.0, 14))
    extra = ["0 0 %.17e %.17e %s" % (x1[0] * UL, x2[0], junk),                                   # ghost zone
             "5 5 %.17e %.17e %s" % (x1[ng + 5] * UL * (1 + 6e-6), x2[ng + 5] * (1 - 6e-6), junk),   # inside 1e-5
             "6 5 %.17e %.17e %s" % (x1[ng + 6] * UL * (1 + 3e-5), x2[ng + 5], junk)]            # outside
    _write_heatcool(tmp_path / "py_heatcool.dat", x1, x2, ng, rng, extra)
    shape = (g.nx2_tot, g.nx1_tot)
    xa, ta, xb, tb = (np.full(shape, -7.0) for _ in range(4))
    path = str(tmp_path / "py_heatcool.dat").encode()
    na = P.pb200_read_heatcool_file(path, C.byref(g), xa.ctypes.data, ta.ctypes.data)
    nb = O.ref_read_heatcool_file(path, C.byref(g), xb.ctypes.data, tb.ctypes.data)
    assert na == nb == 22 * 16 + 1
    assert np.array_equal(xa, xb) and np.array_equal(ta, tb)
    inner = (slice(g.jbeg, g.jend + 1), slice(g.ibeg, g.iend + 1))
    assert xa[inner].min() >= 1.0 and ta[inner].min() >= 1.0e3 and (xa[inner] == 1.0).any() and (ta[inner] == 1.0e3).any()
    # prefactors.dat (line_connect.c:433-437): header, then i rcen j thetacen dens comp_h comp_c xray_h brem_c line_c xi_ion
    rows = []
    for j in range(ng, len(x2) - ng):
        for i in range(ng, len(x1) - ng):
            rows.append("%d %.17e %d %.17e %s" % (i - ng, x1[i] * UL, j - ng, x2[j],
                                                  " ".join("%.17e" % q for q in rng.uniform(0.5, 2.0, 7))))
    rng.shuffle(rows)
    rows = rows[:-5]                                   # five zones without a row keep their preset value
    (tmp_path / "prefactors.dat").write_text("# header\n" + "\n".join(rows) + "\n")
    pa, pb = np.ones((6,) + shape), np.ones((6,) + shape)
    path = str(tmp_path / "prefactors.dat").encode()
    na = P.pb200_read_prefactors_file(path, C.byref(g), pa.ctypes.data)
    nb = O.ref_read_prefactors_file(path, C.byref(g), pb.ctypes.data)
    assert na == nb == 22 * 16 - 5 and np.array_equal(pa, pb)
    # malformed line -> the reference exits; both report -3
    (tmp_path / "bad.dat").write_text("# header\n1 2 3.0\n")
    assert P.pb200_read_prefactors_file(str(tmp_path / "bad.dat").encode(), C.byref(g), pa.ctypes.data) == -3
    assert O.ref_read_prefactors_file(str(tmp_path / "bad.dat").encode(), C.byref(g), pb.ctypes.data) == -3
    assert P.pb200_read_heatcool_file(str(tmp_path / "none.dat").encode(), C.byref(g), xa.ctypes.data, ta.ctypes.data) == -1


def test_python_front_end_round_trip(tmp_path):
    """pluto_sirocco_b200.tables: the files written from the synthetic tables come back as the arrays
    Hydro.set_ldw() is handed in the parity tests (interior zones; ghost zones zero)."""
    from pluto_sirocco_b200 import tables
    g, x1, x2, ng = make_grid(20, 14)
    fr, ft, fp = ldw_flux_tables(x1, x2)
    write_ldw_flux_files(tmp_path, x1, x2, ng, fr, ft, fp)
    got = tables.read_flux_files(tmp_path, x1, x2, ng, LDW_UNITS["length"])
    inner = (slice(None), slice(None), slice(ng, -ng), slice(ng, -ng))
    for a, b in zip(got, (fr, ft, fp)):
        assert a.shape == b.shape and np.array_equal(a[inner], b[inner])
        assert not a[:, :, :ng].any() and not a[:, :, :, :ng].any()
    (tmp_path / "directional_flux_phi.dat").unlink()
    assert tables.read_flux_files(tmp_path, x1, x2, ng, LDW_UNITS["length"])[2] is None
    t, M, lt, lM = ldw_mfit_tables(x1, x2)
    write_ldw_mfit_file(tmp_path, x1, x2, ng, t, M)
    tf, mf = tables.read_mfit_file(tmp_path, x1, x2, ng, LDW_UNITS["length"])
    assert np.array_equal(tf, lt) and np.array_equal(mf[inner], lM[inner])
    with pytest.raises(tables.TableError):
        tables.read_mfit_file(tmp_path / "nowhere", x1, x2, ng, LDW_UNITS["length"])
