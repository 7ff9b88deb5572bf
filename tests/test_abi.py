"""CPU: libplutob200.so loads and exports every symbol include/*.h declares;
host-only entry points behave like the reference's host logic.  No GPU compute here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from pluto_sirocco_b200 import _lib as L
from pluto_sirocco_b200.build import build_library
from oracle import Oracle

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    build_library()
    return L.load()


def declared_symbols():
    text = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))    # every header of the boundary
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libplutob200.so does not export %s" % n
        assert n in L.SYMBOLS, "ctypes table lacks %s" % n
    assert sorted(L.SYMBOLS) == names


def test_version_and_defaults(lib):
    assert lib.pb200_version() == 100
    cfg = L.Config()
    lib.pb200_config_default(C.byref(cfg))
    assert cfg.gamma == pytest.approx(5.0 / 3.0) and cfg.small_density == 1e-12
    assert cfg.solver == L.HLLC and cfg.nghost == 2


def test_config_validation_without_gpu(lib):
    cfg = L.Config()
    lib.pb200_config_default(C.byref(cfg))
    h = C.c_void_p()
    cfg.dimensions = 4
    assert lib.pb200_create(C.byref(cfg), C.byref(h)) == L.EINVAL
    cfg.dimensions = 1
    cfg.solver = 99
    assert lib.pb200_create(C.byref(cfg), C.byref(h)) == L.EINVAL
    assert b"solver" in lib.pb200_last_error()
    cfg.solver = L.HLLC
    cfg.reconstruction = L.PARABOLIC      # needs 3 ghosts
    assert lib.pb200_create(C.byref(cfg), C.byref(h)) == L.EINVAL


def test_no_cpu_fallback(lib):
    """Without a device create() must fail with ENODEV, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = L.Config()
    lib.pb200_config_default(C.byref(cfg))
    cfg.nx[0] = 64
    h = C.c_void_p()
    assert lib.pb200_create(C.byref(cfg), C.byref(h)) == L.ENODEV
    assert b"no CPU fallback" in lib.pb200_last_error()


@pytest.mark.parametrize("inv,cfl,cmv,dt,fdt", [(123.4, 0.8, 1.1, 1e-4, 1e-4), (0.5, 0.3, 1.1, 1.0, 1e-9),
                                                 (7.7, 0.4, 1.0, 2e-3, 1e-4)])
def test_next_time_step_matches_oracle(lib, inv, cfl, cmv, dt, fdt):
    assert lib.pb200_next_time_step(inv, cfl, cmv, dt, fdt) == Oracle.next_time_step(inv, cfl, cmv, dt, fdt)


def test_next_time_step_too_small(lib):
    assert lib.pb200_next_time_step(1e30, 0.4, 1.1, 1.0, 1.0) < 0
