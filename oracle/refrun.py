"""Drive the compiled reference executables in oracle/_ref/<config>/pluto and load their
outputs.  TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py header).

pluto.ini layout follows Src/runtime_setup.c:67-507; `dbl <dt> <dn> single_file` makes the
reference dump d->Vc (interior zones, variable-major, little-endian FP64, bin_io.c:216-261)
every <dn> steps; restart.out carries the exact (t, dt, nstep) of each dump
(structs.h:256-262, restart.c:276-314).
"""
from __future__ import annotations

import os
import struct
import subprocess
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFDIR = HERE / "_ref"


def have_ref(cfg: str) -> bool:
    return (REFDIR / cfg / "pluto").exists()


def write_ini(path, *, grid, cfl=0.4, cfl_max_var=1.1, tstop=1.0, first_dt=1e-4,
              solver="hllc", bcs=("outflow",) * 6, dbl=(-1.0, 1), params=None,
              log=100000):
    """grid = [(lo, n, hi)] * 3 (uniform patches) or full pluto.ini grid strings."""
    L = ["[Grid]", ""]
    for d, g in enumerate(grid):
        if isinstance(g, str):
            L.append("X%d-grid  %s" % (d + 1, g))
        else:
            lo, n, hi = g
            L.append("X%d-grid  1  %r  %d  u  %r" % (d + 1, float(lo), int(n), float(hi)))
    L += ["", "[Chombo Refinement]", "", "Levels 4", "Ref_ratio 2 2 2 2 2",
          "Regrid_interval 2 2 2 2", "Refine_thresh 0.3", "Tag_buffer_size 3",
          "Block_factor 4", "Max_grid_size 32", "Fill_ratio 0.75", "", "[Time]", "",
          "CFL %r" % cfl, "CFL_max_var %r" % cfl_max_var, "tstop %r" % tstop,
          "first_dt %r" % first_dt, "", "[Solver]", "", "Solver %s" % solver, "",
          "[Boundary]", ""]
    names = ["X1-beg", "X1-end", "X2-beg", "X2-end", "X3-beg", "X3-end"]
    for n, b in zip(names, bcs):
        L.append("%s %s" % (n, b))
    L += ["", "[Static Grid Output]", "", "uservar 0",
          "dbl %r %d single_file" % (float(dbl[0]), int(dbl[1])),
          "flt -1.0 -1 single_file", "vtk -1.0 -1 single_file", "tab -1.0 -1",
          "ppm -1.0 -1", "png -1.0 -1", "log %d" % log, "analysis -1.0 -1", "",
          "[Chombo HDF5 output]", "", "Checkpoint_interval -1.0 0",
          "Plot_interval 1.0 0", "", "[Parameters]", ""]
    for k, v in (params or {}).items():
        L.append("%s %r" % (k, v))
    Path(path).write_text("\n".join(L) + "\n")


def read_restart(path):
    """[(nstep, t, dt)] from restart.out: 128-byte records (structs.h:256-262)."""
    raw = Path(path).read_bytes()
    out = []
    for o in range(0, len(raw) - 127, 128):
        nstep = struct.unpack_from("<i", raw, o)[0]
        t, dt = struct.unpack_from("<dd", raw, o + 72)
        out.append((nstep, t, dt))
    return out


def read_dbl(path, nvar, shape):
    """data.NNNN.dbl (single_file) -> array [nvar, nz, ny, nx]."""
    nz, ny, nx = shape
    a = np.fromfile(path, dtype="<f8")
    assert a.size == nvar * nz * ny * nx, (a.size, nvar, shape)
    return a.reshape(nvar, nz, ny, nx)


def run(cfg, workdir, *, shape, nvar=5, maxsteps=None, no_write=False, extra_args=(),
        timeout=300, exe=None, env=None, keep=False, **ini):
    """Run oracle/_ref/<cfg>/pluto in `workdir` with a generated pluto.ini.

    shape = (nz, ny, nx) interior zones.  Returns dict(steps=[(nstep,t,dt)], data=[arrays],
    wall=seconds, log=stdout)."""
    exe = Path(exe) if exe is not None else REFDIR / cfg / "pluto"
    if not exe.exists():
        raise FileNotFoundError("%s missing: run `python oracle/build_ref.py %s`" % (exe, cfg))
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    if not keep:      # keep=True: a restart cycle (-restart) continues from the files of the previous run
        for f in wd.glob("data.*.dbl"):
            f.unlink()
        for f in ("restart.out", "dbl.out", "grid.out"):
            if (wd / f).exists():
                (wd / f).unlink()
    write_ini(wd / "pluto.ini", **ini)
    cmd = [str(exe)]
    if maxsteps is not None:
        cmd += ["-maxsteps", str(maxsteps)]
    if no_write:
        cmd += ["-no-write"]
    cmd += list(extra_args)
    env = dict(os.environ, OMP_NUM_THREADS="1", **(env or {}))
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=wd, capture_output=True, text=True, timeout=timeout, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference run failed (%d):\n%s\n%s" % (r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    res = dict(wall=wall, log=r.stdout, steps=[], data=[])
    if not no_write:
        res["steps"] = read_restart(wd / "restart.out")
        files = sorted(wd.glob("data.*.dbl"))
        res["data"] = [read_dbl(f, nvar, shape) for f in files]
    return res
