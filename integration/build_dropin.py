#!/usr/bin/env python3
"""Link the reference's own driver against libplutob200.so: the drop-in executable.

For a configuration of oracle/build_ref.py this compiles pluto_sirocco_b200/csrc/pluto_shim.c
against the reference headers (+ that configuration's definitions.h) and links it with the
reference objects already built under oracle/_ref/<cfg>/obj -- all of them EXCEPT rk_step.o,
whose only symbol (AdvanceStep) the shim provides -- into integration/_build/<cfg>/pluto_b200.
Needs /root/reference (headers) and oracle/_ref/<cfg>/obj; the result travels to the GPU box.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import build_ref  # noqa: E402

OUT = ROOT / "integration" / "_build"


def build(cfg: str, force: bool = False) -> Path:
    from pluto_sirocco_b200.build import build_library
    lib = build_library()
    build_ref.build(cfg)
    wd = build_ref.OUT / cfg
    out = OUT / cfg
    exe = out / "pluto_b200"
    shim = ROOT / "pluto_sirocco_b200" / "csrc" / "pluto_shim.c"
    if exe.exists() and not force and exe.stat().st_mtime > max(shim.stat().st_mtime, lib.stat().st_mtime):
        return exe
    out.mkdir(parents=True, exist_ok=True)
    s = build_ref.REF / "Src"
    import re
    eos = "Isothermal" if re.search(r"^#define\s+EOS\s+ISOTHERMAL", (wd / "definitions.h").read_text(), re.M) else "Ideal"
    incs = ["-I%s" % wd, "-I%s" % s, "-I%s" % (s / "HD"), "-I%s" % (s / "EOS" / eos),
            "-I%s" % (s / "States"), "-I%s" % (s / "Math_Tools"), "-I%s" % (ROOT / "include")]
    obj = out / "pluto_shim.o"
    r = subprocess.run(["gcc", "-c", "-O2", "-std=gnu17"] + incs + [str(shim), "-o", str(obj)],
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("shim compile failed:\n" + r.stderr[-4000:])
    # read_sirocco_fluxes is wrapped too: the shim can hand the table files to the library's readers
    # (include/pluto_b200_tables.h, exported by libplutob200.so)
    objs = [str(o) for o in sorted((wd / "obj").glob("*.o")) if o.name != "rk_step.o"]
    ldw = (wd / "obj" / "line_connect.o").exists()      # read_sirocco_heatcool exists in line-driven-wind builds only
    cmd = ["gcc"] + objs + [str(obj),
           "-Wl,--wrap=WriteData,--wrap=Analysis,--wrap=SplitSource,--wrap=read_sirocco_fluxes" + (",--wrap=read_sirocco_heatcool" if ldw else ""), "-L%s" % lib.parent, "-lplutob200",
           "-Wl,-rpath,$ORIGIN/../../../pluto_sirocco_b200/lib", "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return exe


if __name__ == "__main__":
    for c in (sys.argv[1:] or ["sod", "sedov3d", "sedov3d_ppm"]):
        print("built", build(c, force=True))
