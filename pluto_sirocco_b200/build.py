"""Build libplutob200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

`python -m pluto_sirocco_b200.build` or build_library().  nvcc cross-compiles without a GPU.
The shared object is kept in pluto_sirocco_b200/lib/ (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libplutob200.so"
SOURCES = ["pb200.cu", "pb200_sweeps.cu", "pb200_gen.cu", "pb200_cool.cu", "pb200_multi.cu", "sirocco_tables.c"]
HEADERS = ["hd_physics.cuh", "pb200_kernels.cuh", "gen_kernels.cuh", "glibc_math.cuh", "glibc_libm_tables.h", "pb200_internal.h", "../../include/pluto_b200.h",
           "../../include/pluto_b200_tables.h"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
# translation units: (object name, source, extra defines).  pb200_sweeps.cu is compiled once per
# (NVAR, BODY_FORCE) pair so that the kernel instantiations build in parallel.
UNITS = [("pb200", "pb200.cu", []), ("pb200_gen", "pb200_gen.cu", []), ("pb200_multi", "pb200_multi.cu", []),
         ("pb200_cool", "pb200_cool.cu", ["-fmad=false"]),      # BLONDIN: no FMA contraction, like the reference build
         ("sirocco_tables", "sirocco_tables.c", [])] + [   # host-only C (table readers): compiled with gcc
    ("sweeps_nv%d_bf%d" % (nv, bf), "pb200_sweeps.cu", ["-DPB_NV=%d" % nv, "-DPB_BF=%d" % bf])
    for nv in (5, 6, 7) for bf in (0, 1)]


# experiment hooks: PB200_EXTRA_DEFS="-DPB_MINBLK=4 ..." and PB200_LIB_NAME=libplutob200_v1.so build a
# variant library next to the default one (select it at run time with PB200_LIB=<path>)
EXTRA_DEFS = os.environ.get("PB200_EXTRA_DEFS", "").split()
if os.environ.get("PB200_LIB_NAME"):
    LIB = LIBDIR / os.environ["PB200_LIB_NAME"]


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        h.update((CSRC / f).read_bytes())
    h.update(" ".join(NVCC_FLAGS + EXTRA_DEFS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / (LIB.stem + ".sha256")
    dig = _digest()
    if LIB.exists() and not force and stamp.exists() and stamp.read_text() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = LIBDIR / ("obj_" + LIB.stem)
    objdir.mkdir(exist_ok=True)

    def cc(unit):
        name, src, defs = unit
        obj = objdir / (name + ".o")
        # per-unit stamp: an unchanged unit (same sources, headers and flags) is not recompiled
        ustamp = objdir / (name + ".sha256")
        udig = hashlib.sha256((dig + " ".join(defs) + src).encode()).hexdigest()
        if obj.exists() and not force and ustamp.exists() and ustamp.read_text() == udig:
            return obj
        if src.endswith(".c"):
            cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=c99", "-fPIC", "-I", str(PKG.parent / "include"),
                   "-c", str(CSRC / src), "-o", str(obj)]
        else:
            cmd = [nvcc] + NVCC_FLAGS + defs + EXTRA_DEFS + ["-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed (%s):\n%s%s" % (name, r.stdout, r.stderr))
        if verbose:
            print(" ".join(cmd))
            print(r.stderr)
        ustamp.write_text(udig)
        return obj

    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(cc, UNITS))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)]
                       + [str(o) for o in objs] + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
