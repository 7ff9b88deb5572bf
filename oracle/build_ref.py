#!/usr/bin/env python3
"""Build the UNMODIFIED reference (PLUTO 4.4-patch3 + sirocco fork) for the hot-path
configurations, straight from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
product path (pluto_sirocco_b200/); only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may use it, and only as the checker / the CPU baseline.

This is our own recipe (plain gcc on an explicit source list) - the reference's
setup.py / curses menu / generated makefile are not run.  The source list mirrors what
Tools/Python/define_problem.py:533-711 + Src/Templates/makefile + Src/HD/makefile +
Src/EOS/Ideal/makefile would select for PHYSICS=HD on a static grid.  The only files
written are under oracle/_ref/<config>/ (git-ignored; they DO travel to the GPU box):
  definitions.h   problem header = the reference's own definitions_NN.h with the
                  overrides listed in CONFIGS (e.g. RECONSTRUCTION PARABOLIC)
  obj/*.o, pluto  the reference executable for that definitions.h
No reference source is copied into the repository.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("PLUTO_DIR", "/root/reference"))
OUT = HERE / "_ref"

# Flags of Config/Linux.gcc.defs:8-9 (the serial gcc configuration of the reference).
CFLAGS = ["-c", "-O3", "-std=c17"]
LDFLAGS = ["-lm"]

# Src/Templates/makefile OBJ lists (static grid, serial).
CORE = """adv_flux arrays array_reconstruct boundary check_states cmd_line_opt debug_tools
entropy_switch failsafe flag_shock flatten fluid_interface_boundary get_nghost
int_bound_reset input_data mappers3D mean_mol_weight parse_file plm_coeffs rbox
reconstruct rotate set_indexes set_geometry set_output tools var_names
bin_io colortable initialize jet_domain main output_log restart ring_average
runtime_setup set_image show_config set_grid startup split_source
write_data write_tab write_img write_vtk write_vtk_proc""".split()
MATH = """math_interp math_lu_decomp math_qr_decomp math_misc math_ode math_quadrature
math_random math_root_finders math_table2D""".split()
HD = """advection_solver ausm eigenv fluxes mappers mappers_loc hll_speed hll hllc
set_solver tvdlf two_shock roe prim_eqn rhs rhs_source""".split()
EOS_IDEAL = ["eos"]
RK = ["rk_step", "update_stage"]

# search path for sources, in make-VPATH order (problem directory first)
def vpath(problem_dir: Path, extra: list[str], eos: str = "Ideal") -> list[Path]:
    s = REF / "Src"
    return [problem_dir, s, s / "Math_Tools", s / "HD", s / "MHD", s / "EOS" / eos,
            s / "States", s / "Time_Stepping"] + [s / e for e in extra]


CONFIGS = {
    # C1: Sod tube, PLM + RK2 (solver chosen in pluto.ini)  Test_Problems/HD/Sod conf 01
    "sod": dict(problem="HD/Sod", defs="definitions_01.h", overrides={}, states="plm"),
    # same problem with PPM + RK3
    "sod_ppm": dict(problem="HD/Sod", defs="definitions_01.h",
                    overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"},
                    states="ppm"),
    # C2: Sedov 3-D Cartesian PLM + RK2   Test_Problems/HD/Sedov conf 05
    "sedov3d": dict(problem="HD/Sedov", defs="definitions_05.h", overrides={}, states="plm"),
    # C5: Sedov 3-D PPM + RK3
    "sedov3d_ppm": dict(problem="HD/Sedov", defs="definitions_05.h",
                        overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"},
                        states="ppm"),
    # Sedov 2-D Cartesian PLM + RK2   (conf 05 with DIMENSIONS 2)
    "sedov2d": dict(problem="HD/Sedov", defs="definitions_05.h", overrides={"DIMENSIONS": "2"},
                    states="plm"),
    # C3: Rayleigh-Taylor with a tracer; user files of THIS repository (oracle/problems/rt:
    # the reference's Rayleigh_Taylor conf 04 is CHAR_LIMITING + AMR oriented and has no tracer)
    "rt3d_vec": dict(local="rt", overrides={}, states="plm"),
    "rt3d_pot": dict(local="rt", overrides={"BODY_FORCE": "POTENTIAL"}, states="plm"),
    "rt2d_vec": dict(local="rt", overrides={"DIMENSIONS": "2"}, states="plm"),
    "rt2d_pot": dict(local="rt", overrides={"DIMENSIONS": "2", "BODY_FORCE": "POTENTIAL",
                                            "LIMITER": "MC_LIM"}, states="plm"),
    "rt1d_vec": dict(local="rt", overrides={"DIMENSIONS": "1"}, states="plm"),
    # C4 building blocks: spherical geometry on stretched grids (oracle/problems/sph)
    "sph2d": dict(local="sph", overrides={}, states="plm"),
    "sph2d_char": dict(local="sph", overrides={"CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM"}, states="plm"),
    "sph2d_flat": dict(local="sph", overrides={"CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM",
                                               "SHOCK_FLATTENING": "MULTID"}, states="plm"),
    "sph2d_entr": dict(local="sph", overrides={"CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM",
                                               "SHOCK_FLATTENING": "MULTID", "ENTROPY_SWITCH": "ALWAYS"},
                       states="plm"),
    "sph2d_sel": dict(local="sph", overrides={"ENTROPY_SWITCH": "SELECTIVE"}, states="plm"),
    "sph2d_oned": dict(local="sph", overrides={"SHOCK_FLATTENING": "ONED"}, states="plm"),
    "sph2d_char_oned": dict(local="sph", overrides={"CHAR_LIMITING": "YES", "LIMITER": "MC_LIM",
                                                    "SHOCK_FLATTENING": "ONED"}, states="plm"),
    "iso2d_oned": dict(local="iso", overrides={"SHOCK_FLATTENING": "ONED"}, states="plm"),
    "sph1d": dict(local="sph", overrides={"DIMENSIONS": "1"}, states="plm"),
    "sph3d": dict(local="sph", overrides={"DIMENSIONS": "3"}, states="plm"),
    # cylindrical (r, z) and polar (r, phi[, z]) geometry (oracle/problems/cyl)
    "cyl2d": dict(local="cyl", overrides={}, states="plm"),
    "cyl2d_nobf": dict(local="cyl", overrides={"BODY_FORCE": "NO"}, states="plm"),
    "cyl2d_flat": dict(local="cyl", overrides={"CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM",
                                               "SHOCK_FLATTENING": "MULTID"}, states="plm"),
    "pol2d": dict(local="cyl", overrides={"GEOMETRY": "POLAR"}, states="plm"),
    "pol3d": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "DIMENSIONS": "3"}, states="plm"),
    # PPM + RK3 on general grids (stretched Cartesian, cylindrical, polar)
    "kh3d_ppm": dict(local="kh", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"}, states="ppm"),
    "cyl2d_ppm": dict(local="cyl", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"}, states="ppm"),
    "cyl2d_ppm_flat": dict(local="cyl", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3",
                                                   "SHOCK_FLATTENING": "MULTID"}, states="ppm"),
    "pol2d_ppm": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "RECONSTRUCTION": "PARABOLIC",
                                              "TIME_STEPPING": "RK3"}, states="ppm"),
    "sph2d_pot": dict(local="sph", overrides={"BODY_FORCE": "POTENTIAL"}, states="plm"),
    "sph3d_pot": dict(local="sph", overrides={"DIMENSIONS": "3", "BODY_FORCE": "(VECTOR+POTENTIAL)"}, states="plm"),
    "pol2d_pot": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "BODY_FORCE": "POTENTIAL"}, states="plm"),
    "sph2d_ppm_char": dict(local="sph", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3",
                                                   "CHAR_LIMITING": "YES", "SHOCK_FLATTENING": "MULTID"}, states="ppm"),
    "iso2d_ppm_char": dict(local="iso", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3",
                                                   "CHAR_LIMITING": "YES"}, states="ppm"),
    "sph2d_ppm": dict(local="sph", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"}, states="ppm"),
    "sph3d_ppm": dict(local="sph", overrides={"DIMENSIONS": "3", "RECONSTRUCTION": "PARABOLIC",
                                              "TIME_STEPPING": "RK3"}, states="ppm"),
    # EOS ISOTHERMAL (oracle/problems/iso): Cartesian 2-D / 3-D, spherical 2-D with gravity
    "iso2d": dict(local="iso", overrides={}, states="plm"),
    "iso2d_flat": dict(local="iso", overrides={"CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM",
                                               "SHOCK_FLATTENING": "MULTID"}, states="plm"),
    "iso3d": dict(local="iso", overrides={"DIMENSIONS": "3"}, states="plm"),
    "iso_sph2d": dict(local="iso", overrides={"GEOMETRY": "SPHERICAL", "BODY_FORCE": "VECTOR",
                                              "CHAR_LIMITING": "YES", "LIMITER": "VANLEER_LIM",
                                              "SHOCK_FLATTENING": "MULTID"}, states="plm"),
    # C4: the line-driven disc wind of the sirocco coupling, UNMODIFIED user files of the reference
    # (Test_Problems/LineDrivenWind/cv_idl: init.c, definitions.h, userdef_output.c)
    "ldw": dict(problem="LineDrivenWind/cv_idl", defs="definitions.h", overrides={}, states="plm",
                extra_vpath=["Cooling/BLONDIN", "LineDriven"], extra_objs=["cooling", "line_connect"]),
    # the same without the BLONDIN source step: isolates the hydro + line-force update
    "ldw_nocool": dict(problem="LineDrivenWind/cv_idl", defs="definitions.h", overrides={"COOLING": "NO"},
                       states="plm", extra_vpath=["LineDriven"], extra_objs=["line_connect"]),
    # the isothermal twin of the line-driven wind problem, UNMODIFIED user files of the reference
    # (Test_Problems/LineDrivenWind/cv_iso: EOS ISOTHERMAL, COOLING NO)
    "ldw_iso": dict(problem="LineDrivenWind/cv_iso", defs="definitions.h", overrides={}, states="plm",
                    extra_vpath=["LineDriven"], extra_objs=["line_connect"]),
    # an arbitrary, time-dependent UserDefBoundary() (oracle/problems/jet): the shim's host-boundary mode
    "jet2d": dict(local="jet", overrides={}, states="plm"),
    "jet2d_ppm": dict(local="jet", overrides={"RECONSTRUCTION": "PARABOLIC", "TIME_STEPPING": "RK3"}, states="ppm"),
    # INTERNAL_BOUNDARY YES with FLAG_INTERNAL_BOUNDARY zones (oracle/problems/tunnel): InternalBoundaryReset()
    "tunnel2d": dict(local="tunnel", overrides={}, states="plm"),
    "tunnel3d_ppm": dict(local="tunnel", overrides={"DIMENSIONS": "3", "RECONSTRUCTION": "PARABOLIC",
                                                    "TIME_STEPPING": "RK3"}, states="ppm"),
    # C3: Kelvin-Helmholtz shear layer with a tracer (oracle/problems/kh)
    "kh3d": dict(local="kh", overrides={}, states="plm"),
    # RING_AVERAGE (Src/ring_average.c) with the POLARAXIS boundary: polar (r, phi[, z]) from r = 0, spherical from theta = 0
    "pol2d_ring": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "BODY_FORCE": "NO", "RING_AVERAGE": "8"}, states="plm"),
    "pol2d_ring_vl": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "BODY_FORCE": "NO", "RING_AVERAGE": "8",
                                                  "RING_AVERAGE_REC": "2"}, states="plm"),
    "pol3d_ring": dict(local="cyl", overrides={"GEOMETRY": "POLAR", "DIMENSIONS": "3", "BODY_FORCE": "NO",
                                               "RING_AVERAGE": "4"}, states="plm"),
    "sph3d_ring": dict(local="sph", overrides={"DIMENSIONS": "3", "RING_AVERAGE": "4", "PHI_PERTURB": "YES"}, states="plm"),
    # the SAME reference sources compiled with FMA contraction (-mfma -ffp-contract=fast): a second, equally
    # valid rounding of the reference.  Long free-running comparisons use |ref_fma - ref| as the yardstick of
    # how far round-off differences are amplified by the flow itself (tests/test_dropin_gpu.py).
    "rt3d_vec_fma": dict(local="rt", overrides={}, states="plm", cflags=["-mfma", "-ffp-contract=fast"]),
    "kh3d_fma": dict(local="kh", overrides={}, states="plm", cflags=["-mfma", "-ffp-contract=fast"]),
}


def patch_definitions(text: str, overrides: dict[str, str]) -> str:
    for key, val in overrides.items():
        pat = re.compile(r"^(#define\s+%s\s+)\S+" % re.escape(key), re.M)
        if pat.search(text):
            text = pat.sub(lambda m: m.group(1) + val, text)
        else:
            text = "#define  %s  %s\n" % (key, val) + text
    return text


def find_source(name: str, paths: list[Path]) -> Path:
    for p in paths:
        f = p / (name + ".c")
        if f.exists():
            return f
    raise FileNotFoundError(name)


def build(cfg_name: str, force: bool = False, verbose: bool = False) -> Path:
    cfg = CONFIGS[cfg_name]
    if "local" in cfg:   # user files written for this repository (init.c, definitions.h)
        problem_dir = HERE / "problems" / cfg["local"]
        cfg = dict(cfg, defs="definitions.h")
    else:
        problem_dir = REF / "Test_Problems" / cfg["problem"]
    wd = OUT / cfg_name
    exe = wd / "pluto"
    if exe.exists() and not force:
        return exe
    if not REF.exists():
        raise RuntimeError("reference tree %s not present; cannot build oracle/_ref" % REF)
    (wd / "obj").mkdir(parents=True, exist_ok=True)
    defs = patch_definitions((problem_dir / cfg["defs"]).read_text(), cfg["overrides"])
    (wd / "definitions.h").write_text(defs)

    extra = cfg.get("extra_vpath", [])
    eos = "Isothermal" if re.search(r"^#define\s+EOS\s+ISOTHERMAL", defs, re.M) else "Ideal"   # makefile: EOS directory
    paths = vpath(problem_dir, extra, eos)
    names = CORE + MATH + HD + EOS_IDEAL + RK + ["init"]
    names += ["plm_states"] if cfg["states"] == "plm" else ["ppm_states", "ppm_coeffs"]
    names += cfg.get("extra_objs", [])
    if not (problem_dir / "userdef_output.c").exists():
        names.append("userdef_output")   # Src/userdef_output.c template
    else:
        names.append("userdef_output")
    s = REF / "Src"
    incs = ["-I%s" % wd, "-I%s" % s, "-I%s" % (s / "HD"), "-I%s" % (s / "EOS" / eos),
            "-I%s" % (s / "States"), "-I%s" % (s / "Math_Tools")]
    incs += ["-I%s" % (s / e) for e in extra]

    def cc(name: str):
        src = find_source(name, paths)
        obj = wd / "obj" / (name + ".o")
        cmd = ["gcc"] + CFLAGS + cfg.get("cflags", []) + incs + [str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        return obj

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, names))
    r = subprocess.run(["gcc"] + [str(o) for o in objs] + LDFLAGS + ["-o", str(exe)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    if verbose:
        print("built", exe)
    return exe


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=list(CONFIGS))
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    for c in a.configs:
        build(c, force=a.force, verbose=True)


if __name__ == "__main__":
    main()
