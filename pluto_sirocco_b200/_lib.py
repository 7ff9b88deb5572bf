"""ctypes binding of the C ABI in include/pluto_b200.h.  No fallback: if the CUDA library
is missing or cannot be loaded this module raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ["PB200_LIB"]) if os.environ.get("PB200_LIB") else PKG / "lib" / "libplutob200.so"

# option codes (include/pluto_b200.h)
CARTESIAN, SPHERICAL = 1, 4
FLAT, LINEAR, PARABOLIC = 1, 2, 3
EULER, RK2, RK3 = 1, 2, 3
TVDLF, HLL, HLLC = 1, 2, 3
LIMITERS = dict(DEFAULT=0, FLAT_LIM=1, MINMOD_LIM=2, VANLEER_LIM=3, MC_LIM=4, VANALBADA_LIM=5,
                OSPRE_LIM=6, UMIST_LIM=7)
BC = dict(outflow=1, reflective=2, axisymmetric=3, eqtsymmetric=4, periodic=5, userdef=8, polaraxis=9,
          neighbour=100)
OK, EINVAL, ENODEV, ECUDA, ENOMEM, ENAN, ENOTSUP = 0, -1, -2, -3, -4, -5, -6


class Config(C.Structure):
    _fields_ = [("dimensions", C.c_int), ("geometry", C.c_int), ("nx", C.c_int * 3),
                ("nghost", C.c_int), ("ntracer", C.c_int), ("reconstruction", C.c_int),
                ("limiter", C.c_int), ("time_stepping", C.c_int), ("solver", C.c_int),
                ("bc", C.c_int * 6), ("gamma", C.c_double), ("small_density", C.c_double),
                ("small_pressure", C.c_double), ("xbeg", C.c_double * 3),
                ("xend", C.c_double * 3), ("device", C.c_int), ("body_force", C.c_int),
                ("char_limiting", C.c_int), ("shock_flattening", C.c_int), ("entropy_switch", C.c_int),
                ("eos", C.c_int), ("ring_average", C.c_int), ("ring_average_rec", C.c_int), ("iso_sound_speed", C.c_double)]


class LdwConfig(C.Structure):
    _fields_ = [("nangles", C.c_int), ("userdef_bc", C.c_int), ("unit_length", C.c_double),
                ("unit_velocity", C.c_double), ("unit_density", C.c_double), ("mu", C.c_double),
                ("krad", C.c_double), ("alpharad", C.c_double), ("dfloor", C.c_double), ("rho_0", C.c_double),
                ("rho_alpha", C.c_double), ("cent_mass", C.c_double), ("disk_mdot", C.c_double),
                ("lx", C.c_double), ("tx", C.c_double), ("t_iso", C.c_double)]


class StepInfo(C.Structure):
    _fields_ = [("invDt_hyp", C.c_double), ("maxMach", C.c_double),
                ("c2p_failures", C.c_ulonglong), ("gpu_ms", C.c_float), ("launches", C.c_int)]


# every symbol include/pluto_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.c_double
_PD = C.POINTER(C.c_double)
_PL = C.POINTER(C.c_long)
SYMBOLS = {
    "pb200_last_error": (C.c_char_p, []),
    "pb200_version": (C.c_int, []),
    "pb200_config_default": (None, [C.POINTER(Config)]),
    "pb200_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "pb200_destroy": (None, [_P]),
    "pb200_shape": (C.c_int, [_P, C.POINTER(C.c_int * 3), C.POINTER(C.c_int)]),
    "pb200_set_grid": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "pb200_set_geometry": (C.c_int, [_P, _P]),
    "pb200_set_grid_uniform": (C.c_int, [_P, _P]),
    "pb200_ppm_coefficients": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P]),
    "pb200_set_body_force_vector": (C.c_int, [_P, C.c_int, _P, C.c_long, C.c_long, C.c_long, C.c_long]),
    "pb200_set_body_force_potential": (C.c_int, [_P, C.c_int, _P, C.c_long, C.c_long, C.c_long, C.c_long]),
    "pb200_cooling_set_tables": (C.c_int, [_P, C.POINTER(C.c_void_p * 7)]),
    "pb200_split_source": (C.c_int, [_P, _D, _D]),
    "pb200_ldw_enable": (C.c_int, [_P, C.POINTER(LdwConfig)]),
    "pb200_ldw_set_fluxes": (C.c_int, [_P, _P, _P, _P]),
    "pb200_ldw_set_mfit": (C.c_int, [_P, C.c_int, _P, _P]),
    # include/pluto_b200_tables.h: host-side SIROCCO table readers (plain C)
    "pb200_flux_file_nangles": (C.c_int, [C.c_char_p]),
    "pb200_read_flux_file": (C.c_long, [C.c_char_p, _P, C.c_int, _P]),
    "pb200_read_mfit_file": (C.c_long, [C.c_char_p, _P, C.POINTER(C.c_int), _P, _P]),
    "pb200_read_heatcool_file": (C.c_long, [C.c_char_p, _P, _P, _P]),
    "pb200_read_prefactors_file": (C.c_long, [C.c_char_p, _P, _P]),
    "pb200_set_internal_boundary_mask": (C.c_int, [_P, _P]),
    "pb200_libm_probe": (C.c_int, [C.c_int, C.c_long, _P, _P, _P]),
    "pb200_upload_vc": (C.c_int, [_P, _P]),
    "pb200_download_vc": (C.c_int, [_P, _P]),
    "pb200_device_vc": (_P, [_P]),
    "pb200_boundary": (C.c_int, [_P]),
    "pb200_advance_step": (C.c_int, [_P, _D, C.POINTER(StepInfo)]),
    "pb200_advance_step_host": (C.c_int, [_P, _P, _D, C.POINTER(StepInfo)]),
    "pb200_set_owned_planes": (C.c_int, [_P, C.c_int, C.c_int]),
    "pb200_host_register": (C.c_int, [_P, C.c_size_t]),
    "pb200_host_unregister": (C.c_int, [_P]),
    "pb200_next_time_step": (_D, [_D, _D, _D, _D, _D]),
    "pb200_integrate": (C.c_int, [_P, C.c_int, _D, _D, _D, _D, _PD, _PD, C.POINTER(StepInfo)]),
    "pb200_halo_layout": (C.c_int, [_P, C.c_int, _PL, _PL, _PL, _PL, _PL, _PL]),
    "pb200_step_begin": (C.c_int, [_P, _D]),
    "pb200_stage_array": (_P, [_P, C.c_int]),
    "pb200_stage": (C.c_int, [_P, C.c_int]),
    "pb200_stage_boundary": (C.c_int, [_P, C.c_int]),
    "pb200_stage_begin": (C.c_int, [_P, C.c_int]),
    "pb200_stage_finish": (C.c_int, [_P, C.c_int]),
    "pb200_stage_download": (C.c_int, [_P, C.c_int, _P]),
    "pb200_stage_upload": (C.c_int, [_P, C.c_int, _P]),
    "pb200_stage_patch_u": (C.c_int, [_P, C.c_long, _P, _P]),
    "pb200_step_end": (C.c_int, [_P, C.POINTER(StepInfo)]),
    "pb200_nstages": (C.c_int, [_P]),
    "pb200_stream": (_P, [_P]),
    "pb200_set_profiling": (C.c_int, [_P, C.c_int]),
    "pb200_kernel_times": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int)]),
    # several GPUs from one host thread (csrc/pb200_multi.cu)
    "pb200_multi_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_int), C.POINTER(_P)]),
    "pb200_multi_destroy": (None, [_P]),
    "pb200_multi_ngpus": (C.c_int, [_P]),
    "pb200_multi_ctx": (_P, [_P, C.c_int]),
    "pb200_multi_slab": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pb200_multi_set_grid": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "pb200_multi_set_body_force_vector": (C.c_int, [_P, C.c_int, _P, C.c_long, C.c_long, C.c_long, C.c_long]),
    "pb200_multi_set_body_force_potential": (C.c_int, [_P, C.c_int, _P, C.c_long, C.c_long, C.c_long, C.c_long]),
    "pb200_multi_upload_vc": (C.c_int, [_P, _P]),
    "pb200_multi_download_vc": (C.c_int, [_P, _P]),
    "pb200_multi_advance_step": (C.c_int, [_P, _D, C.POINTER(StepInfo)]),
    "pb200_multi_advance_step_host": (C.c_int, [_P, _P, _D, C.POINTER(StepInfo)]),
    "pb200_multi_integrate": (C.c_int, [_P, C.c_int, _D, _D, _D, _D, _PD, _PD, C.POINTER(StepInfo)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libplutob200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            "%s not built: run `python -m pluto_sirocco_b200.build` (there is no CPU fallback)"
            % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)     # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pb200 error %d: %s" % (code, msg))
        self.code = code


def check(rc: int):
    if rc < 0:
        raise PB200Error(rc, load().pb200_last_error().decode())
    return rc
