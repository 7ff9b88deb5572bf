"""pluto_sirocco_b200 -- B200-native (sm_100a CUDA) implementation of PLUTO's unsplit HD
update as built in the sirocco-coupled fork, behind the reference's own interface.

Only the hot path lives here: csrc/ (CUDA kernels + C ABI, include/pluto_b200.h) and the
host-side mirror of the reference interface (hydro.py).  There is no CPU fallback."""
from . import _lib, tables
from .hydro import Definitions, Hydro, MultiHydro, Runtime, Simulation, make_grid

__all__ = ["Definitions", "Hydro", "MultiHydro", "Runtime", "Simulation", "make_grid", "_lib", "tables"]
__version__ = "0.1.0"
