"""Debug helper: per-variable error map of one general-grid fixture on the CUDA path."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from common import *
from pluto_sirocco_b200 import Hydro
name = sys.argv[1] if len(sys.argv) > 1 else "ppmg_kh3d_stretched"
g = load_golden(name)
kw = gen_kwargs_from_golden(g)
hk = hydro_kwargs_from_gen(kw)
h = Hydro(**hk)
set_point_mass_gravity(h, float(g["gm"]))
data, steps = g["data"], g["steps"]
h.set_interior(data[0])
h.advance_step(steps[0, 2])
got = h.get_interior()
for nv in range(got.shape[0]):
    e = np.abs(got[nv] - data[1][nv])
    k = np.unravel_index(np.argmax(e), e.shape)
    print(nv, e.max(), k, got[nv][k], data[1][nv][k], "nbad", int((e > 1e-12 * (1 + np.abs(data[1][nv]))).sum()), "of", e.size)
e = np.abs(got - data[1]).max(axis=0)
bad = e > 1e-11
print("bad k:", np.where(bad.any(axis=(1, 2)))[0], "j:", np.where(bad.any(axis=(0, 2)))[0], "i:", np.where(bad.any(axis=(0, 1)))[0])
