import sys
sys.path[:0]=['.','oracle','tests']
import numpy as np, bench
from gen_oracle import GenOracle
from pluto_sirocco_b200 import Hydro
from common import LDW_BCS, LDW_PARAMS, LDW_UNITS, ldw_flux_tables, hydro_kwargs_from_gen
n1, n2 = 1024, 512
grid = [(0.87, n1, 8.7, "r", 1.005), (0.0, n2, float(np.radians(90.0)), "r", 0.995), (0.0, 1, 1.0)]
kw = dict(dimensions=2, grid=grid, geometry="SPHERICAL", gamma=5. / 3., time_stepping="RK2", solver="hll",
          limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1, char_limiting=True,
          shock_flattening=True, entropy_switch=True, nghost=3)
o = GenOracle(**kw); h = Hydro(**hydro_kwargs_from_gen(kw))
x1, x2 = o.x(0), o.x(1)
gm_code = 6.6726e-8 * LDW_PARAMS["CENT_MASS"] / (LDW_UNITS["length"] * LDW_UNITS["velocity"] ** 2)
fr, ft, fp = ldw_flux_tables(x1, x2, roundtrip=False)
for obj in (o, h):
    obj.set_body_force_vector(0, (-gm_code / (x1 * x1)).reshape(1, 1, -1)); obj.set_body_force_vector(1, np.zeros((1, 1, 1))); obj.set_body_force_vector(2, np.zeros((1, 1, 1)))
    obj.set_ldw(params=LDW_PARAMS, units=LDW_UNITS, flux_r=fr, flux_t=ft, flux_p=fp)
v = bench.ldw_state(x1[3:-3], x2[3:-3], LDW_PARAMS, LDW_UNITS)
vc = o.embed(v); h.set_interior(v)
dt = 1e-4
for n in range(4):
    inv, mach, nf = o.advance_step(vc, dt)
    info = h.advance_step(dt)
    got, ref = h.get_interior(), vc[o.interior()]
    for nv in range(6):
        sc = np.abs(ref[1:4]).max() if 1<=nv<=3 else np.abs(ref[nv]).max()
        d = np.abs(got[nv]-ref[nv]); w=np.unravel_index(np.argmax(d), d.shape)
        print(n, nv, "max abs/scale %.3e"%(d.max()/sc), "at", w, "got", got[nv][w], "ref", ref[nv][w], "n>1e-12:", int((d>1e-12*sc).sum()))
    print(n, "invdt", info.invDt_hyp, inv)
    dt = min(0.4/inv, 1.1*dt)
    h.set_interior(ref)
