import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np
from oracle import Oracle
from pluto_sirocco_b200 import Hydro
from pluto_sirocco_b200._lib import PB200Error
kw = dict(dimensions=1, nx=(64, 1, 1), gamma=1.4, bcs=("outflow",) * 6)
h, o = Hydro(**kw), Oracle(**kw)
v = np.zeros((5, 1, 1, 64)); v[0] = 1.0; v[4] = 1e-9
v[1, ..., :32] = 30.0; v[1, ..., 32:] = -30.0
v[1, ..., 20:24] = -25.0
vc = o.embed(v); h.set_interior(v)
np.set_printoptions(linewidth=200, precision=6)
for n in range(6):
    inv, mach, nf = o.advance_step(vc, 1e-3)
    try:
        info = h.advance_step(1e-3)
        print(n, "ok", info.c2p_failures, nf, info.invDt_hyp, inv)
    except PB200Error as e:
        print(n, "ERR", e)
    g = h.get_interior()
    bad = np.argwhere(~np.isfinite(g))
    print(" nonfinite at", bad[:10].tolist())
    for b in bad[:3]:
        i=b[3]; print("  gpu", g[:,0,0,i], "orc", vc[o.interior()][:,0,0,i])
    d=np.abs(g-vc[o.interior()]); print(" maxdiff", np.nanmax(d))
    h.set_interior(vc[o.interior()])
