"""rel. L1 of the drop-in against the stock reference executable as a function of tstop (free running)."""
import sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")]
import numpy as np
import refrun
from common import rel_l1
from test_dropin_gpu import RT, KH
cases = {
    "rt3d_vec": (dict(shape=(16, 64, 24), nvar=6, grid=[(-0.5, 24, 0.5), (-1.0, 64, 1.0), (-0.5, 16, 0.5)], **RT), [1, 2, 3, 4, 5, 6]),
    "kh3d": (dict(shape=(12, 48, 48), nvar=6, grid=[(0.0, 48, 1.0), (-0.5, 48, 0.5), (0.0, 12, 0.25)], **KH), [0.5, 1, 1.5, 2, 3]),
}
for cfg, (kw, ts) in cases.items():
    exe = ROOT / "integration" / "_build" / cfg / "pluto_b200"
    for t in ts:
        k = dict(kw, tstop=float(t))
        with tempfile.TemporaryDirectory() as wd:
            ref = refrun.run(cfg, wd + "/r", solver="hllc", dbl=(1000.0, -1), timeout=900, **k)
            got = refrun.run(cfg, wd + "/g", solver="hllc", dbl=(1000.0, -1), exe=exe, env={"PB200_RESIDENT": "1"}, timeout=900, **k)
        a, b = ref["data"][-1], got["data"][-1]
        print(cfg, "tstop", t, "steps", ref["steps"][-1][0], got["steps"][-1][0], "relL1 %.3e" % rel_l1(b, a),
              "max|vz| ref %.3e" % np.abs(a[3]).max(), "rho range", a[0].min(), a[0].max(), flush=True)
