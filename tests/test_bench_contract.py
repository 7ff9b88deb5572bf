"""CPU: the reference arm of bench.py (the reference's own executable on the host cores) prints ONE
JSON line with the keys of the bench contract; the CUDA arm refuses to run without a device."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    import refrun
    if not refrun.have_ref("sedov3d"):
        pytest.skip("oracle/_ref/sedov3d/pluto not built (needs /root/reference)")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-size", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mzones/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the reference arm runs a bounded sample of THIS arm's workload and carries this arm's config dict
    import argparse
    sys.path.insert(0, str(ROOT))
    import bench
    mine = bench.cart_config(argparse.Namespace(size=512, workload="sedov", recon="LINEAR", rk="RK2", solver="hllc",
                                                state="sedov"), 1, 512 ** 3)
    assert d["config"] == mine
    assert "slab" in d["cpu_baseline"]["sample"] or "16 x 16 x 16" in d["cpu_baseline"]["sample"]


def test_cuda_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
