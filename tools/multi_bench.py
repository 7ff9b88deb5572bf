#!/usr/bin/env python
"""Throughput of the C-side slab decomposition (pb200_multi_*: ONE host thread, N GPUs, NCCL inside the library):
resident steps and the host-buffer call, weak scaling with n^3 zones per GPU stacked along x3.
   python tools/multi_bench.py --ngpus 8 --size 512 [--host-size 256]"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
import numpy as np
import torch
import bench
from pluto_sirocco_b200 import MultiHydro

ap = argparse.ArgumentParser()
ap.add_argument("--ngpus", type=int, default=2)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--host-size", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--recon", default="LINEAR")
ap.add_argument("--rk", default="RK2")
a = ap.parse_args()
out = {"ngpus": a.ngpus}
for tag, n, host in (("resident", a.size, False), ("host", a.host_size, True)):
    if n <= 0:
        continue
    N = a.ngpus
    m = MultiHydro(N, dimensions=3, nx=(n, n, n * N), xend=(1.0, 1.0, float(N)), gamma=1.4, reconstruction=a.recon,
                   time_stepping=a.rk, bcs=bench.SEDOV_BCS)
    pin = torch.empty(m.shape, dtype=torch.float64, pin_memory=True)
    vc = pin.numpy()
    vc[:] = 1.0; vc[1:4] = 0.0
    vc[m.interior()] = bench.sedov_block((n, n, n * N), 0, n)
    m.upload(vc)
    dt = 1e-9
    step = (lambda: m.advance_step_host(vc, dt)) if host else (lambda: m.advance_step(dt))
    for _ in range(3):
        info = step(); dt = min(0.3 / info.invDt_hyp, 1.1 * dt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = []
    for _ in range(a.steps):
        t1 = time.perf_counter()
        info = step(); dt = min(0.3 / info.invDt_hyp, 1.1 * dt)
        per.append(round(1e3 * (time.perf_counter() - t1), 2))
    # every call returns after its own stream synchronisation (pb200_step_end): the sum of the per-call times is the
    # wall time of the loop.  (torch.cuda.synchronize(d) on a device torch has not touched yet initialises a context
    # there - 7 ms each - which an earlier version of this script charged to the steps.)
    wall = sum(per) * 1e-3
    out[tag] = {"zones_per_gpu": n ** 3, "ms_per_step": 1e3 * wall / a.steps, "Mzones_per_s": n ** 3 * N * a.steps / wall / 1e6,
                "gpu_ms_max": info.gpu_ms, "launches_per_step": info.launches, "per_step_ms": per}
    m.close(); del pin
print(json.dumps(out))
