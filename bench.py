#!/usr/bin/env python3
"""bench.py -- zone-updates/s of the unsplit HD update (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU build

A "step" is one full time step (all RK stages, all directions, boundaries, cons<->prim, dt
reduction) of the workload `sedov3d-512^3 PLM+HLLC+RK2` per GPU (BASELINE.json configs[1]);
with N > 1 the blocks are stacked along x3 (weak scaling, slab decomposition, NCCL halo
exchange per stage + one max-allreduce per step).  State arrays (5.5 GB each) are far larger
than the 126 MB L2, so no explicit L2 flush is needed between iterations.

value : device-resident state (inputs in HBM when the timed region starts)
e2e   : the same steps through the host-buffer entry point pb200_advance_step_host(): H2D of
        d->Vc from pinned host memory, AdvanceStep, D2H of d->Vc, every step.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "zone-updates/s"
UNIT = "Mzones/s"
ALG_BYTES_RK2 = 360.0     # FP64, NVAR=5: 160 B (stage 1) + 200 B (stage 2)
ALG_BYTES_RK3 = 560.0
SEDOV_BCS = ("reflective", "outflow") * 3


def workload_name(n, recon, rk, solver, rt=False):
    if rt:
        return "rayleigh-taylor3d-%d^3-global %s+%s+%s, tracer, gravity" % (
            n, {"LINEAR": "PLM", "PARABOLIC": "PPM"}[recon], solver.upper(), rk)
    return "sedov3d-%d^3-per-gpu %s+%s+%s" % (n, {"LINEAR": "PLM", "PARABOLIC": "PPM"}[recon], solver.upper(), rk)


def sedov_block(nx, zoff, nglob, gamma=1.4):
    """Sedov IC of Test_Problems/HD/Sedov/init.c:49-91 (INITIAL_SMOOTHING NO) for the block
    whose x3 index starts at zoff; unit cube spacing 1/nglob in every direction."""
    import numpy as np
    n1, n2, n3 = nx
    x = (np.arange(n1) + 0.5) / nglob
    y = (np.arange(n2) + 0.5) / nglob
    z = (np.arange(n3) + zoff + 0.5) / nglob
    dr = 3.5 / nglob
    vol = 4.0 / 3.0 * np.pi * dr ** 3
    v = np.zeros((5, n3, n2, n1))
    v[0] = 1.0
    v[4] = 1.0e-5
    # the deposition sphere touches only the first few zones
    m = 8
    r = np.sqrt(x[None, None, :m] ** 2 + y[None, :m, None] ** 2 + z[:m, None, None] ** 2)
    v[4, :m, :m, :m] = np.where(r <= dr, (gamma - 1.0) * 1.0 / vol, 1.0e-5)
    return v


def rt_block(nx, zoff, nglob, gamma=5. / 3., eta=2.0, grav=-0.1):
    """Rayleigh-Taylor IC of oracle/problems/rt/init.c (3-D branch) for the block whose x3 index
    starts at zoff: heavy fluid (eta) above light fluid in gravity along x2, hydrostatic pressure,
    single-mode velocity seed, tracer = heavy fluid.  Unit cube centred on the origin."""
    import numpy as np
    n1, n2, n3 = nx
    x = -0.5 + (np.arange(n1) + 0.5) / nglob
    y = -0.5 + (np.arange(n2) + 0.5) / nglob
    z = -0.5 + (np.arange(n3) + zoff + 0.5) / nglob
    v = np.zeros((6, n3, n2, n1))
    heavy = (y >= 0.0).astype(float)[None, :, None]
    v[0] = 1.0 + (eta - 1.0) * heavy
    v[4] = 1.0 / gamma + v[0] * grav * y[None, :, None]
    seed = (1.0 + np.cos(2.0 * np.pi * x))[None, None, :] * (1.0 + np.cos(2.0 * np.pi * z))[:, None, None] * 0.5
    v[2] = -1.e-2 * seed * np.exp(-y * y * 50.0)[None, :, None]
    v[5] = heavy
    return v


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore the pinned host buffers it allocates next: first touch) to the CPUs of the NUMA
    node the GPU hangs off.  With the buffers of all ranks on one socket every H2D / D2H copy of the other socket's
    GPUs crosses the inter-socket link, which caps the host-buffer (e2e) rate of an 8-GPU job far below 8 PCIe links.
    Returns a short description for the JSON line (None when the topology cannot be read)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "numa_node unknown for %s" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "node %d has no CPU this process may use" % node
        os.sched_setaffinity(0, allowed)
        return "GPU %d (%s) -> NUMA node %d, %d CPUs" % (index, bdf, node, len(allowed))
    except Exception as e:      # containers without sysfs, old torch: run unbound
        return "unbound (%s)" % type(e).__name__


def busy_block(nx, seed):
    """"busy" state: the seeded waves + jumps of tests/common.py random_state (shocks, contacts, every limiter and
    solver branch active in every zone neighbourhood), generated on a block of <= 128 zones per direction and
    tiled periodically over the grid."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from common import random_state
    n1, n2, n3 = nx
    b = [min(m, 128) for m in (n3, n2, n1)]
    v = random_state(tuple(b), seed=seed, smooth=False)
    reps = [1] + [-(-m // q) for m, q in zip((n3, n2, n1), b)]
    return np.tile(v, reps)[:, :n3, :n2, :n1]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def count(self, t0=None, t1=None):
        return sum(1 for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1) and len(r.split(",")) >= 7)

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons of the samples taken under load; the count of samples that
        fell inside [t0, t1] (the timed region) is reported separately."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for (_, r) in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": self.count(t0, t1) if t0 is not None else None}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------
#  reference arm: the unmodified reference C build on the host cores
# --------------------------------------------------------------------------------------
def reference_rate(cfg, shape, nproc, warm, steps, solver="hllc"):
    """Mzones/s of `nproc` concurrent serial reference processes (no MPI on this image, so this
    is the communication-free upper bound of an MPI run), each on a block of shape = (nx1, nx2, nx3)
    zones of the Sedov problem.  Timed by differencing a (warm) and a (warm+steps) run so
    initialisation is excluded."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import refrun
    exe = refrun.REFDIR / cfg / "pluto"
    if not exe.exists():
        return None
    n1, n2, n3 = shape

    def launch(maxsteps):
        procs, dirs = [], []
        for p in range(nproc):
            d = tempfile.mkdtemp(prefix="plref_")
            dirs.append(d)
            refrun.write_ini(Path(d) / "pluto.ini", grid=[(0, n1, 1), (0, n2, 1), (0, n3, n3 / float(n1))], cfl=0.3, tstop=0.5,
                             first_dt=1e-9, solver=solver, bcs=SEDOV_BCS, params=dict(ENRG0=1.0, DNST0=1.0, GAMMA=1.4))
            procs.append(subprocess.Popen([str(exe), "-no-write", "-maxsteps", str(maxsteps)], cwd=d,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        t0 = time.perf_counter()
        for pr in procs:
            pr.wait()
        t = time.perf_counter() - t0
        import shutil
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)
        return t

    ta = launch(warm)
    tb = launch(warm + steps)
    dt = tb - ta
    if dt < 0.05 * tb:      # too short to difference reliably: charge the whole second run to its steps
        dt = tb * steps / float(warm + steps)
    return nproc * n1 * n2 * n3 * steps / dt / 1e6, dt


def cart_config(args, world, zones_local, nvar=5):
    """`config` of the Cartesian workloads: the SAME dict in both arms (the reference arm times a bounded
    sample of this workload and says which in cpu_baseline.sample)."""
    n, rt = args.size, args.workload == "rt"
    tot = (n + 4) ** 3 if args.recon == "LINEAR" else (n + 6) ** 3
    return {"workload": workload_name(n, args.recon, args.rk, args.solver, rt), "state": "rt" if rt else args.state,
            "zones_per_gpu": zones_local, "decomposition": "x3 slabs" if world > 1 else "none",
            "l2": "inputs (%.1f GB per state array) larger than L2, no flush needed" % (nvar * tot * 8 / 1e9),
            "boundaries": "periodic x1/x3, reflective x2" if rt else "reflective-beg/outflow-end", "cfl": 0.4 if rt else 0.3}


def run_reference(args):
    """The reference's own CPU implementation of the path (the unmodified executable, oracle/_ref) on all host
    cores, on THIS arm's workload: the n^3-per-GPU Sedov grid is cut into x3 slabs, one serial process per core
    (what its MPI build would do minus the communication); a step of the sample covers min(n^3, budget) zones."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    cfg = "sedov3d" if args.recon == "LINEAR" else "sedov3d_ppm"
    n = args.size
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # bound the run to ~2.5 minutes: K+W steps at ~2.5 Mzones/s per core
    nsteps = args.steps + max(args.warmup, 1)
    budget = int(150.0 * 2.5e6 / nsteps)                       # zones per process and step
    nproc = min(cores, max(1, n // 8))
    planes = max(4, min(n // nproc, budget // (n * n)))
    if args.ref_size:                                           # explicit cubic sample (tests)
        shape = (args.ref_size,) * 3
    else:
        shape = (n, n, planes)
    res = reference_rate(cfg, shape, nproc, max(args.warmup, 1), args.steps, args.solver)
    if res is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/%s/pluto not built on this box" % cfg})
        return 0
    rate, dt = res
    frac = nproc * shape[0] * shape[1] * shape[2] / float(n ** 3 * world)
    sample = ("%d concurrent serial processes of the unmodified reference executable (no MPI on this image: the "
              "communication-free upper bound of its MPI build), each on a %d x %d x %d slab of the workload's grid "
              "(%.0f %% of the %d^3 x %d-GPU zones per step), same problem / solver / reconstruction, %d timed steps "
              "(initialisation excluded by differencing two runs); gcc -O3 -std=c17 (Config/Linux.gcc.defs)"
              % (nproc, shape[0], shape[1], shape[2], 100.0 * frac, n, world, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cart_config(args, world, n ** 3),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
#  short measurements of the other BASELINE configs (reported under "secondary")
# --------------------------------------------------------------------------------------
def _quick_cart(*, workload, size, recon, rk, state, steps, warmup=3, solver="hllc"):
    """Device-resident zone-updates/s of a Cartesian workload on the ranks of this job (all ranks call this).
    Same slab machinery and timing rules as the headline: W warm-up steps, K steps between CUDA events on the
    launching stream, barrier + synchronize on both sides, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pluto_sirocco_b200 import Hydro
    from pluto_sirocco_b200.slab import Slab, SlabHydro
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, rt = size, workload == "rt"
    if rt:
        gnx, bcs = (n, n, n), ("periodic", "periodic", "reflective", "reflective", "periodic", "periodic")
        gxb, gxe, gamma, ntr, bf = (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 5. / 3., 1, 1
        zones_local = n * n * (n // world)
    else:
        gnx, bcs = (n, n, n * world), SEDOV_BCS
        gxb, gxe, gamma, ntr, bf = (0., 0., 0.), (1., 1., float(world)), 1.4, 0, 0
        zones_local = n ** 3
    slab = Slab(rank, world, 3, gnx, gxb, gxe, bcs)
    xb, xe = slab.local_extent()
    h = Hydro(dimensions=3, nx=slab.local_nx(), xbeg=xb, xend=xe, gamma=gamma, reconstruction=recon, time_stepping=rk,
              solver=solver, bcs=slab.local_bcs(), device=local_rank, dx=slab.global_dx(), ntracer=ntr, body_force=bf)
    if rt:
        for comp, val in enumerate((0.0, -0.1, 0.0)):
            h.set_body_force_vector(comp, np.full((1, 1, 1), val))
    sh = SlabHydro(h, slab)
    vc = np.ones(h.shape)
    vc[1:4] = 0.0
    if rt:
        vc[h.interior()] = rt_block(slab.local_nx(), slab.offset, n)
    elif state == "sedov":
        vc[h.interior()] = sedov_block(slab.local_nx(), slab.offset, n)
    else:
        vc[h.interior()] = busy_block(slab.local_nx(), rank)
    h.upload(vc)
    del vc
    cfl, cmv, first_dt = (0.4, 1.1, 1e-3) if rt else (0.3, 1.1, 1e-9)
    g = {"dt": first_dt if (state == "sedov" or rt) else 1e-5}

    def one_step():
        inv, mach, info = sh.advance_step(g["dt"])
        g["dt"] = h.next_time_step(inv, cfl, cmv, g["dt"], first_dt)
        return info

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, warmup)):
        one_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    launches = 0
    ev0.record(sh.stream)
    for _ in range(steps):
        launches += one_step().launches
    ev1.record(sh.stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    nvar = h.nvar
    h.close()
    torch.cuda.empty_cache()
    peak, _ = measured_peaks()
    alg = (ALG_BYTES_RK2 if rk == "RK2" else ALG_BYTES_RK3) * nvar / 5.0
    step_ms = ms / steps
    return {"workload": workload_name(n, recon, rk, solver, rt), "state": "rt" if rt else state, "n_gpus": world,
            "value": zones_local * world * steps / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
            "zones_per_gpu": zones_local, "scaling": "strong" if rt else "weak",
            "roofline_step_frac": alg * zones_local / (step_ms * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_zone_update": alg, "gpu_launches": launches * world}


def quick_sod(steps=400):
    """C1: Test_Problems/HD/Sod (400 zones, PLM + HLLC + RK2), the main loop run by pb200_integrate()."""
    import numpy as np
    import torch
    from pluto_sirocco_b200 import Hydro
    h = Hydro(dimensions=1, nx=(400, 1, 1), gamma=1.4, bcs=("outflow",) * 6, device=int(os.environ.get("LOCAL_RANK", "0")))
    x = (np.arange(400) + 0.5) / 400
    v = np.zeros((5, 1, 1, 400))
    v[0, 0, 0] = np.where(x < 0.5, 1.0, 0.125)
    v[4, 0, 0] = np.where(x < 0.5, 1.0, 0.1)
    h.set_interior(v)
    h.integrate(20, t=0.0, dt=1e-4, tstop=1e9, cfl=0.8, cfl_max_var=1.1, first_dt=1e-4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n, t, dt = h.integrate(steps, t=0.0, dt=1e-4, tstop=1e9, cfl=0.8, cfl_max_var=1.1, first_dt=1e-4)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = h.last.launches
    h.close()
    return {"workload": "sod1d-400 PLM+HLLC+RK2 (Test_Problems/HD/Sod conf 01)", "n_gpus": 1, "value": 400 * n / wall / 1e6,
            "unit": UNIT, "ms_per_step": 1e3 * wall / n, "steps": n, "gpu_launches_per_step": launches,
            "note": "launch/latency bound: 400 zones; wall clock of pb200_integrate() incl. the per-step dt read-back"}


# --------------------------------------------------------------------------------------
#  this repo's arm
# --------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pluto_sirocco_b200 import Hydro
    from pluto_sirocco_b200.slab import Slab, SlabHydro

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa else "off"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keeps NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.size
    rt = args.workload == "rt"
    if rt:
        # BASELINE configs[2]: Rayleigh-Taylor, 3-D, n x n x n GLOBAL zones split into world x3 slabs
        # (1024^3 over 8 GPUs = 1024 x 1024 x 128 per GPU); strong scaling in the driver's terms is not
        # asked for: run it with --gpus 8 --size 1024.  Periodic x1/x3, reflective x2, gravity along x2,
        # one tracer (oracle/problems/rt/init.c is the same set-up for the reference).
        if n % world:
            raise SystemExit("--size must be a multiple of the number of GPUs for --workload rt")
        gnx = (n, n, n)
        bcs = ("periodic", "periodic", "reflective", "reflective", "periodic", "periodic")
        gxb, gxe = (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)
        gamma, ntr, bf = 5. / 3., 1, 1
        zones_local = n * n * (n // world)
    else:
        gnx = (n, n, n * world)
        bcs = SEDOV_BCS
        gxb, gxe = (0., 0., 0.), (1., 1., float(world))
        gamma, ntr, bf = 1.4, 0, 0
        zones_local = n * n * n
    slab = Slab(rank, world, 3, gnx, gxb, gxe, bcs)
    xb, xe = slab.local_extent()
    h = Hydro(dimensions=3, nx=slab.local_nx(), xbeg=xb, xend=xe, gamma=gamma, reconstruction=args.recon,
              time_stepping=args.rk, solver=args.solver, bcs=slab.local_bcs(), device=local_rank,
              dx=slab.global_dx(), ntracer=ntr, body_force=bf)
    if rt:
        for comp, val in enumerate((0.0, -0.1, 0.0)):      # BodyForceVector = (0, GRAV, 0)
            h.set_body_force_vector(comp, np.full((1, 1, 1), val))
    sh = SlabHydro(h, slab)
    zones_total = zones_local * world

    # host state in pinned memory (needed by the e2e leg; also the upload source)
    pin = torch.empty(h.shape, dtype=torch.float64, pin_memory=True)
    vc = pin.numpy()
    vc[:] = 1.0
    vc[1:4] = 0.0
    if rt:
        vc[h.interior()] = rt_block(slab.local_nx(), slab.offset, n)
    elif args.state == "sedov":
        vc[h.interior()] = sedov_block(slab.local_nx(), slab.offset, n)
    else:   # "busy": seeded waves + jumps everywhere: every limiter/solver branch is exercised
        vc[h.interior()] = busy_block(slab.local_nx(), rank)
    h.upload(vc)

    cfl, cmv, first_dt = (0.4, 1.1, 1e-3) if rt else (0.3, 1.1, 1e-9)
    g = {"dt": first_dt if (args.state == "sedov" or rt) else 1e-5}

    def one_step():
        inv, mach, info = sh.advance_step(g["dt"])
        g["dt"] = h.next_time_step(inv, cfl, cmv, g["dt"], first_dt)
        return info

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs a few 100 ms to deliver its first sample: start it early
    for _ in range(args.warmup):
        one_step()
    if rank == 0:
        sampler.rows.clear()     # keep samples taken from here on (all of them under load)
    stream = sh.stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    launches = 0
    t_reg0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        info = one_step()
        launches += info.launches
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    t_reg1 = time.perf_counter()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    # a 20-step region lasts 0.3 s and nvidia-smi may deliver no sample inside it: keep the SAME load
    # running (untimed, same count on every rank) until about 1.5 s of load have been sampled, so that
    # clocks / throttle reasons under load are on record
    extra = 0 if ms >= 1500.0 else int((1500.0 - ms) / (ms / args.steps)) + 1
    for _ in range(extra):
        one_step()
    sync_all()
    clocks = sampler.stop(t_reg0, t_reg1) if rank == 0 else None
    if clocks is not None:
        clocks["untimed_steps_added_for_sampling"] = extra
    value = zones_total * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel times (CUDA events inside the library, on the launching stream) ----
    h.set_profiling(True)
    acc = {}
    nprof = 3
    for _ in range(nprof):
        one_step()
        for (kms, kdir, kstage) in h.kernel_times():
            acc.setdefault((kdir, kstage), []).append(kms)
    h.set_profiling(False)
    kern = {k: sum(v) / len(v) for k, v in acc.items()}
    sweep_ms = sum(kern.values())
    (ddir, dstage), dms = max(kern.items(), key=lambda kv: kv[1])

    # ---- e2e: host buffers, H2D + step + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h.download(vc)
        dt_e = g["dt"]
        deep = None
        if world > 1 and not rt:
            # Host-buffer steps of a slab-decomposed grid WITHOUT an exchange inside the step: every rank's host array
            # carries a deep halo of E = nghost x nstages planes of its neighbours on each cut face, the rank's block
            # (own + halo planes, any fill on the cut faces) goes through the same slab-pipelined call as at N = 1, only
            # the own planes come back (pb200_set_owned_planes), and afterwards the neighbours' fresh edge planes are
            # fetched into the host halo (NCCL device to device, then D2H of those few planes).
            from pluto_sirocco_b200.slab import device_view
            E = h.nghost * h.nstages()
            lo_ext, hi_ext = (E if rank > 0 else 0), (E if rank < world - 1 else 0)
            nloc = n + lo_ext + hi_ext
            he = Hydro(dimensions=3, nx=(n, n, nloc), xbeg=(0., 0., (rank * n - lo_ext) / float(n)),
                       xend=(1., 1., (rank * n + n + hi_ext) / float(n)), gamma=gamma, reconstruction=args.recon,
                       time_stepping=args.rk, solver=args.solver, bcs=SEDOV_BCS, device=local_rank, dx=slab.global_dx())
            he.set_owned_planes(lo_ext, lo_ext + n)
            pin_e = torch.empty(he.shape, dtype=torch.float64, pin_memory=True)
            vce = pin_e.numpy()
            vce[:] = 1.0
            vce[1:4] = 0.0
            vce[he.interior()] = sedov_block((n, n, nloc), rank * n - lo_ext, n)
            ng = he.nghost
            dev = torch.device("cuda", local_rank)
            bufs = [torch.empty((he.nvar, E) + he.shape[2:], dtype=torch.float64, device=dev) for _ in range(4)]
            deep = dict(he=he, vce=vce, pin=pin_e, E=E, lo=lo_ext, hi=hi_ext, ng=ng, bufs=bufs, dev=dev)
            dt_e = first_dt

        def deep_refresh_halo():
            """the neighbours' fresh edge planes -> this rank's host halo planes"""
            he, E, ng, lo_ext, hi_ext = deep["he"], deep["E"], deep["ng"], deep["lo"], deep["hi"]
            view = device_view(he.device_vc_ptr(), he.shape, deep["dev"])
            send_lo, send_hi, recv_lo, recv_hi = deep["bufs"]
            k_own0 = ng + lo_ext
            ops = []
            if hi_ext:
                send_hi.copy_(view[:, k_own0 + n - E:k_own0 + n])
                ops.append(dist.P2POp(dist.isend, send_hi, rank + 1))
            if lo_ext:
                send_lo.copy_(view[:, k_own0:k_own0 + E])
                ops.append(dist.P2POp(dist.isend, send_lo, rank - 1))
            if lo_ext:
                ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1))
            if hi_ext:
                ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1))
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            if lo_ext:
                deep["pin"][:, ng:ng + E].copy_(recv_lo, non_blocking=True)
            if hi_ext:
                deep["pin"][:, k_own0 + n:k_own0 + n + E].copy_(recv_hi, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def e2e_step():
            nonlocal dt_e
            if world == 1:
                info = h.advance_step_host(vc, dt_e)
                inv = info.invDt_hyp
            elif deep is not None:
                info = deep["he"].advance_step_host(deep["vce"], dt_e)
                inv, mach = __import__("pluto_sirocco_b200.slab", fromlist=["allreduce_max"]).allreduce_max(
                    [info.invDt_hyp, info.maxMach], deep["dev"])
                deep_refresh_halo()
            else:
                h.upload(vc)
                inv, mach, info = sh.advance_step(dt_e)
                h.download(vc)
            dt_e = h.next_time_step(inv, cfl, cmv, dt_e, first_dt)

        e2e_step()
        sync_all()
        ne = max(2, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(ne):
            e2e_step()
        ev1.record(stream)
        sync_all()
        wall = time.perf_counter() - t0
        ems = max(ev0.elapsed_time(ev1), wall * 1e3)   # copies are synchronous: wall clock covers them
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = t.item()
        nbytes = int(np.prod(h.shape)) * 8
        up_b, down_b = nbytes * world, nbytes * world
        api = ("pb200_advance_step_host (pinned host d->Vc in, d->Vc out, every step; upload, the RK stages and "
               "download pipelined over slabs of x3 planes)")
        if world > 1 and deep is None:
            api = "per rank: upload of the pinned host d->Vc slab, AdvanceStep with NCCL halo exchange, download, every step"
        if deep is not None:
            he = deep["he"]
            plane_b = he.nvar * he.shape[2] * he.shape[3] * 8
            mine = torch.tensor([float(np.prod(he.shape)) * 8, float(plane_b * (n + deep["lo"] + deep["hi"]))],
                                dtype=torch.float64, device="cuda")
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
            up_b, down_b = int(mine[0].item()), int(mine[1].item())
            api = ("per rank: pb200_advance_step_host on the rank's slab + a deep halo of nghost x nstages = %d planes per cut "
                   "face (no exchange inside the step, slab-pipelined like N = 1, own planes come back: "
                   "pb200_set_owned_planes), then the neighbours' edge planes refresh the host halo" % deep["E"])
            he.close()
        e2e = {"value": zones_total * ne / (ems * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": up_b, "d2h_bytes_per_step": down_b, "steps": ne, "api": api, "host_numa": numa}

    # ---- the other BASELINE configs, a few steps each (device-resident), reported under "secondary" ----
    nvar_head, shape_head = h.nvar, tuple(h.shape)
    secondary = None
    headline_cfg = (not rt and args.size == 512 and args.recon == "LINEAR" and args.rk == "RK2" and args.state == "sedov")
    if headline_cfg and not args.no_secondary:
        h.close()
        torch.cuda.empty_cache()
        secondary = {}
        ks = max(5, min(args.steps, 10))

        def quick_cart(**kw):      # never lose the headline line to a secondary (single rank: no collective can be left hanging)
            try:
                return _quick_cart(**kw)
            except Exception as e:
                if world > 1:
                    raise
                return {"error": repr(e)[:300]}

        # configs[1] again on a state with structure in every zone (the Sedov state is uniform outside the blast)
        secondary["c2_sedov_grid_busy_state"] = quick_cart(workload="sedov", size=512, recon="LINEAR", rk="RK2", state="busy", steps=ks)
        # configs[4]: PPM + HLLC + RK3, 256^3 per GPU, weak scaling
        secondary["c5_ppm_rk3_256"] = quick_cart(workload="sedov", size=256, recon="PARABOLIC", rk="RK3", state="sedov", steps=2 * ks)
        # configs[2]: Rayleigh-Taylor 1024^3 over 8 GPUs (strong-scaling form: n^3 GLOBAL zones over the ranks of this job;
        # the named size needs all 8 GPUs, smaller jobs run 512^3)
        secondary["c3_rayleigh_taylor"] = quick_cart(workload="rt", size=1024 if world == 8 else 512, recon="LINEAR", rk="RK2",
                                                     state="rt", steps=ks)
        if rank == 0:
            secondary["c1_sod_400"] = quick_sod()
        if world == 1:
            # configs[3]: the line-driven wind on the general path (single GPU: replicas only)
            try:
                secondary["c4_line_driven_wind"] = run_ldw(args, as_dict=True)
            except Exception as e:      # never lose the headline line to a secondary
                secondary["c4_line_driven_wind"] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    alg = (ALG_BYTES_RK2 if args.rk == "RK2" else ALG_BYTES_RK3) * nvar_head / 5.0   # 40 B per 5-vector -> 8 B x NVAR
    nlaunch_sweeps = len(kern)
    step_ms = ms / args.steps
    achieved_step = alg * zones_local / (step_ms * 1e-3) / 1e9          # GB/s per GPU, whole step
    achieved_kernel = (alg / nlaunch_sweeps) * zones_local / (dms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("x%d_stage%d" % (ddir + 1, dstage))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved_kernel, "peak": peak, "unit": "GB/s",
                "frac": achieved_kernel / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "x%d sweep, stage %d" % (ddir + 1, dstage), "kernel_ms": dms,
                "algorithmic_bytes_per_launch": (alg / nlaunch_sweeps) * zones_local,
                "step": {"achieved": achieved_step, "frac": achieved_step / peak, "ms": step_ms,
                         "sweep_kernels_ms": sweep_ms, "algorithmic_bytes_per_zone_update": alg},
                "kernels_ms": {"x%d_stage%d" % (k[0] + 1, k[1]): v for k, v in sorted(kern.items())}}

    cpu_baseline = None
    if not args.no_cpu and not rt:
        cfg = "sedov3d" if args.recon == "LINEAR" else "sedov3d_ppm"
        r = reference_rate(cfg, (args.cpu_size,) * 3, 1, 1, args.cpu_steps, args.solver)
        if r is not None:
            cpu_baseline = {"value": r[0], "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": "unmodified reference executable (oracle/_ref/%s), %d^3 zones, %d steps, 1 core, "
                                      "init excluded by differencing two runs" % (cfg, args.cpu_size, args.cpu_steps)}
        else:
            sys.path.insert(0, str(ROOT / "oracle"))
            from oracle import Oracle
            m = 64
            o = Oracle(dimensions=3, nx=(m, m, m), gamma=1.4, reconstruction=args.recon, time_stepping=args.rk,
                       solver=args.solver, bcs=bcs)
            w = o.embed(sedov_block((m, m, m), 0, m))
            t0 = time.perf_counter()
            ns = 8
            o.integrate(w, ns, t=0.0, dt=1e-9, tstop=0.5, cfl=0.3, cfl_max_var=1.1, first_dt=1e-9)
            dtw = time.perf_counter() - t0
            cpu_baseline = {"value": m ** 3 * ns / dtw / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "oracle/hd_oracle.c, %d^3 zones, %d steps, 1 core" % (m, ns)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if rt else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cart_config(args, world, zones_local, nvar_head),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "secondary": secondary}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------
#  secondary workload (BASELINE.json configs[3]): line-driven disc wind, spherical r-theta
# --------------------------------------------------------------------------------------
ALG_BYTES_LDW = 224.0 + 280.0 + 2 * 576.0   # NVAR 7 stages + 36 x (F_r, F_theta) x 8 B per stage


def ldw_state(x1, x2, P, U):
    """Initial condition of Test_Problems/LineDrivenWind/cv_idl/init.c:27-102 (synthetic input of
    that shape): hydrostatic PSD98 disc atmosphere, Keplerian rotation, density floor, tracer."""
    import numpy as np
    G, Rgas, sigma, amu, kB = 6.6726e-8, 8.3144598e7, 5.67051e-5, 1.66053886e-24, 1.3806505e-16
    KELVIN = U["velocity"] ** 2 * amu / kB
    r_WD = x1.min()
    X1, X2 = np.meshgrid(x1, x2, indexing="xy")          # [j][i]
    gm = G * P["CENT_MASS"]
    r = X1 * U["length"]
    teff = (3.0 * gm * P["DISK_MDOT"] / (8.0 * np.pi * sigma)) ** 0.25 * (r_WD * U["length"]) ** -0.75
    temp = teff * (r_WD / X1) ** 0.75
    cs2 = Rgas * temp / 0.6
    with np.errstate(divide="ignore", over="ignore", under="ignore"):
        rho_d = P["RHO_0"] * np.exp(-gm / (2.0 * cs2 * r * np.tan(X2) ** 2)) / U["density"]
    rho_a = P["DFLOOR"] / U["density"]
    disc = rho_d > rho_a
    v = np.zeros((7, 1) + X1.shape)
    v[0, 0] = np.where(disc, rho_d, rho_a)
    v[3, 0] = np.sqrt(gm / r) * np.sin(X2) / U["velocity"]
    v[4, 0] = v[0, 0] * temp / (KELVIN * P["MU"])
    v[5, 0] = disc.astype(float)
    v[6, 0] = 1.0
    return v


def run_ldw(args, as_dict=False):
    """C4: 1024 x 512 spherical r-theta grid, NVAR 7 (tracer + entropy), char-limited PLM (van Leer),
    MULTID flattening, entropy switch, BODY_FORCE VECTOR, VGradCalc + LineForce with 36-angle
    synthetic sirocco tables, cv_idl user boundaries; HLL, RK2 (pluto_sirocco_sub.py:46-118)."""
    import numpy as np
    import torch
    from pluto_sirocco_b200 import Hydro, make_grid
    sys.path.insert(0, str(ROOT / "tests"))
    from common import LDW_BCS, LDW_PARAMS, LDW_UNITS, ldw_flux_tables
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    n1, n2 = args.ldw_size
    grid = [(0.87, n1, 8.7, "r", args.ldw_ratio[0]), (0.0, n2, float(np.radians(90.0)), "r", args.ldw_ratio[1]), (0.0, 1, 1.0)]
    arrays = [make_grid(grid[0], 3), make_grid(grid[1], 3), make_grid(grid[2], 0)]
    h = Hydro(dimensions=2, nx=(n1, n2, 1), xbeg=(0.87, 0.0, 0.0), xend=(8.7, float(np.radians(90.0)), 1.0),
              gamma=5. / 3., solver="hll", limiter="VANLEER_LIM", bcs=LDW_BCS, ntracer=1, body_force=1,
              geometry="SPHERICAL", grid_arrays=arrays, char_limiting=True, shock_flattening=True,
              entropy_switch=True, nghost=3)
    x1, x2 = h.x(0), h.x(1)
    gm_code = 6.6726e-8 * LDW_PARAMS["CENT_MASS"] / (LDW_UNITS["length"] * LDW_UNITS["velocity"] ** 2)
    h.set_body_force_vector(0, (-gm_code / (x1 * x1)).reshape(1, 1, -1))
    h.set_body_force_vector(1, np.zeros((1, 1, 1)))
    h.set_body_force_vector(2, np.zeros((1, 1, 1)))
    fr, ft, fp = ldw_flux_tables(x1, x2, roundtrip=False)
    h.set_ldw(params=LDW_PARAMS, units=LDW_UNITS, flux_r=fr, flux_t=ft, flux_p=fp)
    v = ldw_state(x1[3:-3], x2[3:-3], LDW_PARAMS, LDW_UNITS)
    h.set_interior(v)
    zones = n1 * n2
    cfl, cmv, first_dt = 0.4, 1.1, 1e-4
    g = {"dt": first_dt}

    def one_step():
        info = h.advance_step(g["dt"])
        g["dt"] = h.next_time_step(info.invDt_hyp, cfl, cmv, g["dt"], first_dt)
        return info

    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    sampler.rows.clear()
    stream = torch.cuda.ExternalStream(h.stream_ptr())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    launches = 0
    t_reg0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        launches += one_step().launches
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    t_reg1 = time.perf_counter()
    extra = 0 if (ms >= 1500.0 or as_dict) else min(300, int((1500.0 - ms) / (ms / args.steps)) + 1)   # same load, untimed, for the clock samples
    done = 0
    saved, saved_dt = h.download(), g["dt"]
    try:
        for _ in range(extra):
            one_step(); done += 1
    except Exception:        # the synthetic wind is not meant to be integrated far: stop sampling there
        pass
    extra = done
    torch.cuda.synchronize()
    h.upload(saved)          # the e2e leg continues from the end of the timed region
    g["dt"] = saved_dt
    clocks = sampler.stop(t_reg0, t_reg1)
    clocks["untimed_steps_added_for_sampling"] = extra
    value = zones * args.steps / (ms * 1e-3) / 1e6
    # e2e: host d->Vc in and out every step
    vc = h.download()
    pin = torch.empty(h.shape, dtype=torch.float64, pin_memory=True)
    pin.numpy()[:] = vc
    t0 = time.perf_counter()
    ne = max(2, min(args.steps, 20))
    dt_e = g["dt"]
    for _ in range(ne):
        info = h.advance_step_host(pin.numpy(), dt_e)
        dt_e = h.next_time_step(info.invDt_hyp, cfl, cmv, dt_e, first_dt)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    nbytes = int(np.prod(h.shape)) * 8
    e2e = {"value": zones * ne / wall / 1e6, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
           "steps": ne, "api": "pb200_advance_step_host"}
    peak, peak_src = measured_peaks()
    step_ms = ms / args.steps
    achieved = ALG_BYTES_LDW * zones / (step_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "line-driven-wind spherical r-theta %dx%d PLM(char,VL)+HLL+RK2, 36-angle line force" % (n1, n2),
                       "grid_ratio": list(args.ldw_ratio), "zones_per_gpu": zones, "nvar": 7,
                       "l2": "state + work arrays (%.0f MB) and the 36-angle tables (%.0f MB) exceed L2 together" % (
                           zones * 8 * 7 * 7 / 1e6, zones * 8 * 36 * 4 / 1e6), "cfl": cfl},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "whole step (multi-kernel general path)",
                         "algorithmic_bytes_per_zone_update": ALG_BYTES_LDW},
            "cpu_baseline": None}
    h.close()
    if as_dict:
        return {k: line[k] for k in ("value", "unit", "ms_per_step", "steps", "gpu_launches", "config", "e2e")} | {
            "n_gpus": 1, "roofline_step_frac": line["roofline"]["frac"],
            "algorithmic_bytes_per_zone_update": ALG_BYTES_LDW}
    emit(line)
    return 0


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # libraries (NCCL's version banner, ...) write to fd 1 from C: keep stdout for the JSON line only
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sedov", choices=["sedov", "rt", "ldw"],
                    help="sedov: BASELINE configs[1] (the metric's config; configs[4] with --recon PARABOLIC --rk RK3 "
                         "--size 256); rt: configs[2] (--gpus 8 --size 1024); ldw: configs[3]")
    ap.add_argument("--ldw-size", type=int, nargs=2, default=[1024, 512])
    ap.add_argument("--ldw-ratio", type=float, nargs=2, default=[1.005, 0.995])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="zones per direction per GPU")
    ap.add_argument("--recon", default="LINEAR", choices=["LINEAR", "PARABOLIC"])
    ap.add_argument("--rk", default="RK2", choices=["RK2", "RK3"])
    ap.add_argument("--solver", default="hllc", choices=["hllc", "hll", "tvdlf"])
    ap.add_argument("--state", default="sedov", choices=["sedov", "busy"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--cpu-size", type=int, default=128)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--ref-size", type=int, default=0, help="reference arm: cubic sample of this size instead of x3 slabs of the workload grid")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "ldw":
        return run_ldw(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
