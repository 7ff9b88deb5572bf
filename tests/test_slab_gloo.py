"""CPU, world_size 2 over gloo: the slab decomposition + halo exchange + max-allreduce host
logic (pluto_sirocco_b200/slab.py) driven with the oracle as the per-block compute, compared
with the undecomposed oracle.  Serial == parallel is the reference's own claim
(Src/flag_shock.c:207-219); here it must hold bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dims, gnx, bcs, recon, rk, nsteps, outdir):
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import random_state
    from oracle import Oracle
    from pluto_sirocco_b200.slab import Slab, allreduce_max, exchange_halos

    slab = Slab(rank, world, dims, gnx, (0., 0., 0.), (1., 1., 1.), bcs)
    xb, xe = slab.local_extent()
    o = Oracle(dimensions=dims, nx=slab.local_nx(), xbeg=xb, xend=xe, gamma=1.4, reconstruction=recon,
               time_stepping=rk, solver="hllc", bcs=slab.local_bcs(), dx=slab.global_dx())
    vglob = random_state((gnx[2], gnx[1], gnx[0]), seed=42, smooth=False)
    vc = o.embed(vglob[slab.local_slice()])
    # the oracle is driven stage by stage (orc_step_begin / orc_stage / orc_step_end) so the
    # halo exchange sits where the reference's Boundary() does its MPI exchange
    from oracle import lib as olib
    import ctypes as C
    L = olib()
    t = torch.from_numpy(vc)
    dt = 2e-4
    for n in range(nsteps):
        inv = C.c_double(0.0); mach = C.c_double(0.0)
        st = L.orc_step_begin(C.byref(o.c), vc.ctypes.data_as(C.c_void_p))
        for s in range(1, {"EULER": 1, "RK2": 2, "RK3": 3}[rk] + 1):
            exchange_halos(t, slab, o.nghost)
            L.orc_stage(C.byref(o.c), C.c_void_p(st), vc.ctypes.data_as(C.c_void_p), s, C.c_double(dt),
                        C.byref(inv), C.byref(mach))
        L.orc_step_end(C.c_void_p(st))
        ginv, gmach = allreduce_max([inv.value, mach.value], "cpu")
        dt = min(Oracle.next_time_step(ginv, 0.3, 1.1, dt, 1e-6), 1.1 * dt)
    np.save(os.path.join(outdir, "rank%d.npy" % rank), vc[o.interior()])
    np.save(os.path.join(outdir, "dt%d.npy" % rank), np.array([dt]))
    dist.destroy_process_group()


@pytest.mark.parametrize("dims,gnx,bcs,recon,rk", [
    (3, (12, 10, 16), ("outflow", "reflective", "periodic", "periodic", "reflective", "outflow"), "LINEAR", "RK2"),
    (3, (10, 8, 13), ("periodic",) * 6, "PARABOLIC", "RK3"),
    (2, (24, 17, 1), ("reflective", "outflow", "periodic", "periodic", "outflow", "outflow"), "LINEAR", "RK2"),
])
def test_two_slabs_equal_single_domain(tmp_path, dims, gnx, bcs, recon, rk):
    from common import random_state
    from oracle import Oracle
    world, nsteps = 2, 3
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dims, gnx, bcs, recon, rk, nsteps, str(tmp_path)), nprocs=world, join=True)
    # undecomposed run
    o = Oracle(dimensions=dims, nx=gnx, gamma=1.4, reconstruction=recon, time_stepping=rk, solver="hllc", bcs=bcs)
    v = random_state((gnx[2], gnx[1], gnx[0]), seed=42, smooth=False)
    vc = o.embed(v)
    dt = 2e-4
    for n in range(nsteps):
        inv, mach, nf = o.advance_step(vc, dt)
        dt = min(Oracle.next_time_step(inv, 0.3, 1.1, dt, 1e-6), 1.1 * dt)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world)]
    axis = 3 - (dims - 1)
    glued = np.concatenate(parts, axis=axis)
    assert np.array_equal(glued, vc[o.interior()])
    for r in range(world):
        assert np.load(tmp_path / ("dt%d.npy" % r))[0] == dt


def test_slab_partition_arithmetic():
    from pluto_sirocco_b200.slab import Slab
    bcs = ("outflow",) * 4 + ("periodic", "periodic")
    tot = 0
    for r in range(3):
        s = Slab(r, 3, 3, (8, 8, 10), (0, 0, 0), (1, 1, 2.0), bcs)
        assert s.local_nx()[:2] == (8, 8)
        tot += s.local_n
        lo, hi = s.neighbours()
        assert lo == (r - 1) % 3 and hi == (r + 1) % 3
        assert s.local_bcs()[4:] == ("neighbour", "neighbour")
        xb, xe = s.local_extent()
        assert abs((xe[2] - xb[2]) - 0.2 * s.local_n) < 1e-15
    assert tot == 10
    s = Slab(0, 2, 3, (8, 8, 10), (0, 0, 0), (1, 1, 1), ("outflow",) * 6)
    assert s.neighbours() == (None, 1) and s.local_bcs()[4:] == ("outflow", "neighbour")
    with pytest.raises(ValueError):
        Slab(0, 2, 1, (8, 1, 1), (0, 0, 0), (1, 1, 1), ("outflow",) * 6)
