"""GPU: the C-side slab decomposition (csrc/pb200_multi.cu, one host thread driving N ranks) reproduces
the undecomposed single-context run of the same library bit for bit.  On a 1-GPU box the ranks share
device 0 and the edge planes move by device copies (the slab logic, packing, ordering and ghost handling
are the same code); with >= 2 devices the same cases also run over NCCL (ncclSend/ncclRecv/ncclAllReduce).
Replaces: Src/Parallel/al_exchange_dim.c:64-90, Src/boundary.c:139-158, Src/main.c:288,547."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    (3, (40, 24, 32), ("outflow", "reflective", "periodic", "periodic", "reflective", "outflow"), "LINEAR", "RK2", 0, 0),
    (3, (36, 20, 26), ("periodic",) * 6, "PARABOLIC", "RK3", 0, 0),
    (3, (36, 20, 27), ("periodic",) * 6, "LINEAR", "RK2", 1, 1),       # ragged split, tracer, gravity table
    (3, (36, 20, 26), ("outflow",) * 6, "PARABOLIC", "RK3", 0, 0),
    (2, (48, 40, 1), ("outflow", "outflow", "reflective", "outflow", "outflow", "outflow"), "LINEAR", "RK2", 0, 0),
]


def _run(make, v, nsteps, bf):
    h = make()
    if bf:
        g = np.linspace(-0.3, 0.1, h.tot[2 if h.dimensions == 3 else 1])
        tab = g[:, None, None] if h.dimensions == 3 else g[None, :, None]
        for comp in range(3):
            h.set_body_force_vector(comp, tab if comp == h.dimensions - 1 else np.zeros(1))
    h.set_interior(v)
    dt, dts = 2e-4, []
    for n in range(nsteps):
        info = h.advance_step(dt)
        dts.append((info.invDt_hyp, info.maxMach))
        dt = min(0.3 / info.invDt_hyp, 1.1 * dt)
    out = h.get_interior()
    h.close()
    return out, dts


@pytest.mark.parametrize("ranks,mode", [(2, "shared"), (3, "shared"), (2, "nccl"), (4, "nccl")])
@pytest.mark.parametrize("dims,gnx,bcs,recon,rk,ntr,bf", CASES)
def test_c_slab_decomposition_equals_single_context(cuda_lib, ranks, mode, dims, gnx, bcs, recon, rk, ntr, bf):
    import torch
    from common import random_state, rel_err
    from pluto_sirocco_b200 import Hydro, MultiHydro
    if mode == "nccl" and torch.cuda.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    devices = [0] * ranks if mode == "shared" else list(range(ranks))
    kw = dict(dimensions=dims, nx=gnx, gamma=1.4, reconstruction=recon, time_stepping=rk, solver="hllc", bcs=bcs,
              ntracer=ntr, body_force=1 if bf else 0)
    shape = (gnx[2], gnx[1], gnx[0]) if dims == 3 else (1, gnx[1], gnx[0])
    v = random_state(shape, seed=7, smooth=False)
    if ntr:
        tr = 0.5 + 0.5 * np.sign(v[1:2])
        v = np.concatenate([v, tr], axis=0)
    ref, dts_ref = _run(lambda: Hydro(**kw), v, 4, bf)
    got, dts = _run(lambda: MultiHydro(ranks, devices=devices, **kw), v, 4, bf)
    # Same kernels and data; a zone can sit at an even or an odd iteration of the 2x-unrolled marching
    # loop depending on where its slab starts and ptxas contracts FMAs differently in the two copies of
    # the loop body, so agreement is to the last ulp or two (see tests/test_slab_nccl_gpu.py)
    assert np.allclose(np.array(dts), np.array(dts_ref), rtol=1e-14, atol=0)   # invDt_hyp, maxMach after the max-reduction
    assert rel_err(got, ref) <= 1e-14


def test_c_slab_host_call_equals_resident(cuda_lib):
    """pb200_multi_advance_step_host (the drop-in's call: global d->Vc on the host in and out) equals the
    resident multi-rank step."""
    from common import random_state
    from pluto_sirocco_b200 import MultiHydro
    gnx = (32, 24, 28)
    kw = dict(dimensions=3, nx=gnx, gamma=1.4, bcs=("reflective", "outflow") * 3)
    v = random_state((gnx[2], gnx[1], gnx[0]), seed=3, smooth=False)
    a = MultiHydro(2, devices=[0, 0], **kw)
    a.set_interior(v)
    b = MultiHydro(2, devices=[0, 0], **kw)
    vc = np.ones(b.shape)
    vc[1:4] = 0.0
    vc[b.interior()] = v
    for n in range(3):
        ia = a.advance_step(1e-4)
        ib = b.advance_step_host(vc, 1e-4)
        assert ia.invDt_hyp == ib.invDt_hyp
    assert np.array_equal(a.get_interior(), vc[b.interior()])     # same slabs, same order: bit identical
    a.close(); b.close()


def test_c_slab_refusals(cuda_lib):
    from pluto_sirocco_b200 import MultiHydro
    from pluto_sirocco_b200._lib import PB200Error
    with pytest.raises(PB200Error):      # slabs thinner than 2*nghost
        MultiHydro(4, devices=[0] * 4, dimensions=3, nx=(16, 16, 8))
    with pytest.raises(PB200Error):      # 1-D: replicas only
        MultiHydro(2, devices=[0, 0], dimensions=1, nx=(64, 1, 1))
