"""Slab decomposition across the GPUs of one box: the replacement of Src/Parallel (ArrayLib).

Reference behaviour replaced:
  * al_decompose.c:40,125-158   MPI Cartesian decomposition  -> 1-D slab split along the
    OUTERMOST active direction (x3 in 3-D, x2 in 2-D), whose ghost planes are contiguous per
    variable in Vc[nv][k][j][i] so no pack kernel is needed;
  * boundary.c:139-158 + al_exchange_dim.c:64-78  per-variable MPI_Sendrecv pairs -> one
    grouped batch of NCCL send/recv (torch.distributed.batch_isend_irecv) of packed edge planes per
    stage, overlapped with the fused x1+x2 kernel;
  * main.c:547 (+ :288) MPI_Allreduce(MAX) of invDt_hyp / g_maxMach -> one 2-double
    all_reduce(MAX).
One process per GPU (torchrun).  The same code runs on CPU tensors over gloo, which is how
the host-side logic is tested without GPUs (tests/test_slab_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Slab:
    """Block owned by `rank` out of `world` along direction `sdir`."""
    rank: int
    world: int
    dimensions: int
    global_nx: tuple
    xbeg: tuple
    xend: tuple
    bcs: tuple            # global boundary types (strings), 6 entries

    def __post_init__(self):
        self.sdir = self.dimensions - 1
        n = self.global_nx[self.sdir]
        if self.world > 1 and self.dimensions == 1:
            raise ValueError("1-D grids are not decomposed (replicas only)")
        base, rem = divmod(n, self.world)
        counts = [base + (1 if r < rem else 0) for r in range(self.world)]
        self.counts = counts
        self.offset = sum(counts[:self.rank])
        self.local_n = counts[self.rank]

    # -- what pb200_create / Oracle need for this block -------------------------------------
    def local_nx(self):
        nx = list(self.global_nx)
        nx[self.sdir] = self.local_n
        return tuple(nx)

    def local_extent(self):
        d = self.sdir
        dx = (self.xend[d] - self.xbeg[d]) / self.global_nx[d]
        xb, xe = list(self.xbeg), list(self.xend)
        xb[d] = self.xbeg[d] + self.offset * dx
        xe[d] = self.xbeg[d] + (self.offset + self.local_n) * dx
        return tuple(xb), tuple(xe)

    def global_dx(self):
        """grid->dx of the undecomposed uniform grid (Src/set_grid.c:410): every block must use
        THIS value, not (xend-xbeg)/n of its own extent, to stay bit-identical with a serial run."""
        return tuple((self.xend[d] - self.xbeg[d]) / self.global_nx[d] if d < self.dimensions else 1.0
                     for d in range(3))

    def periodic(self):
        return self.bcs[2 * self.sdir] == "periodic"

    def neighbours(self):
        """(lo_rank, hi_rank) or None where the block touches a physical (non-periodic) boundary."""
        lo = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.periodic() else None)
        hi = self.rank + 1 if self.rank < self.world - 1 else (0 if self.periodic() else None)
        if self.world == 1:
            return None, None
        return lo, hi

    def local_bcs(self):
        b = list(self.bcs)
        lo, hi = self.neighbours()
        if lo is not None:
            b[2 * self.sdir] = "neighbour"
        if hi is not None:
            b[2 * self.sdir + 1] = "neighbour"
        return tuple(b)

    def local_slice(self):
        """slice of the global interior array [nv][nz][ny][nx] owned by this rank."""
        sl = [slice(None)] * 4
        sl[3 - self.sdir] = slice(self.offset, self.offset + self.local_n)
        return tuple(sl)


class HaloExchange:
    """One exchange of the ghost planes of direction slab.sdir of vc[nv][k][j][i] (ghosts included)
    with the neighbouring ranks.  The edge planes of all variables are packed into ONE contiguous
    buffer per face (a strided device copy), so an exchange is at most 2 sends + 2 receives however
    many variables there are - at 256^3 zones per GPU the per-message latency, not the bandwidth, is
    what a stage waits for.  post() packs and posts the transfers (with NCCL they run on NCCL's own
    stream, ordered after the work already enqueued on the current stream, so kernels launched
    between post() and finish() overlap with them); finish() waits and unpacks into the ghost planes.
    Works for CUDA tensors (nccl) and CPU tensors (gloo)."""

    def __init__(self, slab: Slab, nghost: int):
        self.slab, self.ng = slab, nghost
        self._buf = {}

    def _buffers(self, vc):
        key = (vc.data_ptr(), vc.shape)
        if key not in self._buf:
            axis = 3 - self.slab.sdir
            shape = list(vc.shape)
            shape[axis] = self.ng
            self._buf[key] = [torch.empty(shape, dtype=vc.dtype, device=vc.device) for _ in range(4)]
        return self._buf[key]

    def post(self, vc: torch.Tensor):
        lo, hi = self.slab.neighbours()
        self._pending = None
        if lo is None and hi is None:
            return
        axis = 3 - self.slab.sdir                 # position of the split direction in [nv][k][j][i]
        n, ng = vc.shape[axis], self.ng
        send_lo, send_hi, recv_lo, recv_hi = self._buffers(vc)
        ops = []
        # Per peer, NCCL pairs sends and receives in posting order (tags are honoured by gloo only).
        # Post "upward" traffic first (hi edge -> peer's lo ghosts), then "downward", so a periodic
        # pair of 2 ranks (lo == hi) matches correctly as well.
        if hi is not None:
            send_hi.copy_(vc.narrow(axis, n - 2 * ng, ng))
            ops.append(dist.P2POp(dist.isend, send_hi, hi, tag=0))
        if lo is not None:
            send_lo.copy_(vc.narrow(axis, ng, ng))
            ops.append(dist.P2POp(dist.isend, send_lo, lo, tag=1))
        if lo is not None:
            ops.append(dist.P2POp(dist.irecv, recv_lo, lo, tag=0))
        if hi is not None:
            ops.append(dist.P2POp(dist.irecv, recv_hi, hi, tag=1))
        self._pending = (dist.batch_isend_irecv(ops), vc, axis, n, lo, hi)

    def finish(self):
        if self._pending is None:
            return
        reqs, vc, axis, n, lo, hi = self._pending
        for r in reqs:
            r.wait()
        _, _, recv_lo, recv_hi = self._buffers(vc)
        if lo is not None:
            vc.narrow(axis, 0, self.ng).copy_(recv_lo)
        if hi is not None:
            vc.narrow(axis, n - self.ng, self.ng).copy_(recv_hi)
        self._pending = None


def exchange_halos(vc: torch.Tensor, slab: Slab, nghost: int):
    """Blocking form: pack, post, wait, unpack."""
    x = HaloExchange(slab, nghost)
    x.post(vc)
    x.finish()


def allreduce_max(values, device):
    """MPI_Allreduce(MAX) of a few doubles (invDt_hyp, maxMach): main.c:288,547."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


class _DevPtr:
    """__cuda_array_interface__ view of a raw device pointer (zero copy into torch)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr="<f8", data=(int(ptr), False),
                                             version=2, strides=None)


def device_view(ptr: int, shape, device) -> torch.Tensor:
    return torch.as_tensor(_DevPtr(ptr, shape), device=device)


class SlabHydro:
    """AdvanceStep over a slab-decomposed grid: per stage, halo exchange on the array the
    stage sweeps, then the stage; per step, one max-allreduce for dt (NextTimeStep)."""

    def __init__(self, hydro, slab: Slab):
        self.h = hydro
        self.slab = slab
        self.device = torch.device("cuda", hydro.cfg.device)
        self.stream = torch.cuda.ExternalStream(hydro.stream_ptr(), device=self.device)
        self._views = {}
        self._halo = HaloExchange(slab, hydro.nghost)

    def _view(self, ptr):
        if ptr not in self._views:
            self._views[ptr] = device_view(ptr, self.h.shape, self.device)
        return self._views[ptr]

    def advance_step(self, dt):
        h = self.h
        with torch.cuda.stream(self.stream):
            h.step_begin(dt)
            for s in range(1, h.nstages() + 1):
                # the x3 ghost planes are only read by the x3 sweep: fill the physical boundaries,
                # post the exchange, run the fused x1+x2 kernel while the planes travel, then wait
                h.stage_boundary(s)
                self._halo.post(self._view(h.stage_array_ptr(s)))
                h.stage_begin(s)
                self._halo.finish()
                h.stage_finish(s)
            info = h.step_end()
            if self.slab.world > 1:
                inv, mach = allreduce_max([info.invDt_hyp, info.maxMach], self.device)
            else:
                inv, mach = info.invDt_hyp, info.maxMach
        return inv, mach, info
