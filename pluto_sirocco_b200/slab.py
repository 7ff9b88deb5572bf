"""Slab decomposition across the GPUs of one box: the replacement of Src/Parallel (ArrayLib).

Reference behaviour replaced:
  * al_decompose.c:40,125-158   MPI Cartesian decomposition  -> 1-D slab split along the
    OUTERMOST active direction (x3 in 3-D, x2 in 2-D), whose ghost planes are contiguous per
    variable in Vc[nv][k][j][i] so no pack kernel is needed;
  * boundary.c:139-158 + al_exchange_dim.c:64-78  per-variable MPI_Sendrecv pairs -> one
    grouped batch of NCCL send/recv (torch.distributed.batch_isend_irecv) per stage, issued on
    the stream the sweep kernels run on;
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


def post_halo_exchange(vc: torch.Tensor, slab: Slab, nghost: int):
    """Post the exchange of the ghost planes of direction slab.sdir of vc[nv][k][j][i] (ghosts
    included) with the neighbouring ranks' edge planes; returns the outstanding requests (wait on
    them before reading the ghost planes).  With NCCL the transfers run on NCCL's own stream,
    ordered after the work already enqueued on the current stream, so kernels launched between
    post and wait overlap with them.  Works for CUDA tensors (nccl) and CPU tensors (gloo).
    Every plane set is contiguous per variable, so tensors are sent in place (no packing)."""
    lo, hi = slab.neighbours()
    if lo is None and hi is None:
        return []
    axis = 3 - slab.sdir                 # position of the split direction in [nv][k][j][i]
    n = vc.shape[axis]
    ng = nghost
    ops = []
    nvar = vc.shape[0]

    def planes(a, b):
        return [vc[nv].narrow(axis - 1, a, b - a) for nv in range(nvar)]

    lo_ghost, lo_edge = planes(0, ng), planes(ng, 2 * ng)
    hi_edge, hi_ghost = planes(n - 2 * ng, n - ng), planes(n - ng, n)
    for nv in range(nvar):
        for t in (lo_ghost[nv], lo_edge[nv], hi_edge[nv], hi_ghost[nv]):
            assert t.is_contiguous(), "slab planes must be contiguous per variable"
    # Per peer, NCCL pairs sends and receives in posting order (tags are honoured by gloo
    # only).  Post "upward" traffic first (hi_edge -> peer's lo_ghost), then "downward", so a
    # periodic pair of 2 ranks (lo == hi) matches correctly as well.
    if hi is not None:
        for nv in range(nvar):
            ops.append(dist.P2POp(dist.isend, hi_edge[nv], hi, tag=nv))
    if lo is not None:
        for nv in range(nvar):
            ops.append(dist.P2POp(dist.isend, lo_edge[nv], lo, tag=100 + nv))
    if lo is not None:
        for nv in range(nvar):
            ops.append(dist.P2POp(dist.irecv, lo_ghost[nv], lo, tag=nv))
    if hi is not None:
        for nv in range(nvar):
            ops.append(dist.P2POp(dist.irecv, hi_ghost[nv], hi, tag=100 + nv))
    return dist.batch_isend_irecv(ops)


def exchange_halos(vc: torch.Tensor, slab: Slab, nghost: int):
    """Blocking form: post + wait."""
    for r in post_halo_exchange(vc, slab, nghost):
        r.wait()


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
                reqs = post_halo_exchange(self._view(h.stage_array_ptr(s)), self.slab, h.nghost)
                h.stage_begin(s)
                for r in reqs:
                    r.wait()
                h.stage_finish(s)
            info = h.step_end()
            if self.slab.world > 1:
                inv, mach = allreduce_max([info.invDt_hyp, info.maxMach], self.device)
            else:
                inv, mach = info.invDt_hyp, info.maxMach
        return inv, mach, info
