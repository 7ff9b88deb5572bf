// pb200_sweeps.cu -- instantiation and launch of the sweep kernels for ONE (NVAR, BODY_FORCE)
// pair: compile with -DPB_NV=<5|6|7> -DPB_BF=<0|1>.  Solver / reconstruction / limiter are
// run-time options of the reference ([Solver] in pluto.ini) or cheap to carry as template
// parameters, so every combination is instantiated here.
#include <unordered_map>

#include "pb200_internal.h"

using namespace pb;

#ifndef PB_NV
#error "compile with -DPB_NV=5|6|7"
#endif
#ifndef PB_BF
#error "compile with -DPB_BF=0|1"
#endif

// ---- sweeps ------------------------------------------------------------------------------
template <typename K>
static void set_smem(K k, size_t shm) {
  // high-water mark of the opt-in dynamic shared memory per kernel
  static std::unordered_map<const void *, size_t> cur;
  size_t &c = cur[(const void *)k];
  if (shm > c) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm); c = shm; }
}

template <int NV, int RECON, int SOLVER, int LIM, int BF>
static void launch_dir(pb200_ctx *c, int dir, const SweepArgs &a) {
  const Dev &D = c->dev;
  const int slot = (c->profiling && c->nprof < 16) ? c->nprof++ : -1;
  if (slot >= 0) {
    if (!c->pev0[slot]) { cudaEventCreate(&c->pev0[slot]); cudaEventCreate(&c->pev1[slot]); }
    c->pdir[slot] = dir;
    c->pstage[slot] = a.stage;
    cudaEventRecord(c->pev0[slot], c->stream);
  }
  int nx = D.end[0] - D.beg[0] + 1;
  if (dir == 0) {  // DIMENSIONS == 1 only: plain x1 sweep
    constexpr int LO = (RECON == RECON_PARABOLIC) ? 2 : 1;
    constexpr int USE = BX - 1 - LO;
    dim3 grid((nx + USE - 1) / USE, D.end[1] - D.beg[1] + 1, D.end[2] - D.beg[2] + 1);
    sweep_x1<NV, RECON, SOLVER, BF><<<grid, BX, 0, c->stream>>>(D, a);
  } else {
    // dir 1: x1+x2 fused march along x2 ; dir 2: x3 march
    const bool fusex = (dir == 1);
    const bool last = (dir == D.ndim - 1);
    // block width of the fused kernel: fewest warps per row (ties go to the narrower block)
    int bx = BX;
    if (fusex) {
      const int h2 = 2 * recon_xhalo<RECON>();
      const int w128 = ((nx + (128 - h2) - 1) / (128 - h2)) * 4, w192 = ((nx + (192 - h2) - 1) / (192 - h2)) * 6;
      if (w192 < w128) bx = 192;
    }
    const int use = fusex ? bx - 2 * recon_xhalo<RECON>() : BX;
    int npen = D.end[dir] - D.beg[dir] + 1;
    int ntr = (dir == 1) ? (D.end[2] - D.beg[2] + 1) : (D.end[1] - D.beg[1] + 1);
    int nbx = (nx + use - 1) / use;
    // chunk the pencil so that the grid holds several waves of 148 SMs x resident blocks
    long want = 148L * (bx == 192 ? 2 : 3) * 6;
    int nchunk = 1;
    while ((long)nbx * ntr * nchunk < want && npen / (nchunk * 2) >= 32) nchunk *= 2;
    int chunk = (npen + nchunk - 1) / nchunk;
    nchunk = (npen + chunk - 1) / chunk;
    dim3 grid(nbx, ntr, nchunk);
    const bool cdt_in = D.ndim > 1 && a.stage == 1 && !fusex;
    const int nq = ring_nq(NV, fusex, a.comb, cdt_in);
    if (fusex && !last && bx == 192) {
      auto k = sweep_fused<1, true, false, NV, RECON, SOLVER, LIM, BF, 192>;
      size_t shm = sweep_smem_bytes<true, NV, RECON, 192>(nq);
      set_smem(k, shm);
      k<<<grid, 192, shm, c->stream>>>(D, a, chunk);
    } else if (fusex && bx == 192) {
      auto k = sweep_fused<1, true, true, NV, RECON, SOLVER, LIM, BF, 192>;
      size_t shm = sweep_smem_bytes<true, NV, RECON, 192>(nq);
      set_smem(k, shm);
      k<<<grid, 192, shm, c->stream>>>(D, a, chunk);
    } else if (fusex && !last) {
      auto k = sweep_fused<1, true, false, NV, RECON, SOLVER, LIM, BF>;
      size_t shm = sweep_smem_bytes<true, NV, RECON>(nq);
      set_smem(k, shm);
      k<<<grid, BX, shm, c->stream>>>(D, a, chunk);
    } else if (fusex) {
      auto k = sweep_fused<1, true, true, NV, RECON, SOLVER, LIM, BF>;
      size_t shm = sweep_smem_bytes<true, NV, RECON>(nq);
      set_smem(k, shm);
      k<<<grid, BX, shm, c->stream>>>(D, a, chunk);
    } else {
      auto k = sweep_fused<2, false, true, NV, RECON, SOLVER, LIM, BF>;
      size_t shm = sweep_smem_bytes<false, NV, RECON>(nq);
      set_smem(k, shm);
      k<<<grid, BX, shm, c->stream>>>(D, a, chunk);
    }
  }
  if (slot >= 0) cudaEventRecord(c->pev1[slot], c->stream);
  c->launches++;
}

template <int NV, int RECON, int SOLVER, int BF>
static void launch_lim(pb200_ctx *c, int dir, const SweepArgs &a) {
  // LIMITER DEFAULT is compiled in; any other choice takes the run-time limiter switch
  if (RECON != RECON_LINEAR || c->cfg.limiter == PB200_LIM_DEFAULT) launch_dir<NV, RECON, SOLVER, LIM_DEFAULT, BF>(c, dir, a);
  else launch_dir<NV, RECON, SOLVER, LIM_RT, BF>(c, dir, a);
}

template <int NV, int RECON, int BF>
static void launch_solver(pb200_ctx *c, int dir, const SweepArgs &a) {
  switch (c->cfg.solver) {
    case PB200_TVDLF: launch_lim<NV, RECON, SOLVER_TVDLF, BF>(c, dir, a); break;
    case PB200_HLL: launch_lim<NV, RECON, SOLVER_HLL, BF>(c, dir, a); break;
    default: launch_lim<NV, RECON, SOLVER_HLLC, BF>(c, dir, a); break;
  }
}

#define PB_CAT2(a, b, c, d) a##b##c##d
#define PB_CAT(a, b, c, d) PB_CAT2(a, b, c, d)
void PB_CAT(pb200_launch_sweep_nv, PB_NV, _bf, PB_BF)(pb200_ctx *c, int dir, const SweepArgs &a) {
  switch (c->cfg.reconstruction) {
    case PB200_FLAT: launch_solver<PB_NV, RECON_FLAT, PB_BF>(c, dir, a); break;
    case PB200_PARABOLIC: launch_solver<PB_NV, RECON_PARABOLIC, PB_BF>(c, dir, a); break;
    default: launch_solver<PB_NV, RECON_LINEAR, PB_BF>(c, dir, a); break;
  }
}
