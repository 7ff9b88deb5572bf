// pb200_sweeps.cu -- instantiation and launch of the sweep kernels for ONE (NVAR, BODY_FORCE)
// pair: compile with -DPB_NV=<5|6|7> -DPB_BF=<0|1>.  Solver / reconstruction / limiter are
// run-time options of the reference ([Solver] in pluto.ini) or cheap to carry as template
// parameters, so every combination is instantiated here.
#include <cstdlib>
#include <mutex>
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
  // high-water mark of the opt-in dynamic shared memory per (device, kernel): the attribute is
  // per device, and a process may drive several (pb200_multi)
  static std::unordered_map<const void *, size_t> cur[64];
  static std::mutex mtx;          // the worker threads of pb200_multi launch concurrently
  std::lock_guard<std::mutex> lock(mtx);
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &c = cur[dev & 63][(const void *)k];
  if (shm > c) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm); c = shm; }
}

template <int NV, int RECON, int SOLVER, int LIM, int BF>
static void launch_dir(pb200_ctx *c, int dir, const SweepArgs &a) {
  const Dev &D = c->dev;
  cudaStream_t stream = c->stream;
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
    // a launch covers the x3 planes [k0, k1) (all of them except in the slab-wise host pipeline)
    const int nk = (D.ndim == 3) ? a.k1 - a.k0 : 1;
    int npen = (dir == 2) ? nk : (D.end[dir] - D.beg[dir] + 1);
    int ntr = (dir == 1) ? nk : (D.end[1] - D.beg[1] + 1);
    const bool cdt_in = D.ndim > 1 && a.stage == 1 && !fusex;
    const int nq = ring_nq(NV, fusex, a.comb, cdt_in);
    // chunk the pencil so that the grid holds several waves of 148 SMs x resident blocks
    auto chunks = [&](int nbx, int resident, int &chunk) {
      long want = 148L * resident * 6;
      int nchunk = 1;
      while ((long)nbx * ntr * nchunk < want && npen / (nchunk * 2) >= 32) nchunk *= 2;
      chunk = (npen + nchunk - 1) / nchunk;
      return (npen + chunk - 1) / chunk;
    };
    if (fusex) {
      // The fused kernel loses 2*XH threads per block to the x1 halo.  Cover the row with blocks of
      // 128, 160 and 192 threads so that the FEWEST WARPS run (512 zones, PLM: 2x192 + 1x160 = 17
      // warps instead of 5x128 = 20); one launch per block width, at its x1 offset.
      const int h2 = 2 * recon_xhalo<RECON>();
#if PB_REG128
      constexpr int W0 = 256, W1 = 128, W2 = 64;
#else
      constexpr int W0 = 192, W1 = 160, W2 = 128;
#endif
      const int W[3] = {W0, W1, W2};
      int best[3] = {0, 0, (nx + W2 - h2 - 1) / (W2 - h2)}, bestw = best[2] * (W2 / 32);
      for (int n0 = 0; n0 * (W0 - h2) < nx + (W0 - h2); n0++)
        for (int n1 = 0; n0 * (W0 - h2) + n1 * (W1 - h2) < nx + (W1 - h2); n1++) {
          int rest = nx - n0 * (W0 - h2) - n1 * (W1 - h2);
          int n2 = rest > 0 ? (rest + W2 - h2 - 1) / (W2 - h2) : 0;
          int w = n0 * (W0 / 32) + n1 * (W1 / 32) + n2 * (W2 / 32);
          if (w < bestw || (w == bestw && n0 + n1 + n2 < best[0] + best[1] + best[2])) {
            bestw = w; best[0] = n0; best[1] = n1; best[2] = n2;
          }
        }
      // experiment hook: PB200_SMEM_PAD=<bytes> inflates the dynamic shared memory (lowers occupancy)
      static const size_t smem_pad = getenv("PB200_SMEM_PAD") ? (size_t)atol(getenv("PB200_SMEM_PAD")) : 0;
      SweepArgs b = a;
      b.i0 = 0;
      for (int g = 0; g < 3; g++) {
        if (!best[g]) continue;
        int chunk;
        const int nchunk = chunks(best[g], fused_minblk(true, W[g]), chunk);
        dim3 grid(best[g], ntr, nchunk);
#define PB_LAUNCH_FUSED(LASTF, WIDTH)                                                        \
  {                                                                                          \
    auto k = sweep_fused<1, true, LASTF, NV, RECON, SOLVER, LIM, BF, WIDTH>;                 \
    size_t shm = sweep_smem_bytes<true, NV, RECON, WIDTH>(nq) + smem_pad;                    \
    set_smem(k, shm);                                                                        \
    k<<<grid, WIDTH, shm, c->stream>>>(D, b, chunk);                                         \
  }
        if (last) {
          if (g == 0) PB_LAUNCH_FUSED(true, W0) else if (g == 1) PB_LAUNCH_FUSED(true, W1) else PB_LAUNCH_FUSED(true, W2)
        } else {
          if (g == 0) PB_LAUNCH_FUSED(false, W0) else if (g == 1) PB_LAUNCH_FUSED(false, W1) else PB_LAUNCH_FUSED(false, W2)
        }
#undef PB_LAUNCH_FUSED
        b.i0 += best[g] * (W[g] - h2);
        c->launches++;
      }
      c->launches--;   // the common exit counts one
    } else {
      int chunk;
      const int nchunk = chunks((nx + BX - 1) / BX, 3, chunk);
      dim3 grid((nx + BX - 1) / BX, ntr, nchunk);
      auto k = sweep_fused<2, false, true, NV, RECON, SOLVER, LIM, BF>;
      size_t shm = sweep_smem_bytes<false, NV, RECON>(nq);
      set_smem(k, shm);
      k<<<grid, BX, shm, c->stream>>>(D, a, chunk);
    }
  }
  if (slot >= 0) cudaEventRecord(c->pev1[slot], c->stream);
  c->launches++;
}

// PB_HOT_ONLY (kernel experiments, never the shipped library): only NVAR 5 without body force,
// LINEAR + HLLC + LIMITER DEFAULT is instantiated, so that a variant library builds in seconds
#ifdef PB_HOT_ONLY
#include <cstdio>
#include <cstdlib>
#define PB_HOT_REFUSE() do { fprintf(stderr, "PB_HOT_ONLY build: configuration not instantiated\n"); abort(); } while (0)
#endif

template <int NV, int RECON, int SOLVER, int BF>
static void launch_lim(pb200_ctx *c, int dir, const SweepArgs &a) {
#ifdef PB_HOT_ONLY
  if (c->cfg.limiter != PB200_LIM_DEFAULT) PB_HOT_REFUSE();
  launch_dir<NV, RECON, SOLVER, LIM_DEFAULT, BF>(c, dir, a);
#else
  // LIMITER DEFAULT is compiled in; any other choice takes the run-time limiter switch
  if (RECON != RECON_LINEAR || c->cfg.limiter == PB200_LIM_DEFAULT) launch_dir<NV, RECON, SOLVER, LIM_DEFAULT, BF>(c, dir, a);
  else launch_dir<NV, RECON, SOLVER, LIM_RT, BF>(c, dir, a);
#endif
}

template <int NV, int RECON, int BF>
static void launch_solver(pb200_ctx *c, int dir, const SweepArgs &a) {
#ifdef PB_HOT_ONLY
  if (c->cfg.solver != PB200_HLLC) PB_HOT_REFUSE();
  launch_lim<NV, RECON, SOLVER_HLLC, BF>(c, dir, a);
#else
  switch (c->cfg.solver) {
    case PB200_TVDLF: launch_lim<NV, RECON, SOLVER_TVDLF, BF>(c, dir, a); break;
    case PB200_HLL: launch_lim<NV, RECON, SOLVER_HLL, BF>(c, dir, a); break;
    default: launch_lim<NV, RECON, SOLVER_HLLC, BF>(c, dir, a); break;
  }
#endif
}

#define PB_CAT2(a, b, c, d) a##b##c##d
#define PB_CAT(a, b, c, d) PB_CAT2(a, b, c, d)
void PB_CAT(pb200_launch_sweep_nv, PB_NV, _bf, PB_BF)(pb200_ctx *c, int dir, const SweepArgs &a) {
#ifdef PB_HOT_ONLY
#if PB_NV == 5 && PB_BF == 0
  if (c->cfg.reconstruction != PB200_LINEAR && c->cfg.reconstruction != PB200_PARABOLIC) PB_HOT_REFUSE();
  if (c->cfg.reconstruction == PB200_PARABOLIC) launch_solver<PB_NV, RECON_PARABOLIC, PB_BF>(c, dir, a);
  else launch_solver<PB_NV, RECON_LINEAR, PB_BF>(c, dir, a);
#else
  PB_HOT_REFUSE();
#endif
#else
  switch (c->cfg.reconstruction) {
    case PB200_FLAT: launch_solver<PB_NV, RECON_FLAT, PB_BF>(c, dir, a); break;
    case PB200_PARABOLIC: launch_solver<PB_NV, RECON_PARABOLIC, PB_BF>(c, dir, a); break;
    default: launch_solver<PB_NV, RECON_LINEAR, PB_BF>(c, dir, a); break;
  }
#endif
}
